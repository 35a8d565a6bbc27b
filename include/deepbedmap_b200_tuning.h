/*
 * deepbedmap_b200 -- tuning / A-B switches. NOT part of the drop-in boundary (include/deepbedmap_b200.h): nothing a
 * user of the library calls. They exist for the scripts under scripts/ (ablation runs behind profiles/README.md) and
 * for the A/B parity tests that run the same MMAs through two schedules. Results under an ablation mask are invalid.
 */
#ifndef DEEPBEDMAP_B200_TUNING_H_
#define DEEPBEDMAP_B200_TUNING_H_

#ifdef __cplusplus
extern "C" {
#endif

/* key 1: swap LBO/SBO of the per-layer conv kernel's descriptors (bring-up); key 2: per-layer conv kernel in the trunk
 * kernel's 16-channel chunks (bit-exact A/B of the persistent trunk, tests/test_gpu_model.py); key 3: ablation mask of
 * umma_trunk_kernel (1 no dependency wait, 2 no epilogue memory traffic, 4 no TMA loads, 8/16 fence placement,
 * 32/64/128 single traffic classes): timing only. */
int dbm_debug_set(int key, int value);
/* key 1: device buffer [passes][8] of cycle counters written by umma_trunk_kernel (NULL = off) */
int dbm_debug_set_ptr(int key, void* device_ptr);
/* key 1: swap LBO/SBO of the weight-gradient kernel's MN-major descriptors (bring-up) */
int dbm_flat_debug_set(int key, int value);
/* force the image grouping of local_trunk_kernel (0 automatic, 1 one image per CTA, 2 lock-step pair, 3 SOLO) */
int dbm_local_debug_set(int value);
/* tcgen05.mma issue-rate microbenchmark (libdeepbedmap_b200_tuning.so, scripts/umma_rate.py) */
int dbm_debug_umma_rate(int mode, int n, int iters, int per_commit, long long* out_cycles, int grid, void* stream);

#ifdef __cplusplus
}
#endif
#endif
