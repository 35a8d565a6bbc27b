// Pieces shared by the tcgen05 3x3-conv kernels (per-layer kernel and the persistent trunk kernel).
#pragma once
#include <utility>

#include "common.cuh"

namespace dbm {

constexpr int kTile = 16;         // a work item is a kTile x kTile pixel output unit (two M=128 MMA tiles)
constexpr int kHalo = kTile + 2;  // 18: halo tile edge

// slab8 bf16 tensor [N][CS][H][W][8] viewed as 4-D (W*8, H, CS, N); box = 18 px x 18 rows x ck/8 slabs
// (default box: the 18 x 18 halo tile of a 16 x 16 unit)
int make_slab8_tmap(CUtensorMap* tm, const void* base, int N, int CS, int H, int W, int ck, int box_w = kHalo,
                    int box_h = kHalo);

// All MMAs of one pipeline stage: 9 taps x (2 sub-tiles x CK/16 k-steps). The tap loop is kept
// rolled: fully unrolling it makes ptxas hoist all 72 descriptor words, overflow the uniform
// register file and pay R2UR.FILL / MOV.SPILL around every UTCHMMA.
template <int COUT, int CK, int DSTRIDE, int IDX>
__device__ __forceinline__ void issue_one(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t acc) {
  constexpr int KS = CK / 16;
  constexpr int j = IDX / KS, ks = IDX % KS;
  // start-address field advances in 16-byte units: no carry into the LBO field
  constexpr uint32_t a_off = (uint32_t)((2 * ks) * kHalo * kHalo + 8 * j);
  constexpr uint32_t b_off = (uint32_t)((2 * ks) * (COUT / 8) * 8);
  umma_bf16_off<a_off, b_off>(d0 + (uint32_t)(j * DSTRIDE), a_lo, a_hi, b_lo, b_hi, idesc, ks != 0 ? 1u : acc);
}
template <int COUT, int CK, int DSTRIDE, int... IDX>
__device__ __forceinline__ void issue_tap(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t acc, std::integer_sequence<int, IDX...>) {
  (issue_one<COUT, CK, DSTRIDE, IDX>(d0, a_lo, a_hi, b_lo, b_hi, idesc, acc), ...);
}
// DSTRIDE = TMEM column distance between the two sub-tile accumulators
template <int COUT, int CK, int DSTRIDE = COUT>
__device__ __forceinline__ void issue_stage_mmas(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t idesc, uint32_t acc0) {
  constexpr uint32_t kBTap = (uint32_t)((CK / 8) * (COUT / 8) * 8);  // per-tap stride of the packed weights
#pragma unroll 1
  for (uint32_t tap = 0; tap < 9; ++tap) {
    const uint32_t a_tap = a_lo + tap + (tap / 3) * (kHalo - 3);     // ky * 18 + kx
    const uint32_t b_tap = b_lo + tap * kBTap;
    issue_tap<COUT, CK, DSTRIDE>(d0, a_tap, a_hi, b_tap, b_hi, idesc, tap != 0 ? 1u : acc0,
                        std::make_integer_sequence<int, 2 * (CK / 16)>{});
  }
}


// ---- device tables shared by the kernels and the model-level entry points (csrc/gen_api.cu) --------------------
// One MMA pass: out[:, 0:cout] = conv3x3(in[:, 8*in_cs0 : 8*in_cs0 + cin]) with the packed filter.
// mode 0: a layer of its own: all `cout` columns take the fused epilogue (bias, residuals, LeakyReLU, stores).
// Dense-block pairing (see model.py): conv_k and the partial sums of conv_{k+1} over their shared inputs are ONE
// N = 64 pass (mode 1, "head"): columns [0, 32) = conv_k take the epilogue, columns [32, 64) -- a partial
// pre-activation of conv_{k+1} -- STAY IN TENSOR MEMORY; the next table entry (mode 2, "tail") contracts only a_k
// (K = 32 * 9, N = 32) and accumulates onto those very columns, then takes the epilogue of conv_{k+1}. Head and tail
// of a unit run on the same CTA (see the schedule below), so the partial sums never leave the SM: no fp32 stash in
// HBM (round 1 wrote and re-read 128 B per pixel per pair), and the tail's epilogue has nothing to load.
struct TrunkLayer {  // 128 bytes, mirrored by deepbedmap_b200/model.py (TRUNK_LAYER_DTYPE)
  const __nv_bfloat16* wpacked;
  const float* bias;
  __nv_bfloat16* out_bf16;
  float* out_f32;      // slab8f (fp32 [N][8][H][W][8]), 64 channels
  const float* res1;   // slab8f, 4 * res1_cs_total channels
  const float* res2;   // slab8f, 64 channels
  float* stash_out;    // slab8f, cout - cout_main channels
  int cin, cout;       // cout (MMA N) in {32, 64}
  int in_map, in_cs0;  // in_map: 0 = stem output (16 slabs), 1 / 2 = dense-block buffers
  int act, up2;
  int out_cs_total, out_cs0;
  int cout_main, res1_cs_total;
  float beta;
  int mode;            // 0 single pass, 1 pair head, 2 pair tail (must directly follow its head)
  int pad[6];
};
static_assert(sizeof(TrunkLayer) == 128, "TrunkLayer layout is part of the C ABI");


struct PackEntry {  // 48 bytes, mirrored by deepbedmap_b200/model.py (PACK_ENTRY_DTYPE)
  const float* w;
  __nv_bfloat16* out;
  int O, o0, Cin, CinTotal, c0, COUTP, CK, mode;  // mode: see pack_w3x3_table_kernel
};
static_assert(sizeof(PackEntry) == 48, "PackEntry layout is part of the C ABI");


}  // namespace dbm
