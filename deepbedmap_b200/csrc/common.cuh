// Shared helpers for the deepbedmap_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deepbedmap_b200.h"

namespace dbm {

// ---- error plumbing: every C-ABI entry returns 0 on success, message via dbm_last_error ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define DBM_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::dbm::set_error(__VA_ARGS__);      \
      return DBM_ERR_INVALID;             \
    }                                     \
  } while (0)

#define DBM_CUDA(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ::dbm::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
      return DBM_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

constexpr float kLreluSlope = 0.2f;  // srgan_train.py:340

__device__ __forceinline__ float lrelu(float v) { return v >= 0.f ? v : v * kLreluSlope; }

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

int num_sms();
// dbm_set_deterministic(1): every reduction that would otherwise combine partial sums with floating-point atomics (in an
// order the hardware picks) runs in a fixed order instead -- single-contributor launches, ordered two-level sums, or,
// for the deformable scatter, exact 64-bit fixed-point accumulation. The reference trains with
// chainer.global_config.cudnn_deterministic = True (srgan_train.py:69).
bool deterministic();
// library-owned, grow-only device scratch of the deterministic mode (int64 shadow of a scatter target)
int det_scratch(size_t bytes, void** out);
// Co-residency bound of a persistent kernel whose CTAs wait for each other through global flags (umma_trunk_kernel,
// flat_chain_kernel): opts the kernel in to `smem` bytes of dynamic shared memory on the CURRENT device and returns in
// *resident the number of CTAs that device holds at once for (threads, smem) -- occupancy query x SM count, cached per
// (device, kernel, smem). A launch must not use a larger grid: a CTA that is not resident can never publish the flags
// the resident ones spin on. Fails if not even one CTA fits.
int resident_ctas(const void* kernel, int threads, size_t smem, int* resident);
// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (device, kernel): the attribute is per device, a
// per-process flag would leave the second GPU of a multi-device process without it
int ensure_dyn_smem(const void* kernel, size_t smem);

// ---------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 (Blackwell). Raw PTX so that nothing but the CUDA
// toolkit is needed to build.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wall-clock bound for the spins on inter-CTA flags: a co-tenant kernel (NCCL, a side stream) may legitimately hold an
// SM a resident CTA is waiting for, so the bound is 20 s of %globaltimer, not a spin count; a protocol bug still
// traps instead of hanging the GPU box.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct SpinGuard {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  __device__ __forceinline__ bool expired() {
    if ((++spins & 1023u) != 0) return false;
    const unsigned long long now = globaltimer_ns();
    if (t0 == 0) { t0 = now; return false; }
    return now - t0 > 20000000000ull;
  }
};

// Bounded spin: a broken pipeline traps instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("dbm: mbarrier wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- tcgen05 / TMEM ----
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 in, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the descriptors formed inside the asm block from per-stage base words plus
// compile-time start-address offsets (16-byte units): keeps the live uniform-register set small
// (2 adds + 2 moves + UTCHMMA per instruction instead of hoisted/spilled descriptor tables).
template <uint32_t kAOff, uint32_t kBOff>
__device__ __forceinline__ void umma_bf16_off(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 alo, blo;\n\t.reg .b64 ad, bd;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "add.u32 alo, %1, %7;\n\t"
      "add.u32 blo, %3, %8;\n\t"
      "mov.b64 ad, {alo, %2};\n\t"
      "mov.b64 bd, {blo, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "n"(kAOff), "n"(kBOff)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Exactly one lane of a converged warp gets true (keeps the surrounding code warp-uniform so
// descriptors stay in uniform registers; a plain `lane == 0` branch makes ptxas emit
// R2UR + ELECT loops in front of every tcgen05.mma).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xFFFFFFFF;\n\t"
      "@P1 mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
  return ((uint64_t)hi << 32) | (uint64_t)lo;
}
// lo/hi words of the K-major no-swizzle descriptor for a 16-byte-aligned smem address
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);  // version 1 at bit 46
}

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave") canonical layout
//   ((8, n), 2) : ((16 B, SBO), LBO)   -- 8 rows x 16 B core matrices (cute mma_traits_sm100)
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswz(uint32_t saddr, uint32_t lbo_bytes,
                                                           uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace dbm
