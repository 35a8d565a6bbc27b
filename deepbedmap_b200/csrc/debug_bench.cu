// Micro-benchmarks used while tuning (not on the product path): raw tcgen05.mma throughput for
// the shared-memory operand layouts considered for the conv kernel, issued exactly like the
// product kernel does (converged warp, elected lane, descriptors formed from immediates).
#include <utility>

#include "common.cuh"

namespace dbm {

// MODE 0: K-major no-swizzle, halo-tile strides of the conv kernel (A: LBO 5184, SBO 288)
// MODE 1: K-major no-swizzle, dense (A: LBO 2048, SBO 128)
// MODE 2: K-major SWIZZLE_128B (SBO 1024), the layout stock GEMMs use
template <int N, int MODE, int IDX>
__device__ __forceinline__ void rate_one(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc) {
  constexpr int tap = IDX % 9, ks = (IDX / 9) % 2, j = IDX / 18;
  constexpr uint32_t a_off = MODE == 0   ? (uint32_t)(((2 * ks) * 18 + tap / 3) * 18 + tap % 3 + 8 * j)
                             : MODE == 1 ? (uint32_t)((2 * ks) * 128 + j * 1024)
                                         : (uint32_t)(ks * 2 + (tap % 3) * 1024);
  constexpr uint32_t b_off = MODE == 2 ? (uint32_t)(ks * 2 + (tap % 2) * (N * 8))
                                       : (uint32_t)(((tap % 3) * 4 + 2 * ks) * (N / 8) * 8);
  umma_bf16_off<a_off, b_off>(d + (uint32_t)(j * N), a_lo, a_hi, b_lo, b_hi, idesc, 1u);
}
template <int N, int MODE, int... IDX>
__device__ __forceinline__ void rate_all(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, std::integer_sequence<int, IDX...>) {
  (rate_one<N, MODE, IDX>(d, a_lo, a_hi, b_lo, b_hi, idesc), ...);
}

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    uint32_t a_lo, a_hi, b_lo, b_hi;
    if (MODE == 0) {
      a_lo = desc_lo(a0, 5184); a_hi = desc_hi(288);
      b_lo = desc_lo(b0, (N / 8) * 128); b_hi = desc_hi(128);
    } else if (MODE == 1) {
      a_lo = desc_lo(a0, 2048); a_hi = desc_hi(128);
      b_lo = desc_lo(b0, (N / 8) * 128); b_hi = desc_hi(128);
    } else {
      a_lo = desc_lo(a0, 16); a_hi = desc_hi(1024) | (2u << 29);
      b_lo = desc_lo(b0, 16); b_hi = desc_hi(1024) | (2u << 29);
    }
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        rate_all<N, MODE>(tmem, a_lo, a_hi, b_lo, b_hi, idesc, std::make_integer_sequence<int, 36>{});
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int N, int MODE>
static int launch_rate(int iters, long long* out, int grid, cudaStream_t st) {
  const int smem = 160 * 1024 + 2048;
  DBM_CUDA(cudaFuncSetAttribute(umma_rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<N, MODE><<<grid, 128, smem, st>>>(iters, out);
  return check_launch("umma_rate_kernel");
}

}  // namespace dbm

using namespace dbm;

// out_cycles: one int64 per CTA; every CTA issues iters * 36 MMAs (M=128, K=16).
extern "C" int dbm_debug_umma_rate(int mode, int n, int iters, int per_commit, long long* out_cycles, int grid,
                                   cudaStream_t stream) {
  (void)per_commit;
#define DBM_RATE(NN)                                                         \
  if (n == NN) {                                                             \
    if (mode == 0) return launch_rate<NN, 0>(iters, out_cycles, grid, stream); \
    if (mode == 1) return launch_rate<NN, 1>(iters, out_cycles, grid, stream); \
    if (mode == 2) return launch_rate<NN, 2>(iters, out_cycles, grid, stream); \
  }
  DBM_RATE(32) DBM_RATE(64) DBM_RATE(128) DBM_RATE(256)
#undef DBM_RATE
  set_error("umma_rate: unsupported n=%d mode=%d", n, mode);
  return DBM_ERR_INVALID;
}
