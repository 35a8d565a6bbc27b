// Micro-benchmarks used while tuning (not on the product path): raw tcgen05.mma issue rate for
// the shared-memory operand layouts considered for the conv kernel.
#include "common.cuh"

namespace dbm {

// mode 0: K-major no-swizzle, halo-tile strides of the conv kernel (A: LBO 5184, SBO 288)
// mode 1: K-major no-swizzle, dense (LBO 2048, SBO 128)
// mode 2: K-major SWIZZLE_128B (SBO 1024), the layout used by stock GEMMs
template <int N>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int mode, int iters, int per_commit, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<128>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < per_commit; ++j) {
        uint64_t ad, bd;
        const int tap = j % 9, ks = (j / 9) & 1;
        if (mode == 0) {
          ad = umma_desc_kmajor_noswz(a0 + (uint32_t)(((2 * ks) * 18 + tap / 3) * 18 + tap % 3) * 16, 5184, 288);
          bd = umma_desc_kmajor_noswz(b0 + (uint32_t)(((tap % 3) * 4 + 2 * ks) * (N / 8)) * 128, (N / 8) * 128, 128);
        } else if (mode == 1) {
          ad = umma_desc_kmajor_noswz(a0 + (uint32_t)(2 * ks) * 2048, 2048, 128);
          bd = umma_desc_kmajor_noswz(b0 + (uint32_t)(((tap % 3) * 4 + 2 * ks) * (N / 8)) * 128, (N / 8) * 128, 128);
        } else {
          ad = umma_desc_kmajor_noswz(a0 + (uint32_t)ks * 32 + (uint32_t)(tap & 3) * 16384, 16, 1024) | ((uint64_t)2 << 61);
          bd = umma_desc_kmajor_noswz(b0 + (uint32_t)ks * 32 + (uint32_t)(tap & 1) * (N * 128), 16, 1024) | ((uint64_t)2 << 61);
        }
        umma_bf16(tmem, ad, bd, idesc, (it | j) ? 1u : 0u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<128>(tmem);
  }
}

}  // namespace dbm

using namespace dbm;

// out_cycles: one int64 per CTA (grid = number of SMs). Returns total MMAs per CTA via *mmas.
extern "C" int dbm_debug_umma_rate(int mode, int n, int iters, int per_commit, long long* out_cycles, int grid,
                                   cudaStream_t stream) {
  DBM_REQUIRE(n == 32 || n == 64 || n == 128, "umma_rate: N must be 32, 64 or 128");
  const int smem = 96 * 1024 + 2048;
  if (n == 32) {
    DBM_CUDA(cudaFuncSetAttribute(umma_rate_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_rate_kernel<32><<<grid, 128, smem, stream>>>(mode, iters, per_commit, out_cycles);
  } else if (n == 64) {
    DBM_CUDA(cudaFuncSetAttribute(umma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_rate_kernel<64><<<grid, 128, smem, stream>>>(mode, iters, per_commit, out_cycles);
  } else {
    DBM_CUDA(cudaFuncSetAttribute(umma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_rate_kernel<128><<<grid, 128, smem, stream>>>(mode, iters, per_commit, out_cycles);
  }
  return check_launch("umma_rate_kernel");
}
