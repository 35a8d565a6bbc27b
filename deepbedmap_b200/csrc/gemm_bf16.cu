// Batched plain GEMM with bf16 operands rounded on the fly from fp32 memory and fp32 accumulation on the
// tensor cores (warp-level mma through nvcuda::wmma): the three contractions of the TRAINING path's first
// deformable layer (L.DeformableConvolution2D 64 -> 64, srgan_train.py:506-514, and its autograd):
//     y     [px, o ] = sum_k  cols[k, px] * W[o, k]   + b          (M = H W, N = 64,  K = 576)
//     dW    [o,  kk] += sum_p dy[o, p]    * cols[kk, p]            (M = 64,  N = 576, K = H W, reduced over the batch)
//     dcols [px, kk] = sum_o  dy[o, px]   * W[o, kk]               (M = H W, N = 576, K = 64)
// Same arithmetic contract as every other conv of the bf16 training path (DESIGN.md "Numerics"): operands
// rounded to bf16, fp32 accumulation. These GEMMs stream the fp32 cols buffer (382 MB per pass at batch 128):
// the fp32 CUDA-core kernel (gemm_f32.cu) spent 430 / 510 / 430 us on them, this one 250 / 400 / 420 us -- still
// latency bound (one 16-deep K step per barrier, two CTAs per SM), 3-4x the HBM time: open item.
// Operands keep their natural strides: a tile is staged in shared memory in the orientation it has in global
// memory (coalesced loads along the unit-stride dimension) and wmma's row/col-major fragment loads do the rest.
#include <mma.h>

#include <type_traits>

#include "common.cuh"

namespace dbm {

using namespace nvcuda;

struct GemmBf16P {
  const float* A;
  const float* B;
  float* C;
  const float* bias;   // per n, or NULL
  int M, N, K, batch;
  long lda_m, lda_k, ldb_k, ldb_n, ldc_m, ldc_n;
  long a_bs, b_bs, c_bs;
  int act;             // LeakyReLU(0.2) after the bias
  int atomic;          // 1: C += sum over the batch (atomicAdd), K x batch split over gridDim.z
  int kchunks;         // atomic: K chunks per image
};

constexpr int kGBK = 16;   // (32 was measured slower: 175 registers, one CTA per SM)

// A_COL: A(m, k) is unit-stride along m; else along k.  B_ROW: B(k, n) is unit-stride along n; else along k.
template <int BM, int BN, bool A_COL, bool B_ROW>
__global__ void __launch_bounds__(256) gemm_bf16_kernel(const GemmBf16P p) {
  constexpr int WM = BM / 32, WN = BN / 32;          // warps along m / n, 32 x 32 outputs each
  static_assert(WM * WN == 8, "eight warps");
  constexpr int LDA = A_COL ? BM + 8 : kGBK + 8;     // bf16 elements; +8 keeps 16-byte row alignment, skews banks
  constexpr int LDB = B_ROW ? BN + 8 : kGBK + 8;
  constexpr int LDC = BM + 4;                         // epilogue staging, column-major (m fastest)
  constexpr int kAElems = A_COL ? kGBK * LDA : BM * LDA, kBElems = B_ROW ? kGBK * LDB : BN * LDB;
  constexpr int kABBytes = 2 * (kAElems + kBElems) * 2, kCBytes = BN * LDC * 4;
  // the epilogue staging tile re-uses the operand buffers (static shared memory stays under 48 KB)
  __shared__ __align__(128) unsigned char raw[kABBytes > kCBytes ? kABBytes : kCBytes];
  __nv_bfloat16 (*As)[kAElems] = reinterpret_cast<__nv_bfloat16 (*)[kAElems]>(raw);
  __nv_bfloat16 (*Bs)[kBElems] = reinterpret_cast<__nv_bfloat16 (*)[kBElems]>(raw + 2 * kAElems * 2);
  float* Cs = reinterpret_cast<float*>(raw);

  const int t = threadIdx.x, warp = t >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // work range: plain -> one image (blockIdx.z), all of K; atomic -> a slice of (image, K chunk) pairs
  int img0, img1, kc0 = 0, kc1 = 1;
  const int klen = p.atomic ? (p.K + p.kchunks - 1) / p.kchunks : p.K;
  if (p.atomic) {
    const long total = (long)p.batch * p.kchunks;
    const long w0 = total * blockIdx.z / gridDim.z, w1 = total * (blockIdx.z + 1) / gridDim.z;
    img0 = (int)(w0 / p.kchunks); kc0 = (int)(w0 % p.kchunks);
    img1 = (int)((w1 - 1) / p.kchunks); kc1 = (int)((w1 - 1) % p.kchunks) + 1;
    if (w1 <= w0) return;
  } else {
    img0 = img1 = blockIdx.z;
  }

  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) wmma::fill_fragment(acc[i][j], 0.f);

  constexpr int NA = BM * kGBK / 256, NB = BN * kGBK / 256;   // elements per thread per tile
  float ra[NA], rb[NB];
  const float* Ab = nullptr;
  const float* Bb = nullptr;
  auto fetch = [&](int k0, int kend) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = t + 256 * i;
      const int ml = A_COL ? e % BM : e / kGBK, kl = A_COL ? e / BM : e % kGBK;
      const int m = m0 + ml, k = k0 + kl;
      ra[i] = (m < p.M && k < kend) ? __ldg(Ab + m * p.lda_m + k * p.lda_k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = t + 256 * i;
      const int nl = B_ROW ? e % BN : e / kGBK, kl = B_ROW ? e / BN : e % kGBK;
      const int n = n0 + nl, k = k0 + kl;
      rb[i] = (n < p.N && k < kend) ? __ldg(Bb + k * p.ldb_k + n * p.ldb_n) : 0.f;
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = t + 256 * i;
      const int ml = A_COL ? e % BM : e / kGBK, kl = A_COL ? e / BM : e % kGBK;
      As[buf][A_COL ? kl * LDA + ml : ml * LDA + kl] = __float2bfloat16_rn(ra[i]);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = t + 256 * i;
      const int nl = B_ROW ? e % BN : e / kGBK, kl = B_ROW ? e / BN : e % kGBK;
      Bs[buf][B_ROW ? kl * LDB + nl : nl * LDB + kl] = __float2bfloat16_rn(rb[i]);
    }
  };

  int buf = 0;
  for (int img = img0; img <= img1; ++img) {
    Ab = p.A + (long)img * p.a_bs;
    Bb = p.B + (long)img * p.b_bs;
    const int ca = (p.atomic && img == img0) ? kc0 : 0;
    const int cb = (p.atomic && img == img1) ? kc1 : (p.atomic ? p.kchunks : 1);
    const int kbeg = ca * klen, kend = min(p.K, cb * klen);
    if (kbeg >= kend) continue;
    fetch(kbeg, kend);
    for (int k0 = kbeg; k0 < kend; k0 += kGBK) {
      commit(buf);
      __syncthreads();
      if (k0 + kGBK < kend) fetch(k0 + kGBK, kend);
#pragma unroll
      for (int ks = 0; ks < kGBK; ks += 16) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16,
                       typename std::conditional<A_COL, wmma::col_major, wmma::row_major>::type> fa[2];
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16,
                       typename std::conditional<B_ROW, wmma::row_major, wmma::col_major>::type> fb[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ml = wm * 32 + 16 * i;
          wmma::load_matrix_sync(fa[i], A_COL ? &As[buf][ks * LDA + ml] : &As[buf][ml * LDA + ks], LDA);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int nl = wn * 32 + 16 * j;
          wmma::load_matrix_sync(fb[j], B_ROW ? &Bs[buf][ks * LDB + nl] : &Bs[buf][nl * LDB + ks], LDB);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) wmma::mma_sync(acc[i][j], fa[i], fb[j], acc[i][j]);
      }
      buf ^= 1;   // the next commit writes the other buffer: one barrier per K step is enough
    }
  }

  // ---------------- epilogue: stage column-major (m fastest), then bias / act / store along C's unit stride ----------------
  __syncthreads();   // every warp is done with the operand buffers the staging tile overlays
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      wmma::store_matrix_sync(&Cs[(wn * 32 + 16 * j) * LDC + wm * 32 + 16 * i], acc[i][j], LDC, wmma::mem_col_major);
  __syncthreads();
  float* Cb = p.C + (p.atomic ? 0 : (long)blockIdx.z * p.c_bs);
  const bool m_fast = p.ldc_m == 1;
  for (int e = t; e < BM * BN; e += 256) {
    const int ml = m_fast ? e % BM : e / BN, nl = m_fast ? e / BM : e % BN;
    const int m = m0 + ml, n = n0 + nl;
    if (m >= p.M || n >= p.N) continue;
    float v = Cs[nl * LDC + ml];
    if (p.atomic) {
      atomicAdd(Cb + m * p.ldc_m + n * p.ldc_n, v);
    } else {
      if (p.bias) v += __ldg(p.bias + n);
      if (p.act) v = lrelu(v);
      Cb[m * p.ldc_m + n * p.ldc_n] = v;
    }
  }
}

template <int BM, int BN>
static int launch_gemm_bf16(const GemmBf16P& p, bool a_col, bool b_row, dim3 grid, cudaStream_t st) {
  if (a_col && b_row) gemm_bf16_kernel<BM, BN, true, true><<<grid, 256, 0, st>>>(p);
  else if (a_col) gemm_bf16_kernel<BM, BN, true, false><<<grid, 256, 0, st>>>(p);
  else if (b_row) gemm_bf16_kernel<BM, BN, false, true><<<grid, 256, 0, st>>>(p);
  else gemm_bf16_kernel<BM, BN, false, false><<<grid, 256, 0, st>>>(p);
  return check_launch("gemm_bf16_kernel");
}

}  // namespace dbm

using namespace dbm;

// Same contract as dbm_gemm_f32 (strides in elements, batch strides, bias per n, accumulate: 0 overwrite,
// 2 atomicAdd into a batch-reduced C), operands rounded to bf16.
extern "C" int dbm_gemm_bf16(const float* a, long lda_m, long lda_k, long a_batch_stride, const float* b, long ldb_k,
                             long ldb_n, long b_batch_stride, float* c, long ldc_m, long ldc_n, long c_batch_stride,
                             const float* bias, int m, int n, int k, int batch, int act, int accumulate,
                             cudaStream_t st) {
  DBM_REQUIRE(m > 0 && n > 0 && k > 0 && batch > 0, "gemm_bf16: empty problem %dx%dx%d x%d", m, n, k, batch);
  DBM_REQUIRE(accumulate == 0 || accumulate == 2, "gemm_bf16: accumulate must be 0 (overwrite) or 2 (atomic batch sum)");
  DBM_REQUIRE((lda_m == 1 || lda_k == 1) && (ldb_k == 1 || ldb_n == 1) && (ldc_m == 1 || ldc_n == 1),
              "gemm_bf16: every operand needs a unit stride");
  DBM_REQUIRE(batch <= 65535, "gemm_bf16: batch %d too large", batch);
  GemmBf16P p{};
  p.A = a; p.B = b; p.C = c; p.bias = bias;
  p.M = m; p.N = n; p.K = k; p.batch = batch;
  p.lda_m = lda_m; p.lda_k = lda_k; p.ldb_k = ldb_k; p.ldb_n = ldb_n; p.ldc_m = ldc_m; p.ldc_n = ldc_n;
  p.a_bs = a_batch_stride; p.b_bs = b_batch_stride; p.c_bs = c_batch_stride;
  p.act = act; p.atomic = accumulate == 2;
  const bool a_col = lda_m == 1, b_row = ldb_n == 1;
  const bool wide_n = m <= 64;   // M = 64 (weight gradient): 64 x 128 tiles; else 128 x 64
  const int bm = wide_n ? 64 : 128, bn = wide_n ? 128 : 64;
  dim3 grid(ceil_div(m, bm), ceil_div(n, bn), batch);
  if (p.atomic) {
    // split (image, K chunk) pairs over ~4 CTAs per SM in total
    p.kchunks = 1;
    long tiles = (long)grid.x * grid.y;
    long z = (4L * num_sms() + tiles - 1) / tiles;
    if (z < 1) z = 1;
    while ((long)batch * p.kchunks < z && p.kchunks < 64 && k / (p.kchunks * 2) >= 4 * kGBK) p.kchunks *= 2;
    if (z > (long)batch * p.kchunks) z = (long)batch * p.kchunks;
    // K chunks must be whole K tiles so that tiles never straddle a chunk boundary
    const int klen = (k + p.kchunks - 1) / p.kchunks;
    if (klen % kGBK != 0 && p.kchunks > 1) p.kchunks = 1;
    if (z > (long)batch * p.kchunks) z = (long)batch * p.kchunks;
    grid.z = (unsigned)z;
  }
  return wide_n ? launch_gemm_bf16<64, 128>(p, a_col, b_row, grid, st) : launch_gemm_bf16<128, 64>(p, a_col, b_row, grid, st);
}
