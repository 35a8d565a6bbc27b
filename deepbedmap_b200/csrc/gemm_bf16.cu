// Batched plain GEMM with bf16 operands rounded on the fly from fp32 memory and fp32 accumulation on the
// tensor cores (warp-level mma through nvcuda::wmma): the three contractions of the TRAINING path's first
// deformable layer (L.DeformableConvolution2D 64 -> 64, srgan_train.py:506-514, and its autograd):
//     y     [px, o ] = sum_k  cols[k, px] * W[o, k]   + b          (M = H W, N = 64,  K = 576)
//     dW    [o,  kk] += sum_p dy[o, p]    * cols[kk, p]            (M = 64,  N = 576, K = H W, reduced over the batch)
//     dcols [px, kk] = sum_o  dy[o, px]   * W[o, kk]               (M = H W, N = 576, K = 64)
// Same arithmetic contract as every other conv of the bf16 training path (DESIGN.md "Numerics"): operands
// rounded to bf16, fp32 accumulation. These GEMMs stream the fp32 cols buffer (382 MB per pass at batch 128):
// the fp32 CUDA-core kernel (gemm_f32.cu) spent 430 / 510 / 430 us on them; here 250 / 320 / 190 us (general
// kernel, wide-K variant, short-K kernel) -- still 2-3x the HBM time (one K tile per barrier, no async copies).
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

// K depth of a shared-memory tile: 16, except 32 when both operands are unit-stride along k (the weight gradient):
// a 16-deep tile is then a 64-byte run per row -- half-used sectors on the 382 MB cols stream (measured 400 -> 340 us);
// for the other layouts 32 costs registers (175, one CTA per SM) and was slower.
constexpr int kGBK = 16, kGBKWide = 32;

// A_COL: A(m, k) is unit-stride along m; else along k.  B_ROW: B(k, n) is unit-stride along n; else along k.
template <int BM, int BN, bool A_COL, bool B_ROW, int BK>
__global__ void __launch_bounds__(256) gemm_bf16_kernel(const GemmBf16P p) {
  constexpr int WM = BM / 32, WN = BN / 32;          // warps along m / n, 32 x 32 outputs each
  static_assert(WM * WN == 8, "eight warps");
  constexpr int LDA = A_COL ? BM + 8 : BK + 8;     // bf16 elements; +8 keeps 16-byte row alignment, skews banks
  constexpr int LDB = B_ROW ? BN + 8 : BK + 8;
  constexpr int LDC = BM + 4;                         // epilogue staging, column-major (m fastest)
  constexpr int kAElems = A_COL ? BK * LDA : BM * LDA, kBElems = B_ROW ? BK * LDB : BN * LDB;
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

  constexpr int NA = BM * BK / 256, NB = BN * BK / 256;   // elements per thread per tile
  float ra[NA], rb[NB];
  const float* Ab = nullptr;
  const float* Bb = nullptr;
  auto fetch = [&](int k0, int kend) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = t + 256 * i;
      const int ml = A_COL ? e % BM : e / BK, kl = A_COL ? e / BM : e % BK;
      const int m = m0 + ml, k = k0 + kl;
      ra[i] = (m < p.M && k < kend) ? __ldg(Ab + m * p.lda_m + k * p.lda_k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = t + 256 * i;
      const int nl = B_ROW ? e % BN : e / BK, kl = B_ROW ? e / BN : e % BK;
      const int n = n0 + nl, k = k0 + kl;
      rb[i] = (n < p.N && k < kend) ? __ldg(Bb + k * p.ldb_k + n * p.ldb_n) : 0.f;
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = t + 256 * i;
      const int ml = A_COL ? e % BM : e / BK, kl = A_COL ? e / BM : e % BK;
      As[buf][A_COL ? kl * LDA + ml : ml * LDA + kl] = __float2bfloat16_rn(ra[i]);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = t + 256 * i;
      const int nl = B_ROW ? e % BN : e / BK, kl = B_ROW ? e / BN : e % BK;
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
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
      commit(buf);
      __syncthreads();
      if (k0 + BK < kend) fetch(k0 + BK, kend);
#pragma unroll
      for (int ks = 0; ks < BK; ks += 16) {
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

// Short-K form (K <= 64: the cols gradient, K = 64 output channels, N = 576): a CTA keeps its 128 x K block of A
// (unit stride along m) in shared memory and walks over ALL of N in 64-column tiles, B (unit stride along n) fetched
// per tile from L2 -- A is read once, and the kernel is what it should be, a stream of C writes (382 MB at batch 128).
// The general kernel above re-read A per N tile and paid a barrier + a global-load latency per 16-deep K step.
constexpr int kSKMax = 64;
__global__ void __launch_bounds__(256) gemm_bf16_shortk_kernel(const GemmBf16P p) {
  constexpr int BM = 128, BN = 64, LDA = BM + 8, LDB = BN + 8, LDC = BM + 4;
  extern __shared__ __align__(128) unsigned char sk_raw[];       // 60 KB: dynamic (over the 48 KB static limit)
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(sk_raw);                       // col-major: (m, k) at k * LDA + m
  __nv_bfloat16* Bs = As + kSKMax * LDA;                                               // row-major: (k, n) at k * LDB + n
  float* Cs = reinterpret_cast<float*>(sk_raw + (kSKMax * LDA + kSKMax * LDB) * 2);   // col-major staging
  const int t = threadIdx.x, warp = t >> 5;
  const int wm = warp & 3, wn = warp >> 2;                        // 4 x 2 warps, 32 x 32 outputs each
  const int m0 = blockIdx.x * BM;
  const float* Ab = p.A + (long)blockIdx.z * p.a_bs;
  const float* Bb = p.B + (long)blockIdx.z * p.b_bs;
  float* Cb = p.C + (long)blockIdx.z * p.c_bs;
  const int K16 = (p.K + 15) & ~15;
  for (int e = t; e < K16 * BM; e += 256) {
    const int ml = e % BM, kl = e / BM;
    const int m = m0 + ml;
    As[kl * LDA + ml] = __float2bfloat16_rn((m < p.M && kl < p.K) ? __ldg(Ab + m + kl * p.lda_k) : 0.f);
  }
  const bool m_fast = p.ldc_m == 1;
  const bool vec4 = m_fast && p.bias == nullptr && !p.act && (p.M & 3) == 0 && (p.ldc_n & 3) == 0 && (p.c_bs & 3) == 0 &&
                    ((uintptr_t)p.C & 15) == 0;
  for (int n0 = 0; n0 < p.N; n0 += BN) {
    for (int e = t; e < K16 * BN; e += 256) {
      const int nl = e % BN, kl = e / BN;
      const int n = n0 + nl;
      Bs[kl * LDB + nl] = __float2bfloat16_rn((n < p.N && kl < p.K) ? __ldg(Bb + kl * p.ldb_k + n) : 0.f);
    }
    __syncthreads();   // As (first tile) / Bs written; the previous tile's Cs reads are done
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) wmma::fill_fragment(acc[i][j], 0.f);
    for (int ks = 0; ks < K16; ks += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> fa[2];
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) wmma::load_matrix_sync(fa[i], &As[ks * LDA + wm * 32 + 16 * i], LDA);
#pragma unroll
      for (int j = 0; j < 2; ++j) wmma::load_matrix_sync(fb[j], &Bs[ks * LDB + wn * 32 + 16 * j], LDB);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) wmma::mma_sync(acc[i][j], fa[i], fb[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        wmma::store_matrix_sync(&Cs[(wn * 32 + 16 * j) * LDC + wm * 32 + 16 * i], acc[i][j], LDC, wmma::mem_col_major);
    __syncthreads();   // Cs complete; every warp is done reading Bs
    if (vec4) {   // C unit-stride along m, rows 16-byte aligned: 16-byte stores, no bias / activation on this path
      for (int e = t; e < BM * BN / 4; e += 256) {
        const int ml = (e % (BM / 4)) * 4, nl = e / (BM / 4);
        const int m = m0 + ml, n = n0 + nl;
        if (m < p.M && n < p.N)
          *reinterpret_cast<float4*>(Cb + m + n * p.ldc_n) = *reinterpret_cast<const float4*>(&Cs[nl * LDC + ml]);
      }
      continue;
    }
    for (int e = t; e < BM * BN; e += 256) {
      const int ml = m_fast ? e % BM : e / BN, nl = m_fast ? e / BM : e % BN;
      const int m = m0 + ml, n = n0 + nl;
      if (m >= p.M || n >= p.N) continue;
      float v = Cs[nl * LDC + ml];
      if (p.bias) v += __ldg(p.bias + n);
      if (p.act) v = lrelu(v);
      Cb[m * p.ldc_m + n * p.ldc_n] = v;
    }
  }
}

template <int BM, int BN>
static int launch_gemm_bf16(const GemmBf16P& p, bool a_col, bool b_row, dim3 grid, cudaStream_t st) {
  if (a_col && b_row) gemm_bf16_kernel<BM, BN, true, true, kGBK><<<grid, 256, 0, st>>>(p);
  else if (a_col) gemm_bf16_kernel<BM, BN, true, false, kGBK><<<grid, 256, 0, st>>>(p);
  else if (b_row) gemm_bf16_kernel<BM, BN, false, true, kGBK><<<grid, 256, 0, st>>>(p);
  else gemm_bf16_kernel<BM, BN, false, false, kGBKWide><<<grid, 256, 0, st>>>(p);
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
  if (!p.atomic && a_col && b_row && k <= kSKMax && n >= 128) {
    constexpr int kSKSmem = (kSKMax * (128 + 8) + kSKMax * (64 + 8)) * 2 + 64 * (128 + 4) * 4;
    if (int rc = ensure_dyn_smem((const void*)gemm_bf16_shortk_kernel, kSKSmem)) return rc;
    gemm_bf16_shortk_kernel<<<dim3(ceil_div(m, 128), 1, batch), 256, kSKSmem, st>>>(p);
    return check_launch("gemm_bf16_shortk_kernel");
  }
  const bool wide_n = m <= 64;   // M = 64 (weight gradient): 64 x 128 tiles; else 128 x 64
  const int bm = wide_n ? 64 : 128, bn = wide_n ? 128 : 64;
  dim3 grid(ceil_div(m, bm), ceil_div(n, bn), batch);
  if (p.atomic && deterministic()) {
    // batch-reduced C in a fixed order: one launch per image, whole K per CTA -> a single contributor per output
    // element per launch, launches ordered by the stream
    for (int bi = 0; bi < batch; ++bi) {
      GemmBf16P q = p;
      q.A = a + (long)bi * a_batch_stride; q.B = b + (long)bi * b_batch_stride; q.C = c + (long)bi * c_batch_stride;
      q.batch = 1; q.kchunks = 1;
      int rc = wide_n ? launch_gemm_bf16<64, 128>(q, a_col, b_row, dim3(grid.x, grid.y, 1), st)
                      : launch_gemm_bf16<128, 64>(q, a_col, b_row, dim3(grid.x, grid.y, 1), st);
      if (rc) return rc;
    }
    return DBM_OK;
  }
  if (p.atomic) {
    // split (image, K chunk) pairs over ~4 CTAs per SM in total
    p.kchunks = 1;
    long tiles = (long)grid.x * grid.y;
    long z = (4L * num_sms() + tiles - 1) / tiles;
    if (z < 1) z = 1;
    const int bk = (!a_col && !b_row) ? kGBKWide : kGBK;
    while ((long)batch * p.kchunks < z && p.kchunks < 64 && k / (p.kchunks * 2) >= 4 * bk) p.kchunks *= 2;
    if (z > (long)batch * p.kchunks) z = (long)batch * p.kchunks;
    // K chunks must be whole K tiles so that tiles never straddle a chunk boundary
    const int klen = (k + p.kchunks - 1) / p.kchunks;
    if (klen % bk != 0 && p.kchunks > 1) p.kchunks = 1;
    if (z > (long)batch * p.kchunks) z = (long)batch * p.kchunks;
    grid.z = (unsigned)z;
  }
  return wide_n ? launch_gemm_bf16<64, 128>(p, a_col, b_row, grid, st) : launch_gemm_bf16<128, 64>(p, a_col, b_row, grid, st);
}
