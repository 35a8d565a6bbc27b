// Full-precision (fp32 CUDA-core) implicit-GEMM kernels: convolution forward, data gradient,
// weight gradient and plain GEMM. These are the exact-arithmetic path of the product (the
// reference computes in float32, paper/tc-2020-74.tex:629-630): the small-channel stem
// (srgan_train.py:256-266), every discriminator conv (:649-689), the linear layers
// (:694-696) and all backward passes run here; the generator trunk runs on the tcgen05
// kernel in umma_conv3x3.cu when precision == bf16.
//
// One 64x64x16 tiled kernel; the A/B element fetchers are specialised per mode and the
// filter size / stride are template constants so index decomposition is mul-shift only.
#include "common.cuh"

namespace dbm {

enum { kFwd = 0, kDgrad = 1, kWgrad = 2, kPlain = 3 };

struct GemmP {
  int M, N, K;
  int C, H, W, O, P, HO, WO;  // conv geometry (x: [*,C,H,W], y: [*,O,HO,WO])
  long x_bs, y_bs;            // batch strides of x-side / y-side tensors (channel-slice views)
  const float* A;
  const float* B;
  float* Cc;
  const float* bias;
  int act, accumulate, kchunk;  // kchunk: K range per blockIdx.z (split-K => atomicAdd)
  long lda_m, lda_k, ldb_k, ldb_n, ldc_m, ldc_n;  // plain GEMM strides
  long a_bs, b_bs, c_bs;                          // plain GEMM: per-batch (blockIdx.z) strides
  int atomic;                                     // plain GEMM: atomicAdd epilogue (batch-reduced C)
};

constexpr int BK = 16, LDS = 68;

// TM x TN = outputs per thread (tile = 16 TM x 16 TN): the narrow variants serve the generator's
// 32-channel dense-block convs (forward / dgrad with N = 32: TN = 2; wgrad with M = 32: TM = 2)
// without computing a half-empty 64-wide tile.
template <int MODE, int KH, int S, int TM = 4, int TN = 4>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmP p) {
  constexpr int KK = KH * KH;
  constexpr int BM = 16 * TM, BN = 16 * TN;
  static_assert(TM == 4 || MODE == kWgrad, "narrow M tiles: wgrad only");
  static_assert(TN == 4 || MODE == kFwd || MODE == kDgrad, "narrow N tiles: forward / dgrad only");
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = (MODE == kPlain) ? 0 : blockIdx.z * p.kchunk;
  const int kend = (MODE == kPlain) ? p.K : min(p.K, kbeg + p.kchunk);
  const float* __restrict__ Ap = p.A + (MODE == kPlain ? blockIdx.z * p.a_bs : 0);
  const float* __restrict__ Bp = p.B + (MODE == kPlain ? blockIdx.z * p.b_bs : 0);
  float* __restrict__ Cp = p.Cc + (MODE == kPlain ? blockIdx.z * p.c_bs : 0);
  // C-tile thread mapping: the "fast" output dim is interleaved across tx for coalesced stores
  // (plain GEMM: KH == 2 selects the m-fast variant, used when C is unit-stride along m)
  constexpr bool C_MFAST = (MODE == kFwd || MODE == kDgrad || (MODE == kPlain && KH == 2));
  const int tx = t & 15, ty = t >> 4;

  // ---- per-thread loader state hoisted out of the K loop ----
  // A: m-contiguous mapping for fwd/dgrad (thread owns one m), k-contiguous otherwise
  const int HW = p.H * p.W, HOWO = p.HO * p.WO;
  int a_m = 0, a_n = 0, a_h = 0, a_w = 0;
  bool a_ok = false;
  if (MODE == kFwd || MODE == kDgrad) {
    a_m = m0 + (t & 63);
    a_ok = a_m < p.M;
    if (a_ok) {
      const int sp = (MODE == kFwd) ? HOWO : HW;
      const int wd = (MODE == kFwd) ? p.WO : p.W;
      a_n = a_m / sp;
      const int r = a_m - a_n * sp;
      a_h = r / wd;
      a_w = r - a_h * wd;
    }
  }
  // wgrad: the 4 B columns (c,ky,kx) this thread loads are fixed
  int bw_c[4], bw_ky[4], bw_kx[4];
  bool bw_ok[4];
  if (MODE == kWgrad) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + (t >> 4) + 16 * i;
      bw_ok[i] = n < p.N;
      const int c = n / KK, r = n - c * KK;
      bw_c[i] = c; bw_ky[i] = r / KH; bw_kx[i] = r - (r / KH) * KH;
    }
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // Global -> register fetch of the next K tile is issued before the math of the current one
  // (register double buffering): these layers are small (M ~ 10^4), so latency, not FLOPs, bounds them.
  float ra[4], rb[4];
  const bool a_kmajor = (MODE == kFwd || MODE == kDgrad || (MODE == kPlain && p.lda_m == 1));
  const bool b_kmajor = (MODE == kPlain && p.ldb_n == 1);
  auto fetch = [&](int kt) {
    // ---------------- load A tile -> As[k][m] ----------------
    if (MODE == kFwd) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kl = (t >> 6) + 4 * i, k = kt + kl;
        float v = 0.f;
        if (a_ok && k < kend) {
          const int c = k / KK, r = k - c * KK, ky = r / KH, kx = r - ky * KH;
          const int hi = a_h * S - p.P + ky, wi = a_w * S - p.P + kx;
          if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
            v = __ldg(p.A + a_n * p.x_bs + (long)c * HW + hi * p.W + wi);
        }
        ra[i] = v;
      }
    } else if (MODE == kDgrad) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kl = (t >> 6) + 4 * i, k = kt + kl;
        float v = 0.f;
        if (a_ok && k < kend) {
          const int o = k / KK, r = k - o * KK, ky = r / KH, kx = r - ky * KH;
          const int hn = a_h + p.P - ky, wn = a_w + p.P - kx;
          if (hn >= 0 && wn >= 0 && (hn % S) == 0 && (wn % S) == 0) {
            const int ho = hn / S, wo = wn / S;
            if (ho < p.HO && wo < p.WO) v = __ldg(p.A + a_n * p.y_bs + (long)o * HOWO + ho * p.WO + wo);
          }
        }
        ra[i] = v;
      }
    } else if (MODE == kWgrad) {
      const int kl = t & 15, k = kt + kl;
      int n_img = 0, rem = 0;
      const bool kok = k < kend;
      if (kok) { n_img = k / HOWO; rem = k - n_img * HOWO; }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int ml = (t >> 4) + 16 * i, m = m0 + ml;
        float v = 0.f;
        if (kok && m < p.M) v = __ldg(p.A + n_img * p.y_bs + (long)m * HOWO + rem);
        ra[i] = v;
      }
    } else if (p.lda_m == 1) {  // plain, A unit-stride along m
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kl = (t >> 6) + 4 * i, k = kt + kl, m = m0 + (t & 63);
        float v = 0.f;
        if (k < kend && m < p.M) v = __ldg(Ap + m + k * p.lda_k);
        ra[i] = v;
      }
    } else {  // plain, A unit-stride along k
      const int kl = t & 15, k = kt + kl;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ml = (t >> 4) + 16 * i, m = m0 + ml;
        float v = 0.f;
        if (k < kend && m < p.M) v = __ldg(Ap + m * p.lda_m + k * p.lda_k);
        ra[i] = v;
      }
    }
    // ---------------- load B tile -> Bs[k][n] ----------------
    if (MODE == kFwd) {
      const int kl = t & 15, k = kt + kl;
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int nl = (t >> 4) + 16 * i, n = n0 + nl;
        float v = 0.f;
        if (k < kend && n < p.N) v = __ldg(p.B + (long)n * p.K + k);
        rb[i] = v;
      }
    } else if (MODE == kDgrad) {
      const int kl = t & 15, k = kt + kl;
      const int o = k / KK, r = k - o * KK;
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int nl = (t >> 4) + 16 * i, n = n0 + nl;
        float v = 0.f;
        if (k < kend && n < p.N) v = __ldg(p.B + ((long)o * p.C + n) * KK + r);
        rb[i] = v;
      }
    } else if (MODE == kWgrad) {
      const int kl = t & 15, k = kt + kl;
      int n_img = 0, ho = 0, wo = 0;
      const bool kok = k < kend;
      if (kok) {
        n_img = k / HOWO;
        const int rem = k - n_img * HOWO;
        ho = rem / p.WO; wo = rem - ho * p.WO;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nl = (t >> 4) + 16 * i;
        float v = 0.f;
        if (kok && bw_ok[i]) {
          const int hi = ho * S - p.P + bw_ky[i], wi = wo * S - p.P + bw_kx[i];
          if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
            v = __ldg(p.B + n_img * p.x_bs + (long)bw_c[i] * HW + hi * p.W + wi);
        }
        rb[i] = v;
      }
    } else {  // plain: pick the mapping along B's unit-stride dim
      if (p.ldb_n == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kl = (t >> 6) + 4 * i, k = kt + kl, n = n0 + (t & 63);
          float v = 0.f;
          if (k < kend && n < p.N) v = __ldg(Bp + k * p.ldb_k + n);
          rb[i] = v;
        }
      } else {
        const int kl = t & 15, k = kt + kl;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = (t >> 4) + 16 * i, n = n0 + nl;
          float v = 0.f;
          if (k < kend && n < p.N) v = __ldg(Bp + k * p.ldb_k + n * p.ldb_n);
          rb[i] = v;
        }
      }
    }
  };
  auto commit = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (a_kmajor) As[(t >> 6) + 4 * i][t & 63] = ra[i];
      else if (i < TM) As[t & 15][(t >> 4) + 16 * i] = ra[i];
      if (b_kmajor) Bs[(t >> 6) + 4 * i][t & 63] = rb[i];
      else if (i < TN) Bs[t & 15][(t >> 4) + 16 * i] = rb[i];
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int kt = kbeg; kt < kend; kt += BK) {
    commit();
    __syncthreads();
    if (kt + BK < kend) fetch(kt + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      if (C_MFAST) {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][ty * TN + j];
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx + 16 * j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---------------- epilogue ----------------
  const bool atomic = (MODE == kPlain) ? (p.atomic != 0) : (gridDim.z > 1);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (C_MFAST ? tx + 16 * i : ty * TM + i);
    if (m >= p.M) continue;
    long mbase = 0;
    if (MODE == kFwd) {
      const int n_img = m / HOWO, r = m - n_img * HOWO;
      mbase = n_img * p.y_bs + r;
    } else if (MODE == kDgrad) {
      const int n_img = m / HW, r = m - n_img * HW;
      mbase = n_img * p.x_bs + r;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (C_MFAST ? ty * TN + j : tx + 16 * j);
      if (n >= p.N) continue;
      float v = acc[i][j];
      long idx;
      if (MODE == kFwd) idx = mbase + (long)n * HOWO;
      else if (MODE == kDgrad) idx = mbase + (long)n * HW;
      else if (MODE == kWgrad) idx = (long)m * p.N + n;
      else idx = m * p.ldc_m + n * p.ldc_n;
      if (p.bias && (MODE == kPlain || blockIdx.z == 0)) v += __ldg(p.bias + n);
      if (atomic) {
        atomicAdd(Cp + idx, v);
      } else {
        if (p.accumulate) v += Cp[idx];
        if (p.act) v = lrelu(v);
        Cp[idx] = v;
      }
    }
  }
}

template <int MODE>
static int dispatch(const GemmP& p, int kh, int s, int splits, cudaStream_t st) {
  constexpr int BM = 64, BN = 64;
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), splits);
  if (MODE == kPlain) {
    if (p.ldc_m == 1 && p.ldc_n != 1) gemm_f32_kernel<kPlain, 2, 1><<<grid, 256, 0, st>>>(p);
    else gemm_f32_kernel<kPlain, 1, 1><<<grid, 256, 0, st>>>(p);
  } else if (kh == 3 && s == 1) {
    if constexpr (MODE == kFwd || MODE == kDgrad) {
      if (p.N <= 32) {
        gemm_f32_kernel<MODE, 3, 1, 4, 2><<<dim3(grid.x, 1, splits), 256, 0, st>>>(p);
        return check_launch("gemm_f32_kernel");
      }
    }
    if constexpr (MODE == kWgrad) {
      if (p.M <= 32) {
        gemm_f32_kernel<MODE, 3, 1, 2, 4><<<dim3(1, grid.y, splits), 256, 0, st>>>(p);
        return check_launch("gemm_f32_kernel");
      }
    }
    gemm_f32_kernel<MODE, 3, 1><<<grid, 256, 0, st>>>(p);
  } else if (kh == 4 && s == 2) {
    gemm_f32_kernel<MODE, 4, 2><<<grid, 256, 0, st>>>(p);
  } else if (kh == 6 && s == 2) {
    gemm_f32_kernel<MODE, 6, 2><<<grid, 256, 0, st>>>(p);
  } else if (kh == 30 && s == 10) {
    gemm_f32_kernel<MODE, 30, 10><<<grid, 256, 0, st>>>(p);
  } else {
    set_error("conv2d: unsupported (ksize=%d, stride=%d); supported: (3,1) (4,2) (6,2) (30,10)", kh, s);
    return DBM_ERR_INVALID;
  }
  return check_launch("gemm_f32_kernel");
}

static int conv_geom(GemmP& p, int n, int c, int h, int w, int o, int k, int s, int pad, long x_bs, long y_bs) {
  DBM_REQUIRE(n > 0 && c > 0 && o > 0, "conv2d: empty tensor (n=%d c=%d o=%d)", n, c, o);
  DBM_REQUIRE(h + 2 * pad >= k && w + 2 * pad >= k, "conv2d: input %dx%d smaller than kernel %d", h, w, k);
  p.C = c; p.H = h; p.W = w; p.O = o; p.P = pad;
  p.HO = (h + 2 * pad - k) / s + 1;
  p.WO = (w + 2 * pad - k) / s + 1;
  p.x_bs = x_bs ? x_bs : (long)c * h * w;
  p.y_bs = y_bs ? y_bs : (long)o * p.HO * p.WO;
  p.act = 0; p.accumulate = 0; p.bias = nullptr;
  return DBM_OK;
}

}  // namespace dbm

using namespace dbm;

extern "C" int dbm_conv2d_fwd_f32(const float* x, long x_batch_stride, const float* w, const float* bias, float* y,
                                  long y_batch_stride, int n, int c, int h, int wd, int o, int ksize, int stride,
                                  int pad, int act, cudaStream_t st) {
  GemmP p{};
  int rc = conv_geom(p, n, c, h, wd, o, ksize, stride, pad, x_batch_stride, y_batch_stride);
  if (rc) return rc;
  p.M = n * p.HO * p.WO; p.N = o; p.K = c * ksize * ksize; p.kchunk = p.K;
  p.A = x; p.B = w; p.Cc = y; p.bias = bias; p.act = act;
  return dispatch<kFwd>(p, ksize, stride, 1, st);
}

extern "C" int dbm_conv2d_bwd_data_f32(const float* dy, long y_batch_stride, const float* w, float* dx,
                                       long x_batch_stride, int n, int c, int h, int wd, int o, int ksize,
                                       int stride, int pad, int accumulate, cudaStream_t st) {
  GemmP p{};
  int rc = conv_geom(p, n, c, h, wd, o, ksize, stride, pad, x_batch_stride, y_batch_stride);
  if (rc) return rc;
  p.M = n * h * wd; p.N = c; p.K = o * ksize * ksize; p.kchunk = p.K;
  p.A = dy; p.B = w; p.Cc = dx; p.accumulate = accumulate;
  return dispatch<kDgrad>(p, ksize, stride, 1, st);
}

// dW (O,C,k,k) += sum over batch/pixels; db (O) += sum dy. Both ACCUMULATE (zero them first).
extern "C" int dbm_conv2d_bwd_weight_f32(const float* x, long x_batch_stride, const float* dy, long y_batch_stride,
                                         float* dw, int n, int c, int h, int wd, int o, int ksize, int stride,
                                         int pad, cudaStream_t st) {
  GemmP p{};
  int rc = conv_geom(p, n, c, h, wd, o, ksize, stride, pad, x_batch_stride, y_batch_stride);
  if (rc) return rc;
  p.M = o; p.N = c * ksize * ksize; p.K = n * p.HO * p.WO;
  p.A = dy; p.B = x; p.Cc = dw;
  const int BM = p.M <= 32 ? 32 : 64, BN = 64;
  const int tiles = ceil_div(p.M, BM) * ceil_div(p.N, BN);
  int splits = (4 * num_sms() + tiles - 1) / tiles;
  int max_splits = ceil_div(p.K, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (deterministic()) {   // one CTA owns the whole K range of its tile: dW += in a fixed order
    p.kchunk = ceil_div(p.K, BK) * BK;
    p.accumulate = 1;
    return dispatch<kWgrad>(p, ksize, stride, 1, st);
  }
  if (splits < 2) splits = 2;  // always the atomicAdd (accumulating) epilogue
  p.kchunk = ceil_div(ceil_div(p.K, splits), BK) * BK;
  splits = ceil_div(p.K, p.kchunk);
  if (splits < 2) { splits = 2; }
  return dispatch<kWgrad>(p, ksize, stride, splits, st);
}

// Batched C_b[m,n] (+)= sum_k A_b[m,k] B_b[k,n] (+ bias[n]) with arbitrary element strides.
// accumulate: 0 = overwrite, 1 = C += (per batch, non-atomic), 2 = atomicAdd (use when several
// batches reduce into the same C, i.e. c_batch_stride == 0).
extern "C" int dbm_gemm_f32(const float* a, long lda_m, long lda_k, long a_batch_stride, const float* b, long ldb_k,
                            long ldb_n, long b_batch_stride, float* c, long ldc_m, long ldc_n, long c_batch_stride,
                            const float* bias, int m, int n, int k, int batch, int act, int accumulate,
                            cudaStream_t st) {
  DBM_REQUIRE(m > 0 && n > 0 && k > 0 && batch > 0, "gemm: empty problem %dx%dx%d x%d", m, n, k, batch);
  DBM_REQUIRE(batch <= 65535, "gemm: batch %d too large", batch);
  GemmP p{};
  p.M = m; p.N = n; p.K = k; p.kchunk = k;
  p.A = a; p.B = b; p.Cc = c; p.bias = bias; p.act = act; p.accumulate = (accumulate == 1);
  p.atomic = (accumulate == 2);
  p.lda_m = lda_m; p.lda_k = lda_k; p.ldb_k = ldb_k; p.ldb_n = ldb_n; p.ldc_m = ldc_m; p.ldc_n = ldc_n;
  p.a_bs = a_batch_stride; p.b_bs = b_batch_stride; p.c_bs = c_batch_stride;
  p.HO = p.WO = p.H = p.W = 1;
  if (p.atomic && deterministic() && batch > 1) {
    // batch-reduced C: one launch per batch element, each a single contributor per output element, in stream order
    for (int bi = 0; bi < batch; ++bi) {
      GemmP q = p;
      q.A = a + (long)bi * a_batch_stride;
      q.B = b + (long)bi * b_batch_stride;
      q.Cc = c + (long)bi * c_batch_stride;
      int rc = dispatch<kPlain>(q, 1, 1, 1, st);
      if (rc) return rc;
    }
    return DBM_OK;
  }
  return dispatch<kPlain>(p, 1, 1, batch, st);
}
