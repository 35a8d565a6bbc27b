// Fused generator input block for the tensor-core path
// (DeepbedmapInputBlock.forward, srgan_train.py:256-266): the four valid-padded strided convs
//   conv_on_X  1->32 k3 s1 | conv_on_W1 1->32 k30 s10 | conv_on_W2 2->32 k6 s2 | conv_on_W3 1->32 k3 s1
// and F.concat, computed in fp32 on the CUDA cores (raw-metre inputs stay fp32) and written once
// as the 128-channel bf16 slab8 operand of the pre-residual conv. Small-channel direct conv:
// the 100x180 REMA window of an 8x16-pixel output tile and the whole 900x32 filter are staged in
// shared memory; each thread owns 2 pixels x 8 channels (2 scalar + 2 vector LDS per 16 FMA).
#include "common.cuh"

namespace dbm {

constexpr int kSTH = 8, kSTW = 16;                 // output tile
constexpr int kW1Rows = (kSTH - 1) * 10 + 30;      // 100
constexpr int kW1Cols = (kSTW - 1) * 10 + 30;      // 180
constexpr int kXRows = kSTH + 2, kXCols = kSTW + 2;                    // 10 x 18
constexpr int kW2Rows = (kSTH - 1) * 2 + 6, kW2Cols = (kSTW - 1) * 2 + 6;  // 20 x 36
constexpr int kStemSmemFloats = kW1Rows * kW1Cols + 900 * 32 + 2 * kXRows * kXCols + 2 * kW2Rows * kW2Cols + 90 * 32;
constexpr int kStemSmem = kStemSmemFloats * 4;
constexpr int kStemSmallSmem = (2 * kXRows * kXCols + 2 * kW2Rows * kW2Cols + 90 * 32) * 4;   // without the W1 parts

struct StemParams {
  const float *x, *w1, *w2, *w3;   // (N,1,h,w) (N,1,10h,10w) (N,2,2h,2w) (N,1,h,w)
  const float* wt1;                // [900][32]   conv_on_W1 filter, tap-major
  const float* wts;                // [90][32]    conv_on_X (9) | conv_on_W2 (72) | conv_on_W3 (9), tap-major
  const float* bias;               // [128]       X | W1 | W2 | W3
  __nv_bfloat16* out;              // slab8 [N][out_cs_total][H][W][8], channels written at slab out_cs0..+16
  int out_cs_total, out_cs0;
  int N, h, w, H, W, tiles_x, tiles_y;
  int flat_Pg, flat_G0;            // > 0: write the flat-padded layout [16][Pg][8] of umma_flat.cu / umma_local.cu instead
};

__device__ __forceinline__ void store_pair(const StemParams& p, int n, int slab, int y, int x, const float (&a)[2][8],
                                           const float* __restrict__ bias8) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (y < p.H && x + i < p.W) {
      __nv_bfloat162 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = __floats2bfloat162_rn(a[i][2 * k] + bias8[2 * k], a[i][2 * k + 1] + bias8[2 * k + 1]);
      const size_t at = p.flat_Pg > 0
                            ? (size_t)slab * p.flat_Pg + p.flat_G0 + ((size_t)n * (p.H + 2) + y + 1) * (p.W + 2) + x + i + 1
                            : (((size_t)n * p.out_cs_total + p.out_cs0 + slab) * p.H + y) * p.W + x + i;
      *reinterpret_cast<uint4*>(p.out + at * 8) = *reinterpret_cast<uint4*>(v);
    }
  }
}

__device__ __forceinline__ void fma_pair(float (&a)[2][8], float i0, float i1, const float* __restrict__ w8) {
  const float4 wa = *reinterpret_cast<const float4*>(w8);
  const float4 wb = *reinterpret_cast<const float4*>(w8 + 4);
  const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[0][k] = fmaf(i0, ww[k], a[0][k]);
    a[1][k] = fmaf(i1, ww[k], a[1][k]);
  }
}

// kW1 = false: conv_on_X / W2 / W3 only (18 KB of shared memory, several blocks per SM); conv_on_W1 then runs on
// the tensor cores (w1_s2d_split_kernel + umma_conv3x3_kernel, below).
template <bool kW1>
__global__ void __launch_bounds__(256, kW1 ? 1 : 4) stem_kernel(const StemParams p) {
  extern __shared__ __align__(16) float sm[];
  float* s_w1in = sm;                                   // [100][180]
  float* s_w1w = s_w1in + (kW1 ? kW1Rows * kW1Cols : 0);  // [900][32]
  float* s_x = s_w1w + (kW1 ? 900 * 32 : 0);            // [10][18]
  float* s_w3 = s_x + kXRows * kXCols;                  // [10][18]
  float* s_w2 = s_w3 + kXRows * kXCols;                 // [2][20][36]
  float* s_ws = s_w2 + 2 * kW2Rows * kW2Cols;           // [90][32]

  const int t = threadIdx.x;
  int b = blockIdx.x;
  const int tx = b % p.tiles_x; b /= p.tiles_x;
  const int ty = b % p.tiles_y;
  const int n = b / p.tiles_y;
  const int oy0 = ty * kSTH, ox0 = tx * kSTW;

  // ---- stage filters and input windows (zero fill beyond the image; never read for valid outputs)
  if (kW1)
    for (int i = t; i < 900 * 32 / 4; i += 256)
      reinterpret_cast<float4*>(s_w1w)[i] = __ldg(reinterpret_cast<const float4*>(p.wt1) + i);
  for (int i = t; i < 90 * 32 / 4; i += 256)
    reinterpret_cast<float4*>(s_ws)[i] = __ldg(reinterpret_cast<const float4*>(p.wts) + i);
  if (kW1) {
    const int H1 = 10 * p.h, W1 = 10 * p.w;
    const float* src = p.w1 + (size_t)n * H1 * W1;
    const int r0 = oy0 * 10, c0 = ox0 * 10;
    for (int i = t; i < kW1Rows * kW1Cols; i += 256) {
      const int r = i / kW1Cols, c = i - r * kW1Cols;
      const int gr = r0 + r, gc = c0 + c;
      s_w1in[i] = (gr < H1 && gc < W1) ? __ldg(src + (size_t)gr * W1 + gc) : 0.f;
    }
  }
  for (int i = t; i < kXRows * kXCols; i += 256) {
    const int r = i / kXCols, c = i - r * kXCols;
    const int gr = oy0 + r, gc = ox0 + c;
    const bool ok = gr < p.h && gc < p.w;
    s_x[i] = ok ? __ldg(p.x + ((size_t)n * p.h + gr) * p.w + gc) : 0.f;
    s_w3[i] = ok ? __ldg(p.w3 + ((size_t)n * p.h + gr) * p.w + gc) : 0.f;
  }
  {
    const int H2 = 2 * p.h, W2 = 2 * p.w;
    for (int i = t; i < 2 * kW2Rows * kW2Cols; i += 256) {
      const int ch = i / (kW2Rows * kW2Cols), rem = i - ch * (kW2Rows * kW2Cols);
      const int r = rem / kW2Cols, c = rem - r * kW2Cols;
      const int gr = 2 * oy0 + r, gc = 2 * ox0 + c;
      s_w2[i] = (gr < H2 && gc < W2) ? __ldg(p.w2 + (((size_t)n * 2 + ch) * H2 + gr) * W2 + gc) : 0.f;
    }
  }
  __syncthreads();

  const int cg = t & 3;        // 8-channel group of the 32 outputs of each conv
  const int pp = t >> 2;       // pixel pair 0..63
  const int r = pp >> 3, c = (pp & 7) * 2;
  const int y = oy0 + r, x = ox0 + c;
  float a[2][8];

  // ---- conv_on_W1: k30 s10 ----
  if (kW1) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[i][k] = 0.f;
    const float* in = s_w1in + (r * 10) * kW1Cols + c * 10;
    const float* wq = s_w1w + cg * 8;
    // The two pixels' windows are columns [0, 30) and [10, 40) of the same 40 floats: one row of the
    // pair is fetched as ten aligned 16-byte LDS (c is even, so the window starts on a 16-byte
    // boundary) instead of 60 scalar ones -- 70 LDS per 480 FMA; the scalar form (120 per 480)
    // kept the LSU pipe as busy as the FMA pipe.
#pragma unroll 1
    for (int ky = 0; ky < 30; ++ky) {
      float row[40];
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(in + ky * kW1Cols + 4 * q);
        row[4 * q] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int kx = 0; kx < 30; ++kx) fma_pair(a, row[kx], row[kx + 10], wq + (ky * 30 + kx) * 32);
    }
    store_pair(p, n, 4 + cg, y, x, a, p.bias + 32 + cg * 8);
  }

  // ---- conv_on_X: k3 s1 ----
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[i][k] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
      fma_pair(a, s_x[(r + ky) * kXCols + c + kx], s_x[(r + ky) * kXCols + c + kx + 1],
               s_ws + (ky * 3 + kx) * 32 + cg * 8);
  store_pair(p, n, 0 + cg, y, x, a, p.bias + 0 + cg * 8);

  // ---- conv_on_W2: 2 channels, k6 s2 ----
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[i][k] = 0.f;
  for (int ch = 0; ch < 2; ++ch)
    for (int ky = 0; ky < 6; ++ky)
#pragma unroll
      for (int kx = 0; kx < 6; ++kx) {
        const float* in = s_w2 + (ch * kW2Rows + 2 * r + ky) * kW2Cols + 2 * c + kx;
        fma_pair(a, in[0], in[2], s_ws + (9 + ch * 36 + ky * 6 + kx) * 32 + cg * 8);
      }
  store_pair(p, n, 8 + cg, y, x, a, p.bias + 64 + cg * 8);

  // ---- conv_on_W3: k3 s1 ----
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[i][k] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
      fma_pair(a, s_w3[(r + ky) * kXCols + c + kx], s_w3[(r + ky) * kXCols + c + kx + 1],
               s_ws + (81 + ky * 3 + kx) * 32 + cg * 8);
  store_pair(p, n, 12 + cg, y, x, a, p.bias + 96 + cg * 8);
}

// ---- conv_on_W1 (1 -> 32, k30 s10, valid) on the tensor cores -----------------------------------------------
// A k30 s10 convolution is a 3x3 stride-1 valid convolution over the 10x10 space-to-depth of its input: cell
// (cy, cx) holds the 100 values W1[10cy + dy][10cx + dx]. REMA elevations are metres (0 .. 4500) and must not lose
// their low bits to bf16, so every operand is split into two bf16 terms, x = x_hi + x_lo, w = w_hi + w_lo (each
// split exact to 2^-18), and the GEMM contracts the 320-channel operand
//   [x_hi (100 -> 104) | x_lo (104) | x_hi (104) | 0 (8)]  against  [w_hi | w_hi | w_lo | 0]
// = x_hi w_hi + x_lo w_hi + x_hi w_lo in fp32 accumulators: only the x_lo w_lo term (2^-18 relative) is dropped,
// far below the bf16 rounding the 128-channel stem output gets anyway as the pre-residual conv's operand.
constexpr int kW1Jp = 104, kW1Cs = 40;   // padded cell size, slabs of the split operand

__global__ void __launch_bounds__(256) w1_s2d_split_kernel(const float* __restrict__ w1,
                                                           __nv_bfloat16* __restrict__ out, int h, int w) {
  __shared__ float cell[10][32 * 10 + 2];
  const int cx0 = blockIdx.x * 32, cy = blockIdx.y, n = blockIdx.z;
  const int W1 = 10 * w;
  const float* src = w1 + ((size_t)n * 10 * h + (size_t)10 * cy) * W1 + (size_t)10 * cx0;
  const int ncol = min(320, W1 - 10 * cx0);
  for (int i = threadIdx.x; i < 10 * 320; i += 256) {
    const int r = i / 320, c = i - r * 320;
    cell[r][c] = c < ncol ? __ldg(src + (size_t)r * W1 + c) : 0.f;
  }
  __syncthreads();
  const size_t plane = (size_t)h * w;
  for (int i = threadIdx.x; i < 14 * 32; i += 256) {
    const int sl = i >> 5, cx = cx0 + (i & 31);
    if (cx >= w) continue;
    __nv_bfloat16* dst = out + (((size_t)n * kW1Cs) * plane + (size_t)cy * w + cx) * 8;
    if (sl == 13) {   // slab 39: zero padding of the K axis
      *reinterpret_cast<uint4*>(dst + 39 * plane * 8) = make_uint4(0, 0, 0, 0);
      continue;
    }
    __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = 8 * sl + k;
      const float v = j < 100 ? cell[j / 10][(i & 31) * 10 + j % 10] : 0.f;
      hi[k] = __float2bfloat16_rn(v);
      lo[k] = __float2bfloat16_rn(v - __bfloat162float(hi[k]));
    }
    *reinterpret_cast<uint4*>(dst + (size_t)sl * plane * 8) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + (size_t)(13 + sl) * plane * 8) = *reinterpret_cast<uint4*>(lo);
    *reinterpret_cast<uint4*>(dst + (size_t)(26 + sl) * plane * 8) = *reinterpret_cast<uint4*>(hi);
  }
}

// conv_on_W1 filter (32, 1, 30, 30) fp32 -> bf16 UMMA operand image [320/32][9 taps][4][4][8 cout][8 cin] of the
// split contraction above (K index c: group c / 104 in {w_hi, w_hi, w_lo, 0}, cell element j = c % 104).
__global__ void pack_stem_w1_kernel(const float* __restrict__ wf, __nv_bfloat16* __restrict__ out) {
  const int total = 9 * 320 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int c8 = t % 8; t /= 8;
    const int o8 = t % 8; t /= 8;
    const int cg = t % 4; t /= 4;
    const int ksl = t % 4; t /= 4;
    const int tap = t % 9; t /= 9;
    const int kc = t;
    const int o = cg * 8 + o8, c = kc * 32 + ksl * 8 + c8;
    const int g = c / kW1Jp, j = c - g * kW1Jp;
    float v = 0.f;
    if (g < 3 && j < 100) {
      const float wv = wf[o * 900 + (10 * (tap / 3) + j / 10) * 30 + 10 * (tap % 3) + j % 10];
      const float h = __bfloat162float(__float2bfloat16_rn(wv));
      v = g < 2 ? h : wv - h;
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// dst[c][r] = src[r][c]  (filter (32, taps) -> tap-major (taps, 32)); tiny, run once per weight update
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  const int total = rows * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / cols, c = i - r * cols;
    dst[(size_t)c * rows + r] = src[i];
  }
}

}  // namespace dbm

using namespace dbm;

extern "C" int dbm_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t stream) {
  DBM_REQUIRE(rows > 0 && cols > 0, "transpose: empty matrix");
  transpose_kernel<<<ceil_div((long)rows * cols, 256), 256, 0, stream>>>(src, dst, rows, cols);
  return check_launch("transpose");
}

static int stem_launch(const float* x, const float* w1, const float* w2, const float* w3,
                       const float* w1_filter_tapmajor, const float* small_filters_tapmajor, const float* bias128,
                       void* out_slab8, int out_cs_total, int out_cs0, int n, int h, int w, int flat, cudaStream_t stream) {
  DBM_REQUIRE(n > 0 && h >= 3 && w >= 3, "stem: input %dx%d too small (need >= 3x3)", h, w);
  DBM_REQUIRE(out_cs_total >= out_cs0 + 16, "stem: output needs 16 slabs");
  const bool with_w1 = w1 != nullptr;
  if (int rc = ensure_dyn_smem((const void*)stem_kernel<true>, kStemSmem)) return rc;
  StemParams p;
  p.x = x; p.w1 = w1; p.w2 = w2; p.w3 = w3;
  p.wt1 = w1_filter_tapmajor; p.wts = small_filters_tapmajor; p.bias = bias128;
  p.out = (__nv_bfloat16*)out_slab8; p.out_cs_total = out_cs_total; p.out_cs0 = out_cs0;
  p.N = n; p.h = h; p.w = w; p.H = h - 2; p.W = w - 2;
  p.flat_Pg = 0; p.flat_G0 = 0;
  if (flat) {   // geometry of flat_geom() in umma_flat.cu for n images of H x W
    const int halo = p.W + 3;
    p.flat_G0 = (halo + 7) & ~7;
    p.flat_Pg = 2 * p.flat_G0 + 128 * ceil_div((long)n * (p.H + 2) * (p.W + 2), 128);
  }
  p.tiles_x = ceil_div(p.W, kSTW); p.tiles_y = ceil_div(p.H, kSTH);
  const long blocks = (long)n * p.tiles_x * p.tiles_y;
  DBM_REQUIRE(blocks < (1L << 31), "stem: too many tiles");
  if (with_w1)
    stem_kernel<true><<<(int)blocks, 256, kStemSmem, stream>>>(p);
  else
    stem_kernel<false><<<(int)blocks, 256, kStemSmallSmem, stream>>>(p);
  return check_launch("stem_kernel");
}

extern "C" int dbm_stem_w1_s2d(const float* w1, void* s2d_slab8, int n, int h, int w, cudaStream_t stream) {
  DBM_REQUIRE(n > 0 && h >= 3 && w >= 3 && h < 65536 && n < 65536, "stem_w1_s2d: bad shape %d x %d x %d", n, h, w);
  DBM_REQUIRE(((uintptr_t)s2d_slab8 & 15) == 0, "stem_w1_s2d: unaligned output");
  w1_s2d_split_kernel<<<dim3(ceil_div(w, 32), h, n), 256, 0, stream>>>(w1, (__nv_bfloat16*)s2d_slab8, h, w);
  return check_launch("w1_s2d_split_kernel");
}

extern "C" int dbm_pack_stem_w1(const float* w1_filter, void* packed_bf16, cudaStream_t stream) {
  pack_stem_w1_kernel<<<90, 256, 0, stream>>>(w1_filter, (__nv_bfloat16*)packed_bf16);
  return check_launch("pack_stem_w1_kernel");
}

extern "C" int dbm_stem_fwd_slab8(const float* x, const float* w1, const float* w2, const float* w3,
                                  const float* w1_filter_tapmajor, const float* small_filters_tapmajor,
                                  const float* bias128, void* out_slab8, int out_cs_total, int out_cs0, int n, int h,
                                  int w, cudaStream_t stream) {
  return stem_launch(x, w1, w2, w3, w1_filter_tapmajor, small_filters_tapmajor, bias128, out_slab8, out_cs_total, out_cs0,
                     n, h, w, 0, stream);
}

extern "C" int dbm_stem_fwd_flat(const float* x, const float* w1, const float* w2, const float* w3,
                                 const float* w1_filter_tapmajor, const float* small_filters_tapmajor,
                                 const float* bias128, void* out_flat, int n, int h, int w, cudaStream_t stream) {
  return stem_launch(x, w1, w2, w3, w1_filter_tapmajor, small_filters_tapmajor, bias128, out_flat, 16, 0, n, h, w, 1,
                     stream);
}
