// Deformable convolution (L.DeformableConvolution2D, srgan_train.py:506-523, 572-574), exact
// fp32 path: bilinear sampling on the zero-padded input at tap positions displaced by a learned
// offset field (offset channels [0:9] = dx, [9:18] = dy of tap t = ky*3+kx),
// materialised as cols[n][c*9+t][pixel] and contracted with the (O, C*9) filter matrix by the
// fp32 GEMM. The backward kernel scatters d(cols) to d(input) and d(offset).
#include "common.cuh"

namespace dbm {

struct Bilin {
  int x0, y0;
  float fx, fy;
  bool in_range;  // false when the coordinate was clamped (gradient wrt offset is zero)
};

__device__ __forceinline__ Bilin tap_position(const float* __restrict__ off, long off_n, int HW, int p, int t,
                                              int y, int x, int H, int W) {
  // position in the unpadded frame: ox + kx - pad + dx  (pad = 1)
  float px = (float)(x + (t % 3) - 1) + off[off_n + (long)t * HW + p];
  float py = (float)(y + (t / 3) - 1) + off[off_n + (long)(9 + t) * HW + p];
  Bilin b;
  b.in_range = (px >= -2.f && px <= (float)W + 1.f && py >= -2.f && py <= (float)H + 1.f);
  px = fminf(fmaxf(px, -2.f), (float)W + 1.f);
  py = fminf(fmaxf(py, -2.f), (float)H + 1.f);
  const float fx0 = floorf(px), fy0 = floorf(py);
  b.x0 = (int)fx0; b.y0 = (int)fy0;
  b.fx = px - fx0; b.fy = py - fy0;
  return b;
}

__device__ __forceinline__ float at(const float* __restrict__ img, int y, int x, int H, int W) {
  return (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + (long)y * W + x) : 0.f;
}

// cols[n][c*9+t][p]  <- bilinear sample of x[n][c] for tap t at pixel p
__global__ void deform_sample_kernel(const float* __restrict__ x, const float* __restrict__ off,
                                     float* __restrict__ cols, int N, int C, int H, int W) {
  const int HW = H * W;
  const long total = (long)N * 9 * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = i % HW;
    const long r = i / HW;
    const int t = r % 9;
    const int n = r / 9;
    const int y = p / W, xx = p - y * W;
    const Bilin b = tap_position(off, (long)n * 18 * HW, HW, p, t, y, xx, H, W);
    // corner addresses clamped into the image, weights of out-of-image corners zeroed: the 4 x C loads are
    // unconditional and independent, so the channel loop unrolls into batches of loads in flight (the per-load
    // bounds branches of the first version serialised it: 530 us for 382 MB of output at batch 128)
    const bool y0ok = b.y0 >= 0 && b.y0 < H, y1ok = b.y0 + 1 >= 0 && b.y0 + 1 < H;
    const bool x0ok = b.x0 >= 0 && b.x0 < W, x1ok = b.x0 + 1 >= 0 && b.x0 + 1 < W;
    const int xa = min(max(b.x0, 0), W - 1), xb = min(max(b.x0 + 1, 0), W - 1);
    const int ya = min(max(b.y0, 0), H - 1), yb = min(max(b.y0 + 1, 0), H - 1);
    const int o00 = ya * W + xa, o01 = ya * W + xb, o10 = yb * W + xa, o11 = yb * W + xb;
    const float w00 = (y0ok && x0ok) ? (1.f - b.fy) * (1.f - b.fx) : 0.f, w01 = (y0ok && x1ok) ? (1.f - b.fy) * b.fx : 0.f;
    const float w10 = (y1ok && x0ok) ? b.fy * (1.f - b.fx) : 0.f, w11 = (y1ok && x1ok) ? b.fy * b.fx : 0.f;
    const float* img = x + (long)n * C * HW;
    float* out = cols + ((long)n * C * 9 + t) * HW + p;
#pragma unroll 8
    for (int c = 0; c < C; ++c, img += HW, out += 9L * HW)
      *out = w00 * __ldg(img + o00) + w01 * __ldg(img + o01) + w10 * __ldg(img + o10) + w11 * __ldg(img + o11);
  }
}

// ---- deterministic scatter (dbm_set_deterministic): exact 64-bit fixed-point accumulation --------------------------
// Floating-point atomicAdd makes the result depend on the order the hardware serves colliding updates. Integer
// addition is associative, so the scatter accumulates round(v * 2^s) into an int64 shadow of the target and converts
// once at the end: bit-identical from run to run. s is chosen from max|g| of the incoming gradient (one ordered
// max-reduction) so that every addend is below 2^46 -- 2^16 colliding addends fit before 2^62, and the quantum is
// 2^-46 of the largest gradient (fp32 keeps 2^-24).
__global__ void absmax_bits_kernel(const float* __restrict__ x, long n, unsigned int* __restrict__ out) {
  unsigned int m = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(x[i]) & 0x7FFFFFFFu);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);   // max is order-independent
}
__device__ __forceinline__ int det_scale_exp(unsigned int maxbits) {
  const int e = (int)(maxbits >> 23) - 126;   // |g| < 2^e
  const int s = 46 - e;
  return s > 100 ? 100 : (s < -100 ? -100 : s);
}
__device__ __forceinline__ void det_add(long long* shadow, long idx, float v, float scale) {
  atomicAdd(reinterpret_cast<unsigned long long*>(shadow + idx), (unsigned long long)__float2ll_rn(v * scale));
}
// target[i] = (accumulate ? target[i] : 0) + shadow[i] * 2^-s
__global__ void det_finalize_kernel(const long long* __restrict__ shadow, float* __restrict__ target, long n,
                                    const unsigned int* __restrict__ maxbits, int accumulate) {
  const float inv = exp2f((float)-det_scale_exp(*maxbits));
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = (float)((double)shadow[i] * (double)inv);
    target[i] = accumulate ? target[i] + v : v;
  }
}

// Given dcols[n][c*9+t][p]: dx[n][c] += scatter (atomics), doff[n][t|9+t][p] = d/d(position).
// DET: the scatter goes to the int64 shadow `sh` with scale 2^det_scale_exp(*maxbits) instead of dx.
template <bool DET>
__global__ void deform_bwd_kernel(const float* __restrict__ x, const float* __restrict__ off,
                                  const float* __restrict__ dcols, float* __restrict__ dx,
                                  float* __restrict__ doff, int N, int C, int H, int W, long long* __restrict__ sh,
                                  const unsigned int* __restrict__ maxbits) {
  const float dscale = DET ? exp2f((float)det_scale_exp(*maxbits)) : 0.f;
  const int HW = H * W;
  const long total = (long)N * 9 * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = i % HW;
    const long r = i / HW;
    const int t = r % 9;
    const int n = r / 9;
    const int y = p / W, xx = p - y * W;
    const Bilin b = tap_position(off, (long)n * 18 * HW, HW, p, t, y, xx, H, W);
    const float w00 = (1.f - b.fy) * (1.f - b.fx), w01 = (1.f - b.fy) * b.fx;
    const float w10 = b.fy * (1.f - b.fx), w11 = b.fy * b.fx;
    const bool v00 = b.y0 >= 0 && b.y0 < H && b.x0 >= 0 && b.x0 < W;
    const bool v01 = b.y0 >= 0 && b.y0 < H && b.x0 + 1 >= 0 && b.x0 + 1 < W;
    const bool v10 = b.y0 + 1 >= 0 && b.y0 + 1 < H && b.x0 >= 0 && b.x0 < W;
    const bool v11 = b.y0 + 1 >= 0 && b.y0 + 1 < H && b.x0 + 1 >= 0 && b.x0 + 1 < W;
    float gpx = 0.f, gpy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float g = dcols[((long)n * C * 9 + (long)c * 9 + t) * HW + p];
      const long plane = ((long)n * C + c) * HW;
      const float* img = x + plane;
      const float a00 = v00 ? __ldg(img + (long)b.y0 * W + b.x0) : 0.f;
      const float a01 = v01 ? __ldg(img + (long)b.y0 * W + b.x0 + 1) : 0.f;
      const float a10 = v10 ? __ldg(img + (long)(b.y0 + 1) * W + b.x0) : 0.f;
      const float a11 = v11 ? __ldg(img + (long)(b.y0 + 1) * W + b.x0 + 1) : 0.f;
      gpx += g * ((1.f - b.fy) * (a01 - a00) + b.fy * (a11 - a10));
      gpy += g * ((1.f - b.fx) * (a10 - a00) + b.fx * (a11 - a01));
      if (dx) {
        if (DET) {
          if (v00) det_add(sh, plane + (long)b.y0 * W + b.x0, g * w00, dscale);
          if (v01) det_add(sh, plane + (long)b.y0 * W + b.x0 + 1, g * w01, dscale);
          if (v10) det_add(sh, plane + (long)(b.y0 + 1) * W + b.x0, g * w10, dscale);
          if (v11) det_add(sh, plane + (long)(b.y0 + 1) * W + b.x0 + 1, g * w11, dscale);
        } else {
          float* d = dx + plane;
          if (v00) atomicAdd(d + (long)b.y0 * W + b.x0, g * w00);
          if (v01) atomicAdd(d + (long)b.y0 * W + b.x0 + 1, g * w01);
          if (v10) atomicAdd(d + (long)(b.y0 + 1) * W + b.x0, g * w10);
          if (v11) atomicAdd(d + (long)(b.y0 + 1) * W + b.x0 + 1, g * w11);
        }
      }
    }
    if (!b.in_range) { gpx = 0.f; gpy = 0.f; }
    doff[(long)n * 18 * HW + (long)t * HW + p] = gpx;
    doff[(long)n * 18 * HW + (long)(9 + t) * HW + p] = gpy;
  }
}


// ---- single-output deformable layer (final_conv_layer2, 64 -> 1; srgan_train.py:515-523, 574) -------------
// With one output channel the contraction commutes with the bilinear sampler:
//   y[p] = b + sum_t bilin(z_t, pos_t(p)),   z_t[q] = sum_c W[0,c,t] x[c][q]   ("tap projection"),
// so the layer samples 9 projected planes instead of 64 x 9 input planes (64x fewer gathers, no cols buffer),
// and its backward is the transpose: scatter dy into dz_t, then dx[c] = sum_t W[c,t] dz_t and
// dW[c,t] = sum_q x[c][q] dz_t[q]. Only fp32 sums are re-associated.
__global__ void __launch_bounds__(256) deform1_project_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              float* __restrict__ proj, int N, int C, int HW) {
  extern __shared__ float sw[];  // [C][9]
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long total = (long)N * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / HW, px = i - n * HW;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    const float* xp = x + n * C * HW + px;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(xp + (long)c * HW);
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[t] = fmaf(v, sw[c * 9 + t], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) proj[(n * 9 + t) * HW + px] = acc[t];
  }
}

__global__ void __launch_bounds__(256) deform1_sample_kernel(const float* __restrict__ proj, const float* __restrict__ off,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             int N, int H, int W) {
  const int HW = H * W;
  const long total = (long)N * HW;
  const float b0 = bias ? bias[0] : 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = i / HW, p = i - (long)n * HW;
    const int yy = p / W, xx = p - yy * W;
    float acc = b0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const Bilin b = tap_position(off, (long)n * 18 * HW, HW, p, t, yy, xx, H, W);
      const float* img = proj + ((long)n * 9 + t) * HW;
      acc += (1.f - b.fy) * ((1.f - b.fx) * at(img, b.y0, b.x0, H, W) + b.fx * at(img, b.y0, b.x0 + 1, H, W)) +
             b.fy * ((1.f - b.fx) * at(img, b.y0 + 1, b.x0, H, W) + b.fx * at(img, b.y0 + 1, b.x0 + 1, H, W));
    }
    y[i] = acc;
  }
}

// dproj[n][t] += scatter of dy (atomics; dproj zeroed by the caller); doff[n][t | 9+t][p] = d/d(position)
template <bool DET>
__global__ void __launch_bounds__(256) deform1_bwd_scatter_kernel(const float* __restrict__ proj,
                                                                  const float* __restrict__ off,
                                                                  const float* __restrict__ dy, float* __restrict__ dproj,
                                                                  float* __restrict__ doff, int N, int H, int W,
                                                                  long long* __restrict__ sh,
                                                                  const unsigned int* __restrict__ maxbits) {
  const float dscale = DET ? exp2f((float)det_scale_exp(*maxbits)) : 0.f;
  const int HW = H * W;
  const long total = (long)N * 9 * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = i % HW;
    const long r = i / HW;
    const int t = r % 9;
    const int n = r / 9;
    const int y = p / W, xx = p - y * W;
    const Bilin b = tap_position(off, (long)n * 18 * HW, HW, p, t, y, xx, H, W);
    const float g = dy[(long)n * HW + p];
    const long plane = ((long)n * 9 + t) * HW;
    const float* img = proj + plane;
    const bool v00 = b.y0 >= 0 && b.y0 < H && b.x0 >= 0 && b.x0 < W;
    const bool v01 = b.y0 >= 0 && b.y0 < H && b.x0 + 1 >= 0 && b.x0 + 1 < W;
    const bool v10 = b.y0 + 1 >= 0 && b.y0 + 1 < H && b.x0 >= 0 && b.x0 < W;
    const bool v11 = b.y0 + 1 >= 0 && b.y0 + 1 < H && b.x0 + 1 >= 0 && b.x0 + 1 < W;
    const float a00 = v00 ? __ldg(img + (long)b.y0 * W + b.x0) : 0.f;
    const float a01 = v01 ? __ldg(img + (long)b.y0 * W + b.x0 + 1) : 0.f;
    const float a10 = v10 ? __ldg(img + (long)(b.y0 + 1) * W + b.x0) : 0.f;
    const float a11 = v11 ? __ldg(img + (long)(b.y0 + 1) * W + b.x0 + 1) : 0.f;
    float gpx = g * ((1.f - b.fy) * (a01 - a00) + b.fy * (a11 - a10));
    float gpy = g * ((1.f - b.fx) * (a10 - a00) + b.fx * (a11 - a01));
    if (!b.in_range) { gpx = 0.f; gpy = 0.f; }
    doff[(long)n * 18 * HW + (long)t * HW + p] = gpx;
    doff[(long)n * 18 * HW + (long)(9 + t) * HW + p] = gpy;
    if (DET) {
      if (v00) det_add(sh, plane + (long)b.y0 * W + b.x0, g * (1.f - b.fy) * (1.f - b.fx), dscale);
      if (v01) det_add(sh, plane + (long)b.y0 * W + b.x0 + 1, g * (1.f - b.fy) * b.fx, dscale);
      if (v10) det_add(sh, plane + (long)(b.y0 + 1) * W + b.x0, g * b.fy * (1.f - b.fx), dscale);
      if (v11) det_add(sh, plane + (long)(b.y0 + 1) * W + b.x0 + 1, g * b.fy * b.fx, dscale);
    } else {
      float* d = dproj + plane;
      if (v00) atomicAdd(d + (long)b.y0 * W + b.x0, g * (1.f - b.fy) * (1.f - b.fx));
      if (v01) atomicAdd(d + (long)b.y0 * W + b.x0 + 1, g * (1.f - b.fy) * b.fx);
      if (v10) atomicAdd(d + (long)(b.y0 + 1) * W + b.x0, g * b.fy * (1.f - b.fx));
      if (v11) atomicAdd(d + (long)(b.y0 + 1) * W + b.x0 + 1, g * b.fy * b.fx);
    }
  }
}

// dx[n][c][q] (+)= sum_t W[c,t] dproj[n][t][q]
__global__ void __launch_bounds__(256) deform1_bwd_data_kernel(const float* __restrict__ dproj, const float* __restrict__ w,
                                                               float* __restrict__ dx, int N, int C, int HW,
                                                               int accumulate) {
  extern __shared__ float sw[];  // [C][9]
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long total = (long)N * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / HW, px = i - n * HW;
    float dz[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) dz[t] = __ldg(dproj + (n * 9 + t) * HW + px);
    float* dp = dx + n * C * HW + px;
    for (int c = 0; c < C; ++c) {
      float v = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) v = fmaf(dz[t], sw[c * 9 + t], v);
      if (accumulate) v += dp[(long)c * HW];
      dp[(long)c * HW] = v;
    }
  }
}

// dW[c][t] += sum_{n,q} x[n][c][q] dproj[n][t][q]: grid (C, image chunks), one atomicAdd per block and tap
__global__ void __launch_bounds__(256) deform1_bwd_weight_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ dproj, float* __restrict__ dw,
                                                                 int N, int C, int HW) {
  const int c = blockIdx.x;
  const int per = (N + gridDim.y - 1) / gridDim.y;
  const int n0 = blockIdx.y * per, n1 = min(N, n0 + per);
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float* xp = x + ((long)n * C + c) * HW;
    const float* dz = dproj + (long)n * 9 * HW;
    for (int q = threadIdx.x; q < HW; q += blockDim.x) {
      const float v = __ldg(xp + q);
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[t] = fmaf(v, __ldg(dz + (long)t * HW + q), acc[t]);
    }
  }
  __shared__ float red[9][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float v = acc[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[t][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9 && n0 < n1) {
    float v = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += red[threadIdx.x][k];
    atomicAdd(dw + c * 9 + threadIdx.x, v);
  }
}

}  // namespace dbm

using namespace dbm;

extern "C" int dbm_deform_sample_f32(const float* x, const float* offset, float* cols, int n, int c, int h, int w,
                                     cudaStream_t st) {
  DBM_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "deform_sample: empty input");
  const long total = (long)n * 9 * h * w;
  long blocks = (total + 255) / 256;
  if (blocks > (long)num_sms() * 32) blocks = (long)num_sms() * 32;
  deform_sample_kernel<<<(int)blocks, 256, 0, st>>>(x, offset, cols, n, c, h, w);
  return check_launch("deform_sample");
}

extern "C" int dbm_deform_bwd_f32(const float* x, const float* offset, const float* dcols, float* dx, float* doffset,
                                  int n, int c, int h, int w, cudaStream_t st) {
  DBM_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "deform_bwd: empty input");
  const long total = (long)n * 9 * h * w;
  long blocks = (total + 255) / 256;
  if (blocks > (long)num_sms() * 32) blocks = (long)num_sms() * 32;
  if (deterministic() && dx != nullptr) {
    const long nt = (long)n * c * h * w, ng = (long)n * c * 9 * h * w;
    void* scratch;
    int rc = det_scratch((size_t)nt * 8 + 256, &scratch);
    if (rc) return rc;
    unsigned int* maxbits = (unsigned int*)scratch;
    long long* shadow = (long long*)((char*)scratch + 256);
    DBM_CUDA(cudaMemsetAsync(scratch, 0, (size_t)nt * 8 + 256, st));
    absmax_bits_kernel<<<num_sms() * 8, 256, 0, st>>>(dcols, ng, maxbits);
    deform_bwd_kernel<true><<<(int)blocks, 256, 0, st>>>(x, offset, dcols, dx, doffset, n, c, h, w, shadow, maxbits);
    det_finalize_kernel<<<num_sms() * 8, 256, 0, st>>>(shadow, dx, nt, maxbits, 1);
    return check_launch("deform_bwd (deterministic)");
  }
  deform_bwd_kernel<false><<<(int)blocks, 256, 0, st>>>(x, offset, dcols, dx, doffset, n, c, h, w, nullptr, nullptr);
  return check_launch("deform_bwd");
}

static inline int deform1_blocks(long total) {
  long blocks = (total + 255) / 256;
  if (blocks > (long)num_sms() * 16) blocks = (long)num_sms() * 16;
  return (int)(blocks < 1 ? 1 : blocks);
}

extern "C" int dbm_deform1_fwd_f32(const float* x, const float* offset, const float* w, const float* bias, float* y,
                                   float* proj, int n, int c, int h, int wd, cudaStream_t st) {
  DBM_REQUIRE(n > 0 && c > 0 && h > 0 && wd > 0, "deform1_fwd: empty input");
  DBM_REQUIRE(c * 9 * sizeof(float) <= 40 * 1024, "deform1_fwd: too many input channels (%d)", c);
  const long total = (long)n * h * wd;
  deform1_project_kernel<<<deform1_blocks(total), 256, c * 9 * sizeof(float), st>>>(x, w, proj, n, c, h * wd);
  int rc = check_launch("deform1_project");
  if (rc) return rc;
  deform1_sample_kernel<<<deform1_blocks(total), 256, 0, st>>>(proj, offset, bias, y, n, h, wd);
  return check_launch("deform1_sample");
}

extern "C" int dbm_deform1_bwd_f32(const float* x, const float* offset, const float* w, const float* proj,
                                   const float* dy, float* dw, float* dx, int accumulate_dx, float* doffset,
                                   float* dproj_scratch, int n, int c, int h, int wd, cudaStream_t st) {
  DBM_REQUIRE(n > 0 && c > 0 && h > 0 && wd > 0, "deform1_bwd: empty input");
  DBM_REQUIRE(c * 9 * sizeof(float) <= 40 * 1024, "deform1_bwd: too many input channels (%d)", c);
  const int hw = h * wd;
  cudaError_t e = cudaMemsetAsync(dproj_scratch, 0, (size_t)n * 9 * hw * sizeof(float), st);
  DBM_REQUIRE(e == cudaSuccess, "deform1_bwd: memset failed: %s", cudaGetErrorString(e));
  int rc;
  if (deterministic()) {
    const long nt = (long)n * 9 * hw;
    void* scratch;
    rc = det_scratch((size_t)nt * 8 + 256, &scratch);
    if (rc) return rc;
    unsigned int* maxbits = (unsigned int*)scratch;
    long long* shadow = (long long*)((char*)scratch + 256);
    DBM_CUDA(cudaMemsetAsync(scratch, 0, (size_t)nt * 8 + 256, st));
    absmax_bits_kernel<<<num_sms() * 4, 256, 0, st>>>(dy, (long)n * hw, maxbits);
    deform1_bwd_scatter_kernel<true><<<deform1_blocks(nt), 256, 0, st>>>(proj, offset, dy, dproj_scratch, doffset, n, h, wd,
                                                                        shadow, maxbits);
    det_finalize_kernel<<<num_sms() * 4, 256, 0, st>>>(shadow, dproj_scratch, nt, maxbits, 0);
  } else {
    deform1_bwd_scatter_kernel<false><<<deform1_blocks((long)n * 9 * hw), 256, 0, st>>>(proj, offset, dy, dproj_scratch,
                                                                                       doffset, n, h, wd, nullptr, nullptr);
  }
  rc = check_launch("deform1_bwd_scatter");
  if (rc) return rc;
  if (dx) {
    deform1_bwd_data_kernel<<<deform1_blocks((long)n * hw), 256, c * 9 * sizeof(float), st>>>(dproj_scratch, w, dx, n, c,
                                                                                             hw, accumulate_dx);
    rc = check_launch("deform1_bwd_data");
    if (rc) return rc;
  }
  if (dw) {
    int chunks = (4 * num_sms() + c - 1) / c;
    if (chunks > n) chunks = n;
    if (deterministic()) chunks = 1;   // one block per channel: a single contributor per dW element
    deform1_bwd_weight_kernel<<<dim3(c, chunks), 256, 0, st>>>(x, dproj_scratch, dw, n, c, hw);
    rc = check_launch("deform1_bwd_weight");
  }
  return rc;
}
