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
    const float w00 = (1.f - b.fy) * (1.f - b.fx), w01 = (1.f - b.fy) * b.fx;
    const float w10 = b.fy * (1.f - b.fx), w11 = b.fy * b.fx;
    for (int c = 0; c < C; ++c) {
      const float* img = x + ((long)n * C + c) * HW;
      const float v = w00 * at(img, b.y0, b.x0, H, W) + w01 * at(img, b.y0, b.x0 + 1, H, W) +
                      w10 * at(img, b.y0 + 1, b.x0, H, W) + w11 * at(img, b.y0 + 1, b.x0 + 1, H, W);
      cols[((long)n * C * 9 + (long)c * 9 + t) * HW + p] = v;
    }
  }
}

// Given dcols[n][c*9+t][p]: dx[n][c] += scatter (atomics), doff[n][t|9+t][p] = d/d(position).
__global__ void deform_bwd_kernel(const float* __restrict__ x, const float* __restrict__ off,
                                  const float* __restrict__ dcols, float* __restrict__ dx,
                                  float* __restrict__ doff, int N, int C, int H, int W) {
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
        float* d = dx + plane;
        if (v00) atomicAdd(d + (long)b.y0 * W + b.x0, g * w00);
        if (v01) atomicAdd(d + (long)b.y0 * W + b.x0 + 1, g * w01);
        if (v10) atomicAdd(d + (long)(b.y0 + 1) * W + b.x0, g * w10);
        if (v11) atomicAdd(d + (long)(b.y0 + 1) * W + b.x0 + 1, g * w11);
      }
    }
    if (!b.in_range) { gpx = 0.f; gpy = 0.f; }
    doff[(long)n * 18 * HW + (long)t * HW + p] = gpx;
    doff[(long)n * 18 * HW + (long)(9 + t) * HW + p] = gpy;
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
  deform_bwd_kernel<<<(int)blocks, 256, 0, st>>>(x, offset, dcols, dx, doffset, n, c, h, w);
  return check_launch("deform_bwd");
}
