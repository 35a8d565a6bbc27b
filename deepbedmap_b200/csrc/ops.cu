// Element-wise / layout kernels of the generator and discriminator graphs (HBM-bound, vectorised
// where the layout allows). Each entry cites the reference op it replaces.
#include <stdarg.h>

#include "common.cuh"

namespace dbm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static bool g_deterministic = false;
bool deterministic() { return g_deterministic; }

int det_scratch(size_t bytes, void** out) {
  static void* buf = nullptr;
  static size_t cap = 0;
  if (bytes > cap) {
    if (buf) DBM_CUDA(cudaFree(buf));
    buf = nullptr; cap = 0;
    DBM_CUDA(cudaMalloc(&buf, bytes));
    cap = bytes;
  }
  *out = buf;
  return DBM_OK;
}

static long g_launches = 0;   // kernels launched by this library since it was loaded (every launch ends in check_launch)

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return DBM_ERR_CUDA;
  }
  return DBM_OK;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int resident_ctas(const void* kernel, int threads, size_t smem, int* resident) {
  struct Entry { int dev; const void* k; size_t smem; int threads; int value; };
  static Entry cache[64];
  static int used = 0;
  int dev = 0;
  DBM_CUDA(cudaGetDevice(&dev));
  for (int i = 0; i < used; ++i)
    if (cache[i].dev == dev && cache[i].k == kernel && cache[i].smem == smem && cache[i].threads == threads) {
      *resident = cache[i].value;
      return DBM_OK;
    }
  DBM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, sms = 0;
  DBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  DBM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DBM_REQUIRE(per_sm >= 1 && sms >= 1, "persistent kernel cannot be resident (%d CTAs/SM with %d threads, %zu B smem)",
              per_sm, threads, smem);
  *resident = per_sm * sms;
  if (used < 64) cache[used++] = Entry{dev, kernel, smem, threads, *resident};
  return DBM_OK;
}

int ensure_dyn_smem(const void* kernel, size_t smem) {
  struct Entry { int dev; const void* k; size_t smem; };
  static Entry cache[128];
  static int used = 0;
  int dev = 0;
  DBM_CUDA(cudaGetDevice(&dev));
  for (int i = 0; i < used; ++i)
    if (cache[i].dev == dev && cache[i].k == kernel && cache[i].smem >= smem) return DBM_OK;
  DBM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (used < 128) cache[used++] = Entry{dev, kernel, smem};
  return DBM_OK;
}

static inline int ew_grid(long total, int per_block = 256) {
  long b = (total + per_block - 1) / per_block;
  long cap = (long)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---- layout conversion: NCHW fp32 <-> slab8 bf16 / slab4 fp32 -------------------------------
// One thread per (n, slab, pixel): reads V channel planes (coalesced along pixels), writes one
// 16-byte vector.
__global__ void nchw_to_slab8_kernel(const float* __restrict__ src, long src_bs, __nv_bfloat16* __restrict__ dst,
                                     int N, int C, int HW, int dst_cs_total, int dst_cs0) {
  const long total = (long)N * (C / 8) * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = i % HW;
    const long t = i / HW;
    const int cs = t % (C / 8);
    const int n = t / (C / 8);
    const float* s = src + n * src_bs + (long)cs * 8 * HW + px;
    __nv_bfloat162 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __floats2bfloat162_rn(s[(2 * k) * (long)HW], s[(2 * k + 1) * (long)HW]);
    *reinterpret_cast<uint4*>(dst + (((long)n * dst_cs_total + dst_cs0 + cs) * HW + px) * 8) =
        *reinterpret_cast<uint4*>(v);
  }
}
__global__ void slab8_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int src_cs_total, int src_cs0,
                                     float* __restrict__ dst, long dst_bs, int N, int C, int HW) {
  const long total = (long)N * (C / 8) * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = i % HW;
    const long t = i / HW;
    const int cs = t % (C / 8);
    const int n = t / (C / 8);
    uint4 raw = *reinterpret_cast<const uint4*>(src + (((long)n * src_cs_total + src_cs0 + cs) * HW + px) * 8);
    const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(&raw);
    float* d = dst + n * dst_bs + (long)cs * 8 * HW + px;
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k * (long)HW] = __bfloat162float(v[k]);
  }
}
__global__ void nchw_to_slab4_kernel(const float* __restrict__ src, long src_bs, float* __restrict__ dst, int N,
                                     int C, int HW) {
  const long total = (long)N * (C / 4) * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = i % HW;
    const long t = i / HW;
    const int cs = t % (C / 4);
    const int n = t / (C / 4);
    const float* s = src + n * src_bs + (long)cs * 4 * HW + px;
    *reinterpret_cast<float4*>(dst + i * 4) = make_float4(s[0], s[HW], s[2 * (long)HW], s[3 * (long)HW]);
  }
}
__global__ void slab4_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, long dst_bs, int N,
                                     int C, int HW, int c_keep) {
  const long total = (long)N * (C / 4) * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = i % HW;
    const long t = i / HW;
    const int cs = t % (C / 4);
    const int n = t / (C / 4);
    const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
    float* d = dst + n * dst_bs + (long)cs * 4 * HW + px;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (cs * 4 + k < c_keep) d[k * (long)HW] = vv[k];
  }
}

// fp32 NCHW -> split-bf16 slab8 (dbm_trunk_umma_split): per 16 channels the slabs [hi a | hi b | lo a | lo b],
// v = hi + lo, hi = bf16(v), lo = bf16(v - hi)
__global__ void nchw_to_slab8_split_kernel(const float* __restrict__ src, long src_bs, __nv_bfloat16* __restrict__ dst,
                                           int N, int C, int HW) {
  const long total = (long)N * (C / 8) * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = i % HW;
    const long t = i / HW;
    const int cs = t % (C / 8);
    const int n = t / (C / 8);
    const float* s = src + n * src_bs + (long)cs * 8 * HW + px;
    __nv_bfloat162 hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = s[(2 * k) * (long)HW], b = s[(2 * k + 1) * (long)HW];
      hi[k] = __floats2bfloat162_rn(a, b);
      const float2 f = __bfloat1622float2(hi[k]);
      lo[k] = __floats2bfloat162_rn(a - f.x, b - f.y);
    }
    const long ps = ((cs >> 1) << 2) + (cs & 1);   // physical slab of the hi part
    *reinterpret_cast<uint4*>(dst + (((long)n * (C / 4) + ps) * HW + px) * 8) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + (((long)n * (C / 4) + ps + 2) * HW + px) * 8) = *reinterpret_cast<uint4*>(lo);
  }
}
// fp32 slab8f [N][C/8][HW][8] (the trunk kernels' fp32 outputs) -> NCHW
__global__ void slab8f_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, long dst_bs, int N, int C,
                                      int HW) {
  const long total = (long)N * (C / 8) * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = i % HW;
    const long t = i / HW;
    const int cs = t % (C / 8);
    const int n = t / (C / 8);
    const float4 v0 = *reinterpret_cast<const float4*>(src + i * 8);
    const float4 v1 = *reinterpret_cast<const float4*>(src + i * 8 + 4);
    float* d = dst + n * dst_bs + (long)cs * 8 * HW + px;
    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k * (long)HW] = vv[k];
  }
}

// ---- strided element-wise ops on (batch, inner) views of NCHW tensors --------------------------
// out = a*x + b*y   (F.add(a5 * residual_scaling, a0), srgan_train.py:358, 402, 551)
__global__ void axpby_kernel(const float* __restrict__ x, long x_bs, const float* __restrict__ y, long y_bs,
                             float* __restrict__ out, long o_bs, float a, float b, int nb, long inner) {
  const long total = (long)nb * inner;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / inner, r = i - n * inner;
    float v = a * x[n * x_bs + r];
    if (y) v += b * y[n * y_bs + r];
    out[n * o_bs + r] = v;
  }
}
// dx (+)= dy * (y >= 0 ? 1 : slope)   backward of F.leaky_relu given its OUTPUT y
__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, long dy_bs, const float* __restrict__ y, long y_bs,
                                 float* __restrict__ dx, long dx_bs, int nb, long inner, int accumulate) {
  const long total = (long)nb * inner;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / inner, r = i - n * inner;
    float g = dy[n * dy_bs + r];
    if (y[n * y_bs + r] < 0.f) g *= kLreluSlope;
    if (accumulate) g += dx[n * dx_bs + r];
    dx[n * dx_bs + r] = g;
  }
}
__global__ void lrelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
    y[i] = lrelu(x[i]);
}

// F.resize_images(mode="nearest") to exactly 2x and its adjoint (srgan_train.py:556-566)
__global__ void upsample2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long planes, int H, int W) {
  const int Wo = 2 * W, Ho = 2 * H;
  const long total = planes * Ho * Wo;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xo = i % Wo;
    const long t = i / Wo;
    const int yo = t % Ho;
    const long pl = t / Ho;
    y[i] = x[(pl * H + (yo >> 1)) * W + (xo >> 1)];
  }
}
__global__ void upsample2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long planes, int H,
                                     int W) {
  const int Wo = 2 * W;
  const long total = planes * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xi = i % W;
    const long t = i / W;
    const int yi = t % H;
    const long pl = t / H;
    const float* s = dy + (pl * 2 * H + 2 * yi) * Wo + 2 * xi;
    dx[i] = s[0] + s[1] + s[Wo] + s[Wo + 1];
  }
}

// db[o] += sum_{n,pixels} dy[n, o, :]  (bias gradient of L.Convolution2D)
__global__ void bias_grad_kernel(const float* __restrict__ dy, long dy_bs, float* __restrict__ db, int N, int HW) {
  // grid = (channels, image chunks): each block reduces its images' planes, one atomicAdd per block
  const int o = blockIdx.x;
  const int per = (N + gridDim.y - 1) / gridDim.y;
  const int n0 = blockIdx.y * per, n1 = min(N, n0 + per);
  float s = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float* __restrict__ src = dy + (long)n * dy_bs + (long)o * HW;
    for (int r = threadIdx.x; r < HW; r += blockDim.x) s += src[r];
  }
  __shared__ float red[32];
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (threadIdx.x == 0 && n0 < n1) atomicAdd(db + o, s);
  }
}

// Crop + clip + cast of one tile from the device-resident continent grid
// (deepbedmap.py:663-665 clip >= 0, :715-722 crops).  src is (C, Hs, Ws) fp32.
__global__ void crop_clip_kernel(const float* __restrict__ src, int Hs, int Ws, float* __restrict__ dst, int C,
                                 int y0, int x0, int h, int w, int clip0) {
  const long total = (long)C * h * w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xx = i % w;
    const long t = i / w;
    const int yy = t % h;
    const int c = t / h;
    float v = src[((long)c * Hs + (y0 + yy)) * Ws + (x0 + xx)];
    if (clip0) v = fmaxf(v, 0.f);
    dst[i] = v;
  }
}

// Y_hat[ys:ys+hh, xs:xs+ww] = y_pred[cy:cy+hh, cx:cx+ww]   (deepbedmap.py:731-736)
__global__ void place_tile_kernel(const float* __restrict__ tile, int th, int tw, int cy, int cx,
                                  float* __restrict__ canvas, int CH, int CW, int ys, int xs, int hh, int ww) {
  const long total = (long)hh * ww;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xx = i % ww, yy = i / ww;
    canvas[(long)(ys + yy) * CW + (xs + xx)] = tile[(long)(cy + yy) * tw + (cx + xx)];
  }
}
__global__ void fill_kernel(float* __restrict__ p, float v, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}
// Y_hat.astype(np.int16) (deepbedmap.py:751) as NumPy's C cast does it on x86-64: truncate toward zero to a
// 32-bit integer (NaN / |v| >= 2^31 give INT_MIN) and keep the low 16 bits. Four values per thread.
__device__ __forceinline__ short f32_to_i16_numpy(float v) {
  const int i = (v != v || v >= 2147483648.0f || v < -2147483648.0f) ? (int)0x80000000 : (int)v;
  return (short)(i & 0xffff);
}
__global__ void f32_to_i16_kernel(const float* __restrict__ src, short* __restrict__ dst, long n) {
  const long n4 = n >> 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    short4 o;
    o.x = f32_to_i16_numpy(v.x), o.y = f32_to_i16_numpy(v.y), o.z = f32_to_i16_numpy(v.z), o.w = f32_to_i16_numpy(v.w);
    reinterpret_cast<short4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = f32_to_i16_numpy(src[(n4 << 2) + threadIdx.x]);
}
// batch[j] = dataset[index[j]] for rows of `row` floats (on-device minibatch gather, srgan_train.py:132-166)
__global__ void gather_rows_kernel(const float* __restrict__ src, const long* __restrict__ index, float* __restrict__ dst,
                                   long row, int nrows, long src_rows) {
  for (int j = blockIdx.y; j < nrows; j += gridDim.y) {
    const long r = index[j];
    if (r < 0 || r >= src_rows) continue;  // validated on the host; never write from a bad row
    const float* s = src + r * row;
    float* d = dst + (long)j * row;
    if ((row & 3) == 0) {
      for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (row >> 2); i += (long)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    } else {
      for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < row; i += (long)gridDim.x * blockDim.x) d[i] = s[i];
    }
  }
}

}  // namespace dbm

using namespace dbm;

extern "C" const char* dbm_last_error(void) { return g_err; }
extern "C" int dbm_version(void) { return 200; }
extern "C" long dbm_launch_count(void) { return g_launches; }
extern "C" int dbm_set_deterministic(int on) {
  g_deterministic = on != 0;
  return DBM_OK;
}

extern "C" int dbm_nchw_to_slab8(const float* src, long src_batch_stride, void* dst, int n, int c, int h, int w,
                                 int dst_cs_total, int dst_cs0, cudaStream_t st) {
  DBM_REQUIRE(c % 8 == 0, "nchw_to_slab8: C=%d not a multiple of 8", c);
  const long total = (long)n * (c / 8) * h * w;
  nchw_to_slab8_kernel<<<ew_grid(total), 256, 0, st>>>(src, src_batch_stride ? src_batch_stride : (long)c * h * w,
                                                       (__nv_bfloat16*)dst, n, c, h * w, dst_cs_total, dst_cs0);
  return check_launch("nchw_to_slab8");
}
extern "C" int dbm_slab8_to_nchw(const void* src, int src_cs_total, int src_cs0, float* dst, long dst_batch_stride,
                                 int n, int c, int h, int w, cudaStream_t st) {
  DBM_REQUIRE(c % 8 == 0, "slab8_to_nchw: C=%d not a multiple of 8", c);
  const long total = (long)n * (c / 8) * h * w;
  slab8_to_nchw_kernel<<<ew_grid(total), 256, 0, st>>>((const __nv_bfloat16*)src, src_cs_total, src_cs0, dst,
                                                       dst_batch_stride ? dst_batch_stride : (long)c * h * w, n, c,
                                                       h * w);
  return check_launch("slab8_to_nchw");
}
extern "C" int dbm_nchw_to_slab8_split(const float* src, long src_batch_stride, void* dst, int n, int c, int h, int w,
                                       cudaStream_t st) {
  DBM_REQUIRE(c % 16 == 0, "nchw_to_slab8_split: C=%d not a multiple of 16", c);
  const long total = (long)n * (c / 8) * h * w;
  nchw_to_slab8_split_kernel<<<ew_grid(total), 256, 0, st>>>(
      src, src_batch_stride ? src_batch_stride : (long)c * h * w, (__nv_bfloat16*)dst, n, c, h * w);
  return check_launch("nchw_to_slab8_split");
}
extern "C" int dbm_slab8f_to_nchw(const float* src, float* dst, long dst_batch_stride, int n, int c, int h, int w,
                                  cudaStream_t st) {
  DBM_REQUIRE(c % 8 == 0, "slab8f_to_nchw: C=%d not a multiple of 8", c);
  const long total = (long)n * (c / 8) * h * w;
  slab8f_to_nchw_kernel<<<ew_grid(total), 256, 0, st>>>(src, dst, dst_batch_stride ? dst_batch_stride : (long)c * h * w,
                                                        n, c, h * w);
  return check_launch("slab8f_to_nchw");
}
extern "C" int dbm_nchw_to_slab4(const float* src, long src_batch_stride, float* dst, int n, int c, int h, int w,
                                 cudaStream_t st) {
  DBM_REQUIRE(c % 4 == 0, "nchw_to_slab4: C=%d not a multiple of 4", c);
  const long total = (long)n * (c / 4) * h * w;
  nchw_to_slab4_kernel<<<ew_grid(total), 256, 0, st>>>(src, src_batch_stride ? src_batch_stride : (long)c * h * w,
                                                       dst, n, c, h * w);
  return check_launch("nchw_to_slab4");
}
extern "C" int dbm_slab4_to_nchw(const float* src, float* dst, long dst_batch_stride, int n, int c_slab, int c_keep,
                                 int h, int w, cudaStream_t st) {
  DBM_REQUIRE(c_slab % 4 == 0 && c_keep <= c_slab, "slab4_to_nchw: bad channels %d/%d", c_keep, c_slab);
  const long total = (long)n * (c_slab / 4) * h * w;
  slab4_to_nchw_kernel<<<ew_grid(total), 256, 0, st>>>(
      src, dst, dst_batch_stride ? dst_batch_stride : (long)c_keep * h * w, n, c_slab, h * w, c_keep);
  return check_launch("slab4_to_nchw");
}
extern "C" int dbm_axpby_f32(const float* x, long x_bs, const float* y, long y_bs, float* out, long out_bs, float a,
                             float b, int nbatch, long inner, cudaStream_t st) {
  axpby_kernel<<<ew_grid((long)nbatch * inner), 256, 0, st>>>(x, x_bs, y, y_bs, out, out_bs, a, b, nbatch, inner);
  return check_launch("axpby");
}
extern "C" int dbm_lrelu_bwd_f32(const float* dy, long dy_bs, const float* y, long y_bs, float* dx, long dx_bs,
                                 int nbatch, long inner, int accumulate, cudaStream_t st) {
  lrelu_bwd_kernel<<<ew_grid((long)nbatch * inner), 256, 0, st>>>(dy, dy_bs, y, y_bs, dx, dx_bs, nbatch, inner,
                                                                  accumulate);
  return check_launch("lrelu_bwd");
}
extern "C" int dbm_lrelu_fwd_f32(const float* x, float* y, long total, cudaStream_t st) {
  lrelu_fwd_kernel<<<ew_grid(total), 256, 0, st>>>(x, y, total);
  return check_launch("lrelu_fwd");
}
extern "C" int dbm_upsample2_fwd_f32(const float* x, float* y, long planes, int h, int w, cudaStream_t st) {
  upsample2_fwd_kernel<<<ew_grid(planes * 4 * h * w), 256, 0, st>>>(x, y, planes, h, w);
  return check_launch("upsample2_fwd");
}
extern "C" int dbm_upsample2_bwd_f32(const float* dy, float* dx, long planes, int h, int w, cudaStream_t st) {
  upsample2_bwd_kernel<<<ew_grid(planes * h * w), 256, 0, st>>>(dy, dx, planes, h, w);
  return check_launch("upsample2_bwd");
}
extern "C" int dbm_bias_grad_f32(const float* dy, long dy_bs, float* db, int n, int o, int hw, cudaStream_t st) {
  // small planes (9x9 training tiles): fewer threads per block, more image chunks
  const int threads = hw >= 1024 ? 256 : (hw >= 256 ? 128 : 64);
  int chunks = (8 * num_sms() + o - 1) / o;
  if (chunks > n) chunks = n;
  if (chunks < 1 || deterministic()) chunks = 1;   // one block (fixed reduction tree) per channel: a single contributor
  bias_grad_kernel<<<dim3(o, chunks), threads, 0, st>>>(dy, dy_bs ? dy_bs : (long)o * hw, db, n, hw);
  return check_launch("bias_grad");
}
extern "C" int dbm_crop_clip_f32(const float* src, int hs, int ws, float* dst, int c, int y0, int x0, int h, int w,
                                 int clip0, cudaStream_t st) {
  DBM_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + h <= hs && x0 + w <= ws, "crop out of bounds");
  crop_clip_kernel<<<ew_grid((long)c * h * w), 256, 0, st>>>(src, hs, ws, dst, c, y0, x0, h, w, clip0);
  return check_launch("crop_clip");
}
extern "C" int dbm_place_tile_f32(const float* tile, int th, int tw, int cy, int cx, float* canvas, int ch, int cw,
                                  int ys, int xs, int hh, int ww, cudaStream_t st) {
  DBM_REQUIRE(cy >= 0 && cx >= 0 && cy + hh <= th && cx + ww <= tw, "place_tile: source window out of bounds");
  DBM_REQUIRE(ys >= 0 && xs >= 0 && ys + hh <= ch && xs + ww <= cw, "place_tile: canvas window out of bounds");
  if (hh == 0 || ww == 0) return DBM_OK;
  place_tile_kernel<<<ew_grid((long)hh * ww), 256, 0, st>>>(tile, th, tw, cy, cx, canvas, ch, cw, ys, xs, hh, ww);
  return check_launch("place_tile");
}
extern "C" int dbm_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                                size_t rows, cudaStream_t st) {
  if (width_bytes == 0 || rows == 0) return DBM_OK;
  DBM_REQUIRE(dst != nullptr && src != nullptr, "copy2d: null pointer");
  DBM_REQUIRE(dst_pitch >= width_bytes && src_pitch >= width_bytes, "copy2d: pitch smaller than the row width");
  if (dst_pitch == width_bytes && src_pitch == width_bytes) {
    DBM_CUDA(cudaMemcpyAsync(dst, src, width_bytes * rows, cudaMemcpyDefault, st));
  } else {
    DBM_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDefault, st));
  }
  return DBM_OK;
}

extern "C" int dbm_fill_f32(float* p, float v, long n, cudaStream_t st) {
  if (n <= 0) return DBM_OK;
  fill_kernel<<<ew_grid(n), 256, 0, st>>>(p, v, n);
  return check_launch("fill");
}
extern "C" int dbm_f32_to_i16(const float* src, void* dst, long n, cudaStream_t st) {
  if (n <= 0) return DBM_OK;
  DBM_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "f32_to_i16: buffers must be 16/8-byte aligned");
  f32_to_i16_kernel<<<ew_grid((n + 3) / 4), 256, 0, st>>>(src, (short*)dst, n);
  return check_launch("f32_to_i16");
}
extern "C" int dbm_gather_rows_f32(const float* src, long src_rows, const long* index_dev, float* dst, long row,
                                   int nrows, cudaStream_t st) {
  if (nrows <= 0 || row <= 0) return DBM_OK;
  DBM_REQUIRE(src_rows > 0, "gather_rows: empty dataset");
  DBM_REQUIRE((row & 3) != 0 || ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0), "gather_rows: buffers must be 16-byte aligned");
  const int bx = (int)((((row & 3) ? row : (row >> 2)) + 255) / 256);
  gather_rows_kernel<<<dim3(bx < 64 ? (bx < 1 ? 1 : bx) : 64, nrows < 65535 ? nrows : 65535), 256, 0, st>>>(
      src, index_dev, dst, row, nrows, src_rows);
  return check_launch("gather_rows");
}
