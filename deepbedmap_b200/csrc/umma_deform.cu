// Deformable 3x3 convolution for the tensor-core inference path
// (L.DeformableConvolution2D, srgan_train.py:506-523, 572-574; semantics SURVEY App. B.6).
//
//   deform_umma_kernel  (64 -> 64): gather warps bilinearly sample the bf16 slab8 input at the
//       offset-displaced tap positions straight into shared memory in the UMMA K-major
//       core-matrix layout (no im2col buffer in HBM: Chainer materialises 3 GB per continent
//       tile here); one thread issues tcgen05.mma against the SMEM-resident 576x64 filter;
//       epilogue warps add bias, LeakyReLU and store bf16 slab8.
//   deform_out1_kernel  (64 -> 1): the final layer is a 576-long dot product per pixel: CUDA
//       cores, fp32 accumulation, fp32 NCHW output.
#include "common.cuh"

namespace dbm {

// A work item is 32 x 8 = 256 output pixels = two M=128 MMA tiles (rows 0-3 / 4-7 of the item).
// 24 gather warps (6 per SM sub-partition; the 12 of the 128-pixel version issued on 53 % of the
// cycles, ncu r1g) : thread = (pixel, tap slot), all 64 channels of one tap per thread.
constexpr int kDTileW = 32, kDTileH = 8;
constexpr int kDPix = kDTileW * kDTileH;   // 256
constexpr int kDSlots = 3;                 // tap-slot s gathers taps s, s+3, s+6 of every item ...
constexpr int kDBuf = 1;                   // ... into kDBuf stages of its own, in turn
constexpr int kDStages = kDBuf * kDSlots;
constexpr int kDWRing = 3;                 // the 576x64 filter is not resident: its nine 8 KB tap slices stream through a
                                           // three-slot ring, two taps ahead of the MMAs. The 49 KB this frees go to the
                                           // L1 cache the bilinear gathers live on (the shared-memory carve-out drops from
                                           // 172 KB to 123 KB)
constexpr int kDGatherThreads = 3 * kDPix; // 768
constexpr int kDThreads = kDGatherThreads + 128 + 32;  // + 4 epilogue warps + MMA warp
constexpr int kDABytes = kDPix * 64 * 2;               // one tap: 2 M-tiles x 128 px x 64 ch bf16
constexpr int kDBBytes = 9 * 64 * 64 * 2;
constexpr int kDWTapBytes = kDBBytes / 9;              // 8192
constexpr int kDSmem = kDWRing * kDWTapBytes + kDStages * kDABytes + 256 + 9 * 64 * 4 + 1024;
static_assert(kDSmem <= 227 * 1024, "deform_umma_kernel: shared memory");

struct DeformParams {
  int N, H, W;
  int tiles_x, tiles_y, num_items;
  const __nv_bfloat16* x;       // slab8 [N][8][H][W][8]
  const float* off;             // slab4 [N][off_cs][H][W][4], channels 0..17 used
  int off_cs;
  const float* off_nchw;        // alternative (training path): offsets as fp32 NCHW (N,18,H,W); `off` is then unused
  float* out_nchw;              // alternative (training path): fp32 NCHW (N,64,H,W) output, no bf16 rounding; `out` unused
  const __nv_bfloat16* wpacked; // [9][8][8][8][8]
  const float* bias;
  int act;
  __nv_bfloat16* out;           // slab8 [N][out_cs_total][H][W][8] at slab offset out_cs0
  int out_cs_total, out_cs0;
  // optional fused "tap projection" of the FOLLOWING single-output deformable layer (deform_out1_*): proj[n][t][p] =
  // sum_c proj_w[c][t] * out[c][p] computed from the bf16-rounded outputs while they are in registers
  const float* proj_w;          // (1, 64, 3, 3) filter of the next layer, or NULL
  float* proj_out;              // [N][9][H*W]
};

// Sampling position of one tap: the four corner addresses are CLAMPED into the image and the
// bilinear weight of an out-of-image corner is zeroed, so all corner loads are unconditional and
// can be issued back to back (16 independent 16-byte loads in flight per thread).
struct TapPos {
  int o00, o01, o10, o11;   // pixel offsets (y * W + x) of the clamped corners
  float w00, w01, w10, w11;
};

__device__ __forceinline__ TapPos tap_pos(float dx, float dy, int x, int y, int tap, int H, int W) {
  float px = (float)(x + (tap % 3) - 1) + dx;
  float py = (float)(y + (tap / 3) - 1) + dy;
  px = fminf(fmaxf(px, -2.f), (float)W + 1.f);
  py = fminf(fmaxf(py, -2.f), (float)H + 1.f);
  const float fx0 = floorf(px), fy0 = floorf(py);
  const float fx = px - fx0, fy = py - fy0;
  const int x0 = (int)fx0, y0 = (int)fy0;
  const bool y0ok = y0 >= 0 && y0 < H, y1ok = y0 + 1 >= 0 && y0 + 1 < H;
  const bool x0ok = x0 >= 0 && x0 < W, x1ok = x0 + 1 >= 0 && x0 + 1 < W;
  const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
  const int ya = min(max(y0, 0), H - 1), yb = min(max(y0 + 1, 0), H - 1);
  TapPos t;
  t.o00 = ya * W + xa; t.o01 = ya * W + xb; t.o10 = yb * W + xa; t.o11 = yb * W + xb;
  t.w00 = (y0ok && x0ok) ? (1.f - fy) * (1.f - fx) : 0.f;
  t.w01 = (y0ok && x1ok) ? (1.f - fy) * fx : 0.f;
  t.w10 = (y1ok && x0ok) ? fy * (1.f - fx) : 0.f;
  t.w11 = (y1ok && x1ok) ? fy * fx : 0.f;
  return t;
}

// acc[0..7] += w * (the eight bf16 channels of v), as four packed fp32 FMAs (FFMA2: two IEEE fma per instruction --
// the same bits as eight scalar FFMAs, half the issue slots of the kernel's busiest loop)
__device__ __forceinline__ unsigned long long pack_f32x2(uint32_t lo, uint32_t hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void fma8(unsigned long long (&acc)[4], const uint4& v, float w) {
  const unsigned long long w2 = pack_f32x2(__float_as_uint(w), __float_as_uint(w));
  acc[0] = ffma2(pack_f32x2(v.x << 16, v.x & 0xffff0000u), w2, acc[0]);
  acc[1] = ffma2(pack_f32x2(v.y << 16, v.y & 0xffff0000u), w2, acc[1]);
  acc[2] = ffma2(pack_f32x2(v.z << 16, v.z & 0xffff0000u), w2, acc[2]);
  acc[3] = ffma2(pack_f32x2(v.w << 16, v.w & 0xffff0000u), w2, acc[3]);
}

struct Corners {
  uint4 c00, c01, c10, c11;
};
__device__ __forceinline__ Corners load_corners(const __nv_bfloat16* __restrict__ plane, const TapPos& t) {
  const uint4* p = reinterpret_cast<const uint4*>(plane);
  Corners c;
  c.c00 = __ldg(p + t.o00);
  c.c01 = __ldg(p + t.o01);
  c.c10 = __ldg(p + t.o10);
  c.c11 = __ldg(p + t.o11);
  return c;
}
__device__ __forceinline__ void blend8(const Corners& c, const TapPos& t, float (&out)[8]) {
  unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
  fma8(acc, c.c00, t.w00);
  fma8(acc, c.c01, t.w01);
  fma8(acc, c.c10, t.w10);
  fma8(acc, c.c11, t.w11);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float((uint32_t)(acc[i] & 0xffffffffull));
    out[2 * i + 1] = __uint_as_float((uint32_t)(acc[i] >> 32));
  }
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 t3 = __floats2bfloat162_rn(v[6], v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&t0);
  o.y = *reinterpret_cast<uint32_t*>(&t1);
  o.z = *reinterpret_cast<uint32_t*>(&t2);
  o.w = *reinterpret_cast<uint32_t*>(&t3);
  return o;
}

template <bool PROJ>   // PROJ: also compute the following single-output layer's tap projection in the epilogue
__global__ void __launch_bounds__(kDThreads, 1) deform_umma_kernel(const DeformParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smB = smem;                    // [kDWRing][8 KB] filter tap ring
  uint8_t* smA = smem + kDWRing * kDWTapBytes;
  uint64_t* bars = (uint64_t*)(smA + kDStages * kDABytes);
  uint64_t* full = bars;                  // gather -> MMA   (one arrive per gather warp)
  uint64_t* empty = bars + kDStages;      // MMA -> gather   (tcgen05.commit)
  uint64_t* tfull = bars + 2 * kDStages;  // MMA -> epilogue
  uint64_t* tempty = bars + 2 * kDStages + 2;
  uint64_t* wfull = bars + 2 * kDStages + 4;             // [kDWRing] filter tap landed
  uint64_t* wempty = wfull + kDWRing;                    // [kDWRing] the MMAs that read the slot have completed
  uint32_t* tmem_slot = (uint32_t*)(wempty + kDWRing);
  float* sproj = (float*)((uint8_t*)bars + 256);         // [9][64] projection filter, tap-major

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kMmaWarp = kDGatherThreads / 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDStages; ++s) {
      mbar_init(&full[s], kDPix / 32);   // one arrive per gather warp of the slot (256 threads)
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4);
    }
    for (int r = 0; r < kDWRing; ++r) {
      mbar_init(&wfull[r], 1);
      mbar_init(&wempty[r], 1);
    }
    fence_mbar_init();
  }
  if (PROJ)
    for (int i = threadIdx.x; i < 576; i += kDThreads) sproj[(i % 9) * 64 + i / 9] = p.proj_w[i];
  if (warp == kMmaWarp) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int items_per_img = p.tiles_x * p.tiles_y;
  const size_t plane = (size_t)p.H * p.W * 8;  // elements per (n, slab) plane of x

  if (warp < kMmaWarp) {
    // ======================= gather warps =======================
    // thread = (pixel, tap slot): the sampling position of a tap is computed once per pixel and
    // all 8 slabs (64 channels) of that tap are gathered by the same thread.
    const int t = threadIdx.x;
    const int pix = t & (kDPix - 1), slot = t >> 8;
    uint32_t fills = 0;
    // element (pixel m of M-tile j, slab q) of the stage lives at ((j * 8 + q) * 128 + m) * 16 bytes; the slot's fill f
    // goes to stage kDBuf * slot + f % kDBuf
    uint8_t* a0 = smA + (kDBuf * slot) * kDABytes + (pix >> 7) * (kDABytes / 2) + (pix & 127) * 16;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int n = item / items_per_img;
      const int r = item - n * items_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int y = ty * kDTileH + (pix >> 5), x = tx * kDTileW + (pix & 31);
      const bool valid = y < p.H && x < p.W;
      float odx[3], ody[3];
      if (valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int tap = slot + 3 * k;
          if (p.off_nchw != nullptr) {
            const size_t hw = (size_t)p.H * p.W, at = (size_t)y * p.W + x;
            odx[k] = __ldg(p.off_nchw + ((size_t)n * 18 + tap) * hw + at);
            ody[k] = __ldg(p.off_nchw + ((size_t)n * 18 + 9 + tap) * hw + at);
          } else {
            odx[k] = __ldg(p.off + ((((size_t)n * p.off_cs + (tap >> 2)) * p.H + y) * p.W + x) * 4 + (tap & 3));
            ody[k] = __ldg(p.off + ((((size_t)n * p.off_cs + ((9 + tap) >> 2)) * p.H + y) * p.W + x) * 4 + ((9 + tap) & 3));
          }
        }
      }
      const __nv_bfloat16* xin = p.x + (size_t)n * 8 * plane;
#pragma unroll
      for (int k = 0; k < 3; ++k, ++fills) {
        const int tap = slot + 3 * k;
        const int st = kDBuf * slot + (int)(fills % kDBuf);
        uint8_t* a = a0 + (fills % kDBuf) * kDABytes;
        mbar_wait(&empty[st], ((fills / kDBuf) & 1) ^ 1);
        if (valid) {
          const TapPos tp = tap_pos(odx[k], ody[k], x, y, tap, p.H, p.W);
#pragma unroll
          for (int s0 = 0; s0 < 8; s0 += 2) {
            Corners cr[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) cr[q] = load_corners(xin + (s0 + q) * plane, tp);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float v[8];
              blend8(cr[q], tp, v);
              *reinterpret_cast<uint4*>(a + (size_t)(s0 + q) * 2048) = pack8(v);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(a + (size_t)q * 2048) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ======================= MMA issuer (converged warp, one elected lane issues) =======================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    constexpr uint32_t a_hi = desc_hi(128u), b_hi = desc_hi(128u);
    const uint32_t smA_u = smem_u32(smA), smB_u = smem_u32(smB);
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked);
    const int my_items = p.num_items > (int)blockIdx.x ? (p.num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const long total_taps = 9L * my_items;
    // Filter tap g (counted over all of this CTA's items) sits in ring slot g % 3 and is requested two taps ahead:
    // tap g + 2 goes into the slot tap g - 1 was read from, once the MMAs of tap g - 1 have completed (wempty).
    auto load_tap = [&](long gg) {   // elected lane only
      const int r = (int)(gg % kDWRing);
      mbar_arrive_expect_tx(&wfull[r], kDWTapBytes);
      bulk_load(smB + r * kDWTapBytes, wsrc + (size_t)(gg % 9) * kDWTapBytes, kDWTapBytes, &wfull[r]);
    };
    if (elect_one_sync()) {
      if (total_taps > 0) load_tap(0);
      if (total_taps > 1) load_tap(1);
    }
    __syncwarp();
    uint32_t fcount[kDSlots] = {0, 0, 0};   // stages consumed per tap slot
    long g = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(buf * 128);
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap, ++g) {
        const int slot = tap % 3;
        const uint32_t f = fcount[slot]++;
        const int st = kDBuf * slot + (int)(f % kDBuf);
        const int r = (int)(g % kDWRing);
        mbar_wait(&full[st], (f / kDBuf) & 1u);
        mbar_wait(&wfull[r], (uint32_t)((g / kDWRing) & 1));
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smA_u + st * kDABytes, 2048u);
        const uint32_t bt_lo = desc_lo(smB_u + r * kDWTapBytes, 1024u);
        if (elect_one_sync()) {
          const uint32_t acc0 = tap != 0 ? 1u : 0u;
          // M-tile 0 (item rows 0-3), then M-tile 1 (rows 4-7, 16 KB further = 1024 x 16 B): 4 K-steps of 16 channels
          umma_bf16_off<0u, 0u>(d, a_lo, a_hi, bt_lo, b_hi, idesc, acc0);
          umma_bf16_off<256u, 128u>(d, a_lo, a_hi, bt_lo, b_hi, idesc, 1u);
          umma_bf16_off<512u, 256u>(d, a_lo, a_hi, bt_lo, b_hi, idesc, 1u);
          umma_bf16_off<768u, 384u>(d, a_lo, a_hi, bt_lo, b_hi, idesc, 1u);
          umma_bf16_off<1024u, 0u>(d + 64, a_lo, a_hi, bt_lo, b_hi, idesc, acc0);
          umma_bf16_off<1280u, 128u>(d + 64, a_lo, a_hi, bt_lo, b_hi, idesc, 1u);
          umma_bf16_off<1536u, 256u>(d + 64, a_lo, a_hi, bt_lo, b_hi, idesc, 1u);
          umma_bf16_off<1792u, 384u>(d + 64, a_lo, a_hi, bt_lo, b_hi, idesc, 1u);
          umma_commit(&empty[st]);
          umma_commit(&wempty[r]);
          if (tap == 8) umma_commit(&tfull[buf]);
        }
        __syncwarp();
        if (g + 2 < total_taps) {
          // slot of tap g + 2 = slot of tap g - 1 (its (g - 1) / 3-th use): wait for those MMAs, then refill
          if (g >= 1) mbar_wait(&wempty[(g + 2) % kDWRing], (uint32_t)(((g - 1) / kDWRing) & 1));
          if (elect_one_sync()) load_tap(g + 2);
          __syncwarp();
        }
      }
    }
  } else {
    // ======================= epilogue =======================
    const int q = warp & 3;
    const int m = 32 * q + lane;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int n = item / items_per_img;
      const int r = item - n * items_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int x = tx * kDTileW + (m & 31);
      const int buf = it & 1;
      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      float pr[9];
#pragma unroll
      for (int jc = 0; jc < 4; ++jc) {
        const int j = jc >> 1, c0 = (jc & 1) * 32;
        const int y = ty * kDTileH + 4 * j + (m >> 5);
        const bool valid = y < p.H && x < p.W;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * 128 + j * 64 + c0), acc);
        tmem_wait_ld();
        if (PROJ && c0 == 0) {
#pragma unroll
          for (int t = 0; t < 9; ++t) pr[t] = 0.f;
        }
        if (valid) {
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i] = __uint_as_float(acc[8 * s8 + i]) + __ldg(p.bias + c0 + 8 * s8 + i);
              if (p.act) v[i] = lrelu(v[i]);
            }
            if (!PROJ && p.out_nchw != nullptr) {   // consecutive lanes = consecutive x: 128-byte rows per channel
              const size_t hw = (size_t)p.H * p.W;
              float* op = p.out_nchw + ((size_t)n * 64 + c0 + 8 * s8) * hw + (size_t)y * p.W + x;
#pragma unroll
              for (int i = 0; i < 8; ++i) op[(size_t)i * hw] = v[i];
              continue;
            }
            const size_t cs = (size_t)n * p.out_cs_total + (p.out_cs0 + c0 / 8 + s8);
            const uint4 o = pack8(v);
            *reinterpret_cast<uint4*>(p.out + ((cs * p.H + y) * p.W + x) * 8) = o;
            if (PROJ) {
              // same operands (the bf16-rounded outputs) and the same channel order as deform_out1_project_kernel
              const float f[8] = {__uint_as_float(o.x << 16), __uint_as_float(o.x & 0xffff0000u),
                                  __uint_as_float(o.y << 16), __uint_as_float(o.y & 0xffff0000u),
                                  __uint_as_float(o.z << 16), __uint_as_float(o.z & 0xffff0000u),
                                  __uint_as_float(o.w << 16), __uint_as_float(o.w & 0xffff0000u)};
#pragma unroll
              for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int c = 0; c < 8; ++c) pr[t] = fmaf(f[c], sproj[t * 64 + c0 + 8 * s8 + c], pr[t]);
            }
          }
          if (PROJ && c0 == 32) {
            const size_t hw = (size_t)p.H * p.W;
#pragma unroll
            for (int t = 0; t < 9; ++t) p.proj_out[((size_t)n * 9 + t) * hw + (size_t)y * p.W + x] = pr[t];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---- final layer: 64 -> 1 -------------------------------------------------------------------------
// Bilinear sampling is linear, so  y = b + sum_tap sample( sum_c w[c,tap] * x[c] , pos_tap ):
// project the 64 channels onto the 9 taps FIRST (a 1x1 conv 64 -> 9 at integer pixels, 576 MAC per
// pixel, fully coalesced), then sample the nine scalar fields. The gather shrinks from
// 9 taps x 4 corners x 64 channels to 9 x 4 scalars per pixel (the direct form is L1-wavefront and
// issue bound: 288 16-byte gathers + ~5k instructions per pixel).
__global__ void __launch_bounds__(256) deform_out1_project_kernel(const __nv_bfloat16* __restrict__ x,
                                                                  const float* __restrict__ w,  // (1,64,3,3)
                                                                  float* __restrict__ proj,     // [N][9][H*W]
                                                                  int N, int HW) {
  __shared__ float sw[9][64];
  for (int i = threadIdx.x; i < 576; i += blockDim.x) sw[i % 9][i / 9] = w[i];
  __syncthreads();
  const long total = (long)N * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / HW, px = i - n * HW;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
#pragma unroll
    for (int slab = 0; slab < 8; ++slab) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((n * 8 + slab) * HW + px) * 8));
      const float f[8] = {__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                          __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u),
                          __uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u),
                          __uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u)};
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[t] = fmaf(f[c], sw[t][slab * 8 + c], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) proj[(n * 9 + t) * HW + px] = acc[t];
  }
}

__global__ void __launch_bounds__(256) deform_out1_sample_kernel(const float* __restrict__ proj,
                                                                 const float* __restrict__ off, int off_cs,
                                                                 const float* __restrict__ bias,
                                                                 float* __restrict__ y, int N, int H, int W) {
  const int HW = H * W;
  const long total = (long)N * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xx = i % W;
    const long r = i / W;
    const int yy = r % H;
    const long n = r / H;
    float offv[20];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float4 o4 =
          __ldg(reinterpret_cast<const float4*>(off + ((((size_t)n * off_cs + k) * H + yy) * W + xx) * 4));
      offv[4 * k] = o4.x; offv[4 * k + 1] = o4.y; offv[4 * k + 2] = o4.z; offv[4 * k + 3] = o4.w;
    }
    float acc = bias[0];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const TapPos tp = tap_pos(offv[tap], offv[9 + tap], xx, yy, tap, H, W);
      const float* pl = proj + (n * 9 + tap) * HW;
      acc = fmaf(tp.w00, __ldg(pl + tp.o00), acc);
      acc = fmaf(tp.w01, __ldg(pl + tp.o01), acc);
      acc = fmaf(tp.w10, __ldg(pl + tp.o10), acc);
      acc = fmaf(tp.w11, __ldg(pl + tp.o11), acc);
    }
    y[i] = acc;
  }
}

}  // namespace dbm

using namespace dbm;

extern "C" int dbm_deform_conv_umma(const void* x_slab8, const float* offset_slab4, int offset_cs_total,
                                    const void* wpacked_ck64, const float* bias, int n, int h, int w, int act,
                                    void* out_slab8, int out_cs_total, int out_cs0, const float* next_out1_filter,
                                    float* next_out1_proj, cudaStream_t stream) {
  DBM_REQUIRE((next_out1_filter == nullptr) == (next_out1_proj == nullptr),
              "deform_conv_umma: the fused tap projection needs both the filter and the output buffer");
  DBM_REQUIRE(n > 0 && h > 0 && w > 0, "deform_conv_umma: empty input");
  DBM_REQUIRE(offset_cs_total >= 5, "deform_conv_umma: offset tensor needs >= 18 channels (5 slabs)");
  DBM_REQUIRE(((uintptr_t)x_slab8 & 15) == 0 && ((uintptr_t)wpacked_ck64 & 15) == 0 &&
                  ((uintptr_t)offset_slab4 & 15) == 0 && ((uintptr_t)out_slab8 & 15) == 0,
              "deform_conv_umma: unaligned pointer");
  if (int rc = ensure_dyn_smem((const void*)deform_umma_kernel<false>, kDSmem)) return rc;
  if (int rc = ensure_dyn_smem((const void*)deform_umma_kernel<true>, kDSmem)) return rc;
  DeformParams p;
  p.N = n; p.H = h; p.W = w;
  p.tiles_x = ceil_div(w, kDTileW); p.tiles_y = ceil_div(h, kDTileH);
  p.num_items = n * p.tiles_x * p.tiles_y;
  p.x = (const __nv_bfloat16*)x_slab8; p.off = offset_slab4; p.off_cs = offset_cs_total;
  p.wpacked = (const __nv_bfloat16*)wpacked_ck64; p.bias = bias; p.act = act;
  p.out = (__nv_bfloat16*)out_slab8; p.out_cs_total = out_cs_total; p.out_cs0 = out_cs0;
  p.proj_w = next_out1_filter; p.proj_out = next_out1_proj;
  p.off_nchw = nullptr; p.out_nchw = nullptr;
  const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  if (p.proj_w != nullptr) deform_umma_kernel<true><<<grid, kDThreads, kDSmem, stream>>>(p);
  else deform_umma_kernel<false><<<grid, kDThreads, kDSmem, stream>>>(p);
  return check_launch("deform_umma_kernel");
}

// Training-path form of the same kernel (forward of final_conv_layer1 in GeneratorModel.forward_train): offsets read as
// the fp32 NCHW (N,18,H,W) tensor the offset convolution produced, output written as fp32 NCHW (N,64,H,W) without the
// bf16 storage rounding. Same arithmetic as the sampler + bf16 GEMM pair it replaces (bilinear samples and filter
// rounded to bf16, fp32 accumulation), without the 382 MB `cols` round trip through HBM.
extern "C" int dbm_deform_conv_umma_nchw(const void* x_slab8, const float* offset_nchw18, const void* wpacked_ck64,
                                         const float* bias, int n, int h, int w, int act, float* out_nchw,
                                         cudaStream_t stream) {
  DBM_REQUIRE(n > 0 && h > 0 && w > 0, "deform_conv_umma_nchw: empty input");
  DBM_REQUIRE(x_slab8 && offset_nchw18 && wpacked_ck64 && bias && out_nchw, "deform_conv_umma_nchw: null pointer");
  DBM_REQUIRE(((uintptr_t)x_slab8 & 15) == 0 && ((uintptr_t)wpacked_ck64 & 15) == 0,
              "deform_conv_umma_nchw: unaligned pointer");
  if (int rc = ensure_dyn_smem((const void*)deform_umma_kernel<false>, kDSmem)) return rc;
  DeformParams p;
  p.N = n; p.H = h; p.W = w;
  p.tiles_x = ceil_div(w, kDTileW); p.tiles_y = ceil_div(h, kDTileH);
  p.num_items = n * p.tiles_x * p.tiles_y;
  p.x = (const __nv_bfloat16*)x_slab8; p.off = nullptr; p.off_cs = 0;
  p.wpacked = (const __nv_bfloat16*)wpacked_ck64; p.bias = bias; p.act = act;
  p.out = nullptr; p.out_cs_total = 0; p.out_cs0 = 0;
  p.proj_w = nullptr; p.proj_out = nullptr;
  p.off_nchw = offset_nchw18; p.out_nchw = out_nchw;
  const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  deform_umma_kernel<false><<<grid, kDThreads, kDSmem, stream>>>(p);
  return check_launch("deform_umma_kernel");
}

// Re-sampling for the backward of the fused forward above: cols[n][c * 9 + tap][p] (fp32, the operand of the weight
// gradient GEMM) from the SAME bf16 slab8 copy of the input the forward gathered from, with the same 16-byte corner
// loads (one per 8 channels instead of one 4-byte load per channel from the fp32 NCHW tensor: dbm_deform_sample_f32
// took 0.60 ms at batch 128 on the generator's critical path). The values are the forward's samples before their bf16
// rounding, which the GEMM applies.
__global__ void __launch_bounds__(256) deform_sample_slab8_kernel(const __nv_bfloat16* __restrict__ x,
                                                                  const float* __restrict__ off_nchw,
                                                                  float* __restrict__ cols, int N, int H, int W) {
  const long hw = (long)H * W;
  const long total = (long)N * 9 * hw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pp = i % hw;
    const int tap = (int)((i / hw) % 9);
    const long n = i / (9 * hw);
    const int y = (int)(pp / W), xx = (int)(pp - (long)y * W);
    const float dx = __ldg(off_nchw + (n * 18 + tap) * hw + pp);
    const float dy = __ldg(off_nchw + (n * 18 + 9 + tap) * hw + pp);
    const TapPos tp = tap_pos(dx, dy, xx, y, tap, H, W);
    const __nv_bfloat16* xin = x + (size_t)n * 8 * hw * 8;
    float* out = cols + ((size_t)n * 576 + tap) * hw + pp;
#pragma unroll
    for (int s0 = 0; s0 < 8; s0 += 2) {
      Corners cr[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) cr[q] = load_corners(xin + (size_t)(s0 + q) * hw * 8, tp);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v[8];
        blend8(cr[q], tp, v);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) out[(size_t)((s0 + q) * 8 + c8) * 9 * hw] = v[c8];
      }
    }
  }
}

extern "C" int dbm_deform_sample_slab8_f32(const void* x_slab8, const float* offset_nchw18, float* cols, int n, int h,
                                           int w, cudaStream_t stream) {
  DBM_REQUIRE(n > 0 && h > 0 && w > 0 && x_slab8 && offset_nchw18 && cols, "deform_sample_slab8: bad arguments");
  DBM_REQUIRE(((uintptr_t)x_slab8 & 15) == 0, "deform_sample_slab8: unaligned input");
  const long total = (long)n * 9 * h * w;
  long blocks = (total + 255) / 256;
  const long cap = (long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  deform_sample_slab8_kernel<<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)x_slab8, offset_nchw18, cols, n, h, w);
  return check_launch("deform_sample_slab8_kernel");
}

// The sampling half of dbm_deform_conv_out1 alone: the nine projected planes were already produced by the preceding
// layer's fused epilogue (dbm_deform_conv_umma with next_out1_filter).
extern "C" int dbm_deform_out1_sample(const float* proj, const float* offset_slab4, int offset_cs_total,
                                      const float* bias, float* y, int n, int h, int w, cudaStream_t stream) {
  DBM_REQUIRE(n > 0 && h > 0 && w > 0 && proj && y, "deform_out1_sample: empty input");
  DBM_REQUIRE(offset_cs_total >= 5, "deform_out1_sample: offset tensor needs >= 18 channels (5 slabs)");
  const long total = (long)n * h * w;
  long blocks = (total + 255) / 256;
  const long cap = (long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  deform_out1_sample_kernel<<<(int)blocks, 256, 0, stream>>>(proj, offset_slab4, offset_cs_total, bias, y, n, h, w);
  return check_launch("deform_out1_sample_kernel");
}

extern "C" int dbm_deform_conv_out1(const void* x_slab8, const float* offset_slab4, int offset_cs_total,
                                    const float* w_f32, const float* bias, float* y, float* proj_scratch, int n,
                                    int h, int w, cudaStream_t stream) {
  DBM_REQUIRE(n > 0 && h > 0 && w > 0, "deform_conv_out1: empty input");
  DBM_REQUIRE(offset_cs_total >= 5, "deform_conv_out1: offset tensor needs >= 18 channels (5 slabs)");
  DBM_REQUIRE(proj_scratch != nullptr, "deform_conv_out1: needs a scratch buffer of n*9*h*w floats");
  const long total = (long)n * h * w;
  long blocks = (total + 255) / 256;
  const long cap = (long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  deform_out1_project_kernel<<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)x_slab8, w_f32, proj_scratch, n,
                                                              h * w);
  int rc = check_launch("deform_out1_project_kernel");
  if (rc) return rc;
  deform_out1_sample_kernel<<<(int)blocks, 256, 0, stream>>>(proj_scratch, offset_slab4, offset_cs_total, bias, y, n,
                                                             h, w);
  return check_launch("deform_out1_sample_kernel");
}
