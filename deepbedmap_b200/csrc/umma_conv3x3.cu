// 3x3 'same' convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM),
// fed by TMA, with bias / LeakyReLU / residual-scaling / dense-concat slot write / nearest-x2
// replication fused into the epilogue.
//
// Replaces, for the reference's generator trunk (srgan_train.py:339-358, 397-402, 541-568),
// the Chainer op chains  L.Convolution2D(k3,s1,p1) [+ F.leaky_relu] [+ F.concat] [+ *beta, F.add]
// [+ F.resize_images nearest].
//
// Data layout in HBM ("slab" layouts; one 16-byte vector per pixel per slab):
//   bf16 activations  slab8 : [N][C/8][H][W][8]
//   fp32 residuals    slab4 : [N][C/4][H][W][4]
// A work item is a 16x16-pixel output unit of one image = two M=128 UMMA tiles (8 wide x 16
// high each). Its 18x18 halo tile is loaded ONCE per 32-channel chunk by one TMA box
// (out-of-bounds zero fill = the conv's zero padding); the nine filter taps are nine shifted
// UMMA shared-memory descriptors into that same tile (K-major, no-swizzle core matrices:
// 8 consecutive pixels x 8 channels = 128 contiguous bytes), so activations cross L2->SMEM
// 1.27x instead of 9x.
#include "umma_common.cuh"

namespace dbm {

static int g_debug_swap_lbo_sbo = 0;
static int g_debug_ck16 = 0;   // 64-wide layers in 16-channel chunks (the trunk kernel's summation order)
extern int g_trunk_debug;

constexpr int kThreads = 192;          // warp0 TMA, warp1 MMA, warps2-5 epilogue

struct UmmaConvParams {
  int N, H, W, Cin;
  int tiles_x, tiles_y, num_items;
  const __nv_bfloat16* wpacked;  // [Cin/CK][9][CK/8][COUT/8][8 cout][8 cin]
  const float* bias;             // [COUT]
  float beta;
  int act, up2, swap;
  int in_off;                    // 0: 'same' conv (zero padding by TMA out-of-bounds fill); 1: 'valid' conv -- output
                                 // (y, x) of the (H, W) grid is centred on input (y + 1, x + 1) of an (H+2, W+2) one
  __nv_bfloat16* out_bf16;
  int out_cs_total, out_cs0;
  float* out_f32;
  int out_f32_cs_total, out_f32_cs0;
  const float* res1;
  const float* res2;
};

template <int COUT, int CK, int STAGES>
struct UmmaCfg {
  static constexpr int A_BYTES = kHalo * kHalo * CK * 2;
  static constexpr int B_BYTES = 9 * CK * COUT * 2;
  static constexpr int TMEM_COLS = 4 * COUT;  // 2 sub-tiles x 2 accumulator buffers
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 256 + 1024;
};

template <int COUT, int CK, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
umma_conv3x3_kernel(const __grid_constant__ CUtensorMap tmap_in, const UmmaConvParams p) {
  using Cfg = UmmaCfg<COUT, CK, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = (uint64_t*)(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_in);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kc = p.Cin / CK;
  const int items_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int n = item / items_per_img;
        const int r = item - n * items_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        for (int kc = 0; kc < num_kc; ++kc) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES);
          tma_load_4d(smA + s * Cfg::A_BYTES, &tmap_in, &full[s], (tx * kTile - 1 + p.in_off) * 8,
                      ty * kTile - 1 + p.in_off, kc * (CK / 8), n);
          bulk_load(smB + s * Cfg::B_BYTES, p.wpacked + (size_t)kc * (Cfg::B_BYTES / 2), Cfg::B_BYTES,
                    &full[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: converged warp, one elected lane issues =================
    constexpr uint32_t idesc = umma_idesc_bf16(128, COUT);
    uint32_t a_lbo = kHalo * kHalo * 16, a_sbo = kHalo * 16;
    uint32_t b_lbo = (COUT / 8) * 128, b_sbo = 128;
    if (p.swap) {
      uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t;
      t = b_lbo; b_lbo = b_sbo; b_sbo = t;
    }
    const uint32_t a_hi = desc_hi(a_sbo), b_hi = desc_hi(b_sbo);
    const uint32_t smA_u = smem_u32(smA), smB_u = smem_u32(smB);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(buf * 2 * COUT);
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smA_u + s * Cfg::A_BYTES, a_lbo);
        const uint32_t b_lo = desc_lo(smB_u + s * Cfg::B_BYTES, b_lbo);
        const uint32_t acc0 = kc != 0 ? 1u : 0u;
        if (elect_one_sync()) {
          issue_stage_mmas<COUT, CK>(d0, a_lo, a_hi, b_lo, b_hi, idesc, acc0);
          umma_commit(&empty[s]);
          if (kc == num_kc - 1) umma_commit(&tfull[buf]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> HBM =================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = 32 * q + lane;
    const int g = m >> 3, xr = m & 7;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int n = item / items_per_img;
      const int r = item - n * items_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int buf = it & 1;
      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      const int y = ty * kTile + g;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int x = tx * kTile + 8 * j + xr;
        const bool valid = (y < p.H) && (x < p.W);
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * 2 * COUT + j * COUT + c0),
                             acc);
          tmem_wait_ld();
          if (valid) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) + __ldg(p.bias + c0 + i);
            if (p.res1) {
#pragma unroll
              for (int s4 = 0; s4 < 8; ++s4) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(
                    p.res1 + ((((size_t)n * (COUT / 4) + (c0 / 4 + s4)) * p.H + y) * p.W + x) * 4));
                v[4 * s4 + 0] = rr.x + p.beta * v[4 * s4 + 0];
                v[4 * s4 + 1] = rr.y + p.beta * v[4 * s4 + 1];
                v[4 * s4 + 2] = rr.z + p.beta * v[4 * s4 + 2];
                v[4 * s4 + 3] = rr.w + p.beta * v[4 * s4 + 3];
              }
            }
            if (p.res2) {
#pragma unroll
              for (int s4 = 0; s4 < 8; ++s4) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(
                    p.res2 + ((((size_t)n * (COUT / 4) + (c0 / 4 + s4)) * p.H + y) * p.W + x) * 4));
                v[4 * s4 + 0] = rr.x + p.beta * v[4 * s4 + 0];
                v[4 * s4 + 1] = rr.y + p.beta * v[4 * s4 + 1];
                v[4 * s4 + 2] = rr.z + p.beta * v[4 * s4 + 2];
                v[4 * s4 + 3] = rr.w + p.beta * v[4 * s4 + 3];
              }
            }
            if (p.act) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = lrelu(v[i]);
            }
            if (p.out_f32) {
#pragma unroll
              for (int s4 = 0; s4 < 8; ++s4) {
                float4 o = make_float4(v[4 * s4], v[4 * s4 + 1], v[4 * s4 + 2], v[4 * s4 + 3]);
                *reinterpret_cast<float4*>(
                    p.out_f32 +
                    ((((size_t)n * p.out_f32_cs_total + (p.out_f32_cs0 + c0 / 4 + s4)) * p.H + y) * p.W + x) * 4) = o;
              }
            }
            if (p.out_bf16) {
#pragma unroll
              for (int s8 = 0; s8 < 4; ++s8) {
                uint4 o;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * s8 + 0], v[8 * s8 + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * s8 + 2], v[8 * s8 + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * s8 + 4], v[8 * s8 + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * s8 + 6], v[8 * s8 + 7]);
                o.x = *reinterpret_cast<uint32_t*>(&t0);
                o.y = *reinterpret_cast<uint32_t*>(&t1);
                o.z = *reinterpret_cast<uint32_t*>(&t2);
                o.w = *reinterpret_cast<uint32_t*>(&t3);
                const size_t cs = (size_t)n * p.out_cs_total + (p.out_cs0 + c0 / 8 + s8);
                if (!p.up2) {
                  *reinterpret_cast<uint4*>(p.out_bf16 + ((cs * p.H + y) * p.W + x) * 8) = o;
                } else {
                  const int Ho = 2 * p.H, Wo = 2 * p.W;
                  __nv_bfloat16* base = p.out_bf16 + ((cs * Ho + 2 * y) * Wo + 2 * x) * 8;
                  // two adjacent pixels = 32 contiguous bytes per output row
                  *reinterpret_cast<uint4*>(base) = o;
                  *reinterpret_cast<uint4*>(base + 8) = o;
                  *reinterpret_cast<uint4*>(base + (size_t)Wo * 8) = o;
                  *reinterpret_cast<uint4*>(base + (size_t)Wo * 8 + 8) = o;
                }
              }
            }
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- host side --------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// slab8 bf16 tensor [N][CS][H][W][8] viewed as 4-D (W*8, H, CS, N); box = 18 px x 18 rows x CK/8 slabs.
int make_slab8_tmap(CUtensorMap* tm, const void* base, int N, int CS, int H, int W, int ck, int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode();
  DBM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)CS, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)CS * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, (cuuint32_t)(ck / 8), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DBM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for N=%d CS=%d H=%d W=%d", (int)r, N, CS,
              H, W);
  return DBM_OK;
}

template <int COUT, int CK, int STAGES>
static int launch_umma(const CUtensorMap& tm, const UmmaConvParams& p, cudaStream_t st) {
  using Cfg = UmmaCfg<COUT, CK, STAGES>;
  if (int rc = ensure_dyn_smem((const void*)umma_conv3x3_kernel<COUT, CK, STAGES>, Cfg::SMEM)) return rc;
  int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  umma_conv3x3_kernel<COUT, CK, STAGES><<<grid, kThreads, Cfg::SMEM, st>>>(tm, p);
  return check_launch("umma_conv3x3_kernel");
}

// fp32 OIHW 3x3 weights -> bf16 UMMA operand image [Cin/CK][9][CK/8][COUTP/8][8][8]
// Same operand image, filled from a channel slice of a wider filter: rows [o0, o0 + O) of the
// image take w[o][c0 + c][tap] (o < O, c < Cin) of an (O, CinTotal, 3, 3) filter; other rows are
// left untouched, so several filters can be stacked along Cout (dense-block layer pairing).
__global__ void pack_w3x3_slice_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int O, int o0,
                                       int Cin, int CinTotal, int c0, int COUTP, int CK) {
  const long total = (long)9 * Cin * COUTP;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c8 = t % 8; t /= 8;
    const int o8 = t % 8; t /= 8;
    const int cg = t % (COUTP / 8); t /= (COUTP / 8);
    const int ksl = t % (CK / 8); t /= (CK / 8);
    const int tap = t % 9; t /= 9;
    const int kc = (int)t;
    const int o = cg * 8 + o8 - o0;
    const int c = kc * CK + ksl * 8 + c8;
    if (o >= 0 && o < O) out[i] = __float2bfloat16_rn(w[((long)o * CinTotal + c0 + c) * 9 + tap]);
  }
}

// Table-driven form: one launch re-packs every filter of a model (blockIdx.y = table entry). The
// training step re-packs the generator's ~600 operand images after every Adam update; one launch
// instead of ~1200 keeps that off the critical path.
// One thread produces one 16-byte vector (8 consecutive GEMM-K indices c8 of one output row): an eighth of the index
// arithmetic and of the store instructions of the element-per-thread form, which cost the training step 0.38 ms per
// weight update (five launches re-packing ~34 M bf16 values, bound by integer division, not by memory).
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* dst, const float (&v)[8]) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 t3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&t0);
  o.y = *reinterpret_cast<uint32_t*>(&t1);
  o.z = *reinterpret_cast<uint32_t*>(&t2);
  o.w = *reinterpret_cast<uint32_t*>(&t3);
  *reinterpret_cast<uint4*>(dst) = o;
}

__global__ void pack_w3x3_table_kernel(const PackEntry* __restrict__ table) {
  PackEntry e = table[blockIdx.y];
  // mode bit 16: split-bf16 image for dbm_trunk_umma_split -- every 16-channel chunk kc becomes three chunks
  // [w_hi | w_hi | w_lo] (w = w_hi + w_lo, both bf16); forward slices with 8-aligned rows only
  const bool split = (e.mode & 16) != 0;
  e.mode &= 15;
  if (split && !(e.mode == 0 && (e.O & 7) == 0 && (e.o0 & 7) == 0 && e.CK == 16)) return;   // rejected on the host
  if (e.mode == 0 && (e.O & 7) == 0 && (e.o0 & 7) == 0) {
    // stacked forward slices (pair / tail / input-stationary images): visit only this entry's own rows -- the general
    // loop below scans the whole image for every entry that writes into it
    const int og = e.O / 8;
    const long total = (long)9 * (e.Cin / 8) * e.O;   // vectors
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      long t = i;
      const int o8 = t % 8; t /= 8;
      const int cgl = t % og; t /= og;
      const int ksl = t % (e.CK / 8); t /= (e.CK / 8);
      const int tap = t % 9; t /= 9;
      const int kc = (int)t;
      const int o = cgl * 8 + o8;
      const int c = kc * e.CK + ksl * 8;
      const int kcd = split ? 3 * kc : kc;
      const long dst = ((((long)(kcd * 9 + tap) * (e.CK / 8) + ksl) * (e.COUTP / 8) + (e.o0 / 8 + cgl)) * 8 + o8) * 8;
      const float* src = e.w + ((long)o * e.CinTotal + e.c0 + c) * 9 + tap;
      float v[8];
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) v[c8] = src[c8 * 9];
      store8_bf16(e.out + dst, v);
      if (split) {
        const long chunk = (long)9 * (e.CK / 8) * (e.COUTP / 8) * 64;
        store8_bf16(e.out + dst + chunk, v);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) v[c8] -= __bfloat162float(__float2bfloat16_rn(v[c8]));
        store8_bf16(e.out + dst + 2 * chunk, v);
      }
    }
    return;
  }
  const long total = (long)9 * (e.Cin / 8) * e.COUTP;   // vectors: element index = vector * 8 + c8
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int o8 = t % 8; t /= 8;
    const int cg = t % (e.COUTP / 8); t /= (e.COUTP / 8);
    const int ksl = t % (e.CK / 8); t /= (e.CK / 8);
    const int tap = t % 9; t /= 9;
    const int kc = (int)t;
    const int on = cg * 8 + o8;                 // GEMM N index (row of the operand image)
    const int cb = kc * e.CK + ksl * 8;         // GEMM K index of element 0 of the vector
    float v[8];
    if (e.mode == 0) {
      // forward operand; rows [o0, o0 + O) <- w[o][c0 + c][tap], other rows untouched (stacked filters)
      const int o = on - e.o0;
      if (o < 0 || o >= e.O) continue;
      const float* src = e.w + ((long)o * e.CinTotal + e.c0 + cb) * 9 + tap;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) v[c8] = src[c8 * 9];
      store8_bf16(e.out + i * 8, v);
      continue;
    }
    // modes 1-3 write the whole image. o0 = number of REAL filter output channels along the padded
    // output-channel axis (0 = all): the remaining rows / K-lines are zero.
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const int c = cb + c8;
      float x = 0.f;
      if (e.mode == 1) {
        // data-gradient operand of a 3x3 filter w (Oreal, CinTotal, 3, 3): N index = input channel c0 + on,
        // K index c = output channel, taps flipped
        const int kvalid = e.o0 > 0 ? e.o0 : e.Cin;
        if (on < e.O && c < kvalid) x = e.w[((long)c * e.CinTotal + e.c0 + on) * 9 + (8 - tap)];
      } else {
        // 4x4 stride-2 pad-1 filter w4 (Oreal, C, 4, 4), C = CinTotal, embedded as a 3x3 stride-1 filter over the four
        // space-to-depth phases (channel = phase * C + cc): tap (ty, tx) of phase (py, px) is
        // w4[.., 2(ty-1)+py+1, 2(tx-1)+px+1] when that index exists, else 0.
        // mode 2: forward operand (N = output channel on, K = phase channel c; the w pointer is pre-offset per
        // output chunk); mode 3: data-gradient operand (N = phase channel c0 + on, K = output channel c, taps flipped).
        const int oc = e.mode == 2 ? on : c;             // filter output channel
        const int pc = e.mode == 2 ? c : e.c0 + on;      // phase channel
        const int ovalid = e.o0 > 0 ? e.o0 : (e.mode == 2 ? e.O : e.Cin);
        const int t3 = e.mode == 2 ? tap : 8 - tap;
        const int ph = pc / e.CinTotal, cc = pc - ph * e.CinTotal;
        const int ky = 2 * (t3 / 3 - 1) + (ph >> 1) + 1, kx = 2 * (t3 % 3 - 1) + (ph & 1) + 1;
        if (on < e.O && oc < ovalid && ph < 4 && ky >= 0 && ky <= 3 && kx >= 0 && kx <= 3)
          x = e.w[(((long)oc * e.CinTotal + cc) * 4 + ky) * 4 + kx];
      }
      v[c8] = x;
    }
    store8_bf16(e.out + i * 8, v);
  }
}

__global__ void pack_w3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int O, int Cin,
                                 int COUTP, int CK) {
  const long total = (long)9 * Cin * COUTP;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c8 = t % 8; t /= 8;
    const int o8 = t % 8; t /= 8;
    const int cg = t % (COUTP / 8); t /= (COUTP / 8);
    const int ksl = t % (CK / 8); t /= (CK / 8);
    const int tap = t % 9; t /= 9;
    const int kc = (int)t;
    const int o = cg * 8 + o8;
    const int c = kc * CK + ksl * 8 + c8;
    float v = 0.f;
    if (o < O) v = w[((long)o * Cin + c) * 9 + tap];
    out[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace dbm

using namespace dbm;

extern "C" int dbm_debug_set(int key, int value) {
  if (key == 1) g_debug_swap_lbo_sbo = value;
  if (key == 2) g_debug_ck16 = value;
  if (key == 3) g_trunk_debug = value;
  return DBM_OK;
}

extern "C" int dbm_pack_conv3x3_weights(const float* w_oihw, void* packed_bf16, int cout, int cin, int cout_padded,
                                        int ck, cudaStream_t stream) {
  DBM_REQUIRE(ck == 16 || ck == 32 || ck == 64,
              "pack: K-chunk %d must be 16 / 32 (conv3x3_umma, trunk) or 64 (deform_conv_umma)", ck);
  DBM_REQUIRE(cin % ck == 0, "pack: Cin=%d must be a multiple of %d", cin, ck);
  DBM_REQUIRE(cout_padded == 32 || cout_padded == 64, "pack: padded Cout=%d must be 32 or 64", cout_padded);
  DBM_REQUIRE(cout <= cout_padded, "pack: Cout=%d > padded %d", cout, cout_padded);
  const long total = (long)9 * cin * cout_padded;
  pack_w3x3_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(w_oihw, (__nv_bfloat16*)packed_bf16, cout, cin,
                                                             cout_padded, ck);
  return check_launch("pack_w3x3_kernel");
}

extern "C" int dbm_pack_conv3x3_weights_slice(const float* w_oihw, int w_cin_total, int w_c0, void* packed_bf16,
                                              int cout, int cout0, int cin, int cout_padded, int ck,
                                              cudaStream_t stream) {
  DBM_REQUIRE(ck == 16 || ck == 32 || ck == 64, "pack: K-chunk %d must be 16, 32 or 64", ck);
  DBM_REQUIRE(cin % ck == 0 && w_c0 >= 0 && w_c0 + cin <= w_cin_total, "pack: bad channel slice [%d, %d) of %d",
              w_c0, w_c0 + cin, w_cin_total);
  DBM_REQUIRE(cout_padded % 8 == 0 && cout0 >= 0 && cout0 + cout <= cout_padded, "pack: bad Cout rows [%d, %d) of %d",
              cout0, cout0 + cout, cout_padded);
  const long total = (long)9 * cin * cout_padded;
  pack_w3x3_slice_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(w_oihw, (__nv_bfloat16*)packed_bf16, cout, cout0,
                                                                   cin, w_cin_total, w_c0, cout_padded, ck);
  return check_launch("pack_w3x3_slice_kernel");
}

extern "C" int dbm_pack_conv3x3_table(const void* table_dev, int num_entries, long max_elements,
                                      cudaStream_t stream) {
  DBM_REQUIRE(num_entries > 0 && num_entries <= 65535 && max_elements > 0, "pack table: bad size (%d entries)",
              num_entries);
  DBM_REQUIRE(((uintptr_t)table_dev & 7) == 0, "pack table: unaligned");
  int gx = ceil_div(max_elements / 8, 256 * 2);   // one thread per 8-element vector, ~2 vectors per thread
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  pack_w3x3_table_kernel<<<dim3(gx, num_entries), 256, 0, stream>>>((const PackEntry*)table_dev);
  return check_launch("pack_w3x3_table_kernel");
}

extern "C" int dbm_conv3x3_umma(const void* in_slab8, int in_cs_total, int cin, const void* wpacked,
                                const float* bias, int cout_padded, int n, int h, int w, float beta, int act,
                                int up2, void* out_slab8, int out_cs_total, int out_cs0, float* out_f32_slab4,
                                int out_f32_cs_total, int out_f32_cs0, const float* res1_slab4,
                                const float* res2_slab4, cudaStream_t stream) {
  DBM_REQUIRE(cin % 32 == 0 && cin <= in_cs_total * 8, "conv3x3_umma: bad Cin=%d (slabs %d)", cin, in_cs_total);
  DBM_REQUIRE(cout_padded == 32 || cout_padded == 64, "conv3x3_umma: Cout=%d must be 32 or 64", cout_padded);
  DBM_REQUIRE(n > 0 && h > 0 && w > 0, "conv3x3_umma: empty input");
  DBM_REQUIRE(out_slab8 || out_f32_slab4, "conv3x3_umma: no output");
  DBM_REQUIRE(((uintptr_t)in_slab8 & 15) == 0 && ((uintptr_t)wpacked & 15) == 0, "conv3x3_umma: unaligned");
  CUtensorMap tm;
  int rc = make_slab8_tmap(&tm, in_slab8, n, in_cs_total, h, w, 32);
  if (rc) return rc;
  UmmaConvParams p;
  p.N = n; p.H = h; p.W = w; p.Cin = cin;
  p.tiles_x = ceil_div(w, kTile); p.tiles_y = ceil_div(h, kTile);
  p.num_items = n * p.tiles_x * p.tiles_y;
  p.wpacked = (const __nv_bfloat16*)wpacked; p.bias = bias; p.beta = beta; p.act = act; p.up2 = up2;
  p.swap = g_debug_swap_lbo_sbo;
  p.in_off = 0;
  p.out_bf16 = (__nv_bfloat16*)out_slab8; p.out_cs_total = out_cs_total; p.out_cs0 = out_cs0;
  p.out_f32 = out_f32_slab4; p.out_f32_cs_total = out_f32_cs_total; p.out_f32_cs0 = out_f32_cs0;
  p.res1 = res1_slab4; p.res2 = res2_slab4;
  if (g_debug_ck16) {
    rc = make_slab8_tmap(&tm, in_slab8, n, in_cs_total, h, w, 16);
    if (rc) return rc;
    if (cout_padded == 32) return launch_umma<32, 16, 5>(tm, p, stream);
    return launch_umma<64, 16, 5>(tm, p, stream);
  }
  if (cout_padded == 32) return launch_umma<32, 32, 5>(tm, p, stream);
  return launch_umma<64, 32, 3>(tm, p, stream);
}

// 3x3 VALID convolution (no padding) with 32 output channels: out (n, h_in - 2, w_in - 2) bf16 slab8 slot =
// conv(in (n, h_in, w_in) slab8) + bias. Used for conv_on_W1 of the input block (srgan_train.py:231, 259) over the
// 10x10 space-to-depth split operand built by dbm_stem_w1_s2d (csrc/stem.cu).
extern "C" int dbm_conv3x3_umma_valid(const void* in_slab8, int in_cs_total, int cin, const void* wpacked,
                                      const float* bias, int n, int h_in, int w_in, void* out_slab8, int out_cs_total,
                                      int out_cs0, cudaStream_t stream) {
  DBM_REQUIRE(cin % 32 == 0 && cin <= in_cs_total * 8, "conv3x3_umma_valid: bad Cin=%d (slabs %d)", cin, in_cs_total);
  DBM_REQUIRE(n > 0 && h_in >= 3 && w_in >= 3, "conv3x3_umma_valid: input %dx%d too small", h_in, w_in);
  DBM_REQUIRE(out_slab8 != nullptr && out_cs0 + 4 <= out_cs_total, "conv3x3_umma_valid: bad output slot");
  DBM_REQUIRE(((uintptr_t)in_slab8 & 15) == 0 && ((uintptr_t)wpacked & 15) == 0, "conv3x3_umma_valid: unaligned");
  CUtensorMap tm;
  int rc = make_slab8_tmap(&tm, in_slab8, n, in_cs_total, h_in, w_in, 32);
  if (rc) return rc;
  UmmaConvParams p;
  p.N = n; p.H = h_in - 2; p.W = w_in - 2; p.Cin = cin;
  p.tiles_x = ceil_div(p.W, kTile); p.tiles_y = ceil_div(p.H, kTile);
  p.num_items = n * p.tiles_x * p.tiles_y;
  p.wpacked = (const __nv_bfloat16*)wpacked; p.bias = bias; p.beta = 0.f; p.act = 0; p.up2 = 0;
  p.swap = g_debug_swap_lbo_sbo;
  p.in_off = 1;
  p.out_bf16 = (__nv_bfloat16*)out_slab8; p.out_cs_total = out_cs_total; p.out_cs0 = out_cs0;
  p.out_f32 = nullptr; p.out_f32_cs_total = 0; p.out_f32_cs0 = 0;
  p.res1 = nullptr; p.res2 = nullptr;
  return launch_umma<32, 32, 5>(tm, p, stream);
}
