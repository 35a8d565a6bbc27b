// Image-resident trunk forward for SMALL tiles (the reference's 9x9 training / doctest tiles,
// BASELINE.json configs[1], [3], [4]): pre-residual conv -> 3*nb residual dense blocks -> post-residual
// conv (GeneratorModel.forward, srgan_train.py:541-551; RDB :339-358; RRDB :397-402) with the
// activations of an image never leaving the SM.
//
// An image WITH its one-pixel zero border is at most 128 flat positions (the flat-padded layout of
// umma_flat.cu), i.e. ONE M=128 UMMA tile: nothing a conv reads belongs to another CTA, so the
// layer-to-layer dependency that the flat chain kernel pays for with global flags, fences and a
// HBM/L2 round trip per layer (~8 us per layer at batch 128) is a shared-memory write + mbarrier here.
//   * dense-block buffer [a0 | a1 | a2 | a3 | a4] (192 channels bf16) lives in shared memory in the
//     UMMA K-major core-matrix layout ([slab][row][8 ch], tap = start-address shift, as umma_flat.cu);
//   * the contraction is INPUT-stationary: as soon as a_s exists, ONE pass of N = 192 - 32 s columns
//     adds its contribution to every later conv of the block (conv_{s+1} .. conv_5) -- same FLOPs, same
//     per-column accumulation order as the layer-by-layer form, but N = 64..192 instead of 32/64 (a
//     tcgen05.mma re-reads its 4 KB A tile per instruction whatever N is, profiles/README.md) and the
//     partial sums stay in TMEM; pass s completes conv_{s+1}, whose epilogue writes a_{s+1} to smem;
//   * the fp32 residual stream x_j sits in TMEM next to the accumulators; the RRDB input and the
//     pre-residual output (needed once per RRDB / once per image) in a small global scratch;
//   * a CTA carries TWO images in lock step through the same weight stream (5-stage bulk-copy ring,
//     3 taps x 16 channels x N per stage), halving the L2->SMEM filter traffic per image.
// With save pointers set the bf16 activations of every dense block are also written to the flat
// buffers the data-/weight-gradient kernels of umma_flat.cu read (training).
#include "common.cuh"

namespace dbm {

constexpr int kLocThreads = 320;           // warp 0 bulk-copy producer, warp 1 MMA issuer, warps 2-9 epilogue
constexpr int kLocStages = 5;
// SOLO form (batches larger than the SM count): one image per CTA, 192 threads (four epilogue warps), 3 stages,
// 256 TMEM columns and a shared-memory footprint under half an SM, so TWO CTAs share an SM: one image's pass
// hand-off (commit -> epilogue -> fence -> next MMA, ~2/3 of a pass at batch 128) hides under the other's MMAs.
constexpr int kLocSoloThreads = 192;
constexpr int kLocSoloStages = 3;
constexpr int kLocStageBytes = 96 * 192;   // 3 taps x 2 slabs x (N/8) x 128 B at N = 192
constexpr int kLocSlabs = 24;              // 192 channels
constexpr int kLocAcc = 192, kLocXcur = 192, kLocSlot = 256;  // TMEM columns per image slot

enum { kLocPre = 0, kLocAct = 1, kLocRdb = 2, kLocPost = 3,        // forward epilogues
       kLocBPost = 4, kLocBMask = 5, kLocBD1 = 6, kLocBPre = 7 };  // data-gradient epilogues

struct LocalPass {  // 96 bytes; mirrored by deepbedmap_b200/flat.py (LOCAL_PASS_DTYPE)
  const __nv_bfloat16* w;     // [nk][9][2][N/8][8][8] packed filter slice of this pass
  const float* bias;          // forward: bias of the conv this pass completes
  __nv_bfloat16* save;        // flat bf16 slab pointer of the outputs (first slab), or NULL
  float* out_f32;             // kLocPost / kLocBPre: flat fp32 slab4 output
  const __nv_bfloat16* mask;  // backward: flat bf16 activations whose sign selects the LeakyReLU derivative, or NULL
  const float* add_f32;       // kLocBD1: flat fp32 slab4 addend (the a3 = a1 + ... skip, :551), or NULL
  int slab0, nk, N, col0;     // MMA: first input slab, 16-channel K-steps, columns, first accumulator column
  int type, ecol, out_slab, rr;  // epilogue: type, accumulator column of the finished block, smem slab of its output,
                                 // rr = 1 when the block closes (forward) / opens (backward) an RRDB
  float beta;                 // forward: residual scaling; backward: sigma, the scale of the incoming dX
  float scale;                // backward: bf16 output = bf16(scale * v)
  int first;                  // 1: the pass starts a fresh accumulation (first MMA overwrites)
  int pad;
};
static_assert(sizeof(LocalPass) == 96, "LocalPass layout is part of the C ABI");

struct LocalParams {
  const LocalPass* passes;
  int count;
  int n, H, W, Wp, img, halo, G0, Pg, RA;
  const __nv_bfloat16* s0;  // flat bf16 [in_slabs][Pg][8]: stem output (forward) / bf16 d(loss)/d(a3) (backward)
  int in_slabs;             // 16 / 8
  int group;                // images a CTA carries in lock step: 2, or 1 when the batch has fewer images than 2 x SMs
  float* x0;                // scratch [n][16][128][4]: pre-residual output (fp32)                  (forward)
  float* xrr;               // scratch [n][16][128][4]: input of the current RRDB (fp32) / gradient wrt the input
                            // of the RRDB behind the current one (backward)
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint4 pack8f(const float* v) {
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

template <bool BWD, bool SOLO>
__global__ void __launch_bounds__(SOLO ? kLocSoloThreads : kLocThreads, SOLO ? 2 : 1)
local_trunk_kernel(const LocalParams p) {
  constexpr int kStages = SOLO ? kLocSoloStages : kLocStages;
  constexpr int kThreads = SOLO ? kLocSoloThreads : kLocThreads;
  constexpr int kSlots = SOLO ? 1 : 2;          // image slots (operand buffers, TMEM column blocks, epilogue groups)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)smem;
  uint64_t* empty = full + kLocStages;          // (barrier area sized for the larger ring in both forms)
  uint64_t* tfull = empty + kLocStages;   // MMA -> epilogue: a pass is complete
  uint64_t* act_ready = tfull + 1;        // epilogue -> MMA: outputs are in smem, accumulator columns are free
  uint64_t* in_full = act_ready + 1;      // producer -> MMA: the stem outputs of the image pair are in smem
  uint64_t* a_free = in_full + 1;         // MMA -> producer: every MMA of the image pair has read its operands
  uint32_t* tmem_slot = (uint32_t*)(a_free + 1);
  const uint32_t slab_bytes = (uint32_t)p.RA * 16u;
  const uint32_t abuf_bytes = (uint32_t)kLocSlabs * slab_bytes;
  uint8_t* abuf = smem + 256;                       // [2 images][24 slabs][RA rows][16 B]
  uint8_t* stages = abuf + kSlots * abuf_bytes;     // [kStages][kLocStageBytes]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // borders, guard rows and the slabs no bulk copy touches must read as zero
  for (uint32_t i = threadIdx.x; i < kSlots * abuf_bytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(abuf)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    mbar_init(act_ready, 4 * kSlots);
    mbar_init(in_full, 1);
    mbar_init(a_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<SOLO ? 256 : 512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int npairs = (p.n + p.group - 1) / p.group;
  const uint32_t load_rows = (uint32_t)(2 * p.halo + p.img);

  if (warp == 0) {
    // ================= producer: stem outputs of the pair, then the filter stream =================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        if (it > 0) mbar_wait(a_free, (uint32_t)((it - 1) & 1));
        const int nact = min(p.group, p.n - p.group * pair);
        mbar_arrive_expect_tx(in_full, (uint32_t)(nact * p.in_slabs) * load_rows * 16u);
        for (int g = 0; g < nact; ++g) {
          const long pos0 = (long)p.G0 + (long)(p.group * pair + g) * p.img - p.halo;
          for (int sl = 0; sl < p.in_slabs; ++sl)
            bulk_load(abuf + g * abuf_bytes + sl * slab_bytes, p.s0 + ((long)sl * p.Pg + pos0) * 8, load_rows * 16u,
                      in_full);
        }
        for (int l = 0; l < p.count; ++l) {
          const __nv_bfloat16* w = p.passes[l].w;
          const int nk = p.passes[l].nk;
          const uint32_t bytes = 96u * (uint32_t)p.passes[l].N;
          for (int g3 = 0; g3 < 3 * nk; ++g3) {      // (K-step, tap row) granules, contiguous in the packed image
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_load(stages + s * kLocStageBytes, w + (size_t)g3 * (bytes / 2), bytes, &full[s]);
            if (++s == kStages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, one elected lane issues) =================
    const uint32_t a_hi = desc_hi(128), b_hi = desc_hi(128);
    const uint32_t ab_u = smem_u32(abuf), st_u = smem_u32(stages);
    int s = 0; uint32_t ph = 0, actph = 0; int it = 0; long gp = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
      const int nact = min(p.group, p.n - p.group * pair);
      mbar_wait(in_full, (uint32_t)(it & 1));
      for (int l = 0; l < p.count; ++l, ++gp) {
        const int N = p.passes[l].N, nk = p.passes[l].nk, slab0 = p.passes[l].slab0, col0 = p.passes[l].col0;
        const int first = p.passes[l].first;
        const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)N);
        const uint32_t b_lbo = (uint32_t)(N / 8) * 128u;
        const uint32_t b_tap = (2u * b_lbo) >> 4;
        if (gp > 0) {   // the previous pass's outputs are in smem and its accumulator columns have been read
          mbar_wait(act_ready, actph);
          actph ^= 1;
        }
        tc_fence_after();
        for (int kc = 0; kc < nk; ++kc) {
          for (int tg = 0; tg < 3; ++tg) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t b_lo = desc_lo(st_u + s * kLocStageBytes, b_lbo);
            const bool last = (kc == nk - 1) && (tg == 2);
            if (elect_one_sync()) {
              for (int g = 0; g < nact; ++g) {
                const uint32_t a_lo = desc_lo(ab_u + g * abuf_bytes + (uint32_t)(slab0 + 2 * kc) * slab_bytes, slab_bytes);
                const uint32_t d = tmem_base + (uint32_t)(g * kLocSlot + col0);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                  const uint32_t a_off = (uint32_t)(tg * p.Wp + t);   // ky * Wp + kx rows of 16 bytes
                  const uint32_t acc = (first && kc == 0 && tg == 0 && t == 0) ? 0u : 1u;
                  umma_bf16(d, make_desc(a_lo + a_off, a_hi), make_desc(b_lo + (uint32_t)t * b_tap, b_hi), idesc, acc);
                }
              }
              umma_commit(&empty[s]);
              if (last) umma_commit(tfull);
            }
            __syncwarp();
            if (++s == kStages) { s = 0; ph ^= 1; }
          }
        }
      }
      if (elect_one_sync()) umma_commit(a_free);
      __syncwarp();
    }
  } else {
    // ================= epilogue: group g = image slot g; thread = flat position m of the image =================
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;
    const int m = 32 * q + lane;
    const int y = m / p.Wp, x = m - y * p.Wp;
    uint8_t* arow = abuf + g * abuf_bytes + (uint32_t)(p.halo + m) * 16u;
    const uint32_t tbase = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(g * kLocSlot);
    uint32_t tph = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int im = g < p.group ? p.group * pair + g : p.n;   // slot 1 idles when group == 1
      const bool interior = im < p.n && m < p.img && y >= 1 && y <= p.H && x >= 1 && x <= p.W;
      const long gpos = (long)p.G0 + (long)im * p.img + m;                 // flat position
      float* x0p = p.x0 + ((size_t)im * 16 * 128 + m) * 4;                 // + c4 * 512
      float* xrp = p.xrr + ((size_t)im * 16 * 128 + m) * 4;
      if constexpr (BWD) {
        // ---------------- data-gradient chain ----------------
        // accumulator columns = d[a0 | a1 | a2 | a3 | a4] of the current dense block: conv5's data gradient starts
        // it (N = 192), conv_k's adds onto columns [0, 64 + 32 (k - 1)); after conv_k's pass the slot a_{k-1} is final:
        // LeakyReLU derivative (sign of the kept bf16 activation) -> g_{k-1}, the operand of the next pass.
        for (int l = 0; l < p.count; ++l) {
          const LocalPass& P = p.passes[l];
          const int type = P.type;
          const int nh = type == kLocBMask ? 1 : (type == kLocBPre ? 4 : 2);
          const bool has_mask = P.mask != nullptr;
          uint4 mk[4];
          if (has_mask && interior) {
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8)
              mk[s8] = *reinterpret_cast<const uint4*>(P.mask + ((size_t)s8 * p.Pg + gpos) * 8);
          }
          mbar_wait(tfull, tph);
          tph ^= 1;
          tc_fence_after();
          for (int h = 0; h < nh; ++h) {
            uint32_t acc[32];
            float v[32];
            tmem_ld_32x32b_x32(tbase + (uint32_t)(P.ecol + 32 * h), acc);
            if (h == 1 && has_mask && interior) {
#pragma unroll
              for (int s8 = 0; s8 < 4; ++s8)
                mk[s8] = *reinterpret_cast<const uint4*>(P.mask + ((size_t)(4 + s8) * p.Pg + gpos) * 8);
            }
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
            if (type == kLocBPre) {   // d(loss)/d(stem output): fp32, flat slab4
              if (interior) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                  *reinterpret_cast<float4*>(P.out_f32 + ((size_t)(8 * h + c4) * p.Pg + gpos) * 4) =
                      make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
              }
              continue;
            }
            if (type == kLocBD1) {
              // d a0 = (dense-block paths) + sigma * dX_{j+1} (:358) [+ d a3 (:551)] [+ dX of the RRDB behind (:402)]
              uint32_t xc[32];
              tmem_ld_32x32b_x32(tbase + (uint32_t)(kLocXcur + 32 * h), xc);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __fmaf_rn(P.beta, __uint_as_float(xc[i]), v[i]);
              if (interior && P.add_f32 != nullptr) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                  const float4 t = *reinterpret_cast<const float4*>(P.add_f32 + ((size_t)(8 * h + c4) * p.Pg + gpos) * 4);
                  v[4 * c4] += t.x; v[4 * c4 + 1] += t.y; v[4 * c4 + 2] += t.z; v[4 * c4 + 3] += t.w;
                }
              }
              if (interior && P.rr) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                  const float4 t = *reinterpret_cast<const float4*>(xrp + (size_t)(8 * h + c4) * 512);
                  v[4 * c4] += t.x; v[4 * c4 + 1] += t.y; v[4 * c4 + 2] += t.z; v[4 * c4 + 3] += t.w;
                }
              }
            }
            if (has_mask && interior) {
#pragma unroll
              for (int s8 = 0; s8 < 4; ++s8) {
                const uint32_t w4[4] = {mk[s8].x, mk[s8].y, mk[s8].z, mk[s8].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {   // bf16 sign bits: low half = even channel, high half = odd channel
                  if (w4[j] & 0x00008000u) v[8 * s8 + 2 * j] *= kLreluSlope;
                  if (w4[j] & 0x80000000u) v[8 * s8 + 2 * j + 1] *= kLreluSlope;
                }
              }
            }
            if (type == kLocBPost || (type == kLocBD1 && !has_mask)) {
              // gradient wrt the block input: the next (earlier) block's incoming dX; at an RRDB boundary also the
              // skip gradient of the RRDB in front
              uint32_t xs[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) xs[i] = __float_as_uint(v[i]);
              tmem_st_32x32b_x32(tbase + (uint32_t)(kLocXcur + 32 * h), xs);
              if (interior && (type == kLocBPost || P.rr)) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                  *reinterpret_cast<float4*>(xrp + (size_t)(8 * h + c4) * 512) =
                      make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
              }
            }
            const float sc = P.scale;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= sc;
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
              const uint4 o = pack8f(v + 8 * s8);
              if (interior) {
                *reinterpret_cast<uint4*>(arow + (size_t)(P.out_slab + 4 * h + s8) * slab_bytes) = o;
                // kept for the weight gradient (dbm_flat_wgrad); fire-and-forget, the hand-off does not wait for it
                if (P.save != nullptr) *reinterpret_cast<uint4*>(P.save + ((size_t)(4 * h + s8) * p.Pg + gpos) * 8) = o;
              }
            }
          }
          tmem_wait_st();
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(act_ready);
        }
        continue;
      }
      for (int l = 0; l < p.count; ++l) {
        const LocalPass& P = p.passes[l];
        const int type = P.type;
        const int nh = type == kLocAct ? 1 : 2;                            // 32-column halves
        float b0[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(P.bias) + i4);
          b0[4 * i4] = t.x; b0[4 * i4 + 1] = t.y; b0[4 * i4 + 2] = t.z; b0[4 * i4 + 3] = t.w;
        }
        mbar_wait(tfull, tph);
        tph ^= 1;
        tc_fence_after();
        uint4 keep[8];   // bf16 outputs, written to the flat save buffer after the hand-off
        for (int h = 0; h < nh; ++h) {
          uint32_t acc[32];
          float v[32];
          tmem_ld_32x32b_x32(tbase + (uint32_t)(P.ecol + 32 * h), acc);
          if (h == 1) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(P.bias) + 8 + i4);
              b0[4 * i4] = t.x; b0[4 * i4 + 1] = t.y; b0[4 * i4 + 2] = t.z; b0[4 * i4 + 3] = t.w;
            }
          }
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __fadd_rn(__uint_as_float(acc[i]), b0[i]);
          if (type == kLocPre || type == kLocAct) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = lrelu(v[i]);
          }
          if (type == kLocRdb) {
            uint32_t xc[32];
            tmem_ld_32x32b_x32(tbase + (uint32_t)(kLocXcur + 32 * h), xc);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __fmaf_rn(P.beta, v[i], __uint_as_float(xc[i]));   // RDB skip (:358)
            if (P.rr) {                                                                   // RRDB skip (:402)
              if (interior) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                  const float4 t = *reinterpret_cast<const float4*>(xrp + (size_t)(8 * h + c4) * 512);
                  v[4 * c4] = __fmaf_rn(P.beta, v[4 * c4], t.x);
                  v[4 * c4 + 1] = __fmaf_rn(P.beta, v[4 * c4 + 1], t.y);
                  v[4 * c4 + 2] = __fmaf_rn(P.beta, v[4 * c4 + 2], t.z);
                  v[4 * c4 + 3] = __fmaf_rn(P.beta, v[4 * c4 + 3], t.w);
                }
              }
            }
          }
          if (type == kLocPost) {
            if (interior) {
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 t = *reinterpret_cast<const float4*>(x0p + (size_t)(8 * h + c4) * 512);
                v[4 * c4] += t.x; v[4 * c4 + 1] += t.y; v[4 * c4 + 2] += t.z; v[4 * c4 + 3] += t.w;   // a3 = a1 + conv (:551)
                if (P.out_f32 != nullptr)
                  *reinterpret_cast<float4*>(P.out_f32 + ((size_t)(8 * h + c4) * p.Pg + gpos) * 4) =
                      make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
              }
              if (P.save != nullptr) {
                // inference: a3 goes straight to the first upsample conv as bf16 slab8 [n][8][2H][2W][8], nearest x2
                // (F.resize_images, srgan_train.py:556-558) fused into the store
                const int Ho = 2 * p.H, Wo = 2 * p.W;
#pragma unroll
                for (int s8 = 0; s8 < 4; ++s8) {
                  const uint4 o = pack8f(v + 8 * s8);
                  __nv_bfloat16* base =
                      P.save + ((((size_t)im * 8 + 4 * h + s8) * Ho + 2 * (y - 1)) * Wo + 2 * (x - 1)) * 8;
                  *reinterpret_cast<uint4*>(base) = o;
                  *reinterpret_cast<uint4*>(base + 8) = o;
                  *reinterpret_cast<uint4*>(base + (size_t)Wo * 8) = o;
                  *reinterpret_cast<uint4*>(base + (size_t)Wo * 8 + 8) = o;
                }
              }
            }
            continue;
          }
          if (type == kLocPre || type == kLocRdb) {
            // the block output is the next block's residual stream (TMEM) and, when it closes an RRDB (or is
            // the pre-residual conv), the next RRDB's skip input (global scratch)
            uint32_t xs[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) xs[i] = __float_as_uint(v[i]);
            tmem_st_32x32b_x32(tbase + (uint32_t)(kLocXcur + 32 * h), xs);
            if (interior && (type == kLocPre || P.rr)) {
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 t = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
                *reinterpret_cast<float4*>(xrp + (size_t)(8 * h + c4) * 512) = t;
                if (type == kLocPre) *reinterpret_cast<float4*>(x0p + (size_t)(8 * h + c4) * 512) = t;
              }
            }
          }
          // bf16 operand of the following passes
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            const uint4 o = pack8f(v + 8 * s8);
            keep[4 * h + s8] = o;
            if (interior) *reinterpret_cast<uint4*>(arow + (size_t)(P.out_slab + 4 * h + s8) * slab_bytes) = o;
          }
        }
        // hand-off: TMEM reads / writes retired, smem writes visible to the tensor core's (async) proxy
        tmem_wait_st();
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(act_ready);
        if (P.save != nullptr && interior && type != kLocPost) {
#pragma unroll
          for (int s8 = 0; s8 < 8; ++s8)
            if (s8 < 4 * nh) *reinterpret_cast<uint4*>(P.save + ((size_t)s8 * p.Pg + gpos) * 8) = keep[s8];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<SOLO ? 256 : 512>(tmem_base);
  }
}

}  // namespace dbm

using namespace dbm;

static int g_local_group = 0;  // dbm_debug_set(4, v): force images per CTA (tuning)

static int local_trunk_launch(bool bwd, const void* passes_dev, int count, int n, int h, int w, const void* in_flat,
                              float* x0_scratch, float* xrr_scratch, cudaStream_t stream) {
  const char* who = bwd ? "trunk_local_bwd" : "trunk_local_fwd";
  DBM_REQUIRE(passes_dev && count > 0 && n > 0 && h > 0 && w > 0, "%s: empty problem", who);
  DBM_REQUIRE(in_flat && xrr_scratch && (bwd || x0_scratch), "%s: null buffer", who);
  LocalParams p;
  p.passes = (const LocalPass*)passes_dev;
  p.count = count;
  p.n = n; p.H = h; p.W = w; p.Wp = w + 2; p.img = (h + 2) * (w + 2);
  DBM_REQUIRE(p.img <= 128, "%s: a padded image of %dx%d is %d positions; the image-resident kernel holds at most 128 "
              "(use the flat chain / tiled kernels)", who, h + 2, w + 2, p.img);
  p.halo = p.Wp + 1;
  p.G0 = (p.halo + 7) & ~7;                                    // as flat_geom() in umma_flat.cu
  const int tiles = (n * p.img + 127) / 128;
  p.Pg = p.G0 + tiles * 128 + p.G0;
  p.s0 = (const __nv_bfloat16*)in_flat; p.in_slabs = bwd ? 8 : 16;
  p.x0 = x0_scratch; p.xrr = xrr_scratch;
  // Form: batches that fit the SMs take one image per CTA (320 threads, 5 stages); larger ones the SOLO form, two
  // CTAs per SM (dbm_local_debug_set: 1 / 2 = force 1 / 2 images per CTA in lock step, 3 = force SOLO).
  const bool solo = g_local_group == 3 || (g_local_group == 0 && n > num_sms());
  size_t smem;
  if (solo) {
    // slab stride = exactly the rows a bulk copy fills: the <= 2 halo rows an MMA reads past a slab's end alias the
    // next slab's leading (zero) halo rows -- or, after the last slab, ring bytes that only reach discarded rows
    p.RA = 2 * p.halo + p.img;
    p.group = 1;
    smem = 1024 + 256 + (size_t)kLocSlabs * p.RA * 16 + (size_t)kLocSoloStages * kLocStageBytes;
  } else {
    p.RA = 128 + 2 * p.halo;
    p.group = (g_local_group == 1 || g_local_group == 2) ? g_local_group : 1;
    smem = 1024 + 256 + 2 * (size_t)kLocSlabs * p.RA * 16 + (size_t)kLocStages * kLocStageBytes;
  }
  DBM_REQUIRE(smem <= 227 * 1024, "%s: image width %d needs %zu bytes of shared memory", who, w, smem);
  const void* kfn = bwd ? (solo ? (const void*)local_trunk_kernel<true, true> : (const void*)local_trunk_kernel<true, false>)
                        : (solo ? (const void*)local_trunk_kernel<false, true> : (const void*)local_trunk_kernel<false, false>);
  if (int rc = ensure_dyn_smem(kfn, smem)) return rc;
  const int npairs = (n + p.group - 1) / p.group;
  const int cap = solo ? 2 * num_sms() : num_sms();
  const int grid = npairs < cap ? npairs : cap;
  if (bwd && solo) local_trunk_kernel<true, true><<<grid, kLocSoloThreads, smem, stream>>>(p);
  else if (bwd) local_trunk_kernel<true, false><<<grid, kLocThreads, smem, stream>>>(p);
  else if (solo) local_trunk_kernel<false, true><<<grid, kLocSoloThreads, smem, stream>>>(p);
  else local_trunk_kernel<false, false><<<grid, kLocThreads, smem, stream>>>(p);
  return check_launch(bwd ? "local_trunk_kernel<bwd>" : "local_trunk_kernel<fwd>");
}

extern "C" int dbm_local_debug_set(int value) {
  g_local_group = value;
  return DBM_OK;
}

extern "C" int dbm_trunk_local_fwd(const void* passes_dev, int count, int n, int h, int w, const void* s0_flat,
                                   float* x0_scratch, float* xrr_scratch, cudaStream_t stream) {
  return local_trunk_launch(false, passes_dev, count, n, h, w, s0_flat, x0_scratch, xrr_scratch, stream);
}

extern "C" int dbm_trunk_local_bwd(const void* passes_dev, int count, int n, int h, int w, const void* gpost_flat,
                                   float* dxrr_scratch, cudaStream_t stream) {
  return local_trunk_launch(true, passes_dev, count, n, h, w, gpost_flat, nullptr, dxrr_scratch, stream);
}
