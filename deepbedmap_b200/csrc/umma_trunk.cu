// Persistent whole-trunk kernel: pre-residual conv -> 36 residual dense blocks (5 convs each) ->
// post-residual conv (GeneratorModel.forward, srgan_train.py:541-551, RDB :339-358, RRDB :397-402)
// as ONE launch instead of 182.
//
// Every layer is the same tcgen05 implicit-GEMM 3x3 conv as umma_conv3x3_kernel (TMA halo tile,
// nine shifted descriptors, fused epilogue); what changes is the scheduling. All (layer, 32-row x 16-column
// pixel unit) work items are numbered layer-major and dealt round-robin to one resident CTA per
// SM. An item of layer L may start once the <= 9 neighbouring units of layer L-1 (its 1-pixel
// halo) are complete, which each finished item publishes through a global flag
// (st.global -> fence -> atomicAdd / ld.acquire -> fence.proxy.async -> TMA). That removes the
// per-layer launch + pipeline fill/drain (~8 us x 182 at continent-tile size) and the tail wave
// of every layer: CTAs flow into the next layer while the last units of the previous one finish.
#include "umma_common.cuh"

namespace dbm {

constexpr int kTrunkThreads = 320;  // warp0 TMA + dependency wait, warp1 MMA, warps2-9 epilogue
constexpr int kEpiWarps = 8;        // two groups of four warps, one group per TMEM accumulator buffer
// A work item is a 32-row x 16-column pixel unit = four M=128 sub-tiles (8 columns x 16 rows each)
// that share every weight stage: halves the weight traffic (L2 -> SMEM) and the per-item / per-stage
// hand-offs of the 16x16 unit the per-layer kernel uses, and gives the epilogue of the short K=32
// "tail" passes twice the time to drain.
constexpr int kTW = 16, kTH = 32;                    // (a TMA box is at most 256 elements = 32 pixels wide)
constexpr int kHW = kTW + 2, kHH = kTH + 2;          // 18-px x 34-row halo tile
// Every pass streams K in 16-channel chunks: one stage = halo tile (19584 B) + 9 taps x 16 x Cout
// filter slice (<= 18432 B); five stages cover the ~3000-cycle TMA round trip.
constexpr int kTStages = 5;
constexpr int kTABytes = kHW * kHH * 16 * 2;         // 19584
constexpr int kTBBytesMax = 9 * 16 * 64 * 2;         // 18432
constexpr int kMaxTrunkLayers = 512;                 // 23 RRDB (config 5's deepest) = 347 passes
constexpr int kTrunkSmem = kTStages * (kTABytes + kTBBytesMax) + 256 + kMaxTrunkLayers * 28 + kEpiWarps * 256 + 1024;

struct TrunkMaps {  // one 18 px x 34 rows x 16-channel box map per input buffer
  CUtensorMap m[3];
};

// All MMAs of one stage: 9 taps x 4 sub-tiles, K = 16 (one instruction each). Tap loop rolled for
// the same uniform-register reason as issue_stage_mmas.
template <int COUT>
__device__ __forceinline__ void issue_stage_wide(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t acc0) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, COUT);
  constexpr uint32_t kBTap = (uint32_t)(2 * (COUT / 8) * 8);  // per-tap stride of the packed filter, 16-B units
#pragma unroll 1
  for (uint32_t tap = 0; tap < 9; ++tap) {
    const uint32_t a_tap = a_lo + tap + (tap / 3) * (kHW - 3);  // ky * 18 + kx
    const uint32_t b_tap = b_lo + tap * kBTap;
    const uint32_t acc = tap != 0 ? 1u : acc0;
    umma_bf16_off<0, 0>(d0, a_tap, a_hi, b_tap, b_hi, idesc, acc);
    umma_bf16_off<8, 0>(d0 + 64, a_tap, a_hi, b_tap, b_hi, idesc, acc);
    umma_bf16_off<16 * kHW, 0>(d0 + 128, a_tap, a_hi, b_tap, b_hi, idesc, acc);
    umma_bf16_off<16 * kHW + 8, 0>(d0 + 192, a_tap, a_hi, b_tap, b_hi, idesc, acc);
  }
}

extern int g_trunk_debug;
extern unsigned long long* g_trunk_prof;

struct TrunkParams {
  const TrunkLayer* layers;
  int num_layers;
  int N, H, W, tiles_x, tiles_y, items_per_layer;
  unsigned int* done;  // [num_layers][items_per_layer], zeroed before the launch; complete == kEpiWarps
  unsigned long long* prof;  // tuning only (NULL = off): [num_layers][8] cycle counters, see scripts/trunk_ablate.py
  int debug;           // ablation mask for tuning runs (results invalid): 1 no dependency wait, 2 no epilogue
                       // memory traffic (32 fp32 stores only, 64 bf16 stores only, 128 residual loads only), 4 no TMA
                       // loads, 8 / 16 fence placement, 256 all images alias image 0 (what an L2-resident working set would buy)
};

__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// The trunk's fp32 side buffers (residual stream, stash) are "slab8f": fp32 [N][C/8][H][W][8], one
// 32-byte vector per pixel per slab, moved with 256-bit loads / stores (half the LSU instructions of
// the 16-byte slab4 form; the epilogue shares the L1 data path with the UMMA operand fetch).
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld_cg_f8(const float* p) {  // L2-coherent (data written by other SMs)
  F8 r;
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                 "=f"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_f8(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6,
                                      float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3),
               "f"(a4), "f"(a5), "f"(a6), "f"(a7)
               : "memory");
}

// Split-bf16 storage: v = hi + lo with hi = bf16(v), lo = bf16(v - hi) (16 mantissa bits together).
__device__ __forceinline__ void split_bf16x8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __bfloat1622float2(t);
    const __nv_bfloat162 u = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&t);
    l[i] = *reinterpret_cast<const uint32_t*>(&u);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// physical slab of logical 8-channel slab L in a split buffer: per 16 channels [hi a | hi b | lo a | lo b]
__device__ __forceinline__ int split_slab(int L) { return ((L >> 1) << 2) + (L & 1); }

// ---- schedule -----------------------------------------------------------------------------------------------
// The pass table is cut into GROUPS: a pair (head, tail) or a single pass. Within a group the I = N x tiles units are
// dealt to the Gd = gridDim.x resident CTAs in rounds of Gd; a rotation that advances by I mod Gd per group moves the
// CTAs that get the short last round around, as the pass-major numbering of round 1 did. A CTA walks the rounds
// r0, r1, ... of a pair group software-pipelined over the two TMEM accumulator buffers (round k in buffer k & 1):
//     head(r0) head(r1) tail(r0) head(r2) tail(r1) head(r3) tail(r2) ... head(rR-1) tail(rR-2) tail(rR-1)
// so a tail runs one head AND one tail item after its own head: the neighbours' head epilogues (the tail's halo) have
// ~13k cycles to land in L2 before the tail asks for them (their epilogue + flag + TMA round trip is ~8k), and only the
// last tail of a group follows its head closely. Every dependency of an item points to a strictly earlier step of
// another CTA's walk (neighbouring units sit in the same or an adjacent round), so the walk cannot deadlock as long as
// all CTAs are resident.
// All three warp roles walk the same sequence with their own cursor; (group, step) -> item is a pure function.
struct Cursor {
  int grp = 0, step = 0;
};
struct Sched {
  const int* groups;   // shared memory: first table index of the group, bit 30 set for a pair
  int ng, I, Gd, rounds, bid;
  __device__ __forceinline__ bool next(Cursor& c, int& L, int& unit, int& buf) const {
    while (c.grp < ng) {
      const int e = groups[c.grp];
      const bool pair = (e >> 30) & 1;
      const int nsteps = pair ? 2 * rounds : rounds;
      if (c.step >= nsteps) {
        ++c.grp;
        c.step = 0;
        continue;
      }
      const int t = c.step++;
      int round = t, tail = 0;
      if (pair) {
        if (t == 0) round = 0;
        else if (t == nsteps - 1) { round = rounds - 1; tail = 1; }
        else if (t & 1) round = (t + 1) >> 1;
        else { round = (t >> 1) - 1; tail = 1; }
      }
      int r = bid - (int)(((long)c.grp * I) % Gd);
      if (r < 0) r += Gd;
      unit = round * Gd + r;
      if (unit >= I) continue;
      L = (e & 0x3FFFFFFF) + tail;
      buf = round & 1;
      return true;
    }
    return false;
  }
};

// SPLIT: split-bf16 arithmetic (dbm_trunk_umma_split) -- a template parameter so that the bf16 instantiation carries
// none of its address arithmetic (a run-time flag cost the bf16 path 6 %: 739 -> 694 TFLOP/s on the same box)
template <bool SPLIT>
__global__ void __launch_bounds__(kTrunkThreads, 1)
umma_trunk_kernel(const __grid_constant__ TrunkMaps maps, const TrunkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = smem;
  uint8_t* smB = smem + kTStages * kTABytes;
  uint64_t* bars = (uint64_t*)(smem + kTStages * (kTABytes + kTBBytesMax));
  uint64_t* full = bars;
  uint64_t* empty = bars + kTStages;
  uint64_t* tfull = bars + 2 * kTStages;
  uint64_t* tempty = bars + 2 * kTStages + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kTStages + 4);
  int* ng_slot = (int*)(tmem_slot + 1);
  // per-layer scalars the producer / MMA warps need for every item, staged once in shared memory
  // (the epilogue's gpu-scope fences keep invalidating L1, a global read per item costs an L2 trip)
  int4* linfo = (int4*)(smem + kTStages * (kTABytes + kTBBytesMax) + 256);         // {cin, cout, in_map | mode << 8, in_cs0}
  const __nv_bfloat16** lw = (const __nv_bfloat16**)(linfo + kMaxTrunkLayers);
  float* sbias_all = (float*)(lw + kMaxTrunkLayers);  // [kEpiWarps][64]
  int* groups = (int*)(sbias_all + kEpiWarps * 64);   // [kMaxTrunkLayers]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int L = threadIdx.x; L < p.num_layers; L += kTrunkThreads) {
    const TrunkLayer* ly = p.layers + L;
    linfo[L] = make_int4(ly->cin, ly->cout, ly->in_map | (ly->mode << 8), ly->in_cs0);
    lw[L] = ly->wpacked;
  }
  if (threadIdx.x == 0) {
    int ng = 0;
    for (int L = 0; L < p.num_layers; ++L) {
      const int mode = p.layers[L].mode;
      // a pair head (N = 64, the upper 32 columns stay in TMEM) must be followed directly by its tail (N = 32)
      const bool ok = mode == 0 || (mode == 1 && p.layers[L].cout == 64 && L + 1 < p.num_layers &&
                                    p.layers[L + 1].mode == 2 && p.layers[L + 1].cout == 32) ||
                      (mode == 2 && L > 0 && p.layers[L - 1].mode == 1);
      if (!ok) {
        if (blockIdx.x == 0) printf("dbm: trunk pass table entry %d: bad pairing (mode %d)\n", L, mode);
        __trap();
      }
      if (mode == 2) continue;                       // belongs to the head before it
      groups[ng++] = L | (mode == 1 ? (1 << 30) : 0);
    }
    *ng_slot = ng;
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) tma_prefetch_desc(&maps.m[i]);
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], kEpiWarps / 2);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int I = p.items_per_layer;
  const int per_img = p.tiles_x * p.tiles_y;
  Sched sc;
  sc.groups = groups; sc.ng = *ng_slot; sc.I = I; sc.Gd = (int)gridDim.x; sc.bid = (int)blockIdx.x;
  sc.rounds = (I + sc.Gd - 1) / sc.Gd;
  const unsigned int need = (unsigned)(kEpiWarps / 2);   // warps that publish a unit

  if (warp == 0) {
    // ================= dependency wait + TMA producer (converged warp) =================
    // Lanes 0..8 each watch one neighbouring unit of the previous pass. The flags of the NEXT item are requested
    // (relaxed gpu-scope loads, no ordering attached) before this item's stages are issued, so that whenever the
    // previous pass finished long ago the L2 round trip of the dependency check is paid under the TMA loop instead of
    // in front of every item (round 1: 1.3-2.0k cycles of operand wait per item at the MMA issuer). The acquire side
    // is the gpu-scope fence after the observed values.
    int s = 0;
    uint32_t ph = 0;
    auto flag_of = [&](int L, int unit) -> const unsigned int* {
      if (L == 0 || lane >= 9) return nullptr;
      const int n = unit / per_img;
      const int r = unit - n * per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int ny = ty + lane / 3 - 1, nx = tx + lane % 3 - 1;
      if (ny < 0 || ny >= p.tiles_y || nx < 0 || nx >= p.tiles_x) return nullptr;
      return p.done + (size_t)(L - 1) * I + (size_t)n * per_img + ny * p.tiles_x + nx;
    };
    Cursor cur;
    int L, unit, buf;
    bool have = sc.next(cur, L, unit, buf);
    const unsigned int* f_cur = have ? flag_of(L, unit) : nullptr;
    unsigned int v_cur = f_cur != nullptr ? ld_relaxed_gpu(f_cur) : need;
    while (have) {
      const int n = unit / per_img;
      const int r = unit - n * per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int4 li = linfo[L];
      const int cin = li.x, cout = li.y, in_map = li.z & 0xFF, in_cs0 = li.w;
      const __nv_bfloat16* wp = lw[L];
      if (L > 0 && !(p.debug & 1)) {
        if (f_cur != nullptr) {
          SpinGuard guard;
          while (v_cur < need) {
            if (guard.expired()) {
              printf("dbm: trunk dependency timeout layer %d unit %d\n", L, unit);
              __trap();
            }
            __nanosleep(64);
            v_cur = ld_relaxed_gpu(f_cur);
          }
        }
        __syncwarp();
        fence_acq_rel_gpu();   // acquire: orders the observed flags before everything below (all lanes)
      }
      // request the next item's flags now; they are consumed at the top of the next iteration
      int Ln, unitn, bufn;
      have = sc.next(cur, Ln, unitn, bufn);
      f_cur = have ? flag_of(Ln, unitn) : nullptr;
      v_cur = f_cur != nullptr ? ld_relaxed_gpu(f_cur) : need;
      const CUtensorMap* tm = &maps.m[in_map];
      const uint32_t b_bytes = (uint32_t)(9 * 16 * 2) * (uint32_t)cout;
      // split mode: per 16 input channels three K chunks  x_hi.w_hi + x_lo.w_hi + x_hi.w_lo  (the packed filter holds
      // [w_hi | w_hi | w_lo] per chunk; the activation buffers hold [hi a | hi b | lo a | lo b] slabs per 16 channels)
      const int num_kc = SPLIT ? 3 * (cin >> 4) : (cin >> 4);
      for (int kc = 0; kc < num_kc; ++kc) {
        int slab = in_cs0 + kc * 2;
        if constexpr (SPLIT) {
          const int c16 = kc / 3, t = kc - 3 * c16;
          slab = 2 * in_cs0 + 4 * c16 + (t == 1 ? 2 : 0);
        }
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one_sync()) {
          if (p.debug & 4) {
            mbar_arrive(&full[s]);
          } else {
            // order the acquired flags (generic proxy) before the TMA reads (async proxy)
            asm volatile("fence.proxy.async.global;" ::: "memory");
            mbar_arrive_expect_tx(&full[s], (uint32_t)kTABytes + b_bytes);
            tma_load_4d(smA + s * kTABytes, tm, &full[s], (tx * kTW - 1) * 8, ty * kTH - 1, slab,
                        (p.debug & 256) ? 0 : n);
            bulk_load(smB + s * kTBBytesMax, wp + (size_t)kc * (b_bytes / 2), b_bytes, &full[s]);
          }
        }
        __syncwarp();
        if (++s == kTStages) { s = 0; ph ^= 1; }
      }
      L = Ln; unit = unitn; buf = bufn;
    }
  } else if (warp == 1) {
    // ================= MMA issuer: converged warp, one elected lane issues =================
    const uint32_t a_hi = desc_hi(kHW * 16);
    const uint32_t b_hi = desc_hi(128);
    const uint32_t smA_u = smem_u32(smA), smB_u = smem_u32(smB);
    int s = 0;
    uint32_t ph = 0;
    uint32_t use[2] = {0, 0};   // how often each accumulator buffer was handed over (mbarrier phases)
    Cursor cur;
    int L, unit, buf;
    while (sc.next(cur, L, unit, buf)) {
      const int4 li = linfo[L];
      const int cin = li.x, cout = li.y, mode = li.z >> 8;
      const int num_kc = SPLIT ? 3 * (cin >> 4) : (cin >> 4);
      const long long t0 = p.prof ? clock64() : 0;
      mbar_wait(&tempty[buf], (use[buf] & 1) ^ 1);
      ++use[buf];
      tc_fence_after();
      const long long t1 = p.prof ? clock64() : 0;
      long long tw = 0;
      // a tail accumulates onto columns [32, 64) of every sub-tile, where its head left the partial sums
      const uint32_t d0 = tmem_base + (uint32_t)(buf * 256 + (mode == 2 ? 32 : 0));
      for (int kc = 0; kc < num_kc; ++kc) {
        const long long tw0 = p.prof ? clock64() : 0;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (p.prof) tw += clock64() - tw0;
        const uint32_t a_lo = desc_lo(smA_u + s * kTABytes, kHW * kHH * 16);
        const uint32_t acc0 = (kc != 0 || mode == 2) ? 1u : 0u;
        if (cout == 32) {
          const uint32_t b_lo = desc_lo(smB_u + s * kTBBytesMax, 4 * 128);
          if (elect_one_sync()) {
            issue_stage_wide<32>(d0, a_lo, a_hi, b_lo, b_hi, acc0);
            umma_commit(&empty[s]);
            if (kc == num_kc - 1) umma_commit(&tfull[buf]);
          }
        } else {
          const uint32_t b_lo = desc_lo(smB_u + s * kTBBytesMax, 8 * 128);
          if (elect_one_sync()) {
            issue_stage_wide<64>(d0, a_lo, a_hi, b_lo, b_hi, acc0);
            umma_commit(&empty[s]);
            if (kc == num_kc - 1) umma_commit(&tfull[buf]);
          }
        }
        __syncwarp();
        if (++s == kTStages) { s = 0; ph ^= 1; }
      }
      if (p.prof && lane == 0) {
        atomicAdd(p.prof + L * 8 + 0, (unsigned long long)(t1 - t0));       // waiting for a free accumulator
        atomicAdd(p.prof + L * 8 + 1, (unsigned long long)tw);              // waiting for operands
        atomicAdd(p.prof + L * 8 + 2, (unsigned long long)(clock64() - t0)); // whole item at the issuer
        atomicAdd(p.prof + L * 8 + 3, 1ull);
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> HBM, then publish the unit =================
    // Two groups of four warps (warp w may touch TMEM lanes 32*(w%4)..+31). Group e owns TMEM accumulator buffer e
    // and handles the CTA's items on that buffer, so the epilogues of consecutive items (residual loads, stores, the
    // release fence) overlap each other. The accumulator is handed back to the MMA warp as soon as its last column
    // block is in registers.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int m = 32 * q + lane;
    const int gy = m >> 3, xr = m & 7;
    const bool mem = !(p.debug & 2);
    const bool mem_f32 = mem && !(p.debug & 32), mem_bf16 = mem && !(p.debug & 64), mem_res = mem && !(p.debug & 128);
    const size_t plane = (size_t)p.H * p.W;
    float* sbias = sbias_all + (warp - 2) * 64;
    uint32_t use = 0;
    Cursor cur;
    int L, item, buf;
    while (sc.next(cur, L, item, buf)) {
      if (buf != grp) continue;
      const int n_img = item / per_img;
      const int r = item - n_img * per_img;
      const int n = (p.debug & 256) ? 0 : n_img;   // tuning: every image aliases image 0 (L2-resident working set)
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const TrunkLayer ly = p.layers[L];
      const int y0 = ty * kTH + gy;
      const int x0 = tx * kTW + xr;
      // 32-column blocks per sub-tile that take the epilogue now: a pair head leaves its upper 32 columns in TMEM,
      // a pair tail reads exactly those
      const int nbc = ly.mode == 0 ? (ly.cout >> 5) : 1;
      const int col_off = ly.mode == 2 ? 32 : 0;
      const int nblk = 4 * nbc;                // sub-tile-major
      const int cmain = nbc << 5;
      F8 r1[4], r1n[4];
      // this pass's bias, staged per warp in shared memory (a global read per block would cost an
      // L2 round trip each: the gpu-scope fences keep invalidating L1)
      if (lane < cmain) sbias[lane] = __ldg(ly.bias + lane);
      if (lane + 32 < cmain) sbias[lane + 32] = __ldg(ly.bias + lane + 32);
      // Addends (residual stream) were written by the same unit of earlier passes;
      // (L-1, item) complete implies all of those are (dependencies are transitive), and this warp
      // must acquire that flag itself: the producer warp's acquire is only inherited through tfull,
      // which has not been waited on yet.
      if (L > 0 && (ly.res1 || ly.res2) && !(p.debug & 1)) {
        if (lane == 0) {
          const unsigned int* f = p.done + (size_t)(L - 1) * I + item;
          SpinGuard guard;
          while (ld_acquire_gpu(f) < need) {
            if (guard.expired()) {
              printf("dbm: trunk epilogue dependency timeout layer %d item %d\n", L, item);
              __trap();
            }
            __nanosleep(32);
          }
        }
      }
      __syncwarp();
      // res1 of block b is requested one block ahead (block 0: before the accumulator is complete)
      auto load_r1 = [&](int b, F8 (&dst)[4]) {
        const int j = nbc == 2 ? (b >> 1) : b;
        const int c0 = (b - j * nbc) << 5;
        const int x = x0 + 8 * (j & 1), y = y0 + 16 * (j >> 1);
        if (ly.res1 != nullptr && mem_res && y < p.H && x < p.W) {
          const float* rp =
              ly.res1 + (((size_t)n * (ly.res1_cs_total >> 1) + (c0 >> 3)) * plane + (size_t)y * p.W + x) * 8;
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) dst[s8] = ld_cg_f8(rp + (size_t)s8 * plane * 8);
        }
      };
      load_r1(0, r1);
      const long long e0 = p.prof ? clock64() : 0;
      mbar_wait(&tfull[grp], use & 1);
      ++use;
      tc_fence_after();
      const long long e1 = p.prof ? clock64() : 0;
      long long e2 = 0;
      // Passes whose only output is a 32-channel bf16 feature slot (every dense-block conv_1..4: pair heads, pair
      // tails, the unpaired layers) read their sub-tiles out with two tcgen05.ld in flight per wait
      // (two sub-tiles at a time) and hand the accumulator back after the second pair of loads: the buffer is held for
      // ~1k cycles instead of the ~4k of the block-by-block loop below, which matters because in the pair schedule the next head on this very
      // buffer is the next item the MMA warp issues.
      const bool quick = nbc == 1 && ly.res1 == nullptr && ly.res2 == nullptr && ly.out_f32 == nullptr && !ly.up2 &&
                         ly.out_bf16 != nullptr;
      if (quick) {
        const uint32_t tb = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(grp * 256 + col_off);
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // two sub-tiles (64 registers) per round: 168 registers is the cap
          uint32_t acc2[2][32];
          tmem_ld_32x32b_x32(tb + (uint32_t)((2 * half) * 64), acc2[0]);
          tmem_ld_32x32b_x32(tb + (uint32_t)((2 * half + 1) * 64), acc2[1]);
          tmem_wait_ld();
          if (half == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[grp]);
            if (p.prof) e2 = clock64();
          }
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int x = x0 + 8 * jj;
            const int y = y0 + 16 * half;
            if (y < p.H && x < p.W && mem_bf16) {
              const size_t pix = (size_t)y * p.W + x;
#pragma unroll
              for (int s8 = 0; s8 < 4; ++s8) {
                const float4 ba = *reinterpret_cast<const float4*>(sbias + 8 * s8);
                const float4 bb = *reinterpret_cast<const float4*>(sbias + 8 * s8 + 4);
                float v[8] = {__uint_as_float(acc2[jj][8 * s8 + 0]) + ba.x, __uint_as_float(acc2[jj][8 * s8 + 1]) + ba.y,
                              __uint_as_float(acc2[jj][8 * s8 + 2]) + ba.z, __uint_as_float(acc2[jj][8 * s8 + 3]) + ba.w,
                              __uint_as_float(acc2[jj][8 * s8 + 4]) + bb.x, __uint_as_float(acc2[jj][8 * s8 + 5]) + bb.y,
                              __uint_as_float(acc2[jj][8 * s8 + 6]) + bb.z, __uint_as_float(acc2[jj][8 * s8 + 7]) + bb.w};
                if (ly.act) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = lrelu(v[i]);
                }
                if constexpr (SPLIT) {
                  uint4 hi, lo;
                  split_bf16x8(v, hi, lo);
                  const size_t cs = (size_t)n * (2 * ly.out_cs_total) + split_slab(ly.out_cs0 + s8);
                  *reinterpret_cast<uint4*>(ly.out_bf16 + (cs * plane + pix) * 8) = hi;
                  *reinterpret_cast<uint4*>(ly.out_bf16 + ((cs + 2) * plane + pix) * 8) = lo;
                  continue;
                }
                uint4 o;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[2], v[3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[6], v[7]);
                o.x = *reinterpret_cast<uint32_t*>(&t0);
                o.y = *reinterpret_cast<uint32_t*>(&t1);
                o.z = *reinterpret_cast<uint32_t*>(&t2);
                o.w = *reinterpret_cast<uint32_t*>(&t3);
                const size_t cs = (size_t)n * ly.out_cs_total + (ly.out_cs0 + s8);
                *reinterpret_cast<uint4*>(ly.out_bf16 + (cs * plane + pix) * 8) = o;
              }
            }
          }
        }
      } else
      for (int b = 0; b < nblk; ++b) {
        const int j = nbc == 2 ? (b >> 1) : b;
        const int c0 = (b - j * nbc) << 5;
        const int x = x0 + 8 * (j & 1);
        const int y = y0 + 16 * (j >> 1);
        const bool valid = (y < p.H) && (x < p.W);
        const size_t pix = (size_t)y * p.W + x;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(grp * 256 + j * 64 + col_off + c0), acc);
        const bool has1 = ly.res1 != nullptr && valid && mem_res, has2 = ly.res2 != nullptr && valid && mem_res;
        if (b + 1 < nblk) load_r1(b + 1, r1n);
        tmem_wait_ld();
        if (b == nblk - 1) {  // everything this item reads is in registers: release the buffer to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[grp]);
          if (p.prof) e2 = clock64();
        }
        if (valid) {
        float v[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 bb = *reinterpret_cast<const float4*>(sbias + c0 + 4 * i4);
          v[4 * i4 + 0] = __uint_as_float(acc[4 * i4 + 0]) + bb.x;
          v[4 * i4 + 1] = __uint_as_float(acc[4 * i4 + 1]) + bb.y;
          v[4 * i4 + 2] = __uint_as_float(acc[4 * i4 + 2]) + bb.z;
          v[4 * i4 + 3] = __uint_as_float(acc[4 * i4 + 3]) + bb.w;
        }
        if (has1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = r1[i >> 3].v[i & 7] + ly.beta * v[i];
        }
        if (has2) {  // RRDB skip (every third dense block): fetched in place, the item is a long one
          const float* rp = ly.res2 + (((size_t)n * 8 + (c0 >> 3)) * plane + pix) * 8;
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            const F8 rr = ld_cg_f8(rp + (size_t)s8 * plane * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * s8 + i] = rr.v[i] + ly.beta * v[8 * s8 + i];
          }
        }
        if (ly.act) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = lrelu(v[i]);
        }
        if (ly.out_f32 && mem_f32) {
          float* op = ly.out_f32 + (((size_t)n * 8 + (c0 >> 3)) * plane + pix) * 8;
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8)
            st_f8(op + (size_t)s8 * plane * 8, v[8 * s8], v[8 * s8 + 1], v[8 * s8 + 2], v[8 * s8 + 3], v[8 * s8 + 4],
                  v[8 * s8 + 5], v[8 * s8 + 6], v[8 * s8 + 7]);
        }
        if (ly.out_bf16 && mem_bf16) {
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            uint4 o, o_lo = make_uint4(0, 0, 0, 0);
            size_t cs;
            if constexpr (SPLIT) {
              split_bf16x8(v + 8 * s8, o, o_lo);
              cs = (size_t)n * (2 * ly.out_cs_total) + split_slab(ly.out_cs0 + c0 / 8 + s8);
            } else {
              __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * s8 + 0], v[8 * s8 + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * s8 + 2], v[8 * s8 + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * s8 + 4], v[8 * s8 + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * s8 + 6], v[8 * s8 + 7]);
              o.x = *reinterpret_cast<uint32_t*>(&t0);
              o.y = *reinterpret_cast<uint32_t*>(&t1);
              o.z = *reinterpret_cast<uint32_t*>(&t2);
              o.w = *reinterpret_cast<uint32_t*>(&t3);
              cs = (size_t)n * ly.out_cs_total + (ly.out_cs0 + c0 / 8 + s8);
            }
#pragma unroll 1
            for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {   // split: the lo slab lies two slabs further
              const uint4 ov = part ? o_lo : o;
              const size_t csp = cs + 2 * part;
              if (!ly.up2) {
                *reinterpret_cast<uint4*>(ly.out_bf16 + (csp * plane + pix) * 8) = ov;
              } else {
                const int Ho = 2 * p.H, Wo = 2 * p.W;
                __nv_bfloat16* base = ly.out_bf16 + ((csp * Ho + 2 * y) * Wo + 2 * x) * 8;
                *reinterpret_cast<uint4*>(base) = ov;
                *reinterpret_cast<uint4*>(base + 8) = ov;
                *reinterpret_cast<uint4*>(base + (size_t)Wo * 8) = ov;
                *reinterpret_cast<uint4*>(base + (size_t)Wo * 8 + 8) = ov;
              }
            }
          }
        }
        }
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) r1[s8] = r1n[s8];
      }
      // publish the finished unit: the warp's stores are ordered before lane 0's gpu-scope fence by
      // __syncwarp, the fence is cumulative, the flag increment follows it (release pattern)
      if (p.debug & 8) __threadfence();
      __syncwarp();
      if (lane == 0) {
        if (!(p.debug & 16)) __threadfence();
        atomicAdd(p.done + (size_t)L * I + item, 1u);
        if (p.prof && q == 0) {
          atomicAdd(p.prof + L * 8 + 4, (unsigned long long)(e1 - e0));        // waiting for the accumulator
          atomicAdd(p.prof + L * 8 + 5, (unsigned long long)(e2 - e1));        // read-out until TMEM release
          atomicAdd(p.prof + L * 8 + 6, (unsigned long long)(clock64() - e2)); // rest: stores, fence, flag
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace dbm

using namespace dbm;

namespace dbm {
int g_trunk_debug = 0;  // set through dbm_debug_set(3, mask)
unsigned long long* g_trunk_prof = nullptr;
}

extern "C" int dbm_debug_set_ptr(int key, void* ptr) {
  if (key == 1) g_trunk_prof = (unsigned long long*)ptr;
  return DBM_OK;
}

static int trunk_umma_launch(const void* layers_dev, int num_layers, int n, int h, int w, const void* stem_slab8,
                             int stem_cs_total, const void* cat_a_slab8, const void* cat_b_slab8, int cat_cs_total,
                             unsigned int* flags_dev, int split, cudaStream_t stream) {
  DBM_REQUIRE(num_layers > 0 && n > 0 && h > 0 && w > 0, "trunk: empty problem");
  DBM_REQUIRE(num_layers <= kMaxTrunkLayers, "trunk: %d passes exceed the kernel's table of %d", num_layers,
              kMaxTrunkLayers);
  DBM_REQUIRE(((uintptr_t)layers_dev & 7) == 0, "trunk: layer table must be 8-byte aligned");
  TrunkMaps maps;
  const void* bases[3] = {stem_slab8, cat_a_slab8, cat_b_slab8};
  // split buffers hold a hi and a lo slab per logical slab
  const int cs_tot[3] = {stem_cs_total << split, cat_cs_total << split, cat_cs_total << split};
  for (int i = 0; i < 3; ++i) {
    int rc = make_slab8_tmap(&maps.m[i], bases[i], n, cs_tot[i], h, w, 16, kHW, kHH);
    if (rc) return rc;
  }
  TrunkParams p;
  p.layers = (const TrunkLayer*)layers_dev;
  p.num_layers = num_layers;
  p.N = n; p.H = h; p.W = w;
  p.tiles_x = ceil_div(w, kTW); p.tiles_y = ceil_div(h, kTH);
  p.items_per_layer = n * p.tiles_x * p.tiles_y;
  p.done = flags_dev;
  p.debug = g_trunk_debug;
  p.prof = g_trunk_prof;
  DBM_CUDA(cudaMemsetAsync(flags_dev, 0, (size_t)num_layers * p.items_per_layer * sizeof(unsigned int), stream));
  const long total = p.items_per_layer;   // units per pass: the schedule deals them to the CTAs pass group by pass group
  // every CTA must be co-resident (items spin on flags set by other CTAs): the grid never exceeds what the occupancy
  // query says this device holds at once (one CTA per SM with 200+ KB of shared memory)
  int resident = 0;
  const void* kfn = split ? (const void*)umma_trunk_kernel<true> : (const void*)umma_trunk_kernel<false>;
  int rc2 = resident_ctas(kfn, kTrunkThreads, kTrunkSmem, &resident);
  if (rc2) return rc2;
  const int grid = total < resident ? (int)total : resident;
  if (split)
    umma_trunk_kernel<true><<<grid, kTrunkThreads, kTrunkSmem, stream>>>(maps, p);
  else
    umma_trunk_kernel<false><<<grid, kTrunkThreads, kTrunkSmem, stream>>>(maps, p);
  return check_launch("umma_trunk_kernel");
}

extern "C" int dbm_trunk_umma(const void* layers_dev, int num_layers, int n, int h, int w, const void* stem_slab8,
                              int stem_cs_total, const void* cat_a_slab8, const void* cat_b_slab8, int cat_cs_total,
                              unsigned int* flags_dev, cudaStream_t stream) {
  return trunk_umma_launch(layers_dev, num_layers, n, h, w, stem_slab8, stem_cs_total, cat_a_slab8, cat_b_slab8,
                           cat_cs_total, flags_dev, 0, stream);
}

// The same pass table in split-bf16 arithmetic (precision "bf16x3"): every activation and filter is carried as
// hi + lo bf16 terms and each 16-channel K chunk is contracted three times (x_hi w_hi + x_lo w_hi + x_hi w_lo; the
// dropped x_lo w_lo term is 2^-16 relative), accumulating in fp32 -- fp32-grade results on the bf16 tensor pipe at three
// times the MMA work. Buffers: bf16 slab8 with 2 x cs_total slabs, per 16 channels [hi a | hi b | lo a | lo b];
// filters packed by dbm_pack_conv3x3_table entries with mode bit 16 ([w_hi | w_hi | w_lo] per 16-channel chunk).
// Table fields (in_cs0, out_cs0, out_cs_total, cin) stay LOGICAL.
extern "C" int dbm_trunk_umma_split(const void* layers_dev, int num_layers, int n, int h, int w,
                                    const void* stem_slab8, int stem_cs_total, const void* cat_a_slab8,
                                    const void* cat_b_slab8, int cat_cs_total, unsigned int* flags_dev,
                                    cudaStream_t stream) {
  return trunk_umma_launch(layers_dev, num_layers, n, h, w, stem_slab8, stem_cs_total, cat_a_slab8, cat_b_slab8,
                           cat_cs_total, flags_dev, 1, stream);
}
