// Pieces shared by the tcgen05 3x3-conv kernels (per-layer kernel and the persistent trunk kernel).
#pragma once
#include <utility>

#include "common.cuh"

namespace dbm {

constexpr int kTile = 16;         // a work item is a kTile x kTile pixel output unit (two M=128 MMA tiles)
constexpr int kHalo = kTile + 2;  // 18: halo tile edge

// slab8 bf16 tensor [N][CS][H][W][8] viewed as 4-D (W*8, H, CS, N); box = 18 px x 18 rows x ck/8 slabs
// (default box: the 18 x 18 halo tile of a 16 x 16 unit)
int make_slab8_tmap(CUtensorMap* tm, const void* base, int N, int CS, int H, int W, int ck, int box_w = kHalo,
                    int box_h = kHalo);

// All MMAs of one pipeline stage: 9 taps x (2 sub-tiles x CK/16 k-steps). The tap loop is kept
// rolled: fully unrolling it makes ptxas hoist all 72 descriptor words, overflow the uniform
// register file and pay R2UR.FILL / MOV.SPILL around every UTCHMMA.
template <int COUT, int CK, int DSTRIDE, int IDX>
__device__ __forceinline__ void issue_one(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t acc) {
  constexpr int KS = CK / 16;
  constexpr int j = IDX / KS, ks = IDX % KS;
  // start-address field advances in 16-byte units: no carry into the LBO field
  constexpr uint32_t a_off = (uint32_t)((2 * ks) * kHalo * kHalo + 8 * j);
  constexpr uint32_t b_off = (uint32_t)((2 * ks) * (COUT / 8) * 8);
  umma_bf16_off<a_off, b_off>(d0 + (uint32_t)(j * DSTRIDE), a_lo, a_hi, b_lo, b_hi, idesc, ks != 0 ? 1u : acc);
}
template <int COUT, int CK, int DSTRIDE, int... IDX>
__device__ __forceinline__ void issue_tap(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t acc, std::integer_sequence<int, IDX...>) {
  (issue_one<COUT, CK, DSTRIDE, IDX>(d0, a_lo, a_hi, b_lo, b_hi, idesc, acc), ...);
}
// DSTRIDE = TMEM column distance between the two sub-tile accumulators
template <int COUT, int CK, int DSTRIDE = COUT>
__device__ __forceinline__ void issue_stage_mmas(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t idesc, uint32_t acc0) {
  constexpr uint32_t kBTap = (uint32_t)((CK / 8) * (COUT / 8) * 8);  // per-tap stride of the packed weights
#pragma unroll 1
  for (uint32_t tap = 0; tap < 9; ++tap) {
    const uint32_t a_tap = a_lo + tap + (tap / 3) * (kHalo - 3);     // ky * 18 + kx
    const uint32_t b_tap = b_lo + tap * kBTap;
    issue_tap<COUT, CK, DSTRIDE>(d0, a_tap, a_hi, b_tap, b_hi, idesc, tap != 0 ? 1u : acc0,
                        std::make_integer_sequence<int, 2 * (CK / 16)>{});
  }
}

}  // namespace dbm
