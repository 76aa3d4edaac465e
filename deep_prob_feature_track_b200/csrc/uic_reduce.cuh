// Warp-level end of a tile walk: the 27 (+ 12 correction) per-lane sums of a warp become ONE record of PS floats.
// Shared by the persistent kernels (uic_persistent.cu, uic_queue.cu); the launch-per-iteration kernels fold a whole
// CTA instead (reduce_and_finish in uic_forward.cu).
#pragma once
#include <cuda_runtime.h>

#include "dpft_device.cuh"
#include "dpft_records.h"
#include "uic_tile.cuh"

namespace dpft {

// `red`: this warp's [NSUM][33] shared rows; rows 27..38 already hold the sigma-extreme corrections of every lane
// (remove_tru_sigma).  Lanes are combined in fp64, in lane order; the record is written as floats.
template <bool TRU>
__device__ __forceinline__ void flush_warp(TileSums& S, float (*red)[33], float* __restrict__ rec, const int lane) {
  float wmn = 0.f, wmx = 0.f;
  if (TRU) {
    wmn = warp_min(S.vmin);
    wmx = warp_max(S.vmax);
    // only the lanes that sit on the warp's extreme keep what they collected for it
    if (S.vmin != wmn) {
#pragma unroll
      for (int i = 0; i < 6; ++i) red[27 + i][lane] = 0.f;
    }
    if (S.vmax != wmx) {
#pragma unroll
      for (int i = 0; i < 6; ++i) red[33 + i][lane] = 0.f;
    }
  }
#pragma unroll
  for (int e = 0; e < 27; ++e) red[e][lane] = S.acc[e];
  __syncwarp();
  constexpr int NE = TRU ? NSUM : 27;
  for (int e = lane; e < NE; e += 32) {
    double s = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) s += (double)red[e][j];
    rec[e < 27 ? e : e + 2] = (float)s;
  }
  if (TRU && lane == 0) {
    rec[E_VMIN] = wmn;
    rec[E_VMAX] = wmx;
  }
  __syncwarp();
}

}  // namespace dpft
