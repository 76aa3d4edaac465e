// Layout of the partial-sum records the iteration kernels exchange (floats per record: PS).
#pragma once
namespace dpft {
constexpr int PS = 48;      // floats per record
constexpr int NSUM = 39;    // 21 A + 6 b + 6 corr(min) + 6 corr(max) entries that are summed
constexpr int E_VMIN = 27, E_VMAX = 28, E_CMIN = 29, E_CMAX = 35;   // slots of the extremes / corrections
}  // namespace dpft
