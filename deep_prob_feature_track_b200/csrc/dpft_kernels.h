// Launchers shared between translation units of libdpft.
#pragma once
#include <cuda_runtime.h>

namespace dpft {

// gx, gy <- unit Sobel gradient of `planes` images of H x W (reference algorithms.py:1844-1865)
void launch_sobel_unit(const float* img, float* gx, float* gy, int planes, int H, int W, cudaStream_t stream);

}  // namespace dpft
