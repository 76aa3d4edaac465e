// Launchers shared between translation units of libdpft.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpft {

// gx, gy <- unit Sobel gradient of `planes` images of H x W (reference algorithms.py:1844-1865)
void launch_sobel_unit(const float* img, float* gx, float* gy, int planes, int H, int W, cudaStream_t stream);

// V (B,3,H,W) = [x,y,1] depth and N (B,3,H,W) its unit normals (algorithms.py:2148-2171); dmm = order-encoded
// min / max of the whole depth tensor
void launch_vertex_normal(const float* depth, const float* K, const uint32_t* dmm, float* V, float* N, int B, int H,
                          int W, cudaStream_t stream);

// Point-to-plane sums of one iteration into rec (B,28); zeroes rec first (algorithms.py:916-973)
void launch_icp_term(const float* depth0, const float* K, const float* V1, const float* N1, const float* pose,
                     const uint8_t* m0, const uint8_t* m1, float* rec, uint8_t* occ_out, float* r_out, int B, int H,
                     int W, cudaStream_t stream);

}  // namespace dpft
