// micro-benchmark: cost of cooperative_groups grid.sync() on this GPU for a given grid
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void k(int n, int* out) {
  cg::grid_group g = cg::this_grid();
  int acc = 0;
  for (int i = 0; i < n; ++i) { g.sync(); acc += i; }
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = acc;
}
int main() {
  int* out; cudaMalloc(&out, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int per_sm : {1, 2, 4, 8}) {
    int grid = sms * per_sm, n = 1000;
    void* args[] = {&n, &out};
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(128), args, 0, 0);
    cudaEventRecord(a);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(128), args, 0, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("grid %d (%d/SM): %.2f us per grid.sync (%s)\n", grid, per_sm, ms * 1e3 / n, cudaGetErrorString(e));
  }
  return 0;
}
