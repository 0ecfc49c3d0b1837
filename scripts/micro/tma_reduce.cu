// Micro-test: cp.reduce.async.bulk.tensor.3d ... .add (fp32, no swizzle) from shared memory into a global tensor, with in-bounds,
// partially out-of-bounds and negative box origins.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../opticalflowdiffusion_b200/csrc -lcuda
#include <cstdio>
#include <vector>
#include <cuda.h>
#include "fd_tc.cuh"
using namespace fdtc;
thread_local char g_err[256];
void fd_set_error(const char* fmt, ...) {}

constexpr int BW = 36, BH = 5;      // box
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int mode) {
  __shared__ __align__(128) float box[BH * BW];
  for (int i = threadIdx.x; i < BH * BW; i += blockDim.x) box[i] = 1.0f + i * 0.001f;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode == 0)
      asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map),
                   "r"(smem_u32(box)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map),
                   "r"(smem_u32(box)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    tma_store_commit();
    tma_store_wait_all();
  }
}
int main() {
  const int W = 64, H = 24, P = 6;
  float* d;
  cudaMalloc(&d, sizeof(float) * W * H * P);
  CUtensorMap map;
  const uint64_t dims[3] = {W, H, P};
  const uint64_t str[2] = {W * 4, (uint64_t)W * H * 4};
  const uint32_t box[3] = {BW, BH, 1};
  if (make_tmap_f32_plain(&map, d, 3, dims, str, box)) { printf("map failed\n"); return 1; }
  const int cases[][4] = {{0, 0, 0, 1}, {0, 0, 0, 0}, {4, 2, 1, 0}, {40, 22, 2, 0}, {-4, -2, 3, 0}, {-4, -2, 3, 1}};
  for (auto& c : cases) {
    cudaMemset(d, 0, sizeof(float) * W * H * P);
    k<<<1, 128>>>(map, c[0], c[1], c[2], c[3]);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(W * H * P);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    double s = 0; int nz = 0;
    for (float v : h) { s += v; nz += v != 0.f; }
    printf("%s at (%d,%d,%d): %s  nonzero %d sum %.3f\n", c[3] ? "store " : "reduce", c[0], c[1], c[2], cudaGetErrorString(e), nz, s);
    if (e != cudaSuccess) return 2;
  }
  return 0;
}
