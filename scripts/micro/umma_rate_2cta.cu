// Micro-benchmark: sustained rate of SS-mode tcgen05.mma with cta_group::2 (a CTA pair on one TPC, M = 256 = 128 rows per
// CTA, each CTA holding its own A tile and HALF of B in shared memory) for N in {64, 128, 256}, K = 16.
// Companion of umma_rate.cu: does pairing lift the shared-memory operand bound of the N = 64 / N = 128 instructions?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../opticalflowdiffusion_b200/csrc umma_rate_2cta.cu -o umma_rate_2cta
#include <cstdio>
#include "fd_tc.cuh"

using namespace fdtc;

void fd_set_error(const char*, ...) {}
unsigned long long g_fd_launches = 0;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) rate2_kernel(long long* cycles, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 128 * 128;          // A: 128 rows x 128 B; B half: N/2 rows x 128 B
  const uint32_t bar = b_smem + 128 * 128;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (128 + 128) * 128 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  long long t0 = 0;
  if (warp == 1 && lane == 0) {
    t0 = clock64();
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, N);
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma2_bf16(tmem, umma_desc_sw128(a_smem) + (uint64_t)(2 * k), umma_desc_sw128(b_smem) + (uint64_t)(2 * k), idesc, 1u);
      }
      umma2_commit_mc(bar, 3);
    }
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

template <int N>
void run() {
  const int iters = 4000, sms = 148;
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  const int smem = 1024 + (128 + 128) * 128 + 64;
  cudaFuncSetAttribute(rate2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate2_kernel<N><<<sms, 64, smem>>>(d, 100);
  rate2_kernel<N><<<sms, 64, smem>>>(d, iters);
  long long h[148];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("cta_group::2 N=%d: %s\n", N, cudaGetErrorString(e)); return; }
  double avg = 0;
  for (int i = 0; i < sms; i += 2) avg += (double)h[i];
  avg /= (sms / 2);
  const double per = avg / (iters * 4.0);
  const double bytes = (128.0 + N / 2.0) * 16 * 2;
  printf("cta_group::2 M=256 N=%3d K=16: %6.1f cycles / MMA  (tensor-pipe minimum %5.1f per SM, smem operand bytes per CTA %5.0f "
         "-> %5.1f B/clk)  => %4.0f%% of the dense peak\n", N, per, N / 2.0, bytes, bytes / per, 100.0 * (N / 2.0) / per);
  cudaFree(d);
}

int main() {
  run<64>();
  run<128>();
  run<256>();
  return 0;
}
