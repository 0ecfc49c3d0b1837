// Micro-benchmark: sustained rate of SS-mode tcgen05.mma (A and B from shared memory) for M = 128, K = 16 and
// N in {64, 128, 192, 256}: cycles per instruction, measured with clock64 around a long stream of MMAs issued by one
// thread per CTA (one CTA per SM).  Purpose: is the N = 64 instruction bound by the tensor pipe (32 cycles) or by the
// shared-memory operand reads ((128 + N) * 16 * 2 bytes per instruction at 128 B/clk)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../opticalflowdiffusion_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include "fd_tc.cuh"

using namespace fdtc;

void fd_set_error(const char*, ...) {}
unsigned long long g_fd_launches = 0;

template <int N, bool MN_MAJOR>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* cycles, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 128 * 128;          // A: 128 rows x 128 B, B: N rows x 128 B (K = 64)
  const uint32_t bar = b_smem + 256 * 128;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 1 && lane == 0) {
    uint32_t idesc = umma_idesc_bf16(128, N);
    if (MN_MAJOR) idesc |= (1u << 15) | (1u << 16);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t ad, bd;
        if (MN_MAJOR) {
          ad = umma_desc_sw128(a_smem + k * 2048) | ((uint64_t)(8192u >> 4) << 16);
          bd = umma_desc_sw128(b_smem + k * 2048) | ((uint64_t)(8192u >> 4) << 16);
        } else {
          ad = umma_desc_sw128(a_smem) + (uint64_t)(2 * k);
          bd = umma_desc_sw128(b_smem) + (uint64_t)(2 * k);
        }
        umma_bf16(tmem, ad, bd, idesc, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

template <int N, bool MN>
void run(const char* name) {
  const int iters = 4000, sms = 148;
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  const int smem = 1024 + (128 + 256) * 128 + 64;
  cudaFuncSetAttribute(rate_kernel<N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N, MN><<<sms, 64, smem>>>(d, 100);
  rate_kernel<N, MN><<<sms, 64, smem>>>(d, iters);
  long long h[148];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, N, cudaGetErrorString(e)); return; }
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  const double per = avg / (iters * 4.0);
  const double bytes = (128.0 + N) * 16 * 2;
  printf("%s M=128 N=%3d K=16: %6.1f cycles / MMA  (tensor-pipe minimum %5.1f, smem operand bytes %5.0f -> %5.1f B/clk)  "
         "=> %4.0f%% of the dense peak\n", name, N, per, N / 2.0, bytes, bytes / per, 100.0 * (N / 2.0) / per);
  cudaFree(d);
}

int main() {
  run<64, false>("K-major ");
  run<128, false>("K-major ");
  run<192, false>("K-major ");
  run<256, false>("K-major ");
  run<64, true>("MN-major");
  run<128, true>("MN-major");
  run<256, true>("MN-major");
  return 0;
}
