// Micro-benchmark / semantics probe: TS-mode tcgen05.mma (A operand read from TMEM, staged there with tcgen05.cp) against the
// SS-mode instruction on the same shared-memory tiles.
//   1. correctness: D_ss = A * B^T with both operands in shared memory (K-major, SWIZZLE_128B), D_ts = the same product with
//      A copied to TMEM by four tcgen05.cp.128x256b (one per K = 16 slice) -- compared element by element, also with the
//      A tile's start shifted by `shift` pixel rows (the strip convolution's kx taps);
//   2. rate: cycles per M = 128, N = 64, K = 16 instruction in TS mode (SS mode is operand-bound at 48 cycles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../opticalflowdiffusion_b200/csrc umma_ts.cu -o umma_ts
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "fd_tc.cuh"

using namespace fdtc;

void fd_set_error(const char*, ...) {}
unsigned long long g_fd_launches = 0;

__device__ __forceinline__ void utccp_128x256b(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int kRows = 136;     // A strip: 128 + 8 rows so that a shifted start stays inside initialised memory
constexpr int kN = 64;

// a: [kRows][64] bf16 row-major, b: [kN][64]; out_ss / out_ts: [128][kN] fp32
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, float* out_ss, float* out_ts,
                                                      int shift, long long* cycles, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base, b_smem = base + 144 * 128;
  const uint32_t bar = b_smem + 9 * kN * 128;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // canonical K-major SWIZZLE_128B: row r at r * 128 B, 16-byte granule q of the row stored at granule q ^ (r & 7)
  for (int i = threadIdx.x; i < kRows * 8; i += blockDim.x) {
    const int r = i >> 3, q = i & 7;
    *reinterpret_cast<uint4*>(g + r * 128 + ((q ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a + r * 64 + q * 8);
  }
  for (int i = threadIdx.x; i < kN * 8; i += blockDim.x) {
    const int r = i >> 3, q = i & 7;
    *reinterpret_cast<uint4*>(g + 144 * 128 + r * 128 + ((q ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * 64 + q * 8);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t d_ss = tmem, d_ts = tmem + 64, a_t = tmem + 256;
  constexpr uint32_t idesc = umma_idesc_bf16(128, kN);
  if (warp == 1 && elect_one_sync()) {
    const uint64_t ad = umma_desc_sw128(a_smem + shift * 128), bd = umma_desc_sw128(b_smem);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(d_ss, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) utccp_128x256b(a_t + 8 * k, ad + (uint64_t)(2 * k));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ts_bf16(d_ts, a_t + 8 * k, bd + (uint64_t)(2 * k), idesc, k != 0);
    umma_commit(bar);
    mbar_wait(bar, 0);
    if (iters > 0) {
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts_bf16(tmem + 128, a_t + 8 * k, bd + (uint64_t)(2 * k), idesc, 1u);
      }
      umma_commit(bar);
      mbar_wait(bar, 1);
      const long long t1 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k == 0) utccp_128x256b(a_t + 32 + 8 * (it & 3), ad + (uint64_t)(2 * (it & 3)));   // one cp per 4 MMAs
          umma_ts_bf16(tmem + 128, a_t + 8 * k, bd + (uint64_t)(2 * k), idesc, 1u);
        }
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      const long long t2 = clock64();
      // tile-shaped patterns: 36 MMAs (9 different weight tiles) + 12 copies per group
      const uint64_t bd9 = umma_desc_sw128(b_smem);
      for (int it = 0; it < iters / 4; ++it) {           // (a) 36 MMAs, then 12 copies back to back
#pragma unroll
        for (int m = 0; m < 36; ++m) umma_ts_bf16(tmem + 128, a_t + 8 * (m & 3) + 32 * ((m >> 2) % 3), bd9 + (uint64_t)(2 * (m & 3)), idesc, 1u);
#pragma unroll
        for (int c = 0; c < 12; ++c) utccp_128x256b(a_t + 96 + 8 * c, ad + (uint64_t)(2 * (c & 3)));
      }
      umma_commit(bar);
      mbar_wait(bar, 1);
      const long long t3 = clock64();
      for (int it = 0; it < iters / 4; ++it) {           // (b) one copy after every third MMA
#pragma unroll
        for (int m = 0; m < 36; ++m) {
          umma_ts_bf16(tmem + 128, a_t + 8 * (m & 3) + 32 * ((m >> 2) % 3), bd9 + (uint64_t)(2 * (m & 3)), idesc, 1u);
          if (m % 3 == 2) utccp_128x256b(a_t + 96 + 8 * (m / 3), ad + (uint64_t)(2 * ((m / 3) & 3)));
        }
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      const long long t4 = clock64();
      // (c) the strip kernel's exact addressing: accumulators at columns 0 / 64 (alternating per tile), A ring of four
      //     96-column slots from column 128, nine different 8 KB weight tiles, one copy per 3 MMAs into the spare slot
      for (int it = 0; it < iters / 4; ++it) {
        const uint32_t dcol = tmem + (it & 1) * 64;
#pragma unroll
        for (int m = 0; m < 36; ++m) {
          const int tap = m >> 2, ky = tap / 3, kx = tap % 3, k = m & 3;
          const uint32_t a_addr = tmem + 128 + ((it + ky) & 3) * 96 + kx * 32 + k * 8;
          umma_ts_bf16(dcol, a_addr, bd9 + (uint64_t)(tap * (8192 >> 4) + 2 * k), idesc, m != 0);
          if (m % 3 == 2) {
            const int c = m / 3;
            utccp_128x256b(tmem + 128 + ((it + 3) & 3) * 96 + (c >> 2) * 32 + (c & 3) * 8, ad + (uint64_t)((c >> 2) * 8 + 2 * (c & 3)));
          }
        }
      }
      umma_commit(bar);
      mbar_wait(bar, 1);
      const long long t5 = clock64();
      if (blockIdx.x == 0) printf("(c) strip-kernel addressing: %.0f cycles / tile\n", (double)(t5 - t4) / (iters / 4));
      cycles[blockIdx.x * 4] = t1 - t0;
      cycles[blockIdx.x * 4 + 1] = t2 - t1;
      cycles[blockIdx.x * 4 + 2] = t3 - t2;
      cycles[blockIdx.x * 4 + 3] = t4 - t3;
    }
  }
  __syncthreads();
  tc_fence_after();
  if (blockIdx.x == 0) {
    // 4 warps: warp w reads TMEM lanes 32 w .. 32 w + 31
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(d_ss + ((uint32_t)(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out_ss[(warp * 32 + lane) * kN + c * 32 + j] = __uint_as_float(v[j]);
      tmem_ld32(d_ts + ((uint32_t)(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out_ts[(warp * 32 + lane) * kN + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  std::vector<__nv_bfloat16> ha(kRows * 64), hb(kN * 64);
  srand(1);
  for (auto& x : ha) x = __float2bfloat16((float)(rand() % 17 - 8));
  for (auto& x : hb) x = __float2bfloat16((float)(rand() % 9 - 4));
  __nv_bfloat16 *da, *db;
  float *dss, *dts;
  long long* dcyc;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dss, 128 * kN * 4);
  cudaMalloc(&dts, 128 * kN * 4);
  cudaMalloc(&dcyc, 148 * 4 * sizeof(long long));
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 1024 + 144 * 128 + 9 * kN * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int shift = 0; shift <= 2; ++shift) {
    cudaMemset(dss, 0, 128 * kN * 4);
    cudaMemset(dts, 0, 128 * kN * 4);
    probe_kernel<<<1, 128, smem>>>(da, db, dss, dts, shift, dcyc, 0);
    std::vector<float> ss(128 * kN), ts(128 * kN);
    cudaError_t e = cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("shift %d: %s\n", shift, cudaGetErrorString(e)); return 1; }
    int bad_ss = 0, bad_ts = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < kN; ++n) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) ref += __bfloat162float(ha[(m + shift) * 64 + k]) * __bfloat162float(hb[n * 64 + k]);
        if (ss[m * kN + n] != ref) ++bad_ss;
        if (ts[m * kN + n] != ref) {
          if (bad_ts < 4) printf("  ts mismatch m=%d n=%d got %g ref %g (ss %g)\n", m, n, ts[m * kN + n], ref, ss[m * kN + n]);
          ++bad_ts;
        }
      }
    printf("shift %d: SS mismatches %d, TS mismatches %d of %d\n", shift, bad_ss, bad_ts, 128 * kN);
  }
  const int iters = 4000;
  probe_kernel<<<148, 128, smem>>>(da, db, dss, dts, 0, dcyc, 100);
  probe_kernel<<<148, 128, smem>>>(da, db, dss, dts, 0, dcyc, iters);
  long long h[592];
  cudaError_t e = cudaMemcpy(h, dcyc, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("rate: %s\n", cudaGetErrorString(e)); return 1; }
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int i = 0; i < 148; ++i) { a0 += (double)h[4 * i]; a1 += (double)h[4 * i + 1]; a2 += (double)h[4 * i + 2]; a3 += (double)h[4 * i + 3]; }
  printf("TS-mode M=128 N=%d K=16: %.1f cycles / MMA (tensor-pipe minimum %.1f; SS mode measured 48.0)\n", kN, a0 / 148 / (iters * 4.0),
         kN / 2.0);
  printf("TS-mode with one tcgen05.cp.128x256b per 4 MMAs: %.1f cycles / MMA\n", a1 / 148 / (iters * 4.0));
  printf("tile pattern (36 MMAs + 12 copies): copies back to back %.0f cycles / tile, one copy per 3 MMAs %.0f cycles / tile (36 MMAs alone = 1152)\n",
         a2 / 148 / (iters / 4), a3 / 148 / (iters / 4));
  return 0;
}
