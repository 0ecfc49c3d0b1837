// Logging reductions of FlowDiffuser.training_step / validation_step (flow_diffuser.py:221-232, 262-282): for a tensor
// x (B, inner) the four scalars the reference logs with four to seven eager reductions each --
//   torch.min(x), torch.max(x), torch.mean(x), torch.mean(torch.std(x, dim=0))      (std: unbiased, over the batch axis)
// -- in ONE pass over x: a thread owns `inner` positions (grid-stride), reads the B values of a position (coalesced
// across the warp, the second read for the centred sum of squares comes from L1/L2), and the block partials are reduced
// in a fixed order in double precision by a second tiny launch: deterministic, no atomics, no host synchronisation.
// NaN semantics as torch: any NaN makes min, max and mean NaN; B = 1 gives std = NaN (0/0).
#include "fd_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) tensor_stats_kernel(const float* __restrict__ x, int B, long inner,
                                                                float* __restrict__ partials) {
  __shared__ float red[5 * 32];
  float mn = INFINITY, mx = -INFINITY, sum = 0.f, sd = 0.f, nan = 0.f;
  const float inv_b = 1.f / (float)B, inv_bm1 = 1.f / (float)(B - 1);        // B = 1: inf * 0 = NaN, as torch.std
  for (long p = (long)blockIdx.x * kThreads + threadIdx.x; p < inner; p += (long)gridDim.x * kThreads) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float v = __ldg(x + (long)b * inner + p);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
      if (v != v) nan = 1.f;
      s += v;
    }
    const float mean = s * inv_b;
    float m2 = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = __ldg(x + (long)b * inner + p) - mean;
      m2 = fmaf(d, d, m2);
    }
    sum += s;
    sd += sqrtf(m2 * inv_bm1);
  }
  // min / max through the sum helper's shuffle pattern: reduce them separately
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[32], smx[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  float v3[3] = {sum, sd, nan};
  fd_block_sum<3>(v3, red);                      // (contains the __syncthreads that publishes smn / smx)
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    float* o = partials + (long)blockIdx.x * 5;
    o[0] = mn; o[1] = mx; o[2] = v3[0]; o[3] = v3[1]; o[4] = v3[2];
  }
}

__global__ void __launch_bounds__(32) tensor_stats_finalize_kernel(const float* __restrict__ partials, int nblocks, double count,
                                                                   double inner, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  float mn = INFINITY, mx = -INFINITY;
  double sum = 0.0, sd = 0.0, nan = 0.0;
  for (int i = 0; i < nblocks; ++i) {            // fixed order
    const float* p = partials + (long)i * 5;
    mn = fminf(mn, p[0]);
    mx = fmaxf(mx, p[1]);
    sum += (double)p[2];
    sd += (double)p[3];
    nan += (double)p[4];
  }
  const float qnan = __int_as_float(0x7fc00000);
  out[0] = nan > 0.0 ? qnan : mn;
  out[1] = nan > 0.0 ? qnan : mx;
  out[2] = (float)(sum / count);
  out[3] = (float)(sd / inner);
}

int stats_grid(long inner) {
  long blocks = (inner + kThreads - 1) / kThreads;
  const long cap = (long)FD_NUM_SMS * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

size_t fd_tensor_stats_workspace_floats(long inner) { return (size_t)stats_grid(inner) * 5; }

int fd_tensor_stats(const float* x, int B, long inner, float* out4, float* workspace, void* stream) {
  FD_REQUIRE(x && out4 && workspace && B > 0 && inner > 0, "tensor_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = stats_grid(inner);
  tensor_stats_kernel<<<grid, kThreads, 0, st>>>(x, B, inner, workspace);
  FD_LAUNCH_CHECK();
  tensor_stats_finalize_kernel<<<1, 32, 0, st>>>(workspace, grid, (double)B * (double)inner, (double)inner, out4);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

}  // extern "C"
