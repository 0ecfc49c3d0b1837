// Loss kernels of the FlowLearner objective (algorithms/diffusion_animation/flow_learner.py:133-222),
// fp32 NCHW:
//
//  * soft_charb: one (level, offset) term of the multi-scale photometric loss.  Given the raw "soft" splats
//      S = splat(cat(img * e^w, e^w), flow)      and      T = splat(cat(tgt * e, e), 0)          (B, C+1, h, w)
//    it fuses softsplat's normalisation (softsplat_new.py:316-331: x / (norm + 1e-7)), fill_holes_nan (warp.py:273-276:
//    NaN where the splatted weight is not > 0) and nan_charbonnier (warp.py:281-287: mean over the non-NaN pairs of
//    sqrt((a - b)^2 + 1e-6)) into one reduction, and its backward into one pass that yields dL/dS.
//  * edge_smooth: edgeaware_smoothness1 (warp.py:289-303) and its gradient with respect to the flow.
#include "fd_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kNormEps = 0.0000001f;
constexpr float kCharbEps2 = 1e-6f;      // charbonnier(x, alpha = 0.5, eps = 1e-3): (x^2 + eps^2)^0.5

int lgrid(long items) {
  long b = (items + kThreads - 1) / kThreads;
  if (b > FD_NUM_SMS * 8) b = FD_NUM_SMS * 8;
  return (int)(b < 1 ? 1 : b);
}

__global__ void __launch_bounds__(kThreads) soft_charb_fwd_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                                 float* __restrict__ partials, int B, int C, long HW) {
  __shared__ float red[64];
  float acc[2] = {0.f, 0.f};
  const long total = (long)B * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / HW, p = i - b * HW;
    const float* s = S + b * (C + 1) * HW + p;
    const float* t = T + b * (C + 1) * HW + p;
    const float sw = s[(long)C * HW], tw = t[(long)C * HW];
    if (!(sw > 0.f)) continue;                                   // hole -> NaN -> dropped (also drops NaN weights)
    const float sn = sw + kNormEps, tn = tw + kNormEps;
    for (int c = 0; c < C; ++c) {
      const float a = t[(long)c * HW] / tn, w = s[(long)c * HW] / sn;
      const float d = a - w;
      if (d != d) continue;                                      // NaN in either operand
      acc[0] += sqrtf(d * d + kCharbEps2);
      acc[1] += 1.f;
    }
  }
  fd_block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2] = acc[0];
    partials[blockIdx.x * 2 + 1] = acc[1];
  }
}

// fixed-order final reduction in double: sums = {sum, count, sum / count}
__global__ void __launch_bounds__(256) sum2_finalize_kernel(const float* __restrict__ partials, int nblocks, float* __restrict__ sums) {
  __shared__ double sh[2][256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    a += (double)partials[2 * i];
    b += (double)partials[2 * i + 1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sums[0] = (float)sh[0][0];
    sums[1] = (float)sh[1][0];
    sums[2] = (float)(sh[0][0] / sh[1][0]);
  }
}

// K stacked terms (all offsets of one level): partials [K][gridDim.x][2]; level loss = mean_k (sum_k / count_k)
__global__ void __launch_bounds__(kThreads) soft_charb_multi_fwd_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                                       float* __restrict__ partials, int B, int C, long HW) {
  __shared__ float red[64];
  // S, T: the stacked splats in the pixel-interleaved layout of fd_splat_fwd_multi, (K, B, HW, 4) = (r, g, b, weight)
  const int k = blockIdx.y;
  const float4* Sk = reinterpret_cast<const float4*>(S) + (long)k * B * HW;
  const float4* Tk = reinterpret_cast<const float4*>(T) + (long)k * B * HW;
  float acc[2] = {0.f, 0.f};
  const long total = (long)B * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float4 s4 = __ldg(Sk + i), t4 = __ldg(Tk + i);
    const float sv[3] = {s4.x, s4.y, s4.z}, tv[3] = {t4.x, t4.y, t4.z};
    const float sw = s4.w, tw = t4.w;
    if (!(sw > 0.f)) continue;
    const float sn = sw + kNormEps, tn = tw + kNormEps;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = tv[c] / tn - sv[c] / sn;
      if (d != d) continue;
      acc[0] += sqrtf(d * d + kCharbEps2);
      acc[1] += 1.f;
    }
  }
  fd_block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    partials[((long)k * gridDim.x + blockIdx.x) * 2] = acc[0];
    partials[((long)k * gridDim.x + blockIdx.x) * 2 + 1] = acc[1];
  }
}

// one block: sums[k] = {sum_k, count_k, mean_k} for every k in a fixed order, out[0] = mean_k mean_k
__global__ void __launch_bounds__(256) soft_charb_multi_finalize_kernel(const float* __restrict__ partials, int nblocks, int K,
                                                                        float* __restrict__ sums, float* __restrict__ out) {
  __shared__ double sh[256];
  double local = 0.0;
  for (int k = threadIdx.x; k < K; k += 256) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nblocks; ++i) {
      a += (double)partials[((long)k * nblocks + i) * 2];
      b += (double)partials[((long)k * nblocks + i) * 2 + 1];
    }
    sums[k * 3] = (float)a;
    sums[k * 3 + 1] = (float)b;
    sums[k * 3 + 2] = (float)(a / b);
    local += a / b;
  }
  sh[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(sh[0] / (double)K);
}

__global__ void __launch_bounds__(kThreads) soft_charb_multi_bwd_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                                       const float* __restrict__ sums,
                                                                       const float* __restrict__ upstream, float* __restrict__ gS,
                                                                       int B, int C, long HW, int K) {
  const int k = blockIdx.y;
  const float g = upstream[0] / ((float)K * sums[k * 3 + 1]);
  const float4* Sk = reinterpret_cast<const float4*>(S) + (long)k * B * HW;      // interleaved (K, B, HW, 4), like gS
  const float4* Tk = reinterpret_cast<const float4*>(T) + (long)k * B * HW;
  float4* gk = reinterpret_cast<float4*>(gS) + (long)k * B * HW;
  const long total = (long)B * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float4 s4 = __ldg(Sk + i), t4 = __ldg(Tk + i);
    const float svv[3] = {s4.x, s4.y, s4.z}, tvv[3] = {t4.x, t4.y, t4.z};
    const float sw = s4.w, tw = t4.w;
    float gw = 0.f, gc[3] = {0.f, 0.f, 0.f};
    const bool live = sw > 0.f;
    const float sn = sw + kNormEps, tn = tw + kNormEps;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (live) {
        const float sv = svv[c];
        const float d = tvv[c] / tn - sv / sn;
        if (d == d) {
          const float dw = -g * d * rsqrtf(d * d + kCharbEps2);
          gc[c] = dw / sn;
          gw -= dw * sv / (sn * sn);
        }
      }
    }
    gk[i] = make_float4(gc[0], gc[1], gc[2], gw);
  }
}

// dL/dS for L = upstream * sum / count
__global__ void __launch_bounds__(kThreads) soft_charb_bwd_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                                 const float* __restrict__ sums, const float* __restrict__ upstream,
                                                                 float* __restrict__ gS, int B, int C, long HW) {
  const float g = upstream[0] / sums[1];
  const long total = (long)B * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / HW, p = i - b * HW;
    const float* s = S + b * (C + 1) * HW + p;
    const float* t = T + b * (C + 1) * HW + p;
    float* gs = gS + b * (C + 1) * HW + p;
    const float sw = s[(long)C * HW], tw = t[(long)C * HW];
    float gw = 0.f;
    const bool live = sw > 0.f;
    const float sn = sw + kNormEps, tn = tw + kNormEps;
    for (int c = 0; c < C; ++c) {
      float gc = 0.f;
      if (live) {
        const float sv = s[(long)c * HW];
        const float a = t[(long)c * HW] / tn, w = sv / sn;
        const float d = a - w;
        if (d == d) {
          const float dw = -g * d * rsqrtf(d * d + kCharbEps2);      // dL / d(warped)
          gc = dw / sn;
          gw -= dw * sv / (sn * sn);
        }
      }
      gs[(long)c * HW] = gc;
    }
    gs[(long)C * HW] = gw;
  }
}

// ---------------------------------------------------------------------------------------------
// edgeaware_smoothness1(image, flow, edge_weight = 30), warp.py:289-303
//   loss = ( mean_x( exp(-30 mean_c dIx^2) * charb(dFx) ) + mean_y( exp(-30 mean_c dIy^2) * charb(dFy) ) ) / 2
// partial sums: {sum_x, sum_y}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float edge_w(const float* __restrict__ img, long base, long step, int Ci, long HW) {
  float m = 0.f;
  for (int c = 0; c < Ci; ++c) {
    const float d = img[base + (long)c * HW + step] - img[base + (long)c * HW];
    m += d * d;
  }
  return __expf(-30.f * (m / (float)Ci));
}

__global__ void __launch_bounds__(kThreads) edge_smooth_fwd_kernel(const float* __restrict__ img, const float* __restrict__ flow,
                                                                  float* __restrict__ partials, int B, int Ci, int Cf, int H,
                                                                  int W) {
  __shared__ float red[64];
  const long HW = (long)H * W;
  float acc[2] = {0.f, 0.f};
  const long total = (long)B * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / HW, p = i - b * HW;
    const int y = (int)(p / W), x = (int)(p - (long)y * W);
    const long ib = b * Ci * HW + p, fb = b * Cf * HW + p;
    if (x + 1 < W) {
      const float w = edge_w(img, ib, 1, Ci, HW);
      for (int c = 0; c < Cf; ++c) {
        const float d = flow[fb + (long)c * HW + 1] - flow[fb + (long)c * HW];
        acc[0] += w * sqrtf(d * d + kCharbEps2);
      }
    }
    if (y + 1 < H) {
      const float w = edge_w(img, ib, W, Ci, HW);
      for (int c = 0; c < Cf; ++c) {
        const float d = flow[fb + (long)c * HW + W] - flow[fb + (long)c * HW];
        acc[1] += w * sqrtf(d * d + kCharbEps2);
      }
    }
  }
  fd_block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2] = acc[0];
    partials[blockIdx.x * 2 + 1] = acc[1];
  }
}

__global__ void __launch_bounds__(256) edge_smooth_finalize_kernel(const float* __restrict__ partials, int nblocks,
                                                                   float* __restrict__ out, double nx, double ny) {
  __shared__ double sh[2][256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    a += (double)partials[2 * i];
    b += (double)partials[2 * i + 1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)((sh[0][0] / nx + sh[1][0] / ny) * 0.5);
}

// gather form of the gradient: every flow element collects its (up to) four difference terms
__global__ void __launch_bounds__(kThreads) edge_smooth_bwd_kernel(const float* __restrict__ img, const float* __restrict__ flow,
                                                                  const float* __restrict__ upstream, float* __restrict__ gflow,
                                                                  int B, int Ci, int Cf, int H, int W, float inv_nx,
                                                                  float inv_ny) {
  const long HW = (long)H * W;
  const float g = upstream[0] * 0.5f;
  const long total = (long)B * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / HW, p = i - b * HW;
    const int y = (int)(p / W), x = (int)(p - (long)y * W);
    const long ib = b * Ci * HW + p, fb = b * Cf * HW + p;
    const float wr = x + 1 < W ? edge_w(img, ib, 1, Ci, HW) : 0.f;       // pair (x, x+1)
    const float wl = x > 0 ? edge_w(img, ib - 1, 1, Ci, HW) : 0.f;       // pair (x-1, x)
    const float wd = y + 1 < H ? edge_w(img, ib, W, Ci, HW) : 0.f;
    const float wu = y > 0 ? edge_w(img, ib - W, W, Ci, HW) : 0.f;
    for (int c = 0; c < Cf; ++c) {
      const long q = fb + (long)c * HW;
      const float f = flow[q];
      float acc = 0.f;
      if (x + 1 < W) { const float d = flow[q + 1] - f; acc -= inv_nx * wr * d * rsqrtf(d * d + kCharbEps2); }
      if (x > 0) { const float d = f - flow[q - 1]; acc += inv_nx * wl * d * rsqrtf(d * d + kCharbEps2); }
      if (y + 1 < H) { const float d = flow[q + W] - f; acc -= inv_ny * wd * d * rsqrtf(d * d + kCharbEps2); }
      if (y > 0) { const float d = f - flow[q - W]; acc += inv_ny * wu * d * rsqrtf(d * d + kCharbEps2); }
      gflow[q] = g * acc;
    }
  }
}

}  // namespace

extern "C" {

size_t fd_loss_workspace_floats(long items) { return (size_t)lgrid(items) * 2; }

int fd_soft_charb_fwd(const float* S, const float* T, float* sums, float* partials, int B, int C, int HW, void* stream) {
  FD_REQUIRE(S && T && sums && partials && B > 0 && C > 0 && HW > 0, "soft_charb_fwd: bad argument");
  const int grid = lgrid((long)B * HW);
  soft_charb_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(S, T, partials, B, C, (long)HW);
  FD_LAUNCH_CHECK();
  sum2_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, sums);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_soft_charb_bwd(const float* S, const float* T, const float* sums, const float* upstream, float* gS, int B, int C, int HW,
                      void* stream) {
  FD_REQUIRE(S && T && sums && upstream && gS && B > 0 && C > 0 && HW > 0, "soft_charb_bwd: bad argument");
  soft_charb_bwd_kernel<<<lgrid((long)B * HW), kThreads, 0, (cudaStream_t)stream>>>(S, T, sums, upstream, gS, B, C, (long)HW);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

static int multi_grid(long items, int K) {
  long b = (items + kThreads - 1) / kThreads;
  const long cap = (FD_NUM_SMS * 8 + K - 1) / K;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

size_t fd_soft_charb_multi_workspace_floats(int B, int HW, int K) { return (size_t)multi_grid((long)B * HW, K) * K * 2; }

int fd_soft_charb_multi_fwd(const float* S, const float* T, float* sums, float* out, float* partials, int B, int C, int HW, int K,
                            void* stream) {
  FD_REQUIRE(S && T && sums && out && partials && B > 0 && C > 0 && HW > 0 && K > 0 && K <= 65535, "soft_charb_multi_fwd: bad argument");
  FD_REQUIRE(C == 3, "soft_charb_multi_fwd: the stacked splats are pixel-interleaved (r, g, b, weight): C must be 3, got %d", C);
  FD_REQUIRE(((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(T)) & 15) == 0, "soft_charb_multi_fwd: 16-byte aligned S, T");
  const int grid = multi_grid((long)B * HW, K);
  soft_charb_multi_fwd_kernel<<<dim3(grid, K), kThreads, 0, (cudaStream_t)stream>>>(S, T, partials, B, C, (long)HW);
  FD_LAUNCH_CHECK();
  soft_charb_multi_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, K, sums, out);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_soft_charb_multi_bwd(const float* S, const float* T, const float* sums, const float* upstream, float* gS, int B, int C,
                            int HW, int K, void* stream) {
  FD_REQUIRE(S && T && sums && upstream && gS && B > 0 && C > 0 && HW > 0 && K > 0 && K <= 65535, "soft_charb_multi_bwd: bad argument");
  FD_REQUIRE(C == 3, "soft_charb_multi_bwd: the stacked splats are pixel-interleaved (r, g, b, weight): C must be 3, got %d", C);
  FD_REQUIRE(((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(T) | reinterpret_cast<uintptr_t>(gS)) & 15) == 0,
             "soft_charb_multi_bwd: 16-byte aligned S, T, gS");
  soft_charb_multi_bwd_kernel<<<dim3(multi_grid((long)B * HW, K), K), kThreads, 0, (cudaStream_t)stream>>>(S, T, sums, upstream, gS,
                                                                                                          B, C, (long)HW, K);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_edge_smooth_fwd(const float* img, const float* flow, float* out, float* partials, int B, int Ci, int Cf, int H, int W,
                       void* stream) {
  FD_REQUIRE(img && flow && out && partials && B > 0 && Ci > 0 && Cf > 0 && H > 1 && W > 1, "edge_smooth_fwd: bad argument");
  const int grid = lgrid((long)B * H * W);
  edge_smooth_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(img, flow, partials, B, Ci, Cf, H, W);
  FD_LAUNCH_CHECK();
  edge_smooth_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, out, (double)B * Cf * H * (W - 1),
                                                                  (double)B * Cf * (H - 1) * W);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_edge_smooth_bwd(const float* img, const float* flow, const float* upstream, float* gflow, int B, int Ci, int Cf, int H,
                       int W, void* stream) {
  FD_REQUIRE(img && flow && upstream && gflow && B > 0 && Ci > 0 && Cf > 0 && H > 1 && W > 1, "edge_smooth_bwd: bad argument");
  edge_smooth_bwd_kernel<<<lgrid((long)B * H * W), kThreads, 0, (cudaStream_t)stream>>>(
      img, flow, upstream, gflow, B, Ci, Cf, H, W, (float)(1.0 / ((double)B * Cf * H * (W - 1))),
      (float)(1.0 / ((double)B * Cf * (H - 1) * W)));
  FD_LAUNCH_CHECK();
  return FD_OK;
}

}  // extern "C"
