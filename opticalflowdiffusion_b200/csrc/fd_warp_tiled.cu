// Backward bilinear warp (warp.py:95-119) and the fused warp + Charbonnier photometric + end-point-error objective
// (losses.py:3-6,46-47), forward and backward, as TILED kernels with the sampled frame staged in shared memory.
//
// The one-thread-per-pixel-group kernels of fd_warp.cu gather every bilinear tap through L1 / L2 (ncu, round 1: 553 MB of
// L2 -> L1 traffic for 143 MB of DRAM traffic, a 3.9x amplification; 12 scattered red.global per pixel in the backward).
// Here a block owns a 16 x 128 tile of output pixels and works channel by channel:
//   * the (16 + 2 R + 1) x (128 + 2 R + 4) window of the frame plane around the tile (R = 12 pixels: 3 sigma of the config-#4
//     flow) is copied to shared memory with coalesced 128-bit loads (zero outside the image); taps that fall outside the
//     window (large flows) take the global-memory path; both paths apply the reference's validity tests;
//   * backward: the gradient of the frame is accumulated in a second shared-memory window (red.shared.add.f32) and
//     flushed once per tile and channel with 128-bit vector reductions (REDG.E.ADD.F32x4): 1/5 of the reduction
//     operations of the per-tap scatter, all of them coalesced; out-of-window taps fall back to scalar atomics.
// A thread keeps its 8 pixels' taps in compact form (cell index, the two fractions, validity bits: 3 registers per pixel)
// across the channel passes and rebuilds the four weights with the reference's op sequence, so the forward stays
// bit-identical to fd_warp.cu and to the reference (tests/test_gpu_warp.py, tests/test_gpu_headline_parity.py).
#include "fd_warp_common.cuh"

using namespace fdwarp;

namespace {

constexpr int kTH = 16, kTW = 128, kR = 12;
constexpr int kWH = kTH + 2 * kR + 1;          // 41 window rows: taps reach one row below the cell
constexpr int kWW = kTW + 2 * kR + 4;          // 156 window columns (multiple of 4)
constexpr int kWin = kWH * kWW;                // floats per window
constexpr int kThreads = 256;
constexpr int kPx = 8;                         // pixels per thread: 4 consecutive in two rows (r, r + 8)

struct Tile {
  int b, y0, x0;       // image, first row / column of the tile
};

__device__ __forceinline__ Tile tile_of(int blk, int tiles_x, int tiles_y) {
  Tile t;
  t.x0 = (blk % tiles_x) * kTW;
  const int r = blk / tiles_x;
  t.y0 = (r % tiles_y) * kTH;
  t.b = r / tiles_y;
  return t;
}

// compact tap record
struct CTap {
  int cell;            // in-window: (y0 - wy0) * kWW + (x0 - wx0);  else: y0 * W + x0 of the image (may be off-image)
  float wx, ny;        // fractional parts
  uint32_t bits;       // 0..3: okx0 okx1 oky0 oky1, 4: in window, 5: pixel exists
};

__device__ __forceinline__ void make_ctap(float fdx, float fdy, int x, int y, const BwGeom& g, int wy0, int wx0, bool exists,
                                          CTap& c, float& mask) {
  BwTaps t;
  const BwDiv dv = bw_divisors(g);
  bw_taps(fdx, fdy, x, y, g, dv, t);
  mask = bw_mask(t);
  const int ry = t.y0 - wy0, rx = t.x0 - wx0;
  const bool inwin = ry >= 0 && ry + 1 < kWH && rx >= 0 && rx + 1 < kWW;
  c.cell = inwin ? ry * kWW + rx : t.y0 * g.W + t.x0;
  c.wx = t.wx;
  c.ny = t.ny;
  c.bits = (t.okx0 ? 1u : 0u) | (t.okx1 ? 2u : 0u) | (t.oky0 ? 4u : 0u) | (t.oky1 ? 8u : 0u) | (inwin ? 16u : 0u) |
           (exists ? 32u : 0u);
}

struct W4 {
  float nw, ne, sw, se, ex, sy;
};
// the reference's weight sequence (bw_taps): ex = 1 - wx, sy = 1 - ny, products in this order
__device__ __forceinline__ W4 weights_of(const CTap& c) {
  W4 w;
  w.ex = __fsub_rn(1.f, c.wx);
  w.sy = __fsub_rn(1.f, c.ny);
  w.nw = __fmul_rn(w.sy, w.ex);
  w.ne = __fmul_rn(w.sy, c.wx);
  w.sw = __fmul_rn(c.ny, w.ex);
  w.se = __fmul_rn(c.ny, c.wx);
  return w;
}

__device__ __forceinline__ BwVals gather_of(const CTap& c, const float* __restrict__ win, const float* __restrict__ plane, int W) {
  BwVals v;
  if (c.bits & 16u) {
    // same validity tests as the global path: a cell with NO valid column / row is clamped to 0 by bw_taps and must read 0
    const float* p = win + c.cell;
    v.nw = ((c.bits & 5u) == 5u) ? p[0] : 0.f;
    v.ne = ((c.bits & 6u) == 6u) ? p[1] : 0.f;
    v.sw = ((c.bits & 9u) == 9u) ? p[kWW] : 0.f;
    v.se = ((c.bits & 10u) == 10u) ? p[kWW + 1] : 0.f;
  } else {
    const float* r0 = plane + c.cell;
    v.nw = ((c.bits & 5u) == 5u) ? __ldg(r0) : 0.f;
    v.ne = ((c.bits & 6u) == 6u) ? __ldg(r0 + 1) : 0.f;
    v.sw = ((c.bits & 9u) == 9u) ? __ldg(r0 + W) : 0.f;
    v.se = ((c.bits & 10u) == 10u) ? __ldg(r0 + W + 1) : 0.f;
  }
  return v;
}

__device__ __forceinline__ float sample_of(const BwVals& v, const W4& w) {
  float o = __fmul_rn(v.nw, w.nw);
  o = __fmaf_rn(v.ne, w.ne, o);
  o = __fmaf_rn(v.sw, w.sw, o);
  o = __fmaf_rn(v.se, w.se, o);
  return o;
}

// window of one plane -> shared memory (zero outside the image); W % 4 == 0, wx0 % 4 == 0
__device__ __forceinline__ void load_window(float* __restrict__ win, const float* __restrict__ plane, int wy0, int wx0, int H, int W) {
  for (int i = threadIdx.x; i < kWH * (kWW / 4); i += kThreads) {
    const int r = i / (kWW / 4), c4 = i - r * (kWW / 4);
    const int gy = wy0 + r, gx = wx0 + c4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(reinterpret_cast<const float4*>(plane + (long)gy * W + gx));
    *reinterpret_cast<float4*>(win + r * kWW + c4 * 4) = v;
  }
}

// accumulated window -> global plane: one 128-bit vector reduction per non-zero group of four
__device__ __forceinline__ void flush_window(const float* __restrict__ acc, float* __restrict__ plane, int wy0, int wx0, int H, int W) {
  for (int i = threadIdx.x; i < kWH * (kWW / 4); i += kThreads) {
    const int r = i / (kWW / 4), c4 = i - r * (kWW / 4);
    const int gy = wy0 + r, gx = wx0 + c4 * 4;
    if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
    const float4 v = *reinterpret_cast<const float4*>(acc + r * kWW + c4 * 4);
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(plane + (long)gy * W + gx), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
  }
}

__device__ __forceinline__ void scatter_of(const CTap& c, const W4& w, float gv, float* __restrict__ acc, float* __restrict__ plane,
                                           int W) {
  if (gv == 0.f) return;
  if (c.bits & 16u) {
    float* p = acc + c.cell;
    if ((c.bits & 5u) == 5u) atomicAdd(p, gv * w.nw);
    if ((c.bits & 6u) == 6u) atomicAdd(p + 1, gv * w.ne);
    if ((c.bits & 9u) == 9u) atomicAdd(p + kWW, gv * w.sw);
    if ((c.bits & 10u) == 10u) atomicAdd(p + kWW + 1, gv * w.se);
  } else {
    float* r0 = plane + c.cell;
    if ((c.bits & 5u) == 5u) atomicAdd(r0, gv * w.nw);
    if ((c.bits & 6u) == 6u) atomicAdd(r0 + 1, gv * w.ne);
    if ((c.bits & 9u) == 9u) atomicAdd(r0 + W, gv * w.sw);
    if ((c.bits & 10u) == 10u) atomicAdd(r0 + W + 1, gv * w.se);
  }
}

// this thread's pixel j of 8: row r (+8 for j >= 4), column x + (j & 3)
struct Px {
  int y, x;
  bool exists;
};
__device__ __forceinline__ void thread_pixels(const Tile& t, int H, int W, int& ya, int& yb, int& x) {
  const int col4 = threadIdx.x & 31, r8 = threadIdx.x >> 5;
  ya = t.y0 + r8;
  yb = t.y0 + r8 + 8;
  x = t.x0 + col4 * 4;
}

// ---------------------------------------------------------------------------------------------
// forward: out, mask  (MODE 0)   |   photometric / EPE partial sums (MODE 1)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) warp_fwd_tiled_kernel(const float* __restrict__ frame1, const float* __restrict__ frame2,
                                                                     const float* __restrict__ flow, const float* __restrict__ flow_gt,
                                                                     float* __restrict__ out, float* __restrict__ mask_out,
                                                                     float* __restrict__ partials, int C, BwGeom g, int tiles_x,
                                                                     int tiles_y) {
  extern __shared__ __align__(16) float smem[];
  float* win = smem;
  __shared__ float red[3 * 32];
  const int H = g.H, W = g.W;
  const long HW = (long)H * W;
  const Tile t = tile_of(blockIdx.x, tiles_x, tiles_y);
  const int wy0 = t.y0 - kR, wx0 = t.x0 - kR;
  int yy[2], x;
  thread_pixels(t, H, W, yy[0], yy[1], x);
  const bool colok = x < W;                       // W % 4 == 0: the four pixels exist together
  CTap ct[kPx];
  float m[kPx];
  float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const bool ex = colok && yy[h] < H;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, g0 = f0, g1 = f0;
    const long fo = (long)t.b * 2 * HW + (long)yy[h] * W + x;
    if (ex) {
      f0 = __ldg(reinterpret_cast<const float4*>(flow + fo));
      f1 = __ldg(reinterpret_cast<const float4*>(flow + fo + HW));
      if (MODE == 1) {
        g0 = __ldg(reinterpret_cast<const float4*>(flow_gt + fo));
        g1 = __ldg(reinterpret_cast<const float4*>(flow_gt + fo + HW));
      }
    }
    const float a0[4] = {f0.x, f0.y, f0.z, f0.w}, a1[4] = {f1.x, f1.y, f1.z, f1.w};
    const float b0[4] = {g0.x, g0.y, g0.z, g0.w}, b1[4] = {g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      make_ctap(a1[j], a0[j], x + j, yy[h], g, wy0, wx0, ex, ct[h * 4 + j], m[h * 4 + j]);
      if (MODE == 1 && ex) {
        const float du = a0[j] - b0[j], dv = a1[j] - b1[j];
        const float e2 = du * du + dv * dv;
        s[2] += e2 > 0.f ? e2 * rsqrtf(e2) : 0.f;
      }
    }
  }
  for (int c = 0; c < C; ++c) {
    const long po = ((long)t.b * C + c) * HW;
    __syncthreads();                                   // the previous channel's reads of the window are done
    load_window(win, frame2 + po, wy0, wx0, H, W);
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool ex = colok && yy[h] < H;
      if (!ex) continue;
      const long pix = (long)yy[h] * W + x;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const CTap& cj = ct[h * 4 + j];
        o[j] = sample_of(gather_of(cj, win, frame2 + po, W), weights_of(cj));
      }
      if (MODE == 0) {
        *reinterpret_cast<float4*>(out + po + pix) = make_float4(o[0], o[1], o[2], o[3]);
        if (mask_out != nullptr)
          *reinterpret_cast<float4*>(mask_out + po + pix) = make_float4(m[h * 4], m[h * 4 + 1], m[h * 4 + 2], m[h * 4 + 3]);
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(frame1 + po + pix));
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = av[j] - o[j];
          const float q = d * d + 1e-6f;
          s[0] += m[h * 4 + j] * (q * rsqrtf(q));      // sqrt(q), q >= 1e-6 (as photo_epe_fwd_kernel)
          s[1] += m[h * 4 + j];
        }
      }
    }
  }
  if (MODE == 1) {
    fd_block_sum<3>(s, red);
    if (threadIdx.x == 0) {
      partials[blockIdx.x * 3 + 0] = s[0];
      partials[blockIdx.x * 3 + 1] = s[1];
      partials[blockIdx.x * 3 + 2] = s[2];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  MODE 0: of sum(out * gout) (backwarp);  MODE 1: of g_photo * L_photo + g_epe * EPE (fused objective)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) warp_bwd_tiled_kernel(const float* __restrict__ frame1, const float* __restrict__ frame2,
                                                                     const float* __restrict__ flow, const float* __restrict__ flow_gt,
                                                                     const float* __restrict__ gout, const float* __restrict__ sums,
                                                                     float g_photo, float g_epe, float* __restrict__ gflow,
                                                                     float* __restrict__ gframe2, int C, BwGeom g, int tiles_x,
                                                                     int tiles_y) {
  extern __shared__ __align__(16) float smem[];
  float* win = smem;                 // frame window (gathers)
  float* acc = smem + kWin;          // gradient window (scatter), only when gframe2 != nullptr
  const int H = g.H, W = g.W;
  const long HW = (long)H * W;
  const Tile t = tile_of(blockIdx.x, tiles_x, tiles_y);
  const int wy0 = t.y0 - kR, wx0 = t.x0 - kR;
  int yy[2], x;
  thread_pixels(t, H, W, yy[0], yy[1], x);
  const bool colok = x < W;
  float kp = 0.f, ke = 0.f;
  if (MODE == 1) {
    kp = g_photo / __ldg(sums + 1);
    ke = g_epe / __ldg(sums + 3);
  }
  CTap ct[kPx];
  float m[kPx], dix[kPx], diy[kPx];
  float epe_gy[kPx], epe_gx[kPx];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const bool ex = colok && yy[h] < H;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, g0 = f0, g1 = f0;
    const long fo = (long)t.b * 2 * HW + (long)yy[h] * W + x;
    if (ex) {
      f0 = __ldg(reinterpret_cast<const float4*>(flow + fo));
      f1 = __ldg(reinterpret_cast<const float4*>(flow + fo + HW));
      if (MODE == 1) {
        g0 = __ldg(reinterpret_cast<const float4*>(flow_gt + fo));
        g1 = __ldg(reinterpret_cast<const float4*>(flow_gt + fo + HW));
      }
    }
    const float a0[4] = {f0.x, f0.y, f0.z, f0.w}, a1[4] = {f1.x, f1.y, f1.z, f1.w};
    const float b0[4] = {g0.x, g0.y, g0.z, g0.w}, b1[4] = {g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = h * 4 + j;
      make_ctap(a1[j], a0[j], x + j, yy[h], g, wy0, wx0, ex, ct[k], m[k]);
      dix[k] = diy[k] = 0.f;
      epe_gy[k] = epe_gx[k] = 0.f;
      if (MODE == 1) {
        const float du = a0[j] - b0[j], dv = a1[j] - b1[j];
        const float nrm = sqrtf(du * du + dv * dv);
        const float inv = nrm > 0.f ? ke / nrm : 0.f;
        epe_gy[k] = du * inv;
        epe_gx[k] = dv * inv;
      }
    }
  }
  for (int c = 0; c < C; ++c) {
    const long po = ((long)t.b * C + c) * HW;
    __syncthreads();
    load_window(win, frame2 + po, wy0, wx0, H, W);
    if (gframe2 != nullptr)
      for (int i = threadIdx.x; i < kWin / 4; i += kThreads) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool ex = colok && yy[h] < H;
      if (!ex) continue;
      const long pix = (long)yy[h] * W + x;
      float up[4];                                     // dL / d(warped pixel)
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {
        const float4 go = __ldg(reinterpret_cast<const float4*>(gout + po + pix));
        up[0] = go.x; up[1] = go.y; up[2] = go.z; up[3] = go.w;
      } else {
        a = __ldg(reinterpret_cast<const float4*>(frame1 + po + pix));
      }
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = h * 4 + j;
        const CTap& cj = ct[k];
        const W4 w = weights_of(cj);
        const bool need_v = MODE == 1 || gflow != nullptr;
        BwVals v;
        v.nw = v.ne = v.sw = v.se = 0.f;
        if (need_v) v = gather_of(cj, win, frame2 + po, W);
        if (MODE == 1) {
          const float d = av[j] - sample_of(v, w);
          up[j] = -kp * m[k] * d / sqrtf(d * d + 1e-6f);
        }
        if (gflow != nullptr) {
          dix[k] += up[j] * ((v.ne - v.nw) * w.sy + (v.se - v.sw) * cj.ny);
          diy[k] += up[j] * ((v.sw - v.nw) * w.ex + (v.se - v.ne) * cj.wx);
        }
        if (gframe2 != nullptr) scatter_of(cj, w, up[j], acc, gframe2 + po, W);
      }
    }
    if (gframe2 != nullptr) {
      __syncthreads();
      flush_window(acc, gframe2 + po, wy0, wx0, H, W);
    }
  }
  if (gflow != nullptr) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!(colok && yy[h] < H)) continue;
      float gy[4], gx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = h * 4 + j;
        gx[j] = ((dix[k] * g.half_w) / g.wm1n) * 2.f + epe_gx[k];
        gy[j] = ((diy[k] * g.half_h) / g.hm1n) * 2.f + epe_gy[k];
      }
      const long fo = (long)t.b * 2 * HW + (long)yy[h] * W + x;
      *reinterpret_cast<float4*>(gflow + fo) = make_float4(gy[0], gy[1], gy[2], gy[3]);
      *reinterpret_cast<float4*>(gflow + fo + HW) = make_float4(gx[0], gx[1], gx[2], gx[3]);
    }
  }
}

template <class K>
int set_smem(K kernel, int bytes) {
  FD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return FD_OK;
}

}  // namespace

// ---- launchers used by fd_warp.cu's C entry points (W % 4 == 0) ----
int fd_warp_tiles(int B, int H, int W) { return B * ((H + kTH - 1) / kTH) * ((W + kTW - 1) / kTW); }

int fd_warp_fwd_tiled(int mode, const float* frame1, const float* frame2, const float* flow, const float* flow_gt, float* out,
                      float* mask, float* partials, int B, int C, int H, int W, cudaStream_t st) {
  const BwGeom g = make_geom(H, W);
  const int tx = (W + kTW - 1) / kTW, ty = (H + kTH - 1) / kTH;
  const int smem = kWin * 4;
  static bool set = false;
  if (!set) {
    if (int e = set_smem(warp_fwd_tiled_kernel<0>, smem)) return e;
    if (int e = set_smem(warp_fwd_tiled_kernel<1>, smem)) return e;
    set = true;
  }
  if (mode == 0)
    warp_fwd_tiled_kernel<0><<<B * tx * ty, kThreads, smem, st>>>(frame1, frame2, flow, flow_gt, out, mask, partials, C, g, tx, ty);
  else
    warp_fwd_tiled_kernel<1><<<B * tx * ty, kThreads, smem, st>>>(frame1, frame2, flow, flow_gt, out, mask, partials, C, g, tx, ty);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_warp_bwd_tiled(int mode, const float* frame1, const float* frame2, const float* flow, const float* flow_gt, const float* gout,
                      const float* sums, float g_photo, float g_epe, float* gflow, float* gframe2, int B, int C, int H, int W,
                      cudaStream_t st) {
  const BwGeom g = make_geom(H, W);
  const int tx = (W + kTW - 1) / kTW, ty = (H + kTH - 1) / kTH;
  const int smem = 2 * kWin * 4;
  static bool set = false;
  if (!set) {
    if (int e = set_smem(warp_bwd_tiled_kernel<0>, smem)) return e;
    if (int e = set_smem(warp_bwd_tiled_kernel<1>, smem)) return e;
    set = true;
  }
  if (mode == 0)
    warp_bwd_tiled_kernel<0><<<B * tx * ty, kThreads, smem, st>>>(frame1, frame2, flow, flow_gt, gout, sums, g_photo, g_epe, gflow,
                                                                  gframe2, C, g, tx, ty);
  else
    warp_bwd_tiled_kernel<1><<<B * tx * ty, kThreads, smem, st>>>(frame1, frame2, flow, flow_gt, gout, sums, g_photo, g_epe, gflow,
                                                                  gframe2, C, g, tx, ty);
  FD_LAUNCH_CHECK();
  return FD_OK;
}
