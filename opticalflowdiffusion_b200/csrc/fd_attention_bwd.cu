// Backward of the two attention cores (autograd of denoising_diffusion.py:229-243 and :256-267) on
// bf16 NHWC qkv = (N, HW, 384) = [q | k | v], 4 heads x 32.
//
// LinearAttention, with  qs = softmax_d(q) * 32^-0.5,  ks = softmax_pixels(k),  ctx[d,e] = sum_p ks[d,p] v[e,p] / HW,
// out[e,p] = sum_d ctx[d,e] qs[d,p]:
//     dctx[d,e] = sum_p qs[d,p] dout[e,p]                                   (pass 1: reduction over pixels)
//     dqs[d,p]  = sum_e ctx[d,e] dout[e,p] ;   dq = scale * sm (dqs - <sm, dqs>)
//     dv[e,p]   = sum_d ks[d,p] dctx[d,e] / HW
//     dks[d,p]  = sum_e dctx[d,e] v[e,p] / HW ; dk = ks (dks - r[d]),  r[d] = sum_p ks dks = sum_e dctx[d,e] ctx[d,e]
//   so the k softmax over ALL pixels needs no second reduction pass: r comes from the two 32x32 matrices.
//   Pass 2 is one thread per (pixel, head) with the 32x32 matrices broadcast from shared memory.
//
// Attention (flash-style backward, bf16 mma.sync m16n8k16, S / P recomputed from the saved log-sum-exp):
//     D[q] = <dO[q], O[q]> ;  P = exp2(S' - lse) ;  dS = P (dO V^T - D)
//     dQ = scale dS K    (one CTA per 128 queries, streams K/V tiles)
//     dK = scale dS^T Q, dV = P^T dO   (one CTA per 128 keys, streams Q/dO tiles; works on S^T so that the
//     accumulator fragments of P^T / dS^T are directly the A fragments of the two products -- no transposes)
//   No atomics: run-to-run bit-stable.
#include "fd_mma.cuh"

using namespace fdmma;

namespace {

constexpr float kScale = 0.17677669529663687f;          // 32^-0.5
constexpr int kStatsFloats = 2 * kHidden + kHeads * kD * kD;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 p = fd_unpack_bf16(w[e]);
    f[2 * e] = p.x;
    f[2 * e + 1] = p.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = fd_pack_bf16(f[0], f[1]);
  o.y = fd_pack_bf16(f[2], f[3]);
  o.z = fd_pack_bf16(f[4], f[5]);
  o.w = fd_pack_bf16(f[6], f[7]);
  return o;
}

// ------------------------------------------------------------------------------------------------
// linear attention backward, pass 1: dctx[n][head][d][e] += sum over a pixel chunk of qs[d,p] dout[e,p]
// block = 256 threads; tile = 32 pixels staged in shared memory as fp32 (softmaxed q, dout)
// ------------------------------------------------------------------------------------------------
constexpr int kLbTile = 32;

__global__ void __launch_bounds__(256) linattn_bwd_dctx_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                               const __nv_bfloat16* __restrict__ dout,
                                                               float* __restrict__ dctx, int HW, int chunk_px) {
  __shared__ __align__(16) float s_q[kLbTile][kHidden];
  __shared__ __align__(16) float s_do[kLbTile][kHidden];
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const int p_begin = blockIdx.x * chunk_px;
  const int p_end = min(HW, p_begin + chunk_px);
  // staging role: pixel = t / 8, 16-channel slice = t % 8 (two slices = one head)
  const int spx = t >> 3, ssl = t & 7;
  // accumulation role: head, d, 16 e's
  const int head = t >> 6, d = (t & 63) >> 1, e0 = (t & 1) * 16;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int p0 = p_begin; p0 < p_end; p0 += kLbTile) {
    __syncthreads();
    {
      const int p = p0 + spx;
      float q[16], dv[16];
      if (p < p_end) {
        const __nv_bfloat16* qp = qkv + ((long)n * HW + p) * kQkv + ssl * 16;
        const __nv_bfloat16* dp = dout + ((long)n * HW + p) * kHidden + ssl * 16;
        unpack8(__ldg(reinterpret_cast<const uint4*>(qp)), q);
        unpack8(__ldg(reinterpret_cast<const uint4*>(qp + 8)), q + 8);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dp)), dv);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dp + 8)), dv + 8);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { q[j] = 0.f; dv[j] = 0.f; }
      }
      float mx = q[0];
#pragma unroll
      for (int j = 1; j < 16; ++j) mx = fmaxf(mx, q[j]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        q[j] = __expf(q[j] - mx);
        sum += q[j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      const float inv = p < p_end ? kScale / sum : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s_q[spx][ssl * 16 + j] = q[j] * inv;
        s_do[spx][ssl * 16 + j] = dv[j];
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int p = 0; p < kLbTile; ++p) {
      const float qv = s_q[p][head * kD + d];
      const float4* dr = reinterpret_cast<const float4*>(&s_do[p][head * kD + e0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = dr[j];
        acc[4 * j] += qv * v.x;
        acc[4 * j + 1] += qv * v.y;
        acc[4 * j + 2] += qv * v.z;
        acc[4 * j + 3] += qv * v.w;
      }
    }
  }
  float* dst = dctx + (((long)n * kHeads + head) * kD + d) * kD + e0;
#pragma unroll
  for (int j = 0; j < 16; ++j) atomicAdd(dst + j, acc[j]);
}

// ------------------------------------------------------------------------------------------------
// linear attention backward, pass 2: per (pixel, head) thread; block = 32 pixels x 4 heads
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) linattn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                const __nv_bfloat16* __restrict__ dout,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ dctx,
                                                                __nv_bfloat16* __restrict__ dqkv, int HW) {
  __shared__ __align__(16) float s_ctx[kHeads][kD][kD];
  __shared__ __align__(16) float s_dctx[kHeads][kD][kD];
  __shared__ float s_m[kHidden], s_iz[kHidden], s_r[kHidden];
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const float* st = stats + (long)n * kStatsFloats;
  const float inv_hw = 1.f / (float)HW;
  for (int i = t; i < kHeads * kD * kD; i += 128) {
    (&s_ctx[0][0][0])[i] = st[2 * kHidden + i];
    (&s_dctx[0][0][0])[i] = dctx[(long)n * kHeads * kD * kD + i];
  }
  s_m[t] = st[t];
  s_iz[t] = 1.f / st[kHidden + t];
  __syncthreads();
  {
    float r = 0.f;      // r[d] = sum_e dctx[d,e] ctx[d,e]
    const float* a = &s_dctx[0][0][0] + t * kD;
    const float* b = &s_ctx[0][0][0] + t * kD;
#pragma unroll
    for (int e = 0; e < kD; ++e) r += a[e] * b[e];
    s_r[t] = r;
  }
  __syncthreads();
  const int head = t >> 5, lane = t & 31;
  for (int p0 = blockIdx.x * 32; p0 < HW; p0 += gridDim.x * 32) {
    const int p = p0 + lane;
    if (p >= HW) continue;
    const __nv_bfloat16* row = qkv + ((long)n * HW + p) * kQkv + head * kD;
    __nv_bfloat16* drow = dqkv + ((long)n * HW + p) * kQkv + head * kD;
    // ---- phase A: dq
    {
      float sm[kD], dO[kD], dqs[kD];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(row + j * 8)), sm + j * 8);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dout + ((long)n * HW + p) * kHidden + head * kD + j * 8)), dO + j * 8);
      }
      float mx = sm[0];
#pragma unroll
      for (int j = 1; j < kD; ++j) mx = fmaxf(mx, sm[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < kD; ++j) {
        sm[j] = __expf(sm[j] - mx);
        sum += sm[j];
      }
      const float inv = 1.f / sum;
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < kD; ++d) {
        sm[d] *= inv;
        const float4* cr = reinterpret_cast<const float4*>(&s_ctx[head][d][0]);
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 c = cr[j];
          a += c.x * dO[4 * j] + c.y * dO[4 * j + 1] + c.z * dO[4 * j + 2] + c.w * dO[4 * j + 3];
        }
        dqs[d] = a;
        dot += sm[d] * a;
      }
#pragma unroll
      for (int d = 0; d < kD; ++d) dqs[d] = kScale * sm[d] * (dqs[d] - dot);
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(drow + j * 8) = pack8(dqs + j * 8);
    }
    // ---- phase B: dk, dv
    {
      float ks[kD], v[kD], dv[kD], dk[kD];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(row + kHidden + j * 8)), ks + j * 8);
        unpack8(__ldg(reinterpret_cast<const uint4*>(row + 2 * kHidden + j * 8)), v + j * 8);
      }
#pragma unroll
      for (int e = 0; e < kD; ++e) dv[e] = 0.f;
#pragma unroll
      for (int d = 0; d < kD; ++d) {
        const float kv = __expf(ks[d] - s_m[head * kD + d]) * s_iz[head * kD + d];
        const float kvh = kv * inv_hw;
        const float4* dr = reinterpret_cast<const float4*>(&s_dctx[head][d][0]);
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 c = dr[j];
          a += c.x * v[4 * j] + c.y * v[4 * j + 1] + c.z * v[4 * j + 2] + c.w * v[4 * j + 3];
          dv[4 * j] += kvh * c.x;
          dv[4 * j + 1] += kvh * c.y;
          dv[4 * j + 2] += kvh * c.z;
          dv[4 * j + 3] += kvh * c.w;
        }
        dk[d] = kv * (a * inv_hw - s_r[head * kD + d]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(drow + kHidden + j * 8) = pack8(dk + j * 8);
        *reinterpret_cast<uint4*>(drow + 2 * kHidden + j * 8) = pack8(dv + j * 8);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// full attention backward
// ------------------------------------------------------------------------------------------------
constexpr int kRowStride = kD + 8;                 // bf16 per staged row: 80 B, conflict-free ldmatrix
constexpr int kTile = 64;                          // streamed rows per tile
constexpr int kTileElems = kTile * kRowStride;

// D[n][head][p] = sum_d dO[p][head*32+d] * O[p][head*32+d]
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o,
                                                            const __nv_bfloat16* __restrict__ dout, float* __restrict__ D,
                                                            int N, int HW) {
  const long total = (long)N * HW * kHeads;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int head = (int)(i % kHeads);
    const long np = i / kHeads;
    const long n = np / HW, p = np - n * HW;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a[8], b[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(o + np * kHidden + head * kD + j * 8)), a);
      unpack8(__ldg(reinterpret_cast<const uint4*>(dout + np * kHidden + head * kD + j * 8)), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += a[e] * b[e];
    }
    D[(n * kHeads + head) * HW + p] = s;
  }
}

// A fragments (16 rows x 32 d) of rows row0 = base + g, row1 = row0 + 8 straight from global memory
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[2][4], const __nv_bfloat16* base, long row_stride, int row0,
                                             int row1, int HW, int col0, int tq) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const int c0 = col0 + kk * 16 + 2 * tq;
    a[kk][0] = row0 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row0 * row_stride + c0)) : 0u;
    a[kk][1] = row1 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row1 * row_stride + c0)) : 0u;
    a[kk][2] = row0 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row0 * row_stride + c0 + 8)) : 0u;
    a[kk][3] = row1 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row1 * row_stride + c0 + 8)) : 0u;
  }
}

// dQ: CTA = 128 queries (8 warps x 16), streams K / V tiles of 64 keys
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                          const __nv_bfloat16* __restrict__ dout,
                                                          const float* __restrict__ lse, const float* __restrict__ Dv,
                                                          __nv_bfloat16* __restrict__ dqkv, int HW, float scale_log2) {
  __shared__ __align__(16) __nv_bfloat16 s_k[2][kTileElems];
  __shared__ __align__(16) __nv_bfloat16 s_v[2][kTileElems];
  const int n = blockIdx.z, head = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int g = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* base = qkv + (long)n * HW * kQkv;
  const __nv_bfloat16* dbase = dout + (long)n * HW * kHidden;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  uint32_t qa[2][4], da[2][4];
  load_a_frags(qa, base, kQkv, row0, row1, HW, head * kD, tq);
  load_a_frags(da, dbase, kHidden, row0, row1, HW, head * kD, tq);
  const long sb = ((long)n * kHeads + head) * HW;
  const float lse0 = row0 < HW ? lse[sb + row0] : 0.f, lse1 = row1 < HW ? lse[sb + row1] : 0.f;
  const float D0 = row0 < HW ? Dv[sb + row0] : 0.f, D1 = row1 < HW ? Dv[sb + row1] : 0.f;
  float acc[4][4];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[dt][i] = 0.f;

  const uint32_t sk = smem_addr(&s_k[0][0]), sv = smem_addr(&s_v[0][0]);
  auto issue_tile = [&](int k0, int buf) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = it * 256 + t;
      const int key = (idx >> 2) & 63, part = idx & 3, is_v = idx >> 8;
      const int kg = k0 + key;
      const bool ok = kg < HW;
      const __nv_bfloat16* src = base + (long)(ok ? kg : 0) * kQkv + (1 + is_v) * kHidden + head * kD + part * 8;
      cp_async16((is_v ? sv : sk) + (buf * kTileElems + key * kRowStride + part * 8) * 2, src, ok);
    }
    cp_async_commit();
  };
  const int ntiles = (HW + kTile - 1) / kTile;
  issue_tile(0, 0);
  const int mi = lane >> 3, r8 = lane & 7;
  for (int ti = 0; ti < ntiles; ++ti) {
    const int buf = ti & 1, k0 = ti * kTile;
    cp_async_wait<0>();
    __syncthreads();
    if (ti + 1 < ntiles) issue_tile(k0 + kTile, buf ^ 1);
    const uint32_t kb = sk + buf * kTileElems * 2, vb = sv + buf * kTileElems * 2;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[nt][i] = dp[nt][i] = 0.f;
      uint32_t kf[4], vf[4];
      ldmatrix_x4(kf, kb + ((nt * 8 + r8) * kRowStride + mi * 8) * 2);
      ldmatrix_x4(vf, vb + ((nt * 8 + r8) * kRowStride + mi * 8) * 2);
      mma_bf16(s[nt], qa[0], kf[0], kf[1]);
      mma_bf16(s[nt], qa[1], kf[2], kf[3]);
      mma_bf16(dp[nt], da[0], vf[0], vf[1]);
      mma_bf16(dp[nt], da[1], vf[2], vf[3]);
    }
    const bool tail = k0 + kTile > HW;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = !tail || (k0 + nt * 8 + 2 * tq + (i & 1)) < HW;
        const float p = ok ? exp2f(s[nt][i] * scale_log2 - (i < 2 ? lse0 : lse1)) : 0.f;
        s[nt][i] = p * (dp[nt][i] - (i < 2 ? D0 : D1));      // dS
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = fd_pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = fd_pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = fd_pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = fd_pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      uint32_t k01[4], k23[4];
      const uint32_t ka = kb + ((kk * 16 + (mi & 1) * 8 + r8) * kRowStride + (mi >> 1) * 8) * 2;
      ldmatrix_x4_trans(k01, ka);
      ldmatrix_x4_trans(k23, ka + 16 * 2);
      mma_bf16(acc[0], pa, k01[0], k01[1]);
      mma_bf16(acc[1], pa, k01[2], k01[3]);
      mma_bf16(acc[2], pa, k23[0], k23[1]);
      mma_bf16(acc[3], pa, k23[2], k23[3]);
    }
  }
  __nv_bfloat16* ob = dqkv + (long)n * HW * kQkv + head * kD;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    if (row0 < HW)
      *reinterpret_cast<uint32_t*>(ob + (long)row0 * kQkv + dt * 8 + 2 * tq) = fd_pack_bf16(acc[dt][0] * kScale, acc[dt][1] * kScale);
    if (row1 < HW)
      *reinterpret_cast<uint32_t*>(ob + (long)row1 * kQkv + dt * 8 + 2 * tq) = fd_pack_bf16(acc[dt][2] * kScale, acc[dt][3] * kScale);
  }
}

// dK, dV: CTA = 128 keys (8 warps x 16), streams Q / dO tiles of 64 queries (+ their lse / D)
__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           const __nv_bfloat16* __restrict__ dout,
                                                           const float* __restrict__ lse, const float* __restrict__ Dv,
                                                           __nv_bfloat16* __restrict__ dqkv, int HW, float scale_log2) {
  __shared__ __align__(16) __nv_bfloat16 s_q[2][kTileElems];
  __shared__ __align__(16) __nv_bfloat16 s_do[2][kTileElems];
  __shared__ float s_lse[2][kTile], s_D[2][kTile];
  const int n = blockIdx.z, head = blockIdx.y;
  const int key0 = blockIdx.x * 128;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int g = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* base = qkv + (long)n * HW * kQkv;
  const __nv_bfloat16* dbase = dout + (long)n * HW * kHidden;
  const int row0 = key0 + warp * 16 + g, row1 = row0 + 8;
  uint32_t ka[2][4], va[2][4];
  load_a_frags(ka, base, kQkv, row0, row1, HW, kHidden + head * kD, tq);
  load_a_frags(va, base, kQkv, row0, row1, HW, 2 * kHidden + head * kD, tq);
  const long sb = ((long)n * kHeads + head) * HW;
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int i = 0; i < 4; ++i) dk[dt][i] = dv[dt][i] = 0.f;

  const uint32_t sq = smem_addr(&s_q[0][0]), sd = smem_addr(&s_do[0][0]);
  auto issue_tile = [&](int q0, int buf) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = it * 256 + t;
      const int r = (idx >> 2) & 63, part = idx & 3, is_do = idx >> 8;
      const int qg = q0 + r;
      const bool ok = qg < HW;
      const __nv_bfloat16* src = is_do ? dbase + (long)(ok ? qg : 0) * kHidden + head * kD + part * 8
                                       : base + (long)(ok ? qg : 0) * kQkv + head * kD + part * 8;
      cp_async16((is_do ? sd : sq) + (buf * kTileElems + r * kRowStride + part * 8) * 2, src, ok);
    }
    cp_async_commit();
    if (t < kTile) {
      const int qg = q0 + t;
      s_lse[buf][t] = qg < HW ? lse[sb + qg] : INFINITY;      // P = exp2(. - inf) = 0 for rows past the end
      s_D[buf][t] = qg < HW ? Dv[sb + qg] : 0.f;
    }
  };
  const int ntiles = (HW + kTile - 1) / kTile;
  issue_tile(0, 0);
  const int mi = lane >> 3, r8 = lane & 7;
  for (int ti = 0; ti < ntiles; ++ti) {
    const int buf = ti & 1, q0 = ti * kTile;
    cp_async_wait<0>();
    __syncthreads();
    if (ti + 1 < ntiles) issue_tile(q0 + kTile, buf ^ 1);
    const uint32_t qb = sq + buf * kTileElems * 2, db = sd + buf * kTileElems * 2;
    // S^T (16 keys x 64 queries) and dP^T = V dO^T
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[nt][i] = dp[nt][i] = 0.f;
      uint32_t qf[4], df[4];
      ldmatrix_x4(qf, qb + ((nt * 8 + r8) * kRowStride + mi * 8) * 2);
      ldmatrix_x4(df, db + ((nt * 8 + r8) * kRowStride + mi * 8) * 2);
      mma_bf16(s[nt], ka[0], qf[0], qf[1]);
      mma_bf16(s[nt], ka[1], qf[2], qf[3]);
      mma_bf16(dp[nt], va[0], df[0], df[1]);
      mma_bf16(dp[nt], va[1], df[2], df[3]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int qc = nt * 8 + 2 * tq;
      const float l0 = s_lse[buf][qc], l1 = s_lse[buf][qc + 1];
      const float d0 = s_D[buf][qc], d1 = s_D[buf][qc + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = exp2f(s[nt][i] * scale_log2 - ((i & 1) ? l1 : l0));
        s[nt][i] = p;                                           // P^T
        dp[nt][i] = p * (dp[nt][i] - ((i & 1) ? d1 : d0));      // dS^T
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4], sa[4];
      pa[0] = fd_pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = fd_pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = fd_pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = fd_pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      sa[0] = fd_pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
      sa[1] = fd_pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
      sa[2] = fd_pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      sa[3] = fd_pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
      uint32_t b01[4], b23[4];
      const uint32_t off = ((kk * 16 + (mi & 1) * 8 + r8) * kRowStride + (mi >> 1) * 8) * 2;
      ldmatrix_x4_trans(b01, db + off);         // dO[k = query][n = d]
      ldmatrix_x4_trans(b23, db + off + 16 * 2);
      mma_bf16(dv[0], pa, b01[0], b01[1]);
      mma_bf16(dv[1], pa, b01[2], b01[3]);
      mma_bf16(dv[2], pa, b23[0], b23[1]);
      mma_bf16(dv[3], pa, b23[2], b23[3]);
      ldmatrix_x4_trans(b01, qb + off);         // Q[k = query][n = d]
      ldmatrix_x4_trans(b23, qb + off + 16 * 2);
      mma_bf16(dk[0], sa, b01[0], b01[1]);
      mma_bf16(dk[1], sa, b01[2], b01[3]);
      mma_bf16(dk[2], sa, b23[0], b23[1]);
      mma_bf16(dk[3], sa, b23[2], b23[3]);
    }
  }
  __nv_bfloat16* ob = dqkv + (long)n * HW * kQkv + head * kD;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    const int c = dt * 8 + 2 * tq;
    if (row0 < HW) {
      *reinterpret_cast<uint32_t*>(ob + (long)row0 * kQkv + kHidden + c) = fd_pack_bf16(dk[dt][0] * kScale, dk[dt][1] * kScale);
      *reinterpret_cast<uint32_t*>(ob + (long)row0 * kQkv + 2 * kHidden + c) = fd_pack_bf16(dv[dt][0], dv[dt][1]);
    }
    if (row1 < HW) {
      *reinterpret_cast<uint32_t*>(ob + (long)row1 * kQkv + kHidden + c) = fd_pack_bf16(dk[dt][2] * kScale, dk[dt][3] * kScale);
      *reinterpret_cast<uint32_t*>(ob + (long)row1 * kQkv + 2 * kHidden + c) = fd_pack_bf16(dv[dt][2], dv[dt][3]);
    }
  }
}

}  // namespace

extern "C" {

size_t fd_linattn_bwd_workspace_floats(int N, int HW) {
  return (size_t)N * kStatsFloats + (size_t)N * kHeads * kD * kD + fd_linattn_workspace_floats(N, HW);
}

int fd_linattn_bwd(const void* qkv, const void* dout, void* dqkv, float* workspace, int N, int HW, void* stream) {
  FD_REQUIRE(qkv && dout && dqkv && workspace && N > 0 && HW > 0, "linattn_bwd: bad argument");
  FD_REQUIRE(N <= 65535, "linattn_bwd: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  float* stats = workspace;
  float* dctx = stats + (size_t)N * kStatsFloats;
  float* fwd_ws = dctx + (size_t)N * kHeads * kD * kD;
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* dO = static_cast<const __nv_bfloat16*>(dout);
  if (int e = fd_linattn_stats(q + kHidden, kQkv, stats, fwd_ws, N, HW, stream)) return e;
  FD_CUDA(cudaMemsetAsync(dctx, 0, (size_t)N * kHeads * kD * kD * sizeof(float), st));
  int want = (FD_NUM_SMS * 4) / N;
  if (want < 1) want = 1;
  int px = (HW + want - 1) / want;
  px = ((px + kLbTile - 1) / kLbTile) * kLbTile;
  const int chunks = (HW + px - 1) / px;
  linattn_bwd_dctx_kernel<<<dim3(chunks, N), 256, 0, st>>>(q, dO, dctx, HW, px);
  FD_LAUNCH_CHECK();
  int bx = (HW + 31) / 32;
  const int cap = (FD_NUM_SMS * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  linattn_bwd_apply_kernel<<<dim3(bx, N), 128, 0, st>>>(q, dO, stats, dctx, static_cast<__nv_bfloat16*>(dqkv), HW);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

size_t fd_attention_bwd_workspace_floats(int N, int HW) { return (size_t)N * kHeads * HW; }

int fd_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace,
                     int N, int HW, void* stream) {
  FD_REQUIRE(qkv && out && dout && lse && dqkv && workspace && N > 0 && HW > 0, "attention_bwd: bad argument");
  FD_REQUIRE(N <= 65535, "attention_bwd: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const float scale_log2 = kScale * 1.4426950408889634f;
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* dO = static_cast<const __nv_bfloat16*>(dout);
  float* D = workspace;
  long blocks = ((long)N * HW * kHeads + 255) / 256;
  if (blocks > FD_NUM_SMS * 16) blocks = FD_NUM_SMS * 16;
  attn_bwd_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out), dO, D, N, HW);
  FD_LAUNCH_CHECK();
  const dim3 grid((HW + 127) / 128, kHeads, N);
  attn_bwd_dq_kernel<<<grid, 256, 0, st>>>(q, dO, lse, D, static_cast<__nv_bfloat16*>(dqkv), HW, scale_log2);
  FD_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<<<grid, 256, 0, st>>>(q, dO, lse, D, static_cast<__nv_bfloat16*>(dqkv), HW, scale_log2);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

}  // extern "C"
