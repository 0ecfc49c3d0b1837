// Attention cores of the UNet on bf16 NHWC qkv = (N, HW, 384) = [q | k | v], 4 heads x 32.
//
// LinearAttention (denoising_diffusion.py:229-243): softmax(q) over d, softmax(k) over ALL pixels,
//   ctx[d,e] = sum_n k[d,n] v[e,n] / HW,  out[e,n] = sum_d ctx[d,e] q[d,n] * 32^-0.5.
//   O(HW) and bandwidth-bound: three launches
//     (1) per pixel-chunk partial (running max, sum of exp, 32x32 context) with online rescaling,
//     (2) combine chunks -> ctx (N,4,32,32) fp32,
//     (3) apply: per pixel softmax_d(q) and the 32x32 product, ctx broadcast from shared memory.
// Attention (:256-267): softmax(q^T k / sqrt(32)) v over N = HW tokens as a streaming
//   (flash-style) kernel: S and P never leave registers, so the reference's N x N matrix (6.3 GB at
//   batch 8, 440x1024) is never materialised.  bf16 mma.sync m16n8k16 with fp32 accumulation and
//   online softmax; 1.6 % of the forward FLOPs.
#include "fd_common.cuh"

namespace {

constexpr int kHeads = 4;
constexpr int kD = 32;
constexpr int kHidden = kHeads * kD;   // 128
constexpr int kQkv = 3 * kHidden;      // 384

// ------------------------------------------------------------------------------------------------
// linear attention, pass 1
// ------------------------------------------------------------------------------------------------
constexpr int kLaTile = 32;                       // pixels per shared-memory tile
constexpr int kLaPartial = 2 * kHidden + kHeads * kD * kD;   // m[128], s[128], ctx[4][32][32]

__global__ void __launch_bounds__(256) linattn_partial_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                              float* __restrict__ partial, int HW, int chunk_px) {
  __shared__ float s_k[kLaTile][kHidden];   // raw k, then exp(k - m)
  __shared__ float s_v[kLaTile][kHidden];
  __shared__ float s_fac[kHidden];
  const int n = blockIdx.y;
  const int chunk = blockIdx.x;
  const int p_begin = chunk * chunk_px;
  const int p_end = min(HW, p_begin + chunk_px);
  const int t = threadIdx.x;
  const int head = t >> 6, d = (t & 63) >> 1, eh = t & 1;
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  float m_run = -INFINITY, s_run = 0.f;   // threads < 128: channel t
  const __nv_bfloat16* base = qkv + (long)n * HW * kQkv;
  for (int p0 = p_begin; p0 < p_end; p0 += kLaTile) {
    // load k|v of kLaTile pixels: 512 B per pixel = 32 x 16 B
#pragma unroll
    for (int it = 0; it < (kLaTile * 32) / 256; ++it) {
      const int idx = it * 256 + t;
      const int px = idx >> 5, q16 = idx & 31;
      const int p = p0 + px;
      float vals[8];
      if (p < p_end) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(base + (long)p * kQkv + kHidden) + q16);
        const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = fd_unpack_bf16(rw[e]);
          vals[2 * e] = f.x;
          vals[2 * e + 1] = f.y;
        }
      } else {
        const float fill = q16 < 16 ? -INFINITY : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[e] = fill;
      }
      float* dst = q16 < 16 ? &s_k[px][q16 * 8] : &s_v[px][(q16 - 16) * 8];
      *reinterpret_cast<float4*>(dst) = make_float4(vals[0], vals[1], vals[2], vals[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(vals[4], vals[5], vals[6], vals[7]);
    }
    __syncthreads();
    if (t < kHidden) {
      float tmax = -INFINITY;
#pragma unroll 8
      for (int px = 0; px < kLaTile; ++px) tmax = fmaxf(tmax, s_k[px][t]);
      const float m_new = fmaxf(m_run, tmax);
      const float fac = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
      float sum = 0.f;
#pragma unroll 8
      for (int px = 0; px < kLaTile; ++px) {
        const float e = __expf(s_k[px][t] - m_new);
        s_k[px][t] = e;
        sum += e;
      }
      s_run = s_run * fac + sum;
      m_run = m_new;
      s_fac[t] = fac;
    }
    __syncthreads();
    const float fac = s_fac[head * kD + d];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] *= fac;
#pragma unroll 4
    for (int px = 0; px < kLaTile; ++px) {
      const float ke = s_k[px][head * kD + d];
      const float4* vr = reinterpret_cast<const float4*>(&s_v[px][head * kD + eh * 16]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 vv = vr[q];
        acc[q * 4 + 0] += ke * vv.x;
        acc[q * 4 + 1] += ke * vv.y;
        acc[q * 4 + 2] += ke * vv.z;
        acc[q * 4 + 3] += ke * vv.w;
      }
    }
    __syncthreads();
  }
  float* out = partial + ((long)n * gridDim.x + chunk) * kLaPartial;
  if (t < kHidden) {
    out[t] = m_run;
    out[kHidden + t] = s_run;
  }
  float* c = out + 2 * kHidden + (head * kD + d) * kD + eh * 16;
#pragma unroll
  for (int e = 0; e < 16; ++e) c[e] = acc[e];
}

// pass 2: one block per (n, head), thread = (d, e)
__global__ void __launch_bounds__(1024) linattn_combine_kernel(const float* __restrict__ partial, float* __restrict__ ctx,
                                                               int nchunks, float inv_hw) {
  const int n = blockIdx.x / kHeads, head = blockIdx.x % kHeads;
  const int d = threadIdx.x >> 5, e = threadIdx.x & 31;
  const float* base = partial + (long)n * nchunks * kLaPartial;
  float M = -INFINITY;
  for (int c = 0; c < nchunks; ++c) M = fmaxf(M, base[(long)c * kLaPartial + head * kD + d]);
  float S = 0.f, acc = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const float* pc = base + (long)c * kLaPartial;
    const float mc = pc[head * kD + d];
    const float f = (mc == -INFINITY) ? 0.f : __expf(mc - M);
    S += pc[kHidden + head * kD + d] * f;
    acc += pc[2 * kHidden + (head * kD + d) * kD + e] * f;
  }
  ctx[((long)n * kHeads + head) * kD * kD + d * kD + e] = acc / S * inv_hw;
}

// pass 3: warp = (32 pixels, one head); ctx of the sample in shared memory
__global__ void __launch_bounds__(256) linattn_apply_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            const float* __restrict__ ctx,
                                                            __nv_bfloat16* __restrict__ out, int HW, float scale) {
  __shared__ __align__(16) float s_ctx[kHeads * kD * kD];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < kHeads * kD * kD; i += blockDim.x) s_ctx[i] = ctx[(long)n * kHeads * kD * kD + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = warp & 3, grp = warp >> 2;
  const float* cx = s_ctx + head * kD * kD;
  for (int p0 = (blockIdx.x * 2 + grp) * 32; p0 < HW; p0 += gridDim.x * 64) {
    const int p = p0 + lane;
    if (p >= HW) continue;
    const uint4* src = reinterpret_cast<const uint4*>(qkv + ((long)n * HW + p) * kQkv + head * kD);
    float q[kD];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 raw = __ldg(src + i);
      const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = fd_unpack_bf16(rw[e]);
        q[i * 8 + 2 * e] = f.x;
        q[i * 8 + 2 * e + 1] = f.y;
      }
    }
    float mx = q[0];
#pragma unroll
    for (int i = 1; i < kD; ++i) mx = fmaxf(mx, q[i]);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kD; ++i) {
      q[i] = __expf(q[i] - mx);
      sum += q[i];
    }
    const float norm = scale / sum;
    float acc[kD];
#pragma unroll
    for (int e = 0; e < kD; ++e) acc[e] = 0.f;
#pragma unroll
    for (int dd = 0; dd < kD; ++dd) {
      const float qd = q[dd] * norm;
      const float4* row = reinterpret_cast<const float4*>(cx + dd * kD);
#pragma unroll
      for (int e4 = 0; e4 < kD / 4; ++e4) {
        const float4 c4 = row[e4];
        acc[e4 * 4 + 0] += c4.x * qd;
        acc[e4 * 4 + 1] += c4.y * qd;
        acc[e4 * 4 + 2] += c4.z * qd;
        acc[e4 * 4 + 3] += c4.w * qd;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + ((long)n * HW + p) * kHidden + head * kD);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 o;
      o.x = fd_pack_bf16(acc[i * 8 + 0], acc[i * 8 + 1]);
      o.y = fd_pack_bf16(acc[i * 8 + 2], acc[i * 8 + 3]);
      o.z = fd_pack_bf16(acc[i * 8 + 4], acc[i * 8 + 5]);
      o.w = fd_pack_bf16(acc[i * 8 + 6], acc[i * 8 + 7]);
      dst[i] = o;
    }
  }
}

int la_chunks(int N, int HW, int* chunk_px) {
  // enough blocks for ~4 per SM, chunk a multiple of the tile
  int want = (FD_NUM_SMS * 4 + N - 1) / N;
  if (want < 1) want = 1;
  int px = (HW + want - 1) / want;
  px = ((px + kLaTile - 1) / kLaTile) * kLaTile;
  if (px < kLaTile) px = kLaTile;
  *chunk_px = px;
  return (HW + px - 1) / px;
}

// ------------------------------------------------------------------------------------------------
// full attention (flash-style), mma.sync m16n8k16 bf16
// ------------------------------------------------------------------------------------------------
constexpr int kBQ = 64, kBK = 64;
constexpr int kKStride = 40;   // bf16 per K row (32 + 8 pad): conflict-free fragment loads
constexpr int kVStride = 72;   // bf16 per V^T row (64 + 8 pad)

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                        __nv_bfloat16* __restrict__ out, int HW, float scale_log2) {
  __shared__ __align__(16) __nv_bfloat16 s_k[kBK * kKStride];
  __shared__ __align__(16) __nv_bfloat16 s_vt[kD * kVStride];
  const int n = blockIdx.z, head = blockIdx.y;
  const int q0 = blockIdx.x * kBQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* base = qkv + (long)n * HW * kQkv;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;

  uint32_t qa[2][4];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const int c0 = head * kD + kk * 16 + 2 * t;
    qa[kk][0] = row0 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row0 * kQkv + c0)) : 0u;
    qa[kk][1] = row1 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row1 * kQkv + c0)) : 0u;
    qa[kk][2] = row0 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row0 * kQkv + c0 + 8)) : 0u;
    qa[kk][3] = row1 < HW ? __ldg(reinterpret_cast<const uint32_t*>(base + (long)row1 * kQkv + c0 + 8)) : 0u;
  }
  float o[4][4];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[dt][i] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int k0 = 0; k0 < HW; k0 += kBK) {
    // stage K (row-major, padded) and V^T
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = it * 128 + threadIdx.x;     // 0..255
      const int key = idx >> 2, part = idx & 3;   // 4 x 16 B per key
      const int kg = k0 + key;
      uint4 kr = make_uint4(0, 0, 0, 0), vr = make_uint4(0, 0, 0, 0);
      if (kg < HW) {
        const __nv_bfloat16* rowp = base + (long)kg * kQkv + head * kD + part * 8;
        kr = __ldg(reinterpret_cast<const uint4*>(rowp + kHidden));
        vr = __ldg(reinterpret_cast<const uint4*>(rowp + 2 * kHidden));
      }
      *reinterpret_cast<uint4*>(&s_k[key * kKStride + part * 8]) = kr;
      const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vr);
#pragma unroll
      for (int e = 0; e < 8; ++e) s_vt[(part * 8 + e) * kVStride + key] = ve[e];
    }
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&s_k[(nt * 8 + g) * kKStride + kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&s_k[(nt * 8 + g) * kKStride + kk * 16 + 8 + 2 * t]);
        mma_bf16(s[nt], qa[kk], b0, b1);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = k0 + nt * 8 + 2 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = (key + (i & 1)) < HW;
        s[nt][i] = ok ? s[nt][i] * scale_log2 : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = (m0 == -INFINITY) ? 0.f : exp2f(m0 - mn0);
    const float c1 = (m1 == -INFINITY) ? 0.f : exp2f(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    l0 *= c0;
    l1 *= c1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      o[dt][0] *= c0;
      o[dt][1] *= c0;
      o[dt][2] *= c1;
      o[dt][3] *= c1;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mn0);
      s[nt][1] = exp2f(s[nt][1] - mn0);
      s[nt][2] = exp2f(s[nt][2] - mn1);
      s[nt][3] = exp2f(s[nt][3] - mn1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = fd_pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = fd_pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = fd_pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = fd_pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&s_vt[(dt * 8 + g) * kVStride + kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&s_vt[(dt * 8 + g) * kVStride + kk * 16 + 8 + 2 * t]);
        mma_bf16(o[dt], pa, b0, b1);
      }
    }
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __nv_bfloat16* ob = out + (long)n * HW * kHidden + head * kD;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    if (row0 < HW)
      *reinterpret_cast<uint32_t*>(ob + (long)row0 * kHidden + dt * 8 + 2 * t) = fd_pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
    if (row1 < HW)
      *reinterpret_cast<uint32_t*>(ob + (long)row1 * kHidden + dt * 8 + 2 * t) = fd_pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
  }
}

}  // namespace

extern "C" {

size_t fd_linattn_workspace_floats(int N, int HW) {
  int px;
  const int chunks = la_chunks(N, HW, &px);
  return (size_t)N * chunks * kLaPartial + (size_t)N * kHeads * kD * kD;
}

int fd_linattn(const void* qkv, void* out, float* workspace, int N, int HW, void* stream) {
  FD_REQUIRE(qkv && out && workspace && N > 0 && HW > 0, "linattn: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int px;
  const int chunks = la_chunks(N, HW, &px);
  float* partial = workspace;
  float* ctx = workspace + (size_t)N * chunks * kLaPartial;
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  linattn_partial_kernel<<<dim3(chunks, N), 256, 0, st>>>(q, partial, HW, px);
  FD_LAUNCH_CHECK();
  linattn_combine_kernel<<<N * kHeads, 1024, 0, st>>>(partial, ctx, chunks, 1.f / (float)HW);
  FD_LAUNCH_CHECK();
  int bx = (HW + 63) / 64;
  const int cap = (FD_NUM_SMS * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  linattn_apply_kernel<<<dim3(bx, N), 256, 0, st>>>(q, ctx, static_cast<__nv_bfloat16*>(out), HW, 0.17677669529663687f);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_attention(const void* qkv, void* out, int N, int HW, void* stream) {
  FD_REQUIRE(qkv && out && N > 0 && HW > 0, "attention: bad argument");
  FD_REQUIRE(N <= 65535, "attention: batch too large");
  const float scale_log2 = 0.17677669529663687f * 1.4426950408889634f;   // 32^-0.5 * log2(e)
  attention_kernel<<<dim3((HW + kBQ - 1) / kBQ, kHeads, N), 128, 0, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), HW, scale_log2);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

}  // extern "C"
