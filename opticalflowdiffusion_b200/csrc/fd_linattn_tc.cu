// Residual(PreNorm(LinearAttention)) (denoising_diffusion.py:81-87,127-135,216-244) for C in {64, 128} on the 5th-generation
// tensor cores: tcgen05.mma with TMEM accumulators, operands staged by TMA.  Two passes over x and NOTHING else through HBM
// (the unfused form moves LayerNorm(x), the 384-channel qkv tensor and the 128-channel attention output: 13x the bytes).
//
//   y = LayerNorm_g1(x);  q|k|v = W_qkv y;  q = softmax_d(q) 32^-0.5;  k = softmax_pixels(k);  v /= HW
//   ctx[d,e] = sum_p k[p,d] v[p,e]  (per head);  o[p,e] = sum_d ctx[d,e] q[p,d];  out = LayerNorm_g2(W_out o + b) + x
//
// Algebra that removes every intermediate tensor (all exact rewrites of the lines above):
//   * LayerNorm folds into the GEMMs: with W' = W diag(g1), s_j = sum_c W'[j,c],
//       (W y)[p,j] = r_p (sum_c W'[j,c] x[p,c] - mu_p s_j),   mu_p / r_p = mean / rstd of pixel p,
//     so the A operand of the q and k GEMMs is the RAW x tile as TMA lands it; the correction is one FMA in the epilogue.
//   * The softmax over ALL pixels needs no running maximum: |k[p,j]| <= ||W'_j||_2 sqrt(C) because ||y_p||_2 = sqrt(C) after
//     LayerNorm (Cauchy-Schwarz), a weight-only bound m_j.  P[p,d] = exp(k[p,d] - m_d) <= 1 never overflows and the common
//     factor cancels in the normalisation.
//   * v is never formed: sum_p P[p,d] v[p,e] = sum_c W'_v[e,c] G[d,c] with
//       G[d,c] = sum_p (P[p,d] r_p) x[p,c] - T[d],   T[d] = sum_p (P[p,d] r_p) mu_p,   den[d] = sum_p (P[p,d] r_p) sigma_p,
//     i.e. ONE tensor-core GEMM with K = pixels whose B operand is again the raw x tile (pixel rows of 64 channels are the
//     canonical MN-major SWIZZLE_128B layout, as in fd_conv_wgrad.cu); T and den ride on a 16-column GEMM against a tiny
//     tile holding (mu, sigma) split into bf16 hi + lo parts.
//   * to_out folds into the context: o W_out^T = q (ctx W_out^T) = q M, M computed once per sample (128 x C).
//
//   pass 1  linattn_ctx_kernel    x -> per-CTA partial (G, T, den)         [MMA1: k logits, MMA2: P'^T x, P'^T aux]
//           linattn_combine_kernel  partials -> M^T bf16 [N][C][128]         (tiny)
//   pass 2  linattn_apply_kernel  x -> out                                  [MMA1: q logits, MMA2: q_hat M; LayerNorm + x]
//
// Warp roles in both passes (320 threads, one CTA per SM, contiguous tile ranges inside one sample):
//   warp 0 TMA producer, warp 1 TMEM allocator + single-thread tcgen05.mma issuer, warps 2..9 epilogue (TMEM lane quarter =
//   warp % 4; the two warps of a quarter split the columns).
#include "fd_tc.cuh"

using namespace fdtc;

namespace {

constexpr int kTilePx = 128;
constexpr int kLaThreads = 320;
constexpr int kLaEpi = 256;
constexpr float kLog2e = 1.4426950408889634f;

// MN-major SWIZZLE_128B operand (see fd_conv_wgrad.cu): 128-byte rows along M/N, 8 K-rows per 1024-byte atom (SBO),
// successive 64-element M/N chunks `lbo` bytes apart
__device__ __forceinline__ uint64_t la_desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// D fp32, A / B bf16; bit 15 = A is MN-major, bit 16 = B is MN-major
__host__ __device__ constexpr uint32_t la_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float la_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void la_st16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 la_ld16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void la_tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}

// mean / rstd of one pixel row of the landed x tile (64-channel chunks of [128 px][128 B]; the swizzle permutes the 16-byte
// granules inside the row, which a sum does not care about)
template <int C>
__device__ __forceinline__ void la_row_stats(uint32_t x_tile, int row, float eps, float& mu, float& r, float& sigma) {
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int ch = 0; ch < C / 64; ++ch)
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint4 v = la_ld16(x_tile + ch * 16384 + row * 128 + g * 16);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = fd_unpack_bf16(w[e]);
        s += f.x + f.y;
        q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
      }
    }
  mu = s * (1.f / C);
  const float var = fmaxf(q * (1.f / C) - mu * mu, 0.f);
  sigma = sqrtf(var + eps);
  r = 1.f / sigma;
}

// ------------------------------------------------------------------------------------------------
// weight preparation (once per weight update): rows of to_qkv scaled by the PreNorm gain
//   wq, wk  bf16 [128][C]  = W diag(g1) log2(e)   (the softmaxes use ex2)
//   sq, sk  fp32 [128]     = row sums of the bf16-rounded rows (what the tensor core multiplies)
//   mk      fp32 [128]     = sqrt(C) ||wk_j||_2 (1 + 2^-10): upper bound of the k logit, the softmax shift
//   wv      fp32 [128][C]  = W_v diag(g1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) linattn_tc_prep_kernel(const float* __restrict__ wqkv, const float* __restrict__ g1,
                                                              __nv_bfloat16* __restrict__ wq, float* __restrict__ sq,
                                                              __nv_bfloat16* __restrict__ wk, float* __restrict__ sk,
                                                              float* __restrict__ mk, float* __restrict__ wv, int C) {
  __shared__ float red[2][4];
  const int j = blockIdx.x;            // row of to_qkv: [q 0..127 | k 128..255 | v 256..383]
  const int t = threadIdx.x;
  float s = 0.f, q2 = 0.f;
  for (int c = t; c < C; c += 128) {
    const float w = wqkv[(long)j * C + c] * g1[c];
    if (j < 256) {
      const __nv_bfloat16 b = __float2bfloat16(w * kLog2e);
      (j < 128 ? wq : wk)[(long)(j & 127) * C + c] = b;
      const float f = __bfloat162float(b);
      s += f;
      q2 = fmaf(f, f, q2);
    } else {
      wv[(long)(j - 256) * C + c] = w;
    }
  }
  s = fd_warp_sum(s);
  q2 = fd_warp_sum(q2);
  if ((t & 31) == 0) {
    red[0][t >> 5] = s;
    red[1][t >> 5] = q2;
  }
  __syncthreads();
  if (t == 0 && j < 256) {
    const float ss = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    const float qq = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    if (j < 128) {
      sq[j] = ss;
    } else {
      sk[j - 128] = ss;
      mk[j - 128] = sqrtf((float)C) * sqrtf(qq) * (1.f + 1.f / 1024.f) + 1e-3f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------------
template <int C>
struct CtxCfg {
  static constexpr int kChunks = C / 64;
  static constexpr int kXBytes = kChunks * 16384;          // x tile: 64-channel chunks of [128 px][128 B]
  static constexpr int kStages = C == 64 ? 4 : 3;
  static constexpr int kWBytes = kChunks * 16384;          // Wk' [128 rows][C], K-major, 64-channel chunks
  static constexpr int kPBytes = 32768;                    // P' [128 px][128 d]: two 64-d chunks (MN-major A of MMA2)
  static constexpr int kAuxBytes = 4096;                   // aux [16 rows][128 px] K-major: two 64-px chunks of 2 KB
  static constexpr int kOffW = kStages * kXBytes;
  static constexpr int kOffP = kOffW + kWBytes;
  static constexpr int kOffAux = kOffP + 2 * kPBytes;
  static constexpr int kOffConst = kOffAux + 2 * kAuxBytes;   // sk[128], mk[128]
  static constexpr int kOffBar = kOffConst + 1024;
  static constexpr int kSmemBytes = 1024 + kOffBar + 256;
};

struct CtxParams {
  int N, HW, cps, tiles;      // samples, pixels per sample, CTAs per sample, 128-pixel tiles per sample
  float eps;
  const float* sk;
  const float* mk;
  float* partial;             // [N * cps][128 d][C + 4]: G[d][0..C), T[d], den[d]
};

template <int C>
__global__ void __launch_bounds__(kLaThreads, 1)
linattn_ctx_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const CtxParams p) {
  using Cf = CtxCfg<C>;
  constexpr int S = Cf::kStages;
  constexpr int NCH = Cf::kChunks;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t x_smem = base, w_smem = base + Cf::kOffW, p_smem = base + Cf::kOffP, aux_smem = base + Cf::kOffAux;
  float* s_sk = reinterpret_cast<float*>(gbase + Cf::kOffConst);
  float* s_mk = s_sk + 128;
  const uint32_t bar = base + Cf::kOffBar;
  auto xfull = [&](int s) { return bar + 8u * s; };
  auto xempty = [&](int s) { return bar + 8u * (S + s); };
  const uint32_t wfull = bar + 8u * (2 * S);
  auto d1full = [&](int s) { return bar + 8u * (2 * S + 1 + s); };
  auto d1empty = [&](int s) { return bar + 8u * (2 * S + 3 + s); };
  auto pfull = [&](int s) { return bar + 8u * (2 * S + 5 + s); };
  auto pempty = [&](int s) { return bar + 8u * (2 * S + 7 + s); };
  const uint32_t accfull = bar + 8u * (2 * S + 9);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gbase + Cf::kOffBar + 8 * (2 * S + 10));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.cps, jc = blockIdx.x % p.cps;
  const int t0 = (int)((long)jc * p.tiles / p.cps), t1 = (int)((long)(jc + 1) * p.tiles / p.cps);
  const int nt = t1 - t0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < S; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(d1full(s), 1);
      mbar_init(d1empty(s), kLaEpi);
      mbar_init(pfull(s), kLaEpi);
      mbar_init(pempty(s), 1);
    }
    mbar_init(accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // TMEM columns: D1 (k logits) 2 x 128 at 0 / 128, D2 (G) C columns at 256, D3 (T, den) 16 columns at 384
  fd_grid_dependency_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(wfull, Cf::kWBytes);
      for (int ch = 0; ch < NCH; ++ch) tma_load_2d(w_smem + ch * 16384, &map_w, wfull, ch * 64, 0);
      for (int i = 0; i < nt; ++i) {
        const int s = i % S;
        mbar_wait(xempty(s), ((i / S) & 1) ^ 1u);
        mbar_expect_tx(xfull(s), Cf::kXBytes);
        for (int ch = 0; ch < NCH; ++ch)
          tma_load_3d(x_smem + s * Cf::kXBytes + ch * 16384, &map_x, xfull(s), ch * 64, (t0 + i) * kTilePx, n);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t id1 = la_idesc(128, 128, false, false);     // k logits: x (K-major) . Wk'^T (K-major)
      constexpr uint32_t id2 = la_idesc(128, C, true, true);         // G += P'^T (MN-major) . x (MN-major), K = pixels
      constexpr uint32_t id3 = la_idesc(128, 16, true, false);       // T, den += P'^T . aux^T (K-major)
      mbar_wait(wfull, 0);
      const uint64_t wdesc = umma_desc_sw128(w_smem);
      auto mma1 = [&](int i) {
        const int as = i & 1, s = i % S;
        mbar_wait(d1empty(as), ((i >> 1) & 1) ^ 1u);
        mbar_wait(xfull(s), (i / S) & 1);
        tc_fence_after();
        const uint64_t xd = umma_desc_sw128(x_smem + s * Cf::kXBytes);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + as * 128, xd + (uint64_t)(ch * 1024 + 2 * k), wdesc + (uint64_t)(ch * 1024 + 2 * k), id1,
                      (ch | k) != 0 ? 1u : 0u);
        umma_commit(d1full(as));
      };
      if (nt > 0) mma1(0);
      for (int i = 0; i < nt; ++i) {
        if (i + 1 < nt) mma1(i + 1);
        const int pb = i & 1, s = i % S;
        mbar_wait(pfull(pb), (i >> 1) & 1);
        tc_fence_after();
        const uint64_t pd = la_desc_mn(p_smem + pb * Cf::kPBytes, 16384);
        const uint64_t xd = la_desc_mn(x_smem + s * Cf::kXBytes, 16384);
        const uint64_t ad = umma_desc_sw128(aux_smem + pb * Cf::kAuxBytes);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)       // 16 pixels per step = two 8-pixel atoms = 2048 B
          umma_bf16(tmem_base + 256, pd + (uint64_t)(ks * 128), xd + (uint64_t)(ks * 128), id2, (i | ks) != 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16(tmem_base + 384, pd + (uint64_t)(ks * 128), ad + (uint64_t)((ks >> 2) * 128 + 2 * (ks & 3)), id3,
                    (i | ks) != 0 ? 1u : 0u);
        umma_commit(xempty(s));
        umma_commit(pempty(pb));
      }
      umma_commit(accfull);
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps 2..9 =====================
    const int et = threadIdx.x - 64;
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    if (et < 128) s_sk[et] = __ldg(p.sk + et);
    else s_mk[et - 128] = __ldg(p.mk + et - 128);
    for (int i = et; i < 2 * Cf::kAuxBytes / 16; i += kLaEpi) la_st16(aux_smem + i * 16, 0u, 0u, 0u, 0u);   // rows 4..15 stay zero
    named_bar_sync(1, kLaEpi);
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    for (int i = 0; i < nt; ++i) {
      const int s = i % S, pb = i & 1, as = i & 1;
      mbar_wait(xfull(s), (i / S) & 1);
      float mu, r, sigma;
      la_row_stats<C>(x_smem + s * Cf::kXBytes, row, p.eps, mu, r, sigma);
      const bool valid = (t0 + i) * kTilePx + row < p.HW;
      const float nrm = -r * mu;
      const float rv = valid ? r : 0.f;
      mbar_wait(pempty(pb), ((i >> 1) & 1) ^ 1u);
      mbar_wait(d1full(as), (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int col0 = half * 64 + cc * 32;
        uint32_t acc[32];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(as * 128 + col0), acc);
        tmem_ld_wait();
        if (cc == 1) {
          tc_fence_before();
          mbar_arrive(d1empty(as));
        }
        uint32_t pk[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 s4 = *reinterpret_cast<const float4*>(s_sk + col0 + j4 * 4);
          const float4 m4 = *reinterpret_cast<const float4*>(s_mk + col0 + j4 * 4);
          // k' - m = r acc - r mu s_j - m_j  (log2 e folded into Wk'); P' = exp2(.) r_p, zero for rows past the sample
          const float p0 = la_ex2(fmaf(r, __uint_as_float(acc[j4 * 4 + 0]), fmaf(nrm, s4.x, -m4.x))) * rv;
          const float p1 = la_ex2(fmaf(r, __uint_as_float(acc[j4 * 4 + 1]), fmaf(nrm, s4.y, -m4.y))) * rv;
          const float p2 = la_ex2(fmaf(r, __uint_as_float(acc[j4 * 4 + 2]), fmaf(nrm, s4.z, -m4.z))) * rv;
          const float p3 = la_ex2(fmaf(r, __uint_as_float(acc[j4 * 4 + 3]), fmaf(nrm, s4.w, -m4.w))) * rv;
          pk[j4 * 2] = fd_pack_bf16(p0, p1);
          pk[j4 * 2 + 1] = fd_pack_bf16(p2, p3);
        }
        const uint32_t rbase = p_smem + pb * Cf::kPBytes + half * 16384 + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t piece = (uint32_t)(cc * 4 + q) ^ (uint32_t)(row & 7);
          la_st16(rbase + piece * 16u, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      {
        // aux tile, K-major [16 rows][64 px] per 64-pixel chunk: rows 0/1 = mu hi/lo (written by half 0), 2/3 = sigma hi/lo
        const float val = half == 0 ? mu : sigma;
        const __nv_bfloat16 hi = __float2bfloat16(val);
        const __nv_bfloat16 lo = __float2bfloat16(val - __bfloat162float(hi));
        const uint32_t abase = aux_smem + pb * Cf::kAuxBytes + (row >> 6) * 2048 + (row & 7) * 2;
        const int g = (row & 63) >> 3;
        const int r0 = half * 2, r1 = half * 2 + 1;
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(abase + r0 * 128 + ((g ^ r0) & 7) * 16), "h"(*reinterpret_cast<const uint16_t*>(&hi))
                     : "memory");
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(abase + r1 * 128 + ((g ^ r1) & 7) * 16), "h"(*reinterpret_cast<const uint16_t*>(&lo))
                     : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(pfull(pb));
    }
    // ---- the CTA's accumulators -> its partial
    mbar_wait(accfull, 0);
    tc_fence_after();
    float* out = p.partial + ((long)blockIdx.x * 128 + row) * (C + 4);
#pragma unroll
    for (int cb = 0; cb < C / 64; ++cb) {
      const int col0 = half * (C / 2) + cb * 32;
      uint32_t acc[32];
      tmem_ld32(tmem_base + lane_off + (uint32_t)(256 + col0), acc);
      tmem_ld_wait();
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        *reinterpret_cast<float4*>(out + col0 + j4 * 4) =
            make_float4(__uint_as_float(acc[j4 * 4]), __uint_as_float(acc[j4 * 4 + 1]), __uint_as_float(acc[j4 * 4 + 2]),
                        __uint_as_float(acc[j4 * 4 + 3]));
    }
    if (half == 0) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + lane_off + 384u, acc);       // columns 0..3 = T hi/lo, den hi/lo parts (16..31 unused)
      tmem_ld_wait();
      *reinterpret_cast<float4*>(out + C) = make_float4(__uint_as_float(acc[0]) + __uint_as_float(acc[1]),
                                                        __uint_as_float(acc[2]) + __uint_as_float(acc[3]), 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// combine: partials of one sample -> M^T[co][d] = sum_{e in head(d)} ctx[d][e] W_out[co][e]   (bf16 [N][C][128])
//   ctx[d][e] = 32^-0.5 / (HW den[d]) sum_c (G[d][c] - T[d]) W'_v[e][c]
// one block per sample, thread d; partials are added in a fixed order (run-to-run bit-stable)
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) linattn_tc_combine_kernel(const float* __restrict__ partial, const float* __restrict__ wv,
                                                                 const float* __restrict__ wout, __nv_bfloat16* __restrict__ mt,
                                                                 int cps, int HW) {
  const int n = blockIdx.x, d = threadIdx.x;
  float g[C];
#pragma unroll
  for (int c = 0; c < C; ++c) g[c] = 0.f;
  float T = 0.f, den = 0.f;
  for (int j = 0; j < cps; ++j) {
    const float* src = partial + (((long)n * cps + j) * 128 + d) * (C + 4);
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = *reinterpret_cast<const float4*>(src + c4 * 4);
      g[c4 * 4] += v.x;
      g[c4 * 4 + 1] += v.y;
      g[c4 * 4 + 2] += v.z;
      g[c4 * 4 + 3] += v.w;
    }
    T += src[C];
    den += src[C + 1];
  }
#pragma unroll
  for (int c = 0; c < C; ++c) g[c] -= T;
  const float inv = 0.17677669529663687f / (den * (float)HW);      // 32^-0.5 (q scale, :238), 1 / HW (v, :240), softmax denominator
  const int head = d >> 5;
  float ctx[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const float* w = wv + (long)(head * 32 + e) * C;
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) a = fmaf(g[c], __ldg(w + c), a);
    ctx[e] = a * inv;
  }
  for (int co = 0; co < C; ++co) {
    const float* w = wout + (long)co * 128 + head * 32;
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) a = fmaf(ctx[e], __ldg(w + e), a);
    mt[((long)n * C + co) * 128 + d] = __float2bfloat16(a);
  }
}

// ------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------
template <int C>
struct ApTcCfg {
  static constexpr int kChunks = C / 64;
  static constexpr int kXBytes = kChunks * 16384;
  static constexpr int kStages = C == 64 ? 4 : 2;
  static constexpr int kWBytes = kChunks * 16384;          // Wq' [128 rows][C]
  static constexpr int kMBytes = 2 * C * 128;              // M^T [C rows][128 d]: two 64-d chunks of [C][128 B]
  static constexpr int kQBytes = 32768;                    // q_hat [128 px][128 d]: two 64-d chunks (K-major A of MMA2)
  static constexpr int kNQ = C == 64 ? 2 : 1;
  static constexpr int kOBytes = kChunks * 16384;          // output staging: one 64-channel slab per chunk
  static constexpr int kNO = C == 64 ? 2 : 1;
  static constexpr int kOffW = kStages * kXBytes;
  static constexpr int kOffM = kOffW + kWBytes;
  static constexpr int kOffQ = kOffM + kMBytes;
  static constexpr int kOffO = kOffQ + kNQ * kQBytes;
  static constexpr int kOffConst = kOffO + kNO * kOBytes;    // sq[128], bout[C], g2[C]  (<= 1.5 KB) | sx[2][2][128][2] (4 KB)
  static constexpr int kOffSx = kOffConst + 2048;
  static constexpr int kOffBar = kOffSx + 4096;
  static constexpr int kSmemBytes = 1024 + kOffBar + 256;
};

struct ApParams {
  int N, HW, cps, tiles;
  float eps;
  const float* sq;
  const float* bout;
  const float* g2;
};

template <int C>
__global__ void __launch_bounds__(kLaThreads, 1)
linattn_apply_kernel_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                        const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_out, const ApParams p) {
  using Cf = ApTcCfg<C>;
  constexpr int S = Cf::kStages;
  constexpr int NCH = Cf::kChunks;
  constexpr int NQ = Cf::kNQ;
  constexpr int NO = Cf::kNO;
  constexpr int NCB = C / 64;          // 32-column chunks of this thread's half of the output row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t x_smem = base, w_smem = base + Cf::kOffW, m_smem = base + Cf::kOffM, q_smem = base + Cf::kOffQ,
                 o_smem = base + Cf::kOffO;
  float* s_sq = reinterpret_cast<float*>(gbase + Cf::kOffConst);
  float* s_b = s_sq + 128;
  float* s_g2 = s_b + C;
  float2* s_sx = reinterpret_cast<float2*>(gbase + Cf::kOffSx);      // [tile parity][half][row]
  const uint32_t bar = base + Cf::kOffBar;
  auto xfull = [&](int s) { return bar + 8u * s; };
  auto xempty = [&](int s) { return bar + 8u * (S + s); };
  const uint32_t wfull = bar + 8u * (2 * S);
  auto d1full = [&](int s) { return bar + 8u * (2 * S + 1 + s); };
  auto d1empty = [&](int s) { return bar + 8u * (2 * S + 3 + s); };
  auto qfull = [&](int s) { return bar + 8u * (2 * S + 5 + s); };
  auto qempty = [&](int s) { return bar + 8u * (2 * S + 7 + s); };
  auto d2full = [&](int s) { return bar + 8u * (2 * S + 9 + s); };
  auto d2empty = [&](int s) { return bar + 8u * (2 * S + 11 + s); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gbase + Cf::kOffBar + 8 * (2 * S + 13));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.cps, jc = blockIdx.x % p.cps;
  const int t0 = (int)((long)jc * p.tiles / p.cps), t1 = (int)((long)(jc + 1) * p.tiles / p.cps);
  const int nt = t1 - t0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_m);
    tma_prefetch_desc(&map_out);
    for (int s = 0; s < S; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), kLaEpi);
    }
    mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(d1full(s), 1);
      mbar_init(d1empty(s), kLaEpi);
      mbar_init(qfull(s), kLaEpi);
      mbar_init(qempty(s), 1);
      mbar_init(d2full(s), 1);
      mbar_init(d2empty(s), kLaEpi);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // TMEM columns: D1 (q logits) 2 x 128 at 0 / 128, D2 (q_hat M) 2 x C at 256 / 384
  fd_grid_dependency_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(wfull, Cf::kWBytes + Cf::kMBytes);
      for (int ch = 0; ch < NCH; ++ch) tma_load_2d(w_smem + ch * 16384, &map_w, wfull, ch * 64, 0);
      for (int dc = 0; dc < 2; ++dc) tma_load_3d(m_smem + dc * (C * 128), &map_m, wfull, dc * 64, 0, n);
      for (int i = 0; i < nt; ++i) {
        const int s = i % S;
        mbar_wait(xempty(s), ((i / S) & 1) ^ 1u);
        mbar_expect_tx(xfull(s), Cf::kXBytes);
        for (int ch = 0; ch < NCH; ++ch)
          tma_load_3d(x_smem + s * Cf::kXBytes + ch * 16384, &map_x, xfull(s), ch * 64, (t0 + i) * kTilePx, n);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t id1 = la_idesc(128, 128, false, false);
      constexpr uint32_t id2 = la_idesc(128, C, false, false);
      mbar_wait(wfull, 0);
      const uint64_t wdesc = umma_desc_sw128(w_smem), mdesc = umma_desc_sw128(m_smem);
      auto mma1 = [&](int i) {
        const int as = i & 1, s = i % S;
        mbar_wait(d1empty(as), ((i >> 1) & 1) ^ 1u);
        mbar_wait(xfull(s), (i / S) & 1);
        tc_fence_after();
        const uint64_t xd = umma_desc_sw128(x_smem + s * Cf::kXBytes);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + as * 128, xd + (uint64_t)(ch * 1024 + 2 * k), wdesc + (uint64_t)(ch * 1024 + 2 * k), id1,
                      (ch | k) != 0 ? 1u : 0u);
        umma_commit(d1full(as));
      };
      if (nt > 0) mma1(0);
      for (int i = 0; i < nt; ++i) {
        if (i + 1 < nt) mma1(i + 1);
        const int qb = i % NQ, as = i & 1;
        mbar_wait(qfull(qb), (i / NQ) & 1);
        mbar_wait(d2empty(as), ((i >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint64_t qd = umma_desc_sw128(q_smem + qb * Cf::kQBytes);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)       // K = 128 d: two 64-d chunks x four 16-element steps
          umma_bf16(tmem_base + 256 + as * 128, qd + (uint64_t)((ks >> 2) * 1024 + 2 * (ks & 3)),
                    mdesc + (uint64_t)((ks >> 2) * (C * 128 >> 4) + 2 * (ks & 3)), id2, ks != 0 ? 1u : 0u);
        umma_commit(d2full(as));
        umma_commit(qempty(qb));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps 2..9 =====================
    const int ew = warp - 2, et = threadIdx.x - 64;
    const int quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    for (int i = et; i < 128 + 2 * C; i += kLaEpi)
      s_sq[i] = i < 128 ? __ldg(p.sq + i) : (i < 128 + C ? __ldg(p.bout + i - 128) : __ldg(p.g2 + i - 128 - C));
    named_bar_sync(1, kLaEpi);
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t slab_count = 0;

    auto epi_a = [&](int i) {       // q logits -> softmax over d per head -> q_hat tile
      const int s = i % S, qb = i % NQ, as = i & 1;
      mbar_wait(xfull(s), (i / S) & 1);
      float mu, r, sigma;
      la_row_stats<C>(x_smem + s * Cf::kXBytes, row, p.eps, mu, r, sigma);
      const float nrm = -r * mu;
      mbar_wait(qempty(qb), ((i / NQ) & 1) ^ 1u);
      mbar_wait(d1full(as), (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col0 = (half * 2 + hh) * 32;        // one head = 32 columns = one TMEM load
        uint32_t acc[32];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(as * 128 + col0), acc);
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          mbar_arrive(d1empty(as));
        }
        float v[32];
        float mx = -INFINITY;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 s4 = *reinterpret_cast<const float4*>(s_sq + col0 + j4 * 4);
          v[j4 * 4 + 0] = fmaf(r, __uint_as_float(acc[j4 * 4 + 0]), nrm * s4.x);
          v[j4 * 4 + 1] = fmaf(r, __uint_as_float(acc[j4 * 4 + 1]), nrm * s4.y);
          v[j4 * 4 + 2] = fmaf(r, __uint_as_float(acc[j4 * 4 + 2]), nrm * s4.z);
          v[j4 * 4 + 3] = fmaf(r, __uint_as_float(acc[j4 * 4 + 3]), nrm * s4.w);
          mx = fmaxf(mx, fmaxf(fmaxf(v[j4 * 4], v[j4 * 4 + 1]), fmaxf(v[j4 * 4 + 2], v[j4 * 4 + 3])));
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = la_ex2(v[j] - mx);
          sum += v[j];
        }
        const float inv = __fdividef(1.f, sum);
        const uint32_t rbase = q_smem + qb * Cf::kQBytes + half * 16384 + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t piece = (uint32_t)(hh * 4 + q) ^ (uint32_t)(row & 7);
          la_st16(rbase + piece * 16u, fd_pack_bf16(v[q * 8] * inv, v[q * 8 + 1] * inv), fd_pack_bf16(v[q * 8 + 2] * inv, v[q * 8 + 3] * inv),
                  fd_pack_bf16(v[q * 8 + 4] * inv, v[q * 8 + 5] * inv), fd_pack_bf16(v[q * 8 + 6] * inv, v[q * 8 + 7] * inv));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(qfull(qb));
    };

    auto epi_b = [&](int i) {       // q_hat M + b -> LayerNorm_g2 -> + x -> bf16 -> TMA store
      const int s = i % S, as = i & 1;
      mbar_wait(d2full(as), (i >> 1) & 1);
      tc_fence_after();
      float o[NCB * 32];
      float sm = 0.f, sq2 = 0.f;
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        const int col0 = half * (C / 2) + cb * 32;
        uint32_t acc[32];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(256 + as * 128 + col0), acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(acc[j]) + s_b[col0 + j];
          o[cb * 32 + j] = x;
          sm += x;
          sq2 = fmaf(x, x, sq2);
        }
      }
      tc_fence_before();
      mbar_arrive(d2empty(as));
      float2* sx = s_sx + (i & 1) * 256;
      sx[half * 128 + row] = make_float2(sm, sq2);
      const uint32_t ob = slab_count % NO;
      ++slab_count;
      if (ew == 0 && elect_one_sync()) tma_store_wait_read<NO - 1>();      // the store that last used this staging buffer has read it
      named_bar_sync(2, kLaEpi);
      const float2 other = sx[(half ^ 1) * 128 + row];
      const float mean = (sm + other.x) * (1.f / C);
      const float var = fmaxf((sq2 + other.y) * (1.f / C) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        const int col0 = half * (C / 2) + cb * 32;       // absolute output channel of o[cb * 32]
        const int chunk = col0 >> 6, g0 = (col0 & 63) >> 3;
        const uint32_t xrow = x_smem + s * Cf::kXBytes + chunk * 16384 + row * 128;
        const uint32_t orow = o_smem + ob * Cf::kOBytes + chunk * 16384 + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t piece = (uint32_t)(g0 + q) ^ (uint32_t)(row & 7);
          const uint4 xr = la_ld16(xrow + piece * 16u);
          const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
          uint32_t ow[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = fd_unpack_bf16(xw[e]);
            const int j = cb * 32 + q * 8 + e * 2;
            const float y0 = fmaf((o[j] - mean) * rstd, s_g2[col0 + q * 8 + e * 2], f.x);
            const float y1 = fmaf((o[j + 1] - mean) * rstd, s_g2[col0 + q * 8 + e * 2 + 1], f.y);
            ow[e] = fd_pack_bf16(y0, y1);
          }
          la_st16(orow + piece * 16u, ow[0], ow[1], ow[2], ow[3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(xempty(s));              // this thread's reads of the x tile (statistics in epi_a, residual here) are done
      named_bar_sync(3, kLaEpi);
      if (ew == 0 && elect_one_sync()) {
        for (int ch = 0; ch < NCH; ++ch)
          la_tma_store_3d(&map_out, o_smem + ob * Cf::kOBytes + ch * 16384, ch * 64, (t0 + i) * kTilePx, n);
        tma_store_commit();
      }
    };

    if (nt > 0) epi_a(0);
    for (int i = 0; i < nt; ++i) {
      if (i + 1 < nt) epi_a(i + 1);
      epi_b(i);
    }
    __syncwarp();
    if (ew == 0 && elect_one_sync()) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int ctas_per_sample(int N, int tiles) {
  int cps = FD_NUM_SMS / N;
  if (cps < 1) cps = 1;
  if (cps > tiles) cps = tiles;
  return cps;
}

template <int C>
int run_tc(const void* x, const void* wk, const float* sk, const float* mk, const void* wq, const float* sq, const float* wv,
           const float* wout, const float* bout, const float* g2, void* out, float* workspace, int N, int HW, float eps,
           cudaStream_t st) {
  const int tiles = (HW + kTilePx - 1) / kTilePx;
  const int cps = ctas_per_sample(N, tiles);
  float* partial = workspace;
  __nv_bfloat16* mt = reinterpret_cast<__nv_bfloat16*>(workspace + (size_t)N * cps * 128 * (C + 4));
  CUtensorMap map_x, map_wk, map_wq, map_m, map_out;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)HW, (uint64_t)N};
    const uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)HW * C * 2};
    const uint32_t box[3] = {64, (uint32_t)kTilePx, 1};
    if (int e = make_tmap_bf16(&map_x, x, 3, dims, str, box)) return e;
    if (int e = make_tmap_bf16(&map_out, out, 3, dims, str, box)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, 128};
    const uint64_t str[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {64, 128};
    if (int e = make_tmap_bf16(&map_wk, wk, 2, dims, str, box)) return e;
    if (int e = make_tmap_bf16(&map_wq, wq, 2, dims, str, box)) return e;
  }
  {
    const uint64_t dims[3] = {128, (uint64_t)C, (uint64_t)N};
    const uint64_t str[2] = {256, (uint64_t)C * 256};
    const uint32_t box[3] = {64, (uint32_t)C, 1};
    if (int e = make_tmap_bf16(&map_m, mt, 3, dims, str, box)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FD_CUDA(cudaFuncSetAttribute(linattn_ctx_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtxCfg<C>::kSmemBytes));
    FD_CUDA(cudaFuncSetAttribute(linattn_apply_kernel_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, ApTcCfg<C>::kSmemBytes));
    attr_set = true;
  }
  CtxParams cp{N, HW, cps, tiles, eps, sk, mk, partial};
  FD_CUDA(fd_launch_pdl(linattn_ctx_kernel<C>, dim3(N * cps), dim3(kLaThreads), CtxCfg<C>::kSmemBytes, st, map_x, map_wk, cp));
  FD_LAUNCH_CHECK();
  linattn_tc_combine_kernel<C><<<N, 128, 0, st>>>(partial, wv, wout, mt, cps, HW);
  FD_LAUNCH_CHECK();
  ApParams ap{N, HW, cps, tiles, eps, sq, bout, g2};
  FD_CUDA(fd_launch_pdl(linattn_apply_kernel_tc<C>, dim3(N * cps), dim3(kLaThreads), ApTcCfg<C>::kSmemBytes, st, map_x, map_wq, map_m,
                        map_out, ap));
  FD_LAUNCH_CHECK();
  return FD_OK;
}

}  // namespace

extern "C" {

size_t fd_linattn_tc_workspace_floats(int N, int HW, int C) {
  const int tiles = (HW + kTilePx - 1) / kTilePx;
  const int cps = ctas_per_sample(N > 0 ? N : 1, tiles > 0 ? tiles : 1);
  return (size_t)N * cps * 128 * (C + 4) + (size_t)N * C * 128 / 2 + 64;
}

int fd_linattn_tc_prep(const float* wqkv, const float* g1, void* wq, float* sq, void* wk, float* sk, float* mk, float* wv, int C,
                       void* stream) {
  FD_REQUIRE(wqkv && g1 && wq && sq && wk && sk && mk && wv, "linattn_tc_prep: null pointer");
  FD_REQUIRE(C == 64 || C == 128, "linattn_tc_prep: C=%d not in {64,128}", C);
  linattn_tc_prep_kernel<<<384, 128, 0, (cudaStream_t)stream>>>(wqkv, g1, static_cast<__nv_bfloat16*>(wq), sq,
                                                               static_cast<__nv_bfloat16*>(wk), sk, mk, wv, C);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_linattn_tc(const void* x, const void* wk, const float* sk, const float* mk, const void* wq, const float* sq,
                  const float* wv, const float* wout, const float* bout, const float* g2, void* out, float* workspace, int N,
                  int HW, int C, float eps, void* stream) {
  FD_REQUIRE(x && wk && sk && mk && wq && sq && wv && wout && bout && g2 && out && workspace, "linattn_tc: null pointer");
  FD_REQUIRE(N > 0 && HW > 0, "linattn_tc: bad geometry N=%d HW=%d", N, HW);
  FD_REQUIRE(C == 64 || C == 128, "linattn_tc: C=%d not in {64,128}", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 64) return run_tc<64>(x, wk, sk, mk, wq, sq, wv, wout, bout, g2, out, workspace, N, HW, eps, st);
  return run_tc<128>(x, wk, sk, mk, wq, sq, wv, wout, bout, g2, out, workspace, N, HW, eps, st);
}

}  // extern "C"
