// Attention core of the UNet's middle block (denoising_diffusion.py:256-267): softmax(q^T k 32^-0.5) v over all HW tokens,
// 4 heads x 32 channels, as a flash-style streaming kernel on the 5th-generation tensor cores -- S = Q K^T and O += P V are
// tcgen05.mma with their accumulators (and P itself) in tensor memory, Q / K / V tiles staged by TMA.  The reference
// materialises the N x N score matrix (17 GB per sample at 1024x2048); here S and P never leave the SM.
//
//   CTA = (sample, 128-query block), all four heads: the Q / K / V tiles are whole 128-channel rows of the qkv tensor (two
//   64-channel SWIZZLE_128B chunks each), a head is a 32-channel slice inside a chunk: for the K-major operands (Q, K) that
//   is the ordinary K advance inside the swizzle atom, for V (MN-major B operand, K = keys) a start 64 B into the row.
//   Step t = (key block j, head h).  Per step:
//     QK(t):  S[128 q][128 keys] = Q_h K_h^T                 2 MMAs (K = 32), SS form, into S buffer t & 1 (128 TMEM columns)
//     softmax warps (thread = query row): load the row, running maximum with LAZY rescaling (O and l are rescaled only
//             when the maximum grows by more than 2^8, FlashAttention-4 style), P = exp2(s c - m) as bf16 -- computed two at a
//             time by ex2.approx.ftz.bf16x2, because with d = 32 the kernel is bound by the MUFU pipe, not the tensor pipe
//             (128 x 128 exps per 2 x 128 x 128 x 32 MACs) -- and stored back to TMEM over the S it came from
//     PV(t):  O_h[128 q][32] += P V_h                         8 MMAs (K = 16 keys each), A operand from TMEM (TS form)
//             the softmax denominator is summed from the same bf16 P on the CUDA cores (packed bf16x2 tree per 32 keys, fp32
//             across chunks; a first version summed it on the tensor core against a tile of ones, but the 8 extra small MMAs
//             per step made the single issuing thread the bottleneck: 656 us, measured, against 965 us for mma.sync)
//   Two softmax groups of four warps work on alternating steps (S buffers), so the exps of one step overlap the TMEM
//   traffic and MMAs of the other.  Warp 0 = TMA producer (2-stage K / V ring), warp 1 = MMA issuer (descriptors of a step
//   are built in general registers BEFORE its barrier wait, see fd_conv_igemm.cu on the uniform-register scoreboard).
//   TMEM columns: S / P buffers [0, 256), O_h at 256 + 32 h.
//   Measured (b8 x 7040 tokens, mma.sync kernel 965 us): this kernel 664 us = 56 % of the exp bound (1.59e9 exps at the chip's
//   4.26e12 MUFU ops/s = 372 us; the MUFU pipe is 53 % busy: two softmax warps per scheduler do not cover each other's TMEM-load
//   and barrier latencies).  Tried and slower: one softmax group PER HEAD with 64-key steps (16 softmax warps; the 640-thread
//   register allocation leaves 92 registers per thread, the score row spills: 1006 us).
#include <stdlib.h>

#include "fd_tc.cuh"

using namespace fdtc;

namespace {

constexpr int kAtThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2..5 softmax group 0, warps 6..9 group 1
constexpr int kTileBytes = 32768;        // [128 rows][128 channels] bf16 = two 64-channel chunks of 16 KB
constexpr int kKvStages = 2;
constexpr int kOffK = kTileBytes;                               // after the Q tile
constexpr int kOffV = kOffK + kKvStages * kTileBytes;
constexpr int kOffOnes = kOffV + kKvStages * kTileBytes;        // [16 rows][128 B] of bf16 1.0
constexpr int kOffBar = kOffOnes + 2048;
constexpr int kAtSmemBytes = 1024 + kOffBar + 256;
constexpr float kScaleLog2e = 0.17677669529663687f * 1.4426950408889634f;     // 32^-0.5 log2(e)
constexpr float kRescaleThreshold = 8.f;                                       // log2 units

__device__ __forceinline__ uint64_t at_desc_mn(uint32_t saddr) {      // MN-major SWIZZLE_128B, N <= 64 (LBO unused)
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__host__ __device__ constexpr uint32_t at_idesc(int M, int N, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void at_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void at_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void at_tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void at_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void at_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// two exps per MUFU op: exp2 of a packed bf16 pair
__device__ __forceinline__ uint32_t at_ex2_bf16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AtParams {
  int N, HW, qblocks, nkv;
  __nv_bfloat16* out;      // (N, HW, 128)
  float* lse;              // [N][4][HW] log2-domain log-sum-exp of the scaled scores, or null
  int dbg;
};

__global__ void __launch_bounds__(kAtThreads, 1) attention_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                     const __grid_constant__ CUtensorMap map_k,
                                                                     const __grid_constant__ CUtensorMap map_v, const AtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t q_smem = base, k_smem = base + kOffK, v_smem = base + kOffV;
  const uint32_t bar = base + kOffBar;
  const uint32_t qfull = bar;
  auto kvfull = [&](int s) { return bar + 8u * (1 + s); };
  auto kvempty = [&](int s) { return bar + 8u * (3 + s); };
  auto sfull = [&](int s) { return bar + 8u * (5 + s); };
  auto pfull = [&](int s) { return bar + 8u * (7 + s); };
  auto pvdone = [&](int h) { return bar + 8u * (9 + h); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gbase + kOffBar + 8 * 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.qblocks, qb = blockIdx.x % p.qblocks;
  const int nkv = p.nkv, T = 4 * nkv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(qfull, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(kvfull(s), 1);
      mbar_init(kvempty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(sfull(s), 1);
      mbar_init(pfull(s), 128);
    }
    for (int h = 0; h < 4; ++h) mbar_init(pvdone(h), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  fd_grid_dependency_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(qfull, kTileBytes);
      for (int ch = 0; ch < 2; ++ch) tma_load_3d(q_smem + ch * 16384, &map_q, qfull, ch * 64, qb * 128, n);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % kKvStages;
        mbar_wait(kvempty(s), ((j / kKvStages) & 1) ^ 1u);
        mbar_expect_tx(kvfull(s), 2 * kTileBytes);
        for (int ch = 0; ch < 2; ++ch) {
          tma_load_3d(k_smem + s * kTileBytes + ch * 16384, &map_k, kvfull(s), ch * 64, j * 128, n);
          tma_load_3d(v_smem + s * kTileBytes + ch * 16384, &map_v, kvfull(s), ch * 64, j * 128, n);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t id_qk = at_idesc(128, 128, false);
      constexpr uint32_t id_pv = at_idesc(128, 32, true);
      mbar_wait(qfull, 0);
      const uint64_t qd0 = umma_desc_sw128(q_smem), kd0 = umma_desc_sw128(k_smem), vd0 = at_desc_mn(v_smem);
      auto qk = [&](int t) {
        const int h = t & 3, j = t >> 2, s = j % kKvStages;
        // head h: chunk h >> 1 (16384 B = 1024 descriptor units), 64 B (4 units) into the 128-byte rows for odd h, then
        // 2 units per 16 channels; everything in general registers before the wait
        const uint32_t off = (uint32_t)((h >> 1) * 1024 + (h & 1) * 4);
        const uint64_t a0 = opaque64(qd0 + off), a1 = opaque64(qd0 + off + 2);
        const uint64_t b0 = opaque64(kd0 + (uint64_t)(s * (kTileBytes >> 4)) + off), b1 = opaque64(kd0 + (uint64_t)(s * (kTileBytes >> 4)) + off + 2);
        const uint32_t sd = (uint32_t)opaque32((int)(tmem_base + (t & 1) * 128));
        if (h == 0) mbar_wait(kvfull(s), (j / kKvStages) & 1);
        tc_fence_after();
        umma_bf16(sd, a0, b0, id_qk, 0u);
        umma_bf16(sd, a1, b1, id_qk, 1u);
        umma_commit(sfull(t & 1));
      };
      qk(0);
      if (T > 1) qk(1);
      for (int t = 0; t < T; ++t) {
        const int h = t & 3, j = t >> 2, s = j % kKvStages;
        const uint32_t p_tmem = tmem_base + (t & 1) * 128;
        // V_h: chunk h >> 1, 64 B into the rows for odd h; 16 keys (2048 B = 128 units) per K step
        const uint64_t vb = vd0 + (uint64_t)(s * (kTileBytes >> 4) + (h >> 1) * 1024 + (h & 1) * 4);
        uint64_t vd[8];
        uint32_t pa[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          vd[ks] = opaque64(vb + (uint64_t)(ks * 128));
          pa[ks] = (uint32_t)opaque32((int)(p_tmem + ks * 8));
        }
        const uint32_t od = (uint32_t)opaque32((int)(tmem_base + 256 + h * 32));
        mbar_wait(pfull(t & 1), (t >> 1) & 1);
        tc_fence_after();
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) at_umma_ts(od, pa[ks], vd[ks], id_pv, (j | ks) != 0 ? 1u : 0u);
        }
        umma_commit(pvdone(h));
        if (h == 3) umma_commit(kvempty(s));
        if (t + 2 < T) qk(t + 2);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax groups =====================
    const int grp = (warp - 2) >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int q_tok = qb * 128 + row;
    float m[2] = {-INFINITY, -INFINITY};                  // running maxima (scaled, log2 domain) of this group's two heads
    float l[2] = {0.f, 0.f};                              // running denominators
    for (int t = grp; t < T; t += 2) {
      const int h = t & 3, j = t >> 2, hl = h >> 1;
      mbar_wait(sfull(grp), (t >> 1) & 1);
      tc_fence_after();
      const uint32_t sb = tmem_base + lane_off + (uint32_t)(grp * 128);
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(sb + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      const int kleft = p.HW - j * 128;                 // keys of this block that exist (TMA zero-fills the rest)
      if (kleft < 128) {
        // last, ragged key block only (a real branch: as straight-line selects this cost 2 instructions per score in EVERY step)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // -inf: exp2 gives exactly 0.  (s is indexed with compile-time constants only: it has to stay in registers)
            if (c == 0 && i >= kleft) s[i] = 0xFF800000u;
            if (c == 1 && 32 + i >= kleft) s[32 + i] = 0xFF800000u;
            if (c == 2 && 64 + i >= kleft) s[64 + i] = 0xFF800000u;
            if (c == 3 && 96 + i >= kleft) s[96 + i] = 0xFF800000u;
          }
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 128; i += 8) {
#pragma unroll
        for (int a4 = 0; a4 < 4; ++a4)
          mx4[a4] = fmaxf(mx4[a4], fmaxf(__uint_as_float(s[i + 2 * a4]), __uint_as_float(s[i + 2 * a4 + 1])));
      }
      const float bm = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kScaleLog2e;
      // lazy rescaling: the accumulators follow the maximum only when it grew by more than 2^8 (P stays <= 256)
      const bool first = j == 0;
      const bool grow = !first && bm > m[hl] + kRescaleThreshold;
      if (first) m[hl] = bm;
      if (!first) {
        mbar_wait(pvdone(h), (j - 1) & 1);              // PV of this head's previous key block has retired
        if (__any_sync(0xffffffffu, grow)) {
          tc_fence_after();
          const float f = grow ? at_ex2(m[hl] - bm) : 1.f;
          if (grow) m[hl] = bm;
          l[hl] *= f;
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + (uint32_t)(256 + h * 32), o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          at_tmem_st32(tmem_base + lane_off + (uint32_t)(256 + h * 32), o);
        }
      }
      const float nm = -m[hl];
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // (ex2.approx.ftz.bf16x2 was measured first: it compiles to TWO MUFU.EX2.BF16 plus PRMTs per pair, no faster than fp32)
          const float e0 = at_ex2(fmaf(__uint_as_float(s[c * 32 + 2 * i]), kScaleLog2e, nm));
          const float e1 = at_ex2(fmaf(__uint_as_float(s[c * 32 + 2 * i + 1]), kScaleLog2e, nm));
          pk[i] = fd_pack_bf16(e0, e1);
        }
        at_tmem_st16(sb + c * 16, pk);                  // P (bf16 pairs) over the first 64 columns of the S buffer
        // denominator from the SAME bf16 values that multiply V: packed bf16x2 tree over the chunk's 32 values (unbiased
        // roundings of 2^-9, averaged over 4 x nkv chunks), fp32 from there on
        __nv_bfloat162 t8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          t8[i] = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&pk[2 * i]), *reinterpret_cast<const __nv_bfloat162*>(&pk[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 4; ++i) t8[i] = __hadd2(t8[2 * i], t8[2 * i + 1]);
        t8[0] = __hadd2(t8[0], t8[1]);
        t8[2] = __hadd2(t8[2], t8[3]);
        const float2 f = __bfloat1622float2(__hadd2(t8[0], t8[2]));
        lsum += f.x + f.y;
      }
      l[hl] += lsum;
      at_tmem_st_wait();
      tc_fence_before();
      mbar_arrive(pfull(grp));
    }
    // ---- O_h / l_h -> out
    for (int hl = 0; hl < 2; ++hl) {
      const int h = hl * 2 + grp;
      mbar_wait(pvdone(h), (nkv - 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(tmem_base + lane_off + (uint32_t)(256 + h * 32), o);
      tmem_ld_wait();
      if (q_tok < p.HW) {
        const float denom = l[hl];
        const float inv = 1.f / denom;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long)n * p.HW + q_tok) * 128 + h * 32);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 v;
          v.x = fd_pack_bf16(__uint_as_float(o[q4 * 8 + 0]) * inv, __uint_as_float(o[q4 * 8 + 1]) * inv);
          v.y = fd_pack_bf16(__uint_as_float(o[q4 * 8 + 2]) * inv, __uint_as_float(o[q4 * 8 + 3]) * inv);
          v.z = fd_pack_bf16(__uint_as_float(o[q4 * 8 + 4]) * inv, __uint_as_float(o[q4 * 8 + 5]) * inv);
          v.w = fd_pack_bf16(__uint_as_float(o[q4 * 8 + 6]) * inv, __uint_as_float(o[q4 * 8 + 7]) * inv);
          dst[q4] = v;
        }
        if (p.lse != nullptr) p.lse[((long)n * 4 + h) * p.HW + q_tok] = m[hl] + log2f(denom);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int fd_attention_tc_launch(const void* qkv, void* out, float* lse, int N, int HW, cudaStream_t st) {
  const int qblocks = (HW + 127) / 128;
  CUtensorMap map_q, map_k, map_v;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  for (int which = 0; which < 3; ++which) {
    // q | k | v are the three 128-channel thirds of the 384-channel rows
    const uint64_t dims[3] = {128, (uint64_t)HW, (uint64_t)N};
    const uint64_t str[2] = {384 * 2, (uint64_t)HW * 384 * 2};
    const uint32_t box[3] = {64, 128, 1};
    CUtensorMap* m = which == 0 ? &map_q : (which == 1 ? &map_k : &map_v);
    if (int e = make_tmap_bf16(m, base + which * 128, 3, dims, str, box)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FD_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
    attr_set = true;
  }
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("FD_ATTN_DBG"); dbg = e ? atoi(e) : 0; }
  AtParams p{N, HW, qblocks, qblocks, static_cast<__nv_bfloat16*>(out), lse, dbg};
  FD_CUDA(fd_launch_pdl(attention_tc_kernel, dim3(N * qblocks), dim3(kAtThreads), kAtSmemBytes, st, map_q, map_k, map_v, p));
  FD_LAUNCH_CHECK();
  return FD_OK;
}
