// Epilogue shared by the tensor-core convolution kernels (fd_conv_igemm.cu, fd_conv_strip.cu):
// 8 warps drain the fp32 accumulator of a 128 x BLOCK_N tile from TMEM.
//   TMEM lane quarter q = warp % 4 (hardware rule); the two warps of a quarter split the 32-column chunks
//   (even / odd).  Output goes registers -> +bias (+residual) -> bf16 -> 128B-swizzled smem slab (64 channels) ->
//   one TMA store per slab, which also clips rows / columns outside the image.  GroupNorm partial sums
//   (sum, sum of squares per (sample, group); Block.forward :176,181 of the reference) stay in registers across
//   the tiles of one (image, N-tile) and are flushed once: fixed-order reduction in smem, then double atomics.
#pragma once

#include "fd_tc.cuh"

namespace fdtc {

constexpr int kBlockM = 128;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kSlabBytes = kBlockM * 128;           // 128 rows x 64 bf16: one TMA-store slab

struct EpiTile {
  int img, h0, w0, n_tile;
};

struct EpiCtx {
  uint32_t tmem_base;
  uint32_t o_smem;              // NBUF staging slabs, 1024-byte aligned
  uint32_t tfull0, tempty0;     // mbarrier addresses of accumulator stage 0 (stage 1 at +8)
  float* s_bias;                // [2][BLOCK_N]
  float* s_stats;               // [8 warps][16]
  const CUtensorMap* map_out;   // (Cout, W, H, N, 1), box {64, Wt, R, 1, 1}
  const float* bias;
  const __nv_bfloat16* residual;
  double* gn_stats;
  int H, W, Cout, Wt;
  int shuffle_cq;               // > 0: pixel-shuffle store (dgrad of Downsample): map_out is (Cq, 2, W, 2, H), Cq = Cout / 4
  // Output head (BLOCK_N == 64 only; Unet.forward's last two lines, :416-417): when head_out is set the tile is NOT stored;
  // instead the 1x1 final conv (64 -> head_n <= 4 channels, fp32 weights head_w [head_n][64], bias head_b) is applied to the
  // fp32 values of each pixel and written as fp32 NCHW, cropped to the head_h0 x head_w0 window at (head_pt, head_pl).
  float* head_out;
  const float* head_w;
  const float* head_b;
  int head_n, head_h0, head_w0, head_pt, head_pl;
  float* s_head;                // [head_n * 64 weights | 8 warps x 32 lanes x 4 partial dot products]
  int phase;                    // >= 0: sub-pixel phase store (fd_conv_igemm_up): map_out is (Cout, 2, W, 2, N*H), the tile goes to
                                // rows 2 (img H + h) + (phase >> 1), columns 2 w + (phase & 1) of the up-sampled output
  int dbg;                      // diagnostics (FD_CONV_DBG): 8 = barrier handshakes only, no epilogue work
  int tempty_remote;            // != 0 (CTA pairs, non-leader): "accumulator drained" arrives on the LEADER CTA's barrier
  // Residual transform (ResnetBlock with a res_conv, :212-214: out = res_conv(x) + silu(GroupNorm(h2))): when rt_stats is
  // set, `residual` is the RAW conv output h2 of block2.proj and the epilogue applies block2's GroupNorm + SiLU to it
  // on the fly, so the activated tensor never goes through HBM (one read of h2 instead of read + write + read).
  const double* rt_stats;       // [N][8][2] (sum, sum of squares) of the residual tensor per (sample, group), or null
  const float* rt_gamma;        // [Cout]
  const float* rt_beta;         // [Cout]
  float rt_eps;
  float* s_rt;                  // [2 accumulator-stage parities][2][BLOCK_N]: folded a[c], b[c] of the current (image, N-tile)
};

// silu(z) = z sigmoid(z) = h + h tanh(h), h = z / 2: one MUFU op (tanh.approx, 2^-11 relative; outputs are bf16)
__device__ __forceinline__ float epi_silu_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(rank)
      : "memory");
}

// GPT = GroupNorm groups covered by one N-tile (8 when Cout == BLOCK_N, 4 when Cout == 2*BLOCK_N, 0 = no statistics).
// Must be called by warps 2..9 of the CTA (threads 64..319); next_tile(iter, tile) enumerates this CTA's tiles in
// the same order as the MMA warp.
// NACC = accumulator stages in TMEM (tile `iter` uses stage iter % NACC, columns [stage * BLOCK_N, +BLOCK_N); barriers tfull /
// tempty of stage s at +8 s)
template <int BLOCK_N, int GPT, int NBUF, int NACC = 2, class NextTile>
__device__ __forceinline__ void conv_epilogue(const EpiCtx& ec, NextTile next_tile) {
  constexpr int NCHUNK = BLOCK_N / 32;                 // 32-column accumulator chunks per tile
  constexpr int CPW = NCHUNK / 2;                      // chunks per epilogue warp
  constexpr int CPGT = GPT > 0 ? BLOCK_N / GPT : 32;   // columns per group inside the tile
  constexpr int GIC = CPGT < 32 ? 32 / CPGT : 1;       // groups inside one 32-column chunk
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_base = ec.tmem_base;
  const uint32_t o_smem = ec.o_smem;
  float* const s_bias = ec.s_bias;
  float* const s_stats = ec.s_stats;
  {
    // TMEM lane quarter q = warp % 4 (hardware rule); the two warps of a quarter split the 32-column chunks
    // (even / odd).  Output goes registers -> swizzled smem slab (64 channels) -> one TMA store per slab, which
    // also clips rows / columns outside the image.  GroupNorm partial sums stay in registers across the tiles
    // of one (image, N-tile) and are flushed once.
    const int ew = warp - 2;                  // 0..7
    const int et = threadIdx.x - 64;          // 0..255
    const int quarter = warp & 3;
    const int half = ew >> 2;                 // which chunk parity this warp owns
    const int row = quarter * 32 + lane;      // accumulator row = pixel within the tile
    const int rr = row / ec.Wt, ww = row - rr * ec.Wt;
    float st_s[CPW > 0 ? CPW : 1][GIC], st_q[CPW > 0 ? CPW : 1][GIC];
#pragma unroll
    for (int a = 0; a < (CPW > 0 ? CPW : 1); ++a)
#pragma unroll
      for (int b = 0; b < GIC; ++b) st_s[a][b] = st_q[a][b] = 0.f;
    int st_img = -1, st_ntile = 0;
    int rt_img = -1, rt_ntile = -1;
    if (BLOCK_N == 64 && ec.head_out != nullptr) {
      for (int i = et; i < ec.head_n * 64; i += kEpiThreads) ec.s_head[i] = __ldg(ec.head_w + i);
      // (published by the first slab barrier)
    }
    uint32_t slab_count = 0;

    auto flush_stats = [&]() {
      // all epilogue warps call this at the same tile boundary
      if (GPT == 0 || st_img < 0) return;
      float* mine = s_stats + ew * 16;
      if (lane < 16) mine[lane] = 0.f;
      __syncwarp();
#pragma unroll
      for (int a = 0; a < (CPW > 0 ? CPW : 1); ++a)
#pragma unroll
        for (int b = 0; b < GIC; ++b) {
          const float s = fd_warp_sum(st_s[a][b]), q = fd_warp_sum(st_q[a][b]);
          if (lane == 0) {
            const int col = (2 * a + half) * 32 + b * CPGT;      // first column of this partial inside the tile
            const int grp = col / CPGT;
            mine[grp * 2] += s;
            mine[grp * 2 + 1] += q;
          }
          st_s[a][b] = st_q[a][b] = 0.f;
        }
      named_bar_sync(2, kEpiThreads);
      if (et < 2 * GPT) {
        float sv = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < kEpiWarps; ++w8) sv += s_stats[w8 * 16 + et];     // fixed order
        atomicAdd(ec.gn_stats + (long)st_img * 16 + st_ntile * 2 * GPT + et, (double)sv);
      }
      named_bar_sync(2, kEpiThreads);
    };

    EpiTile tc;
    for (int iter = 0; next_tile(iter, tc); ++iter) {
      const int as = iter % NACC;
      const uint32_t aphase = (iter / NACC) & 1;
      const int n_tile = tc.n_tile, img = tc.img, h0 = tc.h0, w0 = tc.w0;
      const int h = h0 + rr, w = w0 + ww;
      const bool valid = (h < ec.H) && (w < ec.W);
      const int n0 = n_tile * BLOCK_N;
      if (GPT > 0 && (img != st_img || n_tile != st_ntile)) {
        flush_stats();
        st_img = img;
        st_ntile = n_tile;
      }
      float* bias_s = s_bias + (as & 1) * BLOCK_N;
      for (int i = et; i < BLOCK_N; i += kEpiThreads) bias_s[i] = ec.bias ? __ldg(ec.bias + n0 + i) : 0.f;
      float* rt_s = ec.s_rt;
      if (ec.rt_stats != nullptr && (img != rt_img || n_tile != rt_ntile)) {
        // Folded once per (image, N-tile), not per tile: the double-precision divisions / square root have a latency of
        // thousands of cycles on this part (measured: +200 us on a full-resolution 1x1 conv when done for every tile).
        // Safe to overwrite in place: every read of the previous coefficients precedes the last slab barrier of the
        // previous tile.  Same folding as gn_silu_kernel, pre-scaled by 1/2 for epi_silu_half.
        rt_img = img;
        rt_ntile = n_tile;
        const int cpg = ec.Cout >> 3;
        const double cnt = (double)ec.H * (double)ec.W * (double)cpg;
        for (int i = et; i < BLOCK_N; i += kEpiThreads) {
          const int c = n0 + i;
          const int g = c / cpg;
          const double s = ec.rt_stats[((long)img * 8 + g) * 2], ss = ec.rt_stats[((long)img * 8 + g) * 2 + 1];
          const double mean = s / cnt;
          double var = ss / cnt - mean * mean;
          if (var < 0.0) var = 0.0;
          const float rstd = (float)(1.0 / sqrt(var + (double)ec.rt_eps));
          const float ga = __ldg(ec.rt_gamma + c) * rstd;
          rt_s[i] = 0.5f * ga;
          rt_s[BLOCK_N + i] = 0.5f * (__ldg(ec.rt_beta + c) - (float)mean * ga);
        }
      }
      // (the slab barrier below also publishes the bias and the folded coefficients)

      // The residual row of the first slab is requested BEFORE waiting for the accumulator, the next slab's while this one
      // is processed: measured on the full-resolution 1x1 res_convs, a load issued after the wait exposes the whole DRAM
      // latency once per tile (8 warps x 64 B in flight per SM = ~3 TB/s over the chip).
      const long pix = ((long)img * ec.H + h) * ec.W + w;
      const __nv_bfloat16* rrow = (ec.residual && valid) ? ec.residual + pix * ec.Cout + n0 : nullptr;
      uint4 rpre[4];
      if (rrow != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; ++q) rpre[q] = __ldg(reinterpret_cast<const uint4*>(rrow + half * 32) + q);
      }
      mbar_wait(ec.tfull0 + 8u * as, aphase);
      tc_fence_after();
      if (ec.dbg & 8) {
        tc_fence_before();
        if (ec.tempty_remote) mbar_arrive_cluster(ec.tempty0 + 8u * as, 0);
        else mbar_arrive(ec.tempty0 + 8u * as);
        continue;
      }
#pragma unroll
      for (int slab = 0; slab < (BLOCK_N + 63) / 64; ++slab) {
        const uint32_t buf = o_smem + (slab_count % NBUF) * kSlabBytes;
        ++slab_count;
        if (ew == 0 && elect_one_sync()) tma_store_wait_read<NBUF - 1>();   // the store that last used this buffer has read it
        named_bar_sync(1, kEpiThreads);
        const int ci = slab * 2 + half;                   // this warp's chunk inside the slab
        const int c = ci * 32;
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + c), acc);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + j4 * 4);
          v[j4 * 4 + 0] = __uint_as_float(acc[j4 * 4 + 0]) + b4.x;
          v[j4 * 4 + 1] = __uint_as_float(acc[j4 * 4 + 1]) + b4.y;
          v[j4 * 4 + 2] = __uint_as_float(acc[j4 * 4 + 2]) + b4.z;
          v[j4 * 4 + 3] = __uint_as_float(acc[j4 * 4 + 3]) + b4.w;
        }
        if (rrow != nullptr) {
          uint4 rcur[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) rcur[q] = rpre[q];
          if (slab + 1 < (BLOCK_N + 63) / 64) {
#pragma unroll
            for (int q = 0; q < 4; ++q) rpre[q] = __ldg(reinterpret_cast<const uint4*>(rrow + c + 64) + q);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 r4 = rcur[q];
            const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
            if (ec.rt_stats != nullptr) {
              const float4 a0 = *reinterpret_cast<const float4*>(rt_s + c + q * 8), a1 = *reinterpret_cast<const float4*>(rt_s + c + q * 8 + 4);
              const float4 b0 = *reinterpret_cast<const float4*>(rt_s + BLOCK_N + c + q * 8),
                           b1 = *reinterpret_cast<const float4*>(rt_s + BLOCK_N + c + q * 8 + 4);
              const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = fd_unpack_bf16(rw[e]);
                v[q * 8 + e * 2] += epi_silu_half(fmaf(aa[e * 2], f.x, bb[e * 2]));
                v[q * 8 + e * 2 + 1] += epi_silu_half(fmaf(aa[e * 2 + 1], f.y, bb[e * 2 + 1]));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = fd_unpack_bf16(rw[e]);
                v[q * 8 + e * 2] += f.x;
                v[q * 8 + e * 2 + 1] += f.y;
              }
            }
          }
        }
        if (GPT > 0 && valid) {
#pragma unroll
          for (int b = 0; b < GIC; ++b) {
            constexpr int span = CPGT < 32 ? CPGT : 32;
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int j = 0; j < span; ++j) {
              const float x = v[b * span + j];
              s += x;
              q = fmaf(x, x, q);
            }
            st_s[slab][b] += s;
            st_q[slab][b] += q;
          }
        }
        if (BLOCK_N == 64 && ec.head_out != nullptr) {
          // this warp holds 32 of the pixel's 64 channels: partial dot products, the other half comes from warp ew ^ 4
          const float* hw = ec.s_head;
          float* hx = ec.s_head + 256 + (ew * 32 + lane) * 4;
          float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            if (o < ec.head_n) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 w4 = *reinterpret_cast<const float4*>(hw + o * 64 + c + j4 * 4);      // (broadcast reads)
                d[o] = fmaf(v[j4 * 4 + 0], w4.x, d[o]);
                d[o] = fmaf(v[j4 * 4 + 1], w4.y, d[o]);
                d[o] = fmaf(v[j4 * 4 + 2], w4.z, d[o]);
                d[o] = fmaf(v[j4 * 4 + 3], w4.w, d[o]);
              }
            }
          }
          *reinterpret_cast<float4*>(hx) = make_float4(d[0], d[1], d[2], d[3]);
          tc_fence_before();
          if (ec.tempty_remote) mbar_arrive_cluster(ec.tempty0 + 8u * as, 0);
          else mbar_arrive(ec.tempty0 + 8u * as);
          named_bar_sync(1, kEpiThreads);
          if (half == 0 && valid) {
            const float4 o4 = *reinterpret_cast<const float4*>(ec.s_head + 256 + ((ew ^ 4) * 32 + lane) * 4);
            const float r4[4] = {d[0] + o4.x, d[1] + o4.y, d[2] + o4.z, d[3] + o4.w};
            const int hh = h - ec.head_pt, wq = w - ec.head_pl;
            if (hh >= 0 && hh < ec.head_h0 && wq >= 0 && wq < ec.head_w0) {
#pragma unroll
              for (int o = 0; o < 4; ++o)
                if (o < ec.head_n)
                  ec.head_out[(((long)img * ec.head_n + o) * ec.head_h0 + hh) * (long)ec.head_w0 + wq] = r4[o] + __ldg(ec.head_b + o);
            }
          }
          continue;       // (the next tile's first slab barrier orders the reuse of the exchange buffer)
        }
        // 64-byte piece of this row inside the 128-byte slab row, 16-byte granules XOR-swizzled by (row & 7)
        const uint32_t rbase = buf + (uint32_t)row * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t piece = (uint32_t)(half * 4 + q) ^ (uint32_t)(row & 7);
          const uint32_t o0 = fd_pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), o1 = fd_pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
          const uint32_t o2 = fd_pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), o3 = fd_pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + piece * 16u), "r"(o0), "r"(o1), "r"(o2),
                       "r"(o3)
                       : "memory");
        }
        fence_proxy_async_smem();
        if (slab == (BLOCK_N + 63) / 64 - 1) {
          tc_fence_before();
          if (ec.tempty_remote) mbar_arrive_cluster(ec.tempty0 + 8u * as, 0);
          else mbar_arrive(ec.tempty0 + 8u * as);     // all TMEM reads of this tile are done
        }
        named_bar_sync(1, kEpiThreads);
        if (ew == 0 && elect_one_sync()) {                 // elect.sync: straight-line UTMASTG (see fd_tc.cuh)
          const int cc = n0 + slab * 64;
          if (ec.shuffle_cq > 0) {
            const int pq = cc / ec.shuffle_cq;          // p1 * 2 + p2
            tma_store_5d(ec.map_out, buf, cc - pq * ec.shuffle_cq, pq & 1, w0, pq >> 1, h0);
          } else if (ec.phase >= 0) {
            tma_store_5d(ec.map_out, buf, cc, ec.phase & 1, w0, ec.phase >> 1, img * ec.H + h0);
          } else {
            tma_store_5d(ec.map_out, buf, cc, w0, h0, img, 0);
          }
          tma_store_commit();
        }
      }
    }
    flush_stats();
    __syncwarp();
    if (ew == 0 && elect_one_sync()) tma_store_wait_all();     // same elected lane as the stores (bulk groups are per thread)
  }
}

}  // namespace fdtc
