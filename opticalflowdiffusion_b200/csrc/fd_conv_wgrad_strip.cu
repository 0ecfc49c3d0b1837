// Weight gradient of the 3x3 / pad 1 convolutions on large images (86 % of the UNet's conv FLOPs), the
// rolling-strip counterpart of fd_conv_strip.cu for the backward pass.
//
// The generic wgrad (fd_conv_wgrad.cu) runs one job per tap, so every source pixel and every dy pixel is pulled
// from L2 nine times: 16-24 KB of L2->smem traffic per 0.5 MMAC block, L2-bound at ~175 TFLOP/s for the Cout = 64
// layers.  Here a CTA owns one (64 input channels) x (64 output channels) block of dW and walks down a 128-pixel
// wide column of the image.  Per output row it TMA-loads ONE 130-pixel source strip (halo included, zero-filled
// outside the image) and ONE 128-pixel dy strip; the nine taps are UMMA A-operands that alias three resident
// source strips at 0/1/2-pixel offsets, exactly like the forward strip kernel, except that here the pixels are
// the REDUCTION dimension: both operands are MN-major SWIZZLE_128B tiles (channel rows of 128 B, 8-pixel atoms
// 1024 B apart), so a 1-pixel shift is a +128 B descriptor start.
//
//   dW[co][tap][ci] = sum_px  src[px + tap][ci] * dy[px][co]
//   M = 128 = two taps' 64 input channels (kx = 0 and kx = 1 of one kernel row: the second 64-row chunk of the
//       A operand starts LBO = 128 B after the first), or one tap (kx = 2) with an ignored upper half;
//   N = 64 output channels;  K = 16 pixels per instruction;  6 fp32 accumulators of 128 x 64 live in TMEM for the
//   whole pixel range (384 columns).  Source and dy strips live in two separate TMA rings (8 x 17 KB, 5 x 16 KB).
//
// Work = (block pair, image, column block, row) units, flattened and cut into equal contiguous ranges, one per
// CTA (one CTA per SM); a CTA flushes its accumulators into dW with fp32 reductions whenever its range crosses
// into another block pair and at the end.  Two MMA issuer warps (accumulators 0-2 / 3-5) keep the tensor pipe fed.
#include "fd_tc.cuh"

using namespace fdtc;

namespace {

constexpr int kTileW = 128;
constexpr int kStripPx = kTileW + 2;
constexpr int kSrcTx = kStripPx * 128;          // bytes TMA delivers per source strip
constexpr int kSrcSlot = 136 * 128;             // multiple of 1024: every slot keeps the swizzle phase
constexpr int kDyTx = kTileW * 128;
constexpr int kNSS = 8;                         // source-strip ring: 3 rows in use + 5 prefetched
constexpr int kNSD = 5;                         // dy-strip ring: 1 row in use + 4 prefetched
constexpr int kRingBytes = kNSS * kSrcSlot + kNSD * kDyTx;     // 221184: TMA latency, not smem bandwidth, is what
                                                // starves the tensor pipe (ncu: 52 % active with 2 rows prefetched)
constexpr int kGroups = 6;
constexpr int kTmemCols = 512;
constexpr int kThreads = 7 * 32;                // warp 0 TMA, warps 1-2 MMA issuers, warps 3-6 flush
constexpr int kSmemBytes = 1024 + kRingBytes + 512;

struct WsParams {
  int N, H, W;
  int Cin, Cout;
  int chunks0;               // 64-channel chunks of src0 (the rest come from src1)
  int ci_chunks, co_chunks, wblocks;
  long rows_per_pair;        // N * wblocks * H
  long total_units;          // pairs * rows_per_pair
  float* dw;
};

struct Item {
  int pair, n, w0, ha, hb;
};

// units [u, end) -> the next item: rows of ONE (pair, image, column block)
__device__ __forceinline__ Item next_item(const WsParams& p, long u, long end) {
  Item it;
  it.pair = (int)(u / p.rows_per_pair);
  long r = u - (long)it.pair * p.rows_per_pair;
  const int h = (int)(r % p.H);
  r /= p.H;
  const int wb = (int)(r % p.wblocks);
  it.n = (int)(r / p.wblocks);
  it.w0 = wb * kTileW;
  it.ha = h;
  long left = end - u;
  it.hb = (int)min((long)p.H, (long)h + left);
  return it;
}

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_strip_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                      const __grid_constant__ CUtensorMap map_dy, const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t dy_base = base + kNSS * kSrcSlot;
  const uint32_t bar_base = base + kRingBytes;
  auto sfull_bar = [&](int s) { return bar_base + 8u * s; };
  auto sempty_bar = [&](int s) { return bar_base + 8u * (kNSS + s); };
  auto dfull_bar = [&](int s) { return bar_base + 8u * (2 * kNSS + s); };
  auto dempty_bar = [&](int s) { return bar_base + 8u * (2 * kNSS + kNSD + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * kNSS + 2 * kNSD);
  const uint32_t acc_empty = bar_base + 8u * (2 * kNSS + 2 * kNSD + 1);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gbase + kRingBytes + 8 * (2 * kNSS + 2 * kNSD + 2) + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const long u_begin = (p.total_units * blockIdx.x) / gridDim.x;
  const long u_end = (p.total_units * (blockIdx.x + 1)) / gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_dy);
    for (int s = 0; s < kNSS; ++s) {
      mbar_init(sfull_bar(s), 1);
      mbar_init(sempty_bar(s), 2);
    }
    for (int s = 0; s < kNSD; ++s) {
      mbar_init(dfull_bar(s), 1);
      mbar_init(dempty_bar(s), 2);
    }
    mbar_init(acc_full, 2);
    mbar_init(acc_empty, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      uint32_t seq = 0, dseq = 0;
      for (long u = u_begin; u < u_end;) {
        const Item it = next_item(p, u, u_end);
        const int ci_chunk = it.pair / p.co_chunks, co_chunk = it.pair % p.co_chunks;
        const CUtensorMap* ma = ci_chunk < p.chunks0 ? &map_a0 : &map_a1;
        const int c0 = (ci_chunk < p.chunks0 ? ci_chunk : ci_chunk - p.chunks0) * 64;
        const int rows = it.hb - it.ha;
        for (int j = 0; j < rows + 2; ++j, ++seq) {
          const int slot = seq % kNSS;
          mbar_wait(sempty_bar(slot), ((seq / kNSS) & 1u) ^ 1u);
          mbar_expect_tx(sfull_bar(slot), kSrcTx);
          tma_load_5d(base + slot * kSrcSlot, ma, sfull_bar(slot), c0, it.w0 - 1, it.ha - 1 + j, it.n, 0);
          if (j < rows) {
            const int ds = dseq % kNSD;
            mbar_wait(dempty_bar(ds), ((dseq / kNSD) & 1u) ^ 1u);
            mbar_expect_tx(dfull_bar(ds), kDyTx);
            tma_load_5d(dy_base + ds * kDyTx, &map_dy, dfull_bar(ds), co_chunk * 64, it.w0, it.ha + j, it.n, 0);
            ++dseq;
          }
        }
        u += rows;
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 -> accumulators 0..2 (taps kx = 0|1 of row ky), warp 2 -> 3..5 (kx = 2)
    constexpr uint32_t idesc = idesc_bf16_mn(128, 64);
    const bool pairs = warp == 1;
    uint32_t seq0 = 0, dseq0 = 0;
    int cur_pair = -1;
    uint32_t flushes = 0;
    bool fresh = true;
    for (long u = u_begin; u < u_end;) {
      const Item it = next_item(p, u, u_end);
      const int rows = it.hb - it.ha;
      if (it.pair != cur_pair) {
        if (cur_pair >= 0) {
          // hand the finished accumulators to the flush warps and wait until they are drained
          if (elect_one_sync()) umma_commit(acc_full);
          __syncwarp();
          mbar_wait(acc_empty, flushes & 1u);
          tc_fence_after();
          ++flushes;
        }
        cur_pair = it.pair;
        fresh = true;
      }
      for (int i = 0; i < rows; ++i) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t s = seq0 + i + ky;
          mbar_wait(sfull_bar(s % kNSS), (s / kNSS) & 1u);
        }
        mbar_wait(dfull_bar((dseq0 + i) % kNSD), ((dseq0 + i) / kNSD) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t dy = dy_base + ((dseq0 + i) % kNSD) * kDyTx;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint32_t strip = base + ((seq0 + i + ky) % kNSS) * kSrcSlot;
            const uint32_t a0 = strip + (pairs ? 0 : 2 * 128);
            const uint32_t lbo = pairs ? 128u : 0u;
            const uint32_t tmem_d = tmem_base + (uint32_t)((pairs ? ky : 3 + ky) * 64);
#pragma unroll
            for (int k = 0; k < kTileW / 16; ++k)
              umma_bf16(tmem_d, desc_mn_sw128(a0 + k * 2048, lbo), desc_mn_sw128(dy + k * 2048, 0u), idesc,
                        (fresh && i == 0 && k == 0) ? 0u : 1u);
          }
          umma_commit(sempty_bar((seq0 + i) % kNSS));
          umma_commit(dempty_bar((dseq0 + i) % kNSD));
          if (i == rows - 1) {
            umma_commit(sempty_bar((seq0 + i + 1) % kNSS));
            umma_commit(sempty_bar((seq0 + i + 2) % kNSS));
          }
        }
        __syncwarp();
      }
      fresh = false;
      seq0 += rows + 2;
      dseq0 += rows;
      u += rows;
    }
    if (cur_pair >= 0 && elect_one_sync()) umma_commit(acc_full);
    __syncwarp();
  } else {
    // ===================== flush: TMEM -> dW (fp32 reductions), once per block pair this CTA touched
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t flushes = 0;
    int cur_pair = -1;
    long u = u_begin;
    while (u < u_end) {
      // advance to the end of the current pair's units inside this CTA's range
      const int pair = (int)(u / p.rows_per_pair);
      const long pair_end = min(u_end, (long)(pair + 1) * p.rows_per_pair);
      cur_pair = pair;
      u = pair_end;
      mbar_wait(acc_full, flushes & 1u);
      tc_fence_after();
      const int ci_chunk = cur_pair / p.co_chunks, co_chunk = cur_pair % p.co_chunks;
      const long ktot = 9L * p.Cin;
#pragma unroll 1
      for (int g = 0; g < kGroups; ++g) {
        const int ky = g % 3;
        const int kx = g < 3 ? (row >> 6) : 2;
        const bool live = g < 3 || row < 64;
        float* dst = p.dw + (long)(co_chunk * 64) * ktot + (long)(ky * 3 + kx) * p.Cin + ci_chunk * 64 + (row & 63);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 64 + c * 32), acc);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (long)(c * 32 + j) * ktot, __uint_as_float(acc[j]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
      ++flushes;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// used by fd_conv_wgrad (fd_conv_wgrad.cu) for mode 0, 3x3, pad 1 on images at least 64 pixels wide
int fd_conv_wgrad_strip_launch(const void* src0, int C0, const void* src1, int C1, const void* dy, float* dw, int N, int H,
                               int W, int Cout, void* stream) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    FD_CUDA(cudaGetDevice(&dev));
    FD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  WsParams p{};
  p.N = N; p.H = H; p.W = W;
  p.Cin = C0 + C1;
  p.Cout = Cout;
  p.chunks0 = C0 / 64;
  p.ci_chunks = (C0 + C1) / 64;
  p.co_chunks = Cout / 64;
  p.wblocks = (W + kTileW - 1) / kTileW;
  p.rows_per_pair = (long)N * p.wblocks * H;
  p.total_units = p.rows_per_pair * p.ci_chunks * p.co_chunks;
  p.dw = dw;
  CUtensorMap ma0, ma1, mdy;
  {
    const uint64_t dims[5] = {(uint64_t)C0, (uint64_t)W, (uint64_t)H, (uint64_t)N, 1};
    const uint64_t str[4] = {(uint64_t)C0 * 2, (uint64_t)W * C0 * 2, (uint64_t)H * W * C0 * 2, (uint64_t)N * H * W * C0 * 2};
    const uint32_t box[5] = {64, (uint32_t)kStripPx, 1, 1, 1};
    if (int e = make_tmap_bf16(&ma0, src0, 5, dims, str, box)) return e;
  }
  if (C1 > 0) {
    const uint64_t dims[5] = {(uint64_t)C1, (uint64_t)W, (uint64_t)H, (uint64_t)N, 1};
    const uint64_t str[4] = {(uint64_t)C1 * 2, (uint64_t)W * C1 * 2, (uint64_t)H * W * C1 * 2, (uint64_t)N * H * W * C1 * 2};
    const uint32_t box[5] = {64, (uint32_t)kStripPx, 1, 1, 1};
    if (int e = make_tmap_bf16(&ma1, src1, 5, dims, str, box)) return e;
  } else {
    ma1 = ma0;
  }
  {
    const uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N, 1};
    const uint64_t str[4] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2,
                             (uint64_t)N * H * W * Cout * 2};
    const uint32_t box[5] = {64, (uint32_t)kTileW, 1, 1, 1};
    if (int e = make_tmap_bf16(&mdy, dy, 5, dims, str, box)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FD_CUDA(cudaFuncSetAttribute(wgrad3x3_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  long grid = sms;
  if (grid > p.total_units) grid = p.total_units;
  wgrad3x3_strip_kernel<<<(unsigned)grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(ma0, ma1, mdy, p);
  FD_LAUNCH_CHECK();
  return FD_OK;
}
