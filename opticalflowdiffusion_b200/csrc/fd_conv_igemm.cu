// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged
// by TMA.  Replaces every nn.Conv2d / WeightStandardizedConv2d call of the reference UNet
// (denoising_diffusion.py:92,98,114,200,222,225,253,254,297,339,354).
//
//   GEMM view:  M = output pixels (tile of 128 = R rows x Wt columns of ONE image),
//               N = output channels (tile BLOCK_N in {64,128,256}),
//               K = taps x input channels, consumed in blocks of 64 channels of one tap.
//
//   A operand:  activations are bf16 NHWC, i.e. every pixel is a 128-byte row per 64 channels.  A
//               K-block of A for tap (ky,kx) is the box {64 ch, Wt, R} of the input shifted by
//               (ky-pad, kx-pad): one TMA box load.  TMA zero-fills out-of-bounds elements, which
//               implements the convolution's zero padding with no branches and no im2col buffer;
//               the 128B-swizzled box lands in shared memory exactly in the canonical K-major
//               UMMA layout.  The skip-connection concat (:405,408,414) is a second tensor map
//               whose channels simply extend K; the pixel-unshuffle downsample (:95-99) is a 5-d
//               view (c, p2, w, p1, h) of the same tensor so its 1x1 conv needs no rearranged copy.
//   B operand:  weights packed bf16 [Cout][K] (K-major), 2-d TMA box {64, BLOCK_N}.
//   D:          fp32 in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of tile i
//               overlaps the MMAs of tile i+1.
//
//   Warp roles (192 threads, persistent CTA per SM, static round-robin tile schedule):
//     warp 0    TMA producer (one elected lane), STAGES-deep mbarrier ring
//     warp 1    TMEM allocator + MMA issuer (one lane issues tcgen05.mma, tcgen05.commit frees slots)
//     warps 2-5 epilogue: tcgen05.ld -> +bias (+residual) -> GroupNorm partial statistics
//               (sum, sum of squares per (sample, group), :176,181) -> bf16 NHWC store
#include "fd_tc.cuh"

using namespace fdtc;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kSlabBytes = kBlockM * 128;           // 128 rows x 64 bf16: one TMA-store slab

struct ConvParams {
  int N, H, W;        // images, output rows, output columns (after flattening / merging)
  int Cout;
  int KW;             // taps per row
  int taps;           // KH*KW (4 for the pixel-unshuffle mode)
  int pad_h, pad_w;
  int mode;           // 0: conv taps with zero padding, 1: pixel-unshuffle + 1x1
  int chunks0, chunks1;  // 64-channel chunks of src0 / src1
  int R, Wt;          // tile = R rows x Wt columns, R*Wt == 128
  int tiles_w, tiles_h, n_tiles, total_tiles;
  const float* bias;
  const __nv_bfloat16* residual;
  double* gn_stats;   // [N][8][2] or null
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int kBTileBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = BLOCK_N == 64 ? 7 : (BLOCK_N == 128 ? 5 : 4);
  static constexpr int kStoreBufs = BLOCK_N == 256 ? 1 : 2;      // output staging slabs (ping-pong when they fit)
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kTailBytes = 256 /*barriers*/ + 2 * BLOCK_N * 4 /*bias*/ + kEpiWarps * 16 * 4 /*stats*/ + 64;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStoreBufs * kSlabBytes + kTailBytes;
};

// GPT = GroupNorm groups covered by one N-tile (8 when Cout == BLOCK_N, 4 when Cout == 2*BLOCK_N, 0 = no statistics)
template <int BLOCK_N, int GPT>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                  const ConvParams p) {
  using C = Cfg<BLOCK_N>;
  constexpr int STAGES = C::kStages;
  constexpr int NBUF = C::kStoreBufs;
  constexpr int NCHUNK = BLOCK_N / 32;                 // 32-column accumulator chunks per tile
  constexpr int CPW = NCHUNK / 2;                      // chunks per epilogue warp (warps split even / odd chunks)
  constexpr int CPGT = GPT > 0 ? BLOCK_N / GPT : 32;   // columns per group inside the tile
  constexpr int GIC = CPGT < 32 ? 32 / CPGT : 1;       // groups inside one 32-column chunk
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base;
  const uint32_t b_smem = base + STAGES * kATileBytes;
  const uint32_t o_smem = base + STAGES * C::kStageBytes;
  const uint32_t bar_base = o_smem + NBUF * kSlabBytes;
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  uint8_t* gtail = gbase + STAGES * C::kStageBytes + NBUF * kSlabBytes + 256;
  float* s_bias = reinterpret_cast<float*>(gtail);                       // [2][BLOCK_N]
  float* s_stats = reinterpret_cast<float*>(gtail + 2 * BLOCK_N * 4);    // [8 warps][16]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gtail + 2 * BLOCK_N * 4 + kEpiWarps * 16 * 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cpt = p.chunks0 + p.chunks1;       // 64-channel chunks per tap
  const int num_kb = p.taps * cpt;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int img = m_tile / tiles_per_img;
        const int rem = m_tile - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.R;
        const int w0 = (rem % p.tiles_w) * p.Wt;
        int tap = 0, chunk = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), C::kStageBytes);
          const CUtensorMap* ma = chunk < p.chunks0 ? &map_a0 : &map_a1;
          const int c0 = (chunk < p.chunks0 ? chunk : chunk - p.chunks0) * kBlockK;
          const uint32_t dst_a = a_smem + stage * kATileBytes;
          if (p.mode == 0) {
            const int ky = tap / p.KW, kx = tap - ky * p.KW;
            tma_load_5d(dst_a, ma, full_bar(stage), c0, w0 + kx - p.pad_w, h0 + ky - p.pad_h, img, 0);
          } else {
            tma_load_5d(dst_a, ma, full_bar(stage), c0, tap & 1, w0, tap >> 1, h0);
          }
          tma_load_2d(b_smem + stage * C::kBTileBytes, &map_b, full_bar(stage), kb * kBlockK, n_tile * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          if (++chunk == cpt) { chunk = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = umma_desc_sw128(a_smem + stage * kATileBytes);
          const uint64_t bdesc = umma_desc_sw128(b_smem + stage * C::kBTileBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 (32 B) along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                      // frees the smem slot when the MMAs retire
          if (kb == num_kb - 1) umma_commit(tfull_bar(as));   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // TMEM lane quarter q = warp % 4 (hardware rule); the two warps of a quarter split the 32-column chunks
    // (even / odd).  Output goes registers -> swizzled smem slab (64 channels) -> one TMA store per slab, which
    // also clips rows / columns outside the image.  GroupNorm partial sums stay in registers across the tiles
    // of one (image, N-tile) and are flushed once.
    const int ew = warp - 2;                  // 0..7
    const int et = threadIdx.x - 64;          // 0..255
    const int quarter = warp & 3;
    const int half = ew >> 2;                 // which chunk parity this warp owns
    const int row = quarter * 32 + lane;      // accumulator row = pixel within the tile
    const int rr = row / p.Wt, ww = row - rr * p.Wt;
    const bool issuer = (et == 0);
    float st_s[CPW > 0 ? CPW : 1][GIC], st_q[CPW > 0 ? CPW : 1][GIC];
#pragma unroll
    for (int a = 0; a < (CPW > 0 ? CPW : 1); ++a)
#pragma unroll
      for (int b = 0; b < GIC; ++b) st_s[a][b] = st_q[a][b] = 0.f;
    int st_img = -1, st_ntile = 0;
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
        atomicAdd(p.gn_stats + (long)st_img * 16 + st_ntile * 2 * GPT + et, (double)sv);
      }
      named_bar_sync(2, kEpiThreads);
    };

    int iter = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.R, w0 = (rem % p.tiles_w) * p.Wt;
      const int h = h0 + rr, w = w0 + ww;
      const bool valid = (h < p.H) && (w < p.W);
      const int n0 = n_tile * BLOCK_N;
      if (GPT > 0 && (img != st_img || n_tile != st_ntile)) {
        flush_stats();
        st_img = img;
        st_ntile = n_tile;
      }
      float* bias_s = s_bias + as * BLOCK_N;
      for (int i = et; i < BLOCK_N; i += kEpiThreads) bias_s[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
      // (the slab barrier below also publishes the bias)

      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const long pix = ((long)img * p.H + h) * p.W + w;
      const __nv_bfloat16* rrow = p.residual ? p.residual + pix * p.Cout + n0 : nullptr;
#pragma unroll
      for (int slab = 0; slab < (BLOCK_N + 63) / 64; ++slab) {
        const uint32_t buf = o_smem + (slab_count % NBUF) * kSlabBytes;
        ++slab_count;
        if (issuer) tma_store_wait_read<NBUF - 1>();      // the store that last used this buffer has read it
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
        if (rrow != nullptr && valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(rrow + c) + q);
            const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = fd_unpack_bf16(rw[e]);
              v[q * 8 + e * 2] += f.x;
              v[q * 8 + e * 2 + 1] += f.y;
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
          mbar_arrive(tempty_bar(as));          // all TMEM reads of this tile are done
        }
        named_bar_sync(1, kEpiThreads);
        if (issuer) {
          if (p.mode == 0)
            tma_store_5d(&map_out, buf, n0 + slab * 64, w0, h0, img, 0);
          else
            tma_store_5d(&map_out, buf, n0 + slab * 64, w0, h0, 0, 0);
          tma_store_commit();
        }
      }
    }
    flush_stats();
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

struct TileShape {
  int R, Wt;
};

// R x Wt = 128 covering an H x W image with the least padding waste (ties -> widest tile)
TileShape pick_tile(int H, int W) {
  TileShape best{1, 128};
  long best_cost = -1;
  for (int wt = 128; wt >= 8; wt >>= 1) {
    const int r = 128 / wt;
    const long cost = (long)((W + wt - 1) / wt) * wt * (long)((H + r - 1) / r) * r;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = TileShape{r, wt};
    }
  }
  return best;
}

template <int BLOCK_N, int GPT>
int launch(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb, const CUtensorMap& mo,
           const ConvParams& p, int sms, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    FD_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, GPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 C::kSmemBytes));
    attr_set = true;
  }
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  conv_igemm_kernel<BLOCK_N, GPT><<<grid, kThreads, C::kSmemBytes, st>>>(ma0, ma1, mb, mo, p);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

}  // namespace

extern "C" {

int fd_conv_igemm(const void* src0, int C0, const void* src1, int C1, const void* wpacked, const float* bias,
                  const void* residual, void* out, double* gn_stats, int N, int H, int W, int Cout, int KH, int KW,
                  int pad_h, int pad_w, int mode, void* stream) {
  FD_REQUIRE(src0 && wpacked && out, "conv_igemm: null pointer");
  FD_REQUIRE(N > 0 && H > 0 && W > 0, "conv_igemm: bad geometry N=%d H=%d W=%d", N, H, W);
  FD_REQUIRE(C0 > 0 && C0 % 64 == 0 && C1 >= 0 && C1 % 64 == 0, "conv_igemm: C0=%d C1=%d must be multiples of 64", C0, C1);
  FD_REQUIRE(src1 != nullptr || C1 == 0, "conv_igemm: C1 > 0 needs src1");
  FD_REQUIRE(Cout > 0 && Cout % 64 == 0, "conv_igemm: Cout=%d must be a multiple of 64", Cout);
  FD_REQUIRE(mode == 0 || mode == 1, "conv_igemm: mode %d", mode);
  FD_REQUIRE(mode == 0 || (C1 == 0 && gn_stats == nullptr), "conv_igemm: mode 1 takes one source and no statistics");
  FD_REQUIRE(KH >= 1 && KW >= 1 && KH * KW <= 64, "conv_igemm: bad kernel %dx%d", KH, KW);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    FD_CUDA(cudaGetDevice(&dev));
    FD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  ConvParams p{};
  p.Cout = Cout;
  p.mode = mode;
  p.chunks0 = C0 / 64;
  p.chunks1 = C1 / 64;
  p.bias = bias;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.gn_stats = gn_stats;
  CUtensorMap ma0, ma1, mb, mo;
  TileShape ts;
  if (mode == 0) {
    int n = N, h = H, w = W;
    if (KH == 1 && KW == 1 && pad_h == 0 && pad_w == 0 && gn_stats == nullptr) {
      // 1x1: no halo, so every pixel of the batch is one long row -> full 128-pixel tiles for any W
      FD_REQUIRE((long)N * H * W < (1L << 31), "conv_igemm: too many pixels");
      w = N * H * W;
      h = 1;
      n = 1;
    }
    p.N = n; p.H = h; p.W = w;
    p.KW = KW;
    p.taps = KH * KW;
    p.pad_h = pad_h;
    p.pad_w = pad_w;
    ts = pick_tile(h, w);
    const uint32_t box[5] = {64, (uint32_t)ts.Wt, (uint32_t)ts.R, 1, 1};
    {
      const uint64_t dims[5] = {(uint64_t)C0, (uint64_t)w, (uint64_t)h, (uint64_t)n, 1};
      const uint64_t str[4] = {(uint64_t)C0 * 2, (uint64_t)w * C0 * 2, (uint64_t)h * w * C0 * 2, (uint64_t)n * h * w * C0 * 2};
      if (int e = make_tmap_bf16(&ma0, src0, 5, dims, str, box)) return e;
    }
    if (C1 > 0) {
      const uint64_t dims[5] = {(uint64_t)C1, (uint64_t)w, (uint64_t)h, (uint64_t)n, 1};
      const uint64_t str[4] = {(uint64_t)C1 * 2, (uint64_t)w * C1 * 2, (uint64_t)h * w * C1 * 2, (uint64_t)n * h * w * C1 * 2};
      if (int e = make_tmap_bf16(&ma1, src1, 5, dims, str, box)) return e;
    } else {
      ma1 = ma0;
    }
  } else {
    // src is (N, 2H, 2W, C0); output rows of all images are merged (no halo -> tiles may span images)
    p.N = 1; p.H = N * H; p.W = W;
    p.KW = 2;
    p.taps = 4;
    ts = pick_tile(p.H, p.W);
    const uint64_t dims[5] = {(uint64_t)C0, 2, (uint64_t)W, 2, (uint64_t)N * H};
    const uint64_t str[4] = {(uint64_t)C0 * 2, (uint64_t)2 * C0 * 2, (uint64_t)2 * W * C0 * 2, (uint64_t)4 * W * C0 * 2};
    const uint32_t box[5] = {64, 1, (uint32_t)ts.Wt, 1, (uint32_t)ts.R};
    if (int e = make_tmap_bf16(&ma0, src0, 5, dims, str, box)) return e;
    ma1 = ma0;
  }
  p.R = ts.R;
  p.Wt = ts.Wt;
  p.tiles_w = (p.W + ts.Wt - 1) / ts.Wt;
  p.tiles_h = (p.H + ts.R - 1) / ts.R;
  const int block_n = (Cout % 256 == 0) ? 256 : ((Cout % 128 == 0) ? 128 : 64);
  p.n_tiles = Cout / block_n;
  const long total = (long)p.N * p.tiles_w * p.tiles_h * p.n_tiles;
  FD_REQUIRE(total < (1L << 31), "conv_igemm: too many tiles");
  p.total_tiles = (int)total;
  {
    const uint64_t K = (uint64_t)p.taps * (C0 + C1);
    const uint64_t dims[2] = {K, (uint64_t)Cout};
    const uint64_t str[1] = {K * 2};
    const uint32_t box[2] = {64, (uint32_t)block_n};
    if (int e = make_tmap_bf16(&mb, wpacked, 2, dims, str, box)) return e;
  }
  {
    // output (N, H, W, Cout) in the same (possibly flattened / merged) geometry: box = one 64-channel slab of a tile
    const uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N, 1};
    const uint64_t str[4] = {(uint64_t)Cout * 2, (uint64_t)p.W * Cout * 2, (uint64_t)p.H * p.W * Cout * 2,
                             (uint64_t)p.N * p.H * p.W * Cout * 2};
    const uint32_t box[5] = {64, (uint32_t)ts.Wt, (uint32_t)ts.R, 1, 1};
    if (int e = make_tmap_bf16(&mo, out, 5, dims, str, box)) return e;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool stats = gn_stats != nullptr;
  if (stats) FD_REQUIRE(p.n_tiles == 1 || p.n_tiles == 2, "conv_igemm: statistics need Cout in {64,128,256,512}");
  if (block_n == 256) {
    if (!stats) return launch<256, 0>(ma0, ma1, mb, mo, p, sms, st);
    return p.n_tiles == 1 ? launch<256, 8>(ma0, ma1, mb, mo, p, sms, st) : launch<256, 4>(ma0, ma1, mb, mo, p, sms, st);
  }
  if (block_n == 128) return stats ? launch<128, 8>(ma0, ma1, mb, mo, p, sms, st) : launch<128, 0>(ma0, ma1, mb, mo, p, sms, st);
  return stats ? launch<64, 8>(ma0, ma1, mb, mo, p, sms, st) : launch<64, 0>(ma0, ma1, mb, mo, p, sms, st);
}

}  // extern "C"
