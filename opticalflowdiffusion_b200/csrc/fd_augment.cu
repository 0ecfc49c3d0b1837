// Training-time augmentation on the GPU: the arithmetic of the reference's Augmentor
// (algorithms/diffusion_animation/augmentation.py:6-76), which applies torchvision transforms item by item from
// Python (7.5 ms of host-bound launches per batch of 8 at 368x768, measured), as four launches over the whole batch.
// The random DECISIONS stay on the host (the same torch / torchvision parameter samplers, in the same order, so a
// given seed gives the same augmentation); the kernels take them as small per-frame / per-item parameter tables:
//
//   frame table  (2B rows: item*2 + {0: img, 1: tgt})   int  [8] = {jitter_on, op0, op1, op2, op3, gray_on, blur_on, -}
//                                                      float [8] = {brightness, contrast, saturation, hue, k_edge, k_center, -, -}
//   item table   (B rows)                               int  [8] = {hflip, vflip, crop_on, top, left, crop_h, crop_w, -}
//
//   (1) contrast_mean   mean grayscale of each jittered frame after the ops that precede `contrast` in its permutation
//   (2) photometric     ColorJitter chain (torchvision adjust_brightness / contrast / saturation / hue, float images
//                       clamped to [0,1] after every op) + Grayscale(3)                     -> frames (B,2,3,H,W)
//   (3) blur3           GaussianBlur(3, sigma) with reflect padding, for the frames that drew it
//   (4) geometric       hflip (negates the LAST flow channel), vflip (negates the second-last), RandomResizedCrop +
//                       bilinear resize back (align_corners = False, no antialias; flow scaled by crop/size) over the
//                       8 channels (img, tgt, flow) of an item
#include "fd_common.cuh"

namespace {

constexpr int kFrameInts = 8, kFrameFloats = 8, kItemInts = 8;

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }

// torchvision _blend for float images: (ratio * a + (1 - ratio) * b).clamp(0, 1)
__device__ __forceinline__ float blend(float a, float b, float ratio) { return clamp01(ratio * a + (1.f - ratio) * b); }

// torchvision adjust_hue on one pixel (_rgb2hsv -> h = (h + f) % 1 -> _hsv2rgb)
__device__ __forceinline__ void hue_shift(float& r, float& g, float& b, float f) {
  const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
  const bool eqc = maxc == minc;
  const float cr = maxc - minc;
  const float s = cr / (eqc ? 1.f : maxc);
  const float div = eqc ? 1.f : cr;
  const float rc = (maxc - r) / div, gc = (maxc - g) / div, bc = (maxc - b) / div;
  const float hr = (maxc == r) ? (bc - gc) : 0.f;
  const float hg = ((maxc == g) && (maxc != r)) ? (2.f + rc - bc) : 0.f;
  const float hb = ((maxc != g) && (maxc != r)) ? (4.f + gc - rc) : 0.f;
  float h = fmodf((hr + hg + hb) / 6.f + 1.f, 1.f);
  h = fmodf(h + f, 1.f);
  if (h < 0.f) h += 1.f;                       // python-style modulo (the result takes the sign of the divisor)
  const float v = maxc;
  const float i_f = floorf(h * 6.f);
  const float fr = h * 6.f - i_f;
  int i = ((int)i_f) % 6;
  if (i < 0) i += 6;
  const float p = clamp01(v * (1.f - s));
  const float q = clamp01(v * (1.f - s * fr));
  const float t = clamp01(v * (1.f - s * (1.f - fr)));
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// ops of the permutation up to (not including) position `stop`; contrast uses `mean`
__device__ __forceinline__ void jitter_ops(float& r, float& g, float& b, const int* __restrict__ fi, const float* __restrict__ ff,
                                           int stop, float mean) {
  for (int k = 0; k < stop; ++k) {
    const int op = fi[1 + k];
    if (op == 0) {
      const float f = ff[0];
      r = blend(r, 0.f, f); g = blend(g, 0.f, f); b = blend(b, 0.f, f);
    } else if (op == 1) {
      const float f = ff[1];
      r = blend(r, mean, f); g = blend(g, mean, f); b = blend(b, mean, f);
    } else if (op == 2) {
      const float f = ff[2];
      const float l = gray_of(r, g, b);
      r = blend(r, l, f); g = blend(g, l, f); b = blend(b, l, f);
    } else {
      hue_shift(r, g, b, ff[3]);
    }
  }
}

__device__ __forceinline__ const float* frame_ptr(const float* img, const float* tgt, int frame, long HW) {
  return ((frame & 1) ? tgt : img) + (long)(frame >> 1) * 3 * HW;
}

__global__ void __launch_bounds__(256) aug_contrast_mean_kernel(const float* __restrict__ img, const float* __restrict__ tgt,
                                                                const int* __restrict__ fints, const float* __restrict__ ffloats,
                                                                float* __restrict__ means, long HW) {
  __shared__ float red[32];
  const int frame = blockIdx.y;
  const int* fi = fints + frame * kFrameInts;
  if (fi[0] == 0) return;
  int stop = 0;
  while (stop < 4 && fi[1 + stop] != 1) ++stop;          // position of `contrast` in the permutation
  const float* ff = ffloats + frame * kFrameFloats;
  const float* src = frame_ptr(img, tgt, frame, HW);
  float s[1] = {0.f};
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long)gridDim.x * blockDim.x) {
    float r = src[p], g = src[HW + p], b = src[2 * HW + p];
    jitter_ops(r, g, b, fi, ff, stop, 0.f);
    s[0] += gray_of(r, g, b);
  }
  fd_block_sum<1>(s, red);
  if (threadIdx.x == 0) atomicAdd(means + frame, s[0] / (float)HW);
}

__global__ void __launch_bounds__(256) aug_photometric_kernel(const float* __restrict__ img, const float* __restrict__ tgt,
                                                              const int* __restrict__ fints, const float* __restrict__ ffloats,
                                                              const float* __restrict__ means, float* __restrict__ out, long HW) {
  const int frame = blockIdx.y;
  const int* fi = fints + frame * kFrameInts;
  const float* ff = ffloats + frame * kFrameFloats;
  const float* src = frame_ptr(img, tgt, frame, HW);
  float* dst = out + (long)frame * 3 * HW;
  const bool jitter = fi[0] != 0, gray = fi[5] != 0;
  const float mean = jitter ? means[frame] : 0.f;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long)gridDim.x * blockDim.x) {
    float r = src[p], g = src[HW + p], b = src[2 * HW + p];
    if (jitter) jitter_ops(r, g, b, fi, ff, 4, mean);
    if (gray) r = g = b = gray_of(r, g, b);
    dst[p] = r;
    dst[HW + p] = g;
    dst[2 * HW + p] = b;
  }
}

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(256) aug_blur3_kernel(const float* __restrict__ in, const int* __restrict__ fints,
                                                        const float* __restrict__ ffloats, float* __restrict__ out, int H, int W) {
  const int frame = blockIdx.y;
  if (fints[frame * kFrameInts + 6] == 0) return;
  const float ke = ffloats[frame * kFrameFloats + 4], kc = ffloats[frame * kFrameFloats + 5];
  const float k1[3] = {ke, kc, ke};
  const long HW = (long)H * W;
  const long total = 3 * HW;
  const float* src = in + (long)frame * 3 * HW;
  float* dst = out + (long)frame * 3 * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i / HW);
    const long p = i - (long)c * HW;
    const int y = (int)(p / W), x = (int)(p - (long)y * W);
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = reflect1(y + dy - 1, H);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = reflect1(x + dx - 1, W);
        acc += (k1[dy] * k1[dx]) * src[(long)c * HW + (long)yy * W + xx];
      }
    }
    dst[i] = acc;
  }
}

// channel ch (0..7) of item `item` at source pixel (y, x), after the photometric stage
__device__ __forceinline__ float geo_src(const float* __restrict__ fa, const float* __restrict__ fb,
                                         const float* __restrict__ flow, const int* __restrict__ fints, int item, int ch,
                                         long HW, long p) {
  if (ch < 6) {
    const int frame = item * 2 + ch / 3;
    const float* base = fints[frame * kFrameInts + 6] ? fb : fa;
    return base[((long)frame * 3 + ch % 3) * HW + p];
  }
  return flow[((long)item * 2 + (ch - 6)) * HW + p];
}

__global__ void __launch_bounds__(256) aug_geometric_kernel(const float* __restrict__ fa, const float* __restrict__ fb,
                                                            const float* __restrict__ flow, const int* __restrict__ fints,
                                                            const int* __restrict__ iints, float* __restrict__ out_img,
                                                            float* __restrict__ out_tgt, float* __restrict__ out_flow, int H,
                                                            int W) {
  const int item = blockIdx.y;
  const int* ii = iints + item * kItemInts;
  const bool hflip = ii[0] != 0, vflip = ii[1] != 0, crop = ii[2] != 0;
  const int top = ii[3], left = ii[4], ch_ = ii[5], cw_ = ii[6];
  const long HW = (long)H * W;
  // flow sign / scale: hflip negates the last flow channel, vflip the second-last, the crop scales (ch6, ch7) by
  // (crop_h / H, crop_w / W) BEFORE resampling (augmentation.py:34-50)
  float s6 = vflip ? -1.f : 1.f, s7 = hflip ? -1.f : 1.f;
  if (crop) {
    s6 *= (float)ch_ / (float)H;
    s7 *= (float)cw_ / (float)W;
  }
  const float rh = crop ? (float)ch_ / (float)H : 1.f, rw = crop ? (float)cw_ / (float)W : 1.f;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long)gridDim.x * blockDim.x) {
    const int y = (int)(p / W), x = (int)(p - (long)y * W);
    // output pixel -> position inside the crop of the flipped image (torch upsample_bilinear2d, align_corners = False)
    int y0 = y, x0 = x, y1 = y, x1 = x;
    float ly = 0.f, lx = 0.f;
    if (crop) {
      const float sy = fmaxf(rh * ((float)y + 0.5f) - 0.5f, 0.f), sx = fmaxf(rw * ((float)x + 0.5f) - 0.5f, 0.f);
      const int iy = (int)sy, ix = (int)sx;
      ly = sy - (float)iy;
      lx = sx - (float)ix;
      y0 = top + iy;
      x0 = left + ix;
      y1 = y0 + (iy < ch_ - 1 ? 1 : 0);
      x1 = x0 + (ix < cw_ - 1 ? 1 : 0);
    }
    // flipped image -> stored image
    if (vflip) { y0 = H - 1 - y0; y1 = H - 1 - y1; }
    if (hflip) { x0 = W - 1 - x0; x1 = W - 1 - x1; }
    const long p00 = (long)y0 * W + x0, p01 = (long)y0 * W + x1, p10 = (long)y1 * W + x0, p11 = (long)y1 * W + x1;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float v;
      if (crop) {
        const float a = geo_src(fa, fb, flow, fints, item, ch, HW, p00), b = geo_src(fa, fb, flow, fints, item, ch, HW, p01);
        const float c = geo_src(fa, fb, flow, fints, item, ch, HW, p10), d = geo_src(fa, fb, flow, fints, item, ch, HW, p11);
        const float sc = ch == 6 ? s6 : (ch == 7 ? s7 : 1.f);
        v = (1.f - ly) * ((1.f - lx) * (a * sc) + lx * (b * sc)) + ly * ((1.f - lx) * (c * sc) + lx * (d * sc));
      } else {
        v = geo_src(fa, fb, flow, fints, item, ch, HW, p00);
        if (ch == 6) v *= s6;
        if (ch == 7) v *= s7;
      }
      if (ch < 3) out_img[((long)item * 3 + ch) * HW + p] = v;
      else if (ch < 6) out_tgt[((long)item * 3 + ch - 3) * HW + p] = v;
      else out_flow[((long)item * 2 + ch - 6) * HW + p] = v;
    }
  }
}

int pgrid(long HW) {
  long b = (HW + 255) / 256;
  if (b > FD_NUM_SMS * 4) b = FD_NUM_SMS * 4;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int fd_aug_photometric(const float* img, const float* tgt, const int* frame_ints, const float* frame_floats, float* means,
                       float* frames_out, int B, int HW, void* stream) {
  FD_REQUIRE(img && tgt && frame_ints && frame_floats && means && frames_out && B > 0 && HW > 0, "aug_photometric: bad argument");
  FD_REQUIRE(2 * B <= 65535, "aug_photometric: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  FD_CUDA(cudaMemsetAsync(means, 0, 2 * B * sizeof(float), st));
  const dim3 grid(pgrid(HW), 2 * B);
  aug_contrast_mean_kernel<<<grid, 256, 0, st>>>(img, tgt, frame_ints, frame_floats, means, (long)HW);
  FD_LAUNCH_CHECK();
  aug_photometric_kernel<<<grid, 256, 0, st>>>(img, tgt, frame_ints, frame_floats, means, frames_out, (long)HW);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_aug_blur3(const float* frames, const int* frame_ints, const float* frame_floats, float* blurred, int B, int H, int W,
                 void* stream) {
  FD_REQUIRE(frames && frame_ints && frame_floats && blurred && B > 0 && H > 1 && W > 1, "aug_blur3: bad argument");
  const dim3 grid(pgrid(3L * H * W), 2 * B);
  aug_blur3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, frame_ints, frame_floats, blurred, H, W);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

int fd_aug_geometric(const float* frames, const float* blurred, const float* flow, const int* frame_ints, const int* item_ints,
                     float* out_img, float* out_tgt, float* out_flow, int B, int H, int W, void* stream) {
  FD_REQUIRE(frames && blurred && flow && frame_ints && item_ints && out_img && out_tgt && out_flow && B > 0 && H > 0 && W > 0,
             "aug_geometric: bad argument");
  FD_REQUIRE(B <= 65535, "aug_geometric: batch too large");
  const dim3 grid(pgrid((long)H * W), B);
  aug_geometric_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, blurred, flow, frame_ints, item_ints, out_img, out_tgt,
                                                              out_flow, H, W);
  FD_LAUNCH_CHECK();
  return FD_OK;
}

}  // extern "C"
