// Strong augmentation, colour jitter of the mixed image.
//
// Reference call site: color_jitter, rsiseg/models/utils/dacs_transforms.py:56-85 —
//   denorm_(data, mean, std); data = kornia.augmentation.ColorJitter(s, s, s, s)(data); renorm_(...)
// kornia is third-party (not under /root/reference, version unpinned): this file implements the
// arithmetic of its 0.6-series ColorJitter as restated in oracle/strong_aug.py (PARITY UNPINNED):
//   0 brightness clamp(x + (f - 1), 0, 1)        1 contrast clamp(x * f, 0, 1)
//   2 saturation rgb->hsv, s = clamp(s * f, 0, 1), hsv->rgb
//   3 hue        rgb->hsv, h = fmod(h + 2 pi f, 2 pi), hsv->rgb
// applied in a drawn order, between the reference's de-normalisation ((x*std + mean) / 255) and
// re-normalisation ((x*255 - mean) / std). The factors and the order are drawn on the host.
//
// One elementwise pass, 12 B read + 12 B written per pixel; the reference makes ~40 ATen passes
// (two HSV round trips with stack/gather temporaries). Every fp32 operation is an explicit IEEE
// intrinsic in the operation order of the torch expressions, so the result follows the CPU
// restatement to the last bit wherever torch's CPU kernels are themselves exactly rounded.
#include <math.h>

#include "common.cuh"

namespace pfst {

constexpr int kCjThreads = 256;
constexpr int kCjMaxImages = 64;

struct CjImage {
  float f[4];        // brightness, contrast, saturation, hue factors
  int32_t order[4];  // transform indices in application order (0..3), -1 = skip
};

struct CjParams {
  const float* in;
  float* out;
  int64_t HW;
  int32_t n_img;
  int32_t denorm;    // 1: 'mean_std', 0: 'none'
  float mean[3], std[3];
  CjImage img[kCjMaxImages];
};

__device__ __forceinline__ float cj_clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// torch.remainder(a, b) for b > 0 (CPU kernel: fmod, then shift a negative result by b)
__device__ __forceinline__ float cj_remainder(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && m < 0.f) m = __fadd_rn(m, b);
  return m;
}

constexpr float kTwoPi = 6.2831855f;   // float32(2 * math.pi)
constexpr float kPi = 3.14159274f;     // float32(math.pi)

__device__ __forceinline__ void cj_rgb2hsv(float r, float g, float b, float& h, float& s, float& v) {
  float mx = r;
  int am = 0;
  if (g > mx) { mx = g; am = 1; }
  if (b > mx) { mx = b; am = 2; }
  const float mn = fminf(fminf(r, g), b);
  float d = __fsub_rn(mx, mn);
  v = mx;
  s = __fdiv_rn(d, __fadd_rn(mx, 1e-8f));
  if (d == 0.f) d = 1.f;
  const float rc = __fsub_rn(mx, r), gc = __fsub_rn(mx, g), bc = __fsub_rn(mx, b);
  float hh;
  if (am == 0) hh = __fsub_rn(bc, gc);
  else if (am == 1) hh = __fadd_rn(__fsub_rn(rc, bc), __fmul_rn(2.0f, d));
  else hh = __fadd_rn(__fsub_rn(gc, rc), __fmul_rn(4.0f, d));
  hh = __fdiv_rn(hh, d);
  hh = cj_remainder(__fdiv_rn(hh, 6.0f), 1.0f);
  h = __fmul_rn(kTwoPi, hh);
}

__device__ __forceinline__ void cj_hsv2rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float hn = __fdiv_rn(h, kTwoPi);
  const float h6 = __fmul_rn(hn, 6.0f);
  const float hi_f = cj_remainder(floorf(h6), 6.0f);
  const float f = __fsub_rn(cj_remainder(h6, 6.0f), hi_f);
  const float p = __fmul_rn(v, __fsub_rn(1.0f, s));
  const float q = __fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(f, s)));
  const float t = __fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(__fsub_rn(1.0f, f), s)));
  const int hi = (int)hi_f;
  // kornia's table: R = (v,q,p,p,t,v)[hi], G = (t,v,v,q,p,p)[hi], B = (p,p,t,v,v,q)[hi]
  switch (hi) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__device__ __forceinline__ void cj_pixel(const CjParams& q, const CjImage& im, float& r, float& g, float& b) {
  if (q.denorm) {
    r = __fdiv_rn(__fadd_rn(__fmul_rn(r, q.std[0]), q.mean[0]), 255.0f);
    g = __fdiv_rn(__fadd_rn(__fmul_rn(g, q.std[1]), q.mean[1]), 255.0f);
    b = __fdiv_rn(__fadd_rn(__fmul_rn(b, q.std[2]), q.mean[2]), 255.0f);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int op = im.order[k];
    if (op == 0) {
      const float db = __fsub_rn(im.f[0], 1.0f);
      r = cj_clamp01(__fadd_rn(r, db)); g = cj_clamp01(__fadd_rn(g, db)); b = cj_clamp01(__fadd_rn(b, db));
    } else if (op == 1) {
      r = cj_clamp01(__fmul_rn(r, im.f[1])); g = cj_clamp01(__fmul_rn(g, im.f[1])); b = cj_clamp01(__fmul_rn(b, im.f[1]));
    } else if (op == 2 || op == 3) {
      float h, s, v;
      cj_rgb2hsv(r, g, b, h, s, v);
      if (op == 2) s = cj_clamp01(__fmul_rn(s, im.f[2]));
      else h = fmodf(__fadd_rn(h, __fmul_rn(__fmul_rn(im.f[3], 2.0f), kPi)), kTwoPi);
      cj_hsv2rgb(h, s, v, r, g, b);
    }
  }
  if (q.denorm) {
    r = __fdiv_rn(__fsub_rn(__fmul_rn(r, 255.0f), q.mean[0]), q.std[0]);
    g = __fdiv_rn(__fsub_rn(__fmul_rn(g, 255.0f), q.mean[1]), q.std[1]);
    b = __fdiv_rn(__fsub_rn(__fmul_rn(b, 255.0f), q.mean[2]), q.std[2]);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kCjThreads)
color_jitter_kernel(const CjParams q) {
  const int img = blockIdx.y;
  const CjImage& im = q.img[img];
  const float* __restrict__ src = q.in + (int64_t)img * 3 * q.HW;
  float* __restrict__ dst = q.out + (int64_t)img * 3 * q.HW;
  const int64_t units = q.HW / VEC;
  for (int64_t u = (int64_t)blockIdx.x * kCjThreads + threadIdx.x; u < units;
       u += (int64_t)gridDim.x * kCjThreads) {
    const int64_t px = u * VEC;
    if (VEC == 4) {
      float4 r = ldg_stream_f4(src + px), g = ldg_stream_f4(src + q.HW + px), b = ldg_stream_f4(src + 2 * q.HW + px);
      cj_pixel(q, im, r.x, g.x, b.x);
      cj_pixel(q, im, r.y, g.y, b.y);
      cj_pixel(q, im, r.z, g.z, b.z);
      cj_pixel(q, im, r.w, g.w, b.w);
      stg_f4(dst + px, r);
      stg_f4(dst + q.HW + px, g);
      stg_f4(dst + 2 * q.HW + px, b);
    } else {
      float r = src[px], g = src[q.HW + px], b = src[2 * q.HW + px];
      cj_pixel(q, im, r, g, b);
      dst[px] = r; dst[q.HW + px] = g; dst[2 * q.HW + px] = b;
    }
  }
}

}  // namespace pfst

extern "C" int pfst_color_jitter(const float* in, float* out, int64_t n_images, int64_t HW,
                                 const float* factors_host, const int32_t* order_host,
                                 const float* mean_host, const float* std_host, int32_t denorm,
                                 void* stream) {
  using namespace pfst;
  if (n_images < 0 || HW < 0 || (denorm != 0 && denorm != 1)) return PFST_ERR_INVALID_ARG;
  if (n_images == 0 || HW == 0) return PFST_OK;
  if (!in || !out || !factors_host || !order_host) return PFST_ERR_INVALID_ARG;
  if (denorm && (!mean_host || !std_host)) return PFST_ERR_INVALID_ARG;
  for (int64_t i = 0; i < 4 * n_images; ++i) {
    if (!(factors_host[i] == factors_host[i])) return PFST_ERR_INVALID_ARG;
    if (order_host[i] < -1 || order_host[i] > 3) return PFST_ERR_INVALID_ARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec4 = (HW % 4 == 0) && aligned16(in) && aligned16(out);
  for (int64_t b0 = 0; b0 < n_images; b0 += kCjMaxImages) {
    CjParams q;
    q.n_img = (int)((n_images - b0) < kCjMaxImages ? (n_images - b0) : kCjMaxImages);
    q.in = in + b0 * 3 * HW;
    q.out = out + b0 * 3 * HW;
    q.HW = HW;
    q.denorm = denorm;
    for (int c = 0; c < 3; ++c) {
      q.mean[c] = denorm ? mean_host[c] : 0.f;
      q.std[c] = denorm ? std_host[c] : 1.f;
    }
    for (int i = 0; i < kCjMaxImages; ++i)
      for (int k = 0; k < 4; ++k) {
        const bool live = i < q.n_img;
        q.img[i].f[k] = live ? factors_host[(b0 + i) * 4 + k] : 1.f;
        q.img[i].order[k] = live ? order_host[(b0 + i) * 4 + k] : -1;
      }
    const int64_t units = HW / (vec4 ? 4 : 1);
    int64_t gx = (units + kCjThreads - 1) / kCjThreads;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)q.n_img);
    if (vec4) color_jitter_kernel<4><<<grid, kCjThreads, 0, s>>>(q);
    else color_jitter_kernel<1><<<grid, kCjThreads, 0, s>>>(q);
    PFST_CHECK_LAUNCH("pfst_color_jitter");
  }
  return PFST_OK;
}
