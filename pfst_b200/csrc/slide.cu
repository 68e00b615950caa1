// Test-time sliding-window accumulation and flip (the tensor arithmetic of
// EncoderDecoder.slide_inference / inference, rsiseg/models/segmentors/encoder_decoder.py:220-263, 312-324).
//
// Reference, per window:   preds += F.pad(crop_seg_logit, ...)   (a full-size padded copy + a full-size add)
//                          count_mat[:, :, y1:y2, x1:x2] += 1
// and at the end:          preds = preds / count_mat ; softmax ; output.flip(dims) per flip direction.
// Here a window touches only its own region (pfst_slide_add: 12 B per window element instead of three
// full-size passes), the count matrix never exists — the windows are a product of row and column
// intervals, so count(y, x) = cnt_y[y] * cnt_x[x], two small host-built vectors — and the division and the
// flips are one pass (pfst_slide_finalize). Bit-exact with the reference: the same fp32 additions in the
// same window order (x + 0 outside a window is x), an IEEE division by an exactly representable count,
// and a flip commutes with the per-pixel soft-max that follows.
#include <math.h>

#include "common.cuh"
#include "exp_exact.cuh"

namespace pfst {

constexpr int kSlThreads = 256;

// preds[b, c, y1 + i, x1 + j] += crop[b, c, i, j]; one thread per 4 consecutive columns when aligned
template <bool VEC4>
__global__ void __launch_bounds__(kSlThreads)
slide_add_kernel(float* __restrict__ preds, const float* __restrict__ crop, int planes, int H, int W, int y1,
                 int x1, int ch, int cw) {
  const int cols = VEC4 ? cw / 4 : cw;
  const int64_t n = (int64_t)planes * ch * cols;
  for (int64_t i = (int64_t)blockIdx.x * kSlThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSlThreads) {
    const int j = (int)(i % cols);
    const int64_t t = i / cols;
    const int r = (int)(t % ch);
    const int64_t p = t / ch;
    const int64_t src = (p * ch + r) * cw, dst = (p * H + y1 + r) * W + x1;
    if (VEC4) {
      const float4 a = *reinterpret_cast<const float4*>(crop + src + 4 * j);
      float4 v = *reinterpret_cast<float4*>(preds + dst + 4 * j);
      v.x = __fadd_rn(v.x, a.x); v.y = __fadd_rn(v.y, a.y); v.z = __fadd_rn(v.z, a.z); v.w = __fadd_rn(v.w, a.w);
      *reinterpret_cast<float4*>(preds + dst + 4 * j) = v;
    } else {
      preds[dst + j] = __fadd_rn(preds[dst + j], crop[src + j]);
    }
  }
}

// out[b, c, y, x] = preds[b, c, ys, xs] / (cnt_y[ys] * cnt_x[xs]),  (ys, xs) = (flip_v ? H-1-y : y, flip_h ? W-1-x : x)
__global__ void __launch_bounds__(kSlThreads)
slide_finalize_kernel(const float* __restrict__ preds, const float* __restrict__ cnt_y,
                      const float* __restrict__ cnt_x, int planes, int H, int W, int flip_h, int flip_v,
                      float* __restrict__ out) {
  const int64_t n = (int64_t)planes * H * W;
  for (int64_t i = (int64_t)blockIdx.x * kSlThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSlThreads) {
    const int x = (int)(i % W);
    const int64_t t = i / W;
    const int y = (int)(t % H);
    const int64_t p = t / H;
    const int ys = flip_v ? H - 1 - y : y, xs = flip_h ? W - 1 - x : x;
    const float c = (cnt_y && cnt_x) ? __fmul_rn(cnt_y[ys], cnt_x[xs]) : 1.f;
    const float v = preds[(p * H + ys) * W + xs];
    out[i] = (cnt_y && cnt_x) ? __fdiv_rn(v, c) : v;
  }
}

// ---- aug_test (encoder_decoder.py:355-373): seg_logit = sum_i softmax(logits_i); seg_logit /= n; argmax ----
// acc[b, c, p] (+)= softmax over c of logits[b, :, p], as torch's CUDA soft-max computes it along a
// non-innermost dim (max, sum of expf(x - max) in class order, expf(x - max) / sum with an IEEE division);
// first != 0 overwrites acc instead of adding (the reference starts from the first augmentation's output).
__global__ void __launch_bounds__(kSlThreads)
softmax_accum_kernel(const float* __restrict__ logits, float* __restrict__ acc, int64_t n_img, int C, int64_t pixels,
                     int first) {
  const int64_t total = n_img * pixels;
  for (int64_t i = (int64_t)blockIdx.x * kSlThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kSlThreads) {
    const int64_t b = i / pixels, p = i - b * pixels;
    const float* x = logits + b * C * pixels + p;
    float* a = acc + b * C * pixels + p;
    float m = x[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, x[c * pixels]);
    if (x[0] != x[0]) m = x[0];
    bool nan = false;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = x[c * pixels];
      nan |= v != v;
      const ExpParts e = exp_split(v - m);
      s = __fmaf_rn(e.scale, e.mant, s);
    }
    for (int c = 0; c < C; ++c) {
      float q = __fdiv_rn(exp_exact(x[c * pixels] - m), s);
      if (nan) q = __int_as_float(0x7fc00000);
      a[c * pixels] = first ? q : __fadd_rn(a[c * pixels], q);
    }
  }
}

// pred[b, p] = argmax_c (acc[b, c, p] / n): first maximum wins, a NaN counts as the maximum (torch.argmax)
__global__ void __launch_bounds__(kSlThreads)
div_argmax_kernel(const float* __restrict__ acc, int64_t n_img, int C, int64_t pixels, float n, int64_t* __restrict__ pred) {
  const int64_t total = n_img * pixels;
  for (int64_t i = (int64_t)blockIdx.x * kSlThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kSlThreads) {
    const int64_t b = i / pixels, p = i - b * pixels;
    const float* a = acc + b * C * pixels + p;
    float m = __fdiv_rn(a[0], n);
    int am = 0;
    for (int c = 1; c < C; ++c) {
      const float v = __fdiv_rn(a[c * pixels], n);
      if (m == m && (v > m || v != v)) { m = v; am = c; }
    }
    pred[i] = am;
  }
}

}  // namespace pfst

extern "C" {

int pfst_softmax_accum(const float* logits, float* acc, int64_t n_images, int32_t C, int64_t pixels, int32_t first,
                       void* stream) {
  if (!logits || !acc || n_images < 0 || C < 1 || pixels < 1) return PFST_ERR_INVALID_ARG;
  if (n_images == 0) return PFST_OK;
  int64_t blocks = (n_images * pixels + pfst::kSlThreads - 1) / pfst::kSlThreads;
  if (blocks > (int64_t)pfst::kNumSMs * 16) blocks = (int64_t)pfst::kNumSMs * 16;
  pfst::softmax_accum_kernel<<<(unsigned)blocks, pfst::kSlThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, acc, n_images, C, pixels, first);
  PFST_CHECK_LAUNCH("pfst_softmax_accum");
  return PFST_OK;
}

int pfst_div_argmax(const float* acc, int64_t n_images, int32_t C, int64_t pixels, float divisor, int64_t* pred,
                    void* stream) {
  if (!acc || !pred || n_images < 0 || C < 1 || pixels < 1) return PFST_ERR_INVALID_ARG;
  if (n_images == 0) return PFST_OK;
  int64_t blocks = (n_images * pixels + pfst::kSlThreads - 1) / pfst::kSlThreads;
  if (blocks > (int64_t)pfst::kNumSMs * 16) blocks = (int64_t)pfst::kNumSMs * 16;
  pfst::div_argmax_kernel<<<(unsigned)blocks, pfst::kSlThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, n_images, C, pixels, divisor, pred);
  PFST_CHECK_LAUNCH("pfst_div_argmax");
  return PFST_OK;
}

int pfst_slide_add(float* preds, const float* crop, int64_t B, int32_t C, int32_t H, int32_t W, int32_t y1,
                   int32_t x1, int32_t ch, int32_t cw, void* stream) {
  if (!preds || !crop || B < 0 || C < 1 || H < 1 || W < 1 || ch < 1 || cw < 1) return PFST_ERR_INVALID_ARG;
  if (y1 < 0 || x1 < 0 || y1 + ch > H || x1 + cw > W) return PFST_ERR_INVALID_ARG;
  if (B * C > 0x7fffffffll) return PFST_ERR_UNSUPPORTED;
  if (B == 0) return PFST_OK;
  const bool vec = cw % 4 == 0 && W % 4 == 0 && x1 % 4 == 0 && pfst::aligned16(preds) && pfst::aligned16(crop);
  const int64_t n = B * C * ch * (vec ? cw / 4 : cw);
  int64_t blocks = (n + pfst::kSlThreads - 1) / pfst::kSlThreads;
  if (blocks > (int64_t)pfst::kNumSMs * 16) blocks = (int64_t)pfst::kNumSMs * 16;
  auto k = vec ? pfst::slide_add_kernel<true> : pfst::slide_add_kernel<false>;
  k<<<(unsigned)blocks, pfst::kSlThreads, 0, static_cast<cudaStream_t>(stream)>>>(preds, crop, (int)(B * C), H, W, y1,
                                                                                 x1, ch, cw);
  PFST_CHECK_LAUNCH("pfst_slide_add");
  return PFST_OK;
}

int pfst_slide_finalize(const float* preds, const float* cnt_y, const float* cnt_x, int64_t B, int32_t C, int32_t H,
                        int32_t W, int32_t flip_h, int32_t flip_v, float* out, void* stream) {
  if (!preds || !out || B < 0 || C < 1 || H < 1 || W < 1 || ((cnt_y == nullptr) != (cnt_x == nullptr)))
    return PFST_ERR_INVALID_ARG;
  if (preds == out && (flip_h || flip_v)) return PFST_ERR_INVALID_ARG;     // a flip cannot run in place
  if (B * C > 0x7fffffffll) return PFST_ERR_UNSUPPORTED;
  if (B == 0) return PFST_OK;
  const int64_t n = B * C * H * W;
  int64_t blocks = (n + pfst::kSlThreads - 1) / pfst::kSlThreads;
  if (blocks > (int64_t)pfst::kNumSMs * 16) blocks = (int64_t)pfst::kNumSMs * 16;
  pfst::slide_finalize_kernel<<<(unsigned)blocks, pfst::kSlThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      preds, cnt_y, cnt_x, (int)(B * C), H, W, flip_h, flip_v, out);
  PFST_CHECK_LAUNCH("pfst_slide_finalize");
  return PFST_OK;
}

}  // extern "C"
