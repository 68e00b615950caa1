// StrongAugmentation — the photometric distortion that produces `target_img_strong_aug`
// (second half).
//
// Reference: rsiseg/datasets/pipelines/transforms.py:1062-1145 (every shipped dataset config):
//   convert(img, alpha, beta) = uint8(clip(float32(img) * alpha + beta, 0, 255))       :1075-1079
//   brightness = convert(beta)            contrast = convert(alpha)                     :1081-1098
//   saturation = bgr2hsv -> S = convert(S, alpha) -> hsv2bgr                            :1100-1110
//   hue        = bgr2hsv -> H = (H + delta) % 180 -> hsv2bgr                            :1112-1121
// applied one after the other on the uint8 HWC image (each hsv step is a lossy round trip through
// uint8), in an order and with parameters drawn on the host. mmcv.bgr2hsv / hsv2bgr are
// cv2.cvtColor on uint8 images; their arithmetic is restated from OpenCV 4.13 and pinned
// bit-exactly against it over every possible colour (oracle/strong_aug.py):
//   * BGR->HSV: integer, 12-bit fixed-point reciprocal tables (kept in __constant__ memory);
//   * HSV->BGR: fp32 with separately rounded operations except the two fused multiply-adds of the
//     reference build, result*255 TRUNCATED inside the 32-pixel SIMD blocks of a row and ROUNDED
//     half-to-even in the row's scalar tail (columns >= W - W % simd).
// Every fp32 operation below is an explicit IEEE intrinsic so that nvcc cannot contract or reorder.
//
// One pass: 3 B read + 3 B written per pixel, all distortions of an image chained in registers
// (the reference makes up to four full passes plus two colour-space round trips on the CPU).
#include "common.cuh"
#include "hsv_tables.cuh"

namespace pfst {

constexpr int kSaThreads = 256;
constexpr int kSaMaxImages = 64;   // per-image op lists travel in the launch parameters
constexpr int kSaMaxOps = 4;       // brightness, contrast, saturation, hue (contrast moves, never doubles)

struct SaOps {
  int32_t code[kSaMaxOps];   // PFST_SA_* (0 = none)
  float p0[kSaMaxOps];       // alpha (convert / saturation) or hue delta
  float p1[kSaMaxOps];       // beta (convert)
};

struct SaParams {
  const uint8_t* in;
  uint8_t* out;
  int64_t pixels;      // H * W per image
  int32_t W;
  int32_t n_img;
  int32_t simd;        // SIMD block width of the reference's HSV->BGR (0: whole rows vectorised)
  SaOps ops[kSaMaxImages];
};

// uint8(clip(float32(x) * alpha + beta, 0, 255)), truncating
__device__ __forceinline__ int sa_convert(int x, float alpha, float beta) {
  float f = __fadd_rn(__fmul_rn((float)x, alpha), beta);
  f = fminf(fmaxf(f, 0.f), 255.f);
  return __float2int_rz(f);
}

__device__ __forceinline__ void sa_bgr2hsv(int b, int g, int r, int& h, int& s, int& v) {
  v = max(max(b, g), r);
  const int diff = v - min(min(b, g), r);
  const int vr = v == r ? -1 : 0;
  const int vg = v == g ? -1 : 0;
  s = (diff * c_sdiv[v] + (1 << 11)) >> 12;
  h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
  h = (h * c_hdiv[diff] + (1 << 11)) >> 12;     // arithmetic shift: h may be negative here
  h += h < 0 ? 180 : 0;
}

__device__ __forceinline__ int sa_to_u8(float x, bool tail) {
  const float f = __fmul_rn(x, 255.0f);
  const int i = tail ? __float2int_rn(f) : __float2int_rz(f);
  return min(max(i, 0), 255);
}

__device__ __forceinline__ void sa_hsv2bgr(int h, int s, int v, bool tail, int& b, int& g, int& r) {
  const float fv = __fmul_rn((float)v, 1.0f / 255.0f);
  if (s == 0) {
    b = g = r = sa_to_u8(fv, tail);
    return;
  }
  const float fs = __fmul_rn((float)s, 1.0f / 255.0f);
  const float fh = __fmul_rn((float)h, 6.0f / 180.0f);
  int sector = __float2int_rd(fh);
  const float fr = __fsub_rn(fh, (float)sector);
  sector = min(max(sector, 0), 5);
  const float t0 = fv;
  const float t1 = __fmul_rn(fv, __fsub_rn(1.0f, fs));
  const float t2 = __fmul_rn(fv, __fmaf_rn(-fs, fr, 1.0f));
  const float t3 = __fmul_rn(fv, __fmaf_rn(-fs, __fsub_rn(1.0f, fr), 1.0f));
  float fb, fg, frd;
  switch (sector) {   // OpenCV sector_data: {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0} -> (b,g,r)
    case 0: fb = t1; fg = t3; frd = t0; break;
    case 1: fb = t1; fg = t0; frd = t2; break;
    case 2: fb = t3; fg = t0; frd = t1; break;
    case 3: fb = t0; fg = t2; frd = t1; break;
    case 4: fb = t0; fg = t1; frd = t3; break;
    default: fb = t2; fg = t1; frd = t0; break;
  }
  b = sa_to_u8(fb, tail);
  g = sa_to_u8(fg, tail);
  r = sa_to_u8(frd, tail);
}

__device__ __forceinline__ void sa_pixel(const SaOps& o, bool tail, int& b, int& g, int& r) {
#pragma unroll
  for (int k = 0; k < kSaMaxOps; ++k) {
    const int code = o.code[k];
    if (code == PFST_SA_CONVERT) {
      b = sa_convert(b, o.p0[k], o.p1[k]);
      g = sa_convert(g, o.p0[k], o.p1[k]);
      r = sa_convert(r, o.p0[k], o.p1[k]);
    } else if (code == PFST_SA_SATURATION || code == PFST_SA_HUE) {
      int h, s, v;
      sa_bgr2hsv(b, g, r, h, s, v);
      if (code == PFST_SA_SATURATION) {
        s = sa_convert(s, o.p0[k], 0.f);
      } else {
        h = (h + (int)o.p0[k]) % 180;          // python %: result takes the divisor's sign
        h += h < 0 ? 180 : 0;
      }
      sa_hsv2bgr(h, s, v, tail, b, g, r);
    }
  }
}

// VEC = 4: a thread owns four consecutive pixels = three aligned 32-bit words; VEC = 1: byte accesses
template <int VEC>
__global__ void __launch_bounds__(kSaThreads)
photometric_u8_kernel(const SaParams q) {
  const int img = blockIdx.y;
  const SaOps& o = q.ops[img];
  const int64_t units = q.pixels / VEC;
  const uint8_t* __restrict__ src = q.in + (int64_t)img * q.pixels * 3;
  uint8_t* __restrict__ dst = q.out + (int64_t)img * q.pixels * 3;
  const int tail_start = q.simd > 0 ? q.W - q.W % q.simd : q.W;
  for (int64_t u = (int64_t)blockIdx.x * kSaThreads + threadIdx.x; u < units;
       u += (int64_t)gridDim.x * kSaThreads) {
    const int64_t px = u * VEC;
    int col = (int)(px % q.W);
    if (VEC == 4) {
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + px * 3);
      uint32_t w[3] = {__ldg(s32), __ldg(s32 + 1), __ldg(s32 + 2)};
      uint8_t bytes[12];
      memcpy(bytes, w, 12);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int b = bytes[3 * k], g = bytes[3 * k + 1], r = bytes[3 * k + 2];
        sa_pixel(o, col >= tail_start, b, g, r);
        bytes[3 * k] = (uint8_t)b; bytes[3 * k + 1] = (uint8_t)g; bytes[3 * k + 2] = (uint8_t)r;
        if (++col == q.W) col = 0;
      }
      memcpy(w, bytes, 12);
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + px * 3);
      d32[0] = w[0]; d32[1] = w[1]; d32[2] = w[2];
    } else {
      int b = src[px * 3], g = src[px * 3 + 1], r = src[px * 3 + 2];
      sa_pixel(o, col >= tail_start, b, g, r);
      dst[px * 3] = (uint8_t)b; dst[px * 3 + 1] = (uint8_t)g; dst[px * 3 + 2] = (uint8_t)r;
    }
  }
}

}  // namespace pfst

extern "C" int pfst_photometric_u8(const uint8_t* in, uint8_t* out, int64_t n_images, int32_t H, int32_t W,
                                   const int32_t* op_codes_host, const float* op_params_host,
                                   int32_t simd_width, void* stream) {
  using namespace pfst;
  if (n_images < 0 || H < 1 || W < 1 || simd_width < 0) return PFST_ERR_INVALID_ARG;
  if (n_images == 0) return PFST_OK;
  if (!in || !out || !op_codes_host || !op_params_host) return PFST_ERR_INVALID_ARG;
  for (int64_t i = 0; i < n_images * kSaMaxOps; ++i) {
    const int32_t c = op_codes_host[i];
    if (c != 0 && c != PFST_SA_CONVERT && c != PFST_SA_SATURATION && c != PFST_SA_HUE) return PFST_ERR_INVALID_ARG;
    const float p0 = op_params_host[2 * i], p1 = op_params_host[2 * i + 1];
    if (!(p0 == p0) || !(p1 == p1)) return PFST_ERR_INVALID_ARG;
    if (c == PFST_SA_HUE && (p0 != (float)(int)p0 || p0 < -100000.f || p0 > 100000.f)) return PFST_ERR_INVALID_ARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t pixels = (int64_t)H * W;
  const bool vec4 = (pixels % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 3u) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 3u) == 0);
  for (int64_t b0 = 0; b0 < n_images; b0 += kSaMaxImages) {
    SaParams q;
    q.n_img = (int)((n_images - b0) < kSaMaxImages ? (n_images - b0) : kSaMaxImages);
    q.in = in + b0 * pixels * 3;
    q.out = out + b0 * pixels * 3;
    q.pixels = pixels;
    q.W = W;
    q.simd = simd_width;
    for (int i = 0; i < kSaMaxImages; ++i)
      for (int k = 0; k < kSaMaxOps; ++k) {
        const int64_t j = (b0 + i) * kSaMaxOps + k;
        const bool live = i < q.n_img;
        q.ops[i].code[k] = live ? op_codes_host[j] : 0;
        q.ops[i].p0[k] = live ? op_params_host[2 * j] : 0.f;
        q.ops[i].p1[k] = live ? op_params_host[2 * j + 1] : 0.f;
      }
    const int64_t units = pixels / (vec4 ? 4 : 1);
    int64_t gx = (units + kSaThreads - 1) / kSaThreads;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)q.n_img);
    if (vec4) photometric_u8_kernel<4><<<grid, kSaThreads, 0, s>>>(q);
    else photometric_u8_kernel<1><<<grid, kSaThreads, 0, s>>>(q);
    PFST_CHECK_LAUNCH("pfst_photometric_u8");
  }
  return PFST_OK;
}
