// P1-P3 — class prototypes: masked segment-reduce of decoder features over
// (pseudo-)labels, prototype EMA, feature-to-prototype distance loss + backward.
//
// north_star extension: the reference ships no prototype code. The
// anchor is PFGST.masked_feat_dist, rsiseg/models/uda/pfgst.py:168-177
//   mean( ||f1 - f2||_2 over channels [mask] )           with f2 = mu[label].
// Label maps are nearest-resampled to the feature grid exactly as the loss does
// (pfgst_loss.py:62).
//
// All kernels are HBM-bound streams over the NCHW feature tensor:
//   accumulate  4*D B/pixel read   (sums (C,D) + counts merged with one atomic per
//               (warp, class, channel): per-lane private accumulators in shared
//               memory, acc[class][lane] — bank = lane, no atomics in the hot loop)
//   distance    fwd 4*D B/pixel read, bwd 4*D read + 4*D written
// The pixel x prototype contraction has arithmetic intensity 2C/4 flop/B (1, 3,
// 16.5 for C = 2, 6, 33) and must hold 1e-5 relative in fp32, which rules out
// TF32/BF16 MMA: it stays an FFMA bandwidth kernel (north_star: tensor cores only
// when D and C make it a real contraction).
#include <math.h>

#include "common.cuh"

namespace pfst {

constexpr int kPrThreads = 256;
constexpr int kPrWarps = kPrThreads / 32;
constexpr int kPrMaxC = 128;
constexpr int kPrPixTile = 8192;   // label bytes staged per pass

__device__ __forceinline__ int pr_nearest(int dst, float scale, int in) {
  const int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}

// label of feature pixel p (0..h*w) of image b, 255 if outside [0,C) or masked out
__device__ __forceinline__ uint8_t pr_label(const int64_t* __restrict__ labels, const float* __restrict__ conf,
                                            float conf_thr, int b, int p, int w, int lab_h, int lab_w,
                                            float sh, float sw, int C) {
  const int y = p / w, x = p - y * w;
  const int64_t o = ((int64_t)b * lab_h + pr_nearest(y, sh, lab_h)) * lab_w + pr_nearest(x, sw, lab_w);
  const int64_t l = labels[o];
  bool ok = l >= 0 && l < C;
  if (conf && ok) ok = conf[o] >= conf_thr;
  return ok ? (uint8_t)l : (uint8_t)255;
}

// grid = (channel groups, B). Every warp takes channels of its group round-robin;
// lanes stride over the image plane with coalesced loads.
__global__ void __launch_bounds__(kPrThreads)
proto_accum_kernel(const float* __restrict__ feats, int B, int D, int h, int w,
                   const int64_t* __restrict__ labels, const float* __restrict__ conf, float conf_thr,
                   int lab_h, int lab_w, int C, int ch_per_block, float* __restrict__ packed) {
  extern __shared__ __align__(16) unsigned char pr_smem[];
  uint8_t* lab_s = pr_smem;                                              // [kPrPixTile]
  float* acc = reinterpret_cast<float*>(pr_smem + kPrPixTile);           // [warps][C][32]
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * ch_per_block;
  const int c1 = min(D, c0 + ch_per_block);
  const int hw = h * w;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sh = (float)lab_h / (float)h, sw = (float)lab_w / (float)w;
  float* my = acc + warp * C * 32;
  for (int i = lane; i < C * 32; i += 32) my[i] = 0.f;

  for (int p0 = 0; p0 < hw; p0 += kPrPixTile) {
    const int np = min(kPrPixTile, hw - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += kPrThreads)
      lab_s[i] = pr_label(labels, conf, conf_thr, b, p0 + i, w, lab_h, lab_w, sh, sw, C);
    __syncthreads();
    // class counts: once per image (channel group 0), same private-accumulator scheme
    if (blockIdx.x == 0 && warp == 0) {
      for (int i = lane; i < np; i += 32) {
        const unsigned l = lab_s[i];
        if (l != 255u) my[l * 32 + lane] += 1.f;
      }
      __syncwarp();
      for (int c = 0; c < C; ++c) {
        const float v = warp_sum(my[c * 32 + lane]);
        my[c * 32 + lane] = 0.f;
        if (lane == 0 && v != 0.f) atomicAdd(&packed[(int64_t)C * D + c], v);
      }
      __syncwarp();
    }
    for (int ch = c0 + warp; ch < c1; ch += kPrWarps) {
      const float* src = feats + ((int64_t)b * D + ch) * hw + p0;
      const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && (np % 4 == 0);
      if (vec) {
        for (int i = lane * 4; i < np; i += 128) {
          const float4 v = ldg_stream_f4(src + i);
          const uchar4 l = *reinterpret_cast<const uchar4*>(lab_s + i);
          if (l.x != 255) my[l.x * 32 + lane] += v.x;
          if (l.y != 255) my[l.y * 32 + lane] += v.y;
          if (l.z != 255) my[l.z * 32 + lane] += v.z;
          if (l.w != 255) my[l.w * 32 + lane] += v.w;
        }
      } else {
        for (int i = lane; i < np; i += 32) {
          const unsigned l = lab_s[i];
          if (l != 255u) my[l * 32 + lane] += __ldg(src + i);
        }
      }
      __syncwarp();
      for (int c = 0; c < C; ++c) {
        const float v = warp_sum(my[c * 32 + lane]);
        my[c * 32 + lane] = 0.f;
        if (lane == 0 && v != 0.f) atomicAdd(&packed[(int64_t)c * D + ch], v);
      }
      __syncwarp();
    }
  }
}

__global__ void proto_finalize_kernel(const float* __restrict__ packed, int C, int D,
                                      const float* __restrict__ mu_prev, const uint8_t* __restrict__ seen_prev,
                                      float a32, float b32, float* __restrict__ mu_out,
                                      int64_t* __restrict__ cnt_out, uint8_t* __restrict__ seen_out) {
  const int c = blockIdx.x;
  const float cnt = packed[(int64_t)C * D + c];
  const bool has = cnt > 0.f;
  const bool seen = seen_prev ? seen_prev[c] != 0 : false;
  const float denom = fmaxf(cnt, 1.f);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float mean = packed[(int64_t)c * D + d] / denom;
    const float prev = mu_prev ? mu_prev[(int64_t)c * D + d] : 0.f;
    float out = prev;
    if (has) out = seen ? __fadd_rn(__fmul_rn(a32, prev), __fmul_rn(b32, mean)) : mean;
    mu_out[(int64_t)c * D + d] = out;
  }
  if (threadIdx.x == 0) {
    if (cnt_out) cnt_out[c] = (int64_t)cnt;
    if (seen_out) seen_out[c] = (has || seen) ? 1 : 0;
  }
}

// ---- distance: block = 128 consecutive pixels of one image x all channels -------
constexpr int kPdPix = 128;

struct PdCtx {
  int b, p0, hw;
  uint8_t lab[4];
};

// BWD = false: dist[n] = ||f_n - mu_y||, block sums -> acc[0] (sum dist), acc[1] (n valid)
// BWD = true : grad[n,d] = g * (f - mu_y) / (dist * n_valid)
template <bool BWD>
__global__ void __launch_bounds__(kPrThreads)
proto_dist_kernel(const float* __restrict__ feats, int B, int D, int h, int w,
                  const int64_t* __restrict__ labels, int lab_h, int lab_w, const float* __restrict__ mu,
                  const uint8_t* __restrict__ seen, int C, float* __restrict__ dist, double* __restrict__ acc,
                  float* __restrict__ loss, const float* __restrict__ grad_loss, float* __restrict__ grad,
                  unsigned* __restrict__ done_counter, int accumulate) {
  extern __shared__ __align__(16) unsigned char pd_smem[];
  float* mu_s = reinterpret_cast<float*>(pd_smem);                 // [C][D+1]  (+1: bank skew)
  float* part = mu_s + (size_t)C * (D + 1);                        // [warps][128]
  const int hw = h * w;
  const int tiles = (hw + kPdPix - 1) / kPdPix;
  const int b = blockIdx.x / tiles, p0 = (blockIdx.x - b * tiles) * kPdPix;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < C * D; i += kPrThreads) mu_s[(i / D) * (D + 1) + (i % D)] = mu[i];
  const float sh = (float)lab_h / (float)h, sw = (float)lab_w / (float)w;
  // this lane's four pixels
  int lab[4];
  bool ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + lane * 4 + i;
    uint8_t l = 255;
    if (p < hw) l = pr_label(labels, nullptr, 0.f, b, p, w, lab_h, lab_w, sh, sw, C);
    if (l != 255 && seen && !seen[l]) l = 255;
    ok[i] = l != 255;
    lab[i] = ok[i] ? l : 0;
  }
  __syncthreads();
  const float* src = feats + (int64_t)b * D * hw + p0 + lane * 4;
  const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(feats) & 15u) == 0);
  const bool any_px = p0 + lane * 4 < hw;

  if (!BWD) {
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    if (any_px)
      for (int d = warp; d < D; d += kPrWarps) {
        float v[4];
        if (vec) {
          const float4 t = ldg_stream_f4(src + (int64_t)d * hw);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = (p0 + lane * 4 + i < hw) ? __ldg(src + (int64_t)d * hw + i) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float df = v[i] - mu_s[lab[i] * (D + 1) + d];
          ss[i] = fmaf(df, df, ss[i]);
        }
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) part[warp * kPdPix + lane * 4 + i] = ss[i];
    __syncthreads();
    double bsum = 0.0, bcnt = 0.0;
    if (warp == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = 0.f;
        for (int wv = 0; wv < kPrWarps; ++wv) s += part[wv * kPdPix + lane * 4 + i];
        const float dn = ok[i] ? sqrtf(s) : 0.f;
        const int p = p0 + lane * 4 + i;
        if (p < hw) dist[(int64_t)b * hw + p] = dn;
        if (ok[i]) { bsum += (double)dn; bcnt += 1.0; }
      }
      bsum = warp_sum(bsum);
      bcnt = warp_sum(bcnt);
      if (lane == 0) {
        if (bcnt != 0.0) { atomicAdd(&acc[0], bsum); atomicAdd(&acc[1], bcnt); }
        __threadfence();
        if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {
          __threadfence();
          const double s = *((volatile double*)&acc[0]), n = *((volatile double*)&acc[1]);
          loss[0] = (float)(s / n);   // mean of an empty selection is NaN, as torch.mean
        }
      }
    }
  } else {
    const float g = grad_loss[0];
    const float nvalid = (float)acc[1];
    float coef[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + lane * 4 + i;
      const float dn = (p < hw) ? dist[(int64_t)b * hw + p] : 0.f;
      coef[i] = (ok[i] && dn > 0.f) ? g / (dn * nvalid) : 0.f;   // torch.norm backward: 0 at 0
    }
    float* dst = grad + (int64_t)b * D * hw + p0 + lane * 4;
    if (any_px)
      for (int d = warp; d < D; d += kPrWarps) {
        float v[4];
        if (vec) {
          const float4 t = ldg_stream_f4(src + (int64_t)d * hw);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = (p0 + lane * 4 + i < hw) ? __ldg(src + (int64_t)d * hw + i) : 0.f;
        }
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (accumulate) {   // grad += ... : the PFGST loss gradient is already in the buffer
          if (vec) {
            const float4 t = *reinterpret_cast<const float4*>(dst + (int64_t)d * hw);
            o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (p0 + lane * 4 + i < hw) o[i] = dst[(int64_t)d * hw + i];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fmaf(coef[i], v[i] - mu_s[lab[i] * (D + 1) + d], o[i]);
        if (vec) {
          __stcs(reinterpret_cast<float4*>(dst + (int64_t)d * hw), make_float4(o[0], o[1], o[2], o[3]));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (p0 + lane * 4 + i < hw) dst[(int64_t)d * hw + i] = o[i];
        }
      }
  }
}

// all-class distances: out[b,c,n] = ||f_n - mu_c||_2
__global__ void __launch_bounds__(kPrThreads)
proto_dist_all_kernel(const float* __restrict__ feats, int B, int D, int h, int w,
                      const float* __restrict__ mu, int C, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char pa_smem[];
  float* mu_s = reinterpret_cast<float*>(pa_smem);     // [C][D]
  const int hw = h * w;
  for (int i = threadIdx.x; i < C * D; i += kPrThreads) mu_s[i] = mu[i];
  __syncthreads();
  const int64_t total = (int64_t)B * hw;
  for (int64_t n = (int64_t)blockIdx.x * kPrThreads + threadIdx.x; n < total;
       n += (int64_t)gridDim.x * kPrThreads) {
    const int b = (int)(n / hw);
    const int p = (int)(n - (int64_t)b * hw);
    const float* src = feats + (int64_t)b * D * hw + p;
    for (int c0 = 0; c0 < C; c0 += 8) {
      float ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int d = 0; d < D; ++d) {
        const float v = __ldg(src + (int64_t)d * hw);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < C) {
            const float df = v - mu_s[(c0 + j) * D + d];
            ss[j] = fmaf(df, df, ss[j]);
          }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < C) out[((int64_t)b * C + c0 + j) * hw + p] = sqrtf(ss[j]);
    }
  }
}

}  // namespace pfst

extern "C" {

int pfst_proto_accum(const float* feats, int64_t B, int32_t D, int32_t h, int32_t w,
                     const int64_t* labels, int32_t lab_h, int32_t lab_w, const float* conf,
                     float conf_thr, int32_t C, float* packed, void* stream) {
  if (!feats || !labels || !packed || B < 0 || D < 1 || h < 1 || w < 1 || lab_h < 1 || lab_w < 1 || C < 1)
    return PFST_ERR_INVALID_ARG;
  if (C > pfst::kPrMaxC || B > 65535) return PFST_ERR_UNSUPPORTED;
  if (B == 0) return PFST_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = pfst::kPrPixTile + (size_t)pfst::kPrWarps * C * 32 * sizeof(float);
  PFST_CUDA_TRY(cudaFuncSetAttribute(pfst::proto_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem), "pfst_proto_accum/attr");
  // channel groups sized so that ~2 blocks per SM are in flight
  int groups = (int)(((int64_t)pfst::kNumSMs * 2 + B - 1) / B);
  int cpb = (D + groups - 1) / groups;
  cpb = ((cpb + pfst::kPrWarps - 1) / pfst::kPrWarps) * pfst::kPrWarps;
  groups = (D + cpb - 1) / cpb;
  pfst::proto_accum_kernel<<<dim3((unsigned)groups, (unsigned)B), pfst::kPrThreads, smem, s>>>(
      feats, (int)B, D, h, w, labels, conf, conf_thr, lab_h, lab_w, C, cpb, packed);
  PFST_CHECK_LAUNCH("pfst_proto_accum");
  return PFST_OK;
}

int pfst_proto_finalize(const float* packed, int32_t C, int32_t D, const float* mu_prev,
                        const uint8_t* seen_prev, float a32, float b32, float* mu_out, int64_t* cnt_out,
                        uint8_t* seen_out, void* stream) {
  if (!packed || !mu_out || C < 1 || D < 1) return PFST_ERR_INVALID_ARG;
  pfst::proto_finalize_kernel<<<(unsigned)C, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      packed, C, D, mu_prev, seen_prev, a32, b32, mu_out, cnt_out, seen_out);
  PFST_CHECK_LAUNCH("pfst_proto_finalize");
  return PFST_OK;
}

static int proto_dist_common(bool bwd, const float* feats, int64_t B, int32_t D, int32_t h, int32_t w,
                             const int64_t* labels, int32_t lab_h, int32_t lab_w, const float* mu,
                             const uint8_t* seen, int32_t C, float* dist, double* acc, float* loss,
                             const float* grad_loss, float* grad, int accumulate, cudaStream_t s) {
  if (!feats || !labels || !mu || !dist || !acc || B < 0 || D < 1 || h < 1 || w < 1 || C < 1)
    return PFST_ERR_INVALID_ARG;
  if (C > pfst::kPrMaxC) return PFST_ERR_UNSUPPORTED;
  const size_t smem = ((size_t)C * (D + 1) + (size_t)pfst::kPrWarps * pfst::kPdPix) * sizeof(float);
  if (smem > 200 * 1024) return PFST_ERR_UNSUPPORTED;
  const int hw = h * w;
  const int64_t grid = B * ((hw + pfst::kPdPix - 1) / pfst::kPdPix);
  if (grid == 0) return PFST_OK;
  if (!bwd) {
    if (!loss) return PFST_ERR_INVALID_ARG;
    PFST_CUDA_TRY(cudaMemsetAsync(acc, 0, 4 * sizeof(double), s), "pfst_proto_dist_fwd/memset");
    auto k = pfst::proto_dist_kernel<false>;
    PFST_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pfst_proto_dist_fwd/attr");
    k<<<(unsigned)grid, pfst::kPrThreads, smem, s>>>(feats, (int)B, D, h, w, labels, lab_h, lab_w, mu, seen, C, dist,
                                                     acc, loss, nullptr, nullptr,
                                                     reinterpret_cast<unsigned*>(acc + 3), 0);
  } else {
    if (!grad_loss || !grad) return PFST_ERR_INVALID_ARG;
    auto k = pfst::proto_dist_kernel<true>;
    PFST_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pfst_proto_dist_bwd/attr");
    k<<<(unsigned)grid, pfst::kPrThreads, smem, s>>>(feats, (int)B, D, h, w, labels, lab_h, lab_w, mu, seen, C, dist,
                                                     acc, nullptr, grad_loss, grad, nullptr, accumulate);
  }
  PFST_CHECK_LAUNCH(bwd ? "pfst_proto_dist_bwd" : "pfst_proto_dist_fwd");
  return PFST_OK;
}

int pfst_proto_dist_fwd(const float* feats, int64_t B, int32_t D, int32_t h, int32_t w,
                        const int64_t* labels, int32_t lab_h, int32_t lab_w, const float* mu,
                        const uint8_t* seen, int32_t C, float* dist, double* acc, float* loss, void* stream) {
  return proto_dist_common(false, feats, B, D, h, w, labels, lab_h, lab_w, mu, seen, C, dist, acc, loss, nullptr,
                           nullptr, 0, static_cast<cudaStream_t>(stream));
}

int pfst_proto_dist_bwd(const float* feats, int64_t B, int32_t D, int32_t h, int32_t w,
                        const int64_t* labels, int32_t lab_h, int32_t lab_w, const float* mu,
                        const uint8_t* seen, int32_t C, const float* dist, const double* acc,
                        const float* grad_loss, float* grad_feats, int32_t accumulate, void* stream) {
  return proto_dist_common(true, feats, B, D, h, w, labels, lab_h, lab_w, mu, seen, C, const_cast<float*>(dist),
                           const_cast<double*>(acc), nullptr, grad_loss, grad_feats, accumulate,
                           static_cast<cudaStream_t>(stream));
}

int pfst_proto_dist_all(const float* feats, int64_t B, int32_t D, int32_t h, int32_t w, const float* mu,
                        int32_t C, float* out, void* stream) {
  if (!feats || !mu || !out || B < 0 || D < 1 || h < 1 || w < 1 || C < 1) return PFST_ERR_INVALID_ARG;
  const size_t smem = (size_t)C * D * sizeof(float);
  if (smem > 200 * 1024) return PFST_ERR_UNSUPPORTED;
  const int64_t total = B * h * w;
  if (total == 0) return PFST_OK;
  PFST_CUDA_TRY(cudaFuncSetAttribute(pfst::proto_dist_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem), "pfst_proto_dist_all/attr");
  int64_t grid = (total + pfst::kPrThreads - 1) / pfst::kPrThreads;
  const int64_t cap = (int64_t)pfst::kNumSMs * 4;
  if (grid > cap) grid = cap;
  pfst::proto_dist_all_kernel<<<(unsigned)grid, pfst::kPrThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      feats, (int)B, D, h, w, mu, C, out);
  PFST_CHECK_LAUNCH("pfst_proto_dist_all");
  return PFST_OK;
}

}  // extern "C"
