// G1 — PFGST.masked_feat_dist (rsiseg/models/uda/pfgst.py:168-177):
//   pw = torch.norm(f1 - f2, dim=1, p=2); if mask: pw = pw[mask.squeeze(1)]; return torch.mean(pw)
// One pass over the two (B,D,h,w) feature maps (8·D B per pixel) instead of a (B,D,h,w) difference
// tensor, a norm kernel, a boolean gather (host sync for its size) and a mean; the backward is one
// elementwise pass: d/df1 = g (f1 - f2) / (||f1 - f2|| n), 0 where the norm is 0 (torch.norm's
// subgradient) or the pixel is masked out; d/df2 = -d/df1.
#include <math.h>

#include "common.cuh"

namespace pfst {

constexpr int kFdPix = 64, kFdGroups = 4;

// acc: double[4] = {sum of selected distances, number of selected pixels, finished blocks, -}
__global__ void __launch_bounds__(kFdPix * kFdGroups)
feat_dist_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const uint8_t* __restrict__ mask,
                     int64_t B, int D, int64_t hw, float* __restrict__ dist, double* __restrict__ acc,
                     float* __restrict__ loss) {
  __shared__ float part[kFdGroups][kFdPix];
  __shared__ double bsum[2];
  const int p = threadIdx.x % kFdPix, g = threadIdx.x / kFdPix;
  const int64_t n = (int64_t)blockIdx.x * kFdPix + p;
  const bool live = n < B * hw;
  float tot = 0.f;
  if (live) {
    const int64_t b = n / hw, r = n - b * hw;
    const float* a = f1 + b * D * hw + r;
    const float* c = f2 + b * D * hw + r;
    float run = 0.f;
    int in_chunk = 0;
    for (int d = g; d < D; d += kFdGroups) {
      const float v = a[(int64_t)d * hw] - c[(int64_t)d * hw];
      run = fmaf(v, v, run);
      if (++in_chunk == 16) { tot += run; run = 0.f; in_chunk = 0; }
    }
    tot += run;
  }
  part[g][p] = tot;
  if (threadIdx.x < 2) bsum[threadIdx.x] = 0.0;
  __syncthreads();
  if (g == 0) {
    float dv = 0.f;
    bool sel = false;
    if (live) {
      dv = sqrtf(((part[0][p] + part[1][p]) + part[2][p]) + part[3][p]);
      sel = mask ? mask[n] != 0 : true;
      dist[n] = sel ? dv : 0.f;
    }
    double s = warp_sum(sel ? (double)dv : 0.0);
    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, sel));
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&bsum[0], s);
      atomicAdd(&bsum[1], (double)cnt);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (bsum[1] != 0.0) {
      atomicAdd(&acc[0], bsum[0]);
      atomicAdd(&acc[1], bsum[1]);
    }
    __threadfence();
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(&acc[2]), 1ull);
    if (done == gridDim.x - 1) {
      __threadfence();
      const double s = *reinterpret_cast<volatile double*>(&acc[0]);
      const double c = *reinterpret_cast<volatile double*>(&acc[1]);
      loss[0] = (float)(s / c);                       // torch.mean of an empty selection is NaN
    }
  }
}

__global__ void feat_dist_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                     const uint8_t* __restrict__ mask, int64_t B, int D, int64_t hw,
                                     const float* __restrict__ dist, const double* __restrict__ acc,
                                     const float* __restrict__ gout, float* __restrict__ g1, float* __restrict__ g2) {
  const int64_t total = B * (int64_t)D * hw;
  const float scale = gout[0] / (float)acc[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ((int64_t)D * hw);
    const int64_t r = i % hw;
    const int64_t n = b * hw + r;
    const float dn = dist[n];
    const bool sel = (mask ? mask[n] != 0 : true) && dn > 0.f;
    const float v = sel ? scale * (f1[i] - f2[i]) / dn : 0.f;
    if (g1) g1[i] = v;
    if (g2) g2[i] = -v;
  }
}

}  // namespace pfst

extern "C" {

int pfst_feat_dist_fwd(const float* f1, const float* f2, const uint8_t* mask, int64_t B, int32_t D, int32_t h,
                       int32_t w, float* dist, double* acc, float* loss, void* stream) {
  if (!f1 || !f2 || !dist || !acc || !loss || B < 0 || D < 1 || h < 1 || w < 1) return PFST_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PFST_CUDA_TRY(cudaMemsetAsync(acc, 0, 4 * sizeof(double), s), "pfst_feat_dist_fwd/memset");
  const int64_t hw = (int64_t)h * w, px = B * hw;
  if (px == 0) return PFST_OK;
  const int64_t grid = (px + pfst::kFdPix - 1) / pfst::kFdPix;
  if (grid > 0x7fffffffll) return PFST_ERR_UNSUPPORTED;
  pfst::feat_dist_fwd_kernel<<<(unsigned)grid, pfst::kFdPix * pfst::kFdGroups, 0, s>>>(f1, f2, mask, B, D, hw, dist,
                                                                                        acc, loss);
  PFST_CHECK_LAUNCH("pfst_feat_dist_fwd");
  return PFST_OK;
}

int pfst_feat_dist_bwd(const float* f1, const float* f2, const uint8_t* mask, int64_t B, int32_t D, int32_t h,
                       int32_t w, const float* dist, const double* acc, const float* grad_loss, float* grad_f1,
                       float* grad_f2, void* stream) {
  if (!f1 || !f2 || !dist || !acc || !grad_loss || (!grad_f1 && !grad_f2) || B < 0 || D < 1 || h < 1 || w < 1)
    return PFST_ERR_INVALID_ARG;
  const int64_t total = B * (int64_t)D * h * w;
  if (total == 0) return PFST_OK;
  const int64_t blocks = (total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  pfst::feat_dist_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      f1, f2, mask, B, D, (int64_t)h * w, dist, acc, grad_loss, grad_f1, grad_f2);
  PFST_CHECK_LAUNCH("pfst_feat_dist_bwd");
  return PFST_OK;
}

}  // extern "C"
