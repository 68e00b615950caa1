// The rest of the offline class-wise pseudo-labelling path:
//   loc_dis_kernel            PseudoLabelingHookV4._cal_loc_dis  (pseudo_labeling_hookv4.py:208-230)
//   gather_rows_kernel +
//   sigma_bisect_kernel       PseudoLabelingHookV4._cal_sigmas   (pseudo_labeling_hookv4.py:232-277)
//   loader_labels_kernel      LoadAnnotationsPseudoLabelsV2.__call__, the label rule (loading.py:474-487)
// The reference runs these on the CPU with nn.Unfold (a 9x copy of every feature map), ~30 full
// passes of exp over the sampled distances per (level, dilation, mean_sim), and numpy per image in
// the data loader. Here: one pass over the features, one launch per bisection step with the
// interval kept on the device (no host sync inside the search), one pass over the logits.
#include <math.h>

#include "common.cuh"

namespace pfst {

// ---- squared distances to the 3x3 dilated neighbours (zero padding) --------------------------
// block = 64 pixels x 4 channel groups; a thread walks the channels c = g, g+4, ... of its pixel
// (loads coalesced along x, neighbours served by L1/L2), partial sums in chunks of 16 channels
// (cascade summation like ATen's sum), groups merged in fixed order through shared memory.
constexpr int kLdPix = 64, kLdGroups = 4;

__global__ void __launch_bounds__(kLdPix * kLdGroups)
loc_dis_kernel(const float* __restrict__ feat, int64_t B, int C, int H, int W, int dil, float* __restrict__ out) {
  __shared__ float part[kLdGroups][9][kLdPix];
  const int p = threadIdx.x % kLdPix, g = threadIdx.x / kLdPix;
  const int64_t hw = (int64_t)H * W;
  const int64_t n = (int64_t)blockIdx.x * kLdPix + p;          // pixel over (B,H,W)
  const bool live = n < B * hw;
  float tot[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) tot[k] = 0.f;
  if (live) {
    const int64_t b = n / hw;
    const int r = (int)(n - b * hw);
    const int y = r / W, x = r - y * W;
    int off[9];
    bool in[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int yy = y + (k / 3 - 1) * dil, xx = x + (k % 3 - 1) * dil;
      in[k] = yy >= 0 && yy < H && xx >= 0 && xx < W;
      off[k] = in[k] ? yy * W + xx : r;
    }
    const float* base = feat + b * C * hw;
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    int in_chunk = 0;
    for (int c = g; c < C; c += kLdGroups) {
      const float* pl = base + (int64_t)c * hw;
      const float v = pl[r];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float d = (in[k] ? pl[off[k]] : 0.f) - v;
        acc[k] = fmaf(d, d, acc[k]);
      }
      if (++in_chunk == 16) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { tot[k] += acc[k]; acc[k] = 0.f; }
        in_chunk = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) tot[k] += acc[k];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) part[g][k][p] = tot[k];
  __syncthreads();
  // 64 pixels x 9 taps = 576 outputs, contiguous in (B,H,W,9)
  for (int i = threadIdx.x; i < kLdPix * 9; i += kLdPix * kLdGroups) {
    const int pp = i / 9, k = i - pp * 9;
    const int64_t nn = (int64_t)blockIdx.x * kLdPix + pp;
    if (nn < B * hw) {
      float s = part[0][k][pp];
#pragma unroll
      for (int gg = 1; gg < kLdGroups; ++gg) s += part[gg][k][pp];
      out[nn * 9 + k] = s;
    }
  }
}

// ---- rows idx[i] of a (N, row) fp32 matrix -> compact (n, row) ---------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, int row,
                                   float* __restrict__ dst) {
  const int64_t total = n * row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row;
    dst[i] = src[idx[r] * row + (i - r * row)];
  }
}

// ---- one bisection step: sigma = (left + right) / 2; mean(exp(-d / sigma^2)) < target ? left = sigma
//      : right = sigma. state = double[4] {left, right, running sum, finished blocks}. ------------
constexpr int kSbThreads = 256;

__global__ void sigma_init_kernel(double* state, double left, double right) {
  state[0] = left; state[1] = right; state[2] = 0.0;
  *reinterpret_cast<unsigned long long*>(&state[3]) = 0ull;
}

__global__ void __launch_bounds__(kSbThreads)
sigma_bisect_kernel(const float* __restrict__ dis, int64_t n, float target, double* __restrict__ state) {
  const double left = *reinterpret_cast<volatile double*>(&state[0]);
  const double right = *reinterpret_cast<volatile double*>(&state[1]);
  const double sigma = (left + right) / 2;
  const float s2 = (float)(sigma * sigma);          // `dis / sigma ** 2`: python double -> fp32 scalar operand
  float acc = 0.f;
  double big = 0.0;
  int cnt = 0;
  const int64_t n4 = n / 4;
  const float4* d4 = reinterpret_cast<const float4*>(dis);
  for (int64_t i = (int64_t)blockIdx.x * kSbThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kSbThreads) {
    const float4 v = d4[i];
    acc += expf(-v.x / s2) + expf(-v.y / s2) + expf(-v.z / s2) + expf(-v.w / s2);
    if (++cnt == 64) { big += (double)acc; acc = 0.f; cnt = 0; }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) acc += expf(-dis[n4 * 4 + threadIdx.x] / s2);
  big += (double)acc;
  big = warp_sum(big);
  __shared__ double wsum[kSbThreads / 32];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = big;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kSbThreads / 32; ++w) t += wsum[w];
    atomicAdd(&state[2], t);
    __threadfence();
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(&state[3]), 1ull);
    last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const double total = *reinterpret_cast<volatile double*>(&state[2]);
    const float mean = (float)(total / (double)n);   // the reference compares an fp32 mean with fp32(mean_sim)
    if (mean < target) state[0] = sigma; else state[1] = sigma;
    state[2] = 0.0;
    *reinterpret_cast<unsigned long long*>(&state[3]) = 0ull;
  }
}

// ---- the loader's label rule, per pixel -------------------------------------------------------
__global__ void loader_labels_kernel(const float* __restrict__ logits, int64_t N, int C, int64_t HW,
                                     const float* __restrict__ thres, int reduce_zero, uint8_t* __restrict__ out) {
  const int64_t total = N * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW;
    const float* z = logits + b * C * HW + (i - b * HW);
    // numpy argmax: first maximum; a NaN counts as the maximum (first NaN wins)
    int best = 0;
    float bv = z[0];
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = z[(int64_t)c * HW];
      if (c > 0 && !(bv != bv) && (v > bv || v != v)) { bv = v; best = c; }
      s = __fadd_rn(s, expf(v));                       // np.exp(logits).sum(axis=0): class order, no max shift
    }
    float ent = 0.f;
    for (int c = 0; c < C; ++c) {
      const float pr = __fdiv_rn(expf(z[(int64_t)c * HW]), s);
      ent = __fadd_rn(ent, __fmul_rn(pr, logf(__fadd_rn(pr, 1e-8f))));
    }
    ent = -ent;
    int lab = (ent < thres[best]) ? best : 255;
    if (reduce_zero) {                                 // loading.py:482-486
      if (lab == 0) lab = 255;
      lab = lab - 1;
      if (lab == 254) lab = 255;
    }
    out[i] = (uint8_t)lab;
  }
}

}  // namespace pfst

extern "C" {

int pfst_loc_dis(const float* feats, int64_t B, int32_t C, int32_t H, int32_t W, int32_t dilation, float* out,
                 void* stream) {
  if (!feats || !out || B < 0 || C < 1 || H < 1 || W < 1 || dilation < 1) return PFST_ERR_INVALID_ARG;
  const int64_t px = B * H * W;
  if (px == 0) return PFST_OK;
  const int64_t grid = (px + pfst::kLdPix - 1) / pfst::kLdPix;
  if (grid > 0x7fffffffll || (int64_t)H * W > 0x7fffffffll) return PFST_ERR_UNSUPPORTED;
  pfst::loc_dis_kernel<<<(unsigned)grid, pfst::kLdPix * pfst::kLdGroups, 0, static_cast<cudaStream_t>(stream)>>>(
      feats, B, C, H, W, dilation, out);
  PFST_CHECK_LAUNCH("pfst_loc_dis");
  return PFST_OK;
}

int pfst_gather_rows(const float* src, const int64_t* idx, int64_t n, int32_t row_floats, float* dst, void* stream) {
  if (!src || !idx || !dst || n < 0 || row_floats < 1) return PFST_ERR_INVALID_ARG;
  if (n == 0) return PFST_OK;
  const int64_t blocks = (n * row_floats + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  pfst::gather_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, idx, n, row_floats, dst);
  PFST_CHECK_LAUNCH("pfst_gather_rows");
  return PFST_OK;
}

int pfst_sigma_bisect(const float* dis, int64_t n, float mean_sim, double left0, double right0, int32_t steps,
                      double* state, void* stream) {
  if (!dis || !state || n < 1 || steps < 0 || steps > 200) return PFST_ERR_INVALID_ARG;
  if (!pfst::aligned16(dis)) return PFST_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pfst::sigma_init_kernel<<<1, 1, 0, s>>>(state, left0, right0);
  const int64_t blocks = (n / 4 + pfst::kSbThreads - 1) / pfst::kSbThreads;
  const unsigned grid = (unsigned)(blocks < 1 ? 1 : (blocks < 148 * 8 ? blocks : 148 * 8));
  for (int i = 0; i < steps; ++i) {
    pfst::sigma_bisect_kernel<<<grid, pfst::kSbThreads, 0, s>>>(dis, n, mean_sim, state);
    PFST_CHECK_LAUNCH("pfst_sigma_bisect");
  }
  return PFST_OK;
}

int pfst_loader_pseudo_labels(const float* logits, int64_t N, int32_t C, int64_t HW, const float* thres,
                              int32_t reduce_zero_label, uint8_t* labels, void* stream) {
  if (!logits || !thres || !labels || N < 0 || C < 1 || C > 255 || HW < 1) return PFST_ERR_INVALID_ARG;
  if (N == 0) return PFST_OK;
  const int64_t blocks = (N * HW + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  pfst::loader_labels_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, N, C, HW, thres,
                                                                                  reduce_zero_label, labels);
  PFST_CHECK_LAUNCH("pfst_loader_pseudo_labels");
  return PFST_OK;
}

}  // extern "C"
