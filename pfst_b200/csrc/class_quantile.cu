// §8(f) rank 4 — offline class-wise pseudo-label thresholds.
//
// Reference: PseudoLabelingHookV4._cal_threshold, rsiseg/core/hook/pseudo_labeling_hookv4.py:173-205
//   logits (B,C,H,W) -> rows (B*H*W, C) -> a random subset idx (host numpy permutation)
//   prob = softmax(row); pred = argmax(prob); ent = sum_c -prob_c * log(prob_c)
//   for every class c and ratio r:  thr[r][c] = sort(ent[pred == c])[int(n_c * r)]   (0 if n_c == 0)
// The reference sorts every class's entropies on the host (numpy). Here:
//   entropy_argmax_kernel   one thread per sampled pixel: pred (u8) + entropy (f32), class counts;
//   quantile_*_kernel       exact k-th order statistic per (class, ratio) by a three-pass radix
//                           select on the IEEE bits of the (non-negative) entropies — 11 + 11 + 10
//                           bit digits, shared-memory privatised histograms — no sort at all.
// Memory-bound: 4*C B per sampled pixel once, then 3 passes over 5 B per sampled pixel.
#include <math.h>

#include "common.cuh"

namespace pfst {

constexpr int kCqThreads = 256;
constexpr int kCqMaxSel = 512;      // (class, ratio) selections
constexpr int kCqBins = 2048;

__global__ void __launch_bounds__(kCqThreads)
entropy_argmax_kernel(const float* __restrict__ logits, int C, int64_t HW, const int64_t* __restrict__ idx,
                      int64_t n, uint8_t* __restrict__ pred, float* __restrict__ ent,
                      unsigned long long* __restrict__ counts) {
  __shared__ unsigned cnt_s[256];
  for (int i = threadIdx.x; i < 256; i += kCqThreads) cnt_s[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kCqThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCqThreads) {
    const int64_t row = idx ? idx[i] : i;               // row of the (B*H*W, C) view = pixel-major index
    const int64_t b = row / HW, p = row - b * HW;
    const float* src = logits + (b * C) * HW + p;
    float m = src[0];
    int am = 0;
    for (int c = 1; c < C; ++c) {
      const float v = src[(int64_t)c * HW];
      if (v > m) { m = v; am = c; }                     // first maximum wins (torch.argmax)
    }
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(src[(int64_t)c * HW] - m);
    float e = 0.f;
    for (int c = 0; c < C; ++c) {
      const float pc = expf(src[(int64_t)c * HW] - m) / s;
      e += -pc * logf(pc);                              // 0 * -inf = NaN, as in the reference
    }
    pred[i] = (uint8_t)am;
    ent[i] = e;
    atomicAdd(&cnt_s[am], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += kCqThreads)
    if (cnt_s[i]) atomicAdd(&counts[i], (unsigned long long)cnt_s[i]);
}

// sortable key of an entropy: non-negative floats order like their bit patterns; -0 -> +0;
// NaN (and anything negative, which cannot occur) sorts last like numpy's sort
__device__ __forceinline__ unsigned cq_key(float e) {
  if (e != e) return 0xffffffffu;
  const unsigned u = __float_as_uint(e);
  return (u & 0x80000000u) ? 0u : u;
}

struct CqState {       // per selection (class * R + ratio)
  unsigned prefix;     // key bits fixed so far (high bits)
  unsigned long long rank;   // residual rank inside the prefix bucket
  int active;
};

// state init: rank = int(n_c * ratio) (python: int * float -> float64 -> trunc)
__global__ void quantile_init_kernel(const unsigned long long* __restrict__ counts, const double* __restrict__ ratios,
                                     int C, int R, CqState* __restrict__ st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * R) return;
  const unsigned long long nc = counts[i / R];
  CqState s;
  s.prefix = 0;
  s.active = nc > 0;
  unsigned long long k = (unsigned long long)((double)nc * ratios[i % R]);
  if (nc > 0 && k >= nc) k = nc - 1;                   // ratio 1.0 would index past the end in the reference
  s.rank = k;
  st[i] = s;
}

// one pass: histogram of digit `shift..shift+bits` of the keys whose higher bits equal the prefix
__global__ void __launch_bounds__(kCqThreads)
quantile_hist_kernel(const uint8_t* __restrict__ pred, const float* __restrict__ ent, int64_t n, int R,
                     const CqState* __restrict__ st, int shift, int bits, unsigned hi_mask,
                     unsigned* __restrict__ hist) {
  for (int64_t i = (int64_t)blockIdx.x * kCqThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCqThreads) {
    const int c = pred[i];
    const unsigned key = cq_key(ent[i]);
    const unsigned digit = (key >> shift) & ((1u << bits) - 1u);
    for (int r = 0; r < R; ++r) {
      const CqState s = st[c * R + r];
      if (s.active && (key & hi_mask) == s.prefix) atomicAdd(&hist[(size_t)(c * R + r) * kCqBins + digit], 1u);
    }
  }
}

// one block per selection: find the digit whose cumulative count passes the residual rank
__global__ void __launch_bounds__(kCqThreads)
quantile_scan_kernel(CqState* __restrict__ st, unsigned* __restrict__ hist, int shift, int bits,
                     float* __restrict__ out, int last) {
  __shared__ unsigned long long part[kCqThreads];
  const int sel = blockIdx.x;
  CqState s = st[sel];
  unsigned* h = hist + (size_t)sel * kCqBins;
  const int nb = 1 << bits, per = (nb + kCqThreads - 1) / kCqThreads;
  unsigned long long local = 0;
  for (int j = 0; j < per; ++j) {
    const int bin = threadIdx.x * per + j;
    if (bin < nb) local += h[bin];
  }
  part[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0 && s.active) {
    unsigned long long cum = 0;
    int t = 0;
    for (; t < kCqThreads; ++t) {
      if (cum + part[t] > s.rank) break;
      cum += part[t];
    }
    if (t == kCqThreads) t = kCqThreads - 1;
    int bin = t * per;
    for (; bin < nb - 1 && bin < (t + 1) * per; ++bin) {
      if (cum + h[bin] > s.rank) break;
      cum += h[bin];
    }
    s.prefix |= (unsigned)bin << shift;
    s.rank -= cum;
    st[sel] = s;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nb; j += kCqThreads) h[j] = 0;     // ready for the next pass
  if (last && threadIdx.x == 0) {
    s = st[sel];
    float v = 0.f;                                                   // class without pixels: 0 (:196-198)
    if (s.active) v = s.prefix == 0xffffffffu ? __int_as_float(0x7fc00000) : __uint_as_float(s.prefix);
    out[sel] = v;
  }
}

}  // namespace pfst

namespace {
struct CqLayout { size_t ent, counts, states, hist, total; };
inline size_t cq_up16(size_t v) { return (v + 15) / 16 * 16; }
// pred u8[n] | ent f32[n] | counts u64[256] | states[C*R] | histograms u32[C*R][2048]
inline CqLayout cq_layout(int64_t n, int C, int R) {
  CqLayout L;
  L.ent = cq_up16((size_t)n);
  L.counts = cq_up16(L.ent + (size_t)n * 4);
  L.states = L.counts + 256 * 8;
  L.hist = cq_up16(L.states + (size_t)C * R * sizeof(pfst::CqState));
  L.total = L.hist + (size_t)C * R * pfst::kCqBins * 4;
  return L;
}
}  // namespace

extern "C" {

int64_t pfst_class_quantile_ws_bytes(int64_t n, int32_t C, int32_t R) {
  if (n < 0 || C < 1 || C > 256 || R < 1 || C * R > pfst::kCqMaxSel) return 0;
  return (int64_t)cq_layout(n, C, R).total;
}

int pfst_class_quantile(const float* logits, int64_t B, int32_t C, int64_t HW, const int64_t* idx, int64_t n,
                        const double* ratios_dev, int32_t R, void* workspace, float* thr_out, void* stream) {
  if (!logits || !ratios_dev || !workspace || !thr_out || B < 0 || C < 1 || HW < 1 || n < 0 || R < 1)
    return PFST_ERR_INVALID_ARG;
  if (C > 256 || C * R > pfst::kCqMaxSel || !pfst::aligned16(workspace)) return PFST_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const CqLayout L = cq_layout(n, C, R);
  uint8_t* pred = ws;
  float* ent = reinterpret_cast<float*>(ws + L.ent);
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(ws + L.counts);
  pfst::CqState* st = reinterpret_cast<pfst::CqState*>(ws + L.states);
  unsigned* hist = reinterpret_cast<unsigned*>(ws + L.hist);
  PFST_CUDA_TRY(cudaMemsetAsync(counts, 0, 256 * 8, s), "pfst_class_quantile/memset");
  PFST_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)C * R * pfst::kCqBins * 4, s), "pfst_class_quantile/memset-hist");
  int64_t grid = (n + pfst::kCqThreads - 1) / pfst::kCqThreads;
  const int64_t cap = (int64_t)pfst::kNumSMs * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  pfst::entropy_argmax_kernel<<<(unsigned)grid, pfst::kCqThreads, 0, s>>>(logits, C, HW, idx, n, pred, ent, counts);
  PFST_CHECK_LAUNCH("pfst_class_quantile/entropy");
  pfst::quantile_init_kernel<<<(C * R + 127) / 128, 128, 0, s>>>(counts, ratios_dev, C, R, st);
  PFST_CHECK_LAUNCH("pfst_class_quantile/init");
  const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
  const unsigned hi_masks[3] = {0u, 0xffe00000u, 0xfffffc00u};
  for (int pass = 0; pass < 3; ++pass) {
    pfst::quantile_hist_kernel<<<(unsigned)grid, pfst::kCqThreads, 0, s>>>(pred, ent, n, R, st, shifts[pass],
                                                                          bits[pass], hi_masks[pass], hist);
    PFST_CHECK_LAUNCH("pfst_class_quantile/hist");
    pfst::quantile_scan_kernel<<<(unsigned)(C * R), pfst::kCqThreads, 0, s>>>(st, hist, shifts[pass], bits[pass],
                                                                             thr_out, pass == 2);
    PFST_CHECK_LAUNCH("pfst_class_quantile/scan");
  }
  return PFST_OK;
}

}  // extern "C"
