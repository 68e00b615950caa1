// V0+V1/V4 — on-device evaluation input path: arg-max of the softmax of the segmentor's
// logits fused with the integer confusion matrix (second half).
//
// Reference: EncoderDecoder.inference / simple_test,
// rsiseg/models/segmentors/encoder_decoder.py:311 (`output = F.softmax(seg_logit, dim=1)`),
// :329-338 (`seg_pred = seg_logit.argmax(dim=1)`; `.cpu().numpy()`), then per image
// dataset.pre_eval -> intersect_and_union (rsiseg/datasets/custom.py:644-682,
// rsiseg/core/evaluation/metrics.py:26-86). The reference writes the (N,C,H,W) softmax,
// an int64 arg-max map, copies it to the host (8 B/px over PCIe) and runs three CPU
// histc per image. Here: ONE pass over the NCHW logits and the label map,
// (4C + sizeof(label)) bytes per pixel, nothing but (C+1)^2 int64 per image leaves the GPU.
//
// Bit-parity rule (same as csrc/pseudo_label.cu): torch.argmax returns the FIRST index of
// the maximum of the softmax OUTPUT. That is the first index of the maximum logit unless an
// earlier class's quotient expf(x_c-m)/s rounds to the same float as 1/s; that needs a logit
// within ~1e-7 of the maximum, so only pixels with an earlier logit within 3e-4 of the
// maximum evaluate the softmax at all (exactly, as torch CUDA computes it). A NaN anywhere in
// the pixel's softmax (NaN or +inf logit, or all -inf) makes every output NaN: index 0.
#include "common.cuh"
#include "exp_exact.cuh"

namespace pfst {

constexpr int kEvThreads = 256;
constexpr int kEvWarps = kEvThreads / 32;
constexpr int kEvWarpHistMaxBins = 1024;    // 8 per-warp histograms <= 32 KB
constexpr int kEvBlockHistMaxBins = 40960;  // one 160 KB histogram per block

struct EvParams {
  const float* logits;   // (N, C, pixels) fp32
  const void* label;     // (N, pixels), nullable together with conf
  int64_t n_images;
  int64_t pixels;
  int32_t C;
  int64_t ignore_index;
  int32_t reduce_zero_label;
  const uint8_t* lut;
  int64_t* conf;         // slots x (C+1)^2 int64, nullable (arg-max only)
  int32_t per_image;
  void* pred_out;        // nullable
  int32_t pred_i64;      // pred_out element type: 1 = int64, 0 = uint8
  int32_t n_hist;        // kEvWarps: per-warp shared; 1: per-block shared; 0: global atomics
  int64_t span_units;    // work units (VEC pixels) per block
};

// Matrix row of a raw label after label_map / reduce_zero_label / ignore
// (metrics.py:66-73): 0..C-1 in range, C out of range, C+1 ignored (never counted).
__device__ __forceinline__ int ev_row(int64_t lab, const EvParams& q, const uint8_t* __restrict__ lut) {
  if (lut && lab >= 0 && lab < 256) lab = lut[lab];
  if (q.reduce_zero_label) lab = (lab == 0 || lab == 255) ? 255 : lab - 1;
  if (lab == q.ignore_index) return q.C + 1;
  return (lab >= 0 && lab < q.C) ? (int)lab : q.C;
}

// Rare path: an earlier class is within 3e-4 of the maximum logit. Evaluates the softmax
// exactly as torch CUDA does (class-order sum of expf(x-m), IEEE division) and returns the
// first index whose probability equals the maximum probability 1/s.
__device__ __noinline__ int ev_tie_label(const float* __restrict__ px, int C, int64_t HW, float m, int am) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    const ExpParts e = exp_split(px[(int64_t)c * HW] - m);
    s = __fmaf_rn(e.scale, e.mant, s);
  }
  const float top = 1.0f / s;
  int label = am;
  for (int c = am - 1; c >= 0; --c) {
    const float d = px[(int64_t)c * HW] - m;
    if (d > -3.0e-4f && exp_exact(d) / s == top) label = c;
  }
  return label;
}

template <int CMAX>
__device__ __forceinline__ int ev_pixel(const float (&x)[CMAX], const float* __restrict__ px, int64_t HW) {
  float m = x[0];
  int am = 0;
#pragma unroll
  for (int c = 1; c < CMAX; ++c)
    if (x[c] > m) { m = x[c]; am = c; }
  // d_c = x_c - m is <= 0, -inf or NaN; the softmax holds a NaN iff some d_c is NaN, and the
  // sum of the d_c is NaN iff one of them is (-inf + -inf stays -inf)
  float dsum = 0.f;
  bool near = false;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    const float d = x[c] - m;
    dsum += d;
    near |= (c < am) && (d > -3.0e-4f);
  }
  if (dsum != dsum) return 0;
  if (near) return ev_tie_label(px, CMAX, HW, m, am);
  return am;
}

// Four consecutive labels kept as loaded (one register for uint8 maps) and unpacked on use.
template <typename LT> struct EvLab4;
template <> struct EvLab4<uint8_t> {
  unsigned w = 0;
  __device__ __forceinline__ void load(const uint8_t* __restrict__ p) { w = __ldcs(reinterpret_cast<const unsigned*>(p)); }
  __device__ __forceinline__ int64_t get(int k) const { return (int64_t)((w >> (8 * k)) & 0xffu); }
};
template <> struct EvLab4<int32_t> {
  int4 v = {0, 0, 0, 0};
  __device__ __forceinline__ void load(const int32_t* __restrict__ p) { v = ldg_stream_i4(p); }
  __device__ __forceinline__ int64_t get(int k) const { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
};
template <> struct EvLab4<int64_t> {
  longlong2 a = {0, 0}, b = {0, 0};
  __device__ __forceinline__ void load(const int64_t* __restrict__ p) { a = ldg_stream_l2(p); b = ldg_stream_l2(p + 2); }
  __device__ __forceinline__ int64_t get(int k) const { return k == 0 ? a.x : (k == 1 ? a.y : (k == 2 ? b.x : b.y)); }
};

// bin (row*(C+1)+col) of a raw label and an in-range prediction; >= bins when ignored
__device__ __forceinline__ unsigned ev_bin(int64_t lab, unsigned pred, const unsigned* row_off,
                                           const EvParams& q) {
  const unsigned roff = ((unsigned long long)lab < 256ull) ? row_off[lab]
                                                           : (unsigned)ev_row(lab, q, nullptr) * (unsigned)(q.C + 1);
  return roff + pred;
}

// One warp-aggregated update: equal bins of the warp collapse into one add by their leader.
// Must be reached by all 32 lanes.
__device__ __forceinline__ void ev_count(unsigned bin, unsigned n, unsigned bins, unsigned* hist,
                                         int64_t* out, int n_hist) {
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin < bins && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
    const unsigned cnt = n * (unsigned)__popc(peers);
    if (n_hist) atomicAdd(&hist[bin], cnt);
    else atomicAdd(reinterpret_cast<unsigned long long*>(out) + bin, (unsigned long long)cnt);
  }
}

// Block-level plumbing shared by both kernels: histogram set-up, per-image flush.
struct EvBlock {
  unsigned* smem;
  unsigned* hist;
  unsigned bins;
  int n_hist;
  bool on;
  __device__ __forceinline__ void init(const EvParams& q, unsigned* smem_, unsigned* row_off) {
    const int tid = threadIdx.x;
    const unsigned C1 = (unsigned)q.C + 1;
    smem = smem_;
    bins = C1 * C1;
    n_hist = q.n_hist;
    on = q.conf != nullptr;
    row_off[tid] = (unsigned)ev_row(tid, q, q.lut) * C1;   // kEvThreads == 256
    for (unsigned i = tid; i < (unsigned)n_hist * bins; i += kEvThreads) smem[i] = 0u;
    hist = smem + (n_hist == kEvWarps ? (unsigned)(tid >> 5) * bins : 0u);
    __syncthreads();
  }
  // add the block's counts into the image's matrix and clear them (all threads)
  __device__ __forceinline__ void flush(int64_t* out) {
    if (!on || n_hist == 0) return;
    __syncthreads();
    for (unsigned b = threadIdx.x; b < bins; b += kEvThreads) {
      unsigned s = 0;
      for (int w = 0; w < n_hist; ++w) {
        s += smem[(unsigned)w * bins + b];
        smem[(unsigned)w * bins + b] = 0u;
      }
      if (s) atomicAdd(reinterpret_cast<unsigned long long*>(out) + b, (unsigned long long)s);
    }
    __syncthreads();
  }
};

// C <= 8, 128-bit loads: four consecutive pixels per thread, next item's loads in flight
// while the current one is ranked and counted. Blocks own contiguous spans of the pixel
// stream and flush their histogram once per image they touch.
// Counting (as csrc/confusion.cu, strategy 0): every thread owns a private column of 16-bit
// counters, priv[bin][tid] — plain LDS/ADD/STS, no atomics, no warp votes, independent of the
// label distribution; columns are summed and flushed per image (and before they can wrap).
template <int C, typename LT>
__global__ void __launch_bounds__(kEvThreads)
argmax_confusion_vec4_kernel(const EvParams q) {
  extern __shared__ __align__(16) unsigned ev_smem[];
  __shared__ unsigned row_off[256];   // raw label byte -> element offset of its matrix row in priv
  constexpr unsigned C1 = C + 1;
  constexpr int kBins = (int)(C1 * C1);            // flushed bins; the scratch row follows them
  constexpr int kRowsTotal = (int)(C1 * (C1 + 1));
  constexpr int kWordsPerBin = kEvThreads / 2;     // two 16-bit counters per 32-bit word
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool on = q.conf != nullptr;
  unsigned short* priv = reinterpret_cast<unsigned short*>(ev_smem);
  row_off[tid] = (unsigned)ev_row(tid, q, q.lut) * C1 * kEvThreads;   // kEvThreads == 256
  if (on)
    for (int i = tid; i < kRowsTotal * kWordsPerBin; i += kEvThreads) ev_smem[i] = 0u;
  __syncthreads();

  // add the block's counts into the image's matrix and clear them (all threads)
  auto flush = [&](int64_t* out) {
    if (!on) return;
    __syncthreads();
    for (int b = warp; b < kBins; b += kEvWarps) {
      unsigned sum = 0;
#pragma unroll
      for (int k = 0; k < kWordsPerBin / 32; ++k) {
        const unsigned w = ev_smem[b * kWordsPerBin + k * 32 + lane];
        sum += (w & 0xffffu) + (w >> 16);
        ev_smem[b * kWordsPerBin + k * 32 + lane] = 0u;
      }
      sum = warp_sum(sum);
      if (lane == 0 && sum) atomicAdd(reinterpret_cast<unsigned long long*>(out) + b, (unsigned long long)sum);
    }
    __syncthreads();
  };
  constexpr int kItersPerFlush = 65535 / 4;   // a private counter gains at most 4 per iteration

  const LT* __restrict__ label = static_cast<const LT*>(q.label);
  const int64_t HW = q.pixels;
  const int64_t upi = HW / 4;   // units per image
  const int64_t total_units = upi * q.n_images;
  int64_t u = (int64_t)blockIdx.x * q.span_units;
  int64_t span_end = u + q.span_units;
  if (span_end > total_units) span_end = total_units;

  float cur[4][C], nxt[4][C];
  EvLab4<LT> lcur, lnxt;

  while (u < span_end) {
    const int64_t img = u / upi;
    const int64_t u0 = img * upi;
    int64_t seg_end = u0 + upi;
    if (seg_end > span_end) seg_end = span_end;
    const float* __restrict__ lg = q.logits + img * (int64_t)C * HW;
    const LT* __restrict__ lb = on ? label + img * HW : nullptr;
    int64_t* out = on ? q.conf + (q.per_image ? img : 0) * (int64_t)kBins : nullptr;

    auto load = [&](int64_t unit, float (&X)[4][C], EvLab4<LT>& L) {
      const int64_t off = (unit - u0) * 4;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 v = ldg_stream_f4(lg + (int64_t)c * HW + off);
        X[0][c] = v.x; X[1][c] = v.y; X[2][c] = v.z; X[3][c] = v.w;
      }
      if (lb) L.load(lb + off);
    };

    if (u + tid < seg_end) load(u + tid, cur, lcur);
    int iters = 0;
    for (int64_t base = u; base < seg_end; base += kEvThreads) {
      if (++iters > kItersPerFlush) {   // block-uniform
        flush(out);
        iters = 1;
      }
      const int64_t unit = base + tid;
      const bool live = unit < seg_end;
      if (unit + kEvThreads < seg_end) load(unit + kEvThreads, nxt, lnxt);

      unsigned pred[4] = {0u, 0u, 0u, 0u};
      if (live) {
        const float* px = lg + (unit - u0) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) pred[k] = (unsigned)ev_pixel<C>(cur[k], px + k, HW);
        if (q.pred_out) {
          const int64_t o = img * HW + (unit - u0) * 4;
          if (q.pred_i64) {
            int64_t* po = static_cast<int64_t*>(q.pred_out) + o;
            stg_l2(po, make_longlong2((int64_t)pred[0], (int64_t)pred[1]));
            stg_l2(po + 2, make_longlong2((int64_t)pred[2], (int64_t)pred[3]));
          } else {
            *reinterpret_cast<unsigned*>(static_cast<uint8_t*>(q.pred_out) + o) =
                pred[0] | (pred[1] << 8) | (pred[2] << 16) | (pred[3] << 24);
          }
        }
        if (on) {
          // all four counter addresses first (independent LDS lookups), then the updates
          unsigned idx[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int64_t lab = lcur.get(k);
            const unsigned roff = ((unsigned long long)lab < 256ull)
                                      ? row_off[lab]
                                      : (unsigned)ev_row(lab, q, nullptr) * C1 * kEvThreads;
            idx[k] = roff + pred[k] * kEvThreads + tid;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) priv[idx[k]] += 1;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) cur[k][c] = nxt[k][c];
      lcur = lnxt;
    }
    flush(out);
    u = seg_end;
  }
}

// Any C, any alignment: one pixel per thread, the pixel's logits are walked twice in
// memory (the second walk is served from L1).
template <typename LT>
__global__ void __launch_bounds__(kEvThreads)
argmax_confusion_generic_kernel(const EvParams q) {
  extern __shared__ __align__(16) unsigned ev_smem[];
  __shared__ unsigned row_off[256];
  EvBlock blk;
  blk.init(q, ev_smem, row_off);
  const int tid = threadIdx.x;
  const LT* __restrict__ label = static_cast<const LT*>(q.label);
  const int C = q.C;
  const int64_t HW = q.pixels;
  const int64_t total_units = HW * q.n_images;
  int64_t u = (int64_t)blockIdx.x * q.span_units;
  int64_t span_end = u + q.span_units;
  if (span_end > total_units) span_end = total_units;

  while (u < span_end) {
    const int64_t img = u / HW;
    const int64_t u0 = img * HW;
    int64_t seg_end = u0 + HW;
    if (seg_end > span_end) seg_end = span_end;
    const float* __restrict__ lg = q.logits + img * (int64_t)C * HW;
    int64_t* out = blk.on ? q.conf + (q.per_image ? img : 0) * (int64_t)blk.bins : nullptr;

    for (int64_t base = u; base < seg_end; base += kEvThreads) {
      const int64_t unit = base + tid;
      const bool live = unit < seg_end;
      unsigned pred = 0u;
      unsigned bin = blk.bins;
      if (live) {
        const float* px = lg + (unit - u0);
        float m = px[0];
        int am = 0;
        for (int c = 1; c < C; ++c) {
          const float v = px[(int64_t)c * HW];
          if (v > m) { m = v; am = c; }
        }
        float dsum = 0.f;
        bool near = false;
        for (int c = 0; c < C; ++c) {
          const float d = px[(int64_t)c * HW] - m;
          dsum += d;
          near |= (c < am) && (d > -3.0e-4f);
        }
        int lab = am;
        if (dsum != dsum) lab = 0;
        else if (near) lab = ev_tie_label(px, C, HW, m, am);
        pred = (unsigned)lab;
        if (q.pred_out) {
          if (q.pred_i64) static_cast<int64_t*>(q.pred_out)[unit] = (int64_t)pred;
          else static_cast<uint8_t*>(q.pred_out)[unit] = (uint8_t)pred;
        }
        if (blk.on) bin = ev_bin((int64_t)label[unit], pred, row_off, q);
      }
      if (blk.on) ev_count(bin, 1u, blk.bins, blk.hist, out, blk.n_hist);
    }
    blk.flush(out);
    u = seg_end;
  }
}

template <typename K>
static int launch_ev(K kernel, EvParams q, int vec, size_t smem, cudaStream_t s) {
  // static shared memory (row_off) counts against the 48 KB default limit too
  if (smem + 2048 > 48 * 1024)
    PFST_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "pfst_argmax_confusion/attr");
  int occ = 0;
  PFST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kEvThreads, smem),
                "pfst_argmax_confusion/occupancy");
  if (occ < 1) return PFST_ERR_UNSUPPORTED;
  // one resident wave; every block owns one contiguous span of whole tiles
  const int64_t total_units = (q.pixels / vec) * q.n_images;
  const int64_t tile = kEvThreads;
  int64_t grid = (total_units + tile - 1) / tile;
  const int64_t cap = (int64_t)kNumSMs * occ;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  int64_t span = (total_units + grid - 1) / grid;
  span = (span + tile - 1) / tile * tile;
  grid = (total_units + span - 1) / span;
  q.span_units = span;
  kernel<<<(unsigned)grid, kEvThreads, smem, s>>>(q);
  PFST_CHECK_LAUNCH("pfst_argmax_confusion");
  return PFST_OK;
}

// private 16-bit counter columns of the vec4 kernel: (C+1)(C+2) bins x 256 threads
static size_t ev_private_smem(const EvParams& q) {
  return q.conf ? (size_t)(q.C + 1) * (q.C + 2) * kEvThreads * sizeof(unsigned short) : 0;
}

template <typename LT>
static int launch_ev_generic(EvParams q, cudaStream_t s) {
  const int64_t bins = (int64_t)(q.C + 1) * (q.C + 1);
  q.n_hist = !q.conf ? 0 : (bins <= kEvWarpHistMaxBins ? kEvWarps : (bins <= kEvBlockHistMaxBins ? 1 : 0));
  return launch_ev(argmax_confusion_generic_kernel<LT>, q, 1, (size_t)q.n_hist * bins * sizeof(unsigned), s);
}

template <typename LT>
static int dispatch_ev(const EvParams& q, cudaStream_t s) {
  const bool vec4 = q.C <= 8 && (q.pixels % 4 == 0) && aligned16(q.logits) &&
                    (!q.label || aligned16(q.label)) && (!q.pred_out || aligned16(q.pred_out));
  if (vec4) {
    const size_t smem = ev_private_smem(q);
    switch (q.C) {
      case 1: return launch_ev(argmax_confusion_vec4_kernel<1, LT>, q, 4, smem, s);
      case 2: return launch_ev(argmax_confusion_vec4_kernel<2, LT>, q, 4, smem, s);
      case 3: return launch_ev(argmax_confusion_vec4_kernel<3, LT>, q, 4, smem, s);
      case 4: return launch_ev(argmax_confusion_vec4_kernel<4, LT>, q, 4, smem, s);
      case 5: return launch_ev(argmax_confusion_vec4_kernel<5, LT>, q, 4, smem, s);
      case 6: return launch_ev(argmax_confusion_vec4_kernel<6, LT>, q, 4, smem, s);
      case 7: return launch_ev(argmax_confusion_vec4_kernel<7, LT>, q, 4, smem, s);
      default: return launch_ev(argmax_confusion_vec4_kernel<8, LT>, q, 4, smem, s);
    }
  }
  return launch_ev_generic<LT>(q, s);
}

}  // namespace pfst

extern "C" int pfst_argmax_confusion(const float* logits, int64_t n_images, int32_t C, int64_t pixels,
                                     const void* label, int32_t label_dtype, int64_t ignore_index,
                                     int32_t reduce_zero_label, const uint8_t* lut, int64_t* conf,
                                     int32_t per_image, void* pred_out, int32_t pred_dtype,
                                     void* stream) {
  if (n_images < 0 || pixels < 0 || C < 1 || C > 255) return PFST_ERR_INVALID_ARG;
  if (pred_out && pred_dtype != PFST_DT_U8 && pred_dtype != PFST_DT_I64) return PFST_ERR_INVALID_ARG;
  if (n_images == 0 || pixels == 0) return PFST_OK;               // empty tensors may carry NULL pointers
  if (!logits) return PFST_ERR_INVALID_ARG;
  if (!conf && !pred_out) return PFST_ERR_INVALID_ARG;            // nothing to produce
  if ((conf != nullptr) != (label != nullptr)) return PFST_ERR_INVALID_ARG;
  if (pixels > 0x7fffffffll) return PFST_ERR_UNSUPPORTED;         // 32-bit per-image counters
  pfst::EvParams q;
  q.logits = logits;
  q.label = label;
  q.n_images = n_images;
  q.pixels = pixels;
  q.C = C;
  q.ignore_index = ignore_index;
  q.reduce_zero_label = reduce_zero_label;
  q.lut = lut;
  q.conf = conf;
  q.per_image = per_image;
  q.pred_out = pred_out;
  q.pred_i64 = pred_dtype == PFST_DT_I64 ? 1 : 0;
  q.n_hist = 0;
  q.span_units = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!label) return pfst::dispatch_ev<uint8_t>(q, s);
  switch (label_dtype) {
    case PFST_DT_U8: return pfst::dispatch_ev<uint8_t>(q, s);
    case PFST_DT_I32: return pfst::dispatch_ev<int32_t>(q, s);
    case PFST_DT_I64: return pfst::dispatch_ev<int64_t>(q, s);
    default: return PFST_ERR_INVALID_ARG;
  }
}
