// V1/V4 — integer confusion matrix / histc area vectors for mIoU evaluation.
//
// Reference: intersect_and_union, rsiseg/core/evaluation/metrics.py:26-86 (three
// float32 torch.histc per image on the CPU) and the integer confusion matrix
// np.bincount(n*gt+pred) of tools/confusion_matrix.py:46-65 /
// tests/test_metrics.py:9-28.
//
// One pass over (pred, label); HBM-bound at sizeof(pred)+sizeof(label) bytes per
// pixel (9 B for the reference's int64 pred + uint8 label). Output is an
// (C+1)x(C+1) int64 matrix per slot: row/col C collect out-of-range values so
// that the histc areas (which count pred and label independently) are exact.
//
// Histogramming strategy (SURVEY.md §7 "histogram contention"):
//   * bins=(C+1)^2 <= 96 (C <= 8: ISPRS 6 classes, Inria 2): every thread owns a
//     private column of 32-bit counters in shared memory, hist[bin][tid] — bank
//     = tid, so updates are conflict-free plain LDS/ADD/STS with NO atomics even
//     for worst-case uniformly random labels;
//   * larger C: one shared histogram per block, warp-aggregated (match.any)
//     shared atomics; beyond the shared-memory budget, warp-aggregated global
//     atomics.
// Blocks own a contiguous span of the pixel stream and flush their histogram
// once per image they touch, so global atomics are O(grid * bins).
#include "common.cuh"

namespace pfst {

constexpr int kCfThreads = 256;
constexpr int kCfUnroll = 4;
constexpr int kCfPrivateMaxBins = 96;          // 96 KB of private counters per block
constexpr int kCfSharedMaxBins = 40960;        // 160 KB shared histogram

struct CfParams {
  const void* pred;
  const void* label;
  int64_t n_images;     // number of output slots touched (1 if !per_image)
  int64_t pixels;       // pixels per slot
  int32_t C;
  int64_t ignore_index;
  int32_t reduce_zero_label;
  const uint8_t* lut;
  int64_t* conf;
  int64_t span_units;   // units per block
};

template <typename T, int N>
__device__ __forceinline__ void load_units(const T* __restrict__ p, T (&out)[N]) {
  constexpr int BYTES = N * (int)sizeof(T);
  if constexpr (BYTES % 16 == 0) {
    uint4 tmp[BYTES / 16];
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i) tmp[i] = __ldcs(reinterpret_cast<const uint4*>(p) + i);
    memcpy(out, tmp, BYTES);
  } else if constexpr (BYTES == 8) {
    uint2 tmp = __ldcs(reinterpret_cast<const uint2*>(p));
    memcpy(out, &tmp, 8);
  } else if constexpr (BYTES == 4) {
    unsigned tmp = __ldcs(reinterpret_cast<const unsigned*>(p));
    memcpy(out, &tmp, 4);
  } else if constexpr (BYTES == 2) {
    unsigned short tmp = __ldcs(reinterpret_cast<const unsigned short*>(p));
    memcpy(out, &tmp, 2);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = p[i];
  }
}

// returns bin index or -1 when the pixel is ignored
__device__ __forceinline__ int cf_bin(int64_t pred, int64_t lab, const CfParams& q,
                                      const uint8_t* __restrict__ lut_s) {
  if (lut_s && lab >= 0 && lab < 256) lab = lut_s[lab];
  if (q.reduce_zero_label) lab = (lab == 0 || lab == 255) ? 255 : lab - 1;
  if (lab == q.ignore_index) return -1;
  const int C = q.C;
  const int row = (lab >= 0 && lab < C) ? (int)lab : C;
  const int col = (pred >= 0 && pred < C) ? (int)pred : C;
  return row * (C + 1) + col;
}

// STRAT 0: private per-thread counters; 1: shared atomics; 2: global atomics
template <typename PT, typename LT, int UNIT, int STRAT>
__global__ void __launch_bounds__(kCfThreads)
confusion_kernel(const CfParams q) {
  extern __shared__ __align__(16) unsigned cf_smem[];
  __shared__ uint8_t lut_s[256];
  const int tid = threadIdx.x;
  const int bins = (q.C + 1) * (q.C + 1);
  const uint8_t* lut = nullptr;
  if (q.lut) {
    lut_s[tid] = q.lut[tid];  // kCfThreads == 256
    lut = lut_s;
  }
  const int hist_words = STRAT == 0 ? bins * kCfThreads : (STRAT == 1 ? bins : 0);
  for (int i = tid; i < hist_words; i += kCfThreads) cf_smem[i] = 0u;
  __syncthreads();

  const PT* __restrict__ pred = static_cast<const PT*>(q.pred);
  const LT* __restrict__ label = static_cast<const LT*>(q.label);
  const int64_t upi = q.pixels / UNIT;  // units per image (pixels % UNIT == 0 by dispatch)
  const int64_t total_units = upi * q.n_images;
  int64_t u = (int64_t)blockIdx.x * q.span_units;
  int64_t span_end = u + q.span_units;
  if (span_end > total_units) span_end = total_units;

  while (u < span_end) {
    const int64_t img = u / upi;
    int64_t seg_end = (img + 1) * upi;
    if (seg_end > span_end) seg_end = span_end;
    int64_t* out = q.conf + img * bins;

    for (int64_t base = u; base < seg_end; base += kCfThreads * kCfUnroll) {
      PT pv[kCfUnroll][UNIT];
      LT lv[kCfUnroll][UNIT];
#pragma unroll
      for (int j = 0; j < kCfUnroll; ++j) {
        const int64_t unit = base + j * kCfThreads + tid;
        if (unit < seg_end) {
          load_units<PT, UNIT>(pred + unit * UNIT, pv[j]);
          load_units<LT, UNIT>(label + unit * UNIT, lv[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < kCfUnroll; ++j) {
        const int64_t unit = base + j * kCfThreads + tid;
        const bool live = unit < seg_end;
#pragma unroll
        for (int k = 0; k < UNIT; ++k) {
          const int bin = live ? cf_bin((int64_t)pv[j][k], (int64_t)lv[j][k], q, lut) : -1;
          if (STRAT == 0) {
            if (bin >= 0) cf_smem[bin * kCfThreads + tid] += 1u;
          } else {
            // warp-aggregate equal bins, one atomic per distinct bin per warp
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            const int leader = __ffs(peers) - 1;
            if (bin >= 0 && (tid & 31) == leader) {
              if (STRAT == 1) atomicAdd(&cf_smem[bin], (unsigned)__popc(peers));
              else atomicAdd(reinterpret_cast<unsigned long long*>(out) + bin,
                             (unsigned long long)__popc(peers));
            }
          }
        }
      }
    }

    // flush this image's counts
    if (STRAT == 0) {
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      for (int b = warp; b < bins; b += kCfThreads / 32) {
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < kCfThreads / 32; ++k) {
          s += cf_smem[b * kCfThreads + k * 32 + lane];
          cf_smem[b * kCfThreads + k * 32 + lane] = 0u;
        }
        s = warp_sum(s);
        if (lane == 0 && s)
          atomicAdd(reinterpret_cast<unsigned long long*>(out) + b, (unsigned long long)s);
      }
      __syncthreads();
    } else if (STRAT == 1) {
      __syncthreads();
      for (int b = tid; b < bins; b += kCfThreads) {
        const unsigned s = cf_smem[b];
        if (s) {
          atomicAdd(reinterpret_cast<unsigned long long*>(out) + b, (unsigned long long)s);
          cf_smem[b] = 0u;
        }
      }
      __syncthreads();
    }
    u = seg_end;
  }
}

template <typename PT, typename LT, int UNIT, int STRAT>
static int launch_cf_strat(CfParams q, size_t smem, cudaStream_t s) {
  auto k = confusion_kernel<PT, LT, UNIT, STRAT>;
  if (smem > 0)
    PFST_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "pfst_confusion_accum/attr");
  int occ = 0;
  PFST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kCfThreads, smem),
                "pfst_confusion_accum/occupancy");
  if (occ < 1) return PFST_ERR_UNSUPPORTED;
  // persistent-style grid: one resident wave; every block owns one contiguous span
  const int64_t total_units = (q.pixels / UNIT) * q.n_images;
  const int64_t tile = (int64_t)kCfThreads * kCfUnroll;
  int64_t grid = (total_units + tile - 1) / tile;
  const int64_t cap = (int64_t)kNumSMs * occ;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  int64_t span = (total_units + grid - 1) / grid;
  span = (span + tile - 1) / tile * tile;
  grid = (total_units + span - 1) / span;
  q.span_units = span;
  k<<<(unsigned)grid, kCfThreads, smem, s>>>(q);
  PFST_CHECK_LAUNCH("pfst_confusion_accum");
  return PFST_OK;
}

template <typename PT, typename LT, int UNIT>
static int launch_cf(const CfParams& q, cudaStream_t s) {
  const int bins = (q.C + 1) * (q.C + 1);
  if (bins <= kCfPrivateMaxBins)
    return launch_cf_strat<PT, LT, UNIT, 0>(q, (size_t)bins * kCfThreads * sizeof(unsigned), s);
  if (bins <= kCfSharedMaxBins)
    return launch_cf_strat<PT, LT, UNIT, 1>(q, (size_t)bins * sizeof(unsigned), s);
  return launch_cf_strat<PT, LT, UNIT, 2>(q, 0, s);
}

template <typename PT, typename LT>
static int dispatch_unit(const CfParams& q, cudaStream_t s) {
  // UNIT pixels per thread-load such that the wider of the two operands is one
  // 128-bit access; falls back to scalar loads for ragged / unaligned inputs.
  constexpr int W = sizeof(PT) > sizeof(LT) ? sizeof(PT) : sizeof(LT);
  constexpr int UNIT = 16 / W;
  const bool ok = (q.pixels % UNIT == 0) && aligned16(q.pred) && aligned16(q.label);
  if (ok) return launch_cf<PT, LT, UNIT>(q, s);
  return launch_cf<PT, LT, 1>(q, s);
}

template <typename PT>
static int dispatch_label(const CfParams& q, int label_dtype, cudaStream_t s) {
  switch (label_dtype) {
    case PFST_DT_U8: return dispatch_unit<PT, uint8_t>(q, s);
    case PFST_DT_I32: return dispatch_unit<PT, int32_t>(q, s);
    case PFST_DT_I64: return dispatch_unit<PT, int64_t>(q, s);
    default: return PFST_ERR_INVALID_ARG;
  }
}

}  // namespace pfst

extern "C" int pfst_confusion_accum(const void* pred, int32_t pred_dtype, const void* label,
                                    int32_t label_dtype, int64_t n_images, int64_t pixels,
                                    int32_t C, int64_t ignore_index, int32_t reduce_zero_label,
                                    const uint8_t* lut, int64_t* conf, int32_t per_image,
                                    void* stream) {
  if (n_images < 0 || pixels < 0 || C < 1 || C > 255 || !conf) return PFST_ERR_INVALID_ARG;
  if (n_images == 0 || pixels == 0) return PFST_OK;
  if (!pred || !label) return PFST_ERR_INVALID_ARG;
  pfst::CfParams q;
  q.pred = pred;
  q.label = label;
  if (per_image) {
    q.n_images = n_images;
    q.pixels = pixels;
  } else {
    q.n_images = 1;
    q.pixels = n_images * pixels;
  }
  q.C = C;
  q.ignore_index = ignore_index;
  q.reduce_zero_label = reduce_zero_label;
  q.lut = lut;
  q.conf = conf;
  q.span_units = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (pred_dtype) {
    case PFST_DT_U8: return pfst::dispatch_label<uint8_t>(q, label_dtype, s);
    case PFST_DT_I32: return pfst::dispatch_label<int32_t>(q, label_dtype, s);
    case PFST_DT_I64: return pfst::dispatch_label<int64_t>(q, label_dtype, s);
    default: return PFST_ERR_INVALID_ARG;
  }
}
