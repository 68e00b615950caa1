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
// Label handling (label_map LUT, reduce_zero_label, ignore_index, range check) is
// folded into ONE 256-entry shared-memory table built per block: raw label byte
// -> element offset of the matrix row inside the block histogram (ignored labels
// point at a scratch row that is never flushed). Per pixel that is one LDS.
//
// Histogramming strategy (SURVEY.md §7 "histogram contention"):
//   * (C+1)(C+2) <= 192 counters (C <= 12: ISPRS 6 classes, Inria 2): every thread
//     owns a private column of 16-bit counters, hist[bin][tid] (flushed before they
//     can wrap), so updates are plain LDS/ADD/STS with NO atomics and no dependence
//     on the label distribution, even for worst-case uniformly random labels;
//   * larger C: one shared histogram per block, warp-aggregated (match.any)
//     shared atomics; beyond the shared-memory budget, warp-aggregated global
//     atomics.
// Blocks own a contiguous span of the pixel stream (persistent grid, one resident
// wave), prefetch the next tile while counting the current one, and flush their
// histogram once per image they touch, so global atomics are O(grid * bins).
#include <stdlib.h>

#include "common.cuh"

namespace pfst {

constexpr int kCfThreads = 256;
constexpr int kCfUnroll = 4;
#ifndef PFST_CF_PRIV_ATOMIC
#define PFST_CF_PRIV_ATOMIC 0
#endif
// private counters: 16-bit LDS/ADD/STS (0) or 32-bit fire-and-forget shared atomics (1)
constexpr bool kCfPrivAtomic = PFST_CF_PRIV_ATOMIC != 0;
constexpr int kCfPrivWordsPerBin = kCfPrivAtomic ? 256 : 128;   // 32-bit words per bin (256 threads)
constexpr int kCfPrivateMaxRows = 96 * 1024 / (4 * kCfPrivWordsPerBin);  // 96 KB per block
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

// Row of a label value after label_map / reduce_zero_label / ignore: 0..C-1 in
// range, C out of range, C+1 ignored (scratch row).
__device__ __forceinline__ int cf_row(int64_t lab, const CfParams& q, const uint8_t* __restrict__ lut) {
  if (lut && lab >= 0 && lab < 256) lab = lut[lab];
  if (q.reduce_zero_label) lab = (lab == 0 || lab == 255) ? 255 : lab - 1;
  if (lab == q.ignore_index) return q.C + 1;
  return (lab >= 0 && lab < q.C) ? (int)lab : q.C;
}

template <typename PT>
__device__ __forceinline__ unsigned cf_col(PT pred, unsigned C) {
  if constexpr (sizeof(PT) == 8) {
    return ((unsigned long long)pred < (unsigned long long)C) ? (unsigned)pred : C;
  } else {
    return min((unsigned)pred, C);   // negative int32 wraps to a large unsigned -> C
  }
}

// STRAT 0: private per-thread counters; 1: shared atomics; 2: global atomics
template <typename PT, typename LT, int UNIT, int STRAT>
__global__ void __launch_bounds__(kCfThreads)
confusion_kernel(const CfParams q) {
  extern __shared__ __align__(16) unsigned cf_smem[];
  __shared__ unsigned row_off[256];   // raw label byte -> element offset of its matrix row
  const int tid = threadIdx.x;
  const unsigned C = (unsigned)q.C, C1 = C + 1;
  const int bins = (int)(C1 * C1);            // flushed bins; the scratch row follows them
  const int rows_total = (int)(C1 * (C1 + 1));
  // element stride between consecutive bins: STRAT 0 interleaves the 256 private columns
  constexpr unsigned kBinStride = STRAT == 0 ? kCfThreads : 1;
  row_off[tid] = (unsigned)cf_row(tid, q, q.lut) * C1 * kBinStride;   // kCfThreads == 256
  // STRAT 0 packs two 16-bit private counters per 32-bit word
  unsigned short* priv = reinterpret_cast<unsigned short*>(cf_smem);
  const int hist_words = STRAT == 0 ? rows_total * kCfPrivWordsPerBin : (STRAT == 1 ? rows_total : 0);
  for (int i = tid; i < hist_words; i += kCfThreads) cf_smem[i] = 0u;
  __syncthreads();

  const PT* __restrict__ pred = static_cast<const PT*>(q.pred);
  const LT* __restrict__ label = static_cast<const LT*>(q.label);
  const int64_t upi = q.pixels / UNIT;  // units per image (pixels % UNIT == 0 by dispatch)
  const int64_t total_units = upi * q.n_images;
  int64_t u = (int64_t)blockIdx.x * q.span_units;
  int64_t span_end = u + q.span_units;
  if (span_end > total_units) span_end = total_units;
  constexpr int64_t kTile = (int64_t)kCfThreads * kCfUnroll;

  PT pv[kCfUnroll][UNIT], pn[kCfUnroll][UNIT];
  LT lv[kCfUnroll][UNIT], ln[kCfUnroll][UNIT];
  auto load_tile = [&](int64_t base, int64_t seg_end, PT (&P)[kCfUnroll][UNIT], LT (&L)[kCfUnroll][UNIT]) {
#pragma unroll
    for (int j = 0; j < kCfUnroll; ++j) {
      const int64_t unit = base + j * kCfThreads + tid;
      if (unit < seg_end) {
        load_units<PT, UNIT>(pred + unit * UNIT, P[j]);
        load_units<LT, UNIT>(label + unit * UNIT, L[j]);
      }
    }
  };

  // add the block's counts into the image's matrix and clear them
  auto flush = [&](int64_t* out) {
    if (STRAT == 0) {
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      for (int b = warp; b < rows_total; b += kCfThreads / 32) {
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < kCfPrivWordsPerBin / 32; ++k) {
          const unsigned w = cf_smem[b * kCfPrivWordsPerBin + k * 32 + lane];
          s += kCfPrivAtomic ? w : (w & 0xffffu) + (w >> 16);
          cf_smem[b * kCfPrivWordsPerBin + k * 32 + lane] = 0u;
        }
        s = warp_sum(s);
        if (lane == 0 && s && b < bins)
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
  };
  // a private 16-bit counter gains at most kCfUnroll*UNIT per tile
  constexpr int kTilesPerFlush = kCfPrivAtomic ? 0x7fffffff : 65535 / (kCfUnroll * UNIT);

  while (u < span_end) {
    const int64_t img = u / upi;
    int64_t seg_end = (img + 1) * upi;
    if (seg_end > span_end) seg_end = span_end;
    int64_t* out = q.conf + img * bins;

    load_tile(u, seg_end, pv, lv);
    int tiles_since_flush = 0;
    for (int64_t base = u; base < seg_end; base += kTile) {
      if (STRAT == 0 && ++tiles_since_flush > kTilesPerFlush) {
        flush(out);
        tiles_since_flush = 1;
      }
      // software pipeline: next tile's loads are in flight while this one is counted
      if (base + kTile < seg_end) load_tile(base + kTile, seg_end, pn, ln);
      // 1) all bin indices of the tile first (independent LDS lookups, batched) ...
      unsigned idx[kCfUnroll][UNIT];
      const unsigned scratch = C1 * C1 * kBinStride;
#pragma unroll
      for (int j = 0; j < kCfUnroll; ++j) {
        const bool live = base + j * kCfThreads + tid < seg_end;
#pragma unroll
        for (int k = 0; k < UNIT; ++k) {
          unsigned roff;
          if constexpr (sizeof(LT) == 1) {
            roff = row_off[lv[j][k]];
          } else {
            const int64_t lab = (int64_t)lv[j][k];
            roff = ((unsigned long long)lab < 256ull) ? row_off[lab]
                                                      : (unsigned)cf_row(lab, q, nullptr) * C1 * kBinStride;
          }
          idx[j][k] = live ? roff + cf_col<PT>(pv[j][k], C) * kBinStride + (STRAT == 0 ? tid : 0) : scratch + (STRAT == 0 ? tid : 0);
        }
      }
      // 2) ... then the counter updates
#pragma unroll
      for (int j = 0; j < kCfUnroll; ++j) {
#pragma unroll
        for (int k = 0; k < UNIT; ++k) {
          if (STRAT == 0) {
            if (kCfPrivAtomic) atomicAdd(&cf_smem[idx[j][k]], 1u);
            else priv[idx[j][k]] += 1;
          } else {
            // warp-aggregate equal bins, one atomic per distinct bin per warp
            const unsigned peers = __match_any_sync(0xffffffffu, idx[j][k]);
            const int leader = __ffs(peers) - 1;
            if ((int)idx[j][k] < bins && (tid & 31) == leader) {
              if (STRAT == 1) atomicAdd(&cf_smem[idx[j][k]], (unsigned)__popc(peers));
              else atomicAdd(reinterpret_cast<unsigned long long*>(out) + idx[j][k],
                             (unsigned long long)__popc(peers));
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kCfUnroll; ++j)
#pragma unroll
        for (int k = 0; k < UNIT; ++k) { pv[j][k] = pn[j][k]; lv[j][k] = ln[j][k]; }
    }

    flush(out);
    u = seg_end;
  }
}

template <typename PT, typename LT, int UNIT, int STRAT>
static int launch_cf_strat(CfParams q, size_t smem, cudaStream_t s) {
  auto k = confusion_kernel<PT, LT, UNIT, STRAT>;
  if (smem > 0) {
    PFST_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "pfst_confusion_accum/attr");
    static const bool max_carve = getenv("PFST_CONF_MAX_CARVEOUT") != nullptr;   // A/B switch (round-1 setting)
    if (max_carve)
      PFST_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared),
                    "pfst_confusion_accum/carveout");
  }
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
  const int rows_total = (q.C + 1) * (q.C + 2);   // bins + scratch row
  if (rows_total <= kCfPrivateMaxRows)
    return launch_cf_strat<PT, LT, UNIT, 0>(q, (size_t)rows_total * kCfPrivWordsPerBin * sizeof(unsigned), s);
  if (rows_total <= kCfSharedMaxBins)
    return launch_cf_strat<PT, LT, UNIT, 1>(q, (size_t)rows_total * sizeof(unsigned), s);
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
