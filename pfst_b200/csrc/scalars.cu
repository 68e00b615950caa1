// Log-variable bookkeeping without host round trips.
// BaseSegmentor._parse_losses (rsiseg/models/segmentors/base.py:177-222) sums the entries whose
// key contains 'loss' with Python's sum() — a left-to-right chain of fp32 adds starting from
// integer 0 — divides every log variable by the world size, all-reduces it and calls .item() on
// each one (a device->host sync per variable). Here ONE single-thread kernel reads the 0-dim
// device scalars through their pointers, writes them (divided by the world size) into a row of
// a persistent device ledger and produces the same left-to-right sum; the ledger row is
// all-reduced once per iteration and read back once per log interval.
#include "common.cuh"

namespace pfst {

constexpr int kMaxScalars = 32;

struct ScalarArgs {
  const float* p[kMaxScalars];
  float w[kMaxScalars];
};

__global__ void gather_scalars_kernel(ScalarArgs a, int n, unsigned sum_mask, float divisor,
                                      float* __restrict__ row, float* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float t = 0.f;                       // Python: 0 + v0 == v0 exactly
  for (int i = 0; i < n; ++i) {
    const float v = *a.p[i];
    if (row) row[i] = __fdiv_rn(v, divisor);
    if ((sum_mask >> i) & 1u) t = __fadd_rn(t, a.w[i] == 1.f ? v : __fmul_rn(v, a.w[i]));
  }
  if (row) row[n] = __fdiv_rn(t, divisor);
  if (total) *total = t;
}

// Several _parse_losses calls of one iteration in ONE launch: entry i belongs to the current
// segment (= one call); flags bit0: the key contains 'loss' (entry enters the segment's sum),
// bit1: last entry of its segment — the segment's sum is then written behind its entries (the
// call's 'loss' log variable) and w[i] * sum is added to the grand total (total_loss of
// pfgst.py:237,310,342: 0 + clean + mix * trg_loss_weight + aux ...), all left to right in fp32.
struct SegmentArgs {
  const float* p[kMaxScalars];
  float w[kMaxScalars];
  unsigned char flags[kMaxScalars];
};

__global__ void gather_segments_kernel(SegmentArgs a, int n, float divisor, float* __restrict__ row,
                                       float* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float t = 0.f, grand = 0.f;
  int o = 0;
  for (int i = 0; i < n; ++i) {
    const float v = *a.p[i];
    row[o++] = __fdiv_rn(v, divisor);
    if (a.flags[i] & 1) t = __fadd_rn(t, v);
    if (a.flags[i] & 2) {
      row[o++] = __fdiv_rn(t, divisor);
      grand = __fadd_rn(grand, __fmul_rn(t, a.w[i]));
      t = 0.f;
    }
  }
  if (total) *total = grand;
}

__global__ void pack_scalars_kernel(ScalarArgs a, int n, float* __restrict__ out) {
  const int i = threadIdx.x;
  if (i < n) out[i] = a.p[i] ? __fmul_rn(*a.p[i], a.w[i]) : 0.f;
}

}  // namespace pfst

extern "C" {

int pfst_gather_scalars(const float* const* ptrs_host, const float* weights_host, int32_t n, uint32_t sum_mask,
                        float divisor, float* row_out, float* total_out, void* stream) {
  if (!ptrs_host || n < 0 || n > pfst::kMaxScalars || !(divisor > 0.f)) return PFST_ERR_INVALID_ARG;
  pfst::ScalarArgs a;
  for (int i = 0; i < pfst::kMaxScalars; ++i) {
    a.p[i] = i < n ? ptrs_host[i] : nullptr;
    a.w[i] = (i < n && weights_host) ? weights_host[i] : 1.f;
    if (i < n && !a.p[i]) return PFST_ERR_INVALID_ARG;
  }
  pfst::gather_scalars_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, n, sum_mask, divisor, row_out,
                                                                               total_out);
  PFST_CHECK_LAUNCH("pfst_gather_scalars");
  return PFST_OK;
}

int pfst_gather_segments(const float* const* ptrs_host, const float* weights_host, const uint8_t* flags_host,
                         int32_t n, float divisor, float* row_out, float* total_out, void* stream) {
  if (!ptrs_host || !weights_host || !flags_host || !row_out || n < 1 || n > pfst::kMaxScalars || !(divisor > 0.f))
    return PFST_ERR_INVALID_ARG;
  pfst::SegmentArgs a;
  for (int i = 0; i < pfst::kMaxScalars; ++i) {
    a.p[i] = i < n ? ptrs_host[i] : nullptr;
    a.w[i] = i < n ? weights_host[i] : 0.f;
    a.flags[i] = i < n ? flags_host[i] : 0;
    if (i < n && !a.p[i]) return PFST_ERR_INVALID_ARG;
  }
  pfst::gather_segments_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, n, divisor, row_out, total_out);
  PFST_CHECK_LAUNCH("pfst_gather_segments");
  return PFST_OK;
}

int pfst_pack_scalars(const float* const* ptrs_host, const float* weights_host, int32_t n, float* out,
                      void* stream) {
  if (!ptrs_host || !out || n < 1 || n > pfst::kMaxScalars) return PFST_ERR_INVALID_ARG;
  pfst::ScalarArgs a;
  for (int i = 0; i < pfst::kMaxScalars; ++i) {
    a.p[i] = i < n ? ptrs_host[i] : nullptr;
    a.w[i] = (i < n && weights_host) ? weights_host[i] : 1.f;
  }
  pfst::pack_scalars_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, n, out);
  PFST_CHECK_LAUNCH("pfst_pack_scalars");
  return PFST_OK;
}

}  // extern "C"
