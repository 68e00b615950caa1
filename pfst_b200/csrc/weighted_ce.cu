// §8(f) rank 1 — pixel-weighted cross-entropy of a decode head, fused with the bilinear
// up-sampling of its logits, the top-1 accuracy and the gradient w.r.t. the LOW-resolution
// logits.
//
// Reference: BaseDecodeHead.losses, rsiseg/models/decode_heads/decode_head.py:249-283
//   seg_logit = resize(seg_logit, size=label.shape[2:], mode='bilinear', align_corners=False)
//   loss_ce   = loss_weight * mean_over_ALL_pixels( CE(seg_logit, label, ignore_index) * seg_weight )
//               (cross_entropy_loss.py:45-63 + utils.py:48-79: reduction 'mean', avg_non_ignore False)
//   acc_seg   = accuracy(seg_logit, label, ignore_index)           (accuracy.py:6-59, top-1)
// The reference materialises the (B,C,H,W) up-sampled logits (50 MB at cfg2), their
// log-softmax and, in backward, both gradients. Here one kernel reads the low-resolution
// logits (3 MB), the labels and the pixel weights once and produces the loss, the accuracy
// and d loss / d low-res logits: a block owns an 8x8 tile of low-res pixels (+1 halo) in
// shared memory, evaluates the (8s)^2 high-res pixels it covers (bilinear taps from shared
// memory, torch's align_corners=False arithmetic), and scatters the per-class gradients back
// into the shared tile; only the tile (100*C values) goes to global memory with atomics.
// HBM-bound in principle: (8 + 4 [+4]) B per high-res pixel.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace pfst {

constexpr int kCeTile = 8;                       // low-res pixels per tile side
constexpr int kCeHalo = kCeTile + 2;             // + one neighbour on each side
constexpr int kCeThreads = 256;
constexpr int kCeMaxC = 64;

struct CeParams {
  const float* logits;       // (B, C, lh, lw)
  const int64_t* labels;     // (B, H, W)
  const float* weight;       // (B, H, W) or null
  const float* class_weight; // (C) or null
  int B, C, lh, lw, H, W, s; // H = s * lh, W = s * lw
  int64_t ignore_index;
  float loss_weight;
  float* grad;               // (B, C, lh, lw) or null: d(loss_weight * mean loss) / d logits
  double* stats;             // [0] loss sum, [1] correct, [2] valid, [3] block counter
  float* out;                // [0] loss_weight * mean, [1] acc_seg
};

// torch upsample_bilinear2d, align_corners=False: src = max(scale * (dst + 0.5) - 0.5, 0)
__device__ __forceinline__ void ce_src(int dst, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
  float r = scale * ((float)dst + 0.5f) - 0.5f;
  r = r < 0.f ? 0.f : r;
  i0 = (int)r;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = r - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(kCeThreads)
weighted_ce_kernel(const CeParams P) {
  extern __shared__ __align__(16) float ce_smem[];
  float* z_s = ce_smem;                                   // [C][kCeHalo*kCeHalo] low-res logits
  float* g_s = ce_smem + (size_t)P.C * kCeHalo * kCeHalo;   // same shape: gradient accumulator
  __shared__ double red[3][kCeThreads / 32];
  const int b = blockIdx.z;
  const int ly0 = blockIdx.y * kCeTile - 1, lx0 = blockIdx.x * kCeTile - 1;   // halo origin (may be -1)
  const int64_t lplane = (int64_t)P.lh * P.lw;
  const float* zb = P.logits + (int64_t)b * P.C * lplane;
  for (int i = threadIdx.x; i < P.C * kCeHalo * kCeHalo; i += kCeThreads) {
    const int c = i / (kCeHalo * kCeHalo), r = i - c * (kCeHalo * kCeHalo);
    const int ly = ly0 + r / kCeHalo, lx = lx0 + r % kCeHalo;
    z_s[i] = (ly >= 0 && ly < P.lh && lx >= 0 && lx < P.lw) ? zb[c * lplane + (int64_t)ly * P.lw + lx] : 0.f;
    g_s[i] = 0.f;
  }
  __syncthreads();

  const float sch = (float)P.lh / (float)P.H, scw = (float)P.lw / (float)P.W;
  const int hy0 = blockIdx.y * kCeTile * P.s, hx0 = blockIdx.x * kCeTile * P.s;
  const int side = kCeTile * P.s;
  const float gscale = P.loss_weight / (float)((double)P.B * P.H * P.W);
  double loss = 0.0, correct = 0.0, valid = 0.0;
  for (int p = threadIdx.x; p < side * side; p += kCeThreads) {
    const int y = hy0 + p / side, x = hx0 + p % side;
    if (y >= P.H || x >= P.W) continue;
    const int64_t pix = ((int64_t)b * P.H + y) * P.W + x;
    const int64_t lab = P.labels[pix];
    int y0, y1, x0, x1;
    float hy0l, hy1l, wx0l, wx1l;
    ce_src(y, sch, P.lh, y0, y1, hy0l, hy1l);
    ce_src(x, scw, P.lw, x0, x1, wx0l, wx1l);
    const int i00 = (y0 - ly0) * kCeHalo + (x0 - lx0), i01 = (y0 - ly0) * kCeHalo + (x1 - lx0);
    const int i10 = (y1 - ly0) * kCeHalo + (x0 - lx0), i11 = (y1 - ly0) * kCeHalo + (x1 - lx0);
    auto up = [&](int c) {
      const float* z = z_s + c * (kCeHalo * kCeHalo);
      return hy0l * (wx0l * z[i00] + wx1l * z[i01]) + hy1l * (wx0l * z[i10] + wx1l * z[i11]);
    };
    float m = -INFINITY;
    int arg = 0;
    for (int c = 0; c < P.C; ++c) {
      const float v = up(c);
      if (v > m) { m = v; arg = c; }        // first maximum wins
    }
    const bool ign = lab == P.ignore_index || lab < 0 || lab >= P.C;
    if (!ign) {
      valid += 1.0;
      if (arg == (int)lab) correct += 1.0;
    }
    float sum = 0.f, vlab = 0.f;
    for (int c = 0; c < P.C; ++c) {
      const float v = up(c);
      sum += expf(v - m);
      if (c == (int)lab) vlab = v;
    }
    if (ign) continue;
    float wpx = P.weight ? P.weight[pix] : 1.f;
    if (P.class_weight) wpx *= P.class_weight[lab];
    loss += (double)(((m + logf(sum)) - vlab) * wpx);
    if (P.grad) {
      const float inv = 1.f / sum, k = gscale * wpx;
      const float t00 = hy0l * wx0l, t01 = hy0l * wx1l, t10 = hy1l * wx0l, t11 = hy1l * wx1l;
      for (int c = 0; c < P.C; ++c) {
        const float g = k * (expf(up(c) - m) * inv - (c == (int)lab ? 1.f : 0.f));
        float* gs = g_s + c * (kCeHalo * kCeHalo);
        atomicAdd(gs + i00, t00 * g);
        atomicAdd(gs + i01, t01 * g);
        atomicAdd(gs + i10, t10 * g);
        atomicAdd(gs + i11, t11 * g);
      }
    }
  }

  // gradient tile -> global (halo pixels belong to neighbouring tiles too: atomics)
  __syncthreads();
  if (P.grad) {
    float* gb = P.grad + (int64_t)b * P.C * lplane;
    for (int i = threadIdx.x; i < P.C * kCeHalo * kCeHalo; i += kCeThreads) {
      const float v = g_s[i];
      if (v == 0.f) continue;
      const int c = i / (kCeHalo * kCeHalo), r = i - c * (kCeHalo * kCeHalo);
      const int ly = ly0 + r / kCeHalo, lx = lx0 + r % kCeHalo;
      if (ly >= 0 && ly < P.lh && lx >= 0 && lx < P.lw) atomicAdd(gb + c * lplane + (int64_t)ly * P.lw + lx, v);
    }
  }
  // statistics: block partials in fp64, last block finalises on the device
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  loss = warp_sum(loss); correct = warp_sum(correct); valid = warp_sum(valid);
  if (lane == 0) { red[0][warp] = loss; red[1][warp] = correct; red[2][warp] = valid; }
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int wv = 0; wv < kCeThreads / 32; ++wv) v += red[threadIdx.x][wv];
    if (v != 0.0) atomicAdd(&P.stats[threadIdx.x], v);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    is_last = atomicAdd(reinterpret_cast<unsigned*>(P.stats + 3), 1u) == gridDim.x * gridDim.y * gridDim.z - 1;
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    const double s = *((volatile double*)&P.stats[0]);
    const double nc = *((volatile double*)&P.stats[1]), nv = *((volatile double*)&P.stats[2]);
    P.out[0] = P.loss_weight * (float)(s / ((double)P.B * P.H * P.W));
    // accuracy.py:50-58: (correct + eps) * (100 / (valid + eps)), eps = float32 epsilon
    const float eps = FLT_EPSILON;
    P.out[1] = ((float)nc + eps) * (float)(100.0 / (nv + (double)eps));
  }
}

}  // namespace pfst

extern "C" {

int pfst_weighted_ce(const float* logits, const int64_t* labels, const float* weight, const float* class_weight,
                     int64_t B, int32_t C, int32_t lh, int32_t lw, int32_t H, int32_t W, int64_t ignore_index,
                     float loss_weight, float* grad_logits, double* stats, float* out2, void* stream) {
  if (!logits || !labels || !stats || !out2 || B < 0 || C < 1 || lh < 1 || lw < 1 || H < 1 || W < 1)
    return PFST_ERR_INVALID_ARG;
  if (C > pfst::kCeMaxC || B > 65535) return PFST_ERR_UNSUPPORTED;
  // the fused kernel covers integer up-sampling factors (H/4 logits -> H in every shipped config)
  if (H % lh != 0 || W % lw != 0 || H / lh != W / lw) return PFST_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PFST_CUDA_TRY(cudaMemsetAsync(stats, 0, 4 * sizeof(double), s), "pfst_weighted_ce/memset");
  if (grad_logits)
    PFST_CUDA_TRY(cudaMemsetAsync(grad_logits, 0, sizeof(float) * (size_t)B * C * lh * lw, s),
                  "pfst_weighted_ce/memset-grad");
  if (B == 0) return PFST_OK;
  pfst::CeParams P{logits, labels, weight, class_weight, (int)B, C, lh, lw, H, W, H / lh,
                   ignore_index, loss_weight, grad_logits, stats, out2};
  const dim3 grid((unsigned)((lw + pfst::kCeTile - 1) / pfst::kCeTile),
                  (unsigned)((lh + pfst::kCeTile - 1) / pfst::kCeTile), (unsigned)B);
  if (grid.y > 65535) return PFST_ERR_UNSUPPORTED;
  const size_t smem = (size_t)2 * C * pfst::kCeHalo * pfst::kCeHalo * sizeof(float);
  PFST_CUDA_TRY(cudaFuncSetAttribute(pfst::weighted_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                "pfst_weighted_ce/attr");
  pfst::weighted_ce_kernel<<<grid, pfst::kCeThreads, smem, s>>>(P);
  PFST_CHECK_LAUNCH("pfst_weighted_ce");
  return PFST_OK;
}

}  // extern "C"
