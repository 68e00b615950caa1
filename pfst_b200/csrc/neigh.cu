// L2 (+ its backward) — dilated 3x3 neighbourhood dot products of decoder features.
//
// Reference: PFGSTLoss.get_sim_feat, rsiseg/models/losses/pfgst_loss.py:181-201
//   unf = nn.Unfold(k=3, dilation=d, padding=d)(feats)      # materialises 9x the tensor
//   sim = F.cosine_similarity(unf, feats.unsqueeze(4), dim=1)
// (>90 % of the loss time in the reference: 604 MB of im2col per tensor, four
// full passes, col2im in backward.) Here the feature tensor is read ONCE:
//
//   forward  (pfst_neigh_dots): per pixel n the squared norm and the four
//            "forward" dot products x_n . x_{n+delta}, delta in {(0,+d), (+d,-d),
//            (+d,0), (+d,+d)}. cos(n, n+delta_k) for all nine taps follows from
//            these five maps by symmetry (cos(n,m) = cos(m,n)), so the loss kernel
//            gathers them; out-of-image taps are the unfold's zero padding.
//   backward (pfst_neigh_grad): grad_x[n,c] = sum_k coef[n,k] * x[n+delta_k, c],
//            the gather form of d(sum cos)/dx with per-pixel coefficients computed
//            by the loss backward kernel (no atomics, no col2im).
//
// Both are HBM-bound: 4*D bytes per pixel read (+ 4*D written in backward).
// Blackwell path: (W,H,D,B) TMA tensor map, 32x16-pixel tiles with the dilation
// halo (16-byte aligned box start), 8-channel boxes, hardware zero fill for the padding, 4-stage
// full/empty mbarrier ring, one producer warp + four consumer warps; each
// consumer thread owns a 1x4 pixel strip and reads shared memory as conflict-free
// 128-bit rows. Channel ranges are split across CTAs (ksplit) so that one
// resident wave covers the GPU; partial maps are summed by the consumer kernel in a
// fixed order (deterministic, no atomics).
// Shapes TMA cannot describe (W % 4 != 0, e.g. SeasonNet's 15x15 maps) take a
// plain coalesced kernel with the same outputs.
#include "common.cuh"
#include "tma.cuh"

namespace pfst {

constexpr int kNbTW = 32;          // tile width (pixels)
constexpr int kNbTH = 16;          // tile height
constexpr int kNbCH = 8;           // channels per TMA box / pipeline stage
constexpr int kNbStages = 4;
constexpr int kNbConsumers = 128;  // 4 warps, one 1x4 strip per thread
constexpr int kNbThreads = kNbConsumers + 32;

// TMA requires the box to start on a 16-byte boundary in the innermost dimension
// (x_start % 4 == 0 for fp32), so the left halo is always kNbHL = 4 columns wide
// (>= every supported dilation) and the box is 32 + 4 + 4 = 40 columns.
constexpr int kNbHL = 4;
template <int DIL>
struct NbGeom {
  static_assert(DIL <= kNbHL, "dilation larger than the aligned halo");
  static constexpr int RS = kNbTW + 2 * kNbHL;                // smem row stride = box width (40)
  static constexpr int FWD_ROWS = kNbTH + DIL;               // rows y0 .. y0+TH-1+d
  static constexpr int BWD_ROWS = kNbTH + 2 * DIL;           // rows y0-d .. y0+TH-1+d
};

struct NeighMaps {
  CUtensorMap m[2];
};

struct NbTile {
  int split, t, b, y0, x0, c_begin, c_end;
};

__device__ __forceinline__ NbTile nb_decode(int n_units, int B, int h, int w, int D, int ksplit) {
  // blockIdx.x -> (unit = tensor or 0, b, tile_y, tile_x, split)
  const int tiles_x = (w + kNbTW - 1) / kNbTW, tiles_y = (h + kNbTH - 1) / kNbTH;
  int idx = blockIdx.x;
  NbTile o;
  o.split = idx % ksplit; idx /= ksplit;
  o.x0 = (idx % tiles_x) * kNbTW; idx /= tiles_x;
  o.y0 = (idx % tiles_y) * kNbTH; idx /= tiles_y;
  o.b = idx % B;
  o.t = idx / B;
  const int chunks = (D + kNbCH - 1) / kNbCH;
  o.c_begin = (int)((int64_t)o.split * chunks / ksplit);
  o.c_end = (int)((int64_t)(o.split + 1) * chunks / ksplit);
  (void)n_units;
  return o;
}

__device__ __forceinline__ void nb_init_barriers(uint64_t* full_bar, uint64_t* empty_bar) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kNbStages; ++s) {
      mbar_init(&full_bar[s], 1);                  // producer's arrive.expect_tx
      mbar_init(&empty_bar[s], kNbConsumers / 32); // one arrive per consumer warp
    }
    mbar_fence_init();
  }
  __syncthreads();
}

template <int ROWS, int RS>
__device__ __forceinline__ void nb_produce(const CUtensorMap* map, float* stage_buf, uint64_t* full_bar,
                                           uint64_t* empty_bar, const NbTile& tl, int x_start,
                                           int y_start) {
  constexpr int kStageFloats = kNbCH * ROWS * RS;
  constexpr uint32_t kStageBytes = kStageFloats * sizeof(float);
  tma_prefetch_desc(map);
  int it = 0;
  for (int c = tl.c_begin; c < tl.c_end; ++c, ++it) {
    const int s = it % kNbStages;
    const uint32_t ph = (uint32_t)(it / kNbStages) & 1u;
    if (it >= kNbStages) mbar_wait(&empty_bar[s], ph ^ 1u);
    mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
    tma_load_4d(stage_buf + (size_t)s * kStageFloats, map, &full_bar[s], x_start, y_start, c * kNbCH,
                tl.b);
  }
}

// ---------------------------------------------------------------- forward (TMA)
template <int DIL>
__global__ void __launch_bounds__(kNbThreads)
neigh_dots_tma_kernel(const __grid_constant__ NeighMaps maps, int n_tensors, int B, int D, int h, int w,
                      int ksplit, float* __restrict__ dots) {
  using G = NbGeom<DIL>;
  constexpr int kStageFloats = kNbCH * G::FWD_ROWS * G::RS;
  extern __shared__ __align__(128) unsigned char nb_smem[];
  float* stage_buf = reinterpret_cast<float*>(nb_smem);
  __shared__ uint64_t full_bar[kNbStages], empty_bar[kNbStages];
  const NbTile tl = nb_decode(n_tensors, B, h, w, D, ksplit);
  nb_init_barriers(full_bar, empty_bar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kNbConsumers / 32) {
    if (lane == 0)
      nb_produce<G::FWD_ROWS, G::RS>(&maps.m[tl.t], stage_buf, full_bar, empty_bar, tl, tl.x0 - kNbHL, tl.y0);
    return;
  }
  const int lx = (threadIdx.x & 7) * 4, ly = threadIdx.x >> 3;  // strip origin inside the tile
  float acc[5][4];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;

  int it = 0;
  for (int c = tl.c_begin; c < tl.c_end; ++c, ++it) {
    const int s = it % kNbStages;
    mbar_wait(&full_bar[s], (uint32_t)(it / kNbStages) & 1u);
    const float* buf = stage_buf + (size_t)s * kStageFloats + ly * G::RS + lx;
#pragma unroll
    for (int ch = 0; ch < kNbCH; ++ch) {
      // smem column of image column x is (x - x0 + 4); this strip's pixels sit at lx+4+i
      const float* ra = buf + ch * G::FWD_ROWS * G::RS;   // row y
      const float* rb = ra + DIL * G::RS;                 // row y + d
      float A[8], Bv[12];                                 // A[j] = col lx+4+j ; Bv[j] = col lx+j
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const float4 fa = *reinterpret_cast<const float4*>(ra + 4 + 4 * v);
        A[4 * v] = fa.x; A[4 * v + 1] = fa.y; A[4 * v + 2] = fa.z; A[4 * v + 3] = fa.w;
      }
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const float4 fb = *reinterpret_cast<const float4*>(rb + 4 * v);
        Bv[4 * v] = fb.x; Bv[4 * v + 1] = fb.y; Bv[4 * v + 2] = fb.z; Bv[4 * v + 3] = fb.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = A[i];
        acc[0][i] = fmaf(a, a, acc[0][i]);
        acc[1][i] = fmaf(a, A[i + DIL], acc[1][i]);             // ( 0, +d)
        acc[2][i] = fmaf(a, Bv[4 + i - DIL], acc[2][i]);        // (+d, -d)
        acc[3][i] = fmaf(a, Bv[4 + i], acc[3][i]);              // (+d,  0)
        acc[4][i] = fmaf(a, Bv[4 + i + DIL], acc[4][i]);        // (+d, +d)
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  const int y = tl.y0 + ly;
  if (y < h) {
    float* out = dots + ((((int64_t)tl.split * n_tensors + tl.t) * B + tl.b) * 5) * h * w + (int64_t)y * w;
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = tl.x0 + lx + i;
        if (x < w) out[(int64_t)k * h * w + x] = acc[k][i];
      }
  }
}

// --------------------------------------------------------------- backward (TMA)
template <int DIL>
__global__ void __launch_bounds__(kNbThreads)
neigh_grad_tma_kernel(const __grid_constant__ NeighMaps maps, const float* __restrict__ coef, int B, int D,
                      int h, int w, int ksplit, float* __restrict__ grad) {
  using G = NbGeom<DIL>;
  constexpr int kStageFloats = kNbCH * G::BWD_ROWS * G::RS;
  extern __shared__ __align__(128) unsigned char nb_smem[];
  float* stage_buf = reinterpret_cast<float*>(nb_smem);
  __shared__ uint64_t full_bar[kNbStages], empty_bar[kNbStages];
  const NbTile tl = nb_decode(1, B, h, w, D, ksplit);
  nb_init_barriers(full_bar, empty_bar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kNbConsumers / 32) {
    if (lane == 0)
      nb_produce<G::BWD_ROWS, G::RS>(&maps.m[0], stage_buf, full_bar, empty_bar, tl, tl.x0 - kNbHL,
                                     tl.y0 - DIL);
    return;
  }
  const int lx = (threadIdx.x & 7) * 4, ly = threadIdx.x >> 3;
  const int y = tl.y0 + ly, x = tl.x0 + lx;
  const bool live = (y < h) && (x < w);   // w % 4 == 0: a strip is entirely inside or outside
  float K[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) v = *reinterpret_cast<const float4*>(coef + (((int64_t)tl.b * 9 + k) * h + y) * w + x);
    K[k][0] = v.x; K[k][1] = v.y; K[k][2] = v.z; K[k][3] = v.w;
  }
  float* gout = grad + (((int64_t)tl.b * D) * h + y) * w + x;
  const int64_t plane = (int64_t)h * w;

  int it = 0;
  for (int c = tl.c_begin; c < tl.c_end; ++c, ++it) {
    const int s = it % kNbStages;
    mbar_wait(&full_bar[s], (uint32_t)(it / kNbStages) & 1u);
    const float* buf = stage_buf + (size_t)s * kStageFloats + ly * G::RS + lx;
#pragma unroll
    for (int ch = 0; ch < kNbCH; ++ch) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* r = buf + (ch * G::BWD_ROWS + ky * DIL) * G::RS;
        float R[12];                                      // R[j] = col lx+j  (pixel i at 4+i)
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const float4 f = *reinterpret_cast<const float4*>(r + 4 * v);
          R[4 * v] = f.x; R[4 * v + 1] = f.y; R[4 * v + 2] = f.z; R[4 * v + 3] = f.w;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = fmaf(K[ky * 3 + kx][i], R[4 + i + (kx - 1) * DIL], o[i]);
      }
      const int cg = c * kNbCH + ch;
      if (live && cg < D) __stcs(reinterpret_cast<float4*>(gout + cg * plane), make_float4(o[0], o[1], o[2], o[3]));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
}

// ------------------------------------------------- generic (any shape) kernels
__global__ void __launch_bounds__(128)
neigh_dots_generic_kernel(const float* __restrict__ xa, const float* __restrict__ xb, int n_tensors, int B,
                          int D, int h, int w, int dil, int ksplit, float* __restrict__ dots) {
  const int64_t plane = (int64_t)h * w;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int z = blockIdx.y;
  const int split = z % ksplit; z /= ksplit;
  const int b = z % B, t = z / B;
  if (p >= plane) return;
  const int y = p / w, x = p - y * w;
  const int c_begin = (int)((int64_t)split * D / ksplit), c_end = (int)((int64_t)(split + 1) * D / ksplit);
  const float* src = (t == 0 ? xa : xb) + ((int64_t)b * D) * plane + p;
  const bool in1 = x + dil < w, inr = y + dil < h;
  const bool in2 = inr && x - dil >= 0, in4 = inr && x + dil < w;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
  for (int c = c_begin; c < c_end; ++c) {
    const float* q = src + (int64_t)c * plane;
    const float v = __ldg(q);
    a0 = fmaf(v, v, a0);
    if (in1) a1 = fmaf(v, __ldg(q + dil), a1);
    if (in2) a2 = fmaf(v, __ldg(q + (int64_t)dil * w - dil), a2);
    if (inr) a3 = fmaf(v, __ldg(q + (int64_t)dil * w), a3);
    if (in4) a4 = fmaf(v, __ldg(q + (int64_t)dil * w + dil), a4);
  }
  float* out = dots + ((((int64_t)split * n_tensors + t) * B + b) * 5) * plane + p;
  out[0] = a0; out[plane] = a1; out[2 * plane] = a2; out[3 * plane] = a3; out[4 * plane] = a4;
}

__global__ void __launch_bounds__(128)
neigh_grad_generic_kernel(const float* __restrict__ x, const float* __restrict__ coef, int B, int D, int h,
                          int w, int dil, float* __restrict__ grad) {
  const int64_t plane = (int64_t)h * w;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= plane) return;
  const int y = p / w, xx = p - y * w;
  float K[9];
  int off[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + (k / 3 - 1) * dil, xk = xx + (k % 3 - 1) * dil;
    const bool in = yy >= 0 && yy < h && xk >= 0 && xk < w;
    K[k] = in ? coef[((int64_t)b * 9 + k) * plane + p] : 0.f;
    off[k] = in ? (yy - y) * w + (xk - xx) : 0;
  }
  const float* src = x + ((int64_t)b * D) * plane + p;
  float* dst = grad + ((int64_t)b * D) * plane + p;
  for (int c = blockIdx.z; c < D; c += gridDim.z) {
    const float* q = src + (int64_t)c * plane;
    float o = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) o = fmaf(K[k], __ldg(q + off[k]), o);
    dst[(int64_t)c * plane] = o;
  }
}

static bool nb_tma_ok(const void* p0, const void* p1, int w, int dil) {
  return (w % 4 == 0) && aligned16(p0) && (!p1 || aligned16(p1)) && (dil == 1 || dil == 2 || dil == 4) &&
         get_encode_tiled() != nullptr;
}

static int nb_splits(int64_t units, int h, int w, int D) {
  const int64_t tiles = (int64_t)((w + kNbTW - 1) / kNbTW) * ((h + kNbTH - 1) / kNbTH);
  const int64_t items = units * tiles;
  const int chunks = (D + kNbCH - 1) / kNbCH;
  const int64_t slots = (int64_t)kNumSMs * 2;
  int ks = 1;
  while (ks < 8 && ks * 2 <= chunks && items * ks * 2 <= slots) ks *= 2;
  return ks;
}

template <int DIL>
static int launch_dots_tma(const float* xa, const float* xb, int T, int B, int D, int h, int w, int ks,
                           float* dots, cudaStream_t s) {
  using G = NbGeom<DIL>;
  NeighMaps maps;
  if (!make_nchw_tensor_map(&maps.m[0], xa, B, D, h, w, G::RS, G::FWD_ROWS, kNbCH)) return PFST_ERR_CUDA;
  if (!make_nchw_tensor_map(&maps.m[1], xb ? xb : xa, B, D, h, w, G::RS, G::FWD_ROWS, kNbCH)) return PFST_ERR_CUDA;
  const size_t smem = (size_t)kNbStages * kNbCH * G::FWD_ROWS * G::RS * sizeof(float);
  auto k = neigh_dots_tma_kernel<DIL>;
  PFST_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pfst_neigh_dots/attr");
  const int64_t tiles = (int64_t)((w + kNbTW - 1) / kNbTW) * ((h + kNbTH - 1) / kNbTH);
  const int64_t grid = (int64_t)T * B * tiles * ks;
  k<<<(unsigned)grid, kNbThreads, smem, s>>>(maps, T, B, D, h, w, ks, dots);
  PFST_CHECK_LAUNCH("pfst_neigh_dots");
  return PFST_OK;
}

template <int DIL>
static int launch_grad_tma(const float* x, const float* coef, int B, int D, int h, int w, int ks, float* grad,
                           cudaStream_t s) {
  using G = NbGeom<DIL>;
  NeighMaps maps;
  if (!make_nchw_tensor_map(&maps.m[0], x, B, D, h, w, G::RS, G::BWD_ROWS, kNbCH)) return PFST_ERR_CUDA;
  maps.m[1] = maps.m[0];
  const size_t smem = (size_t)kNbStages * kNbCH * G::BWD_ROWS * G::RS * sizeof(float);
  auto k = neigh_grad_tma_kernel<DIL>;
  PFST_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pfst_neigh_grad/attr");
  const int64_t tiles = (int64_t)((w + kNbTW - 1) / kNbTW) * ((h + kNbTH - 1) / kNbTH);
  const int64_t grid = (int64_t)B * tiles * ks;
  k<<<(unsigned)grid, kNbThreads, smem, s>>>(maps, coef, B, D, h, w, ks, grad);
  PFST_CHECK_LAUNCH("pfst_neigh_grad");
  return PFST_OK;
}

}  // namespace pfst

extern "C" {

int32_t pfst_neigh_dots_splits(int64_t n_tensors, int64_t B, int32_t D, int32_t h, int32_t w) {
  if (n_tensors < 1 || B < 1 || D < 1 || h < 1 || w < 1) return 1;
  return pfst::nb_splits(n_tensors * B, h, w, D);
}

int pfst_neigh_dots(const float* x_a, const float* x_b, int64_t B, int32_t D, int32_t h, int32_t w,
                    int32_t dilation, float* dots, void* stream) {
  if (!x_a || !dots || B < 0 || D < 1 || h < 1 || w < 1 || dilation < 1) return PFST_ERR_INVALID_ARG;
  if (B == 0) return PFST_OK;
  if (B > 0x7fffffffll / 4) return PFST_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int T = x_b ? 2 : 1;
  const int ks = pfst::nb_splits((int64_t)T * B, h, w, D);
  if (pfst::nb_tma_ok(x_a, x_b, w, dilation)) {
    switch (dilation) {
      case 1: return pfst::launch_dots_tma<1>(x_a, x_b, T, (int)B, D, h, w, ks, dots, s);
      case 2: return pfst::launch_dots_tma<2>(x_a, x_b, T, (int)B, D, h, w, ks, dots, s);
      default: return pfst::launch_dots_tma<4>(x_a, x_b, T, (int)B, D, h, w, ks, dots, s);
    }
  }
  const int64_t plane = (int64_t)h * w;
  const dim3 grid((unsigned)((plane + 127) / 128), (unsigned)(T * B * ks), 1);
  if (grid.y > 65535) return PFST_ERR_UNSUPPORTED;
  pfst::neigh_dots_generic_kernel<<<grid, 128, 0, s>>>(x_a, x_b, T, (int)B, D, h, w, dilation, ks, dots);
  PFST_CHECK_LAUNCH("pfst_neigh_dots");
  return PFST_OK;
}

int pfst_neigh_grad(const float* x, const float* coef, int64_t B, int32_t D, int32_t h, int32_t w,
                    int32_t dilation, float* grad_x, void* stream) {
  if (!x || !coef || !grad_x || B < 0 || D < 1 || h < 1 || w < 1 || dilation < 1) return PFST_ERR_INVALID_ARG;
  if (B == 0) return PFST_OK;
  if (B > 65535) return PFST_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pfst::nb_tma_ok(x, grad_x, w, dilation) && pfst::aligned16(coef)) {
    const int ks = pfst::nb_splits(B, h, w, D);
    switch (dilation) {
      case 1: return pfst::launch_grad_tma<1>(x, coef, (int)B, D, h, w, ks, grad_x, s);
      case 2: return pfst::launch_grad_tma<2>(x, coef, (int)B, D, h, w, ks, grad_x, s);
      default: return pfst::launch_grad_tma<4>(x, coef, (int)B, D, h, w, ks, grad_x, s);
    }
  }
  const int64_t plane = (int64_t)h * w;
  const unsigned gx = (unsigned)((plane + 127) / 128);
  unsigned gz = (unsigned)(((int64_t)pfst::kNumSMs * 16) / ((int64_t)gx * B) + 1);
  if (gz > (unsigned)D) gz = (unsigned)D;
  if (gz > 64) gz = 64;
  pfst::neigh_grad_generic_kernel<<<dim3(gx, (unsigned)B, gz), 128, 0, s>>>(x, coef, (int)B, D, h, w, dilation,
                                                                             grad_x);
  PFST_CHECK_LAUNCH("pfst_neigh_grad");
  return PFST_OK;
}

}  // extern "C"
