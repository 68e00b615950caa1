// Strong augmentation, Gaussian blur of the mixed image.
//
// Reference call site: gaussian_blur, rsiseg/models/utils/dacs_transforms.py:88-107 —
//   sigma = np.random.uniform(0.15, 1.15); k = odd(~0.1 * size) per axis;
//   data = kornia.filters.GaussianBlur2d(kernel_size=(k_y,k_x), sigma=(sigma,sigma))(data)
// kornia is a third-party dependency (not under /root/reference, version unpinned): this file
// implements its published algorithm — normalised 1-D Gaussians g(x)=exp(-x^2/(2 sigma^2)),
// x = t - k//2, outer product, 'reflect' border, per-channel correlation (oracle/strong_aug.py).
// On the CPU that is a 51x51 direct convolution per 512^2 image (0.7 s per image); here it is
// ONE separable pass per tile: (4 + 4) bytes per pixel and channel of HBM traffic.
//
// Exactness notes:
//   * the weights are normalised by the sum over ALL k taps, like kornia;
//   * taps whose weight is below 2^-30 of the centre tap (|x| > 6.449 sigma) are skipped: their
//     total contribution (< 3e-9 relative) is a twentieth of one fp32 rounding of the sum, and for
//     |x| > 14.4 sigma kornia's own fp32 weights are exactly zero. For the reference's sigma
//     range that leaves <= 15 of the 51 taps;
//   * accumulation is fp32 FMA in tap order, horizontal pass first (kornia's separable order).
//
// Tiling: a block owns a 64x64 output tile of one (image, channel) plane. The tile plus its
// halo is staged in shared memory with the reflect index map, the horizontal pass writes a
// second shared array, the vertical pass writes global memory. Both passes are register
// tiled (8 outputs per thread sliding over the taps: 2 shared loads per 8 FMAs) and bank
// conflict free (lanes walk rows in the horizontal pass — odd pitch — and columns in the
// vertical pass).
#include <type_traits>

#include "common.cuh"

namespace pfst {

constexpr int kBlThreads = 256;
constexpr int kBlTile = 64;        // output tile edge
constexpr int kBlOut = 8;          // outputs per thread and task
constexpr int kBlMaxImages = 64;   // sigmas carried in the launch parameters
constexpr int kBlMaxTaps = 2 * 80 + 1;
constexpr int kBlSlack = 4;        // finite values behind every staged row / column (see blur_slide)
// taps with |x| > kBlCutoff * sigma weigh less than 2^-30 of the centre tap: sqrt(2 * 30 * ln 2)
constexpr float kBlCutoff = 6.4489403f;

struct BlurParams {
  const float* in;
  float* out;
  int n_img, C, H, W;
  int ky, kx;       // full kernel sizes (odd)
  int ry, rx;       // effective radii actually evaluated
  int tiles_y, tiles_x;
  float sigma[kBlMaxImages];
};

__device__ __forceinline__ int reflect_index(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);   // only out-of-tile (masked) positions can still be outside
}

__device__ __forceinline__ void cp_async4(unsigned smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// Normalised weights of the taps |x| <= r of a k-tap Gaussian, zero-padded to a multiple of four
// taps (one warp).
__device__ __forceinline__ void blur_weights(float* w, int k, int r, float sigma) {
  const int lane = threadIdx.x & 31;
  const float inv = 1.0f / (2.0f * sigma * sigma);
  const int half = k / 2;
  float part = 0.f;
  for (int t = lane; t < k; t += 32) {
    const float x = (float)(t - half);
    part += expf(-(x * x) * inv);
  }
  const float total = warp_sum(part);
  const int taps = 2 * r + 1, taps4 = (taps + 3) & ~3;
  for (int t = lane; t < taps4; t += 32) {
    const float x = (float)(t - r);
    w[t] = t < taps ? expf(-(x * x) * inv) / total : 0.f;
  }
}

// acc[j] = sum_t w[t] * p[(j + t) * stride], j < 8, four taps per iteration over a register window
// of 8 + 4 values (taps4 is a multiple of four, the padding weights are zero; the window reads up
// to four finite slack values behind the last real one).
// NG > 0: the number of four-tap groups is a compile-time constant — the tap loop is fully unrolled,
// the window shift becomes register renaming (round-2 candidate: the runtime loop spends a third of
// its instructions on window moves and loop control). NG == 0: runtime group count.
template <int STRIDE_IS_ONE, int NG>
__device__ __forceinline__ void blur_slide(const float* __restrict__ p, int stride, const float* __restrict__ w,
                                           int taps4, float (&acc)[kBlOut]) {
  const int st = STRIDE_IS_ONE ? 1 : stride;
  float ext[kBlOut + 4];
#pragma unroll
  for (int j = 0; j < kBlOut; ++j) {
    acc[j] = 0.f;
    ext[j] = p[j * st];
  }
  const float* nxt = p + kBlOut * st;
  auto group = [&](int t) {
    const float4 wt = *reinterpret_cast<const float4*>(w + t);
#pragma unroll
    for (int k = 0; k < 4; ++k) ext[kBlOut + k] = nxt[k * st];
    nxt += 4 * st;
#pragma unroll
    for (int j = 0; j < kBlOut; ++j) acc[j] = __fmaf_rn(wt.x, ext[j], acc[j]);
#pragma unroll
    for (int j = 0; j < kBlOut; ++j) acc[j] = __fmaf_rn(wt.y, ext[j + 1], acc[j]);
#pragma unroll
    for (int j = 0; j < kBlOut; ++j) acc[j] = __fmaf_rn(wt.z, ext[j + 2], acc[j]);
#pragma unroll
    for (int j = 0; j < kBlOut; ++j) acc[j] = __fmaf_rn(wt.w, ext[j + 3], acc[j]);
#pragma unroll
    for (int j = 0; j < kBlOut; ++j) ext[j] = ext[j + 4];
  };
  if (NG > 0) {
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) group(4 * gi);
  } else {
#pragma unroll 2
    for (int t = 0; t < taps4; t += 4) group(t);
  }
}

// one pass of a block: `tasks` (row or column, 8-output chunk) pairs, group count dispatched once
template <int STRIDE_IS_ONE, typename F>
__device__ __forceinline__ void blur_pass(int taps4, F&& body) {
  switch (taps4 >> 2) {
    case 1: body(std::integral_constant<int, 1>()); break;
    case 2: body(std::integral_constant<int, 2>()); break;
    case 3: body(std::integral_constant<int, 3>()); break;
    case 4: body(std::integral_constant<int, 4>()); break;
    default: body(std::integral_constant<int, 0>()); break;
  }
}

__global__ void __launch_bounds__(kBlThreads)
gaussian_blur_kernel(const BlurParams q) {
  extern __shared__ __align__(16) float bl_smem[];
  __shared__ __align__(16) float wy[kBlMaxTaps + 3], wx[kBlMaxTaps + 3];
  const int tid = threadIdx.x;
  const int rx = q.rx, ry = q.ry;
  const int cols = kBlTile + 2 * rx, rows = kBlTile + 2 * ry;
  const int PA = (cols + kBlSlack) | 1; // odd pitches: conflict-free row walks; >= cols + slack
  constexpr int PB = kBlTile + 1;
  float* A = bl_smem;                   // rows x PA  input tile + halo
  float* Bm = bl_smem + rows * PA + 8;  // (rows + slack) x PB  horizontally blurred

  // tile coordinates
  int tile = blockIdx.x;
  const int tx = tile % q.tiles_x; tile /= q.tiles_x;
  const int ty = tile % q.tiles_y; tile /= q.tiles_y;
  const int c = tile % q.C;
  const int b = tile / q.C;
  const int x_org = tx * kBlTile, y_org = ty * kBlTile;
  const float* __restrict__ src = q.in + ((int64_t)b * q.C + c) * q.H * (int64_t)q.W;
  float* __restrict__ dst = q.out + ((int64_t)b * q.C + c) * q.H * (int64_t)q.W;

  // this image's own effective radii (<= the launch's, which size the halo): the result of an
  // image does not depend on which other images share its launch
  const float sigma = q.sigma[b];
  const int r_img = (int)floorf(kBlCutoff * sigma);
  const int rxi = min(rx, r_img), ryi = min(ry, r_img);
  if (tid < 32) blur_weights(wx, q.kx, rxi, sigma);
  else if (tid < 64) blur_weights(wy, q.ky, ryi, sigma);

  // stage the tile and its halo (reflect border) with 4-byte cp.async copies: global -> shared
  // without a register round trip, every row of the tile in flight at once, ONE wait per block.
  // One warp per row: the row index is reflected once, a lane copies two interior columns and one
  // of the first 32 halo columns (their reflected x is row-invariant) — ~18 instructions per row
  // and thread. (A flat element loop spends ~45 integer instructions per element on a division,
  // two reflections and 64-bit addressing, which made the staging the bulk of the kernel; register
  // staging pays one DRAM latency per batch of rows.)
  {
    const int lane = tid & 31, wrp = tid >> 5;
    constexpr int kWarps = kBlThreads / 32;
    const unsigned As = (unsigned)__cvta_generic_to_shared(A);
    const bool fast = x_org + kBlTile <= q.W;
    if (fast) {
      const int hc0 = lane < rx ? lane : lane + kBlTile;        // left halo [0,rx), right [rx+64, cols)
      const bool has_h = lane < 2 * rx;
      const int gxh = reflect_index(x_org - rx + hc0, q.W);
      const unsigned o0 = 4u * (unsigned)(rx + lane), oh = 4u * (unsigned)hc0;
      const float* __restrict__ in0 = src + x_org + lane;
      const float* __restrict__ inh = src + gxh;
#pragma unroll 2
      for (int r = wrp; r < rows; r += kWarps) {
        const int64_t goff = (int64_t)reflect_index(y_org - ry + r, q.H) * q.W;
        const unsigned a = As + 4u * (unsigned)(r * PA);
        cp_async4(a + o0, in0 + goff);
        cp_async4(a + o0 + 128u, in0 + goff + 32);
        if (has_h) cp_async4(a + oh, inh + goff);
      }
      if (2 * rx > 32) {   // wide kernels: the rest of the halo, element by element
        for (int r = wrp; r < rows; r += kWarps) {
          const float* __restrict__ row = src + (int64_t)reflect_index(y_org - ry + r, q.H) * q.W;
          for (int hc = 32 + lane; hc < 2 * rx; hc += 32) {
            const int cc = hc < rx ? hc : hc + kBlTile;
            cp_async4(As + 4u * (unsigned)(r * PA + cc), row + reflect_index(x_org - rx + cc, q.W));
          }
        }
      }
    } else {
      for (int r = wrp; r < rows; r += kWarps) {
        const float* __restrict__ row = src + (int64_t)reflect_index(y_org - ry + r, q.H) * q.W;
        for (int cc = lane; cc < cols; cc += 32)
          cp_async4(As + 4u * (unsigned)(r * PA + cc), row + reflect_index(x_org - rx + cc, q.W));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // zero slack behind every row of A and below Bm: read by the last slide of a task, multiplied by
  // the zero padding weights (must be finite)
  for (int i = tid; i < rows * kBlSlack; i += kBlThreads) A[(i / kBlSlack) * PA + cols + (i % kBlSlack)] = 0.f;
  for (int i = tid; i < kBlSlack * PB; i += kBlThreads) Bm[rows * PB + i] = 0.f;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // horizontal pass: task = (row, 8-column chunk); lanes of a warp take consecutive rows
  {
    const int tasks = rows * (kBlTile / kBlOut);
    const int t4 = (2 * rxi + 4) & ~3;
    blur_pass<1>(t4, [&](auto ng) {
      for (int t = tid; t < tasks; t += kBlThreads) {
        const int chunk = t / rows, r = t - chunk * rows;
        float acc[kBlOut];
        blur_slide<1, decltype(ng)::value>(A + r * PA + chunk * kBlOut + (rx - rxi), 1, wx, t4, acc);
#pragma unroll
        for (int j = 0; j < kBlOut; ++j) Bm[r * PB + chunk * kBlOut + j] = acc[j];
      }
    });
  }
  __syncthreads();

  // vertical pass: task = (column, 8-row chunk); lanes take consecutive columns
  {
    const int tasks = kBlTile * (kBlTile / kBlOut);
    const int t4 = (2 * ryi + 4) & ~3;
    blur_pass<0>(t4, [&](auto ng) {
      for (int t = tid; t < tasks; t += kBlThreads) {
        const int chunk = t / kBlTile, cc = t - chunk * kBlTile;
        float acc[kBlOut];
        blur_slide<0, decltype(ng)::value>(Bm + (chunk * kBlOut + (ry - ryi)) * PB + cc, PB, wy, t4, acc);
        const int gx = x_org + cc;
        if (gx < q.W) {
#pragma unroll
          for (int j = 0; j < kBlOut; ++j) {
            const int gy = y_org + chunk * kBlOut + j;
            if (gy < q.H) dst[(int64_t)gy * q.W + gx] = acc[j];
          }
        }
      }
    });
  }
}

static size_t blur_smem_bytes(int ry, int rx) {
  const int cols = kBlTile + 2 * rx, rows = kBlTile + 2 * ry;
  const int PA = (cols + kBlSlack) | 1;
  return ((size_t)rows * PA + 8 + (size_t)(rows + kBlSlack) * (kBlTile + 1)) * sizeof(float);
}

// taps with |x| <= r are evaluated: weights below 2^-30 of the centre are dropped
static int blur_radius(int k, float sigma_max) {
  const int half = k / 2;
  const double r = floor((double)kBlCutoff * (double)sigma_max) + 1.0;   // >= the kernel's per-image floorf
  return r < (double)half ? (int)r : half;
}

}  // namespace pfst

extern "C" int pfst_gaussian_blur(const float* in, float* out, int64_t n_images, int32_t C, int32_t H,
                                  int32_t W, int32_t ksize_y, int32_t ksize_x, const float* sigma_host,
                                  void* stream) {
  using namespace pfst;
  if (n_images < 0 || C < 1 || H < 1 || W < 1 || !sigma_host) return PFST_ERR_INVALID_ARG;
  if (ksize_y < 1 || ksize_x < 1 || (ksize_y % 2) == 0 || (ksize_x % 2) == 0) return PFST_ERR_INVALID_ARG;
  // 'reflect' needs pad < size (torch raises otherwise)
  if (ksize_y / 2 >= H || ksize_x / 2 >= W) return PFST_ERR_INVALID_ARG;
  if (n_images == 0) return PFST_OK;
  if (!in || !out || in == out) return PFST_ERR_INVALID_ARG;
  for (int64_t i = 0; i < n_images; ++i)
    if (!(sigma_host[i] > 0.f) || !(sigma_host[i] < 1e6f)) return PFST_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t plane = (int64_t)C * H * W;
  for (int64_t b0 = 0; b0 < n_images; b0 += kBlMaxImages) {
    BlurParams q;
    q.n_img = (int)((n_images - b0) < kBlMaxImages ? (n_images - b0) : kBlMaxImages);
    float smax = 0.f;
    for (int i = 0; i < kBlMaxImages; ++i) {
      q.sigma[i] = i < q.n_img ? sigma_host[b0 + i] : 1.f;
      if (i < q.n_img && q.sigma[i] > smax) smax = q.sigma[i];
    }
    q.in = in + b0 * plane;
    q.out = out + b0 * plane;
    q.C = C; q.H = H; q.W = W;
    q.ky = ksize_y; q.kx = ksize_x;
    q.ry = blur_radius(ksize_y, smax);
    q.rx = blur_radius(ksize_x, smax);
    if (2 * q.ry + 1 > kBlMaxTaps || 2 * q.rx + 1 > kBlMaxTaps) return PFST_ERR_UNSUPPORTED;
    q.tiles_y = (H + kBlTile - 1) / kBlTile;
    q.tiles_x = (W + kBlTile - 1) / kBlTile;
    const size_t smem = blur_smem_bytes(q.ry, q.rx);
    if (smem > 200 * 1024) return PFST_ERR_UNSUPPORTED;
    if (smem + 4096 > 48 * 1024)   // the static weight arrays count against the 48 KB default too
      PFST_CUDA_TRY(cudaFuncSetAttribute(gaussian_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem), "pfst_gaussian_blur/attr");
    const int64_t grid = (int64_t)q.n_img * C * q.tiles_y * q.tiles_x;
    if (grid > 0x7fffffffll) return PFST_ERR_UNSUPPORTED;
    gaussian_blur_kernel<<<(unsigned)grid, kBlThreads, smem, s>>>(q);
    PFST_CHECK_LAUNCH("pfst_gaussian_blur");
  }
  return PFST_OK;
}
