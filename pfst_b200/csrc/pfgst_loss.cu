// L1, L3-L6 — the per-pixel part of PFGSTLoss, forward statistics and backward maps.
//
// Reference: rsiseg/models/losses/pfgst_loss.py
//   forward :44-140, get_cross_prob_map_diag :142-159, get_sim_losses :203-234
// (SURVEY.md Appendix B restates it per pixel). The reference runs ~60 ATen
// kernels with 3+ host syncs (boolean-mask gathers, `if ignore_mask.sum() > 1`);
// here ONE kernel reduces everything into 9 fp64 sums (the last block turns them
// into the six losses on the device) and ONE backward kernel produces
//   * coef (B,9,fh,fw): per-pixel coefficients of the gather-form gradient of the
//     source cosine statistics w.r.t. x_src (consumed by pfst_neigh_grad), and
//   * grad_logits: d(loss_sim_pos + loss_sim_neg)/d logits_trg through p only
//     (q is detached: detach_unfold=True, configs/pfst/*.py:44).
//
// Everything here works on maps of (B, few, g, g) floats — a few MB; the cost is
// launch latency, not bandwidth. The heavy tensors are touched by neigh.cu only.
#include <math.h>

#include "common.cuh"

namespace pfst {

constexpr int kLossThreads = 64;
constexpr int kMaxC = 64;   // classes held in registers per pixel

struct LossParams {
  // neighbourhood dot maps from pfst_neigh_dots: (ksplit, 2, B, 5, fh, fw); tensor 0 = x_ema, 1 = x_src
  const float* dots;
  int ksplit;
  int B, fh, fw, up;          // feature grid, loss grid = feature grid * up
  const float* logits;        // (B, C, lh, lw) student logits on the mixed image
  int C, lh, lw;
  float lscale_h, lscale_w;   // torch nearest: src = min(floor(dst * scale), in - 1)
  const int64_t* gt;          // (B,1,gt_h,gt_w) source labels
  const int64_t* mix;         // (B,1,gt_h,gt_w) ClassMix masks
  int gt_h, gt_w;
  float gscale_h, gscale_w;
  int gh, gw;                 // loss grid
  int dil;                    // dilation on the loss grid (feature-grid dilation = dil / up)
  int top_k;
  float w_src_pos, w_src_neg, w_src_pos_std, w_src_neg_std, w_sim_pos, w_sim_neg;
  // workspace written by the prep kernel, read by the statistics / backward kernels
  float* dm;                  // (2, B, 5, fh, fw)  dots summed over the channel splits
  float* invn;                // (2, B, fh, fw)     1 / max(|x|, 1e-8)
  float* prob;                // (B, C, gh, gw)     softmax of the resampled logits
  uint8_t* lab;               // (B, gh, gw)        resampled source label (0..255)
  uint8_t* flags;             // (B, gh, gw)        bit0 = gt != 255, bit1 = target pixel (mix mask == 0)
};

__device__ __forceinline__ int nearest_src(int dst, float scale, int in) {
  const int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}

__host__ __device__ inline size_t ws_floats(int B, int C, int fh, int fw, int up) {
  const size_t fplane = (size_t)fh * fw, gplane = fplane * up * up;
  return (size_t)2 * B * 5 * fplane + (size_t)2 * B * fplane + (size_t)B * C * gplane;
}

// ---- prep: everything that is per-pixel (not per-neighbourhood), computed once --------
__global__ void __launch_bounds__(kLossThreads)
pfgst_loss_prep_kernel(const LossParams P) {
  const int64_t fplane = (int64_t)P.fh * P.fw, gplane = (int64_t)P.gh * P.gw;
  const int64_t nfeat = (int64_t)2 * P.B * fplane, nloss = (int64_t)P.B * gplane;
  const int64_t i = (int64_t)blockIdx.x * kLossThreads + threadIdx.x;
  if (i < nfeat) {
    // (tensor t, image b, pixel r): merge the channel-split partial maps in a fixed order
    const int64_t tb = i / fplane, r = i - tb * fplane;
    const int64_t split_stride = (int64_t)2 * P.B * 5 * fplane;
    float v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float* src = P.dots + (tb * 5 + k) * fplane + r;
      float a = 0.f;
      for (int sp = 0; sp < P.ksplit; ++sp) a += src[sp * split_stride];
      v[k] = a;
      P.dm[(tb * 5 + k) * fplane + r] = a;
    }
    P.invn[i] = 1.f / fmaxf(sqrtf(v[0]), 1e-8f);
  }
  if (i < nloss) {
    const int b = (int)(i / gplane);
    const int r = (int)(i - (int64_t)b * gplane);
    const int y = r / P.gw, x = r - y * P.gw;
    const int sy = nearest_src(y, P.gscale_h, P.gt_h), sx = nearest_src(x, P.gscale_w, P.gt_w);
    const int64_t go = ((int64_t)b * P.gt_h + sy) * P.gt_w + sx;
    const int64_t g = P.gt[go];
    P.lab[i] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
    P.flags[i] = (uint8_t)((g != 255 ? 1 : 0) | (P.mix[go] <= 0 ? 2 : 0));   // (1 - mix) > 0.5
    // softmax of the nearest-resampled logits (pfgst_loss.py:57, 145)
    const int ly = nearest_src(y, P.lscale_h, P.lh), lx = nearest_src(x, P.lscale_w, P.lw);
    const float* z = P.logits + ((int64_t)b * P.C * P.lh + ly) * P.lw + lx;
    const int64_t lplane = (int64_t)P.lh * P.lw;
    float m = -INFINITY;
    for (int c = 0; c < P.C; ++c) m = fmaxf(m, z[c * lplane]);
    float s = 0.f;
    for (int c = 0; c < P.C; ++c) s += expf(z[c * lplane] - m);
    const float inv = 1.f / s;
    float* po = P.prob + (int64_t)b * P.C * gplane + r;
    for (int c = 0; c < P.C; ++c) po[c * gplane] = expf(z[c * lplane] - m) * inv;
  }
}

// Everything one loss-grid pixel needs from its 3x3 dilated neighbourhood.
struct PixelNb {
  bool inb[9];        // tap inside the loss grid
  float s_ema[9];     // cos(x_ema[n], x_ema[n+delta_k]), 0 outside
  float s_src[9];
  float inv_n_src;    // 1 / max(|x_src[n]|, eps)
  float inv_m_src[9]; // 1 / max(|x_src[n+delta_k]|, eps)
  bool valid_src;     // gt != 255
  bool pos_pair[9];   // unfold(gt)[k] == gt  (zero padding reads as class 0)
  bool nb_valid[9];   // valid_src of the in-bounds neighbour
  bool in_mk;         // valid_src && eroded target mask
  bool eroded;
};

__device__ __forceinline__ void load_pixel(const LossParams& P, int b, int y, int x, PixelNb& o) {
  const int64_t fplane = (int64_t)P.fh * P.fw, gplane = (int64_t)P.gh * P.gw;
  const int fy = y / P.up, fx = x / P.up, fd = P.dil / P.up;
  const int64_t fn = (int64_t)fy * P.fw + fx;
  const float* dme = P.dm + ((int64_t)(0 * P.B + b) * 5) * fplane;
  const float* dms = P.dm + ((int64_t)(1 * P.B + b) * 5) * fplane;
  const float* ine = P.invn + (int64_t)(0 * P.B + b) * fplane;
  const float* ins = P.invn + (int64_t)(1 * P.B + b) * fplane;
  const uint8_t* lab = P.lab + (int64_t)b * gplane;
  const uint8_t* flg = P.flags + (int64_t)b * gplane;
  const float inv_ne = ine[fn], inv_ns = ins[fn];
  o.inv_n_src = inv_ns;
  const int g0 = lab[(int64_t)y * P.gw + x];
  o.valid_src = (flg[(int64_t)y * P.gw + x] & 1) != 0;
  bool er = true;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int oy = (k / 3 - 1), ox = (k % 3 - 1);
    const int yy = y + oy * P.dil, xx = x + ox * P.dil;
    const bool in = yy >= 0 && yy < P.gh && xx >= 0 && xx < P.gw;
    o.inb[k] = in;
    float se = 0.f, ss = 0.f, invm = 0.f;
    int gk = 0;        // zero padding of unfold(gt.float())
    bool trg = false, nbv = false;
    if (in) {
      const int64_t fm = (int64_t)(fy + oy * fd) * P.fw + (fx + ox * fd);
      float de, ds;
      if (k == 4)      { de = dme[fn]; ds = dms[fn]; }
      // forward taps (k > 4) are stored at n, backward taps at the neighbour (symmetry)
      else if (k > 4)  { de = dme[(k - 4) * fplane + fn]; ds = dms[(k - 4) * fplane + fn]; }
      else             { de = dme[(4 - k) * fplane + fm]; ds = dms[(4 - k) * fplane + fm]; }
      invm = ins[fm];
      se = de * (inv_ne * ine[fm]);
      ss = ds * (inv_ns * invm);
      const int64_t gm = (int64_t)yy * P.gw + xx;
      gk = lab[gm];
      const unsigned f = flg[gm];
      nbv = (f & 1u) != 0;
      trg = (f & 2u) != 0;
    }
    o.s_ema[k] = se;
    o.s_src[k] = ss;
    o.inv_m_src[k] = invm;
    o.pos_pair[k] = gk == g0;
    o.nb_valid[k] = nbv;
    er = er && in && trg;
  }
  o.eroded = er;
  o.in_mk = er && o.valid_src;
}

// softmax probabilities of loss-grid pixel (y,x), from the prep map
__device__ __forceinline__ void softmax_at(const LossParams& P, int b, int y, int x, float* p) {
  const int64_t gplane = (int64_t)P.gh * P.gw;
  const float* src = P.prob + (int64_t)b * P.C * gplane + (int64_t)y * P.gw + x;
  for (int c = 0; c < P.C; ++c) p[c] = src[c * gplane];
}

// rank of tap k among the nine similarities: number of taps strictly "before" it
// in descending order, ties broken by the lower index.
__device__ __forceinline__ void tap_ranks(const float (&s)[9], int (&rank_desc)[9], int (&rank_asc)[9]) {
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    int rd = 0, ra = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      if (j == k) continue;
      rd += (s[j] > s[k] || (s[j] == s[k] && j < k)) ? 1 : 0;
      ra += (s[j] < s[k] || (s[j] == s[k] && j < k)) ? 1 : 0;
    }
    rank_desc[k] = rd;
    rank_asc[k] = ra;
  }
}

// stats layout (fp64): 0 n_pos, 1 sum_pos, 2 sumsq_pos, 3 n_neg, 4 sum_neg, 5 sumsq_neg,
//                      6 |Mk|, 7 sum loc_pos, 8 sum loc_neg
constexpr int kNumStats = 9;

__device__ __forceinline__ void finalize_losses(const LossParams& P, const double* st, float* losses) {
  const double n_pos = st[0], n_neg = st[3], mk = st[6];
  const double mean_pos = st[1] / n_pos, mean_neg = st[4] / n_neg;
  const double var_pos = (st[2] - n_pos * mean_pos * mean_pos) / (n_pos - 1.0);
  const double var_neg = (st[5] - n_neg * mean_neg * mean_neg) / (n_neg - 1.0);
  losses[0] = (float)(-mean_pos) * P.w_src_pos;
  losses[1] = (float)(mean_neg) * P.w_src_neg;
  losses[2] = (float)sqrt(var_pos > 0.0 || var_pos != var_pos ? var_pos : 0.0) * P.w_src_pos_std;
  losses[3] = (float)sqrt(var_neg > 0.0 || var_neg != var_neg ? var_neg : 0.0) * P.w_src_neg_std;
  const bool any = mk > 1.0;   // pfgst_loss.py:227  `if ignore_mask.sum() > 1`
  losses[4] = any ? (float)(st[7] / (mk * (double)(P.top_k + 1))) * P.w_sim_pos : 0.f;
  losses[5] = any ? (float)(st[8] / (mk * (double)P.top_k)) * P.w_sim_neg : 0.f;
}

__global__ void __launch_bounds__(kLossThreads)
pfgst_loss_fwd_kernel(const LossParams P, double* __restrict__ stats, float* __restrict__ losses,
                      float* __restrict__ density, uint8_t* __restrict__ eroded_out,
                      unsigned* __restrict__ done_counter) {
  __shared__ double red[kNumStats][kLossThreads / 32];
  const int64_t plane = (int64_t)P.gh * P.gw;
  const int64_t total = (int64_t)P.B * plane;
  double acc[kNumStats];
#pragma unroll
  for (int i = 0; i < kNumStats; ++i) acc[i] = 0.0;

  for (int64_t n = (int64_t)blockIdx.x * kLossThreads + threadIdx.x; n < total;
       n += (int64_t)gridDim.x * kLossThreads) {
    const int b = (int)(n / plane);
    const int r = (int)(n - (int64_t)b * plane);
    const int y = r / P.gw, x = r - y * P.gw;
    PixelNb px;
    load_pixel(P, b, y, x, px);

    // L4: source pair statistics (pfgst_loss.py:85-113)
    if (px.valid_src) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double s = (double)px.s_src[k];
        if (px.pos_pair[k]) { acc[0] += 1.0; acc[1] += s; acc[2] += s * s; }
        else                { acc[3] += 1.0; acc[4] += s; acc[5] += s * s; }
      }
    }
    float mean_e = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) mean_e += px.s_ema[k];
    if (density) density[n] = 1.f - mean_e / 9.f;
    if (eroded_out) eroded_out[n] = px.eroded ? 1 : 0;

    // L3 + L5: target consistency terms on Mk
    if (px.in_mk) {
      float p[kMaxC], q[kMaxC];
      softmax_at(P, b, y, x, p);
      int rd[9], ra[9];
      tap_ranks(px.s_ema, rd, ra);
      float lp = 0.f, ln = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const bool top = rd[k] < P.top_k + 1, bot = ra[k] < P.top_k;
        if (!top && !bot) continue;
        // every tap of an Mk pixel is in bounds (eroded mask)
        softmax_at(P, b, y + (k / 3 - 1) * P.dil, x + (k % 3 - 1) * P.dil, q);
        float cp = 0.f;
        for (int c = 0; c < P.C; ++c) cp = fmaf(p[c], q[c], cp);
        if (top) lp += px.s_ema[k] * (-cp);
        if (bot) ln += (1.f - px.s_ema[k]) * (-(1.f - cp));
      }
      acc[6] += 1.0;
      acc[7] += (double)lp;
      acc[8] += (double)ln;
    }
  }

  // block reduction -> one fp64 atomic per statistic per block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kNumStats; ++i) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < kNumStats) {
    double v = 0.0;
    for (int wv = 0; wv < kLossThreads / 32; ++wv) v += red[threadIdx.x][wv];
    if (v != 0.0) atomicAdd(&stats[threadIdx.x], v);
  }
  // last block finalises the six losses on the device (no host round trip)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double st[kNumStats];
    for (int i = 0; i < kNumStats; ++i) st[i] = *((volatile double*)&stats[i]);
    finalize_losses(P, st, losses);
  }
}

// One thread per FEATURE-grid pixel (loops over its up x up loss-grid pixels).
__global__ void __launch_bounds__(kLossThreads)
pfgst_loss_bwd_kernel(const LossParams P, const double* __restrict__ stats, const float* __restrict__ gout,
                      float* __restrict__ coef, float* __restrict__ grad_logits) {
  const int64_t fplane = (int64_t)P.fh * P.fw;
  const int64_t total = (int64_t)P.B * fplane;
  const int64_t n = (int64_t)blockIdx.x * kLossThreads + threadIdx.x;
  if (n >= total) return;
  const int b = (int)(n / fplane);
  const int r = (int)(n - (int64_t)b * fplane);
  const int fy = r / P.fw, fx = r - fy * P.fw;

  const double n_pos = stats[0], n_neg = stats[3], mk = stats[6];
  const double mean_pos = stats[1] / n_pos, mean_neg = stats[4] / n_neg;
  const double std_pos = sqrt(fmax((stats[2] - n_pos * mean_pos * mean_pos) / (n_pos - 1.0), 0.0));
  const double std_neg = sqrt(fmax((stats[5] - n_neg * mean_neg * mean_neg) / (n_neg - 1.0), 0.0));
  // d loss / d S for a positive / negative source pair:  a + c * (S - mean)
  const float a_pos = (float)(-(double)gout[0] * P.w_src_pos / n_pos);
  const float c_pos = (float)((double)gout[2] * P.w_src_pos_std / ((n_pos - 1.0) * std_pos));
  const float a_neg = (float)((double)gout[1] * P.w_src_neg / n_neg);
  const float c_neg = (float)((double)gout[3] * P.w_src_neg_std / ((n_neg - 1.0) * std_neg));
  const float fmean_pos = (float)mean_pos, fmean_neg = (float)mean_neg;
  const bool any = mk > 1.0;
  const float g_pos = any ? (float)((double)gout[4] * P.w_sim_pos / (mk * (double)(P.top_k + 1))) : 0.f;
  const float g_neg = any ? (float)((double)gout[5] * P.w_sim_neg / (mk * (double)P.top_k)) : 0.f;

  float cf[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) cf[k] = 0.f;

  for (int uy = 0; uy < P.up; ++uy)
    for (int ux = 0; ux < P.up; ++ux) {
      const int y = fy * P.up + uy, x = fx * P.up + ux;
      PixelNb px;
      load_pixel(P, b, y, x, px);
      // --- x_src: gather-form coefficients ---
      float bsum = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (k == 4 || !px.inb[k]) continue;
        const float S = px.s_src[k];
        const float g = px.pos_pair[k] ? a_pos + c_pos * (S - fmean_pos) : a_neg + c_neg * (S - fmean_neg);
        // the pair (n, m) is counted once from n (if n is valid) and once from m (if m is valid)
        const float W = g * ((px.valid_src ? 1.f : 0.f) + (px.nb_valid[k] ? 1.f : 0.f));
        cf[k] += W * px.inv_n_src * px.inv_m_src[k];
        bsum += W * S;
      }
      cf[4] -= bsum * px.inv_n_src * px.inv_n_src;

      // --- logits_trg: through p only (q detached) ---
      if (grad_logits && px.in_mk && any) {
        float p[kMaxC], q[kMaxC], dp[kMaxC];
        softmax_at(P, b, y, x, p);
        for (int c = 0; c < P.C; ++c) dp[c] = 0.f;
        int rd[9], ra[9];
        tap_ranks(px.s_ema, rd, ra);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const bool top = rd[k] < P.top_k + 1, bot = ra[k] < P.top_k;
          if (!top && !bot) continue;
          // d/dcp of  -S*cp  and of  -(1-S)*(1-cp)
          const float dcp = (top ? -px.s_ema[k] * g_pos : 0.f) + (bot ? (1.f - px.s_ema[k]) * g_neg : 0.f);
          softmax_at(P, b, y + (k / 3 - 1) * P.dil, x + (k % 3 - 1) * P.dil, q);
          for (int c = 0; c < P.C; ++c) dp[c] = fmaf(dcp, q[c], dp[c]);
        }
        float dot = 0.f;
        for (int c = 0; c < P.C; ++c) dot = fmaf(p[c], dp[c], dot);
        const int sy = nearest_src(y, P.lscale_h, P.lh), sx = nearest_src(x, P.lscale_w, P.lw);
        float* gz = grad_logits + ((int64_t)b * P.C * P.lh + sy) * P.lw + sx;
        const int64_t lplane = (int64_t)P.lh * P.lw;
        // several loss pixels can map to one logit only when the logits are UP-sampled
        // (lscale < 1); accumulate then, plain store otherwise
        const bool shared_src = P.lscale_h < 1.f || P.lscale_w < 1.f;
        for (int c = 0; c < P.C; ++c) {
          const float v = p[c] * (dp[c] - dot);
          if (shared_src) atomicAdd(gz + c * lplane, v); else gz[c * lplane] = v;
        }
      }
    }
#pragma unroll
  for (int k = 0; k < 9; ++k) coef[((int64_t)b * 9 + k) * fplane + r] = cf[k];
}

static int fill_params(LossParams& P, const float* dots, int ksplit, int64_t B, int fh, int fw, int up,
                       const float* logits, int C, int lh, int lw, float lsh, float lsw, const int64_t* gt,
                       const int64_t* mix, int gt_h, int gt_w, int dil, int top_k, const float* w6, void* ws) {
  if (!dots || !logits || !gt || !mix || !w6 || !ws) return PFST_ERR_INVALID_ARG;
  if (B < 0 || B > 0x7fffffff || fh < 1 || fw < 1 || up < 1 || C < 1 || lh < 1 || lw < 1 || gt_h < 1 || gt_w < 1)
    return PFST_ERR_INVALID_ARG;
  if (C > kMaxC) return PFST_ERR_UNSUPPORTED;
  if (dil < 1 || dil % up != 0 || top_k < 1 || top_k > 4 || ksplit < 1) return PFST_ERR_INVALID_ARG;
  P.dots = dots; P.ksplit = ksplit; P.B = (int)B; P.fh = fh; P.fw = fw; P.up = up;
  P.logits = logits; P.C = C; P.lh = lh; P.lw = lw; P.lscale_h = lsh; P.lscale_w = lsw;
  P.gt = gt; P.mix = mix; P.gt_h = gt_h; P.gt_w = gt_w;
  P.gh = fh * up; P.gw = fw * up;
  // F.interpolate(size=...) nearest: scale = (float)in / out
  P.gscale_h = (float)gt_h / (float)P.gh; P.gscale_w = (float)gt_w / (float)P.gw;
  P.dil = dil; P.top_k = top_k;
  P.w_src_pos = w6[0]; P.w_src_neg = w6[1]; P.w_src_pos_std = w6[2]; P.w_src_neg_std = w6[3];
  P.w_sim_pos = w6[4]; P.w_sim_neg = w6[5];
  const size_t fplane = (size_t)fh * fw, gplane = (size_t)P.gh * P.gw;
  P.dm = static_cast<float*>(ws);
  P.invn = P.dm + (size_t)2 * B * 5 * fplane;
  P.prob = P.invn + (size_t)2 * B * fplane;
  P.lab = reinterpret_cast<uint8_t*>(P.prob + (size_t)B * C * gplane);
  P.flags = P.lab + (size_t)B * gplane;
  return PFST_OK;
}

}  // namespace pfst

extern "C" {

int64_t pfst_pfgst_loss_ws_bytes(int64_t B, int32_t C, int32_t fh, int32_t fw, int32_t up) {
  if (B < 0 || C < 1 || fh < 1 || fw < 1 || up < 1) return 0;
  const size_t gplane = (size_t)fh * fw * up * up;
  return (int64_t)(pfst::ws_floats((int)B, C, fh, fw, up) * sizeof(float) + 2 * (size_t)B * gplane + 16);
}

int pfst_pfgst_loss_fwd(const float* dots, int32_t ksplit, int64_t B, int32_t fh, int32_t fw, int32_t up,
                        const float* logits, int32_t C, int32_t lh, int32_t lw, float lscale_h,
                        float lscale_w, const int64_t* gt, const int64_t* mix, int32_t gt_h, int32_t gt_w,
                        int32_t dilation, int32_t top_k, const float* weights6_host, void* workspace,
                        double* stats, float* losses, float* density, uint8_t* eroded, void* stream) {
  pfst::LossParams P;
  const int rc = pfst::fill_params(P, dots, ksplit, B, fh, fw, up, logits, C, lh, lw, lscale_h, lscale_w, gt,
                                   mix, gt_h, gt_w, dilation, top_k, weights6_host, workspace);
  if (rc != PFST_OK) return rc;
  if (!stats || !losses) return PFST_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // stats[0..8] fp64 sums, stats[15] doubles as the block-completion counter
  PFST_CUDA_TRY(cudaMemsetAsync(stats, 0, 16 * sizeof(double), s), "pfst_pfgst_loss_fwd/memset");
  const int64_t total = (int64_t)P.B * P.gh * P.gw;
  {
    const int64_t nfeat = (int64_t)2 * P.B * fh * fw;
    const int64_t n = nfeat > total ? nfeat : total;
    if (n == 0) return PFST_OK;
    pfst::pfgst_loss_prep_kernel<<<(unsigned)((n + pfst::kLossThreads - 1) / pfst::kLossThreads),
                                   pfst::kLossThreads, 0, s>>>(P);
    PFST_CHECK_LAUNCH("pfst_pfgst_loss_fwd/prep");
  }
  int64_t grid = (total + pfst::kLossThreads - 1) / pfst::kLossThreads;
  const int64_t cap = (int64_t)pfst::kNumSMs * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  pfst::pfgst_loss_fwd_kernel<<<(unsigned)grid, pfst::kLossThreads, 0, s>>>(
      P, stats, losses, density, eroded, reinterpret_cast<unsigned*>(stats + 15));
  PFST_CHECK_LAUNCH("pfst_pfgst_loss_fwd");
  return PFST_OK;
}

int pfst_pfgst_loss_bwd(const float* dots, int32_t ksplit, int64_t B, int32_t fh, int32_t fw, int32_t up,
                        const float* logits, int32_t C, int32_t lh, int32_t lw, float lscale_h,
                        float lscale_w, const int64_t* gt, const int64_t* mix, int32_t gt_h, int32_t gt_w,
                        int32_t dilation, int32_t top_k, const float* weights6_host, const void* workspace,
                        const double* stats, const float* grad_losses, float* coef, float* grad_logits,
                        void* stream) {
  pfst::LossParams P;
  const int rc = pfst::fill_params(P, dots, ksplit, B, fh, fw, up, logits, C, lh, lw, lscale_h, lscale_w, gt,
                                   mix, gt_h, gt_w, dilation, top_k, weights6_host, const_cast<void*>(workspace));
  if (rc != PFST_OK) return rc;
  if (!stats || !grad_losses || !coef) return PFST_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (grad_logits)
    PFST_CUDA_TRY(cudaMemsetAsync(grad_logits, 0, sizeof(float) * (size_t)P.B * C * lh * lw, s),
                  "pfst_pfgst_loss_bwd/memset");
  const int64_t total = (int64_t)P.B * fh * fw;
  if (total == 0) return PFST_OK;
  const int64_t grid = (total + pfst::kLossThreads - 1) / pfst::kLossThreads;
  pfst::pfgst_loss_bwd_kernel<<<(unsigned)grid, pfst::kLossThreads, 0, s>>>(P, stats, grad_losses, coef,
                                                                           grad_logits);
  PFST_CHECK_LAUNCH("pfst_pfgst_loss_bwd");
  return PFST_OK;
}

}  // extern "C"
