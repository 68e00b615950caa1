// Bisect probe for the TMA tile pipeline (development tool, not part of the library).
// usage: tma_probe <variant>   variant bits: 1 = prefetch desc, 2 = producer lanes exit early,
//                                            4 = map inside struct array (dynamic index)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pfst_b200/csrc/tma.cuh"
using namespace pfst;

struct Maps { CUtensorMap m[2]; };
constexpr int BW = 36, BH = 18, BC = 8, STAGES = 2;

template <int VAR>
__global__ void __launch_bounds__(160) probe(const __grid_constant__ Maps maps, const __grid_constant__ CUtensorMap single,
                                             int t, int chunks, float* out, int mode, unsigned bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* buf = reinterpret_cast<float*>(smem);
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 4); }
    mbar_fence_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CUtensorMap* map = (VAR & 4) ? &maps.m[t] : &single;
  constexpr int STAGE_FLOATS = BW * BH * BC;
  if (warp == 4) {
    if (lane == 0) {
      if (VAR & 1) tma_prefetch_desc(map);
      for (int it = 0; it < chunks; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        if (it >= STAGES) mbar_wait(&empty_bar[s], ph ^ 1);
        if (mode == 1) { mbar_arrive(&full_bar[s]); }
        else {
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          tma_load_4d(buf + s * STAGE_FLOATS, map, &full_bar[s], mode == 2 ? 0 : -2, 0, it * BC, 0);
        }
      }
    }
    if (VAR & 2) return;
  } else {
    float acc = 0.f;
    for (int it = 0; it < chunks; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      for (int i = threadIdx.x; i < STAGE_FLOATS; i += 128) acc += buf[s * STAGE_FLOATS + i];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    atomicAdd(out, acc);
  }
}

int main(int argc, char** argv) {
  const int var = argc > 1 ? atoi(argv[1]) : 0;
  const int mode = argc > 2 ? atoi(argv[2]) : 0;
  const int bw = argc > 3 ? atoi(argv[3]) : BW, bh = argc > 4 ? atoi(argv[4]) : BH;
  const int B = 2, D = 32, H = 16, W = 16;
  std::vector<float> h((size_t)B * D * H * W, 1.0f);
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 4);
  cudaMemset(out, 0, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  Maps maps;
  bool ok = make_nchw_tensor_map(&maps.m[0], d, B, D, H, W, bw, bh, BC);
  maps.m[1] = maps.m[0];
  printf("variant %d encode ok=%d\n", var, (int)ok);
  const size_t smem = STAGES * BW * BH * BC * 4;
#define RUN(V) { cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                 probe<V><<<1, 160, smem>>>(maps, maps.m[0], 1, D / BC, out, mode, (unsigned)(bw * bh * BC * 4)); }
  switch (var) { case 0: RUN(0) break; case 1: RUN(1) break; case 2: RUN(2) break; case 3: RUN(3) break;
                 case 4: RUN(4) break; case 5: RUN(5) break; case 6: RUN(6) break; default: RUN(7) break; }
  cudaError_t e = cudaDeviceSynchronize();
  float r = 0;
  cudaMemcpy(&r, out, 4, cudaMemcpyDeviceToHost);
  // expected: in-image elements of the box: rows 0..15 (of 18), cols 2..17 -> 16 of 36, per channel; 32 channels
  printf("mode %d box %dx%d variant %d: %s  sum=%.1f (expect %.1f)\n", mode, bw, bh, var, cudaGetErrorString(e), r, 16.0 * 16 * 32);
  return e != cudaSuccess;
}
