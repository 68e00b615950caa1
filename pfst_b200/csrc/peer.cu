// Peer board (P2, multi-rank): the one cross-image reduction on the step's critical path —
// the packed prototype [sums | counts] buffer, 12-68 KB — done as a ONE-SHOT all-reduce over
// NVLink peer memory inside the prototype-finalise kernel, instead of a library collective
// between two kernels.
//
// Every rank owns a board (cudaMalloc, exported with a CUDA IPC handle, mapped by every
// peer):   inbox[2 parities][N source ranks][stride floats]  |  flags[N source ranks][C] u64
// Block c of rank r (one block per class):
//   1. pushes its chunk (D sums + 1 count of class c) into inbox[parity][r] of EVERY rank
//      (plain remote stores — fire and forget), fences at system scope, then writes the
//      step token into flags[r][c] of every rank (st.release.sys);
//   2. waits until its own flags[0..N)[c] carry the token (ld.acquire.sys; a bounded spin);
//   3. sums the N chunks of its own inbox in fixed rank order 0..N-1 — so the prototypes are
//      bit-identical on all ranks by construction — and finalises class c exactly like
//      proto_finalize_kernel (proto.cu), on the device-resident iteration counter.
// No block depends on another block of any grid being resident, only on stores that peers
// issue before they wait: the exchange cannot deadlock on occupancy. The parity (iteration
// & 1) double-buffers the inbox: a rank can run at most one step ahead of a peer (it needs the
// peer's token of step t to finish step t), so the chunk of step t+1 never overwrites data of
// step t that the peer still has to read, and tokens are monotonic (wait is `>= token`).
#include <string.h>

#include "common.cuh"

namespace pfst {

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerThreads = 128;

// status: int64[2] = {0 | 1 = a wait timed out (results of that step are invalid), last token}
__global__ void __launch_bounds__(kPeerThreads)
proto_finalize_peer_kernel(float* __restrict__ packed, int C, int D, const float* mu_prev,
                           const uint8_t* seen_prev, double alpha, long long* iter_state, float* mu_out,
                           int64_t* __restrict__ cnt_out, uint8_t* seen_out,
                           const unsigned long long* __restrict__ boards, int rank, int nranks,
                           long long stride, long long flag_off_bytes, long long* status,
                           unsigned long long timeout_ns) {
  const int c = blockIdx.x, tid = threadIdx.x;
  const long long it = *reinterpret_cast<volatile long long*>(iter_state);
  float a32 = 0.f, b32 = 1.f;
  if (it > 0) {
    double a = 1.0 - 1.0 / (double)(it + 1);
    if (alpha < a) a = alpha;
    a32 = (float)a;
    b32 = (float)(1.0 - a);
  }
  const unsigned long long token = (unsigned long long)it + 1ull;
  const long long par = it & 1;
  __shared__ unsigned long long s_board[kPeerMaxRanks];
  if (tid < nranks) s_board[tid] = boards[tid];
  __syncthreads();

  // 1. push: my chunk of class c into every rank's inbox[par][rank] (my own board included)
  const long long src_off = (par * nranks + rank) * stride;
  const float my_cnt = packed[(int64_t)C * D + c];
  for (int d = tid; d < D; d += kPeerThreads) {
    const float v = packed[(int64_t)c * D + d];
    for (int q = 0; q < nranks; ++q)
      (reinterpret_cast<float*>(s_board[q]) + src_off)[(int64_t)c * D + d] = v;
  }
  if (tid < nranks) (reinterpret_cast<float*>(s_board[tid]) + src_off)[(int64_t)C * D + c] = my_cnt;
  __threadfence_system();
  __syncthreads();
  if (tid < nranks) {
    unsigned long long* flags =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(s_board[tid]) + flag_off_bytes);
    st_release_sys_u64(flags + (int64_t)rank * C + c, token);
  }

  // 2. wait for the chunk of class c from every rank
  if (tid < nranks) {
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(s_board[rank]) + flag_off_bytes) +
        (int64_t)tid * C + c;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys_u64(mine) < token) {
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
        atomicExch(reinterpret_cast<unsigned long long*>(status), 1ull);
        break;
      }
    }
  }
  __syncthreads();

  // 3. reduce in rank order, then the finalise arithmetic of proto_finalize_kernel
  const float* in0 = reinterpret_cast<const float*>(s_board[rank]) + par * nranks * stride;
  float cnt = 0.f;
  for (int r = 0; r < nranks; ++r) cnt += ld_relaxed_sys_f32(in0 + r * stride + (int64_t)C * D + c);
  const bool has = cnt > 0.f;
  const bool seen = seen_prev ? seen_prev[c] != 0 : false;
  const float denom = fmaxf(cnt, 1.f);
  for (int d = tid; d < D; d += kPeerThreads) {
    float sum = 0.f;
    for (int r = 0; r < nranks; ++r) sum += ld_relaxed_sys_f32(in0 + r * stride + (int64_t)c * D + d);
    const float mean = sum / denom;
    const float prev = mu_prev ? mu_prev[(int64_t)c * D + d] : 0.f;
    float out = prev;
    if (has) out = seen ? __fadd_rn(__fmul_rn(a32, prev), __fmul_rn(b32, mean)) : mean;
    mu_out[(int64_t)c * D + d] = out;
    packed[(int64_t)c * D + d] = 0.f;          // my accumulator is consumed: ready for the next step
  }
  __syncthreads();
  if (tid == 0) {
    if (cnt_out) cnt_out[c] = (int64_t)cnt;
    if (seen_out) seen_out[c] = (has || seen) ? 1 : 0;
    packed[(int64_t)C * D + c] = 0.f;
    __threadfence();
    if (atomicAdd(reinterpret_cast<unsigned long long*>(iter_state + 1), 1ull) == gridDim.x - 1) {
      iter_state[1] = 0;
      iter_state[0] = it + 1;
      status[1] = (long long)token;
    }
  }
}

}  // namespace pfst

extern "C" {

int64_t pfst_peer_board_bytes(int32_t C, int32_t D, int32_t nranks, int64_t* stride_out,
                              int64_t* flag_offset_out) {
  if (C < 1 || D < 1 || nranks < 1 || nranks > pfst::kPeerMaxRanks) return PFST_ERR_INVALID_ARG;
  const int64_t stride = (((int64_t)C * D + C + 3) / 4) * 4;       // floats per (parity, source rank)
  const int64_t flag_off = 2 * (int64_t)nranks * stride * 4;       // 16-byte aligned by construction
  if (stride_out) *stride_out = stride;
  if (flag_offset_out) *flag_offset_out = flag_off;
  return flag_off + (int64_t)nranks * C * 8;
}

int pfst_peer_alloc(int64_t bytes, void** ptr_out, void* ipc_handle64) {
  if (bytes <= 0 || !ptr_out || !ipc_handle64) return PFST_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  PFST_CUDA_TRY(cudaMalloc(&p, (size_t)bytes), "pfst_peer_alloc/cudaMalloc");
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();   // set-up time only: the zeroed flags must be visible to peers
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    pfst::set_last_cuda_error(e, "pfst_peer_alloc/cudaIpcGetMemHandle");
    cudaFree(p);
    cudaGetLastError();
    return PFST_ERR_CUDA;
  }
  memcpy(ipc_handle64, &h, 64);
  *ptr_out = p;
  return PFST_OK;
}

int pfst_peer_open(const void* ipc_handle64, void** ptr_out) {
  if (!ipc_handle64 || !ptr_out) return PFST_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, 64);
  void* p = nullptr;
  PFST_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "pfst_peer_open/cudaIpcOpenMemHandle");
  *ptr_out = p;
  return PFST_OK;
}

int pfst_peer_close(void* ptr) {
  if (!ptr) return PFST_ERR_INVALID_ARG;
  PFST_CUDA_TRY(cudaIpcCloseMemHandle(ptr), "pfst_peer_close/cudaIpcCloseMemHandle");
  return PFST_OK;
}

int pfst_peer_free(void* ptr) {
  if (!ptr) return PFST_ERR_INVALID_ARG;
  PFST_CUDA_TRY(cudaFree(ptr), "pfst_peer_free/cudaFree");
  return PFST_OK;
}

int pfst_proto_finalize_peer(float* packed, int32_t C, int32_t D, const float* mu_prev,
                             const uint8_t* seen_prev, double alpha, int64_t* iter_state, float* mu_out,
                             int64_t* cnt_out, uint8_t* seen_out, const uint64_t* boards, int32_t rank,
                             int32_t nranks, int64_t* status, int64_t timeout_ns, void* stream) {
  if (!packed || !mu_out || !iter_state || !boards || !status || C < 1 || D < 1 || nranks < 1 ||
      rank < 0 || rank >= nranks || timeout_ns <= 0)
    return PFST_ERR_INVALID_ARG;
  if (nranks > pfst::kPeerMaxRanks) return PFST_ERR_UNSUPPORTED;
  int64_t stride = 0, flag_off = 0;
  pfst_peer_board_bytes(C, D, nranks, &stride, &flag_off);
  pfst::proto_finalize_peer_kernel<<<(unsigned)C, pfst::kPeerThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      packed, C, D, mu_prev, seen_prev, alpha, reinterpret_cast<long long*>(iter_state), mu_out, cnt_out, seen_out,
      reinterpret_cast<const unsigned long long*>(boards), rank, nranks, stride, flag_off,
      reinterpret_cast<long long*>(status), (unsigned long long)timeout_ns);
  PFST_CHECK_LAUNCH("pfst_proto_finalize_peer");
  return PFST_OK;
}

}  // extern "C"
