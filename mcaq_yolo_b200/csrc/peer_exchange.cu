// C ABI of the peer-memory range exchange: buffer management (cudaMalloc + CUDA IPC so that each
// process can map the buffers of the other ranks of the node) and a stand-alone merge kernel.
#include "peer_exchange.cuh"

namespace mcaq {

static long long g_timeout_ns = 2000000000LL;
long long xchg_timeout_ns() { return g_timeout_ns; }

// on a timeout `packed` is left untouched (callers pre-fill it with this rank's own ranges when they
// want a defined fallback) and the error word carries the step
__global__ void xchg_merge_kernel(const float* local, int world, int C, float* packed, long long budget_ns) {
  const int e = xchg_step(local);
  const bool ok = xchg_wait(local, world, e, threadIdx.x, budget_ns);
  if (!__syncthreads_and(ok)) return;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) packed[i] = xchg_min(local, world, 2 * C, e, i);
}

// single-rank publish (what K2's first CTA does), used to drive the protocol from tests / host code
__global__ void xchg_publish_kernel(XchgPeers px, const float* packed, int C) {
  __shared__ int e_s;
  if (threadIdx.x == 0) {
    int* ep = reinterpret_cast<int*>(px.base[px.rank]);
    const int e = *ep + 1;
    *ep = e;
    e_s = e;
  }
  __syncthreads();
  const int e = e_s;
  for (int p = 0; p < px.world; ++p) {
    float* dst = px.base[p] + XCHG_SLOTS + (long long)((e & 1) * px.world + px.rank) * 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) dst[i] = packed[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < px.world)
    st_release_sys(reinterpret_cast<int*>(px.base[threadIdx.x]) + XCHG_FLAGS + 8 * (e & 1) + px.rank, e);
}

}  // namespace mcaq

using namespace mcaq;

extern "C" long long mcaq_xchg_bytes(int C, int world) {
  if (C <= 0 || world <= 0 || world > XCHG_MAX_RANKS) return MCAQ_EINVAL;
  return (long long)(XCHG_SLOTS + 2LL * world * 2 * C) * 4;
}

extern "C" int mcaq_xchg_alloc(long long bytes, void** out) {
  if (!out || bytes <= 0) return MCAQ_EINVAL;
  cudaError_t e = cudaMalloc(out, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*out, 0, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaDeviceSynchronize();
}

extern "C" int mcaq_xchg_free(void* p) { return p ? (int)cudaFree(p) : 0; }

extern "C" int mcaq_xchg_export(void* p, void* handle64) {
  if (!p || !handle64) return MCAQ_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p);
}

extern "C" int mcaq_xchg_open(const void* handle64, void** out) {
  if (!handle64 || !out) return MCAQ_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" void mcaq_xchg_set_timeout_ms(int ms) { g_timeout_ns = ms > 0 ? (long long)ms * 1000000LL : 2000000000LL; }

// Synchronising read of the buffer's error word: *step = 0 when every wait so far was satisfied, else the
// first exchange step whose wait timed out (that launch kept this rank's own ranges).  Clears the word.
extern "C" int mcaq_xchg_error(void* local, int* step) {
  if (!local || !step) return MCAQ_EINVAL;
  int* w = reinterpret_cast<int*>(local) + XCHG_ERROR;
  cudaError_t e = cudaMemcpy(step, w, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return (int)e;
  if (*step != 0) e = cudaMemset(w, 0, sizeof(int));
  return (int)e;
}

extern "C" int mcaq_xchg_close(void* p) { return p ? (int)cudaIpcCloseMemHandle(p) : 0; }

static int fill_peers(XchgPeers& px, void* const* peers, int rank, int world) {
  if (!peers || world <= 0 || world > XCHG_MAX_RANKS || rank < 0 || rank >= world) return MCAQ_EINVAL;
  for (int i = 0; i < XCHG_MAX_RANKS; ++i) px.base[i] = i < world ? reinterpret_cast<float*>(peers[i]) : nullptr;
  for (int i = 0; i < world; ++i)
    if (!px.base[i]) return MCAQ_EINVAL;
  px.rank = rank;
  px.world = world;
  px.timeout_ns = g_timeout_ns;
  return 0;
}

extern "C" int mcaq_xchg_publish(void* const* peers, int rank, int world, const float* packed, int C, void* stream) {
  XchgPeers px;
  int rc = fill_peers(px, peers, rank, world);
  if (rc) return rc;
  if (!packed || C <= 0) return MCAQ_EINVAL;
  xchg_publish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(px, packed, C);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_xchg_merge(const void* local, int world, int C, float* packed, void* stream) {
  if (!local || !packed || C <= 0 || world <= 0 || world > XCHG_MAX_RANKS) return MCAQ_EINVAL;
  xchg_merge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(local), world, C, packed,
                                                         g_timeout_ns);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
