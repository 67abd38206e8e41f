// Per-channel range exchange between the GPUs of one node over peer memory (NVLink / NVSwitch).
//
// The batch is sharded over R ranks (one process per GPU); the only value the inference path has
// to merge is the per-channel [min, -max] vector (2C floats).  Instead of a collective
// launch between K2 and K3, the exchange is folded into the two kernels:
//   * K2's first CTA stores this rank's vector into slot `rank` of EVERY rank's exchange buffer
//     (plain stores through the peer mapping), fences, then publishes the step number in every
//     rank's flag word with a system-scope release store -- at the START of the kernel;
//   * the same CTA, at the END of the kernel (≈ 100 us later, when the peers' vectors have long
//     arrived), acquires the R flag words of its LOCAL buffer (spins until all carry the current
//     step), takes the minimum over the R slots and writes it over the `packed` vector K3 reads.
// Exactly one CTA per launch can ever wait, and only after its own work: CTAs of a wide kernel
// spinning on a peer could occupy the SMs the peer-facing publisher of another in-flight step needs
// (a cross-GPU resource deadlock when several steps are in flight).
// Slots and flags are double buffered by step parity: a rank can run at most one step ahead of a
// peer (its next merge waits for the peer's next publish), so a slot is never overwritten while it
// can still be read.  No host involvement, no extra launch, capturable in a CUDA graph.
//
// A wait is BOUNDED: a peer that died, or ranks that issued different numbers of launches (uneven last
// batch, rank-0-only validation), would otherwise leave the waiting CTA spinning until the process is
// killed.  After `timeout_ns` (default 2 s, mcaq_xchg_set_timeout_ms) the waiter records the step in the
// buffer's error word, the launch keeps THIS rank's own ranges (no merge) and finishes; the host reads the
// error word with mcaq_xchg_error (peer.RangeExchange.check raises).  Every rank must run the same
// sequence of exchange steps.
//
// Buffer layout (4-byte words):  [0] step counter of the owning rank   [1] error word (0, or the first
// step whose wait timed out)   [16 + 8*par + q] flag of
// rank q   [32 + ((par*R + q) * 2C) ...] slot of rank q, par = step & 1.
#pragma once
#include "common.cuh"

namespace mcaq {

constexpr int XCHG_MAX_RANKS = 8;
constexpr int XCHG_FLAGS = 16;
constexpr int XCHG_SLOTS = 32;
constexpr int XCHG_ERROR = 1;

struct XchgPeers {
  float* base[XCHG_MAX_RANKS];   // exchange buffer of every rank (own entry = local memory)
  int rank, world;
  long long timeout_ns;          // budget of one wait (<= 0: library default)
};

long long xchg_timeout_ns();     // host: current budget (peer_exchange.cu)

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// current step of the local buffer (written by this rank's K2 earlier in stream order)
__device__ __forceinline__ int xchg_step(const float* local) {
  return *reinterpret_cast<const volatile int*>(local);
}

// thread q < world spins until rank q has published step e in the local buffer, for at most budget_ns;
// returns false (and records e in the error word) on a timeout.  Callers combine the per-thread results
// with __syncthreads_and and skip the merge when any wait failed.
__device__ __forceinline__ bool xchg_wait(const float* local, int world, int e, int q, long long budget_ns) {
  bool ok = true;
  if (q < world) {
    const int* f = reinterpret_cast<const int*>(local) + XCHG_FLAGS + 8 * (e & 1) + q;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) != e) {
      __nanosleep(64);
      if ((long long)(global_ns() - t0) > budget_ns) { ok = false; break; }
    }
    if (!ok) atomicCAS(reinterpret_cast<int*>(const_cast<float*>(local)) + XCHG_ERROR, 0, e);
  }
  return ok;
}

// min over ranks of element idx (< 2C) of the step-e slots of the local buffer
__device__ __forceinline__ float xchg_min(const float* local, int world, int C2, int e, int idx) {
  const volatile float* s = local + XCHG_SLOTS + (long long)((e & 1) * world) * C2 + idx;
  float m = s[0];
  for (int q = 1; q < world; ++q) m = fminf(m, s[(long long)q * C2]);
  return m;
}

}  // namespace mcaq
