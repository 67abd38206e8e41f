// Geometry and per-tile context shared by the K3 kernels (tile_quantize.cu, tile_quantize_train.cu).
#pragma once
#include "common.cuh"

namespace mcaq {

struct QGeom {
  int B, C, H, W, HW, Ht, Wt;
  float sy, sx;              // (float)Ht/H, (float)Wt/W
  long long nvec_total;      // B * HW / VEC
  int nvec;                  // HW / VEC
};

__device__ __forceinline__ void bit_limits(int bidx, float& qmin, float& qmax) {
  const int half = 1 << (bidx + 1);            // 2^(bits-1)
  qmin = -(float)half;
  qmax = (float)(half - 1);
}

// 128-thread CTAs, six per SM (80 registers): the same 768 resident threads as three CTAs of 256, but the finer grain
// packs around the resident morphology CTAs (16 K registers each) -- whole step 0.0820 -> 0.0803 ms
// (profiles/r02_k3_variants.txt)
#ifndef K3_THREADS
#define K3_THREADS 128
#endif
constexpr int QV_THREADS = K3_THREADS;
constexpr int QV_CHUNK = 16;
#ifndef K3_UNROLL
#define K3_UNROLL 8
#endif
constexpr int QV_UNROLL = K3_UNROLL;
constexpr int QV_ROW = QV_CHUNK + 1;


struct FracCtx {
  int lo_idx, hi_idx;     // table rows of floor(b) and min(floor(b)+1, 8)
  float f, omf;           // frac and (1 - frac)
};

__device__ __forceinline__ FracCtx frac_ctx(float bits) {
  FracCtx fc;
  const float bf = floorf(bits);
  fc.f = __fsub_rn(bits, bf);
  fc.omf = __fsub_rn(1.f, fc.f);
  int lo = (int)bf;
  lo = lo < 2 ? 2 : (lo > 8 ? 8 : lo);
  fc.lo_idx = lo - 2;
  fc.hi_idx = (lo + 1 <= 8) ? lo - 1 : lo - 2;   // q_hi = q_lo when floor(b)+1 > 8
  return fc;
}


static inline QGeom make_geom(int B, int C, int H, int W, int Ht, int Wt, int VEC) {
  QGeom g;
  g.B = B; g.C = C; g.H = H; g.W = W; g.HW = H * W; g.Ht = Ht; g.Wt = Wt;
  g.sy = (float)Ht / (float)H;
  g.sx = (float)Wt / (float)W;
  g.nvec = (H * W) / VEC;
  g.nvec_total = (long long)B * g.nvec;
  return g;
}

// inference vector path: every aligned 4-pixel segment must lie in one row and one tile
static inline bool seg_ok(const void* a, const void* b, const void* mask, const void* codes, int HW, int W, int Wt,
                   int VEC) {
  return ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)mask & 15) == 0 &&
         ((uintptr_t)codes & 7) == 0 && HW % VEC == 0 && W % 4 == 0 && W % Wt == 0 && (W / Wt) % 4 == 0;
}

// Sum of v over maximal runs of consecutive lanes holding the same id; the first lane of a run adds
// it to dst[id] (id < 0: nothing).  Whole warp must call.  Runs are numbered by counting run heads,
// so two separate runs with the same id (a tile met again one pixel row later) stay separate.
__device__ __forceinline__ void run_reduce_atomic(float v, int id, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int idp = __shfl_up_sync(0xffffffffu, id, 1);
  const bool head = lane == 0 || idp != id;
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float t = __shfl_down_sync(0xffffffffu, v, off);
    const int rn = __shfl_down_sync(0xffffffffu, run, off);
    if (lane + off < 32 && rn == run) v += t;
  }
  if (id >= 0 && head) atomicAdd(dst + id, v);
}

// training vector path (tile_quantize_train.cu); teacher / kd_* may be NULL (no distillation term)
bool train_vec_ok(const void* a, const void* b, const void* c, const void* mask, const void* dmask,
                  const void* teacher, int dtype, int H, int W, int Wt);
int train_fwd_vec(const void* x, void* y, int dtype, int B, int C, int H, int W, const float* bit_map, int Ht,
                  int Wt, const float* qtable, const float* mask, const float* teacher, double* kd_sum,
                  cudaStream_t st);
int train_bwd_vec(const void* gy, const void* x, void* gx, int dtype, int B, int C, int H, int W,
                  const float* bit_map, int Ht, int Wt, const float* qtable, const float* mask,
                  const float* teacher, const float* kd_coef, float* dbit, float* dmask, cudaStream_t st);

}  // namespace mcaq
