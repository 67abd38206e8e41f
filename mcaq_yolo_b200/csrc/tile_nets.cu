// Standalone launches of the tile-level stages (one CTA per image); the device code lives in
// tile_nets.cuh and is shared with the fused per-image kernel (morph_fused.cu).
#include "tile_nets.cuh"

namespace mcaq {

constexpr int TN_THREADS = 256;
constexpr int TN_WARPS = TN_THREADS / 32;

__global__ void __launch_bounds__(TN_THREADS)
complexity_kernel(const float* __restrict__ phi, int ht, int wt, const float* __restrict__ cmlp,
                  float* __restrict__ raw_out, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int ntiles = ht * wt;
  float* scratch = sm;                                  // max(TN_WARPS * NET_WARP_SCRATCH, 25 * ntiles)
  const int nscr = TN_WARPS * NET_WARP_SCRATCH > 25 * ntiles ? TN_WARPS * NET_WARP_SCRATCH : 25 * ntiles;
  float* craw = scratch + ((nscr + 3) & ~3);
  float* cfin = craw + ntiles;
  const int b = blockIdx.x;
  complexity_mlp_warps(phi + (long long)b * ntiles * 8, 0, ntiles, cmlp, scratch, craw,
                       raw_out ? raw_out + (long long)b * ntiles : nullptr);
  __syncthreads();
  bilateral_range(craw, ht, wt, 0, ntiles, scratch, cfin, out + (long long)b * ntiles);
}

__global__ void __launch_bounds__(TN_THREADS)
mapper_mlp_kernel(const float* __restrict__ cmap, int ntiles, const float* __restrict__ mp, float temperature,
                  int use_t, int continuous, float lo, float hi, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* scratch = sm;                                  // TN_WARPS * NET_WARP_SCRATCH
  float* bits_s = scratch + TN_WARPS * NET_WARP_SCRATCH;
  const int b = blockIdx.x;
  mapper_mlp_warps(cmap + (long long)b * ntiles, 0, ntiles, mp, scratch, temperature, use_t, continuous, lo, hi,
                   bits_s, out + (long long)b * ntiles);
}

// Builds the step table of mapper_steps_range by bisection on the fp32 bit pattern of c in [0, 1],
// evaluating the mapper with mapper_mlp_warps itself (all 8 steps advance together, 31 rounds).
__global__ void __launch_bounds__(TN_THREADS)
mapper_steps_kernel(const float* __restrict__ mp, float temperature, int use_t, float lo, float hi,
                    float* __restrict__ steps) {
  extern __shared__ __align__(16) float sm[];
  float* scratch = sm;                                  // TN_WARPS * NET_WARP_SCRATCH
  float* cand = scratch + TN_WARPS * NET_WARP_SCRATCH;  // [8]
  float* bits = cand + 8;                               // [8]
  __shared__ unsigned L[MAPPER_STEPS], R[MAPPER_STEPS];
  __shared__ int mode[MAPPER_STEPS];                    // 0 bisect, 1 always, 2 never
  const int tid = threadIdx.x;
  const int nsteps = (int)__fsub_rn(hi, lo);
  if (tid < 8) cand[tid] = tid == 0 ? 0.f : 1.f;
  __syncthreads();
  mapper_mlp_warps(cand, 0, 8, mp, scratch, temperature, use_t, 0, lo, hi, bits, nullptr);
  __syncthreads();
  if (tid < MAPPER_STEPS) {
    const float target = __fadd_rn(lo, (float)(tid + 1));
    const float b0 = bits[0], b1 = bits[1];
    mode[tid] = (tid >= nsteps || !(b1 >= target)) ? 2 : (b0 >= target ? 1 : 0);
    L[tid] = 0u;                                        // bits(L) <  target
    R[tid] = 0x3f800000u;                               // bits(R) >= target
  }
  __syncthreads();
  for (int it = 0; it < 31; ++it) {
    if (tid < MAPPER_STEPS) cand[tid] = __uint_as_float((L[tid] + R[tid]) >> 1);
    __syncthreads();
    mapper_mlp_warps(cand, 0, 8, mp, scratch, temperature, use_t, 0, lo, hi, bits, nullptr);
    __syncthreads();
    if (tid < MAPPER_STEPS && mode[tid] == 0) {
      const unsigned mid = (L[tid] + R[tid]) >> 1;
      if (mid != L[tid]) {
        if (bits[tid] >= __fadd_rn(lo, (float)(tid + 1))) R[tid] = mid; else L[tid] = mid;
      }
    }
    __syncthreads();
  }
  if (tid < MAPPER_STEPS)
    steps[tid] = mode[tid] == 1 ? 0.f : (mode[tid] == 2 ? INFINITY : __uint_as_float(R[tid]));
  __syncthreads();
  if (tid == 0) {
    bool ok = nsteps >= 1 && nsteps <= MAPPER_STEPS;
    float prev = 0.f;
    for (int k = 0; k < MAPPER_STEPS; ++k) {
      const float v = mode[k] == 1 ? 0.f : (mode[k] == 2 ? INFINITY : __uint_as_float(R[k]));
      if (v < prev) ok = false;
      prev = v;
    }
    steps[8] = ok ? 1.f : 0.f;
    steps[9] = use_t ? temperature : 0.f;
    steps[10] = lo;
    steps[11] = hi;
  }
}

__global__ void __launch_bounds__(TN_THREADS)
mapper_linear_kernel(const float* __restrict__ cmap, int ntiles, float temperature, int use_t,
                     int continuous, float lo, float hi, float eps_spread, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* cm = sm;                        // [ntiles] staged complexity
  float* bits_s = cm + ntiles;
  float* sel = bits_s + ntiles;          // 4
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) cm[t] = cmap[(long long)b * ntiles + t];
  __syncthreads();
  mapper_linear_range(cm, ntiles, sel, 0, ntiles, temperature, use_t, continuous, lo, hi, eps_spread, bits_s,
                      out + (long long)b * ntiles);
}

__global__ void __launch_bounds__(TN_THREADS)
soft_mask_kernel(const float* __restrict__ bit_map, int Ht, int Wt, const float* __restrict__ abs_plane,
                 int C, int H, int W, const float* __restrict__ prm, float* __restrict__ tiles_out,
                 float* __restrict__ mask) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, nt = Ht * Wt;
  float* P = sm;                         // 196
  float* act = P + SOFTMASK_SMEM_FLOATS; // [nt]
  float* bn = act + nt;
  float* an = bn + nt;
  float* mt = an + nt;
  float* bits = mt + nt;                 // [nt] staged copy of the bit map
  copy_params(prm, P, SOFTMASK_SMEM_FLOATS);
  for (int t = threadIdx.x; t < nt; t += blockDim.x) bits[t] = bit_map[(long long)b * nt + t];
  softmask_act_generic(abs_plane + (long long)b * H * W, C, H, W, Ht, Wt, 0, Ht, act);
  __syncthreads();
  float amax = -INFINITY;
  for (int t = threadIdx.x & 31; t < nt; t += 32) amax = fmaxf(amax, act[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  softmask_head_range(bits, act, amax, Ht, Wt, 0, nt, P, bn, an, mt,
                      tiles_out ? tiles_out + (long long)b * nt : nullptr);
  softmask_plane_rows(mt, P, H, W, Ht, Wt, 0, H, mask + (long long)b * H * W);
}

}  // namespace mcaq

using namespace mcaq;

extern "C" int mcaq_complexity(const float* phi, int B, int ht, int wt, const float* cmlp, const float* consts,
                               float* complexity_raw, float* complexity, void* stream) {
  (void)consts;   // the stencil constants are compiled in (mcaq_consts.cuh); kept for ABI stability
  if (!phi || !cmlp || !complexity || B <= 0 || ht <= 0 || wt <= 0) return MCAQ_EINVAL;
  const int nt_ = ht * wt;
  const int nscr = TN_WARPS * NET_WARP_SCRATCH > 25 * nt_ ? TN_WARPS * NET_WARP_SCRATCH : 25 * nt_;
  const size_t smem = (size_t)(((nscr + 3) & ~3) + 2 * nt_) * 4;
  if ((uintptr_t)cmlp & 15) return MCAQ_EALIGN;
  if (smem > 200 * 1024) return MCAQ_ETOOBIG;
  if (smem > 48 * 1024) cudaFuncSetAttribute(complexity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  complexity_kernel<<<B, TN_THREADS, smem, (cudaStream_t)stream>>>(phi, ht, wt, cmlp, complexity_raw, complexity);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_mapper_steps(const float* mapper, float temperature, int use_temperature, float min_bits,
                                 float max_bits, float* steps, void* stream) {
  if (!mapper || !steps || !(max_bits > min_bits)) return MCAQ_EINVAL;
  if (((uintptr_t)mapper | (uintptr_t)steps) & 15) return MCAQ_EALIGN;
  const size_t smem = (size_t)(TN_WARPS * NET_WARP_SCRATCH + 16) * sizeof(float);
  mapper_steps_kernel<<<1, TN_THREADS, smem, (cudaStream_t)stream>>>(mapper, temperature, use_temperature,
                                                                      min_bits, max_bits, steps);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_bit_mapper(const float* complexity, int B, int ht, int wt, const float* mapper,
                               float temperature, int use_temperature, int continuous, float min_bits,
                               float max_bits, float eps_spread, float* bit_map, void* stream) {
  if (!complexity || !bit_map || B <= 0 || ht <= 0 || wt <= 0) return MCAQ_EINVAL;
  const int ntiles = ht * wt;
  cudaStream_t st = (cudaStream_t)stream;
  if (mapper) {
    const size_t smem = (size_t)(TN_WARPS * NET_WARP_SCRATCH + ntiles) * 4;
    if ((uintptr_t)mapper & 15) return MCAQ_EALIGN;
    if (smem > 200 * 1024) return MCAQ_ETOOBIG;
    if (smem > 48 * 1024) cudaFuncSetAttribute(mapper_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mapper_mlp_kernel<<<B, TN_THREADS, smem, st>>>(complexity, ntiles, mapper, temperature, use_temperature,
                                                   continuous, min_bits, max_bits, bit_map);
  } else {
    const size_t smem = (size_t)(2 * ntiles + 4) * 4;
    if (smem > 200 * 1024) return MCAQ_ETOOBIG;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(mapper_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mapper_linear_kernel<<<B, TN_THREADS, smem, st>>>(complexity, ntiles, temperature, use_temperature,
                                                      continuous, min_bits, max_bits, eps_spread, bit_map);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_soft_mask(const float* bit_map, int Ht, int Wt, const float* abs_plane, int B, int C, int H,
                              int W, const float* softmask, float* mask_tiles, float* mask, void* stream) {
  if (!bit_map || !abs_plane || !softmask || !mask || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0)
    return MCAQ_EINVAL;
  const size_t smem = (size_t)(SOFTMASK_SMEM_FLOATS + 5 * Ht * Wt) * 4;
  if ((uintptr_t)softmask & 15) return MCAQ_EALIGN;
  if (smem > 200 * 1024) return MCAQ_ETOOBIG;
  if (smem > 48 * 1024) cudaFuncSetAttribute(soft_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  soft_mask_kernel<<<B, TN_THREADS, smem, (cudaStream_t)stream>>>(bit_map, Ht, Wt, abs_plane, C, H, W, softmask,
                                                                 mask_tiles, mask);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
