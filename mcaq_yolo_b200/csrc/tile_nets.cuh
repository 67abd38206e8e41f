// Device functions for the tile-level stages (complexity MLP + bilateral, bit mappers, soft
// mask).  Used by the standalone kernels in tile_nets.cu and by the fused per-image kernel in
// morph_fused.cu.
//
// The two MLPs run as warp-private pipelines: a warp carries a group of NET_TT tiles through all
// layers (lane = output unit, activations in a per-warp shared-memory scratch read back as
// broadcast LDS.128), so there is no CTA barrier between layers and the NET_TT independent FMA
// chains hide each other's latency.  Parameter blocks arrive in the layout the host packs
// (mcaq_yolo_b200/constants.py: weights transposed to [k][unit], BatchNorm folded) and are read
// in place through L1 (coalesced per-lane loads; every CTA shares the same 30 KB).
//
// Arithmetic = oracle/mcaq_oracle.py: nn.Linear / conv = FMA chain over k from 0, bias last;
// LayerNorm statistics = 32-lane xor-butterfly tree (element k and k+32 pre-added for D = 64);
// transcendentals in fp64, rounded once.
#pragma once
#include "common.cuh"
#include "mcaq_consts.cuh"

namespace mcaq {

// fp64 transcendentals rounded once to fp32 (the oracle's contract).  Deliberately NOT inlined: each
// expansion is ~1.5 KB of straight-line code used once per call site, and the per-image kernel is
// instruction-fetch bound.
// exp: k = rint(x log2 e), r = x - k ln2 (two-part ln2, exact products), exp(r) by the degree-13 Taylor
// polynomial (|r| <= 0.347: truncation 4e-18), scaled by 2^k -- about 1 ulp in fp64 like the library exp at
// a fifth of its instructions (the bilateral filter evaluates 25 per tile).
static __device__ __noinline__ float exp_f64(float xf) {
  const double x = (double)xf;
  if (!(fabs(x) < 700.0)) return (float)exp(x);            // inf / nan / beyond the scaling range: library
  const int k = __double2int_rn(x * 1.4426950408889634074);
  const double kd = (double)k;
  const double r = fma(-kd, 1.90821492927058770002e-10, fma(-kd, 6.93147180369123816490e-01, x));
  double p = 1.6059043836821613e-10;                        // 1 / 13!
  p = fma(p, r, 2.08767569878681e-09);                      // 1 / 12!
  p = fma(p, r, 2.505210838544172e-08);                     // 1 / 11!
  p = fma(p, r, 2.755731922398589e-07);                     // 1 / 10!
  p = fma(p, r, 2.7557319223985893e-06);                    // 1 / 9!
  p = fma(p, r, 2.48015873015873e-05);                      // 1 / 8!
  p = fma(p, r, 1.984126984126984e-04);                     // 1 / 7!
  p = fma(p, r, 1.388888888888889e-03);                     // 1 / 6!
  p = fma(p, r, 8.333333333333333e-03);                     // 1 / 5!
  p = fma(p, r, 4.1666666666666664e-02);                    // 1 / 4!
  p = fma(p, r, 1.6666666666666666e-01);                    // 1 / 3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return (float)(p * __hiloint2double((k + 1023) << 20, 0));
}
static __device__ __noinline__ float log1p_f64(float x) { return (float)log1p((double)x); }

__device__ __forceinline__ float sigmoid_exact(float z) {
  const float e = exp_f64(-z);
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

// n / d for 0 <= n, n * d < 2^32, with magic = 2^32 / d + 1 (one IMAD.HI instead of the division sequence)
// (d = 1 gives magic 0 = "2^32": the quotient is n itself)
__device__ __forceinline__ uint32_t div_magic(int d) { return 0xffffffffu / (uint32_t)d + 1u; }
__device__ __forceinline__ int fast_div(int n, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)n, magic) : n; }

// sum over the 32 lanes in xor-butterfly order; every lane returns the same bits
__device__ __forceinline__ float warp_tree_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- parameter blocks (floats, every block a multiple of 4 so 16-byte copies work) -------------
// complexity MLP: W0t[8][64] | b0[64] g1[64] be1[64] | W3t[64][32] | b3[32] g4[32] be4[32] | W6[32] b6 pad3
constexpr int CMLP_SMEM_FLOATS = 512 + 192 + 2048 + 96 + 36;             // 2884 == MCAQ_CMLP_FLOATS
// mapper: W0t[3][32] | b0 a0 be0 [32] | W3t[32][64] | b3 a3 be3 [64] | W6t[64][32] | b6 a6 be6 [32] | W9[32] b9 pad3
constexpr int MAPPER_SMEM_FLOATS = 96 + 96 + 2048 + 192 + 2048 + 96 + 36;   // 4612 == MCAQ_MAPPER_FLOATS
// soft mask: W0[8][2][3][3] | b0[8] | W2[2][8] | b2[2] | smooth[5][5] | pad1
constexpr int SOFTMASK_SMEM_FLOATS = 196;                                   // == MCAQ_SOFTMASK_FLOATS

constexpr int NET_TT = 4;                                    // tiles a warp carries through the mapper at once
constexpr int CM_TT = 8;                                     // ... and through the complexity MLP (packed pairs)
// floats of scratch per warp: mapper NET_TT * (64 + 32 + 32 + 4) = 528, complexity MLP CM_TT * (64 + 32) = 768
constexpr int NET_WARP_SCRATCH = CM_TT * (64 + 32);

// plain cooperative copy global -> shared (n % 4 == 0, both 16-byte aligned)
__device__ __forceinline__ void copy_params(const float* __restrict__ src, float* dst, int n) {
  for (int i = threadIdx.x; i < n / 4; i += blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
}

// asynchronous copy global -> shared (LDGSTS, 16 bytes per operation); pair with cp_async_wait_all
__device__ __forceinline__ void copy_params_async(const float* __restrict__ src, float* dst, int n) {
  for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 4 * i);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 4 * i) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// LayerNorm + ReLU of one value per lane (D = 32) or two values per lane (D = 64)
__device__ __forceinline__ void ln64_relu(float& a0, float& a1, float g0, float g1, float be0, float be1) {
  const float mean = __fmul_rn(warp_tree_sum(__fadd_rn(a0, a1)), 0.015625f);        // / 64, exact scaling
  const float d0 = __fsub_rn(a0, mean), d1 = __fsub_rn(a1, mean);
  const float var = __fmul_rn(warp_tree_sum(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1))), 0.015625f);
  const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-5f)));
  a0 = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d0, rstd), g0), be0), 0.f);
  a1 = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d1, rstd), g1), be1), 0.f);
}
__device__ __forceinline__ float ln32_relu(float a0, float g, float be) {
  const float mean = __fmul_rn(warp_tree_sum(a0), 0.03125f);                          // / 32, exact scaling
  const float d0 = __fsub_rn(a0, mean);
  const float var = __fmul_rn(warp_tree_sum(__fmul_rn(d0, d0)), 0.03125f);
  const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-5f)));
  return fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d0, rstd), g), be), 0.f);
}

// The same LayerNorm + ReLU for N tiles at once: the N butterflies run as independent shuffle chains and
// the 1 / sqrt(var + eps) of tile j (an IEEE division and square root: slow-path calls) is evaluated once
// by lane j for all tiles in parallel, then broadcast.  Per tile exactly the operations of ln64_relu /
// ln32_relu, so the values are bit-identical.
template <int N>
__device__ __forceinline__ void warp_tree_sum_n(float (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = __fadd_rn(v[j], __shfl_xor_sync(0xffffffffu, v[j], o));
  }
}
template <int N>
__device__ __forceinline__ void rstd_n(const float (&var)[N], float (&rstd)[N]) {
  static_assert(N <= 32, "one lane per tile");
  const int lane = threadIdx.x & 31;
  float mine = var[0];
#pragma unroll
  for (int j = 1; j < N; ++j) mine = (lane == j) ? var[j] : mine;
  const float r = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mine, 1e-5f)));
#pragma unroll
  for (int j = 0; j < N; ++j) rstd[j] = __shfl_sync(0xffffffffu, r, j);
}
template <int N>
__device__ __forceinline__ void ln64_relu_n(float2 (&a)[N], float g0, float g1, float be0, float be1) {
  float s[N], rstd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = __fadd_rn(a[j].x, a[j].y);
  warp_tree_sum_n<N>(s);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float mean = __fmul_rn(s[j], 0.015625f);
    a[j].x = __fsub_rn(a[j].x, mean);
    a[j].y = __fsub_rn(a[j].y, mean);
    s[j] = __fadd_rn(__fmul_rn(a[j].x, a[j].x), __fmul_rn(a[j].y, a[j].y));
  }
  warp_tree_sum_n<N>(s);
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = __fmul_rn(s[j], 0.015625f);
  rstd_n<N>(s, rstd);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    a[j].x = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(a[j].x, rstd[j]), g0), be0), 0.f);
    a[j].y = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(a[j].y, rstd[j]), g1), be1), 0.f);
  }
}
template <int N>
__device__ __forceinline__ void ln32_relu_n(float (&a)[N], float g, float be) {
  float s[N], rstd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = a[j];
  warp_tree_sum_n<N>(s);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    a[j] = __fsub_rn(a[j], __fmul_rn(s[j], 0.03125f));
    s[j] = __fmul_rn(a[j], a[j]);
  }
  warp_tree_sum_n<N>(s);
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = __fmul_rn(s[j], 0.03125f);
  rstd_n<N>(s, rstd);
#pragma unroll
  for (int j = 0; j < N; ++j) a[j] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(a[j], rstd[j]), g), be), 0.f);
}

// out[j] (+)= sum_k act[j][k] * Wt[k][unit]: FMA chain over k from 0 for NET_TT tiles at once.
// act rows are 16-byte aligned shared memory (broadcast LDS.128), Wt column reads are conflict-free.
template <int K_IN, int N_OUT>
__device__ __forceinline__ void dense_warp(const float* act, int act_stride, const float* __restrict__ Wt, int unit,
                                           float (&acc)[NET_TT]) {
#pragma unroll
  for (int j = 0; j < NET_TT; ++j) acc[j] = 0.f;
#pragma unroll 4
  for (int k4 = 0; k4 < K_IN / 4; ++k4) {
    const float w0 = __ldg(Wt + (4 * k4 + 0) * N_OUT + unit), w1 = __ldg(Wt + (4 * k4 + 1) * N_OUT + unit);
    const float w2 = __ldg(Wt + (4 * k4 + 2) * N_OUT + unit), w3 = __ldg(Wt + (4 * k4 + 3) * N_OUT + unit);
#pragma unroll
    for (int j = 0; j < NET_TT; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(act + j * act_stride + 4 * k4);
      acc[j] = fmaf(a.x, w0, acc[j]);
      acc[j] = fmaf(a.y, w1, acc[j]);
      acc[j] = fmaf(a.z, w2, acc[j]);
      acc[j] = fmaf(a.w, w3, acc[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// complexity MLP 8 -> 64 (LN, ReLU) -> 32 (LN, ReLU) -> 1, sigmoid  (morphology.py:81-97)
// craw[t] for t in [t_lo, t_hi).  phi: [ntiles][8] floats, 16-byte aligned rows (shared or global);
// w: CMLP block, global or shared memory (plain loads; the fused kernel stages it into shared memory:
// L1 starts cold in every launch and layer 2 would walk the block in 16 dependent L2 round trips);
// scratch: NET_WARP_SCRATCH floats of shared memory per warp.  Warp-level only: the
// caller synchronises the CTA before anyone reads craw.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void complexity_mlp_warps(const float* phi, int t_lo, int t_hi,
                                                     const float* __restrict__ w, float* scratch_all, float* craw,
                                                     float* __restrict__ raw_out, long long* dbg = nullptr) {
  // dbg (tools/stage_clocks.py): clock64() after layer 1 / LayerNorm + park / layer 2 of warp 0's first group
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#define CM_DBG(k) do { if (dbg && threadIdx.x == 0 && tg == t_lo) dbg[k] = clock64(); } while (0)
  const float* W0t = w;
  const float* b0 = w + 512;
  const float* g1 = w + 576;
  const float* be1 = w + 640;
  const float* W3t = w + 704;
  const float* b3 = w + 2752;
  const float* g4 = w + 2784;
  const float* be4 = w + 2816;
  const float* W6 = w + 2848;
  const float b6 = *(w + 2880);
  // a warp carries CM_TT = 8 tiles; activations of layer 1 are parked as h1[k][tile] so that one
  // LDS.128 yields two (tile, tile + 1) pairs for the packed FMAs of layer 2
  float* h1 = scratch_all + warp * NET_WARP_SCRATCH;      // [64][CM_TT]
  float* h2 = h1 + 64 * CM_TT;                            // [CM_TT][32]
  for (int tg = t_lo + warp * CM_TT; tg < t_hi; tg += nwarps * CM_TT) {
    // layer 1 (8 -> 64): lane owns units lane and lane + 32 = one packed accumulator per tile
    float2 a[CM_TT];
    {
      float2 w01[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) w01[k] = make_float2(*(W0t + k * 64 + lane), *(W0t + k * 64 + 32 + lane));
#pragma unroll
      for (int j = 0; j < CM_TT; ++j) {
        const float4* pr = reinterpret_cast<const float4*>(phi + min(tg + j, t_hi - 1) * 8);
        const float4 pa = pr[0], pb = pr[1];
        float2 acc = make_float2(0.f, 0.f);
        acc = ffma2(splat2(pa.x), w01[0], acc);
        acc = ffma2(splat2(pa.y), w01[1], acc);
        acc = ffma2(splat2(pa.z), w01[2], acc);
        acc = ffma2(splat2(pa.w), w01[3], acc);
        acc = ffma2(splat2(pb.x), w01[4], acc);
        acc = ffma2(splat2(pb.y), w01[5], acc);
        acc = ffma2(splat2(pb.z), w01[6], acc);
        acc = ffma2(splat2(pb.w), w01[7], acc);
        a[j] = acc;
      }
    }
    CM_DBG(13);
    {
      const float bb0 = *(b0 + lane), bb1 = *(b0 + lane + 32);
      const float gg0 = *(g1 + lane), gg1 = *(g1 + lane + 32), ee0 = *(be1 + lane), ee1 = *(be1 + lane + 32);
#pragma unroll
      for (int j = 0; j < CM_TT; ++j) a[j] = make_float2(__fadd_rn(a[j].x, bb0), __fadd_rn(a[j].y, bb1));
      ln64_relu_n<CM_TT>(a, gg0, gg1, ee0, ee1);
      float4* r0 = reinterpret_cast<float4*>(h1 + lane * CM_TT);
      float4* r1 = reinterpret_cast<float4*>(h1 + (lane + 32) * CM_TT);
      r0[0] = make_float4(a[0].x, a[1].x, a[2].x, a[3].x);
      r0[1] = make_float4(a[4].x, a[5].x, a[6].x, a[7].x);
      r1[0] = make_float4(a[0].y, a[1].y, a[2].y, a[3].y);
      r1[1] = make_float4(a[4].y, a[5].y, a[6].y, a[7].y);
    }
    __syncwarp();
    CM_DBG(14);
    // layer 2 (64 -> 32): lane = unit, packed accumulators over tile pairs, weights one block ahead
    float2 acc[CM_TT / 2];
#pragma unroll
    for (int p = 0; p < CM_TT / 2; ++p) acc[p] = make_float2(0.f, 0.f);
    float wc[4], wn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) wc[q] = *(W3t + q * 32 + lane);
#pragma unroll 1
    for (int k4 = 0; k4 < 16; ++k4) {
      const int kn = k4 < 15 ? k4 + 1 : k4;
#pragma unroll
      for (int q = 0; q < 4; ++q) wn[q] = *(W3t + (4 * kn + q) * 32 + lane);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4* hr = reinterpret_cast<const float4*>(h1 + (4 * k4 + q) * CM_TT);
        const float4 ha = hr[0], hb = hr[1];
        const float2 ws = splat2(wc[q]);
        acc[0] = ffma2(make_float2(ha.x, ha.y), ws, acc[0]);
        acc[1] = ffma2(make_float2(ha.z, ha.w), ws, acc[1]);
        acc[2] = ffma2(make_float2(hb.x, hb.y), ws, acc[2]);
        acc[3] = ffma2(make_float2(hb.z, hb.w), ws, acc[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) wc[q] = wn[q];
    }
    CM_DBG(15);
    {
      const float bb = *(b3 + lane), gg = *(g4 + lane), ee = *(be4 + lane);
      float v[CM_TT];
#pragma unroll
      for (int p = 0; p < CM_TT / 2; ++p) {
        v[2 * p] = __fadd_rn(acc[p].x, bb);
        v[2 * p + 1] = __fadd_rn(acc[p].y, bb);
      }
      ln32_relu_n<CM_TT>(v, gg, ee);
#pragma unroll
      for (int j = 0; j < CM_TT; ++j) h2[j * 32 + lane] = v[j];
    }
    __syncwarp();
    if (lane < CM_TT && tg + lane < t_hi) {                // layer 3: 32 -> 1 and sigmoid, one lane per tile
      float z = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 av = *reinterpret_cast<const float4*>(h2 + lane * 32 + 4 * k4);
        const float4 ww = *reinterpret_cast<const float4*>(W6 + 4 * k4);
        z = fmaf(av.x, ww.x, z); z = fmaf(av.y, ww.y, z); z = fmaf(av.z, ww.z, z); z = fmaf(av.w, ww.w, z);
      }
      const float c = sigmoid_exact(__fadd_rn(z, b6));
      craw[tg + lane] = c;
      if (raw_out) raw_out[tg + lane] = c;
    }
    __syncwarp();
  }
}
#undef CM_DBG

// 5x5 bilateral filter with replicate padding (morphology.py:309-354) and clamp, for tiles
// [t_lo, t_hi); craw must hold ALL tiles of the image.  wgt: 25*(t_hi-t_lo) floats of scratch.
// Block-level (contains barriers).
__device__ __forceinline__ void bilateral_range(const float* craw, int ht, int wt, int t_lo, int t_hi, float* wgt,
                                                float* cfin, float* __restrict__ out) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int n = t_hi - t_lo;
  const uint32_t mwt = div_magic(wt);
  for (int o = tid; o < n * 25; o += NT) {        // range weights for all (tile, tap) pairs in parallel
    const int tl = o / 25, tap = o - tl * 25;
    const int t = t_lo + tl;
    const int y = fast_div(t, mwt), x = t - y * wt;
    const int ky = tap / 5, kx = tap - ky * 5;
    const int yy = min(max(y + ky - 2, 0), ht - 1), xx = min(max(x + kx - 2, 0), wt - 1);
    const float d = __fsub_rn(craw[yy * wt + xx], craw[t]);
    // -(d^2) / (2 sigma_r^2): Markstein's correction with the constant's reciprocal equals div.rn in the
    // normal range (tests/test_gpu_division.py); zero / tiny numerators take the division
    const float nd = -__fmul_rn(d, d);
    const float arg = (nd <= -1e-30f) ? div_markstein(nd, kc::BILAT_DEN, 50.0f) : __fdiv_rn(nd, kc::BILAT_DEN);
    wgt[o] = __fmul_rn(kc::BILAT[tap], exp_f64(arg));
  }
  __syncthreads();
  for (int tl = tid; tl < n; tl += NT) {          // ordered accumulation of the 25 taps
    const int t = t_lo + tl;
    const int y = fast_div(t, mwt), x = t - y * wt;
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int yy = min(max(y + ky - 2, 0), ht - 1);
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int xx = min(max(x + kx - 2, 0), wt - 1);
        const float wg = wgt[tl * 25 + ky * 5 + kx];
        num = __fadd_rn(num, __fmul_rn(wg, craw[yy * wt + xx]));
        den = __fadd_rn(den, wg);
      }
    }
    const float r = fminf(fmaxf(__fdiv_rn(num, __fadd_rn(den, 1e-8f)), 0.f), 1.f);
    cfin[t] = r;
    if (out) out[t] = r;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// bit mappers (eval).  finish: temperature, straight-through clamp / round (forward values)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float finish_bits(float bits, float temperature, int use_t, int continuous,
                                             float lo, float hi) {
  if (use_t) bits = __fmul_rn(bits, temperature);
  const float cl = fminf(fmaxf(bits, lo), hi);
  bits = __fadd_rn(bits, __fsub_rn(cl, bits));
  if (!continuous) bits = __fadd_rn(bits, __fsub_rn(rintf(bits), bits));
  return bits;
}

// ComplexityToBitMappingNetwork (bit_allocation.py:218-280, eval BN folded): bits_s[t] for t in
// [t_lo, t_hi); cmap indexed by absolute tile.  Warp-level only, like complexity_mlp_warps.
__device__ __forceinline__ void mapper_mlp_warps(const float* cmap, int t_lo, int t_hi, const float* __restrict__ w,
                                                 float* scratch_all, float temperature, int use_t, int continuous,
                                                 float lo, float hi, float* bits_s, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* W0t = w;                  // [3][32]
  const float* v0 = w + 96;              // b, alpha, beta [32]
  const float* W3t = w + 192;            // [32][64]
  const float* v3 = w + 2240;            // b, alpha, beta [64]
  const float* W6t = w + 2432;           // [64][32]
  const float* v6 = w + 4480;            // b, alpha, beta [32]
  const float* W9 = w + 4576;            // [32], b9
  float* g1 = scratch_all + warp * NET_WARP_SCRATCH;      // [TT][64]
  float* g0 = g1 + NET_TT * 64;                           // [TT][32]
  float* g2 = g0 + NET_TT * 32;                           // [TT][32]
  float* zin = g2 + NET_TT * 32;                          // [TT][4]
  for (int tg = t_lo + warp * NET_TT; tg < t_hi; tg += nwarps * NET_TT) {
    if (lane < NET_TT) {                                  // z0 = [c, c^2, log1p(c)]  (Eq.13)
      const float c = fminf(fmaxf(cmap[min(tg + lane, t_hi - 1)], 0.f), 1.f);
      *reinterpret_cast<float4*>(zin + lane * 4) = make_float4(c, __fmul_rn(c, c), log1p_f64(c), 0.f);
    }
    __syncwarp();
    {                                                     // 3 -> 32, BN, ReLU
      const float w0 = __ldg(W0t + lane), w1 = __ldg(W0t + 32 + lane), w2 = __ldg(W0t + 64 + lane);
      const float bb = __ldg(v0 + lane), al = __ldg(v0 + 32 + lane), be = __ldg(v0 + 64 + lane);
#pragma unroll
      for (int j = 0; j < NET_TT; ++j) {
        const float4 z = *reinterpret_cast<const float4*>(zin + j * 4);
        float acc = __fmul_rn(z.x, w0);
        acc = fmaf(z.y, w1, acc);
        acc = fmaf(z.z, w2, acc);
        const float x = __fadd_rn(acc, bb);
        g0[j * 32 + lane] = fmaxf(__fadd_rn(__fmul_rn(x, al), be), 0.f);
      }
    }
    __syncwarp();
    {                                                     // 32 -> 64, BN, ReLU (two units per lane)
      float acc0[NET_TT], acc1[NET_TT];
      dense_warp<32, 64>(g0, 32, W3t, lane, acc0);
      dense_warp<32, 64>(g0, 32, W3t, lane + 32, acc1);
      const float b_0 = __ldg(v3 + lane), a_0 = __ldg(v3 + 64 + lane), e_0 = __ldg(v3 + 128 + lane);
      const float b_1 = __ldg(v3 + 32 + lane), a_1 = __ldg(v3 + 96 + lane), e_1 = __ldg(v3 + 160 + lane);
#pragma unroll
      for (int j = 0; j < NET_TT; ++j) {
        g1[j * 64 + lane] = fmaxf(__fadd_rn(__fmul_rn(__fadd_rn(acc0[j], b_0), a_0), e_0), 0.f);
        g1[j * 64 + 32 + lane] = fmaxf(__fadd_rn(__fmul_rn(__fadd_rn(acc1[j], b_1), a_1), e_1), 0.f);
      }
    }
    __syncwarp();
    {                                                     // 64 -> 32, BN, ReLU
      float acc[NET_TT];
      dense_warp<64, 32>(g1, 64, W6t, lane, acc);
      const float bb = __ldg(v6 + lane), al = __ldg(v6 + 32 + lane), be = __ldg(v6 + 64 + lane);
#pragma unroll
      for (int j = 0; j < NET_TT; ++j)
        g2[j * 32 + lane] = fmaxf(__fadd_rn(__fmul_rn(__fadd_rn(acc[j], bb), al), be), 0.f);
    }
    __syncwarp();
    if (lane < NET_TT && tg + lane < t_hi) {              // 32 -> 1, sigmoid, Eq.17, temperature / STE
      float acc = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 a = *reinterpret_cast<const float4*>(g2 + lane * 32 + 4 * k4);
        const float4 ww = __ldg(reinterpret_cast<const float4*>(W9 + 4 * k4));
        acc = fmaf(a.x, ww.x, acc); acc = fmaf(a.y, ww.y, acc); acc = fmaf(a.z, ww.z, acc); acc = fmaf(a.w, ww.w, acc);
      }
      const float s = sigmoid_exact(__fadd_rn(acc, __ldg(W9 + 32)));
      const float bits = finish_bits(__fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), s)), temperature, use_t,
                                     continuous, lo, hi);
      bits_s[tg + lane] = bits;
      if (out) out[tg + lane] = bits;
    }
    __syncwarp();
  }
}

// Step table of the MLP mapper (eval, integer output).  With non-negative weights (Eq.18) and folded
// BN scales the network is monotone in c, so bits(c) = lo + #{k : c >= steps[k]} where steps[k] is the
// smallest c in [0, 1] (by fp32 bit pattern) at which mapper_mlp_warps itself returns >= lo + 1 + k
// (0: always, +inf: never).  The table is built by bisection WITH mapper_mlp_warps
// (tile_nets.cu: mapper_steps_kernel), so both agree everywhere except possibly within a few ulps of a
// step, where the continuous bit value sits on a .5 rounding boundary (SURVEY 7.3 ambiguity set).
// steps[8] == 1 marks a valid (monotone) table.
constexpr int MAPPER_STEPS = 8;
constexpr int MAPPER_STEPS_FLOATS = 12;                       // 8 steps | valid | temperature | lo | hi
__device__ __forceinline__ void mapper_steps_range(const float* cmap, int t_lo, int t_hi,
                                                   const float* __restrict__ steps, float lo, float* bits_s,
                                                   float* __restrict__ out) {
  float st[MAPPER_STEPS];
#pragma unroll
  for (int k = 0; k < MAPPER_STEPS; ++k) st[k] = __ldg(steps + k);
  for (int t = t_lo + threadIdx.x; t < t_hi; t += blockDim.x) {
    const float c = fminf(fmaxf(cmap[t], 0.f), 1.f);
    int n = 0;
#pragma unroll
    for (int k = 0; k < MAPPER_STEPS; ++k) n += (c >= st[k]) ? 1 : 0;
    const float b = __fadd_rn(lo, (float)n);
    bits_s[t] = b;
    if (out) out[t] = b;
  }
}

// torch.quantile(q, 'linear'): fp32 rank, torch.lerp formula on the two neighbouring order statistics
__device__ __forceinline__ void quantile_ranks(int n, float q, int& lo, int& hi, float& w) {
  const float rank = __fmul_rn(q, (float)(n - 1));
  lo = (int)floorf(rank);
  hi = (int)ceilf(rank);
  w = __fsub_rn(rank, (float)lo);
}
__device__ __forceinline__ float quantile_lerp(float a, float bb, float w) {
  const float diff = __fsub_rn(bb, a);
  if (w < 0.5f) return __fadd_rn(a, __fmul_rn(w, diff));
  return __fsub_rn(bb, __fmul_rn(diff, __fsub_rn(1.0f, w)));
}

// LinearBitMapper (bit_allocation.py:42-80): 2 % / 98 % quantiles over ALL tiles of the image
// (cmap: [ntiles]) by rank counting (each element counts the elements ordered before it; the four
// wanted order statistics are written by their owners), then bits for [t_lo, t_hi).  sel: 4 floats.
// Block-level (contains barriers).
__device__ __forceinline__ void mapper_linear_range(const float* cmap, int ntiles, float* sel, int t_lo, int t_hi,
                                                    float temperature, int use_t, int continuous, float lo,
                                                    float hi, float eps_spread, float* bits_s,
                                                    float* __restrict__ out) {
  const int tid = threadIdx.x, NT = blockDim.x;
  int l0, h0, l1, h1;
  float w0, w1;
  quantile_ranks(ntiles, 0.02f, l0, h0, w0);
  quantile_ranks(ntiles, 0.98f, l1, h1, w1);
  for (int i = tid; i < ntiles; i += NT) {
    const float v = cmap[i];
    int r = 0;
    for (int j = 0; j < ntiles; ++j) {
      const float u = cmap[j];
      r += (u < v || (u == v && j < i)) ? 1 : 0;
    }
    if (r == l0) sel[0] = v;
    if (r == h0) sel[1] = v;
    if (r == l1) sel[2] = v;
    if (r == h1) sel[3] = v;
  }
  __syncthreads();
  const float qlo = quantile_lerp(sel[0], sel[1], w0);
  const float qhi = quantile_lerp(sel[2], sel[3], w1);
  const float spread = __fsub_rn(qhi, qlo);
  for (int t = t_lo + tid; t < t_hi; t += NT) {
    const float v = cmap[t];
    float rel = __fdiv_rn(__fsub_rn(v, qlo), __fadd_rn(spread, 1e-8f));
    rel = fminf(fmaxf(rel, 0.f), 1.f);
    const float cn = spread > eps_spread ? rel : fminf(fmaxf(v, 0.f), 1.f);
    const float bits = finish_bits(__fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), cn)), temperature, use_t,
                                   continuous, lo, hi);
    bits_s[t] = bits;
    if (out) out[t] = bits;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// learned soft mask (quantization.py:213-239)
// packed params: W0[8][2][3][3] b0[8] W2[2][8] b2[2] smooth[5][5]  (195 floats + 1 pad)
// ---------------------------------------------------------------------------------------------
// (a) generic tile activity: act[t] = mean over the tile's adaptive-pool window of sum_c|x| / C for
//     tile rows [ty_lo, ty_hi): column sums top-to-bottom, then left-to-right (oracle order), one
//     thread per tile.  Used when the window grid is not the analyzer's tile grid.
__device__ __forceinline__ void softmask_act_generic(const float* __restrict__ ap, int C, int H, int W, int Ht,
                                                     int Wt, int ty_lo, int ty_hi, float* act) {
  const float fC = (float)C;
  const float rC = __frcp_rn(fC);
  for (int t = ty_lo * Wt + threadIdx.x; t < ty_hi * Wt; t += blockDim.x) {
    const int i = t / Wt, j = t - i * Wt;
    const int ys = (i * H) / Ht, ye = ((i + 1) * H + Ht - 1) / Ht;
    const int xs = (j * W) / Wt, xe = ((j + 1) * W + Wt - 1) / Wt;
    float tot = 0.f;
    for (int x = xs; x < xe; ++x) {
      float s = 0.f;
      for (int y = ys; y < ye; ++y) {
        const float v = __ldg(ap + y * W + x);
        const float av = fabsf(v);
        const float q = (av >= 1e-30f && av < 1e27f) ? div_markstein(v, fC, rC) : __fdiv_rn(v, fC);
        s = __fadd_rn(s, q);
      }
      tot = __fadd_rn(tot, s);
    }
    act[t] = __fdiv_rn(tot, (float)((ye - ys) * (xe - xs)));
  }
}

// (b) tile head for tiles [t_lo, t_hi): act / (amax + 1e-8), bits -> [0,1], conv3x3(2->8)+ReLU,
//     conv1x1(8->2), softmax channel 0.  bits/act hold ALL tiles; P: the parameter block in shared
//     memory; bn/an: nt floats each of scratch.  Block-level (contains barriers).
__device__ __forceinline__ void softmask_head_range(const float* bits, const float* act, float amax, int Ht, int Wt,
                                                    int t_lo, int t_hi, const float* P, float* bn, float* an,
                                                    float* mt, float* __restrict__ tiles_out) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int nt = Ht * Wt;
  const float aden = __fadd_rn(amax, 1e-8f);
  for (int t = tid; t < nt; t += NT) {
    an[t] = __fdiv_rn(act[t], aden);
    bn[t] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bits[t], 2.0f), 6.0f), 0.f), 1.f);
  }
  __syncthreads();
  const float* W0 = P;
  const float* b0 = P + 144;
  const float* W2 = P + 152;
  const float* b2 = P + 168;
  const uint32_t mwt = div_magic(Wt);
  for (int t = t_lo + tid; t < t_hi; t += NT) {
    const int i = fast_div(t, mwt), j = t - i * Wt;
    float hid[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int ic = 0; ic < 2; ++ic) {
        const float* src = ic == 0 ? bn : an;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = i + ky - 1;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = j + kx - 1;
            if (yy >= 0 && yy < Ht && xx >= 0 && xx < Wt)
              acc = fmaf(src[yy * Wt + xx], W0[((o * 2 + ic) * 3 + ky) * 3 + kx], acc);
          }
        }
      }
      hid[o] = fmaxf(__fadd_rn(acc, b0[o]), 0.f);
    }
    float lg[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int ic = 0; ic < 8; ++ic) acc = fmaf(hid[ic], W2[o * 8 + ic], acc);
      lg[o] = __fadd_rn(acc, b2[o]);
    }
    const float mx = fmaxf(lg[0], lg[1]);
    const float e0 = exp_f64(__fsub_rn(lg[0], mx));
    const float e1 = exp_f64(__fsub_rn(lg[1], mx));
    const float m = __fdiv_rn(e0, __fadd_rn(e0, e1));
    mt[t] = m;
    if (tiles_out) tiles_out[t] = m;
  }
  __syncthreads();
}

// (c) generic: rows [h_lo, h_hi) of m = smooth5x5(nearest_upsample(mt)), replicate padding, FMA
//     chain over taps in row-major order.  mt holds ALL tiles; P holds the params (smooth kernel at +170).
__device__ __forceinline__ void softmask_plane_rows(const float* mt, const float* P, int H, int W, int Ht, int Wt,
                                                    int h_lo, int h_hi, float* __restrict__ mo) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const float* ks = P + 170;
  const float sy = (float)Ht / (float)H, sx = (float)Wt / (float)W;
  for (int p = h_lo * W + tid; p < h_hi * W; p += NT) {
    const int h = p / W, w = p - h * W;
    int ix[5];
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) ix[kx] = nearest_src(min(max(w + kx - 2, 0), W - 1), sx, Wt);
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int iy = nearest_src(min(max(h + ky - 2, 0), H - 1), sy, Ht);
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) acc = fmaf(mt[iy * Wt + ix[kx]], ks[ky * 5 + kx], acc);
    }
    mo[p] = acc;
  }
}

// (c') tile-aligned planes (H == Ht*tile, W == Wt*tile, tile a power of two >= 4): inside a tile the
//     25 taps of a pixel come from at most 3x3 neighbouring tiles and the pattern depends only on the
//     pixel's offset class (0, 1, interior, tile-2, tile-1) per axis, so every tile has <= 25 distinct
//     mask values.  cls[(t - t_lo)*25 + cy*5 + cx] is that value (same FMA chain as the generic path).
__device__ __forceinline__ int mask_class(int d, int tile) { return d < 2 ? d : (d >= tile - 2 ? d - (tile - 5) : 2); }

__device__ __forceinline__ void softmask_class_table(const float* mt, const float* P, int Ht, int Wt, int tile,
                                                     int t_lo, int t_hi, float* cls) {
  // one thread per tile: 3x3 tile neighbourhood (replicate clamped) and the 5x5 kernel in registers,
  // all 25 classes unrolled so every tap is a register-register FFMA
  const uint32_t mwt = div_magic(Wt);
  for (int t = t_lo + threadIdx.x; t < t_hi; t += blockDim.x) {
    const int ty = fast_div(t, mwt), tx = t - ty * Wt;
    float n[3][3], ks[25];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = min(max(ty + dy - 1, 0), Ht - 1);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) n[dy][dx] = mt[yy * Wt + min(max(tx + dx - 1, 0), Wt - 1)];
    }
#pragma unroll
    for (int i = 0; i < 25; ++i) ks[i] = P[170 + i];
    float* dst = cls + (t - t_lo) * 25;
#pragma unroll
    for (int cy = 0; cy < 5; ++cy) {
#pragma unroll
      for (int cx = 0; cx < 5; ++cx) {
        if (tile == 4 && (cy == 2 || cx == 2)) continue;      // a 4-pixel tile has no interior class
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
          // tile-row offset of tap ky for class cy: the tile above for (cy=0: ky<2, cy=1: ky<1), the
          // tile below for (cy=3: ky>3, cy=4: ky>2)
          const int oy = (ky < 2 - cy) ? -1 : ((ky > 6 - cy) ? 1 : 0);
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const int ox = (kx < 2 - cx) ? -1 : ((kx > 6 - cx) ? 1 : 0);
            acc = fmaf(n[1 + oy][1 + ox], ks[ky * 5 + kx], acc);
          }
        }
        dst[cy * 5 + cx] = acc;
      }
    }
  }
}

// rows [h_lo, h_hi) of m from the class table (tiles [t_lo, ...) start at tile row h_lo / tile);
// one float4 (4 pixels of one tile) per thread-iteration.
__device__ __forceinline__ void softmask_plane_from_classes(const float* cls, int W, int Wt, int tile, int tshift,
                                                            int t_lo, int h_lo, int h_hi, float* __restrict__ mo) {
  const int W4 = W >> 2;
  const int n4 = (h_hi - h_lo) * W4;
  const uint32_t mw4 = div_magic(W4);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int hr = fast_div(i, mw4), w = (i - hr * W4) << 2;
    const int h = h_lo + hr;
    const int t = (h >> tshift) * Wt + (w >> tshift);
    const float* c = cls + (t - t_lo) * 25 + mask_class(h & (tile - 1), tile) * 5;
    const int d = w & (tile - 1);
    float4 v;
    v.x = c[mask_class(d, tile)];
    v.y = c[mask_class(d + 1, tile)];
    v.z = c[mask_class(d + 2, tile)];
    v.w = c[mask_class(d + 3, tile)];
    *reinterpret_cast<float4*>(mo + (long long)h * W + w) = v;
  }
}

}  // namespace mcaq
