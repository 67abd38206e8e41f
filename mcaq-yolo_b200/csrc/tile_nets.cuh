// Block-cooperative device functions for the tile-level stages (complexity MLP + bilateral,
// bit mappers, soft mask).  Used by the standalone kernels in tile_nets.cu and by the fused
// per-image kernel in morph_fused.cu.  All threads of the CTA must call each function; `sm`
// arguments are caller-provided shared-memory scratch of the documented size.
//
// Arithmetic = oracle/mcaq_oracle.py: nn.Linear / conv = FMA chain over k from 0, bias last;
// LayerNorm statistics = 32-lane xor-butterfly tree (element k and k+32 pre-added for D = 64);
// transcendentals in fp64, rounded once.
#pragma once
#include "common.cuh"
#include "mcaq_consts.cuh"

namespace mcaq {

__device__ __forceinline__ float sigmoid_exact(float z) {
  const float e = (float)exp((double)(-z));
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

// sum over the 32 lanes in xor-butterfly order; every lane returns the same bits
__device__ __forceinline__ float warp_tree_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// out(t, m) = sum_k in[t][k] * Wt[k][m] as an FMA chain over k from 0 (nn.Linear's order), for nb
// tiles.  A thread owns unit m for TT = 8 consecutive tiles: per 4 k it issues 8 LDS.128 of
// activations (warp-broadcast: the 32 lanes of a warp share the tile group) and 4 conflict-free
// weight loads for 32 FMAs.  in rows must be 16-byte aligned (in_stride % 4 == 0).
template <int K_IN, int N_OUT, typename Epi>
__device__ __forceinline__ void dense_tiled(const float* in, int in_stride, int nb, const float* Wt, Epi epi) {
  constexpr int TT = 8;
  static_assert(K_IN % 4 == 0 && N_OUT % 32 == 0, "dense_tiled shape");
  const int ngroups = (nb + TT - 1) / TT;
  for (int task = threadIdx.x; task < ngroups * N_OUT; task += blockDim.x) {
    const int gi = task / N_OUT, m = task - gi * N_OUT;
    const int tb = gi * TT;
    float acc[TT];
    const float4* ip[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      acc[j] = 0.f;
      ip[j] = reinterpret_cast<const float4*>(in + min(tb + j, nb - 1) * in_stride);
    }
#pragma unroll 2
    for (int k4 = 0; k4 < K_IN / 4; ++k4) {
      const float w0 = Wt[(4 * k4 + 0) * N_OUT + m], w1 = Wt[(4 * k4 + 1) * N_OUT + m];
      const float w2 = Wt[(4 * k4 + 2) * N_OUT + m], w3 = Wt[(4 * k4 + 3) * N_OUT + m];
#pragma unroll
      for (int j = 0; j < TT; ++j) {
        const float4 a = ip[j][k4];
        acc[j] = fmaf(a.x, w0, acc[j]);
        acc[j] = fmaf(a.y, w1, acc[j]);
        acc[j] = fmaf(a.z, w2, acc[j]);
        acc[j] = fmaf(a.w, w3, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < TT; ++j)
      if (tb + j < nb) epi(tb + j, m, acc[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// complexity MLP 8 -> 64 (LN, ReLU) -> 32 (LN, ReLU) -> 1, sigmoid; 5x5 bilateral; clamp
// smem need: cpx_scratch_floats(ntiles) + 2*ntiles
// ---------------------------------------------------------------------------------------------
constexpr int CMLP_SMEM_FLOATS = 512 + 192 + 2048 + 96 + 36;   // W0t b0 g1 be1 | W3t b3 g4 be4 | W6 b6

__device__ __forceinline__ void complexity_load_weights(const float* __restrict__ cmlp, float* w) {
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int i = tid; i < 512; i += NT) { const int m = i >> 3, k = i & 7; w[k * 64 + m] = __ldg(cmlp + i); }
  for (int i = tid; i < 192; i += NT) w[512 + i] = __ldg(cmlp + 512 + i);
  for (int i = tid; i < 2048; i += NT) { const int m = i >> 6, k = i & 63; w[704 + k * 32 + m] = __ldg(cmlp + 704 + i); }
  for (int i = tid; i < 96; i += NT) w[2752 + i] = __ldg(cmlp + 2752 + i);
  for (int i = tid; i < 33; i += NT) w[2848 + i] = __ldg(cmlp + 2848 + i);
}

// Tiles are processed in batches of NET_TB; within a batch every (tile, unit) output is one
// thread's FMA chain, so the dense layers run at CTA width instead of one warp per tile.
// All functions take a tile range [t_lo, t_hi): the fused kernel splits an image's tiles over the
// CTAs of a cluster and all-gathers the results through distributed shared memory.
constexpr int NET_TB = 128;
constexpr int ROW64 = 68, ROW32 = 36;                  // padded, 16-byte aligned activation rows
constexpr int CPX_ACT_FLOATS = NET_TB * (ROW64 + ROW32);   // h1 [TB][68], h2 [TB][36]

__host__ __device__ __forceinline__ int cpx_scratch_floats(int ntiles) {
  const int act = CPX_ACT_FLOATS, bil = ntiles * 25;
  return CMLP_SMEM_FLOATS + (act > bil ? act : bil);
}

// craw[t] = sigmoid(MLP(phi[t])) for t in [t_lo, t_hi).  phi: [ntiles][8] (global or shared),
// weights already in w, act: CPX_ACT_FLOATS scratch.
__device__ __forceinline__ void complexity_mlp_range(const float* phi, int t_lo, int t_hi, const float* w,
                                                     float* act, float* craw, float* __restrict__ raw_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
  const float* W0t = w;
  const float* b0 = w + 512;
  const float* g1 = w + 576;
  const float* be1 = w + 640;
  const float* W3t = w + 704;
  const float* b3 = w + 2752;
  const float* g4 = w + 2784;
  const float* be4 = w + 2816;
  const float* W6 = w + 2848;
  const float b6 = w[2880];
  float* h1 = act;                       // [TB][68]
  float* h2 = act + NET_TB * ROW64;      // [TB][36]
  for (int t0 = t_lo; t0 < t_hi; t0 += NET_TB) {
    const int nb = min(NET_TB, t_hi - t0);
    for (int o = tid; o < nb * 64; o += NT) {     // layer 1: 8 -> 64
      const int t = o >> 6, m = o & 63;
      const float* in = phi + (t0 + t) * 8;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(in[k], W0t[k * 64 + m], acc);
      h1[t * ROW64 + m] = __fadd_rn(acc, b0[m]);
    }
    __syncthreads();
    for (int t = warp; t < nb; t += nwarps) {     // LayerNorm(64) + ReLU, one warp per tile
      const float a0 = h1[t * ROW64 + lane], a1 = h1[t * ROW64 + lane + 32];
      const float mean = __fdiv_rn(warp_tree_sum(__fadd_rn(a0, a1)), 64.f);
      const float d0 = __fsub_rn(a0, mean), d1 = __fsub_rn(a1, mean);
      const float var = __fdiv_rn(warp_tree_sum(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1))), 64.f);
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-5f)));
      h1[t * ROW64 + lane] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d0, rstd), g1[lane]), be1[lane]), 0.f);
      h1[t * ROW64 + lane + 32] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d1, rstd), g1[lane + 32]), be1[lane + 32]), 0.f);
    }
    __syncthreads();
    // layer 2: 64 -> 32, register tiled (8 tiles x 1 unit per thread)
    dense_tiled<64, 32>(h1, ROW64, nb, W3t, [&](int t, int m, float acc) { h2[t * ROW32 + m] = __fadd_rn(acc, b3[m]); });
    __syncthreads();
    for (int t = warp; t < nb; t += nwarps) {     // LayerNorm(32) + ReLU
      const float a0 = h2[t * ROW32 + lane];
      const float mean = __fdiv_rn(warp_tree_sum(a0), 32.f);
      const float d0 = __fsub_rn(a0, mean);
      const float var = __fdiv_rn(warp_tree_sum(__fmul_rn(d0, d0)), 32.f);
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-5f)));
      h2[t * ROW32 + lane] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d0, rstd), g4[lane]), be4[lane]), 0.f);
    }
    __syncthreads();
    for (int t = tid; t < nb; t += NT) {          // layer 3: 32 -> 1 and sigmoid, one thread per tile
      float z = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) z = fmaf(h2[t * ROW32 + k], W6[k], z);
      const float c = sigmoid_exact(__fadd_rn(z, b6));
      craw[t0 + t] = c;
      if (raw_out) raw_out[t0 + t] = c;
    }
    __syncthreads();
  }
}

// 5x5 bilateral filter with replicate padding (morphology.py:309-354) and clamp, for tiles
// [t_lo, t_hi); craw must hold ALL tiles of the image.  wgt: 25*(t_hi-t_lo) floats of scratch.
__device__ __forceinline__ void bilateral_range(const float* craw, int ht, int wt, int t_lo, int t_hi, float* wgt,
                                                float* cfin, float* __restrict__ out) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int n = t_hi - t_lo;
  for (int o = tid; o < n * 25; o += NT) {        // range weights for all (tile, tap) pairs in parallel
    const int tl = o / 25, tap = o - tl * 25;
    const int t = t_lo + tl;
    const int y = t / wt, x = t - y * wt;
    const int ky = tap / 5, kx = tap - ky * 5;
    const int yy = min(max(y + ky - 2, 0), ht - 1), xx = min(max(x + kx - 2, 0), wt - 1);
    const float d = __fsub_rn(craw[yy * wt + xx], craw[t]);
    const float arg = __fdiv_rn(-__fmul_rn(d, d), kc::BILAT_DEN);
    wgt[o] = __fmul_rn(kc::BILAT[tap], (float)exp((double)arg));
  }
  __syncthreads();
  for (int tl = tid; tl < n; tl += NT) {          // ordered accumulation of the 25 taps
    const int t = t_lo + tl;
    const int y = t / wt, x = t - y * wt;
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

constexpr int MAPPER_SMEM_FLOATS = 192 + 2048 + 192 + 2048 + 96 + 36;   // W0 v0 | W3t v3 | W6t v6 | W9 b9

__device__ __forceinline__ void mapper_load_weights(const float* __restrict__ mp, float* w) {
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int i = tid; i < 192; i += NT) w[i] = __ldg(mp + i);
  for (int i = tid; i < 2048; i += NT) { const int m = i >> 5, k = i & 31; w[192 + k * 64 + m] = __ldg(mp + 192 + i); }
  for (int i = tid; i < 192; i += NT) w[2240 + i] = __ldg(mp + 2240 + i);
  for (int i = tid; i < 2048; i += NT) { const int m = i >> 6, k = i & 63; w[2432 + k * 32 + m] = __ldg(mp + 2432 + i); }
  for (int i = tid; i < 96; i += NT) w[4480 + i] = __ldg(mp + 4480 + i);
  for (int i = tid; i < 33; i += NT) w[4576 + i] = __ldg(mp + 4576 + i);
}

constexpr int MAP_ACT_FLOATS = NET_TB * (4 + ROW32 + ROW64 + ROW32);   // zin [TB][4], g0 [TB][36], g1 [TB][68], g2 [TB][36]

// bits_s[t] for t in [t_lo, t_hi); cmap indexed by absolute tile
__device__ __forceinline__ void mapper_mlp_range(const float* cmap, int t_lo, int t_hi, const float* w, float* act,
                                                 float temperature, int use_t, int continuous,
                                                 float lo, float hi, float* bits_s, float* __restrict__ out) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const float* W0 = w;
  const float* v0 = w + 96;
  const float* W3t = w + 192;
  const float* v3 = w + 2240;
  const float* W6t = w + 2432;
  const float* v6 = w + 4480;
  const float* W9 = w + 4576;
  float* zin = act;                      // [TB][4]
  float* g0 = zin + NET_TB * 4;          // [TB][36]
  float* g1 = g0 + NET_TB * ROW32;       // [TB][68]
  float* g2 = g1 + NET_TB * ROW64;       // [TB][36]
  for (int t0 = t_lo; t0 < t_hi; t0 += NET_TB) {
    const int nb = min(NET_TB, t_hi - t0);
    for (int t = tid; t < nb; t += NT) {          // z0 = [c, c^2, log1p(c)]  (Eq.13)
      const float c = fminf(fmaxf(cmap[t0 + t], 0.f), 1.f);
      zin[t * 4 + 0] = c;
      zin[t * 4 + 1] = __fmul_rn(c, c);
      zin[t * 4 + 2] = (float)log1p((double)c);
    }
    __syncthreads();
    for (int o = tid; o < nb * 32; o += NT) {     // 3 -> 32, BN, ReLU
      const int t = o >> 5, m = o & 31;
      float acc = __fmul_rn(zin[t * 4 + 0], W0[m * 3 + 0]);
      acc = fmaf(zin[t * 4 + 1], W0[m * 3 + 1], acc);
      acc = fmaf(zin[t * 4 + 2], W0[m * 3 + 2], acc);
      const float x = __fadd_rn(acc, v0[m]);
      g0[t * ROW32 + m] = fmaxf(__fadd_rn(__fmul_rn(x, v0[32 + m]), v0[64 + m]), 0.f);
    }
    __syncthreads();
    dense_tiled<32, 64>(g0, ROW32, nb, W3t, [&](int t, int m, float acc) {      // 32 -> 64, BN, ReLU
      const float x = __fadd_rn(acc, v3[m]);
      g1[t * ROW64 + m] = fmaxf(__fadd_rn(__fmul_rn(x, v3[64 + m]), v3[128 + m]), 0.f);
    });
    __syncthreads();
    dense_tiled<64, 32>(g1, ROW64, nb, W6t, [&](int t, int m, float acc) {      // 64 -> 32, BN, ReLU
      const float x = __fadd_rn(acc, v6[m]);
      g2[t * ROW32 + m] = fmaxf(__fadd_rn(__fmul_rn(x, v6[32 + m]), v6[64 + m]), 0.f);
    });
    __syncthreads();
    for (int t = tid; t < nb; t += NT) {          // 32 -> 1, sigmoid, Eq.17, temperature / STE
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc = fmaf(g2[t * ROW32 + k], W9[k], acc);
      const float s = sigmoid_exact(__fadd_rn(acc, W9[32]));
      const float bits = finish_bits(__fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), s)), temperature, use_t,
                                     continuous, lo, hi);
      bits_s[t0 + t] = bits;
      if (out) out[t0 + t] = bits;
    }
    __syncthreads();
  }
}

// torch.quantile(q, 'linear') on a sorted row: fp32 rank, torch.lerp formula
__device__ __forceinline__ float quantile_sorted(const float* srt, int n, float q) {
  const float rank = __fmul_rn(q, (float)(n - 1));
  const int lo = (int)floorf(rank), hi = (int)ceilf(rank);
  const float w = __fsub_rn(rank, (float)lo);
  const float a = srt[lo], bb = srt[hi];
  const float diff = __fsub_rn(bb, a);
  if (w < 0.5f) return __fadd_rn(a, __fmul_rn(w, diff));
  return __fsub_rn(bb, __fmul_rn(diff, __fsub_rn(1.0f, w)));
}

// LinearBitMapper: quantiles over ALL tiles of the image (cmap: [ntiles]); writes bits for
// [t_lo, t_hi).  srt: npow2 floats of scratch.
__device__ __forceinline__ void mapper_linear_range(const float* cmap, int ntiles, int npow2, float* srt,
                                                    int t_lo, int t_hi, float temperature, int use_t,
                                                    int continuous, float lo, float hi, float eps_spread,
                                                    float* bits_s, float* __restrict__ out) {
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int i = tid; i < npow2; i += NT) srt[i] = i < ntiles ? cmap[i] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = srt[i], bb = srt[ixj];
          const bool up = (i & k) == 0;
          if ((a > bb) == up) { srt[i] = bb; srt[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  const float qlo = quantile_sorted(srt, ntiles, 0.02f);
  const float qhi = quantile_sorted(srt, ntiles, 0.98f);
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
// learned soft mask (quantization.py:213-239), in three range-aware steps
// packed params: W0[8][2][3][3] b0[8] W2[2][8] b2[2] smooth[5][5]  (195 floats)
// ---------------------------------------------------------------------------------------------
// (a) act[t] = mean over the tile's adaptive window of sum_c|x| / C, for tile rows [ty_lo, ty_hi);
//     returns the maximum over those tiles (every thread gets it).  rows: H*Wt floats, red: 32.
__device__ __forceinline__ float softmask_act_range(const float* __restrict__ ap, int C, int H, int W, int Ht,
                                                    int Wt, int ty_lo, int ty_hi, float* rows, float* red,
                                                    float* act) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
  const float fC = (float)C;
  const float rC = __frcp_rn(fC);
  const int y_lo = (ty_lo * H) / Ht, y_hi = (ty_hi * H + Ht - 1) / Ht;
  for (int i = y_lo * Wt + tid; i < y_hi * Wt; i += NT) {
    const int y = i / Wt, j = i - y * Wt;
    const int xs = (j * W) / Wt, xe = ((j + 1) * W + Wt - 1) / Wt;
    float s = 0.f;
    for (int x0 = xs; x0 < xe; x0 += 8) {          // 8 independent loads, then the ordered sum
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (x0 + u < xe) ? __ldg(ap + y * W + x0 + u) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (x0 + u < xe) {
          const float av = fabsf(v[u]);
          const float q = (av >= 1e-30f && av < 1e27f) ? div_markstein(v[u], fC, rC) : __fdiv_rn(v[u], fC);
          s = __fadd_rn(s, q);
        }
    }
    rows[i] = s;
  }
  __syncthreads();
  float lmax = -INFINITY;
  for (int t = ty_lo * Wt + tid; t < ty_hi * Wt; t += NT) {
    const int i = t / Wt, j = t - i * Wt;
    const int ys = (i * H) / Ht, ye = ((i + 1) * H + Ht - 1) / Ht;
    const int xs = (j * W) / Wt, xe = ((j + 1) * W + Wt - 1) / Wt;
    float s = 0.f;
    for (int y = ys; y < ye; ++y) s = __fadd_rn(s, rows[y * Wt + j]);
    const float a = __fdiv_rn(s, (float)((ye - ys) * (xe - xs)));
    act[t] = a;
    lmax = fmaxf(lmax, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float amax = red[0];
  for (int w = 1; w < nwarps; ++w) amax = fmaxf(amax, red[w]);
  __syncthreads();
  return amax;
}

// (b) tile head for tiles [t_lo, t_hi): act / (amax + 1e-8), bits -> [0,1], conv3x3(2->8)+ReLU,
//     conv1x1(8->2), softmax channel 0.  bits/act hold ALL tiles; P: 196 floats (params loaded
//     here), bn/an: nt floats each of scratch.
__device__ __forceinline__ void softmask_head_range(const float* bits, const float* act, float amax, int Ht, int Wt,
                                                    int t_lo, int t_hi, const float* __restrict__ prm, float* P,
                                                    float* bn, float* an, float* mt, float* __restrict__ tiles_out) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int nt = Ht * Wt;
  for (int i = tid; i < 195; i += NT) P[i] = __ldg(prm + i);
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
  for (int t = t_lo + tid; t < t_hi; t += NT) {
    const int i = t / Wt, j = t - i * Wt;
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
    const float e0 = (float)exp((double)__fsub_rn(lg[0], mx));
    const float e1 = (float)exp((double)__fsub_rn(lg[1], mx));
    const float m = __fdiv_rn(e0, __fadd_rn(e0, e1));
    mt[t] = m;
    if (tiles_out) tiles_out[t] = m;
  }
  __syncthreads();
}

// (c) rows [h_lo, h_hi) of m = smooth5x5(nearest_upsample(mt)), replicate padding, FMA chain over
//     taps in row-major order.  mt holds ALL tiles; P holds the params (smooth kernel at +170).
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

}  // namespace mcaq
