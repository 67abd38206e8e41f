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

// ---------------------------------------------------------------------------------------------
// complexity MLP 8 -> 64 (LN, ReLU) -> 32 (LN, ReLU) -> 1, sigmoid; 5x5 bilateral; clamp
// smem need: CMLP_SMEM_FLOATS + nwarps*96 + 2*ntiles
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

// phi: [ntiles][8] (global or shared).  craw/cfin: [ntiles] shared.  Weights already loaded into w.
__device__ __forceinline__ void complexity_block(const float* phi, int ht, int wt, const float* w, float* wbuf,
                                                 float* craw, float* cfin, float* __restrict__ raw_out,
                                                 float* __restrict__ out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
  const int ntiles = ht * wt;
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
  float* h1 = wbuf + warp * 96;
  float* h2 = h1 + 64;
  for (int t = warp; t < ntiles; t += nwarps) {
    float in[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = phi[t * 8 + k];
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fmaf(in[k], W0t[k * 64 + lane], a0);
      a1 = fmaf(in[k], W0t[k * 64 + lane + 32], a1);
    }
    a0 = __fadd_rn(a0, b0[lane]);
    a1 = __fadd_rn(a1, b0[lane + 32]);
    // LayerNorm(64), eps 1e-5
    float mean = __fdiv_rn(warp_tree_sum(__fadd_rn(a0, a1)), 64.f);
    float d0 = __fsub_rn(a0, mean), d1 = __fsub_rn(a1, mean);
    float var = __fdiv_rn(warp_tree_sum(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1))), 64.f);
    float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-5f)));
    h1[lane] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d0, rstd), g1[lane]), be1[lane]), 0.f);
    h1[lane + 32] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d1, rstd), g1[lane + 32]), be1[lane + 32]), 0.f);
    __syncwarp();
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) acc = fmaf(h1[k], W3t[k * 32 + lane], acc);
    acc = __fadd_rn(acc, b3[lane]);
    mean = __fdiv_rn(warp_tree_sum(acc), 32.f);
    d0 = __fsub_rn(acc, mean);
    var = __fdiv_rn(warp_tree_sum(__fmul_rn(d0, d0)), 32.f);
    rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-5f)));
    h2[lane] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(d0, rstd), g4[lane]), be4[lane]), 0.f);
    __syncwarp();
    if (lane == 0) {
      float z = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) z = fmaf(h2[k], W6[k], z);
      craw[t] = __fadd_rn(z, b6);                 // pre-sigmoid; finished below, one thread per tile
    }
    __syncwarp();
  }
  __syncthreads();
  for (int t = tid; t < ntiles; t += NT) {
    const float c = sigmoid_exact(craw[t]);
    craw[t] = c;
    if (raw_out) raw_out[t] = c;
  }
  __syncthreads();
  // 5x5 bilateral filter, replicate padding (morphology.py:309-354), then clamp to [0,1]
  for (int t = tid; t < ntiles; t += NT) {
    const int y = t / wt, x = t - y * wt;
    const float c = craw[t];
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int yy = min(max(y + ky - 2, 0), ht - 1);
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int xx = min(max(x + kx - 2, 0), wt - 1);
        const float p = craw[yy * wt + xx];
        const float d = __fsub_rn(p, c);
        const float arg = __fdiv_rn(-__fmul_rn(d, d), kc::BILAT_DEN);
        const float rw = (float)exp((double)arg);
        const float wgt = __fmul_rn(kc::BILAT[ky * 5 + kx], rw);
        num = __fadd_rn(num, __fmul_rn(wgt, p));
        den = __fadd_rn(den, wgt);
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

// cmap: [ntiles] (global or shared); zbuf: [ntiles] shared scratch; bits_s: [ntiles] shared result
// wbuf: nwarps*128 floats
__device__ __forceinline__ void mapper_mlp_block(const float* cmap, int ntiles, const float* w, float* wbuf,
                                                 float* zbuf, float temperature, int use_t, int continuous,
                                                 float lo, float hi, float* bits_s, float* __restrict__ out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
  const float* W0 = w;
  const float* v0 = w + 96;
  const float* W3t = w + 192;
  const float* v3 = w + 2240;
  const float* W6t = w + 2432;
  const float* v6 = w + 4480;
  const float* W9 = w + 4576;
  // z0 = [c, c^2, log1p(c)] per tile, one thread per tile (fp64 log1p)
  for (int t = tid; t < ntiles; t += NT) {
    const float c = fminf(fmaxf(cmap[t], 0.f), 1.f);
    zbuf[t] = (float)log1p((double)c);
  }
  __syncthreads();
  float* h0 = wbuf + warp * 128;
  float* h1 = h0 + 32;
  float* h2 = h1 + 64;
  for (int t = warp; t < ntiles; t += nwarps) {
    const float c = fminf(fmaxf(cmap[t], 0.f), 1.f);
    const float z1 = __fmul_rn(c, c), z2 = zbuf[t];
    {
      float acc = __fmul_rn(c, W0[lane * 3 + 0]);
      acc = fmaf(z1, W0[lane * 3 + 1], acc);
      acc = fmaf(z2, W0[lane * 3 + 2], acc);
      const float x = __fadd_rn(acc, v0[lane]);
      h0[lane] = fmaxf(__fadd_rn(__fmul_rn(x, v0[32 + lane]), v0[64 + lane]), 0.f);
    }
    __syncwarp();
    {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float hv = h0[k];
        a0 = fmaf(hv, W3t[k * 64 + lane], a0);
        a1 = fmaf(hv, W3t[k * 64 + lane + 32], a1);
      }
      const float x0 = __fadd_rn(a0, v3[lane]), x1 = __fadd_rn(a1, v3[lane + 32]);
      h1[lane] = fmaxf(__fadd_rn(__fmul_rn(x0, v3[64 + lane]), v3[128 + lane]), 0.f);
      h1[lane + 32] = fmaxf(__fadd_rn(__fmul_rn(x1, v3[96 + lane]), v3[160 + lane]), 0.f);
    }
    __syncwarp();
    {
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < 64; ++k) acc = fmaf(h1[k], W6t[k * 32 + lane], acc);
      const float x = __fadd_rn(acc, v6[lane]);
      h2[lane] = fmaxf(__fadd_rn(__fmul_rn(x, v6[32 + lane]), v6[64 + lane]), 0.f);
    }
    __syncwarp();
    if (lane == 0) {
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc = fmaf(h2[k], W9[k], acc);
      bits_s[t] = __fadd_rn(acc, W9[32]);          // logit; finished below
    }
    __syncwarp();
  }
  __syncthreads();
  for (int t = tid; t < ntiles; t += NT) {
    const float s = sigmoid_exact(bits_s[t]);
    const float bits = finish_bits(__fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), s)), temperature, use_t,
                                   continuous, lo, hi);
    bits_s[t] = bits;
    if (out) out[t] = bits;
  }
  __syncthreads();
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

// cmap: [ntiles]; srt: [npow2] shared scratch; bits_s: [ntiles] shared
__device__ __forceinline__ void mapper_linear_block(const float* cmap, int ntiles, int npow2, float* srt,
                                                    float temperature, int use_t, int continuous, float lo,
                                                    float hi, float eps_spread, float* bits_s,
                                                    float* __restrict__ out) {
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
  for (int t = tid; t < ntiles; t += NT) {
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
// packed params: W0[8][2][3][3] b0[8] W2[2][8] b2[2] smooth[5][5]  (195 floats)
// smem need: 196 + 3*Ht*Wt + H*Wt + 32
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void soft_mask_block(const float* bits /*[Ht*Wt], global or shared*/, int Ht, int Wt,
                                                const float* __restrict__ ap /*[H*W] sum_c|x| of this image*/,
                                                int C, int H, int W, const float* __restrict__ prm, float* sm,
                                                float* __restrict__ tiles_out, float* __restrict__ mo) {
  float* P = sm;                         // 196
  const int nt = Ht * Wt;
  float* act = P + 196;                  // [nt]
  float* bn = act + nt;                  // [nt]
  float* mt = bn + nt;                   // [nt]
  float* rows = mt + nt;                 // [H*Wt]
  float* red = rows + H * Wt;            // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
  for (int i = tid; i < 195; i += NT) P[i] = __ldg(prm + i);
  const float fC = (float)C;
  for (int i = tid; i < H * Wt; i += NT) {
    const int y = i / Wt, j = i - y * Wt;
    const int xs = (j * W) / Wt, xe = ((j + 1) * W + Wt - 1) / Wt;
    float s = 0.f;
    for (int x = xs; x < xe; ++x) s = __fadd_rn(s, __fdiv_rn(__ldg(ap + y * W + x), fC));
    rows[i] = s;
  }
  __syncthreads();
  float lmax = -INFINITY;
  for (int t = tid; t < nt; t += NT) {
    const int i = t / Wt, j = t - i * Wt;
    const int ys = (i * H) / Ht, ye = ((i + 1) * H + Ht - 1) / Ht;
    const int xs = (j * W) / Wt, xe = ((j + 1) * W + Wt - 1) / Wt;
    float s = 0.f;
    for (int y = ys; y < ye; ++y) s = __fadd_rn(s, rows[y * Wt + j]);
    const float a = __fdiv_rn(s, (float)((ye - ys) * (xe - xs)));
    act[t] = a;
    lmax = fmaxf(lmax, a);
    bn[t] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bits[t], 2.0f), 6.0f), 0.f), 1.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float amax = red[0];
  for (int w = 1; w < nwarps; ++w) amax = fmaxf(amax, red[w]);
  const float aden = __fadd_rn(amax, 1e-8f);
  __syncthreads();
  for (int t = tid; t < nt; t += NT) act[t] = __fdiv_rn(act[t], aden);
  __syncthreads();
  const float* W0 = P;
  const float* b0 = P + 144;
  const float* W2 = P + 152;
  const float* b2 = P + 168;
  for (int t = tid; t < nt; t += NT) {
    const int i = t / Wt, j = t - i * Wt;
    float hid[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int ic = 0; ic < 2; ++ic) {
        const float* src = ic == 0 ? bn : act;
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
  const float* ks = P + 170;
  const float sy = (float)Ht / (float)H, sx = (float)Wt / (float)W;
  for (int p = tid; p < H * W; p += NT) {
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
  __syncthreads();
}

}  // namespace mcaq
