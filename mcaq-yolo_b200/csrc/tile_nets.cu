// Tile-level stages of the hot path (tensors of ht*wt <= a few hundred values per image):
//   mcaq_complexity : phi(8) -> complexity MLP (LayerNorm) -> sigmoid -> 5x5 bilateral -> clamp
//   mcaq_bit_mapper : complexity -> bits (monotone MLP with eval BatchNorm, or linear/quantile)
//   mcaq_soft_mask  : bits + tile activity -> conv net -> softmax -> nearest upsample -> 5x5 smooth
// One CTA per image; the dense layers run one warp per tile with the layer width spread over
// lanes.  All arithmetic mirrors oracle/mcaq_oracle.py: nn.Linear / conv = FMA chain over k from
// 0 with the bias added last, everything else separately rounded, transcendentals in fp64.
#include "common.cuh"

namespace mcaq {

constexpr int TN_THREADS = 256;
constexpr int TN_WARPS = TN_THREADS / 32;

__device__ __forceinline__ float sigmoid_exact(float z) {
  const float e = (float)exp((double)(-z));
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

// LayerNorm statistics over buf[0..D) in sequential order (every lane computes the same values)
__device__ __forceinline__ void ln_stats(const float* buf, int D, float eps, float& mean, float& rstd) {
  float s = 0.f;
  for (int k = 0; k < D; ++k) s = __fadd_rn(s, buf[k]);
  mean = __fdiv_rn(s, (float)D);
  float v = 0.f;
  for (int k = 0; k < D; ++k) {
    const float d = __fsub_rn(buf[k], mean);
    v = __fadd_rn(v, __fmul_rn(d, d));
  }
  const float var = __fdiv_rn(v, (float)D);
  rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
}

// ---------------------------------------------------------------------------------------------
// complexity MLP: 8 -> 64 (LN, ReLU) -> 32 (LN, ReLU) -> 1 (sigmoid)       morphology.py:81-90
// packed params: W0[64][8] b0[64] g1[64] be1[64] W3[32][64] b3[32] g4[32] be4[32] W6[32] b6[1]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TN_THREADS)
complexity_kernel(const float* __restrict__ phi, int ht, int wt, const float* __restrict__ cmlp,
                  const float* __restrict__ consts, float* __restrict__ raw_out, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* W0t = sm;                 // [8][64]
  float* b0 = W0t + 512;           // 64
  float* g1 = b0 + 64;
  float* be1 = g1 + 64;
  float* W3t = be1 + 64;           // [64][32]
  float* b3 = W3t + 2048;          // 32
  float* g4 = b3 + 32;
  float* be4 = g4 + 32;
  float* W6 = be4 + 32;            // 32
  float* b6 = W6 + 32;             // 1 (+3 pad)
  float* bil = b6 + 4;             // 25 spatial weights + 1 range denominator (+2 pad)
  float* wbuf = bil + 28;          // per warp: 64 + 32
  const int ntiles = ht * wt;
  float* craw = wbuf + TN_WARPS * 96;   // [ntiles]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  for (int i = tid; i < 512; i += TN_THREADS) { const int m = i >> 3, k = i & 7; W0t[k * 64 + m] = cmlp[i]; }
  for (int i = tid; i < 192; i += TN_THREADS) b0[i] = cmlp[512 + i];
  for (int i = tid; i < 2048; i += TN_THREADS) { const int m = i >> 6, k = i & 63; W3t[k * 32 + m] = cmlp[704 + i]; }
  for (int i = tid; i < 96; i += TN_THREADS) b3[i] = cmlp[2752 + i];
  for (int i = tid; i < 33; i += TN_THREADS) W6[i] = cmlp[2848 + i];
  for (int i = tid; i < 25; i += TN_THREADS) bil[i] = consts[146 + i];
  if (tid == 0) bil[25] = consts[184];
  __syncthreads();

  float* h1 = wbuf + warp * 96;
  float* h2 = h1 + 64;
  const float* ph = phi + (long long)b * ntiles * 8;
  for (int t = warp; t < ntiles; t += TN_WARPS) {
    float in[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = __ldg(ph + t * 8 + k);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = lane + 32 * u;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(in[k], W0t[k * 64 + m], acc);
      h1[m] = __fadd_rn(acc, b0[m]);
    }
    __syncwarp();
    float mean, rstd;
    ln_stats(h1, 64, 1e-5f, mean, rstd);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = lane + 32 * u;
      const float d = __fsub_rn(h1[m], mean);
      const float y = __fadd_rn(__fmul_rn(__fmul_rn(d, rstd), g1[m]), be1[m]);
      h1[m] = fmaxf(y, 0.f);
    }
    __syncwarp();
    {
      float acc = 0.f;
      for (int k = 0; k < 64; ++k) acc = fmaf(h1[k], W3t[k * 32 + lane], acc);
      h2[lane] = __fadd_rn(acc, b3[lane]);
    }
    __syncwarp();
    ln_stats(h2, 32, 1e-5f, mean, rstd);
    __syncwarp();
    {
      const float d = __fsub_rn(h2[lane], mean);
      const float y = __fadd_rn(__fmul_rn(__fmul_rn(d, rstd), g4[lane]), be4[lane]);
      h2[lane] = fmaxf(y, 0.f);
    }
    __syncwarp();
    if (lane == 0) {
      float acc = 0.f;
      for (int k = 0; k < 32; ++k) acc = fmaf(h2[k], W6[k], acc);
      const float c = sigmoid_exact(__fadd_rn(acc, b6[0]));
      craw[t] = c;
      if (raw_out) raw_out[(long long)b * ntiles + t] = c;
    }
    __syncwarp();
  }
  __syncthreads();

  // 5x5 bilateral filter, replicate padding (morphology.py:309-354), then clamp to [0,1]
  const float rden = bil[25];
  for (int t = tid; t < ntiles; t += TN_THREADS) {
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
        const float arg = __fdiv_rn(-__fmul_rn(d, d), rden);
        const float rw = (float)exp((double)arg);
        const float wgt = __fmul_rn(bil[ky * 5 + kx], rw);
        num = __fadd_rn(num, __fmul_rn(wgt, p));
        den = __fadd_rn(den, wgt);
      }
    }
    float r = __fdiv_rn(num, __fadd_rn(den, 1e-8f));
    out[(long long)b * ntiles + t] = fminf(fmaxf(r, 0.f), 1.f);
  }
}

// ---------------------------------------------------------------------------------------------
// bit mappers (eval)                                              bit_allocation.py:42-80, 218-280
// packed MLP params (BatchNorm folded to y = x*alpha + beta):
//   W0[32][3] b0[32] a1[32] be1[32] W3[64][32] b3[64] a4[64] be4[64] W6[32][64] b6[32] a7[32] be7[32]
//   W9[32] b9[1]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float finish_bits(float bits, float temperature, int use_t, int continuous,
                                             float lo, float hi) {
  if (use_t) bits = __fmul_rn(bits, temperature);
  const float cl = fminf(fmaxf(bits, lo), hi);
  bits = __fadd_rn(bits, __fsub_rn(cl, bits));                 // straight-through clamp, forward value
  if (!continuous) bits = __fadd_rn(bits, __fsub_rn(rintf(bits), bits));
  return bits;
}

__global__ void __launch_bounds__(TN_THREADS)
mapper_mlp_kernel(const float* __restrict__ cmap, int ntiles, const float* __restrict__ mp, float temperature,
                  int use_t, int continuous, float lo, float hi, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* W0 = sm;                  // [32][3] as is (96)
  float* v0 = W0 + 96;             // b0 a1 be1 (96)
  float* W3t = v0 + 96;            // [32][64]
  float* v3 = W3t + 2048;          // b3 a4 be4 (192)
  float* W6t = v3 + 192;           // [64][32]
  float* v6 = W6t + 2048;          // b6 a7 be7 (96)
  float* W9 = v6 + 96;             // 32 + b9
  float* wbuf = W9 + 36;           // per warp: 32 + 64 + 32
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  for (int i = tid; i < 192; i += TN_THREADS) W0[i] = mp[i];
  for (int i = tid; i < 2048; i += TN_THREADS) { const int m = i >> 5, k = i & 31; W3t[k * 64 + m] = mp[192 + i]; }
  for (int i = tid; i < 192; i += TN_THREADS) v3[i] = mp[2240 + i];
  for (int i = tid; i < 2048; i += TN_THREADS) { const int m = i >> 6, k = i & 63; W6t[k * 32 + m] = mp[2432 + i]; }
  for (int i = tid; i < 96; i += TN_THREADS) v6[i] = mp[4480 + i];
  for (int i = tid; i < 33; i += TN_THREADS) W9[i] = mp[4576 + i];
  __syncthreads();
  float* h0 = wbuf + warp * 128;
  float* h1 = h0 + 32;
  float* h2 = h1 + 64;
  for (int t = warp; t < ntiles; t += TN_WARPS) {
    float c = __ldg(cmap + (long long)b * ntiles + t);
    c = fminf(fmaxf(c, 0.f), 1.f);
    const float z0 = c, z1 = __fmul_rn(c, c), z2 = (float)log1p((double)c);
    {
      float acc = __fmul_rn(z0, W0[lane * 3 + 0]);              // fma(z0, w, 0)
      acc = fmaf(z1, W0[lane * 3 + 1], acc);
      acc = fmaf(z2, W0[lane * 3 + 2], acc);
      const float x = __fadd_rn(acc, v0[lane]);
      h0[lane] = fmaxf(__fadd_rn(__fmul_rn(x, v0[32 + lane]), v0[64 + lane]), 0.f);
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = lane + 32 * u;
      float acc = 0.f;
      for (int k = 0; k < 32; ++k) acc = fmaf(h0[k], W3t[k * 64 + m], acc);
      const float x = __fadd_rn(acc, v3[m]);
      h1[m] = fmaxf(__fadd_rn(__fmul_rn(x, v3[64 + m]), v3[128 + m]), 0.f);
    }
    __syncwarp();
    {
      float acc = 0.f;
      for (int k = 0; k < 64; ++k) acc = fmaf(h1[k], W6t[k * 32 + lane], acc);
      const float x = __fadd_rn(acc, v6[lane]);
      h2[lane] = fmaxf(__fadd_rn(__fmul_rn(x, v6[32 + lane]), v6[64 + lane]), 0.f);
    }
    __syncwarp();
    if (lane == 0) {
      float acc = 0.f;
      for (int k = 0; k < 32; ++k) acc = fmaf(h2[k], W9[k], acc);
      const float s = sigmoid_exact(__fadd_rn(acc, W9[32]));
      const float bits = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), s));
      out[(long long)b * ntiles + t] = finish_bits(bits, temperature, use_t, continuous, lo, hi);
    }
    __syncwarp();
  }
}

// torch.quantile(q, interpolation='linear') on a sorted row: fp32 rank, torch.lerp formula
__device__ __forceinline__ float quantile_sorted(const float* srt, int n, float q) {
  const float rank = __fmul_rn(q, (float)(n - 1));
  const int lo = (int)floorf(rank), hi = (int)ceilf(rank);
  const float w = __fsub_rn(rank, (float)lo);
  const float a = srt[lo], bb = srt[hi];
  const float diff = __fsub_rn(bb, a);
  if (w < 0.5f) return __fadd_rn(a, __fmul_rn(w, diff));
  return __fsub_rn(bb, __fmul_rn(diff, __fsub_rn(1.0f, w)));
}

__global__ void __launch_bounds__(TN_THREADS)
mapper_linear_kernel(const float* __restrict__ cmap, int ntiles, int npow2, float temperature, int use_t,
                     int continuous, float lo, float hi, float eps_spread, float* __restrict__ out) {
  extern __shared__ float srt[];   // [npow2]
  const int tid = threadIdx.x, b = blockIdx.x;
  const float* c = cmap + (long long)b * ntiles;
  for (int i = tid; i < npow2; i += TN_THREADS) srt[i] = i < ntiles ? c[i] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += TN_THREADS) {
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
  for (int t = tid; t < ntiles; t += TN_THREADS) {
    const float v = c[t];
    float rel = __fdiv_rn(__fsub_rn(v, qlo), __fadd_rn(spread, 1e-8f));
    rel = fminf(fmaxf(rel, 0.f), 1.f);
    const float cn = spread > eps_spread ? rel : fminf(fmaxf(v, 0.f), 1.f);
    const float bits = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), cn));
    out[(long long)b * ntiles + t] = finish_bits(bits, temperature, use_t, continuous, lo, hi);
  }
}

// ---------------------------------------------------------------------------------------------
// learned soft mask                                                       quantization.py:213-239
// packed params: W0[8][2][3][3] b0[8] W2[2][8] b2[2] smooth[5][5]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TN_THREADS)
soft_mask_kernel(const float* __restrict__ bit_map, int Ht, int Wt, const float* __restrict__ abs_plane,
                 int C, int H, int W, const float* __restrict__ prm, float* __restrict__ tiles_out,
                 float* __restrict__ mask) {
  extern __shared__ float sm[];
  float* P = sm;                         // 195 (+1)
  float* act = P + 196;                  // [Ht*Wt]
  const int nt = Ht * Wt;
  float* bn = act + nt;                  // [Ht*Wt] normalised bits
  float* mt = bn + nt;                   // [Ht*Wt] tile mask
  float* rows = mt + nt;                 // [H*Wt]
  float* red = rows + H * Wt;            // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  for (int i = tid; i < 195; i += TN_THREADS) P[i] = prm[i];
  const float* ap = abs_plane + (long long)b * H * W;
  const float fC = (float)C;
  // per (row, window column): sum over the window's columns, left to right
  for (int i = tid; i < H * Wt; i += TN_THREADS) {
    const int y = i / Wt, j = i - y * Wt;
    const int xs = (j * W) / Wt, xe = ((j + 1) * W + Wt - 1) / Wt;
    float s = 0.f;
    for (int x = xs; x < xe; ++x) s = __fadd_rn(s, __fdiv_rn(ap[y * W + x], fC));
    rows[i] = s;
  }
  __syncthreads();
  float lmax = -INFINITY;
  for (int t = tid; t < nt; t += TN_THREADS) {
    const int i = t / Wt, j = t - i * Wt;
    const int ys = (i * H) / Ht, ye = ((i + 1) * H + Ht - 1) / Ht;
    const int xs = (j * W) / Wt, xe = ((j + 1) * W + Wt - 1) / Wt;
    float s = 0.f;
    for (int y = ys; y < ye; ++y) s = __fadd_rn(s, rows[y * Wt + j]);
    const float a = __fdiv_rn(s, (float)((ye - ys) * (xe - xs)));
    act[t] = a;
    lmax = fmaxf(lmax, a);
    const float bits = __ldg(bit_map + (long long)b * nt + t);
    bn[t] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bits, 2.0f), 6.0f), 0.f), 1.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float amax = red[0];
  for (int w = 1; w < TN_WARPS; ++w) amax = fmaxf(amax, red[w]);
  const float aden = __fadd_rn(amax, 1e-8f);
  __syncthreads();
  for (int t = tid; t < nt; t += TN_THREADS) act[t] = __fdiv_rn(act[t], aden);
  __syncthreads();
  // conv3x3(2->8, zero pad) + ReLU, conv1x1(8->2), softmax channel 0
  const float* W0 = P;            // [8][2][3][3]
  const float* b0 = P + 144;
  const float* W2 = P + 152;      // [2][8]
  const float* b2 = P + 168;
  for (int t = tid; t < nt; t += TN_THREADS) {
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
    if (tiles_out) tiles_out[(long long)b * nt + t] = m;
  }
  __syncthreads();
  // nearest upsample + replicate-padded 5x5 Gaussian (FMA chain in row-major tap order)
  const float* ks = P + 170;
  const float sy = (float)Ht / (float)H, sx = (float)Wt / (float)W;
  float* mo = mask + (long long)b * H * W;
  for (int p = tid; p < H * W; p += TN_THREADS) {
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

using namespace mcaq;

extern "C" int mcaq_complexity(const float* phi, int B, int ht, int wt, const float* cmlp, const float* consts,
                               float* complexity_raw, float* complexity, void* stream) {
  if (!phi || !cmlp || !consts || !complexity || B <= 0 || ht <= 0 || wt <= 0) return MCAQ_EINVAL;
  const size_t smem = (size_t)(512 + 192 + 2048 + 96 + 32 + 4 + 28 + TN_WARPS * 96 + ht * wt) * 4;
  if (smem > 200 * 1024) return MCAQ_ETOOBIG;
  if (smem > 48 * 1024) cudaFuncSetAttribute(complexity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  complexity_kernel<<<B, TN_THREADS, smem, (cudaStream_t)stream>>>(phi, ht, wt, cmlp, consts, complexity_raw,
                                                                  complexity);
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
    const size_t smem = (size_t)(96 + 96 + 2048 + 192 + 2048 + 96 + 36 + TN_WARPS * 128) * 4;
    mapper_mlp_kernel<<<B, TN_THREADS, smem, st>>>(complexity, ntiles, mapper, temperature, use_temperature,
                                                   continuous, min_bits, max_bits, bit_map);
  } else {
    int npow2 = 1;
    while (npow2 < ntiles) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * 4;
    if (smem > 200 * 1024) return MCAQ_ETOOBIG;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(mapper_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mapper_linear_kernel<<<B, TN_THREADS, smem, st>>>(complexity, ntiles, npow2, temperature, use_temperature,
                                                      continuous, min_bits, max_bits, eps_spread, bit_map);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_soft_mask(const float* bit_map, int Ht, int Wt, const float* abs_plane, int B, int C, int H,
                              int W, const float* softmask, float* mask_tiles, float* mask, void* stream) {
  if (!bit_map || !abs_plane || !softmask || !mask || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0)
    return MCAQ_EINVAL;
  const size_t smem = (size_t)(196 + 3 * Ht * Wt + H * Wt + 32) * 4;
  if (smem > 200 * 1024) return MCAQ_ETOOBIG;
  if (smem > 48 * 1024) cudaFuncSetAttribute(soft_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  soft_mask_kernel<<<B, TN_THREADS, smem, (cudaStream_t)stream>>>(bit_map, Ht, Wt, abs_plane, C, H, W, softmask,
                                                                 mask_tiles, mask);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
