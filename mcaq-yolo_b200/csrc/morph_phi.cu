// K2: per-image morphology on the on-chip gray plane -> phi(8) per tile.
//
// One CTA per image.  The cropped gray plane (Hc x Wc fp32) lives in shared memory together
// with one scratch plane; binary maps (adaptive mask, Canny strong/weak/edge) are 1-bit planes
// (32 pixels per word, built with warp ballots) so that hysteresis, erosion, Euler quads and
// box counting are word-parallel bit operations.  Stage list (each separated by a CTA barrier):
//
//   S0  gray = sum/C, per-image min/max            S7  Otsu: fp64 warp scan, first-argmax
//   S1  normalise to [0,1]                         S8  |Sobel(255*blur)| -> P0
//   S2  P1 = 255*gray                              S9  NMS + double threshold -> strong/weak bits
//   S3  11x11 adaptive threshold -> BIN bits       S10 8 constrained dilations (hysteresis)
//   S4  LBP histograms, Sobel(gray) row sums       S11 tile counts: edge, area, perimeter, Euler,
//   S5  phi2 (entropy), phi3 (gradient variance)       dyadic box counts
//   S6  5x5 blur -> P1, 256-bin histogram          S12 phi1, phi4, phi5, interactions -> HBM
//
// Arithmetic follows oracle/mcaq_oracle.py exactly: stencils are FMA chains over taps in
// row-major order from 0, everything else is separately rounded fp32; per-tile transcendentals
// are evaluated in fp64 and rounded once (B200 has full-rate-class FP64, and there are only a
// few thousand per image).
#include "common.cuh"

namespace mcaq {

// offsets into the constant block (MCAQ_CONSTS_FLOATS floats), see mcaq_b200/constants.py
constexpr int K_CANNY = 0;        // 25
constexpr int K_ADAPT = 25;       // 121
constexpr int K_BILAT = 146;      // 25
constexpr int K_FLOG = 171;       // 5: log(2,4,8,16,32)
constexpr int K_FW = 176;         // 5: exp(-0.1 i)
constexpr int K_RAD2DEG = 181;
constexpr int K_FOURPI = 182;
constexpr int K_LOG2_10 = 183;
constexpr int K_BILAT_DEN = 184;

constexpr int MORPH_THREADS = 512;

struct MorphGeom {
  int B, C, H, W, tile, ht, wt, Hc, Wc, WW, ntiles, S;
  // shared-memory layout (offsets in 4-byte words from the start of dynamic smem)
  int np1;          // words reserved for P1 (>= Hc*Wc)
  int off_lbp;      // 10 ints per tile; inside P1 when it fits next to rowsum, else separate
  int off_rowsum;   // Hc*wt float4, inside P1
  int off_tail;     // BIN, STRONG, WEAK, phis, hist, consts, scratch
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// word k of row r of a bit plane, 0 outside the plane
__device__ __forceinline__ uint32_t bp_get(const uint32_t* bp, int r, int k, int Hc, int WW) {
  return (r < 0 || r >= Hc || k < 0 || k >= WW) ? 0u : bp[r * WW + k];
}

// Sobel responses of plane p (scaled by `mul`, one rounding) at (r, x) with zero padding;
// accumulation order = FMA chain over the 3x3 taps in row-major order (zero taps are no-ops).
__device__ __forceinline__ void sobel_at(const float* p, int r, int x, int Hc, int Wc, float mul, bool scaled,
                                         float& gx, float& gy) {
  float v[3][3];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int rr = r + dy - 1, xx = x + dx - 1;
      float t = 0.f;
      if (rr >= 0 && rr < Hc && xx >= 0 && xx < Wc) {
        t = p[rr * Wc + xx];
        if (scaled) t = __fmul_rn(t, mul);
      }
      v[dy][dx] = t;
    }
  }
  // kx = [-1 0 1; -2 0 2; -1 0 1]
  float a = __fmul_rn(v[0][0], -1.f);
  a = fmaf(v[0][2], 1.f, a);
  a = fmaf(v[1][0], -2.f, a);
  a = fmaf(v[1][2], 2.f, a);
  a = fmaf(v[2][0], -1.f, a);
  gx = fmaf(v[2][2], 1.f, a);
  // ky = [-1 -2 -1; 0 0 0; 1 2 1]
  float c = __fmul_rn(v[0][0], -1.f);
  c = fmaf(v[0][1], -2.f, c);
  c = fmaf(v[0][2], -1.f, c);
  c = fmaf(v[2][0], 1.f, c);
  c = fmaf(v[2][1], 2.f, c);
  gy = fmaf(v[2][2], 1.f, c);
}

__global__ void __launch_bounds__(MORPH_THREADS, 1)
morph_phi_kernel(const float* __restrict__ sum_plane, MorphGeom g, const float* __restrict__ consts,
                 float* __restrict__ phi_out, float* __restrict__ gray_dbg, uint32_t* __restrict__ edge_dbg,
                 uint32_t* __restrict__ bin_dbg, int* __restrict__ lbp_dbg, int* __restrict__ counts_dbg,
                 long long* __restrict__ clk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NP = g.Hc * g.Wc;
  const int NW = g.Hc * g.WW;
  float* P0 = reinterpret_cast<float*>(smem_raw);
  float* P1 = P0 + NP;
  uint32_t* BIN = reinterpret_cast<uint32_t*>(P0 + g.off_tail);
  uint32_t* STRONG = BIN + NW;
  uint32_t* WEAK = STRONG + NW;
  float* phis = reinterpret_cast<float*>(WEAK + NW);          // [ntiles][5]
  int* hist = reinterpret_cast<int*>(phis + g.ntiles * 5);    // [256]
  float* kc = reinterpret_cast<float*>(hist + 256);           // constants copy [192]
  float* red = kc + MCAQ_CONSTS_FLOATS;                        // [64] reduction scratch
  // P1 is free during S4-S5: the tile-row sums (and the LBP histograms when they fit) live there
  int* lbp_hist = reinterpret_cast<int*>(P0 + g.off_lbp);      // [ntiles][10]
  float* rowsum = P0 + g.off_rowsum;                           // [Hc][wt][4]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = blockDim.x, nwarps = NT >> 5;
  const int b = blockIdx.x;
  const int tile = g.tile, Hc = g.Hc, Wc = g.Wc, WW = g.WW, wt = g.wt;
  const int tshift = 31 - __clz(tile);
  const float ntile2 = (float)(tile * tile);

#define STAGE_CLOCK(k) do { if (clk && tid == 0) clk[(long long)b * 16 + (k)] = clock64(); } while (0)
  STAGE_CLOCK(0);
  for (int i = tid; i < MCAQ_CONSTS_FLOATS; i += NT) kc[i] = consts[i];

  // ---- S0: gray = sum / C over the cropped plane, per-image min / max --------------------
  const float* sp = sum_plane + (long long)b * g.H * g.W;
  const float fC = (float)g.C;
  float lmin = INFINITY, lmax = -INFINITY;
  for (int i = tid; i < NP; i += NT) {
    const int r = i / Wc, x = i - r * Wc;
    const float v = __fdiv_rn(sp[r * g.W + x], fC);
    P0[i] = v;
    lmin = fminf(lmin, v);
    lmax = fmaxf(lmax, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if (lane == 0) { red[warp] = lmin; red[32 + warp] = lmax; }
  __syncthreads();
  float gmin = red[0], gmax = red[32];
  for (int w = 1; w < nwarps; ++w) { gmin = fminf(gmin, red[w]); gmax = fmaxf(gmax, red[32 + w]); }
  // ---- S1 + S2: normalise (morphology.py:378-383), P1 = 255 * gray ------------------------
  const float den = __fadd_rn(__fsub_rn(gmax, gmin), 1e-8f);
  for (int i = tid; i < NP; i += NT) {
    const float v = __fdiv_rn(__fsub_rn(P0[i], gmin), den);
    P0[i] = v;
    P1[i] = __fmul_rn(v, 255.f);
    if (gray_dbg) gray_dbg[(long long)b * NP + i] = v;
  }
  for (int i = tid; i < 256; i += NT) hist[i] = 0;
  __syncthreads();

  STAGE_CLOCK(1);
  // ---- S3: adaptive threshold, 11x11 Gaussian mean with replicate borders -----------------
  //      (morphology.py:550-573): bin = g255 > local_mean - 2
  for (int slot = warp; slot < NW; slot += nwarps) {
    const int r = slot / WW, k = slot - r * WW;
    const int x = 32 * k + lane;
    const bool valid = x < Wc;
    bool bit = false;
    if (valid) {
      int xc[11];
#pragma unroll
      for (int j = 0; j < 11; ++j) xc[j] = clampi(x + j - 5, 0, Wc - 1);
      float acc = 0.f;
      const float* wk = kc + K_ADAPT;
#pragma unroll 1
      for (int ky = 0; ky < 11; ++ky) {
        const float* row = P1 + clampi(r + ky - 5, 0, Hc - 1) * Wc;
#pragma unroll
        for (int kx = 0; kx < 11; ++kx) acc = fmaf(row[xc[kx]], wk[ky * 11 + kx], acc);
      }
      bit = P1[r * Wc + x] > __fsub_rn(acc, 2.0f);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) BIN[slot] = word;
  }
  __syncthreads();
  STAGE_CLOCK(2);
  for (int i = tid; i < g.ntiles * 10; i += NT) lbp_hist[i] = 0;
  __syncthreads();

  // ---- S4: uniform-LBP histograms + Sobel(gray) tile-row sums ------------------------------
  for (int slot = warp; slot < NW; slot += nwarps) {
    const int r = slot / WW, k = slot - r * WW;
    const int x = 32 * k + lane;
    const bool valid = x < Wc;
    float gx = 0.f, gy = 0.f;
    if (valid) {
      const float c = P0[r * Wc + x];
      const int ru = clampi(r - 1, 0, Hc - 1), rd = clampi(r + 1, 0, Hc - 1);
      const int xl = clampi(x - 1, 0, Wc - 1), xr = clampi(x + 1, 0, Wc - 1);
      // neighbour order (-1,-1),(-1,0),(-1,1),(0,1),(1,1),(1,0),(1,-1),(0,-1)  (morphology.py:634)
      uint32_t code = 0;
      code |= (P0[ru * Wc + xl] >= c) << 0;
      code |= (P0[ru * Wc + x] >= c) << 1;
      code |= (P0[ru * Wc + xr] >= c) << 2;
      code |= (P0[r * Wc + xr] >= c) << 3;
      code |= (P0[rd * Wc + xr] >= c) << 4;
      code |= (P0[rd * Wc + x] >= c) << 5;
      code |= (P0[rd * Wc + xl] >= c) << 6;
      code |= (P0[r * Wc + xl] >= c) << 7;
      const uint32_t rot = ((code << 1) | (code >> 7)) & 0xffu;
      const int trans = __popc(code ^ rot);
      const int label = trans <= 2 ? __popc(code) : 9;
      const int t = (r >> tshift) * wt + (x >> tshift);
      atomicAdd(&lbp_hist[t * 10 + label], 1);
      sobel_at(P0, r, x, Hc, Wc, 1.f, false, gx, gy);
    }
    // sequential left-to-right sum of each tile-row segment (leaders: lane % tile == 0)
    float s0 = gx, s1 = __fmul_rn(gx, gx), s2 = gy, s3 = __fmul_rn(gy, gy);
    const float q0 = s0, q1 = s1, q2 = s2, q3 = s3;
    for (int j = 1; j < tile; ++j) {
      s0 = __fadd_rn(s0, __shfl_down_sync(0xffffffffu, q0, j));
      s1 = __fadd_rn(s1, __shfl_down_sync(0xffffffffu, q1, j));
      s2 = __fadd_rn(s2, __shfl_down_sync(0xffffffffu, q2, j));
      s3 = __fadd_rn(s3, __shfl_down_sync(0xffffffffu, q3, j));
    }
    if (valid && (lane & (tile - 1)) == 0) {
      float4* dst = reinterpret_cast<float4*>(rowsum) + (r * wt + (x >> tshift));
      *dst = make_float4(s0, s1, s2, s3);
    }
  }
  __syncthreads();

  STAGE_CLOCK(3);
  // ---- S5: phi2 (LBP entropy, morphology.py:648-652) and phi3 (654-670) per tile -----------
  for (int t = tid; t < g.ntiles; t += NT) {
    const int ty = t / wt, tx = t - ty * wt;
    float ent = 0.f;
#pragma unroll
    for (int kk = 0; kk < 10; ++kk) {
      const int cnt = lbp_hist[t * 10 + kk];
      if (lbp_dbg) lbp_dbg[((long long)b * g.ntiles + t) * 10 + kk] = cnt;
      const float p = __fdiv_rn((float)cnt, ntile2);
      const float lg = (float)log2((double)__fadd_rn(p, 1e-10f));
      ent = __fadd_rn(ent, __fmul_rn(p, lg));
    }
    phis[t * 5 + 1] = __fdiv_rn(-ent, kc[K_LOG2_10]);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int y = 0; y < tile; ++y) {
      const float4 v = reinterpret_cast<const float4*>(rowsum)[(ty * tile + y) * wt + tx];
      a0 = __fadd_rn(a0, v.x); a1 = __fadd_rn(a1, v.y); a2 = __fadd_rn(a2, v.z); a3 = __fadd_rn(a3, v.w);
    }
    const float mx_ = __fdiv_rn(a0, ntile2), mx2 = __fdiv_rn(a1, ntile2);
    const float my_ = __fdiv_rn(a2, ntile2), my2 = __fdiv_rn(a3, ntile2);
    const float vx = fmaxf(__fsub_rn(mx2, __fmul_rn(mx_, mx_)), 0.f);
    const float vy = fmaxf(__fsub_rn(my2, __fmul_rn(my_, my_)), 0.f);
    const float v = __fadd_rn(vx, vy);
    phis[t * 5 + 2] = __fdiv_rn(v, __fadd_rn(v, 1.0f));
  }
  __syncthreads();

  STAGE_CLOCK(4);
  // ---- S6: 5x5 Gaussian blur (zero padding) -> P1, Otsu histogram (morphology.py:485-493) --
  for (int slot = warp; slot < NW; slot += nwarps) {
    const int r = slot / WW, k = slot - r * WW;
    const int x = 32 * k + lane;
    if (x < Wc) {
      float acc = 0.f;
      const float* wk = kc + K_CANNY;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        const int rr = r + ky - 2;
        if (rr < 0 || rr >= Hc) continue;          // fma(0, w, acc) == acc
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const int xx = x + kx - 2;
          if (xx >= 0 && xx < Wc) acc = fmaf(P0[rr * Wc + xx], wk[ky * 5 + kx], acc);
        }
      }
      P1[r * Wc + x] = acc;
      if (acc >= 0.f && acc <= 1.f) {              // torch.histc(bins=256, min=0, max=1)
        int bin = (int)__fmul_rn(acc, 256.f);
        if (bin == 256) bin = 255;
        atomicAdd(&hist[bin], 1);
      }
    }
  }
  __syncthreads();

  STAGE_CLOCK(5);
  // ---- S7: Otsu threshold (morphology.py:397-418), one warp ----------------------------------
  if (warp == 0) {
    int cnt[8];
    int tot_i = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { cnt[j] = hist[lane * 8 + j]; tot_i += cnt[j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot_i += __shfl_xor_sync(0xffffffffu, tot_i, o);
    const float tot = fmaxf((float)tot_i, 1.0f);
    float p[8], pc[8];
    double so = 0.0, sm = 0.0;                      // fp64 sums of fp32 terms are exact here
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p[j] = __fdiv_rn((float)cnt[j], tot);
      const float center = __fdiv_rn(__fadd_rn((float)(lane * 8 + j), 0.5f), 256.f);
      pc[j] = __fmul_rn(p[j], center);
      so += (double)p[j];
      sm += (double)pc[j];
    }
    double io = so, im = sm;                        // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double to = __shfl_up_sync(0xffffffffu, io, o);
      const double tm = __shfl_up_sync(0xffffffffu, im, o);
      if (lane >= o) { io += to; im += tm; }
    }
    const float mu_t = (float)__shfl_sync(0xffffffffu, im, 31);
    double ro = io - so, rm = im - sm;              // exclusive prefix
    float best = -INFINITY;
    int best_i = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ro += (double)p[j];
      rm += (double)pc[j];
      const float omega = (float)ro, mu = (float)rm;
      float num = __fsub_rn(__fmul_rn(mu_t, omega), mu);
      num = __fmul_rn(num, num);
      const float dn = __fadd_rn(__fmul_rn(omega, __fsub_rn(1.0f, omega)), 1e-12f);
      const float sig = __fdiv_rn(num, dn);
      if (sig > best) { best = sig; best_i = lane * 8 + j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    if (lane == 0) {
      const float thr = __fdiv_rn(__fadd_rn((float)best_i, 0.5f), 256.f);
      red[0] = __fmul_rn(thr, 255.f);               // thr255
      red[1] = __int_as_float(best_i);
    }
  }
  __syncthreads();
  const float thr255 = red[0];
  const int otsu_bin = __float_as_int(red[1]);
  const float thr_lo = __fmul_rn(0.5f, thr255);

  STAGE_CLOCK(6);
  // ---- S8: L1 gradient magnitude of Sobel(255 * blur) -> P0 (morphology.py:496-497) -------
  for (int i = tid; i < NP; i += NT) {
    const int r = i / Wc, x = i - r * Wc;
    float gx, gy;
    sobel_at(P1, r, x, Hc, Wc, 255.f, true, gx, gy);
    P0[i] = __fadd_rn(fabsf(gx), fabsf(gy));
  }
  __syncthreads();

  STAGE_CLOCK(7);
  // ---- S9: non-maximum suppression + double threshold (morphology.py:426-449, 500-502) ----
  const float rad2deg = kc[K_RAD2DEG];
  for (int slot = warp; slot < NW; slot += nwarps) {
    const int r = slot / WW, k = slot - r * WW;
    const int x = 32 * k + lane;
    bool st = false, wk = false;
    if (x < Wc) {
      float gx, gy;
      sobel_at(P1, r, x, Hc, Wc, 255.f, true, gx, gy);
      const float mag = P0[r * Wc + x];
      float ang = __fmul_rn(atan2f(gy, gx), rad2deg);
      if (ang < 0.f) ang = __fadd_rn(ang, 180.f);
      int dy1, dx1;
      if (ang < 22.5f || ang >= 157.5f) { dy1 = 0; dx1 = 1; }
      else if (ang < 67.5f) { dy1 = -1; dx1 = 1; }
      else if (ang < 112.5f) { dy1 = -1; dx1 = 0; }
      else { dy1 = -1; dx1 = -1; }
      const float n1 = P0[clampi(r + dy1, 0, Hc - 1) * Wc + clampi(x + dx1, 0, Wc - 1)];
      const float n2 = P0[clampi(r - dy1, 0, Hc - 1) * Wc + clampi(x - dx1, 0, Wc - 1)];
      const float nms = (mag >= n1 && mag >= n2) ? mag : 0.f;
      st = nms > thr255;
      wk = nms > thr_lo;
    }
    const uint32_t ws = __ballot_sync(0xffffffffu, st);
    const uint32_t ww = __ballot_sync(0xffffffffu, wk);
    if (lane == 0) { STRONG[slot] = ws; WEAK[slot] = ww; }
  }
  __syncthreads();

  STAGE_CLOCK(8);
  // ---- S10: hysteresis = 8 constrained 3x3 dilations (morphology.py:504-509) ---------------
  uint32_t* EA = STRONG;
  uint32_t* EB = reinterpret_cast<uint32_t*>(P1);               // P1 is dead from here on
  for (int it = 0; it < 8; ++it) {
    for (int i = tid; i < NW; i += NT) {
      const int r = i / WW, k = i - r * WW;
      uint32_t d = 0;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const uint32_t c = bp_get(EA, r + dy, k, Hc, WW);
        const uint32_t l = bp_get(EA, r + dy, k - 1, Hc, WW);
        const uint32_t rr = bp_get(EA, r + dy, k + 1, Hc, WW);
        d |= c | (c << 1) | (l >> 31) | (c >> 1) | (rr << 31);
      }
      EB[i] = EA[i] | (WEAK[i] & d);
    }
    __syncthreads();
    uint32_t* t = EA; EA = EB; EB = t;
  }
  const uint32_t* EDGE = EA;                                      // == STRONG after 8 swaps

  STAGE_CLOCK(9);
  // ---- S11: integer tile counts ------------------------------------------------------------
  // acc[t][0..8] = edge, area, perim, euler_x4, N_2, N_4, N_8, N_16, N_32   (aliases P0)
  int* acc = reinterpret_cast<int*>(P0);
  for (int i = tid; i < g.ntiles * 9; i += NT) acc[i] = 0;
  __syncthreads();
  const int segs = 32 >> tshift;                                  // tile segments per word
  const uint32_t segmask = tile == 32 ? 0xffffffffu : ((1u << tile) - 1u);
  for (int i = tid; i < NW; i += NT) {
    const int r = i / WW, k = i - r * WW;
    const int nbits = min(32, Wc - 32 * k);
    const uint32_t vmask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u);
    const uint32_t e = EDGE[i];
    const uint32_t m = BIN[i];
    // erosion: 3x3 AND, out-of-image neighbours ignored (morphology.py:726)
    uint32_t er = 0xffffffffu;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int rr = r + dy;
      if (rr < 0 || rr >= Hc) continue;
      const uint32_t c = BIN[rr * WW + k];
      const uint32_t l = k > 0 ? BIN[rr * WW + k - 1] : 0xffffffffu;
      const uint32_t rn = k + 1 < WW ? BIN[rr * WW + k + 1] : 0xffffffffu;
      uint32_t left = (c << 1) | (l >> 31);
      uint32_t right = (c >> 1) | (rn << 31);
      if (nbits < 32) right |= (1u << (nbits - 1));             // x+1 == Wc is outside
      er &= c & left & right;
    }
    const uint32_t bnd = m & ~er & vmask;
    // Euler quads (morphology.py:694-702): a=(i-1,j-1) b=(i-1,j) c=(i,j-1) d=(i,j), zero padded
    const uint32_t U = r > 0 ? BIN[(r - 1) * WW + k] : 0u;
    const uint32_t Up = (r > 0 && k > 0) ? BIN[(r - 1) * WW + k - 1] : 0u;
    const uint32_t Cp = k > 0 ? BIN[r * WW + k - 1] : 0u;
    const uint32_t qa = (U << 1) | (Up >> 31), qb = U, qc = (m << 1) | (Cp >> 31), qd = m;
    const uint32_t x1 = qa ^ qb, c1 = qa & qb, x2 = qc ^ qd, c2 = qc & qd;
    const uint32_t odd = x1 ^ x2, anyc = c1 | c2;
    const uint32_t Q1 = odd & ~anyc & vmask, Q3 = odd & anyc & vmask;
    const uint32_t QD = ((qb & qc & ~qa & ~qd) | (qa & qd & ~qb & ~qc)) & vmask;
    const int ty = r >> tshift;
    for (int s = 0; s < segs; ++s) {
      const int x0 = 32 * k + s * tile;
      if (x0 >= Wc) break;
      const int t = ty * wt + (x0 >> tshift);
      const int sh = s * tile;
      const int ne = __popc((e >> sh) & segmask);
      const int na = __popc((m >> sh) & segmask);
      const int np = __popc((bnd >> sh) & segmask);
      const int e4 = __popc((Q1 >> sh) & segmask) - __popc((Q3 >> sh) & segmask) -
                     2 * __popc((QD >> sh) & segmask);
      if (ne) atomicAdd(&acc[t * 9 + 0], ne);
      if (na) atomicAdd(&acc[t * 9 + 1], na);
      if (np) atomicAdd(&acc[t * 9 + 2], np);
      if (e4) atomicAdd(&acc[t * 9 + 3], e4);
    }
  }
  STAGE_CLOCK(10);
  // dyadic box counts (morphology.py:595-601): one thread per (tile, box row of the scale)
  for (int t = tid; t < g.ntiles; t += NT) {
    const int ty = t / wt, tx = t - ty * wt;
    const int x0 = tx * tile, k = x0 >> 5, sh = x0 & 31;
    int sidx = 0;
    for (int s = 2; s <= tile; s <<= 1, ++sidx) {
      // mask with one bit per box column: bits 0, s, 2s, ...
      uint32_t colmask = 0;
      for (int q = 0; q < tile; q += s) colmask |= (1u << q);
      int n = 0;
      for (int y0 = 0; y0 < tile; y0 += s) {
        uint32_t o = 0;
        for (int y = 0; y < s; ++y) o |= (EDGE[(ty * tile + y0 + y) * WW + k] >> sh) & segmask;
        for (int d = 1; d < s; d <<= 1) o |= o >> d;            // bit q = OR of bits q..q+s-1
        n += __popc(o & colmask);
      }
      acc[t * 9 + 4 + sidx] = n;                                 // only this thread writes it
    }
  }
  __syncthreads();

  STAGE_CLOCK(11);
  // ---- S12: phi1, phi4, phi5, interactions (morphology.py:852-864) --------------------------
  for (int t = tid; t < g.ntiles; t += NT) {
    const int* a = acc + t * 9;
    // phi1: weighted LSQ slope of log(N_s + 1) on log s (morphology.py:603-621)
    const int S = g.S;
    float y[5];
    for (int i = 0; i < S; ++i) y[i] = (float)log((double)__fadd_rn((float)a[4 + i], 1.0f));
    const float* lx = kc + K_FLOG;
    const float* lw = kc + K_FW;
    float w_sum = 0.f, sx = 0.f, sy = 0.f;
    for (int i = 0; i < S; ++i) {
      w_sum = __fadd_rn(w_sum, lw[i]);
      sx = __fadd_rn(sx, __fmul_rn(lw[i], lx[i]));
      sy = __fadd_rn(sy, __fmul_rn(lw[i], y[i]));
    }
    const float x_mean = __fdiv_rn(sx, w_sum), y_mean = __fdiv_rn(sy, w_sum);
    float cov = 0.f, var = 0.f;
    for (int i = 0; i < S; ++i) {
      const float dx = __fsub_rn(lx[i], x_mean);
      cov = __fadd_rn(cov, __fmul_rn(__fmul_rn(lw[i], dx), __fsub_rn(y[i], y_mean)));
      var = __fadd_rn(var, __fmul_rn(lw[i], __fmul_rn(dx, dx)));
    }
    float df = -__fdiv_rn(cov, __fadd_rn(var, 1e-12f));
    df = fminf(fmaxf(df, 1.0f), 2.0f);
    const float p1 = S < 2 ? 0.5f : __fdiv_rn(df, 2.0f);
    const float p2 = phis[t * 5 + 1], p3 = phis[t * 5 + 2];
    const float p4 = __fdiv_rn((float)a[0], ntile2);
    // phi5 (morphology.py:729-738)
    const float area = (float)a[1], perim = (float)a[2];
    float ic = __fdiv_rn(__fmul_rn(perim, perim), __fadd_rn(__fmul_rn(kc[K_FOURPI], area), 1e-6f));
    const float K = fmaxf(rintf(__fdiv_rn((float)a[3], 4.0f)), 1.0f);
    ic = __fdiv_rn(ic, K);
    float p5 = __fsub_rn(1.0f, __fdiv_rn(1.0f, fmaxf(ic, 1.0f)));
    if (a[1] <= 0) p5 = 0.f;
    float* o = phi_out + ((long long)b * g.ntiles + t) * 8;
    o[0] = p1; o[1] = p2; o[2] = p3; o[3] = p4; o[4] = p5;
    o[5] = __fmul_rn(p1, p2);
    o[6] = __fmul_rn(p3, p3);
    o[7] = __fsqrt_rn(__fadd_rn(__fmul_rn(p4, p5), 1e-12f));
    if (counts_dbg) {
      int* c = counts_dbg + ((long long)b * g.ntiles + t) * 12;
      for (int i = 0; i < 9; ++i) c[i] = a[i];
      c[9] = otsu_bin; c[10] = 0; c[11] = 0;
    }
  }
  STAGE_CLOCK(12);
  if (edge_dbg)
    for (int i = tid; i < NW; i += NT) edge_dbg[(long long)b * NW + i] = EDGE[i];
  if (bin_dbg)
    for (int i = tid; i < NW; i += NT) bin_dbg[(long long)b * NW + i] = BIN[i];
}

}  // namespace mcaq

using namespace mcaq;

static long long* g_stage_clk = nullptr;
// debug: device buffer of 16 clock64() stamps per image (NULL disables); see tools/stage_clocks.py
extern "C" void mcaq_debug_stage_clocks(long long* dev_buf) { g_stage_clk = dev_buf; }

extern "C" int mcaq_morph_phi(const float* sum_plane, int B, int C, int H, int W, int grid_size,
                              const float* consts, float* phi, float* gray_dbg, uint32_t* edge_bits_dbg,
                              uint32_t* bin_bits_dbg, int32_t* lbp_hist_dbg, int32_t* counts_dbg, void* stream) {
  if (!sum_plane || !consts || !phi || B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0) return MCAQ_EINVAL;
  MorphGeom g;
  g.B = B; g.C = C; g.H = H; g.W = W;
  g.tile = mcaq_tile_size(H, grid_size);
  g.ht = H / g.tile; g.wt = W / g.tile;
  if (g.ht <= 0 || g.wt <= 0 || g.tile > 32) return MCAQ_EINVAL;
  g.Hc = g.ht * g.tile; g.Wc = g.wt * g.tile;
  g.WW = (g.Wc + 31) / 32;
  g.ntiles = g.ht * g.wt;
  g.S = 0;
  for (int s = 2; s <= g.tile; s <<= 1) g.S++;
  const size_t NP = (size_t)g.Hc * g.Wc, NW = (size_t)g.Hc * g.WW;
  // word offsets: [P0: NP][P1: np1][lbp if separate][BIN NW][STRONG NW][WEAK NW][phis][hist][consts][red]
  const size_t rowsum_w = (size_t)g.Hc * g.wt * 4, lbp_w = (size_t)g.ntiles * 10;
  size_t np1 = NP < rowsum_w ? rowsum_w : NP;
  np1 = (np1 + 3) & ~(size_t)3;
  const size_t np0 = (NP + 3) & ~(size_t)3;       // keeps P1 / rowsum 16-byte aligned
  if (np0 != NP) return MCAQ_ETOOBIG;             // Hc*Wc is a multiple of 16 (tile >= 4): cannot happen
  g.off_rowsum = (int)np0;
  size_t tail = np0 + np1;
  if (rowsum_w + lbp_w <= np1) {
    g.off_lbp = (int)(np0 + rowsum_w);
  } else {
    g.off_lbp = (int)tail;
    tail += (lbp_w + 3) & ~(size_t)3;
  }
  g.np1 = (int)np1;
  g.off_tail = (int)tail;
  const size_t smem = (tail + 3 * NW + (size_t)g.ntiles * 5 + 256 + MCAQ_CONSTS_FLOATS + 64) * 4;
  if (smem > 227 * 1024) return MCAQ_ETOOBIG;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(morph_phi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int threads = MORPH_THREADS;
  if (NP <= 1024) threads = 128;
  else if (NP <= 4096) threads = 256;
  morph_phi_kernel<<<B, threads, smem, st>>>(sum_plane, g, consts, phi, gray_dbg, edge_bits_dbg, bin_bits_dbg,
                                             lbp_hist_dbg, counts_dbg, g_stage_clk);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
