// K2: per-image morphology on the on-chip gray plane -> phi(8) per tile, optionally followed in
// the same launch by the complexity MLP + bilateral filter, the bit mapper and the soft mask
// (the whole "between the two HBM sweeps" part of the hook in one kernel).
//
// One CTA per image.  The cropped gray plane (Hc x Wc fp32) lives in shared memory with one
// scratch plane; binary maps (adaptive mask, Canny strong/weak/edge) are 1-bit planes (32 pixels
// per word, built with warp ballots) so hysteresis, erosion, Euler quads and box counting are
// word-parallel bit operations.  Stages (CTA barrier between each):
//
//   S0  gray = sum/C, per-image min/max            S7  Otsu (fp64 warp scan, first argmax) on warp 0
//   S1  normalise to [0,1], P1 = 255*gray              while the other warps do S8
//   S3  11x11 adaptive threshold -> BIN bits       S8  |Sobel(255*blur)| -> P0
//   S4  LBP histograms, Sobel(gray) row sums       S9  NMS + double threshold -> strong/weak bits
//   S5  phi2 (entropy), phi3 (gradient variance)   S10 8 constrained dilations (hysteresis)
//   S6  5x5 blur -> P1, 256-bin histogram          S11 tile counts: edge, area, perimeter, Euler, boxes
//                                                  S12 phi1, phi4, phi5, interactions
//   N1  complexity MLP + bilateral   N2  bit mapper   N3  soft mask tiles + full-resolution m
//
// The 11x11 and 5x5 stencils are register-tiled (4 output rows per thread share their input
// rows) with the taps as FFMA immediates (mcaq_consts.cuh), accumulation order per output =
// row-major FMA chain from 0, exactly the oracle's.  Arithmetic is otherwise separately rounded
// fp32; log tables are fp64-rounded literals; Otsu sums are exact in fp64.
#include <cooperative_groups.h>

#include "tile_nets.cuh"

namespace cg = cooperative_groups;

namespace mcaq {

constexpr int MORPH_THREADS = 512;
constexpr int RT = 4;    // output rows per thread in the stencil stages

struct MorphGeom {
  int B, C, H, W, tile, ht, wt, Hc, Wc, WW, ntiles, S;
  // shared-memory layout, offsets in 4-byte words
  int off_lbp;      // 10 ints per tile (inside P1 when it fits next to rowsum, else separate)
  int off_rowsum;   // Hc*wt float4, inside P1
  int off_tail;     // BIN, STRONG, WEAK, phis5, phi8, cfin, bits, craw, act, mt, hist, hloc, red, mmx, luts
  int ns;           // CTAs per image (thread-block cluster size): the image's tile rows are split
};

struct FusedArgs {
  MorphGeom g;
  const float* sum_plane;
  const float* abs_plane;      // for the soft mask (NULL: no mask stage)
  int* keys;                   // K1 range keys: decoded into `packed` and re-armed by CTA 0 (NULL: skip)
  float* packed;
  const float* cmlp;           // NULL: stop after phi
  const float* mapper;         // packed MLP mapper, or NULL with linear_mapper != 0
  const float* softmask;
  int run_mapper, linear_mapper, use_t, continuous;
  float temperature, lo, hi, eps_spread;
  float* phi;                  // (B, ntiles, 8) or NULL
  float* complexity_raw;
  float* complexity;
  float* bit_map;
  float* mask_tiles;
  float* mask;                 // (B, H, W)
  float* gray_dbg;
  uint32_t* edge_dbg;
  uint32_t* bin_dbg;
  int* lbp_dbg;
  int* counts_dbg;
  long long* clk;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ uint32_t bp_get(const uint32_t* bp, int r, int k, int Hc, int WW) {
  return (r < 0 || r >= Hc || k < 0 || k >= WW) ? 0u : bp[r * WW + k];
}

// Sobel responses at (r, x) with zero padding; FMA chain over the 3x3 taps in row-major order.
template <bool SCALED>
__device__ __forceinline__ void sobel_at(const float* p, int r, int x, int Hc, int Wc, float& gx, float& gy) {
  float v[3][3];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int rr = r + dy - 1;
    const bool rok = rr >= 0 && rr < Hc;
    const float* row = p + clampi(rr, 0, Hc - 1) * Wc;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      if (dy == 1 && dx == 1) { v[1][1] = 0.f; continue; }
      const int xx = x + dx - 1;
      float t = row[clampi(xx, 0, Wc - 1)];
      if (SCALED) t = __fmul_rn(t, 255.f);
      v[dy][dx] = (rok && xx >= 0 && xx < Wc) ? t : 0.f;
    }
  }
  float a = __fmul_rn(v[0][0], -1.f);
  a = fmaf(v[0][2], 1.f, a);
  a = fmaf(v[1][0], -2.f, a);
  a = fmaf(v[1][2], 2.f, a);
  a = fmaf(v[2][0], -1.f, a);
  gx = fmaf(v[2][2], 1.f, a);
  float c = __fmul_rn(v[0][0], -1.f);
  c = fmaf(v[0][1], -2.f, c);
  c = fmaf(v[0][2], -1.f, c);
  c = fmaf(v[2][0], 1.f, c);
  c = fmaf(v[2][1], 2.f, c);
  gy = fmaf(v[2][2], 1.f, c);
}

// RN(x / d) with a precomputed rinv = RN(1/d): Markstein's correction in the normal range (swept
// against div.rn in tests/test_gpu_division.py), plain division for zero / tiny / huge numerators.
__device__ __forceinline__ float div_exact(float x, float d, float rinv) {
  const float ax = fabsf(x);
  if (ax >= 1e-30f && ax < 1e27f) return div_markstein(x, d, rinv);
  return __fdiv_rn(x, d);
}

// direction bin of the NMS (morphology.py:430-444).  Slope tests with a 1e-5 relative guard band
// decide all but boundary cases; those take the literal atan2f path.
__device__ __forceinline__ int nms_bin(float gx, float gy) {
  const float ax = fabsf(gx), ay = fabsf(gy);
  const float a = 0.41421356f * ax;          // tan(22.5 deg)
  const float b = 2.41421356f * ax;          // tan(67.5 deg)
  if (ay < a * 0.99999f) return 0;
  if (ay > a * 1.00001f && ay < b * 0.99999f) return ((gx > 0.f) == (gy > 0.f)) ? 1 : 3;
  if (ay > b * 1.00001f) return 2;
  float ang = __fmul_rn(atan2f(gy, gx), kc::RAD2DEG);
  if (ang < 0.f) ang = __fadd_rn(ang, 180.f);
  if (ang < 22.5f || ang >= 157.5f) return 0;
  if (ang < 67.5f) return 1;
  if (ang < 112.5f) return 2;
  return 3;
}

__global__ void __launch_bounds__(MORPH_THREADS, 1)
morph_fused_kernel(const FusedArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const MorphGeom& g = A.g;
  const int NP = g.Hc * g.Wc;
  const int NW = g.Hc * g.WW;
  float* P0 = reinterpret_cast<float*>(smem_raw);
  float* P1 = P0 + NP;
  uint32_t* BIN = reinterpret_cast<uint32_t*>(P0 + g.off_tail);
  uint32_t* STRONG = BIN + NW;
  uint32_t* WEAK = STRONG + NW;
  float* phis = reinterpret_cast<float*>(WEAK + NW);          // [ntiles][5]  (phi2, phi3 parked here)
  float* phi8 = phis + g.ntiles * 5;                          // [ntiles][8]
  float* cfin = phi8 + g.ntiles * 8;                          // [ntiles] complexity
  float* bits_s = cfin + g.ntiles;                            // [ntiles] bits
  float* craw_s = bits_s + g.ntiles;                          // [ntiles] complexity before the bilateral
  float* act_s = craw_s + g.ntiles;                           // [ntiles] tile activity (soft mask)
  float* mt_s = act_s + g.ntiles;                             // [ntiles] tile mask
  int* hist = reinterpret_cast<int*>(mt_s + g.ntiles);        // [256] whole-image Otsu histogram
  int* hloc = hist + 256;                                     // [256] this CTA's band
  float* red = reinterpret_cast<float*>(hloc + 256);          // [64]
  float* mmx = red + 64;                                      // [16] per-rank min/max | per-rank act max
  float* lutn = mmx + 16;                                     // [260] log(N + 1)
  float* lutp = lutn + 260;                                   // [tile^2 + 1] log2(k / tile^2 + 1e-10)
  int* lbp_hist = reinterpret_cast<int*>(P0 + g.off_lbp);     // [ntiles][10]
  float* rowsum = P0 + g.off_rowsum;                          // [Hc][wt][4]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = blockDim.x, nwarps = NT >> 5;
  const int tile = g.tile, Hc = g.Hc, Wc = g.Wc, WW = g.WW, wt = g.wt;
  // an image is split over the ns CTAs of a thread-block cluster by tile rows; every CTA keeps
  // full-size planes and all-gathers the bands the others produced through DSMEM
  cg::cluster_group cl = cg::this_cluster();
  const int ns = g.ns;
  const int rank = ns > 1 ? (int)cl.block_rank() : 0;
  const int b = blockIdx.x / ns;
  const int tr0 = (g.ht * rank) / ns, tr1 = (g.ht * (rank + 1)) / ns;   // own tile rows
  const int r_lo = tr0 * tile, r_hi = tr1 * tile;                       // own pixel rows
  const int t_lo = tr0 * wt, t_hi = tr1 * wt;                           // own tiles
  auto csync = [&]() { if (ns > 1) cl.sync(); else __syncthreads(); };
  auto publish = [&](auto* base, int off, int n) {                      // my [off, off+n) -> every peer
    for (int pr = 1; pr < ns; ++pr) {
      auto* dst = cl.map_shared_rank(base, (rank + pr) % ns);
      for (int i = tid; i < n; i += NT) dst[off + i] = base[off + i];
    }
  };
  const int tshift = 31 - __clz(tile);
  const float ntile2 = (float)(tile * tile);
  long long* clk = A.clk;
#define STAGE_CLOCK(k) do { if (clk && tid == 0 && rank == 0) clk[(long long)b * 16 + (k)] = clock64(); } while (0)
  STAGE_CLOCK(0);
  {
    const float* src = g.tile == 4 ? kc::LOG2P_4 : (g.tile == 8 ? kc::LOG2P_8 : (g.tile == 16 ? kc::LOG2P_16 : kc::LOG2P_32));
    for (int i = tid; i < 257; i += NT) lutn[i] = __ldg(kc::LOGN1 + i);
    for (int i = tid; i <= g.tile * g.tile; i += NT) lutp[i] = __ldg(src + i);
  }

  // K1 -> K3 hand-off of the per-channel ranges: decode the atomics' integer keys to floats and
  // re-arm the keys for the next sweep (stream order: K1 done, K3 not started)
  if (A.keys && blockIdx.x == 0) {
    for (int c = tid; c < g.C; c += NT) {
      A.packed[c] = key_float(A.keys[c]);
      A.packed[g.C + c] = -key_float(A.keys[g.C + c]);
      A.keys[c] = MCAQ_KEY_POS_INF;
      A.keys[g.C + c] = MCAQ_KEY_NEG_INF;
    }
  }

  // ---- S0: gray = sum / C over the cropped plane, per-image min / max --------------------
  const float* sp = A.sum_plane + (long long)b * g.H * g.W;
  const float fC = (float)g.C;
  const float rC = __frcp_rn(fC);
  float lmin = INFINITY, lmax = -INFINITY;
  for (int i = tid; i < 512; i += NT) hist[i] = 0;               // hist + hloc
  if (ns > 1) cl.sync();      // every CTA of the cluster is resident before the first DSMEM store
  const int band_lo = r_lo * Wc, band_hi = r_hi * Wc;
  for (int i0 = band_lo + tid; i0 < band_hi; i0 += 4 * NT) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * NT;
      if (i < band_hi) { const int r = i / Wc, x = i - r * Wc; v[u] = __ldg(sp + r * g.W + x); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * NT;
      if (i < band_hi) {
        const float q = div_exact(v[u], fC, rC);
        P0[i] = q;
        lmin = fminf(lmin, q);
        lmax = fmaxf(lmax, q);
      }
    }
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
  if (ns > 1) {                                                   // all-gather raw gray bands and band min/max
    if (tid == 0)
      for (int pr = 0; pr < ns; ++pr) {
        float* m = cl.map_shared_rank(mmx, pr);
        m[2 * rank] = gmin;
        m[2 * rank + 1] = gmax;
      }
    publish(P0, band_lo, band_hi - band_lo);
    cl.sync();
    gmin = mmx[0]; gmax = mmx[1];
    for (int pr = 1; pr < ns; ++pr) { gmin = fminf(gmin, mmx[2 * pr]); gmax = fmaxf(gmax, mmx[2 * pr + 1]); }
  }
  // ---- S1: normalise (morphology.py:378-383), P1 = 255 * gray ------------------------------
  const float den = __fadd_rn(__fsub_rn(gmax, gmin), 1e-8f);
  const float rden = __frcp_rn(den);
  for (int i = tid; i < NP; i += NT) {
    const float v = div_exact(__fsub_rn(P0[i], gmin), den, rden);
    P0[i] = v;
    P1[i] = __fmul_rn(v, 255.f);
    if (A.gray_dbg && rank == 0) A.gray_dbg[(long long)b * NP + i] = v;
  }
  __syncthreads();
  STAGE_CLOCK(1);

  // ---- S3: adaptive threshold, 11x11 Gaussian mean, replicate borders (morphology.py:550-573)
  //      register tile: RT output rows share their 14 input rows; taps are immediates
  const int nrg = (r_hi - r_lo) / RT;                            // band rows % 4 == 0 (tile >= 4)
  for (int task = warp; task < nrg * WW; task += nwarps) {
    const int rg = task / WW, k = task - rg * WW;
    const int r0 = r_lo + rg * RT;
    const int x = 32 * k + lane;
    const bool valid = x < Wc;
    int xc[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) xc[j] = clampi(x + j - 5, 0, Wc - 1);
    float acc[RT];
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[j] = 0.f;
#pragma unroll
    for (int rr = 0; rr < RT + 10; ++rr) {
      const float* row = P1 + clampi(r0 - 5 + rr, 0, Hc - 1) * Wc;
      float v[11];
#pragma unroll
      for (int kx = 0; kx < 11; ++kx) v[kx] = row[xc[kx]];
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int ky = rr - j;
        if (ky >= 0 && ky < 11) {
#pragma unroll
          for (int kx = 0; kx < 11; ++kx) acc[j] = fmaf(v[kx], kc::ADAPT[ky * 11 + kx], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      const bool bit = valid && (P1[(r0 + j) * Wc + min(x, Wc - 1)] > __fsub_rn(acc[j], 2.0f));
      const uint32_t word = __ballot_sync(0xffffffffu, bit);
      if (lane == 0) BIN[(r0 + j) * WW + k] = word;
    }
  }
  __syncthreads();
  if (ns > 1) publish(BIN, r_lo * WW, (r_hi - r_lo) * WW);        // visible to the peers after the next cluster sync
  STAGE_CLOCK(2);
  for (int i = tid; i < g.ntiles * 10; i += NT) lbp_hist[i] = 0;
  __syncthreads();

  // ---- S4: uniform-LBP histograms + Sobel(gray) tile-row sums ------------------------------
  //      one 3x3 neighbourhood load serves both: LBP sees replicate borders, Sobel zero borders
  const int nslot_b = (r_hi - r_lo) * WW;
  for (int slot = warp; slot < nslot_b; slot += nwarps) {
    const int r = r_lo + slot / WW, k = slot % WW;
    const int x = 32 * k + lane;
    const bool valid = x < Wc;
    float gx = 0.f, gy = 0.f;
    int key = -1;
    if (valid) {
      const bool up = r > 0, dn = r + 1 < Hc, lf = x > 0, rt = x + 1 < Wc;
      const float* r0p = P0 + (up ? r - 1 : r) * Wc;
      const float* r1p = P0 + r * Wc;
      const float* r2p = P0 + (dn ? r + 1 : r) * Wc;
      const int xl = lf ? x - 1 : x, xr = rt ? x + 1 : x;
      const float v00 = r0p[xl], v01 = r0p[x], v02 = r0p[xr];
      const float v10 = r1p[xl], c = r1p[x], v12 = r1p[xr];
      const float v20 = r2p[xl], v21 = r2p[x], v22 = r2p[xr];
      // neighbour order (-1,-1),(-1,0),(-1,1),(0,1),(1,1),(1,0),(1,-1),(0,-1)  (morphology.py:634)
      uint32_t code = (uint32_t)(v00 >= c) | ((uint32_t)(v01 >= c) << 1) | ((uint32_t)(v02 >= c) << 2) |
                      ((uint32_t)(v12 >= c) << 3) | ((uint32_t)(v22 >= c) << 4) | ((uint32_t)(v21 >= c) << 5) |
                      ((uint32_t)(v20 >= c) << 6) | ((uint32_t)(v10 >= c) << 7);
      const uint32_t rot = ((code << 1) | (code >> 7)) & 0xffu;
      const int label = __popc(code ^ rot) <= 2 ? __popc(code) : 9;
      key = ((r >> tshift) * wt + (x >> tshift)) * 10 + label;
      // Sobel with zero padding: FMA chain over taps in row-major order (zero taps are no-ops)
      const float z00 = (up && lf) ? v00 : 0.f, z01 = up ? v01 : 0.f, z02 = (up && rt) ? v02 : 0.f;
      const float z10 = lf ? v10 : 0.f, z12 = rt ? v12 : 0.f;
      const float z20 = (dn && lf) ? v20 : 0.f, z21 = dn ? v21 : 0.f, z22 = (dn && rt) ? v22 : 0.f;
      float a = __fmul_rn(z00, -1.f);
      a = fmaf(z02, 1.f, a); a = fmaf(z10, -2.f, a); a = fmaf(z12, 2.f, a); a = fmaf(z20, -1.f, a);
      gx = fmaf(z22, 1.f, a);
      float cc = __fmul_rn(z00, -1.f);
      cc = fmaf(z01, -2.f, cc); cc = fmaf(z02, -1.f, cc); cc = fmaf(z20, 1.f, cc); cc = fmaf(z21, 2.f, cc);
      gy = fmaf(z22, 1.f, cc);
    }
    {   // warp-aggregated histogram update: one shared-memory atomic per distinct (tile, label)
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(&lbp_hist[key], __popc(peers));
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
    if (valid && (lane & (tile - 1)) == 0)
      reinterpret_cast<float4*>(rowsum)[r * wt + (x >> tshift)] = make_float4(s0, s1, s2, s3);
  }
  __syncthreads();
  STAGE_CLOCK(3);

  // ---- S5: phi2 (LBP entropy, morphology.py:648-652) and phi3 (654-670) per tile -----------
  {
    for (int t = t_lo + tid; t < t_hi; t += NT) {
      const int ty = t / wt, tx = t - ty * wt;
      float ent = 0.f;
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) {
        const int cnt = lbp_hist[t * 10 + kk];
        if (A.lbp_dbg) A.lbp_dbg[((long long)b * g.ntiles + t) * 10 + kk] = cnt;
        const float p = __fdiv_rn((float)cnt, ntile2);
        ent = __fadd_rn(ent, __fmul_rn(p, lutp[cnt]));             // log2(p + 1e-10)
      }
      phis[t * 5 + 1] = __fdiv_rn(-ent, kc::LOG2_10);
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
  }
  csync();           // every CTA is done with P1 = 255*gray; BIN bands have landed
  STAGE_CLOCK(4);

  // ---- S6: 5x5 Gaussian blur (zero padding) -> P1, Otsu histogram (morphology.py:485-493) --
  float* P1r[3] = {P1, P1, P1};
  float* P0r[3] = {P0, P0, P0};
  for (int pr = 1; pr < ns; ++pr) {
    P1r[pr - 1] = cl.map_shared_rank(P1, (rank + pr) % ns);
    P0r[pr - 1] = cl.map_shared_rank(P0, (rank + pr) % ns);
  }
  for (int task = warp; task < nrg * WW; task += nwarps) {
    const int rg = task / WW, k = task - rg * WW;
    const int r0 = r_lo + rg * RT;
    const int x = 32 * k + lane;
    const bool valid = x < Wc;
    int xc[5];
    bool xok[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { const int xx = x + j - 2; xok[j] = xx >= 0 && xx < Wc; xc[j] = clampi(xx, 0, Wc - 1); }
    float acc[RT];
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[j] = 0.f;
#pragma unroll
    for (int rr = 0; rr < RT + 4; ++rr) {
      const int r = r0 - 2 + rr;
      const bool rok = r >= 0 && r < Hc;
      const float* row = P0 + clampi(r, 0, Hc - 1) * Wc;
      float v[5];
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) v[kx] = (rok && xok[kx]) ? row[xc[kx]] : 0.f;   // fma(0, w, acc) == acc
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int ky = rr - j;
        if (ky >= 0 && ky < 5) {
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) acc[j] = fmaf(v[kx], kc::CANNY[ky * 5 + kx], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      int bin = -1;
      if (valid) {
        P1[(r0 + j) * Wc + x] = acc[j];
        for (int pr = 1; pr < ns; ++pr) P1r[pr - 1][(r0 + j) * Wc + x] = acc[j];
        if (acc[j] >= 0.f && acc[j] <= 1.f) {          // torch.histc(bins=256, min=0, max=1)
          bin = (int)__fmul_rn(acc[j], 256.f);
          if (bin == 256) bin = 255;
        }
      }
      const unsigned peers = __match_any_sync(0xffffffffu, bin);     // smooth images: many equal bins
      if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(&hloc[bin], __popc(peers));
    }
  }
  __syncthreads();
  for (int pr = 0; pr < ns; ++pr) {                               // integer counts: order-free, exact
    int* hp = ns > 1 ? cl.map_shared_rank(hist, pr) : hist;
    for (int i = tid; i < 256; i += NT)
      if (hloc[i]) atomicAdd(&hp[i], hloc[i]);
  }
  csync();           // blurred bands and the whole-image histogram are complete everywhere
  STAGE_CLOCK(5);

  // ---- S7 (warp 0): Otsu threshold (morphology.py:397-418)  ||  S8 (other warps): magnitude ---
  if (warp == 0) {
    int cnt[8];
    int tot_i = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { cnt[j] = hist[lane * 8 + j]; tot_i += cnt[j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot_i += __shfl_xor_sync(0xffffffffu, tot_i, o);
    const float tot = fmaxf((float)tot_i, 1.0f);
    float p[8], pc[8];
    double so = 0.0, sm = 0.0;                      // fp64 sums of these fp32 terms are exact
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p[j] = __fdiv_rn((float)cnt[j], tot);
      const float center = (float)(2 * (lane * 8 + j) + 1) * 0.001953125f;   // (i + 0.5) / 256, exact
      pc[j] = __fmul_rn(p[j], center);
      so += (double)p[j];
      sm += (double)pc[j];
    }
    double io = so, im = sm;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double to = __shfl_up_sync(0xffffffffu, io, o);
      const double tm = __shfl_up_sync(0xffffffffu, im, o);
      if (lane >= o) { io += to; im += tm; }
    }
    const float mu_t = (float)__shfl_sync(0xffffffffu, im, 31);
    double ro = io - so, rm = im - sm;
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
      const float thr = (float)(2 * best_i + 1) * 0.001953125f;
      red[0] = __fmul_rn(thr, 255.f);
      red[1] = __int_as_float(best_i);
    }
  } else {
    // ---- S8: L1 gradient magnitude of Sobel(255 * blur) -> P0 (morphology.py:496-497) and the
    //      NMS direction bin (2 bits per pixel, parked in the not-yet-used STRONG/WEAK planes)
    for (int slot = warp - 1; slot < nslot_b; slot += nwarps - 1) {
      const int r = r_lo + slot / WW, k = slot % WW;
      const int x = 32 * k + lane;
      int bin = 0;
      if (x < Wc) {
        float gx, gy;
        sobel_at<true>(P1, r, x, Hc, Wc, gx, gy);
        const float mg = __fadd_rn(fabsf(gx), fabsf(gy));
        P0[r * Wc + x] = mg;
        for (int pr = 1; pr < ns; ++pr) P0r[pr - 1][r * Wc + x] = mg;
        bin = nms_bin(gx, gy);
      }
      const uint32_t d0 = __ballot_sync(0xffffffffu, bin & 1);
      const uint32_t d1 = __ballot_sync(0xffffffffu, bin & 2);
      if (lane == 0) { STRONG[r * WW + k] = d0; WEAK[r * WW + k] = d1; }
    }
  }
  csync();           // magnitude bands have landed everywhere
  STAGE_CLOCK(6);
  const float thr255 = red[0];
  const int otsu_bin = __float_as_int(red[1]);
  const float thr_lo = __fmul_rn(0.5f, thr255);

  // ---- S9: non-maximum suppression + double threshold (morphology.py:426-449, 500-502) ----
  for (int bslot = warp; bslot < nslot_b; bslot += nwarps) {
    const int r = r_lo + bslot / WW, k = bslot % WW;
    const int slot = r * WW + k;
    const int x = 32 * k + lane;
    const uint32_t d0 = STRONG[slot], d1 = WEAK[slot];     // direction bits of this slot (same warp rewrites it)
    bool st = false, wk = false;
    if (x < Wc) {
      const float mag = P0[r * Wc + x];
      if (mag > thr_lo) {                           // below the weak threshold the pixel is irrelevant
        const int bin = ((d0 >> lane) & 1) | (((d1 >> lane) & 1) << 1);
        const int dy1 = bin == 0 ? 0 : -1;
        const int dx1 = bin == 0 ? 1 : (bin == 1 ? 1 : (bin == 2 ? 0 : -1));
        const float n1 = P0[clampi(r + dy1, 0, Hc - 1) * Wc + clampi(x + dx1, 0, Wc - 1)];
        const float n2 = P0[clampi(r - dy1, 0, Hc - 1) * Wc + clampi(x - dx1, 0, Wc - 1)];
        const float nms = (mag >= n1 && mag >= n2) ? mag : 0.f;
        st = nms > thr255;
        wk = nms > thr_lo;
      }
    }
    __syncwarp();
    const uint32_t ws = __ballot_sync(0xffffffffu, st);
    const uint32_t ww = __ballot_sync(0xffffffffu, wk);
    if (lane == 0) { STRONG[slot] = ws; WEAK[slot] = ww; }
  }
  __syncthreads();
  if (ns > 1) {
    publish(STRONG, r_lo * WW, (r_hi - r_lo) * WW);
    publish(WEAK, r_lo * WW, (r_hi - r_lo) * WW);
    cl.sync();       // full strong / weak planes everywhere; hysteresis then runs redundantly per CTA
  }
  STAGE_CLOCK(7);

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
  STAGE_CLOCK(8);

  // ---- S11: integer tile counts ------------------------------------------------------------
  // acc[t][0..8] = edge, area, perim, euler_x4, N_2, N_4, N_8, N_16, N_32   (aliases P0)
  int* acc = reinterpret_cast<int*>(P0);
  for (int i = tid; i < g.ntiles * 9; i += NT) acc[i] = 0;
  __syncthreads();
  const int segs = 32 >> tshift;
  const uint32_t segmask = tile == 32 ? 0xffffffffu : ((1u << tile) - 1u);
  for (int i = r_lo * WW + tid; i < r_hi * WW; i += NT) {
    const int r = i / WW, k = i - r * WW;
    const int nbits = min(32, Wc - 32 * k);
    const uint32_t vmask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u);
    const uint32_t e = EDGE[i];
    const uint32_t m = BIN[i];
    uint32_t er = 0xffffffffu;                                    // 3x3 erosion, outside ignored
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int rr = r + dy;
      if (rr < 0 || rr >= Hc) continue;
      const uint32_t c = BIN[rr * WW + k];
      const uint32_t l = k > 0 ? BIN[rr * WW + k - 1] : 0xffffffffu;
      const uint32_t rn = k + 1 < WW ? BIN[rr * WW + k + 1] : 0xffffffffu;
      const uint32_t left = (c << 1) | (l >> 31);
      uint32_t right = (c >> 1) | (rn << 31);
      if (nbits < 32) right |= (1u << (nbits - 1));
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
  STAGE_CLOCK(9);
  // dyadic box counts (morphology.py:595-601): one thread per (tile, scale)
  for (int ts = t_lo * g.S + tid; ts < t_hi * g.S; ts += NT) {
    const int t = ts / g.S, sidx = ts - t * g.S;
    const int s = 2 << sidx;
    const int ty = t / wt, tx = t - ty * wt;
    const int x0 = tx * tile, k = x0 >> 5, sh = x0 & 31;
    uint32_t colmask = 0;
    for (int q = 0; q < tile; q += s) colmask |= (1u << q);
    int n = 0;
    for (int y0 = 0; y0 < tile; y0 += s) {
      uint32_t o = 0;
      for (int y = 0; y < s; ++y) o |= (EDGE[(ty * tile + y0 + y) * WW + k] >> sh) & segmask;
      for (int d = 1; d < s; d <<= 1) o |= o >> d;
      n += __popc(o & colmask);
    }
    acc[t * 9 + 4 + sidx] = n;
  }
  __syncthreads();
  STAGE_CLOCK(10);

  // ---- S12: phi1, phi4, phi5, interactions (morphology.py:852-864) --------------------------
  for (int t = t_lo + tid; t < t_hi; t += NT) {
    const int* a = acc + t * 9;
    const int S = g.S;
    float y[5];
    for (int i = 0; i < S; ++i) y[i] = lutn[a[4 + i]];                         // log(N_s + 1)
    float w_sum = 0.f, sx = 0.f, sy = 0.f;
    for (int i = 0; i < S; ++i) {
      w_sum = __fadd_rn(w_sum, kc::FW[i]);
      sx = __fadd_rn(sx, __fmul_rn(kc::FW[i], kc::FLOG[i]));
      sy = __fadd_rn(sy, __fmul_rn(kc::FW[i], y[i]));
    }
    const float x_mean = __fdiv_rn(sx, w_sum), y_mean = __fdiv_rn(sy, w_sum);
    float cov = 0.f, var = 0.f;
    for (int i = 0; i < S; ++i) {
      const float dx = __fsub_rn(kc::FLOG[i], x_mean);
      cov = __fadd_rn(cov, __fmul_rn(__fmul_rn(kc::FW[i], dx), __fsub_rn(y[i], y_mean)));
      var = __fadd_rn(var, __fmul_rn(kc::FW[i], __fmul_rn(dx, dx)));
    }
    float df = -__fdiv_rn(cov, __fadd_rn(var, 1e-12f));
    df = fminf(fmaxf(df, 1.0f), 2.0f);
    const float p1 = S < 2 ? 0.5f : __fdiv_rn(df, 2.0f);
    const float p2 = phis[t * 5 + 1], p3 = phis[t * 5 + 2];
    const float p4 = __fdiv_rn((float)a[0], ntile2);
    const float area = (float)a[1], perim = (float)a[2];
    float ic = __fdiv_rn(__fmul_rn(perim, perim), __fadd_rn(__fmul_rn(kc::FOUR_PI, area), 1e-6f));
    const float K = fmaxf(rintf(__fdiv_rn((float)a[3], 4.0f)), 1.0f);
    ic = __fdiv_rn(ic, K);
    float p5 = __fsub_rn(1.0f, __fdiv_rn(1.0f, fmaxf(ic, 1.0f)));
    if (a[1] <= 0) p5 = 0.f;
    float o[8];
    o[0] = p1; o[1] = p2; o[2] = p3; o[3] = p4; o[4] = p5;
    o[5] = __fmul_rn(p1, p2);
    o[6] = __fmul_rn(p3, p3);
    o[7] = __fsqrt_rn(__fadd_rn(__fmul_rn(p4, p5), 1e-12f));
#pragma unroll
    for (int i = 0; i < 8; ++i) phi8[t * 8 + i] = o[i];
    if (A.phi) {
      float4* dst = reinterpret_cast<float4*>(A.phi + ((long long)b * g.ntiles + t) * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (A.counts_dbg) {
      int* c = A.counts_dbg + ((long long)b * g.ntiles + t) * 12;
      for (int i = 0; i < 9; ++i) c[i] = a[i];
      c[9] = otsu_bin; c[10] = 0; c[11] = 0;
    }
  }
  if (A.edge_dbg && rank == 0)
    for (int i = tid; i < NW; i += NT) A.edge_dbg[(long long)b * NW + i] = EDGE[i];
  if (A.bin_dbg && rank == 0)
    for (int i = tid; i < NW; i += NT) A.bin_dbg[(long long)b * NW + i] = BIN[i];
  __syncthreads();
  STAGE_CLOCK(11);
  if (!A.cmlp) return;                       // last remote access was before the strong/weak sync

  // ---- N1: complexity MLP (own tiles) -> all-gather -> bilateral (own tiles) -> all-gather ----
  float* scratch = P0;                       // the planes are dead from here on
  const int nt = g.ntiles;
  {
    float* w = scratch;
    float* act = w + CMLP_SMEM_FLOATS;
    complexity_load_weights(A.cmlp, w);
    __syncthreads();
    complexity_mlp_range(phi8, t_lo, t_hi, w, act, craw_s,
                         A.complexity_raw ? A.complexity_raw + (long long)b * nt : nullptr);
    if (ns > 1) { publish(craw_s, t_lo, t_hi - t_lo); cl.sync(); }
    bilateral_range(craw_s, g.ht, g.wt, t_lo, t_hi, act, cfin, A.complexity ? A.complexity + (long long)b * nt : nullptr);
    if (ns > 1) { publish(cfin, t_lo, t_hi - t_lo); cl.sync(); }
  }
  STAGE_CLOCK(12);
  if (!A.run_mapper) return;
  // ---- N2: bit mapper (own tiles) --------------------------------------------------------------
  float* bout = A.bit_map ? A.bit_map + (long long)b * nt : nullptr;
  if (A.linear_mapper) {
    int npow2 = 1;
    while (npow2 < nt) npow2 <<= 1;
    mapper_linear_range(cfin, nt, npow2, scratch, t_lo, t_hi, A.temperature, A.use_t, A.continuous, A.lo, A.hi,
                        A.eps_spread, bits_s, bout);
  } else {
    float* w = scratch;
    float* act = w + MAPPER_SMEM_FLOATS;
    mapper_load_weights(A.mapper, w);
    __syncthreads();
    mapper_mlp_range(cfin, t_lo, t_hi, w, act, A.temperature, A.use_t, A.continuous, A.lo, A.hi, bits_s, bout);
  }
  STAGE_CLOCK(13);
  if (!A.softmask || !A.abs_plane) {
    if (ns > 1) cl.sync();                   // nobody leaves while a peer may still write into it
    return;
  }
  // ---- N3: soft mask (quantization.py:213-239): tile head (own tiles) + m rows (own band) ------
  {
    float* rows = scratch;                   // [H*wt]
    float* P = rows + g.H * g.wt;            // 196
    float* bn = P + 196;                     // [nt]
    float* an = bn + nt;                     // [nt]
    float amax = softmask_act_range(A.abs_plane + (long long)b * g.H * g.W, g.C, g.H, g.W, g.ht, g.wt, tr0, tr1,
                                    rows, red, act_s);
    if (ns > 1) {
      if (tid == 0)
        for (int pr = 0; pr < ns; ++pr) cl.map_shared_rank(mmx, pr)[8 + rank] = amax;
      publish(bits_s, t_lo, t_hi - t_lo);
      publish(act_s, t_lo, t_hi - t_lo);
      cl.sync();
      amax = mmx[8];
      for (int pr = 1; pr < ns; ++pr) amax = fmaxf(amax, mmx[8 + pr]);
    }
    softmask_head_range(bits_s, act_s, amax, g.ht, g.wt, t_lo, t_hi, A.softmask, P, bn, an, mt_s,
                        A.mask_tiles ? A.mask_tiles + (long long)b * nt : nullptr);
    if (ns > 1) { publish(mt_s, t_lo, t_hi - t_lo); cl.sync(); }
    softmask_plane_rows(mt_s, P, g.H, g.W, g.ht, g.wt, (g.H * rank) / ns, (g.H * (rank + 1)) / ns,
                        A.mask + (long long)b * g.H * g.W);
  }
  STAGE_CLOCK(14);
}

}  // namespace mcaq

using namespace mcaq;

static long long* g_stage_clk = nullptr;
extern "C" void mcaq_debug_stage_clocks(long long* dev_buf) { g_stage_clk = dev_buf; }

static int max_i(int a, int b) { return a > b ? a : b; }

static int g_force_split = 0;
// debug / tuning: force the number of CTAs per image (0 = automatic)
extern "C" void mcaq_debug_cluster_split(int ns) { g_force_split = ns; }

// CTAs per image: tile rows are split over a thread-block cluster (portable size <= 8)
static int pick_split(int ht) {
  int ns = ht >= 4 ? 4 : (ht >= 2 ? 2 : 1);
  if (g_force_split == 1 || g_force_split == 2 || g_force_split == 4) ns = g_force_split;
  while (ns > ht) ns >>= 1;
  return ns < 1 ? 1 : ns;
}


// geometry + shared-memory layout; returns bytes of dynamic smem or a negative error
static long long plan(MorphGeom& g, int B, int C, int H, int W, int grid_size, bool nets, int threads) {
  g.B = B; g.C = C; g.H = H; g.W = W;
  g.tile = mcaq_tile_size(H, grid_size);
  g.ht = H / g.tile; g.wt = W / g.tile;
  if (g.ht <= 0 || g.wt <= 0 || g.tile > 32 || g.tile < 4) return MCAQ_EINVAL;
  g.Hc = g.ht * g.tile; g.Wc = g.wt * g.tile;
  g.WW = (g.Wc + 31) / 32;
  g.ntiles = g.ht * g.wt;
  g.S = 0;
  for (int s = 2; s <= g.tile; s <<= 1) g.S++;
  g.ns = pick_split(g.ht);
  const long long NP = (long long)g.Hc * g.Wc, NW = (long long)g.Hc * g.WW;
  const long long rowsum_w = (long long)g.Hc * g.wt * 4, lbp_w = (long long)g.ntiles * 10;
  long long np1 = NP < rowsum_w ? rowsum_w : NP;
  np1 = (np1 + 3) & ~3LL;
  g.off_rowsum = (int)NP;                       // NP % 16 == 0
  long long tail = NP + np1;
  if (rowsum_w + lbp_w <= np1) {
    g.off_lbp = (int)(NP + rowsum_w);
  } else {
    g.off_lbp = (int)tail;
    tail += (lbp_w + 3) & ~3LL;
  }
  if (nets) {
    (void)threads;
    int npow2 = 1;
    while (npow2 < g.ntiles) npow2 <<= 1;
    long long need = CMLP_SMEM_FLOATS + max_i(CPX_ACT_FLOATS, 25 * g.ntiles);
    need = max_i((int)need, MAPPER_SMEM_FLOATS + MAP_ACT_FLOATS);
    need = max_i((int)need, npow2);
    need = max_i((int)need, H * g.wt + 196 + 2 * g.ntiles);
    if (tail < need) tail = (need + 3) & ~3LL;
  }
  g.off_tail = (int)tail;
  const long long words = tail + 3 * NW + (long long)g.ntiles * (5 + 8 + 5) + 512 + 64 + 16 + 260 +
                          g.tile * g.tile + 4;
  return words * 4;
}

// threads per CTA from the pixels of the largest band
static int pick_threads(const MorphGeom& g) {
  const long long band = (long long)((g.ht + g.ns - 1) / g.ns) * g.tile * g.Wc;
  return band <= 640 ? 128 : (band <= 4096 ? 256 : MORPH_THREADS);
}

static int launch_fused(FusedArgs& A, long long smem, int threads, cudaStream_t st) {
  if (smem > 227 * 1024) return MCAQ_ETOOBIG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(morph_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  A.clk = g_stage_clk;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(A.g.B * A.g.ns));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)A.g.ns;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, morph_fused_kernel, A);
  if (e != cudaSuccess) return (int)e;
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_morph_phi(const float* sum_plane, int B, int C, int H, int W, int grid_size,
                              const float* consts, float* phi, float* gray_dbg, uint32_t* edge_bits_dbg,
                              uint32_t* bin_bits_dbg, int32_t* lbp_hist_dbg, int32_t* counts_dbg, void* stream) {
  (void)consts;   // compiled in (mcaq_consts.cuh); argument kept for ABI stability
  if (!sum_plane || !phi || B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0) return MCAQ_EINVAL;
  FusedArgs A = {};
  const long long smem = plan(A.g, B, C, H, W, grid_size, false, 0);
  if (smem < 0) return (int)smem;
  const int threads = pick_threads(A.g);
  A.sum_plane = sum_plane;
  A.phi = phi;
  A.gray_dbg = gray_dbg; A.edge_dbg = edge_bits_dbg; A.bin_dbg = bin_bits_dbg;
  A.lbp_dbg = lbp_hist_dbg; A.counts_dbg = counts_dbg;
  return launch_fused(A, smem, threads, (cudaStream_t)stream);
}

extern "C" int mcaq_morph_fused(const float* sum_plane, const float* abs_plane, int B, int C, int H, int W,
                                int grid_size, int32_t* keys, float* packed_ranges, const float* cmlp,
                                const float* mapper, int linear_mapper, const float* softmask,
                                float temperature, int use_temperature, int continuous, float min_bits,
                                float max_bits, float eps_spread, float* phi, float* complexity,
                                float* bit_map, float* mask, void* stream) {
  if (!sum_plane || !cmlp || !complexity || !bit_map || B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0)
    return MCAQ_EINVAL;
  if (!mapper && !linear_mapper) return MCAQ_EINVAL;
  if (softmask && (!abs_plane || !mask)) return MCAQ_EINVAL;
  if (keys && !packed_ranges) return MCAQ_EINVAL;
  FusedArgs A = {};
  const long long smem = plan(A.g, B, C, H, W, grid_size, true, 0);
  if (smem < 0) return (int)smem;
  const int threads = pick_threads(A.g);
  A.sum_plane = sum_plane; A.abs_plane = abs_plane;
  A.keys = keys; A.packed = packed_ranges;
  A.cmlp = cmlp; A.mapper = mapper; A.softmask = softmask;
  A.run_mapper = 1; A.linear_mapper = linear_mapper; A.use_t = use_temperature; A.continuous = continuous;
  A.temperature = temperature; A.lo = min_bits; A.hi = max_bits; A.eps_spread = eps_spread;
  A.phi = phi; A.complexity = complexity; A.bit_map = bit_map; A.mask = mask;
  return launch_fused(A, smem, threads, (cudaStream_t)stream);
}
