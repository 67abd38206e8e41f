// K2: per-image morphology on the on-chip gray plane -> phi(8) per tile, optionally followed in
// the same launch by the complexity MLP + bilateral filter, the bit mapper and the soft mask
// (the whole "between the two HBM sweeps" part of the hook in one kernel).
//
// An image is split by tile rows over the `ns` CTAs of a thread-block cluster.  Each CTA keeps only
// its band of the float planes (plus the stencil halos, recomputed locally) in shared memory:
//   G   normalised gray, rows [r_lo-5, r_hi+5), two zero columns each side (zero rows outside the image)
//   BL  255 * blur5x5(G), rows [r_lo-2, r_hi+2), one zero column each side
//   MAG |Sobel(BL)| (L1),  rows [r_lo-1, r_hi+1)
// Binary maps (adaptive mask, NMS direction, Canny strong / weak / edge) are 1-bit planes (32 pixels
// per word, built with warp ballots) indexed by image row; the strong / weak / adaptive planes are
// all-gathered through distributed shared memory, so hysteresis, erosion, Euler quads and box
// counting are word-parallel bit operations on whole-image planes.
//
// Every pixel stage is organised as "one thread owns a column, a warp owns 32 adjacent columns and
// walks a run of rows": stencil inputs slide through registers, there is no per-pixel index
// arithmetic, and tile statistics are accumulated down the column in registers and combined across
// the tile's lanes once per tile.  Stages between two barriers (blur + Otsu histogram, adaptive
// threshold, LBP + gradient variance, soft-mask activity) are independent warp tasks.
//
//   L   gray = sum / C (band + halo), band min / max      -> cluster exchange of min / max
//   N   normalise in place (morphology.py:378-383)
//   T1  tasks: 5x5 blur -> BL + histogram | 11x11 adaptive threshold -> BIN | LBP + Sobel(G) -> phi2, phi3
//              | tile activity of sum_c|x| (soft mask)
//   T2  |Sobel(BL)| -> MAG, NMS direction bits         -> cluster sync (histogram complete)
//   T3  Otsu (every warp, fp64 scan), NMS + double threshold -> STRONG / WEAK  -> cluster sync
//   T4  hysteresis: 8 constrained dilations held in registers (lane = row) -> EDGE
//   T5  integer tile counts (edge, area, perimeter, Euler, boxes) -> phi1, phi4, phi5, interactions
//   N1  complexity MLP -> gather -> bilateral   N2  bit mapper   N3  soft-mask head -> gather -> m
//
// Arithmetic contract = oracle/mcaq_oracle.py: stencils are FMA chains over the taps in row-major
// order from 0 (taps are FFMA immediates from mcaq_consts.cuh), everything else separately rounded
// fp32; float tile sums are column sums top-to-bottom added left-to-right; log tables are
// fp64-rounded literals; Otsu sums are exact in fp64.
#include <cooperative_groups.h>

#include "peer_exchange.cuh"
#include "tile_nets.cuh"
#include "morph_common.cuh"

namespace cg = cooperative_groups;

namespace mcaq {

#ifndef MCAQ_MORPH_MAX_THREADS
#define MCAQ_MORPH_MAX_THREADS 512
#endif
constexpr int MORPH_MAX_THREADS = MCAQ_MORPH_MAX_THREADS;

struct MorphGeom {
  int B, C, H, W, tile, ht, wt, Hc, Wc, WW, ntiles, S;
  int ns;             // CTAs per image (thread-block cluster size)
  int band_max;       // largest band (rows) of a CTA
  int gs, bs;         // row strides of G (Wc + 4) and BL (Wc + 2)
  int aligned;        // H == Hc && W == Wc: soft-mask windows are the analyzer's tiles
  // shared-memory layout, offsets in 4-byte words
  int off_bl, off_mag, off_h, off_bits, off_tiles, off_w, words;
  int hs;             // row stride of the horizontally filtered plane (Wc + 1: odd)
  unsigned m_ww, m_hseg;   // 2^32 / d + 1 for d = WW and the horizontal tasks per row group (fast_div)
  int max_own;        // most tiles a CTA owns
  int threads;        // CTA size (fixes the per-warp scratch of the tile networks)
};

struct FusedArgs {
  MorphGeom g;
  const float* sum_plane;
  const float* abs_plane;      // for the soft mask (NULL: no mask stage)
  int* keys;                   // K1 range keys: decoded into `packed` and re-armed by CTA 0 (NULL: skip)
  float* packed;
  XchgPeers px;                // world > 1: CTA 0 also publishes `packed` to every rank (peer_exchange.cuh)
  const float* cmlp;           // NULL: stop after phi
  const float* mapper;         // packed MLP mapper, or NULL with linear_mapper != 0
  const float* steps;          // step table of the MLP mapper (tile_nets.cuh), or NULL
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

struct Ctx {
  // planes (biased so that [r * stride + x] works with image coordinates)
  const float* Gp; int gs;
  float* BLp; int bs;
  float* MAGp;
  float* Hp; int hs;       // horizontally filtered gray (adaptive threshold), biased like Gp without pad columns
  uint32_t *BIN, *DIR0, *DIR1, *STRONG, *WEAK, *EDGE;
  int Hc, Wc, WW, tile, tshift, wt;
  int r_lo, r_hi;
  int ns, rank;
};

// ---- T1a: 5x5 Gaussian blur (zero padding) of rows [r0, r0+RT) -> BL = 255 * blur; Otsu histogram
//      of the blurred value for band rows (morphology.py:485-493).  Equal bins of vertically adjacent
//      pixels are merged before the shared-memory atomic.
template <int RT>
__device__ __forceinline__ void task_blur(const Ctx& c, int r0, int rend, int k, int lane, int* hloc) {
  const int x = 32 * k + lane;
  const bool valid = x < c.Wc;
  const int xr = valid ? x : c.Wc - 1;
  const float* base = c.Gp + (r0 - 2) * c.gs + xr - 2;
  float acc[RT];
#pragma unroll
  for (int j = 0; j < RT; ++j) acc[j] = 0.f;
#pragma unroll
  for (int rr = 0; rr < RT + 4; ++rr) {
    float v[5];
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) v[kx] = base[rr * c.gs + kx];
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      const int ky = rr - j;
      if (ky >= 0 && ky < 5) {
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) acc[j] = fmaf(v[kx], kc::CANNY[ky * 5 + kx], acc[j]);
      }
    }
  }
  int cur = -1, cnt = 0;
#pragma unroll
  for (int j = 0; j < RT; ++j) {
    const int r = r0 + j;
    if (valid && r < rend) {
      c.BLp[r * c.bs + x] = __fmul_rn(acc[j], 255.f);
      if (r >= c.r_lo && r < c.r_hi && acc[j] >= 0.f && acc[j] <= 1.f) {   // torch.histc(bins=256, min=0, max=1)
        int bin = (int)__fmul_rn(acc[j], 256.f);
        bin = min(bin, 255);
        if (bin == cur) {
          ++cnt;
        } else {
          if (cnt) atomicAdd(&hloc[cur], cnt);
          cur = bin;
          cnt = 1;
        }
      }
    }
  }
  if (cnt) atomicAdd(&hloc[cur], cnt);
}

// ---- T1b: adaptive threshold, 11x11 Gaussian mean of 255*G with replicate borders, rows [r0, r0+RT)
//      (morphology.py:550-573) -> BIN words (own plane and every peer's).
// The reference's value is the 121-tap FMA chain in row-major order; only the SIGN of
// 255 g - (mean - 2) is kept.  The 2-D kernel is the fp32 outer product of an 11-tap vector, so a
// separable evaluation (11 + 11 FMAs) differs from the chain by at most
//   121 u 255 (chain) + 22 u 255 (separable) + 3 u 255 (scalings)  <  2.3e-3      (u = 2^-24, g in [0,1]),
// and decides every pixel whose margin exceeds 4e-3; the few pixels inside the guard band (about one
// in a thousand) take the literal chain, so the bit plane is exactly the reference's.
// literal reference arithmetic for one pixel (not inlined: rare)
static __device__ __noinline__ bool adaptive_exact(const float* Gp, int gs, int Hc, int Wc, int r, int x) {
  float acc = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 11; ++ky) {
    const float* row = Gp + clampi(r - 5 + ky, 0, Hc - 1) * gs;
#pragma unroll
    for (int kx = 0; kx < 11; ++kx)
      acc = fmaf(__fmul_rn(row[clampi(x + kx - 5, 0, Wc - 1)], 255.f), __fmul_rn(ADAPT1[ky], ADAPT1[kx]), acc);
  }
  return __fmul_rn(Gp[r * gs + x], 255.f) > __fsub_rn(acc, 2.0f);
}

// Horizontal 11-tap pass (replicate borders) of image rows lr0 + lane (< r_end) for columns [x0, x1), x1 - x0 a
// multiple of 8: a thread owns a row and produces 8 consecutive outputs from 18 inputs held in registers.
// The column clamps are warp-uniform, rows of G have an odd stride (conflict-free for lane = row).
__device__ __forceinline__ void task_adapt_h(const Ctx& c, int r_begin, int r_end, int x0, int x1, int lane) {
  const int r = min(r_begin + lane, r_end - 1);
  const bool ok = r_begin + lane < r_end;
  const float* row = c.Gp + r * c.gs;
  float* out = c.Hp + r * c.hs;
#pragma unroll 1
  for (int xb = x0; xb < x1; xb += 8) {
    float v[18];
    if (xb >= 5 && xb + 12 < c.Wc) {                      // interior block (warp-uniform): immediate offsets
#pragma unroll
      for (int j = 0; j < 18; ++j) v[j] = row[xb + j - 5];
    } else {
#pragma unroll
      for (int j = 0; j < 18; ++j) v[j] = row[clampi(xb + j - 5, 0, c.Wc - 1)];
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int kx = 0; kx < 11; ++kx) acc = fmaf(v[o + kx], ADAPT1[kx], acc);
      if (ok && xb + o < c.Wc) out[xb + o] = acc;
    }
  }
}

// Rows [r0, min(r0 + RT, r1)) of the columns of word k: the vertical 11-tap pass over the horizontally
// filtered plane (RT + 10 input rows per RT output rows, RT independent accumulators), then the sign.
template <int RT>
__device__ __forceinline__ void task_adaptive(const Ctx& c, cg::cluster_group& cl, int r0, int r1, int k, int lane) {
  const int x = 32 * k + lane;
  const bool valid = x < c.Wc;
  const int xq = valid ? x : c.Wc - 1;
  float acc[RT];
#pragma unroll
  for (int j = 0; j < RT; ++j) acc[j] = 0.f;
#pragma unroll
  for (int rr = 0; rr < RT + 10; ++rr) {
    const float h = c.Hp[clampi(r0 - 5 + rr, 0, c.Hc - 1) * c.hs + xq];
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      const int ky = rr - j;
      if (ky >= 0 && ky < 11) acc[j] = fmaf(h, ADAPT1[ky], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < RT; ++j) {
    const int r = r0 + j;
    if (r < r1) {                                           // warp-uniform
      const float ctr = __fmul_rn(c.Gp[r * c.gs + xq], 255.f);
      const float d = __fsub_rn(ctr, __fsub_rn(__fmul_rn(acc[j], 255.f), 2.0f));
      bool bit = d > 0.f;
      if (fabsf(d) <= ADAPT_GUARD) bit = adaptive_exact(c.Gp, c.gs, c.Hc, c.Wc, r, xq);
      const uint32_t word = __ballot_sync(0xffffffffu, bit && valid);
      if (lane < c.ns) {
        uint32_t* dst = lane == 0 ? c.BIN : cl.map_shared_rank(c.BIN, (c.rank + lane) % c.ns);
        dst[r * c.WW + k] = word;
      }
    }
  }
}

// ---- T1c: uniform-LBP histogram (morphology.py:623-652) and Sobel(G) statistics (654-670) of the
//      tiles of tile row ty under word k -> phi2, phi3.
__device__ __forceinline__ void task_lbp_var(const Ctx& c, int ty, int k, int lane, const float* lutp, float* phis,
                                             const unsigned long long* __restrict__ lbp_inc,
                                             int* __restrict__ lbp_dbg) {
  const int x = 32 * k + lane;
  const bool valid = x < c.Wc;
  const int xr = valid ? x : c.Wc - 1;
  const bool x0 = xr == 0, xN = xr == c.Wc - 1;
  const int tile = c.tile;
  const int r0 = ty * tile;
  const float* p = c.Gp + (r0 - 1) * c.gs + xr;
  // window rows: a = above, m = current, b = below; L / R zero padded, Lr / Rr replicate padded
  float aL = p[-1], aC = p[0], aR = p[1];
  p += c.gs;
  float mL = p[-1], mC = p[0], mR = p[1];
  float aLr = x0 ? aC : aL, aRr = xN ? aC : aR;
  float mLr = x0 ? mC : mL, mRr = xN ? mC : mR;
  const bool top = r0 == 0, bot = r0 + tile == c.Hc;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  unsigned long long cnt = 0ull;                          // ten 6-bit counters (a column has <= 32 pixels)
  // one pixel of the column: LBP label through the 256-entry table of counter increments
  // (1 << 6 * label), Sobel statistics, window shift.  rt / rb: replicate padding in y.
  auto step = [&](bool rt, bool rb) {
    p += c.gs;
    const float bL = p[-1], bC = p[0], bR = p[1];
    const float bLr = x0 ? bC : bL, bRr = xN ? bC : bR;
    const float uL = rt ? mLr : aLr, uC = rt ? mC : aC, uR = rt ? mRr : aRr;
    const float dL = rb ? mLr : bLr, dC = rb ? mC : bC, dR = rb ? mRr : bRr;
    // neighbour order (-1,-1),(-1,0),(-1,1),(0,1),(1,1),(1,0),(1,-1),(0,-1)  (morphology.py:634)
    const uint32_t code = (uint32_t)(uL >= mC) | ((uint32_t)(uC >= mC) << 1) | ((uint32_t)(uR >= mC) << 2) |
                          ((uint32_t)(mRr >= mC) << 3) | ((uint32_t)(dR >= mC) << 4) | ((uint32_t)(dC >= mC) << 5) |
                          ((uint32_t)(dL >= mC) << 6) | ((uint32_t)(mLr >= mC) << 7);
    cnt += lbp_inc[code];
    float gx, gy;
    sobel3(aL, aC, aR, mL, mR, bL, bC, bR, gx, gy);
    s0 = __fadd_rn(s0, gx);
    s1 = __fadd_rn(s1, __fmul_rn(gx, gx));
    s2 = __fadd_rn(s2, gy);
    s3 = __fadd_rn(s3, __fmul_rn(gy, gy));
    aL = mL; aC = mC; aR = mR; aLr = mLr; aRr = mRr;
    mL = bL; mC = bC; mR = bR; mLr = bLr; mRr = bRr;
  };
  // tiles are 4, 8, 16 or 32 rows: chunks of four unrolled rows, border handling only in the first / last
  for (int j = 0; j < tile; j += 4) {
    step(top && j == 0, false);
    step(false, false);
    step(false, false);
    step(false, bot && j + 4 == tile);
  }
  // column sums left-to-right over the tile's lanes (leader = first lane of the tile)
  const float q0 = s0, q1 = s1, q2 = s2, q3 = s3;
  for (int j = 1; j < tile; ++j) {
    s0 = __fadd_rn(s0, __shfl_down_sync(0xffffffffu, q0, j));
    s1 = __fadd_rn(s1, __shfl_down_sync(0xffffffffu, q1, j));
    s2 = __fadd_rn(s2, __shfl_down_sync(0xffffffffu, q2, j));
    s3 = __fadd_rn(s3, __shfl_down_sync(0xffffffffu, q3, j));
  }
  // label counts: five registers of two 16-bit counters, butterfly over the tile's lanes
  uint32_t h[5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
    h[i] = (uint32_t)((cnt >> (12 * i)) & 63ull) | ((uint32_t)((cnt >> (12 * i + 6)) & 63ull) << 16);
  if (!valid) {
#pragma unroll
    for (int i = 0; i < 5; ++i) h[i] = 0;
  }
  for (int o = tile >> 1; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 5; ++i) h[i] += __shfl_xor_sync(0xffffffffu, h[i], o);
  }
  if (valid && (lane & (tile - 1)) == 0) {
    const int t = ty * c.wt + (x >> c.tshift);
    const float rtile2 = 1.0f / (float)(tile * tile);        // tile^2 is a power of two: x * rtile2 == x / tile^2
    float ent = 0.f;
#pragma unroll
    for (int kk = 0; kk < 10; ++kk) {
      const int n = (h[kk >> 1] >> (16 * (kk & 1))) & 0xffff;
      if (lbp_dbg) lbp_dbg[t * 10 + kk] = n;
      const float pr = __fmul_rn((float)n, rtile2);
      ent = __fadd_rn(ent, __fmul_rn(pr, lutp[n]));                // log2(p + 1e-10)
    }
    phis[t * 2 + 0] = __fdiv_rn(-ent, kc::LOG2_10);
    const float mx_ = __fmul_rn(s0, rtile2), mx2 = __fmul_rn(s1, rtile2);
    const float my_ = __fmul_rn(s2, rtile2), my2 = __fmul_rn(s3, rtile2);
    const float vx = fmaxf(__fsub_rn(mx2, __fmul_rn(mx_, mx_)), 0.f);
    const float vy = fmaxf(__fsub_rn(my2, __fmul_rn(my_, my_)), 0.f);
    const float v = __fadd_rn(vx, vy);
    phis[t * 2 + 1] = __fdiv_rn(v, __fadd_rn(v, 1.0f));
  }
}

// ---- T1d: tile activity of the soft mask (quantization.py:224-226) for tile row ty under word k when
//      the adaptive-pool windows are the analyzer's tiles: mean over the tile of sum_c|x| / C.
__device__ __forceinline__ void task_act(const Ctx& c, const float* __restrict__ ap, int W, float fC, float rC,
                                         bool cpow2, int ty, int k, int lane, float* act) {
  const int x = 32 * k + lane;
  const bool valid = x < c.Wc;
  const int xr = valid ? x : c.Wc - 1;
  const int tile = c.tile;
  const float* p = ap + (long long)(ty * tile) * W + xr;
  float s = 0.f;
  for (int j0 = 0; j0 < tile; j0 += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(p + (j0 + u) * W);
#pragma unroll
    for (int u = 0; u < 4; ++u) s = __fadd_rn(s, div_channels(v[u], fC, rC, cpow2));
  }
  const float q = s;
  for (int j = 1; j < tile; ++j) s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, q, j));
  if (valid && (lane & (tile - 1)) == 0)
    act[ty * c.wt + (x >> c.tshift)] = __fmul_rn(s, 1.0f / (float)(tile * tile));   // exact scaling
}

// ---- T2: L1 magnitude of Sobel(BL) (morphology.py:496-497) for rows [r0, min(r0+RT, rend)) -> MAG and
//      the NMS direction bin as two bit planes.
template <int RT>
__device__ __forceinline__ void task_mag(const Ctx& c, int r0, int rend, int k, int lane) {
  const int x = 32 * k + lane;
  const bool valid = x < c.Wc;
  const int xr = valid ? x : c.Wc - 1;
  const float* p = c.BLp + (r0 - 1) * c.bs + xr;
  float aL = p[-1], aC = p[0], aR = p[1];
  p += c.bs;
  float mL = p[-1], mC = p[0], mR = p[1];
#pragma unroll
  for (int j = 0; j < RT; ++j) {
    const int r = r0 + j;
    if (r < rend) {                                   // warp-uniform
      p += c.bs;
      const float bL = p[-1], bC = p[0], bR = p[1];
      float gx, gy;
      sobel3(aL, aC, aR, mL, mR, bL, bC, bR, gx, gy);
      int bin = 0;
      if (valid) {
        c.MAGp[r * c.Wc + x] = __fadd_rn(fabsf(gx), fabsf(gy));
        bin = nms_bin(gx, gy);
      }
      const uint32_t d0 = __ballot_sync(0xffffffffu, bin & 1);
      const uint32_t d1 = __ballot_sync(0xffffffffu, bin & 2);
      if (lane == 0) { c.DIR0[r * c.WW + k] = d0; c.DIR1[r * c.WW + k] = d1; }
      aL = mL; aC = mC; aR = mR;
      mL = bL; mC = bC; mR = bR;
      (void)mC;
    }
  }
}

// ---- T3: non-maximum suppression + double threshold (morphology.py:426-449, 500-502) for band rows
//      [r0, r0+RT) -> STRONG / WEAK words (own planes and every peer's).
template <int RT>
__device__ __forceinline__ void task_nms(const Ctx& c, cg::cluster_group& cl, int r0, int k, int lane, float thr_hi,
                                         float thr_lo) {
  const int x = 32 * k + lane;
  const bool valid = x < c.Wc;
  const int xr = valid ? x : c.Wc - 1;
  const int xl = max(xr - 1, 0), xg = min(xr + 1, c.Wc - 1);
  const float* ra = c.MAGp + max(r0 - 1, 0) * c.Wc;
  float aL = ra[xl], aC = ra[xr], aR = ra[xg];
  const float* rm = c.MAGp + r0 * c.Wc;
  float mL = rm[xl], mC = rm[xr], mR = rm[xg];
#pragma unroll
  for (int j = 0; j < RT; ++j) {
    const int r = r0 + j;
    const float* rb = c.MAGp + min(r + 1, c.Hc - 1) * c.Wc;
    const float bL = rb[xl], bC = rb[xr], bR = rb[xg];
    const uint32_t d0 = (c.DIR0[r * c.WW + k] >> lane) & 1u, d1 = (c.DIR1[r * c.WW + k] >> lane) & 1u;
    // bin 0: (0,+1)/(0,-1)   1: (-1,+1)/(+1,-1)   2: (-1,0)/(+1,0)   3: (-1,-1)/(+1,+1)
    const float n1 = d1 ? (d0 ? aL : aC) : (d0 ? aR : mR);
    const float n2 = d1 ? (d0 ? bR : bC) : (d0 ? bL : mL);
    const float nms = (mC >= n1 && mC >= n2) ? mC : 0.f;
    const uint32_t ws = __ballot_sync(0xffffffffu, valid && nms > thr_hi);
    const uint32_t ww = __ballot_sync(0xffffffffu, valid && nms > thr_lo);
    if (lane < c.ns) {
      const int pr = (c.rank + lane) % c.ns;
      uint32_t* ds = lane == 0 ? c.STRONG : cl.map_shared_rank(c.STRONG, pr);
      uint32_t* dw = lane == 0 ? c.WEAK : cl.map_shared_rank(c.WEAK, pr);
      ds[r * c.WW + k] = ws;
      dw[r * c.WW + k] = ww;
    }
    aL = mL; aC = mC; aR = mR;
    mL = bL; mC = bC; mR = bR;
  }
}

// ---- T4: hysteresis = 8 constrained 3x3 dilations (morphology.py:504-509) of a block of 32 rows held
//      in registers (lane = row R0 + lane, up to 5 words per row); after 8 steps the 16 middle rows are
//      exact, so blocks advance by 16 rows.
__device__ __forceinline__ void task_hysteresis(const Ctx& c, int R0, int lane) {
  const int row = R0 + lane;
  const bool rok = row >= 0 && row < c.Hc;
  uint32_t e[5], w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const bool ok = rok && k < c.WW;
    e[k] = ok ? c.STRONG[row * c.WW + k] : 0u;
    w[k] = ok ? c.WEAK[row * c.WW + k] : 0u;
  }
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    uint32_t h[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const uint32_t cc = e[k];
      const uint32_t l = k > 0 ? e[k - 1] : 0u;
      const uint32_t rn = k < 4 ? e[k + 1] : 0u;
      h[k] = cc | (cc << 1) | (l >> 31) | (cc >> 1) | (rn << 31);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (k < c.WW) {                                   // warp-uniform
        uint32_t up = __shfl_up_sync(0xffffffffu, h[k], 1);
        uint32_t dn = __shfl_down_sync(0xffffffffu, h[k], 1);
        if (lane == 0) up = 0u;
        if (lane == 31) dn = 0u;
        e[k] |= w[k] & (h[k] | up | dn);
      }
    }
  }
  if (lane >= 8 && lane < 24 && row >= c.r_lo && row < c.r_hi) {
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (k < c.WW) c.EDGE[row * c.WW + k] = e[k];
  }
}

// 64 registers per thread (4 CTAs of 256 threads fit the register file): the kernel is issue / latency
// bound, so resident CTAs of other images, scales and steps -- and the HBM-bound K1 / K3 CTAs --
// fill its idle issue slots
__global__ void __launch_bounds__(MORPH_MAX_THREADS, 1024 / MORPH_MAX_THREADS)
morph_fused_kernel(const FusedArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const MorphGeom& g = A.g;
  float* S = reinterpret_cast<float*>(smem_raw);
  const int NW = g.Hc * g.WW;
  float* G = S;
  float* BL = S + g.off_bl;
  float* MAG = S;                                              // aliases G (dead after T1)
  uint32_t* BIN = reinterpret_cast<uint32_t*>(S + g.off_bits);
  uint32_t* DIR0 = BIN + NW;
  uint32_t* DIR1 = DIR0 + NW;
  uint32_t* STRONG = DIR1 + NW;
  uint32_t* WEAK = STRONG + NW;
  uint32_t* EDGE = WEAK + NW;
  float* phi8 = S + g.off_tiles;                               // [ntiles][8]  (16-byte aligned rows)
  float* phis = phi8 + g.ntiles * 8;                           // [ntiles][2]  phi2, phi3
  float* craw_s = phis + g.ntiles * 2;                         // [ntiles] complexity before the bilateral
  float* cfin = craw_s + g.ntiles;                             // [ntiles] complexity
  float* bits_s = cfin + g.ntiles;                             // [ntiles] bits
  float* act_s = bits_s + g.ntiles;                            // [ntiles] tile activity (soft mask)
  float* mt_s = act_s + g.ntiles;                              // [ntiles] tile mask
  int* acc = reinterpret_cast<int*>(mt_s + g.ntiles);          // [ntiles][9] integer tile counts
  int* hist = acc + g.ntiles * 9;                              // [256] whole-image Otsu histogram
  int* hloc = hist + 256;                                      // [256] this CTA's band
  float* red = reinterpret_cast<float*>(hloc + 256);           // [64]
  float* mmx = red + 64;                                       // [16] per-rank min / max
  float* lutn = mmx + 16;                                      // [260] log(N + 1)
  float* lutp = lutn + 260;                                    // [tile^2 + 1 (+pad)] log2(k / tile^2 + 1e-10)
  unsigned long long* lbp_inc = reinterpret_cast<unsigned long long*>(S + g.off_w);   // [256] LBP code -> 1 << 6 * label
  // parameter blocks: the complexity-MLP block (11.5 KB) is staged into the dead plane area right before
  // N1 (L1 starts cold in every launch), the soft-mask block over the dead LUT; the mapper block is only
  // read (through L1) when its step table is absent or invalid
  const float* w_cmlp = A.cmlp;
  const float* w_map = A.mapper;
  float* w_sm = lutn;                                          // soft-mask block (196 floats) staged over the dead LUT

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = blockDim.x, nwarps = NT >> 5;
  const int tile = g.tile, Hc = g.Hc, Wc = g.Wc, WW = g.WW, wt = g.wt;
  cg::cluster_group cl = cg::this_cluster();
  const int ns = g.ns;
  const int rank = ns > 1 ? (int)cl.block_rank() : 0;
  const int b = blockIdx.x / ns;
  const int tr0 = (g.ht * rank) / ns, tr1 = (g.ht * (rank + 1)) / ns;   // own tile rows
  const int r_lo = tr0 * tile, r_hi = tr1 * tile;                       // own pixel rows
  const int t_lo = tr0 * wt, t_hi = tr1 * wt;                           // own tiles
  const int tshift = 31 - __clz(tile);
  const float ntile2 = (float)(tile * tile);
  long long* clk = A.clk;
#define STAGE_CLOCK(k) do { if (clk && tid == 0 && rank == 0) clk[(long long)b * 16 + (k)] = clock64(); } while (0)
  STAGE_CLOCK(0);
  if (ns > 1) cl.barrier_arrive();      // paired with the wait before the first DSMEM store

  {
    const float* src = tile == 4 ? kc::LOG2P_4 : (tile == 8 ? kc::LOG2P_8 : (tile == 16 ? kc::LOG2P_16 : kc::LOG2P_32));
    for (int i = tid; i < 257; i += NT) lutn[i] = __ldg(kc::LOGN1 + i);
    for (int i = tid; i <= tile * tile; i += NT) lutp[i] = __ldg(src + i);
    for (int code = tid; code < 256; code += NT) {              // uniform LBP label (morphology.py:640-650)
      const uint32_t rot = (((uint32_t)code << 1) | ((uint32_t)code >> 7)) & 0xffu;
      const int label = __popc((uint32_t)code ^ rot) <= 2 ? __popc((uint32_t)code) : 9;
      lbp_inc[code] = 1ull << (6 * label);
    }
  }
  // K1 -> K3 hand-off of the per-channel ranges: decode the atomics' integer keys to floats and
  // re-arm the keys for the next sweep (stream order: K1 done, K3 not started)
  int xstep = 0;                                   // multi-GPU: exchange step published by this launch
  if (A.keys && blockIdx.x == 0) {
    for (int ch = tid; ch < g.C; ch += NT) {
      A.packed[ch] = key_float(A.keys[ch]);
      A.packed[g.C + ch] = -key_float(A.keys[g.C + ch]);
      A.keys[ch] = MCAQ_KEY_POS_INF;
      A.keys[g.C + ch] = MCAQ_KEY_NEG_INF;
    }
    if (A.px.world > 1) {
      // multi-GPU: this rank's [min, -max] goes into slot `rank` of every rank's exchange buffer,
      // then the step number is released system-wide; this same CTA merges at the end of the kernel
      // (peer_exchange.cuh)
      int* stepw = reinterpret_cast<int*>(red);
      if (tid == 0) {
        int* ep = reinterpret_cast<int*>(A.px.base[A.px.rank]);
        const int e = *ep + 1;
        *ep = e;
        stepw[0] = e;
      }
      __syncthreads();
      const int e = stepw[0];
      xstep = e;
      for (int p = 0; p < A.px.world; ++p) {
        float* dst = A.px.base[p] + XCHG_SLOTS + (long long)((e & 1) * A.px.world + A.px.rank) * 2 * g.C;
        for (int i = tid; i < 2 * g.C; i += NT) dst[i] = A.packed[i];       // own stores, same thread
      }
      __threadfence_system();
      __syncthreads();
      if (tid < A.px.world)
        st_release_sys(reinterpret_cast<int*>(A.px.base[tid]) + XCHG_FLAGS + 8 * (e & 1) + A.px.rank, e);
      __syncthreads();               // `red` is reused below
    }
  }

  // multi-GPU: the launch's first CTA finishes by waiting for every rank's ranges of this step and
  // writing their minimum over `packed` (what K3 reads).  Only this one CTA per launch ever spins, and it
  // does so after its own work, so a waiting kernel can never starve the peers' publishers of SMs.
  auto merge_ranges = [&]() {
    if (xstep == 0) return;
    const float* local = A.px.base[A.px.rank];
    __syncthreads();
    const bool ok = xchg_wait(local, A.px.world, xstep, tid, A.px.timeout_ns);
    if (!__syncthreads_and(ok)) return;     // timed out: `packed` keeps this rank's ranges, error word set
    for (int i = tid; i < 2 * g.C; i += NT) A.packed[i] = xchg_min(local, A.px.world, 2 * g.C, xstep, i);
  };

  Ctx c;
  c.gs = g.gs; c.bs = g.bs;
  c.Gp = G + (5 - r_lo) * g.gs + 2;
  c.BLp = BL + (2 - r_lo) * g.bs + 1;
  c.MAGp = MAG + (1 - r_lo) * Wc;
  c.Hp = S + g.off_h + (5 - r_lo) * g.hs; c.hs = g.hs;
  c.BIN = BIN; c.DIR0 = DIR0; c.DIR1 = DIR1; c.STRONG = STRONG; c.WEAK = WEAK; c.EDGE = EDGE;
  c.Hc = Hc; c.Wc = Wc; c.WW = WW; c.tile = tile; c.tshift = tshift; c.wt = wt;
  c.r_lo = r_lo; c.r_hi = r_hi; c.ns = ns; c.rank = rank;

  // ---- L: gray = sum / C for rows [r_lo-5, r_hi+5) (zero outside the image and in the pad
  //      columns), min / max over the band; BL and the histograms start as zeros -------------------
  const float* sp = A.sum_plane + (long long)b * g.H * g.W;
  const float fC = (float)g.C;
  const float rC = __frcp_rn(fC);
  const bool cpow2 = (g.C & (g.C - 1)) == 0;
  float lmin = INFINITY, lmax = -INFINITY;
  for (int i = tid; i < 512; i += NT) hist[i] = 0;               // hist + hloc
  {
    const int nbl = (r_hi - r_lo + 4) * g.bs;
    for (int i = tid; i < nbl; i += NT) BL[i] = 0.f;
  }
  // band min / max of gray = sum / C straight from global memory (the plane is L2 resident: K1 just
  // wrote it); the values are formed again, identically, when G is filled below
  const bool vec4 = (g.W & 3) == 0 && (reinterpret_cast<uintptr_t>(sp) & 15) == 0;
  if (vec4 && g.W == Wc) {                                        // band rows are one contiguous run
    const float4* p4 = reinterpret_cast<const float4*>(sp + (long long)r_lo * g.W);
    const int n4 = ((r_hi - r_lo) * Wc) >> 2;
    for (int i = tid; i < n4; i += NT) {
      const float4 v = __ldg(p4 + i);
      const float q0 = div_channels(v.x, fC, rC, cpow2), q1 = div_channels(v.y, fC, rC, cpow2);
      const float q2 = div_channels(v.z, fC, rC, cpow2), q3 = div_channels(v.w, fC, rC, cpow2);
      lmin = fminf(fminf(lmin, fminf(q0, q1)), fminf(q2, q3));
      lmax = fmaxf(fmaxf(lmax, fmaxf(q0, q1)), fmaxf(q2, q3));
    }
  } else {
    for (int r = r_lo + warp; r < r_hi; r += nwarps) {
      for (int x = lane; x < Wc; x += 32) {
        const float q = div_channels(__ldg(sp + (long long)r * g.W + x), fC, rC, cpow2);
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
  if (ns > 1) {                                                   // all-gather the band min / max
    cl.barrier_wait();                                            // every CTA of the cluster is resident
    if (tid < ns) {
      float* m = cl.map_shared_rank(mmx, tid);
      m[2 * rank] = gmin;
      m[2 * rank + 1] = gmax;
    }
    cl.sync();
    gmin = mmx[0]; gmax = mmx[1];
    for (int pr = 1; pr < ns; ++pr) { gmin = fminf(gmin, mmx[2 * pr]); gmax = fmaxf(gmax, mmx[2 * pr + 1]); }
  }
  STAGE_CLOCK(1);
  // ---- N: G = normalised gray (morphology.py:378-383) for rows [r_lo-5, r_hi+5): zero rows outside the
  //      image, zero pad columns -------------------------------------------------------------------
  {
    const float den = __fadd_rn(__fsub_rn(gmax, gmin), 1e-8f);
    const float rden = __frcp_rn(den);
    const int nrows = r_hi - r_lo + 10;
    const int npad = g.gs - Wc;
    for (int i = tid; i < nrows * npad; i += NT) {                // pad columns: 2 left, the rest right
      const int lr = i / npad, k = i - lr * npad;
      G[lr * g.gs + (k < 2 ? k : Wc + k)] = 0.f;
    }
    if (vec4) {
      const int W4 = Wc >> 2;
      const uint32_t m4 = div_magic(W4);
      for (int i = tid; i < nrows * W4; i += NT) {
        const int lr = fast_div(i, m4), x = (i - lr * W4) << 2;
        const int r = r_lo - 5 + lr;
        float* dst = G + lr * g.gs + 2 + x;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (r >= 0 && r < Hc) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(sp + (long long)r * g.W + x));
          o[0] = div_exact(__fsub_rn(div_channels(v.x, fC, rC, cpow2), gmin), den, rden);
          o[1] = div_exact(__fsub_rn(div_channels(v.y, fC, rC, cpow2), gmin), den, rden);
          o[2] = div_exact(__fsub_rn(div_channels(v.z, fC, rC, cpow2), gmin), den, rden);
          o[3] = div_exact(__fsub_rn(div_channels(v.w, fC, rC, cpow2), gmin), den, rden);
          if (A.gray_dbg && r >= r_lo && r < r_hi)
            *reinterpret_cast<float4*>(A.gray_dbg + (long long)b * Hc * Wc + r * Wc + x) = make_float4(o[0], o[1], o[2], o[3]);
        }
        dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3];
      }
    } else {
      for (int lr = warp; lr < nrows; lr += nwarps) {
        const int r = r_lo - 5 + lr;
        const bool rin = r >= 0 && r < Hc;
        float* row = G + lr * g.gs + 2;
        for (int x = lane; x < Wc; x += 32) {
          float v = 0.f;
          if (rin) {
            v = div_exact(__fsub_rn(div_channels(__ldg(sp + (long long)r * g.W + x), fC, rC, cpow2), gmin), den, rden);
            if (A.gray_dbg && r >= r_lo && r < r_hi) A.gray_dbg[(long long)b * Hc * Wc + r * Wc + x] = v;
          }
          row[x] = v;
        }
      }
    }
  }
  __syncthreads();
  STAGE_CLOCK(2);

  // ---- T1: independent warp tasks on G, two phases (the vertical half of the adaptive threshold needs
  //      the horizontally filtered plane) ------------------------------------------------------------
  {
    const int b_lo = max(r_lo - 2, 0), b_hi = min(r_hi + 2, Hc);          // blur rows
    const int nblur = ((b_hi - b_lo + 7) >> 3) * WW;
#ifndef MCAQ_ADAPT_RT
#define MCAQ_ADAPT_RT 8
#endif
    constexpr int RTa = MCAQ_ADAPT_RT;   // output rows per vertical task: RTa + 10 input rows for RTa outputs
    const int nadapt = ((r_hi - r_lo + RTa - 1) / RTa) * WW;
    const int h_lo = max(r_lo - 5, 0), h_hi = min(r_hi + 5, Hc);          // rows of the horizontal pass
    constexpr int HSEG = 40;                                              // columns per horizontal task
    const int nhseg = (Wc + HSEG - 1) / HSEG;
    const int nhor = ((h_hi - h_lo + 31) >> 5) * nhseg;
    const int nlbp = (tr1 - tr0) * WW;
    const bool fast_act = A.softmask && A.abs_plane && g.aligned;
    const int nact = fast_act ? nlbp : 0;
    const float* ap = A.abs_plane ? A.abs_plane + (long long)b * g.H * g.W : nullptr;
    // phase a: horizontal pass | LBP + gradient variance | tile activity
    for (int task = warp; task < nhor + nlbp + nact; task += nwarps) {
      if (task < nhor) {
        const int rg = fast_div(task, g.m_hseg), sg = task - rg * nhseg;
        task_adapt_h(c, h_lo + 32 * rg, h_hi, sg * HSEG, min((sg + 1) * HSEG, Wc), lane);
      } else if (task < nhor + nlbp) {
        const int q = task - nhor;
        const int tyl = fast_div(q, g.m_ww), k = q - tyl * WW;
        task_lbp_var(c, tr0 + tyl, k, lane, lutp, phis, lbp_inc,
                     A.lbp_dbg ? A.lbp_dbg + (long long)b * g.ntiles * 10 : nullptr);
      } else {
        const int q = task - nhor - nlbp;
        const int tyl = fast_div(q, g.m_ww), k = q - tyl * WW;
        task_act(c, ap, g.W, fC, rC, cpow2, tr0 + tyl, k, lane, act_s);
      }
    }
    if (A.softmask && A.abs_plane && !g.aligned)
      softmask_act_generic(ap, g.C, g.H, g.W, g.ht, g.wt, tr0, tr1, act_s);
    __syncthreads();
    // phase b: vertical pass + sign -> BIN | 5x5 blur -> BL + histogram
    for (int task = warp; task < nadapt + nblur; task += nwarps) {
      if (task < nadapt) {
        const int rg = fast_div(task, g.m_ww), k = task - rg * WW;
        task_adaptive<RTa>(c, cl, r_lo + rg * RTa, r_hi, k, lane);
      } else {
        const int q = task - nadapt;
        const int rg = fast_div(q, g.m_ww), k = q - rg * WW;
        task_blur<8>(c, b_lo + rg * 8, b_hi, k, lane, hloc);
      }
    }
  }
  __syncthreads();
  for (int pr = 0; pr < ns; ++pr) {                               // integer counts: order-free, exact
    int* hp = ns > 1 ? cl.map_shared_rank(hist, pr) : hist;
    for (int i = tid; i < 256; i += NT)
      if (hloc[i]) atomicAdd(&hp[i], hloc[i]);
  }
  STAGE_CLOCK(3);

  // ---- T2: gradient magnitude + direction for rows [r_lo-1, r_hi+1) ------------------------------
  // One CTA per image: the histogram is complete after one more barrier, so ONE warp evaluates Otsu's threshold
  // (16 IEEE divisions per lane, an fp64 scan) while the others start on the magnitude tasks, which are handed out
  // through a shared counter; with a cluster the histogram only completes at the cluster barrier below and every
  // warp evaluates the threshold itself.
  int* t2ctr = reinterpret_cast<int*>(red) + 62;
  float* t2thr = red + 60;
  if (ns == 1) {
    if (tid == 0) *t2ctr = 0;
    __syncthreads();
  }
  {
    const int m_lo = max(r_lo - 1, 0), m_hi = min(r_hi + 1, Hc);
    const int nmag = ((m_hi - m_lo + 7) >> 3) * WW;
    if (ns == 1) {
      if (warp == nwarps - 1) {
        float th;
        int ob;
        otsu_warp(hist, lane, th, ob);
        if (lane == 0) { t2thr[0] = th; t2thr[1] = __int_as_float(ob); }
      }
      for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(t2ctr, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= nmag) break;
        const int rg = fast_div(task, g.m_ww), k = task - rg * WW;
        task_mag<8>(c, m_lo + rg * 8, m_hi, k, lane);
      }
    } else {
      for (int task = warp; task < nmag; task += nwarps) {
        const int rg = fast_div(task, g.m_ww), k = task - rg * WW;
        task_mag<8>(c, m_lo + rg * 8, m_hi, k, lane);
      }
    }
  }
  if (ns > 1) cl.sync(); else __syncthreads();       // MAG / DIR complete, whole-image histogram complete
  STAGE_CLOCK(4);

  // ---- T3: Otsu, NMS + double threshold ------------------------------------------------------------
  float thr255;
  int otsu_bin;
  if (ns == 1) { thr255 = t2thr[0]; otsu_bin = __float_as_int(t2thr[1]); }
  else otsu_warp(hist, lane, thr255, otsu_bin);
  {
    const float thr_lo = __fmul_rn(0.5f, thr255);
    const int RTn = tile >= 8 ? 8 : 4;
    const int nnms = ((r_hi - r_lo) / RTn) * WW;
    for (int task = warp; task < nnms; task += nwarps) {
      const int rg = fast_div(task, g.m_ww), k = task - rg * WW;
      if (RTn == 8) task_nms<8>(c, cl, r_lo + rg * 8, k, lane, thr255, thr_lo);
      else task_nms<4>(c, cl, r_lo + rg * 4, k, lane, thr255, thr_lo);
    }
  }
  if (ns > 1) cl.sync(); else __syncthreads();       // strong / weak / adaptive planes complete everywhere
  STAGE_CLOCK(5);

  // ---- T4: hysteresis ------------------------------------------------------------------------------
  {
    const int nblk = (r_hi - r_lo + 15) >> 4;
    for (int task = warp; task < nblk; task += nwarps) task_hysteresis(c, r_lo - 8 + 16 * task, lane);
  }
  for (int i = t_lo * 9 + tid; i < t_hi * 9; i += NT) acc[i] = 0;
  __syncthreads();
  STAGE_CLOCK(6);

  // ---- T5: integer tile counts -------------------------------------------------------------------
  // acc[t][0..8] = edge, area, perim, euler_x4, N_2, N_4, N_8, N_16, N_32
  const int segs = 32 >> tshift;
  const uint32_t segmask = tile == 32 ? 0xffffffffu : ((1u << tile) - 1u);
  for (int i = r_lo * WW + tid; i < r_hi * WW; i += NT) {
    {
      const int r = fast_div(i, g.m_ww), k = i - r * WW;
      const int nbits = min(32, Wc - 32 * k);
      const uint32_t vmask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u);
      const uint32_t e = EDGE[i];
      const uint32_t m = BIN[i];
      uint32_t er = 0xffffffffu;                                    // 3x3 erosion, outside ignored
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int rr = r + dy;
        if (rr < 0 || rr >= Hc) continue;
        const uint32_t cc = BIN[rr * WW + k];
        const uint32_t l = k > 0 ? BIN[rr * WW + k - 1] : 0xffffffffu;
        const uint32_t rn = k + 1 < WW ? BIN[rr * WW + k + 1] : 0xffffffffu;
        const uint32_t left = (cc << 1) | (l >> 31);
        uint32_t right = (cc >> 1) | (rn << 31);
        if (nbits < 32) right |= (1u << (nbits - 1));
        er &= cc & left & right;
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
  }
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
  STAGE_CLOCK(7);

  // ---- phi1, phi4, phi5, interactions (morphology.py:852-864) --------------------------------------
  for (int t = t_lo + tid; t < t_hi; t += NT) {
    const int* a = acc + t * 9;
    const int Sn = g.S;
    float y[5];
#pragma unroll 1
    for (int i = 0; i < Sn; ++i) y[i] = lutn[a[4 + i]];                         // log(N_s + 1)
    float w_sum = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll 1
    for (int i = 0; i < Sn; ++i) {
      w_sum = __fadd_rn(w_sum, kc::FW[i]);
      sx = __fadd_rn(sx, __fmul_rn(kc::FW[i], kc::FLOG[i]));
      sy = __fadd_rn(sy, __fmul_rn(kc::FW[i], y[i]));
    }
    const float x_mean = __fdiv_rn(sx, w_sum), y_mean = __fdiv_rn(sy, w_sum);
    float cov = 0.f, var = 0.f;
#pragma unroll 1
    for (int i = 0; i < Sn; ++i) {
      const float dx = __fsub_rn(kc::FLOG[i], x_mean);
      cov = __fadd_rn(cov, __fmul_rn(__fmul_rn(kc::FW[i], dx), __fsub_rn(y[i], y_mean)));
      var = __fadd_rn(var, __fmul_rn(kc::FW[i], __fmul_rn(dx, dx)));
    }
    float df = -__fdiv_rn(cov, __fadd_rn(var, 1e-12f));
    df = fminf(fmaxf(df, 1.0f), 2.0f);
    const float p1 = Sn < 2 ? 0.5f : __fdiv_rn(df, 2.0f);
    const float p2 = phis[t * 2 + 0], p3 = phis[t * 2 + 1];
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
    float4* ds = reinterpret_cast<float4*>(phi8 + t * 8);
    ds[0] = make_float4(o[0], o[1], o[2], o[3]);
    ds[1] = make_float4(o[4], o[5], o[6], o[7]);
    if (A.phi) {
      float4* dst = reinterpret_cast<float4*>(A.phi + ((long long)b * g.ntiles + t) * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (A.counts_dbg) {
      int* cd = A.counts_dbg + ((long long)b * g.ntiles + t) * 12;
      for (int i = 0; i < 9; ++i) cd[i] = a[i];
      cd[9] = otsu_bin; cd[10] = 0; cd[11] = 0;
    }
  }
  if (A.edge_dbg)
    for (int i = r_lo * WW + tid; i < r_hi * WW; i += NT) A.edge_dbg[(long long)b * NW + i] = EDGE[i];
  if (A.bin_dbg)
    for (int i = r_lo * WW + tid; i < r_hi * WW; i += NT) A.bin_dbg[(long long)b * NW + i] = BIN[i];
  if (!A.cmlp) return;                       // last remote access was before the strong / weak sync
  __syncthreads();                           // phi8 complete; lutn / lutp are dead from here on
  if (A.softmask) copy_params(A.softmask, w_sm, SOFTMASK_SMEM_FLOATS);
  STAGE_CLOCK(8);

  // all-gather of a per-tile array: own tiles -> every peer
  auto publish = [&](float* base) {
    const int n = t_hi - t_lo;
    for (int i = tid; i < n * (ns - 1); i += NT) {
      const int pr = 1 + i / n, o = t_lo + (i - (pr - 1) * n);
      cl.map_shared_rank(base, (rank + pr) % ns)[o] = base[o];
    }
  };

  // ---- N1: complexity MLP (own tiles) -> all-gather -> bilateral (own tiles) -------------------------
  float* scratch = S;                        // the float planes are dead from here on
  const int nt = g.ntiles;
  float* w_cmlp_s = S + nwarps * NET_WARP_SCRATCH;   // staged behind the per-warp scratch
  copy_params(w_cmlp, w_cmlp_s, CMLP_SMEM_FLOATS);
  __syncthreads();
  complexity_mlp_warps(phi8, t_lo, t_hi, w_cmlp_s, scratch, craw_s,
                       A.complexity_raw ? A.complexity_raw + (long long)b * nt : nullptr,
#ifdef MCAQ_T1_PROF
                       nullptr);
#else
                       (clk && rank == 0) ? clk + (long long)b * 16 : nullptr);
#endif
  if (ns > 1) { __syncthreads(); publish(craw_s); cl.sync(); } else __syncthreads();
  STAGE_CLOCK(9);
  bilateral_range(craw_s, g.ht, g.wt, t_lo, t_hi, scratch, cfin, A.complexity ? A.complexity + (long long)b * nt : nullptr);
  STAGE_CLOCK(10);
  if (!A.run_mapper) return;
  // ---- N2: bit mapper (own tiles) --------------------------------------------------------------------
  float* bout = A.bit_map ? A.bit_map + (long long)b * nt : nullptr;
  if (A.linear_mapper) {
    if (ns > 1) { publish(cfin); cl.sync(); }
    mapper_linear_range(cfin, nt, red, t_lo, t_hi, A.temperature, A.use_t, A.continuous, A.lo, A.hi, A.eps_spread,
                        bits_s, bout);
  } else if (A.steps && __ldg(A.steps + MAPPER_STEPS) == 1.f) {
    mapper_steps_range(cfin, t_lo, t_hi, A.steps, A.lo, bits_s, bout);     // staircase of the same network
    __syncthreads();
  } else {
    mapper_mlp_warps(cfin, t_lo, t_hi, w_map, scratch, A.temperature, A.use_t, A.continuous, A.lo, A.hi, bits_s, bout);
    __syncthreads();
  }
  STAGE_CLOCK(11);
  if (!A.softmask || !A.abs_plane) {
    if (ns > 1) cl.sync();                   // nobody leaves while a peer may still write into it
    merge_ranges();
    return;
  }
  // ---- N3: soft mask (quantization.py:213-239): tile head (own tiles) + m rows (own band) -----------
  {
    if (ns > 1) { publish(bits_s); publish(act_s); cl.sync(); }
    float amax = -INFINITY;                                    // every warp: max over all tiles
    for (int t = lane; t < nt; t += 32) amax = fmaxf(amax, act_s[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    float* bn = scratch;                     // [nt]
    float* an = bn + nt;                     // [nt]
    float* cls = an + nt;                    // [own tiles][25]
    softmask_head_range(bits_s, act_s, amax, g.ht, g.wt, t_lo, t_hi, w_sm, bn, an, mt_s,
                        A.mask_tiles ? A.mask_tiles + (long long)b * nt : nullptr);
    if (ns > 1) { publish(mt_s); cl.sync(); }
    float* mo = A.mask + (long long)b * g.H * g.W;
    if (g.aligned) {
      softmask_class_table(mt_s, w_sm, g.ht, g.wt, tile, t_lo, t_hi, cls);
      __syncthreads();
      softmask_plane_from_classes(cls, g.W, g.wt, tile, tshift, t_lo, r_lo, r_hi, mo);
    } else {
      softmask_plane_rows(mt_s, w_sm, g.H, g.W, g.ht, g.wt, (g.H * rank) / ns, (g.H * (rank + 1)) / ns, mo);
    }
  }
  STAGE_CLOCK(12);
  merge_ranges();
}

}  // namespace mcaq

using namespace mcaq;

static long long* g_stage_clk = nullptr;
extern "C" void mcaq_debug_stage_clocks(long long* dev_buf) { g_stage_clk = dev_buf; }

static int max_i(int a, int b) { return a > b ? a : b; }

static int g_force_split = 0;
// debug / tuning: force the number of CTAs per image (0 = automatic)
extern "C" void mcaq_debug_cluster_split(int ns) { g_force_split = ns; }

static int g_sm_count = 0;
static int g_latency_mode = 0;
// split policy: 0 (default) = throughput -- fewest CTAs per image that still fill about half the GPU,
// for callers that keep several launches in flight; 1 = latency -- split every image over as many
// cluster CTAs as fit two per SM, for a serial caller (the forward hook inside a model)
extern "C" void mcaq_morph_policy(int latency) { g_latency_mode = latency ? 1 : 0; }

// CTAs per image: a single CTA per image costs the least SM time (no halo recomputation, no
// cluster barriers), so an image is split by tile rows over a thread-block cluster (portable size
// <= 8) only while the launch would otherwise leave most SMs idle (2 * B * ns <= SM count / 2), and
// never below one 8-row run per CTA
static int pick_split(int B, int ht, int tile) {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sm_count <= 0)
      g_sm_count = 148;
  }
  int ns = 1;
  if (g_force_split == 1 || g_force_split == 2 || g_force_split == 4 || g_force_split == 8) {
    ns = g_force_split;
  } else {
    const int cap = g_latency_mode ? 2 * g_sm_count : g_sm_count / 2;      // CTAs of one launch
    while (ns < 8 && B * ns * 2 <= cap && (ht / (ns * 2)) * tile >= 8) ns *= 2;
  }
  while (ns > ht) ns >>= 1;
  return ns < 1 ? 1 : ns;
}

static int g_force_threads = 0;
// debug / tuning: force the CTA size of the morphology kernel (0 = automatic)
extern "C" void mcaq_debug_morph_threads(int n) { g_force_threads = n; }

static int pick_threads(const MorphGeom& g) {
  if (g_force_threads >= 64 && g_force_threads <= MORPH_MAX_THREADS && g_force_threads % 32 == 0) return g_force_threads;
  const long long band = (long long)g.band_max * g.Wc;
  return band <= 512 ? 128 : 256;
}

// shared-memory layout for a given split; returns bytes of dynamic smem
static long long layout(MorphGeom& g) {
  g.band_max = ((g.ht + g.ns - 1) / g.ns) * g.tile;
  g.threads = pick_threads(g);
  g.max_own = ((g.ht + g.ns - 1) / g.ns) * g.wt;
  // float planes; a partial 8-row run of a stencil task reads at most 7 rows past its plane (results
  // discarded), which stays inside the following plane / the margin after BL
  const long long wG = (long long)(g.band_max + 10) * g.gs;      // also holds MAG ((band+2) * Wc) after T1
  const long long wBL = (long long)(g.band_max + 4) * g.bs + 8LL * g.gs;
  // the same region later holds the per-warp net scratch, bilateral weights, mask classes
  const long long nets = max_i((g.threads / 32) * NET_WARP_SCRATCH + CMLP_SMEM_FLOATS, 2 * g.ntiles + 25 * g.max_own);
  g.off_bl = (int)((wG + 3) & ~3LL);
  g.off_mag = 0;
  g.off_h = (int)((g.off_bl + wBL + 3) & ~3LL);
  const long long wH = (long long)(g.band_max + 10) * g.hs;      // rows [r_lo-5, r_hi+5) of the horizontal pass
  long long bits0 = (g.off_h + wH + 3) & ~3LL;
  if (bits0 < ((nets + 3) & ~3LL)) bits0 = (nets + 3) & ~3LL;
  g.off_bits = (int)bits0;
  const long long tiles0 = (bits0 + 6LL * g.Hc * g.WW + 3) & ~3LL;
  g.off_tiles = (int)tiles0;
  const long long tw = (long long)g.ntiles * (2 + 8 + 5 + 9) + 512 + 64 + 16 + 260 + g.tile * g.tile + 4;
  const long long w0 = (tiles0 + tw + 3) & ~3LL;
  g.off_w = (int)w0;                                   // 256 x 8-byte LBP increment table
  g.words = (int)(w0 + 512);
  return (long long)g.words * 4;
}

// geometry + split + shared-memory layout; returns bytes of dynamic smem or a negative error
static long long plan(MorphGeom& g, int B, int C, int H, int W, int grid_size) {
  g.B = B; g.C = C; g.H = H; g.W = W;
  g.tile = mcaq_tile_size(H, grid_size);
  g.ht = H / g.tile; g.wt = W / g.tile;
  if (g.ht <= 0 || g.wt <= 0 || g.tile > 32 || g.tile < 4) return MCAQ_EINVAL;
  g.Hc = g.ht * g.tile; g.Wc = g.wt * g.tile;
  g.WW = (g.Wc + 31) / 32;
  if (g.WW > 5) return MCAQ_ETOOBIG;          // hysteresis keeps a row of <= 5 words in registers
  g.ntiles = g.ht * g.wt;
  g.S = 0;
  for (int s = 2; s <= g.tile; s <<= 1) g.S++;
  g.aligned = (H == g.Hc && W == g.Wc) ? 1 : 0;
  g.m_ww = 0xffffffffu / (unsigned)g.WW + 1u;
  g.m_hseg = 0xffffffffu / (unsigned)((g.Wc + 39) / 40) + 1u;
  g.gs = g.Wc + 5;          // two zero columns each side; odd, so lane = row accesses are conflict-free
  g.hs = g.Wc + 1;
  g.bs = g.Wc + 2;
  g.ns = pick_split(B, g.ht, g.tile);
  long long bytes = layout(g);
  // planes too large for two CTAs per SM (or for one at all): split the image further -- but not below 32-row
  // bands while one CTA per SM still fits: every band recomputes a 10-row stencil halo, and at 20-row bands that
  // halo doubles the stencil work (measured on the 160 x 160 maps of YOLOv8s @ 1280: 4 CTAs per image beat 8)
  while (bytes > 112 * 1024 && g.ns * 2 <= 8 && g.ns * 2 <= g.ht && !g_force_split &&
         (g.band_max / 2 >= 32 || bytes > 200 * 1024)) {
    g.ns *= 2;
    bytes = layout(g);
  }
  if (bytes > 227 * 1024) return MCAQ_ETOOBIG;
  return bytes;
}

// threads per CTA from the pixels of the largest band
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

// 1 when the fused per-image kernel covers this geometry, 0 when the plane pipeline (morph_planes.cu) must
// take it (more than 160 columns, tiles above 32 pixels, or a plane beyond the shared-memory budget)
extern "C" int mcaq_morph_fits(int B, int C, int H, int W, int grid_size) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0) return 0;
  MorphGeom g;
  return plan(g, B, C, H, W, grid_size) >= 0 ? 1 : 0;
}

extern "C" int mcaq_morph_phi(const float* sum_plane, int B, int C, int H, int W, int grid_size,
                              const float* consts, float* phi, float* gray_dbg, uint32_t* edge_bits_dbg,
                              uint32_t* bin_bits_dbg, int32_t* lbp_hist_dbg, int32_t* counts_dbg, void* stream) {
  (void)consts;   // compiled in (mcaq_consts.cuh); argument kept for ABI stability
  if (!sum_plane || !phi || B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0) return MCAQ_EINVAL;
  FusedArgs A = {};
  const long long smem = plan(A.g, B, C, H, W, grid_size);
  if (smem < 0) return (int)smem;
  const int threads = A.g.threads;
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
  return mcaq_morph_fused_xchg(sum_plane, abs_plane, B, C, H, W, grid_size, keys, packed_ranges, cmlp, mapper,
                               linear_mapper, softmask, temperature, use_temperature, continuous, min_bits,
                               max_bits, eps_spread, phi, complexity, bit_map, mask, nullptr, 0, 1, stream);
}

extern "C" int mcaq_morph_fused_xchg(const float* sum_plane, const float* abs_plane, int B, int C, int H, int W,
                                     int grid_size, int32_t* keys, float* packed_ranges, const float* cmlp,
                                     const float* mapper, int linear_mapper, const float* softmask,
                                     float temperature, int use_temperature, int continuous, float min_bits,
                                     float max_bits, float eps_spread, float* phi, float* complexity,
                                     float* bit_map, float* mask, void* const* xchg_peers, int xchg_rank,
                                     int xchg_world, void* stream) {
  if (!sum_plane || !cmlp || !complexity || !bit_map || B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0)
    return MCAQ_EINVAL;
  if (!mapper && !linear_mapper) return MCAQ_EINVAL;
  if (softmask && (!abs_plane || !mask)) return MCAQ_EINVAL;
  if (keys && !packed_ranges) return MCAQ_EINVAL;
  if (((uintptr_t)cmlp | (uintptr_t)mapper | (uintptr_t)softmask | (uintptr_t)mask) & 15) return MCAQ_EALIGN;
  FusedArgs A = {};
  const long long smem = plan(A.g, B, C, H, W, grid_size);
  if (smem < 0) return (int)smem;
  const int threads = A.g.threads;
  A.sum_plane = sum_plane; A.abs_plane = abs_plane;
  A.keys = keys; A.packed = packed_ranges;
  if (xchg_world > 1) {
    if (!keys || !xchg_peers || xchg_world > XCHG_MAX_RANKS || xchg_rank < 0 || xchg_rank >= xchg_world)
      return MCAQ_EINVAL;
    for (int i = 0; i < xchg_world; ++i) {
      if (!xchg_peers[i]) return MCAQ_EINVAL;
      A.px.base[i] = reinterpret_cast<float*>(xchg_peers[i]);
    }
    A.px.rank = xchg_rank;
    A.px.world = xchg_world;
    A.px.timeout_ns = xchg_timeout_ns();
  }
  A.cmlp = cmlp; A.mapper = mapper; A.softmask = softmask;
  // linear_mapper == 2: MLP mapper whose block is followed by its step table (mcaq_mapper_steps)
  A.steps = (linear_mapper == 2 && mapper && !continuous) ? mapper + MAPPER_SMEM_FLOATS : nullptr;
  if (linear_mapper == 2) linear_mapper = 0;
  A.run_mapper = 1; A.linear_mapper = linear_mapper; A.use_t = use_temperature; A.continuous = continuous;
  A.temperature = temperature; A.lo = min_bits; A.hi = max_bits; A.eps_spread = eps_spread;
  A.phi = phi; A.complexity = complexity; A.bit_map = bit_map; A.mask = mask;
  return launch_fused(A, smem, threads, (cudaStream_t)stream);
}
