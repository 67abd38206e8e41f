// K2 for image-sized planes: phi(8) per tile when the plane does not fit the fused per-image kernel
// (more than 160 columns, or tiles larger than 32 pixels) -- the curriculum scoring path of the reference
// (utils/dataset.py:345-353 -> core/morphology.py:923-937 on raw 640x640 / 1280x1280 images, tile 64 / 128).
//
// Same arithmetic contract as morph_fused.cu (oracle/mcaq_oracle.py), organised as a short pipeline over
// L2-resident planes instead of one CTA per image.  Two quantities are global per image and force the
// split points: the gray min / max (before anything) and the Otsu threshold of the blurred plane (before
// the non-maximum suppression):
//
//   P0 plane_minmax   min / max of sum / C over the cropped plane          -> 2 keys per image
//   PA stage A        per region (one tile, or a 32x32 block of smaller tiles) with a 6-pixel halo of
//                     normalised gray in shared memory: 5x5 blur -> BL plane (global) + Otsu histogram,
//                     11x11 adaptive threshold (separable pass + guard band + literal 121-tap chain inside
//                     it) -> area / perimeter / Euler quads, LBP histograms + Sobel statistics -> phi2, phi3
//   PO otsu           one warp per image                                   -> thresholds
//   PB stage B        per region with a 10-pixel halo of BL: |Sobel| + direction, NMS + double threshold,
//                     8 constrained dilations on bit rows (the 8-pixel halo makes the tile exact), edge and
//                     dyadic box counts -> phi1, phi4, phi5, interactions -> phi
//
// Halo arithmetic: blur 2 + Sobel 1 + NMS 1 + hysteresis 8 = 12 pixels of gray, split here
// as 2 (stage A, around BL pixels) and 10 (stage B, of BL).
#include "morph_common.cuh"

namespace mcaq {

struct PlaneGeom {
  int B, C, H, W, tile, tshift, ht, wt, Hc, Wc, ntiles, S;
  int R, rshift;    // region edge = max(tile, 32), log2
  int nry, nrx;     // regions per image
  int tpr;          // tiles per region edge
};

struct PlaneWs {
  int* mm;          // [B][2]   ordered keys of the gray min / max
  int* hist;        // [B][256] Otsu histogram of the blurred plane
  float* thr;       // [B][2]   thr_hi (x255), thr_lo
  int* otsu_bin;    // [B]
  float* bl;        // [B][Hc][Wc]   255 * blur
  float* tA;        // [B][ntiles][2]  phi2, phi3
  int* tI;          // [B][ntiles][3]  area, perimeter, 4 * Euler mass
};

__global__ void plane_init_kernel(PlaneGeom g, PlaneWs ws) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.B) { ws.mm[2 * i] = MCAQ_KEY_POS_INF; ws.mm[2 * i + 1] = MCAQ_KEY_NEG_INF; }
  if (i < g.B * 256) ws.hist[i] = 0;
}

// P0: grid (row chunks, B)
__global__ void __launch_bounds__(256) plane_minmax_kernel(const float* __restrict__ sum_plane, PlaneGeom g, PlaneWs ws) {
  const int b = blockIdx.y;
  const float fC = (float)g.C, rC = __frcp_rn(fC);
  const bool pow2 = (g.C & (g.C - 1)) == 0;
  const float* sp = sum_plane + (long long)b * g.H * g.W;
  float lmin = INFINITY, lmax = -INFINITY;
  const int r0 = blockIdx.x * 16, r1 = min(r0 + 16, g.Hc);
  for (int r = r0 + (threadIdx.x >> 5); r < r1; r += 8)
    for (int c = threadIdx.x & 31; c < g.Wc; c += 32) {
      const float q = div_channels(__ldg(sp + (long long)r * g.W + c), fC, rC, pow2);
      lmin = fminf(lmin, q);
      lmax = fmaxf(lmax, q);
    }
  const int kmin = __reduce_min_sync(0xffffffffu, float_key(lmin));
  const int kmax = __reduce_max_sync(0xffffffffu, float_key(lmax));
  if ((threadIdx.x & 31) == 0) { atomicMin(ws.mm + 2 * b, kmin); atomicMax(ws.mm + 2 * b + 1, kmax); }
}

// uniform LBP label of an 8-bit code (morphology.py:640-650)
__device__ __forceinline__ int lbp_label(uint32_t code) {
  const uint32_t rot = ((code << 1) | (code >> 7)) & 0xffu;
  return __popc(code ^ rot) <= 2 ? __popc(code) : 9;
}

// ---- stage A ------------------------------------------------------------------------------------------
// shared memory (floats unless noted), R = region edge, HA = 6:
//   Gw  [(R+12)][(R+13)]   normalised gray, window coords [-6, R+6)^2, REPLICATE filled (clamped image coords)
//   Hh  [(R+12)][(R+2)]    horizontal 11-tap pass, rows [-6, R+6), cols [-1, R+1)
//   cs  [4][tpr][R]        column sums of gx, gx^2, gy, gy^2 per tile row
//   Bw  bytes [(R+2)][(R+2)]  adaptive mask on [-1, R+1)^2: 0 / 1, 2 = outside the image
//   lb  int [ntl][10] | ct int [ntl][3] | hs int [256]
constexpr int HA = 6;

__global__ void __launch_bounds__(256)
plane_stage_a_kernel(const float* __restrict__ sum_plane, PlaneGeom g, PlaneWs ws, float* __restrict__ gray_dbg,
                     uint32_t* __restrict__ bin_dbg, int* __restrict__ lbp_dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = g.R, T = g.tile, tpr = g.tpr, ntl = tpr * tpr;
  const int gw = R + 2 * HA + 1, hw = R + 2;
  float* Gw = reinterpret_cast<float*>(smem_raw);
  float* Hh = Gw + (R + 2 * HA) * gw;
  float* cs = Hh + (R + 2 * HA) * hw;
  int* lb = reinterpret_cast<int*>(cs + 4 * tpr * R);
  int* ct = lb + ntl * 10;
  int* hs = ct + ntl * 3;
  unsigned char* Bw = reinterpret_cast<unsigned char*>(hs + 256);
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const int b = blockIdx.y;
  const int ry = blockIdx.x / g.nrx, rx = blockIdx.x - ry * g.nrx;
  const int ry0 = ry * R, rx0 = rx * R;
  const int Hc = g.Hc, Wc = g.Wc;
  const float fC = (float)g.C, rC = __frcp_rn(fC);
  const bool pow2 = (g.C & (g.C - 1)) == 0;
  const float gmin = key_float(ws.mm[2 * b]), gmax = key_float(ws.mm[2 * b + 1]);
  const float den = __fadd_rn(__fsub_rn(gmax, gmin), 1e-8f), rden = __frcp_rn(den);
  const float* sp = sum_plane + (long long)b * g.H * g.W;

  for (int i = tid; i < ntl * 13 + 256; i += NT) lb[i] = 0;          // lb, ct, hs are contiguous
  // ---- S1: window of normalised gray, replicate filled -----------------------------------------------
  {
    const int ww = R + 2 * HA;
    for (int i = tid; i < ww * ww; i += NT) {
      const int wy = i / ww, wx = i - wy * ww;
      const int r = clampi(ry0 + wy - HA, 0, Hc - 1), c = clampi(rx0 + wx - HA, 0, Wc - 1);
      const float q = div_channels(__ldg(sp + (long long)r * g.W + c), fC, rC, pow2);
      const float v = div_exact(__fsub_rn(q, gmin), den, rden);
      Gw[wy * gw + wx] = v;
      if (gray_dbg && wy >= HA && wy < HA + R && wx >= HA && wx < HA + R && ry0 + wy - HA < Hc && rx0 + wx - HA < Wc)
        gray_dbg[((long long)b * Hc + r) * Wc + c] = v;
    }
  }
  __syncthreads();
  const float* G0 = Gw + HA * gw + HA;                 // G0[wy * gw + wx], window coords from (0, 0)
  // ---- S2a: 5x5 blur (zero padding), BL = 255 * blur, Otsu histogram ---------------------------------
  {
    const bool border = ry0 < 2 || rx0 < 2 || ry0 + R + 2 > Hc || rx0 + R + 2 > Wc;
    float* bl = ws.bl + (long long)b * Hc * Wc;
    for (int i = tid; i < R * R; i += NT) {
      const int wy = i >> g.rshift, wx = i & (R - 1);
      const int r = ry0 + wy, c = rx0 + wx;
      if (r >= Hc || c >= Wc) continue;
      const float* p = G0 + (wy - 2) * gw + wx - 2;
      float acc = 0.f;
      if (!border) {
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) acc = fmaf(p[ky * gw + kx], kc::CANNY[ky * 5 + kx], acc);
      } else {
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const bool in = r + ky - 2 >= 0 && r + ky - 2 < Hc && c + kx - 2 >= 0 && c + kx - 2 < Wc;
            acc = fmaf(in ? p[ky * gw + kx] : 0.f, kc::CANNY[ky * 5 + kx], acc);
          }
      }
      bl[(long long)r * Wc + c] = __fmul_rn(acc, 255.f);
      if (acc >= 0.f && acc <= 1.f) atomicAdd(&hs[min((int)__fmul_rn(acc, 256.f), 255)], 1);   // torch.histc
    }
  }
  // ---- S2b: horizontal 11-tap pass (replicate borders are in Gw) --------------------------------------
  for (int i = tid; i < (R + 2 * HA) * hw; i += NT) {
    const int wy = i / hw, wx = i - wy * hw;          // row wy - HA, column wx - 1
    const float* p = Gw + wy * gw + (wx - 1 + HA) - 5;
    float acc = 0.f;
#pragma unroll
    for (int kx = 0; kx < 11; ++kx) acc = fmaf(p[kx], ADAPT1[kx], acc);
    Hh[wy * hw + wx] = acc;
  }
  // ---- S2c: LBP labels + Sobel(G) statistics, one thread per column of a tile row ----------------------
  for (int task = tid; task < tpr * R; task += NT) {
    const int tyl = task >> g.rshift, wx = task & (R - 1);
    const int c = rx0 + wx, r0 = ry0 + tyl * T;
    if (c >= Wc || r0 >= Hc) continue;
    const bool zl = c == 0, zr = c == Wc - 1;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    unsigned long long c0 = 0ull, c1 = 0ull;              // 2 x five 12-bit counters (a column has <= 128 pixels)
    for (int j = 0; j < T; ++j) {
      const int r = r0 + j;
      const float* p = G0 + (tyl * T + j) * gw + wx;
      const float uL = p[-gw - 1], uC = p[-gw], uR = p[-gw + 1], mL = p[-1], mC = p[0], mR = p[1];
      const float dL = p[gw - 1], dC = p[gw], dR = p[gw + 1];
      const uint32_t code = (uint32_t)(uL >= mC) | ((uint32_t)(uC >= mC) << 1) | ((uint32_t)(uR >= mC) << 2) |
                            ((uint32_t)(mR >= mC) << 3) | ((uint32_t)(dR >= mC) << 4) | ((uint32_t)(dC >= mC) << 5) |
                            ((uint32_t)(dL >= mC) << 6) | ((uint32_t)(mL >= mC) << 7);
      const int label = lbp_label(code);
      if (label < 5) c0 += 1ull << (12 * label); else c1 += 1ull << (12 * (label - 5));
      // Sobel with ZERO padding: neighbours outside the image count as 0
      const bool zt = r == 0, zb = r == Hc - 1;
      float gx, gy;
      sobel3((zt || zl) ? 0.f : uL, zt ? 0.f : uC, (zt || zr) ? 0.f : uR, zl ? 0.f : mL, zr ? 0.f : mR,
             (zb || zl) ? 0.f : dL, zb ? 0.f : dC, (zb || zr) ? 0.f : dR, gx, gy);
      s0 = __fadd_rn(s0, gx);
      s1 = __fadd_rn(s1, __fmul_rn(gx, gx));
      s2 = __fadd_rn(s2, gy);
      s3 = __fadd_rn(s3, __fmul_rn(gy, gy));
    }
    float* q = cs + tyl * R + wx;
    q[0] = s0; q[tpr * R] = s1; q[2 * tpr * R] = s2; q[3 * tpr * R] = s3;
    int* h = lb + (tyl * tpr + (wx >> g.tshift)) * 10;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int n0 = (int)((c0 >> (12 * k)) & 0xfffull), n1 = (int)((c1 >> (12 * k)) & 0xfffull);
      if (n0) atomicAdd(h + k, n0);
      if (n1) atomicAdd(h + 5 + k, n1);
    }
  }
  __syncthreads();
  // ---- S3: vertical pass + sign -> Bw on [-1, R+1)^2 -----------------------------------------------------
  for (int i = tid; i < hw * hw; i += NT) {
    const int wy = i / hw, wx = i - wy * hw;          // window coords (wy - 1, wx - 1)
    const int r = ry0 + wy - 1, c = rx0 + wx - 1;
    unsigned char out = 2;
    if (r >= 0 && r < Hc && c >= 0 && c < Wc) {
      const float* hp = Hh + wy * hw + wx;            // row index (wy - 1) - 5 + HA = wy
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 11; ++ky) acc = fmaf(hp[ky * hw], ADAPT1[ky], acc);
      const float* gp = G0 + (wy - 1) * gw + (wx - 1);
      const float ctr = __fmul_rn(gp[0], 255.f);
      const float d = __fsub_rn(ctr, __fsub_rn(__fmul_rn(acc, 255.f), 2.0f));
      bool bit = d > 0.f;
      if (fabsf(d) <= ADAPT_GUARD) {                  // literal reference chain (morphology.py:550-573)
        float e = 0.f;
#pragma unroll 1
        for (int ky = 0; ky < 11; ++ky)
#pragma unroll
          for (int kx = 0; kx < 11; ++kx)
            e = fmaf(__fmul_rn(gp[(ky - 5) * gw + kx - 5], 255.f), __fmul_rn(ADAPT1[ky], ADAPT1[kx]), e);
        bit = ctr > __fsub_rn(e, 2.0f);
      }
      out = bit ? 1 : 0;
    }
    Bw[i] = out;
  }
  __syncthreads();
  // ---- S4: area / perimeter / Euler quads per tile -------------------------------------------------------
  {
    const unsigned char* B0 = Bw + hw + 1;            // B0[wy * hw + wx], window coords from (0, 0)
    uint32_t* bd = bin_dbg ? bin_dbg + (long long)b * Hc * ((Wc + 31) >> 5) : nullptr;
    for (int i = tid; i < R * R; i += NT) {           // R is a multiple of 32: a warp walks 32 pixels of one row
      const int wy = i >> g.rshift, wx = i & (R - 1);
      const int r = ry0 + wy, c = rx0 + wx;
      const bool in = r < Hc && c < Wc;
      int na = 0, np_ = 0, e4 = 0;
      if (in) {
        const unsigned char* p = B0 + wy * hw + wx;
        const int m = p[0] == 1;
        bool er = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) er = er && (p[dy * hw + dx] != 0);      // outside (2) is ignored
        na = m;
        np_ = m && !er;
        const int qa = p[-hw - 1] == 1, qb = p[-hw] == 1, qc = p[-1] == 1, qd = m;
        const int idx = qa + 2 * qb + 4 * qc + 8 * qd;
        // Q1 = {1,2,4,8}, Q3 = {7,11,13,14}, QD = {6,9}  (morphology.py:694-702)
        e4 = ((0x0116 >> idx) & 1) - ((0x6880 >> idx) & 1) - 2 * ((0x0240 >> idx) & 1);
      }
      if (bd) {
        const uint32_t word = __ballot_sync(0xffffffffu, in && na);
        if (lane == 0 && r < Hc && c < Wc) bd[(long long)r * ((Wc + 31) >> 5) + (c >> 5)] = word;
      }
      if (T >= 32) {                                  // the warp's 32 pixels lie in one tile
        na = __reduce_add_sync(0xffffffffu, na);
        np_ = __reduce_add_sync(0xffffffffu, np_);
        e4 = __reduce_add_sync(0xffffffffu, e4);
        if (lane == 0) {
          int* q = ct + ((wy >> g.tshift) * tpr + (wx >> g.tshift)) * 3;
          if (na) atomicAdd(q, na);
          if (np_) atomicAdd(q + 1, np_);
          if (e4) atomicAdd(q + 2, e4);
        }
      } else if (in) {
        int* q = ct + ((wy >> g.tshift) * tpr + (wx >> g.tshift)) * 3;
        if (na) atomicAdd(q, na);
        if (np_) atomicAdd(q + 1, np_);
        if (e4) atomicAdd(q + 2, e4);
      }
    }
  }
  __syncthreads();
  // ---- outputs: histogram, per-tile phi2 / phi3 / counts ---------------------------------------------------
  for (int i = tid; i < 256; i += NT)
    if (hs[i]) atomicAdd(ws.hist + b * 256 + i, hs[i]);
  for (int tl = tid; tl < ntl; tl += NT) {
    const int tyl = tl / tpr, txl = tl - tyl * tpr;
    const int ty = ry * tpr + tyl, tx = rx * tpr + txl;
    if (ty >= g.ht || tx >= g.wt) continue;
    const int t = ty * g.wt + tx;
    const float rtile2 = 1.0f / (float)(T * T);                       // power of two: exact scaling
    float ent = 0.f;
    for (int k = 0; k < 10; ++k) {
      const int n = lb[tl * 10 + k];
      if (lbp_dbg) lbp_dbg[((long long)b * g.ntiles + t) * 10 + k] = n;
      const float pr = __fmul_rn((float)n, rtile2);
      const float lg = (float)log2((double)__fadd_rn(pr, 1e-10f));    // log2(p + 1e-10), fp64 rounded once
      ent = __fadd_rn(ent, __fmul_rn(pr, lg));
    }
    float s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                      // column sums left to right
      const float* q = cs + (k * tpr + tyl) * R + txl * T;
      float a = q[0];
      for (int j = 1; j < T; ++j) a = __fadd_rn(a, q[j]);
      s[k] = a;
    }
    const float mx_ = __fmul_rn(s[0], rtile2), mx2 = __fmul_rn(s[1], rtile2);
    const float my_ = __fmul_rn(s[2], rtile2), my2 = __fmul_rn(s[3], rtile2);
    const float vx = fmaxf(__fsub_rn(mx2, __fmul_rn(mx_, mx_)), 0.f);
    const float vy = fmaxf(__fsub_rn(my2, __fmul_rn(my_, my_)), 0.f);
    const float v = __fadd_rn(vx, vy);
    float* oa = ws.tA + ((long long)b * g.ntiles + t) * 2;
    oa[0] = __fdiv_rn(-ent, kc::LOG2_10);
    oa[1] = __fdiv_rn(v, __fadd_rn(v, 1.0f));
    int* oi = ws.tI + ((long long)b * g.ntiles + t) * 3;
    oi[0] = ct[tl * 3]; oi[1] = ct[tl * 3 + 1]; oi[2] = ct[tl * 3 + 2];
  }
}

__global__ void plane_otsu_kernel(PlaneGeom g, PlaneWs ws) {
  const int b = blockIdx.x, lane = threadIdx.x;
  float thr255;
  int bin;
  otsu_warp(ws.hist + b * 256, lane, thr255, bin);
  if (lane == 0) {
    ws.thr[2 * b] = thr255;
    ws.thr[2 * b + 1] = __fmul_rn(0.5f, thr255);
    ws.otsu_bin[b] = bin;
  }
}

// ---- stage B ------------------------------------------------------------------------------------------
// shared memory, R = region edge, HB = 10:
//   BLw [(R+20)][(R+21)]  255 * blur on [-10, R+10)^2, zero outside the image
//   MGw [(R+18)][(R+19)]  |Sobel| (L1) on [-9, R+9)^2          DRw bytes, same extent: NMS direction bin
//   SW, WK, E0, E1 uint32 [(R+16)][WB]  strong / weak / edge bit rows on [-8, R+8)^2, WB = ceil((R+16)/32)
//   Et uint32 [R][R/32]   edge bits of the region, word aligned    nbox int [ntl][8]   ecnt int [ntl]
constexpr int HB = 10;

__global__ void __launch_bounds__(256)
plane_stage_b_kernel(PlaneGeom g, PlaneWs ws, float* __restrict__ phi, uint32_t* __restrict__ edge_dbg,
                     int* __restrict__ counts_dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = g.R, T = g.tile, tpr = g.tpr, ntl = tpr * tpr;
  const int bw = R + 2 * HB + 1, mw = R + 2 * HB - 1, ew = R + 16, WB = (ew + 31) >> 5, RW = R >> 5;
  float* BLw = reinterpret_cast<float*>(smem_raw);
  float* MGw = BLw + (R + 2 * HB) * bw;
  uint32_t* SW = reinterpret_cast<uint32_t*>(MGw + (R + 2 * HB - 2) * mw);
  uint32_t* WK = SW + ew * WB;
  uint32_t* E0 = WK + ew * WB;
  uint32_t* E1 = E0 + ew * WB;
  uint32_t* Et = E1 + ew * WB;
  int* nbox = reinterpret_cast<int*>(Et + R * RW);
  int* ecnt = nbox + ntl * 8;
  unsigned char* DRw = reinterpret_cast<unsigned char*>(ecnt + ntl);
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  const int b = blockIdx.y;
  const int ry = blockIdx.x / g.nrx, rx = blockIdx.x - ry * g.nrx;
  const int ry0 = ry * R, rx0 = rx * R;
  const int Hc = g.Hc, Wc = g.Wc;
  const float* bl = ws.bl + (long long)b * Hc * Wc;
  const float thr_hi = ws.thr[2 * b], thr_lo = ws.thr[2 * b + 1];

  for (int i = tid; i < ntl * 9; i += NT) nbox[i] = 0;                 // nbox + ecnt
  // ---- S1: BL window, zero outside the image ---------------------------------------------------------
  {
    const int ww = R + 2 * HB;
    for (int i = tid; i < ww * ww; i += NT) {
      const int wy = i / ww, wx = i - wy * ww;
      const int r = ry0 + wy - HB, c = rx0 + wx - HB;
      BLw[wy * bw + wx] = (r >= 0 && r < Hc && c >= 0 && c < Wc) ? __ldg(bl + (long long)r * Wc + c) : 0.f;
    }
  }
  __syncthreads();
  // ---- S2: L1 magnitude of Sobel(BL) and direction bin on [-9, R+9)^2 (morphology.py:496-497, 430-444) ---
  {
    const int ww = R + 2 * HB - 2;
    for (int i = tid; i < ww * ww; i += NT) {
      const int wy = i / ww, wx = i - wy * ww;          // window coords (wy - 9, wx - 9)
      const float* p = BLw + (wy + 1) * bw + wx + 1;
      float gx, gy;
      sobel3(p[-bw - 1], p[-bw], p[-bw + 1], p[-1], p[1], p[bw - 1], p[bw], p[bw + 1], gx, gy);
      MGw[wy * mw + wx] = __fadd_rn(fabsf(gx), fabsf(gy));
      DRw[wy * mw + wx] = (unsigned char)nms_bin(gx, gy);
    }
  }
  __syncthreads();
  // ---- S3: NMS (replicate-padded neighbours) + double threshold -> bit rows on [-8, R+8)^2 ----------------
  for (int task = warp; task < ew * WB; task += nwarps) {
    const int wy = task / WB, k = task - wy * WB;
    const int wx = 32 * k + lane;                        // window coords (wy - 8, wx - 8)
    const int r = ry0 + wy - 8, c = rx0 + wx - 8;
    bool st = false, wk = false;
    if (wx < ew && r >= 0 && r < Hc && c >= 0 && c < Wc) {
      // neighbour coordinates clamped to the image, then to window indices of MGw (origin -9)
      const int ru = max(r - 1, 0) - ry0 + 9, rd = min(r + 1, Hc - 1) - ry0 + 9, rm = r - ry0 + 9;
      const int cl = max(c - 1, 0) - rx0 + 9, cr = min(c + 1, Wc - 1) - rx0 + 9, cm = c - rx0 + 9;
      const float mC = MGw[rm * mw + cm];
      const int bin = DRw[rm * mw + cm];
      // bin 0: (0,+1)/(0,-1)   1: (-1,+1)/(+1,-1)   2: (-1,0)/(+1,0)   3: (-1,-1)/(+1,+1)
      float n1, n2;
      if (bin == 0) { n1 = MGw[rm * mw + cr]; n2 = MGw[rm * mw + cl]; }
      else if (bin == 1) { n1 = MGw[ru * mw + cr]; n2 = MGw[rd * mw + cl]; }
      else if (bin == 2) { n1 = MGw[ru * mw + cm]; n2 = MGw[rd * mw + cm]; }
      else { n1 = MGw[ru * mw + cl]; n2 = MGw[rd * mw + cr]; }
      const float nms = (mC >= n1 && mC >= n2) ? mC : 0.f;
      st = nms > thr_hi;
      wk = nms > thr_lo;
    }
    const uint32_t s_ = __ballot_sync(0xffffffffu, st), w_ = __ballot_sync(0xffffffffu, wk);
    if (lane == 0) { SW[task] = s_; WK[task] = w_; E0[task] = s_; }
  }
  __syncthreads();
  // ---- S4: 8 constrained dilations (morphology.py:504-509) on the bit rows ---------------------------------
  {
    uint32_t* src = E0;
    uint32_t* dst = E1;
    for (int it = 0; it < 8; ++it) {
      for (int i = tid; i < ew * WB; i += NT) {
        const int wy = i / WB, k = i - wy * WB;
        uint32_t h = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int y = wy + dy;
          if (y < 0 || y >= ew) continue;
          const uint32_t cc = src[y * WB + k];
          const uint32_t l = k > 0 ? src[y * WB + k - 1] : 0u;
          const uint32_t rn = k + 1 < WB ? src[y * WB + k + 1] : 0u;
          h |= cc | (cc << 1) | (l >> 31) | (cc >> 1) | (rn << 31);
        }
        dst[i] = src[i] | (WK[i] & h);
      }
      __syncthreads();
      uint32_t* t = src; src = dst; dst = t;
    }
    // after 8 swaps the result is back in E0 (src)
    // ---- region rows, word aligned: Et[y][k] = window bits [8 + 32k, 8 + 32k + 32) of row y + 8 ----------
    uint32_t* ed = edge_dbg ? edge_dbg + (long long)b * Hc * ((Wc + 31) >> 5) : nullptr;
    for (int i = tid; i < R * RW; i += NT) {
      const int y = i / RW, k = i - y * RW;
      const uint32_t lo = src[(y + 8) * WB + k], hi = k + 1 < WB ? src[(y + 8) * WB + k + 1] : 0u;
      const uint32_t w = (lo >> 8) | (hi << 24);
      Et[i] = w;
      if (ed && ry0 + y < Hc && rx0 + 32 * k < Wc) ed[(long long)(ry0 + y) * ((Wc + 31) >> 5) + ((rx0 >> 5) + k)] = w;
    }
  }
  __syncthreads();
  // ---- S5: edge count and dyadic box counts per tile (morphology.py:595-601) -------------------------------
  {
    const uint32_t segmask = T >= 32 ? 0xffffffffu : ((1u << T) - 1u);
    const int wpt = T >= 32 ? T >> 5 : 1;                              // words per tile row
    // tasks: (tile, scale index, block of s rows); scale index S = the edge count (s = 1 row at a time)
    for (int tl = 0; tl < ntl; ++tl) {
      const int tyl = tl / tpr, txl = tl - tyl * tpr;
      const int k0 = (txl * T) >> 5, sh = (txl * T) & 31, y00 = tyl * T;
      for (int y = tid; y < T; y += NT) {                              // edge count
        int n = 0;
        for (int q = 0; q < wpt; ++q) n += __popc((Et[(y00 + y) * RW + k0 + q] >> sh) & segmask);
        if (n) atomicAdd(ecnt + tl, n);
      }
      for (int sidx = 0; sidx < g.S; ++sidx) {
        const int s = 2 << sidx;
        const int nblk = T / s;                                        // row blocks of s rows
        for (int yb = tid; yb < nblk; yb += NT) {
          int n = 0;
          if (s <= 32) {
            uint32_t colmask = 0;
            for (int q = 0; q < 32 && q < T; q += s) colmask |= 1u << q;
            for (int q = 0; q < wpt; ++q) {
              uint32_t o = 0;
              for (int y = 0; y < s; ++y) o |= (Et[(y00 + yb * s + y) * RW + k0 + q] >> sh) & segmask;
              for (int d = 1; d < s; d <<= 1) o |= o >> d;
              n += __popc(o & colmask);
            }
          } else {                                                     // s = 64, 128: whole words
            const int wps = s >> 5;
            for (int q = 0; q < wpt; q += wps) {
              uint32_t o = 0;
              for (int y = 0; y < s; ++y)
                for (int u = 0; u < wps; ++u) o |= Et[(y00 + yb * s + y) * RW + k0 + q + u];
              n += o != 0u;
            }
          }
          if (n) atomicAdd(nbox + tl * 8 + sidx, n);
        }
      }
    }
  }
  __syncthreads();
  // ---- S6: phi1, phi4, phi5, interactions (morphology.py:852-864) --------------------------------------------
  for (int tl = tid; tl < ntl; tl += NT) {
    const int tyl = tl / tpr, txl = tl - tyl * tpr;
    const int ty = ry * tpr + tyl, tx = rx * tpr + txl;
    if (ty >= g.ht || tx >= g.wt) continue;
    const int t = ty * g.wt + tx;
    const int Sn = g.S;
    float y[8], fl[8], fw[8];
    for (int i = 0; i < Sn; ++i) {
      y[i] = (float)log((double)__fadd_rn((float)nbox[tl * 8 + i], 1.0f));          // log(N_s + 1)
      fl[i] = (float)log((double)(float)(2 << i));                                   // log s
      fw[i] = (float)exp((double)__fmul_rn(-0.1f, (float)i));                        // exp(-0.1 i)
    }
    float w_sum = 0.f, sx = 0.f, sy = 0.f;
    for (int i = 0; i < Sn; ++i) {
      w_sum = __fadd_rn(w_sum, fw[i]);
      sx = __fadd_rn(sx, __fmul_rn(fw[i], fl[i]));
      sy = __fadd_rn(sy, __fmul_rn(fw[i], y[i]));
    }
    const float x_mean = __fdiv_rn(sx, w_sum), y_mean = __fdiv_rn(sy, w_sum);
    float cov = 0.f, var = 0.f;
    for (int i = 0; i < Sn; ++i) {
      const float dx = __fsub_rn(fl[i], x_mean);
      cov = __fadd_rn(cov, __fmul_rn(__fmul_rn(fw[i], dx), __fsub_rn(y[i], y_mean)));
      var = __fadd_rn(var, __fmul_rn(fw[i], __fmul_rn(dx, dx)));
    }
    float df = -__fdiv_rn(cov, __fadd_rn(var, 1e-12f));
    df = fminf(fmaxf(df, 1.0f), 2.0f);
    const float p1 = Sn < 2 ? 0.5f : __fdiv_rn(df, 2.0f);
    const float* ta = ws.tA + ((long long)b * g.ntiles + t) * 2;
    const int* ti = ws.tI + ((long long)b * g.ntiles + t) * 3;
    const float p2 = ta[0], p3 = ta[1];
    const float p4 = __fdiv_rn((float)ecnt[tl], (float)(T * T));
    const float area = (float)ti[0], perim = (float)ti[1];
    float ic = __fdiv_rn(__fmul_rn(perim, perim), __fadd_rn(__fmul_rn(kc::FOUR_PI, area), 1e-6f));
    const float K = fmaxf(rintf(__fdiv_rn((float)ti[2], 4.0f)), 1.0f);
    ic = __fdiv_rn(ic, K);
    float p5 = __fsub_rn(1.0f, __fdiv_rn(1.0f, fmaxf(ic, 1.0f)));
    if (ti[0] <= 0) p5 = 0.f;
    float4* dst = reinterpret_cast<float4*>(phi + ((long long)b * g.ntiles + t) * 8);
    dst[0] = make_float4(p1, p2, p3, p4);
    dst[1] = make_float4(p5, __fmul_rn(p1, p2), __fmul_rn(p3, p3), __fsqrt_rn(__fadd_rn(__fmul_rn(p4, p5), 1e-12f)));
    if (counts_dbg) {
      int* cd = counts_dbg + ((long long)b * g.ntiles + t) * 12;
      cd[0] = ecnt[tl]; cd[1] = ti[0]; cd[2] = ti[1]; cd[3] = ti[2];
      for (int i = 0; i < 7; ++i) cd[4 + i] = i < Sn ? nbox[tl * 8 + i] : 0;
      cd[11] = ws.otsu_bin[b];
    }
  }
}

static int ilog2(int v) { int s = 0; while ((1 << (s + 1)) <= v) ++s; return s; }

static int plan_planes(PlaneGeom& g, int B, int C, int H, int W, int grid_size) {
  g.B = B; g.C = C; g.H = H; g.W = W;
  g.tile = mcaq_tile_size(H, grid_size);
  if (g.tile < 4 || g.tile > 128) return MCAQ_ETOOBIG;
  g.tshift = ilog2(g.tile);
  g.ht = H / g.tile; g.wt = W / g.tile;
  if (g.ht <= 0 || g.wt <= 0) return MCAQ_EINVAL;
  g.Hc = g.ht * g.tile; g.Wc = g.wt * g.tile;
  g.ntiles = g.ht * g.wt;
  g.S = 0;
  for (int s = 2; s <= g.tile; s <<= 1) g.S++;
  g.R = g.tile >= 32 ? g.tile : 32;
  g.rshift = ilog2(g.R);
  g.tpr = g.R / g.tile;
  g.nry = (g.Hc + g.R - 1) / g.R;
  g.nrx = (g.Wc + g.R - 1) / g.R;
  return 0;
}

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// workspace carve-up; returns total bytes
static size_t carve(const PlaneGeom& g, unsigned char* base, PlaneWs& ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) { unsigned char* p = base ? base + off : nullptr; off += align16(bytes); return p; };
  ws.mm = reinterpret_cast<int*>(take((size_t)g.B * 2 * 4));
  ws.hist = reinterpret_cast<int*>(take((size_t)g.B * 256 * 4));
  ws.thr = reinterpret_cast<float*>(take((size_t)g.B * 2 * 4));
  ws.otsu_bin = reinterpret_cast<int*>(take((size_t)g.B * 4));
  ws.bl = reinterpret_cast<float*>(take((size_t)g.B * g.Hc * g.Wc * 4));
  ws.tA = reinterpret_cast<float*>(take((size_t)g.B * g.ntiles * 2 * 4));
  ws.tI = reinterpret_cast<int*>(take((size_t)g.B * g.ntiles * 3 * 4));
  return off;
}

}  // namespace mcaq

using namespace mcaq;

// bytes of workspace mcaq_morph_phi_planes needs for this geometry (negative: MCAQ_E*)
extern "C" long long mcaq_morph_planes_workspace(int B, int C, int H, int W, int grid_size) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0) return MCAQ_EINVAL;
  PlaneGeom g;
  const int rc = plan_planes(g, B, C, H, W, grid_size);
  if (rc) return rc;
  PlaneWs ws;
  return (long long)carve(g, nullptr, ws);
}

extern "C" int mcaq_morph_phi_planes(const float* sum_plane, int B, int C, int H, int W, int grid_size,
                                     void* workspace, long long workspace_bytes, float* phi, float* gray_dbg,
                                     uint32_t* edge_bits_dbg, uint32_t* bin_bits_dbg, int32_t* lbp_hist_dbg,
                                     int32_t* counts_dbg, void* stream) {
  if (!sum_plane || !phi || !workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0 || grid_size <= 0) return MCAQ_EINVAL;
  if (((uintptr_t)workspace | (uintptr_t)phi) & 15) return MCAQ_EALIGN;
  PlaneGeom g;
  int rc = plan_planes(g, B, C, H, W, grid_size);
  if (rc) return rc;
  PlaneWs ws;
  if ((long long)carve(g, reinterpret_cast<unsigned char*>(workspace), ws) > workspace_bytes) return MCAQ_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = g.R, ntl = g.tpr * g.tpr;
  const size_t smem_a = ((size_t)(R + 2 * HA) * (R + 2 * HA + 1) + (size_t)(R + 2 * HA) * (R + 2) + (size_t)4 * g.tpr * R) * 4 +
                        ((size_t)ntl * 13 + 256) * 4 + (size_t)(R + 2) * (R + 2);
  const int ew = R + 16, WB = (ew + 31) >> 5;
  const size_t smem_b = ((size_t)(R + 2 * HB) * (R + 2 * HB + 1) + (size_t)(R + 2 * HB - 2) * (R + 2 * HB - 1)) * 4 +
                        ((size_t)4 * ew * WB + (size_t)R * (R >> 5) + (size_t)ntl * 9) * 4 +
                        (size_t)(R + 2 * HB - 2) * (R + 2 * HB - 1);
  if (smem_a > 227 * 1024 || smem_b > 227 * 1024) return MCAQ_ETOOBIG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(plane_stage_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(plane_stage_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  plane_init_kernel<<<(B * 256 + 255) / 256, 256, 0, st>>>(g, ws);
  plane_minmax_kernel<<<dim3((unsigned)((g.Hc + 15) / 16), (unsigned)B), 256, 0, st>>>(sum_plane, g, ws);
  const dim3 grid((unsigned)(g.nry * g.nrx), (unsigned)B);
  plane_stage_a_kernel<<<grid, 256, smem_a, st>>>(sum_plane, g, ws, gray_dbg, bin_bits_dbg, lbp_hist_dbg);
  plane_otsu_kernel<<<B, 32, 0, st>>>(g, ws);
  plane_stage_b_kernel<<<grid, 256, smem_b, st>>>(g, ws, phi, edge_bits_dbg, counts_dbg);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
