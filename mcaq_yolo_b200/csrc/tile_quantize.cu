// K3: fused bit-map quantize / de-quantize (+ soft-mask multiply) over an NCHW feature map,
// plus the fractional-bit training forward and its straight-through backward.
//
// Thread = one 16-byte pixel vector (4 fp32 / 8 bf16 pixels of one row) x a chunk of channels:
// the tile's bit-width, the mask vector and the (qmin, qmax) pair are computed once per thread
// and reused for every channel; per channel the thread does one LDG.128, a float2 table lookup
// {scale, zero_point}[bits][c] (L1-resident, warp-uniform when the warp's pixels share a
// bit-width), the IEEE quantise/de-quantise chain and one STG.128.
//
// HBM traffic: reads x once and writes y once (2*s bytes per element); mask / bit map / table
// are 1/C of that and stay in L1/L2.
#include "common.cuh"
#include "peer_exchange.cuh"
#include "tile_quantize.cuh"

namespace mcaq {

constexpr int QCHUNK = 16;   // channels per thread
constexpr int QUNROLL = 4;   // independent loads in flight per thread
constexpr int QTHREADS = 128;

template <typename T, int VEC>
__device__ __forceinline__ void load_elems(const T* p, float* f, bool inplace) {
  if (VEC == 1) { f[0] = Elem<T>::load1(p); return; }
  uint4 r = inplace ? ldg_plain(p) : ldg_stream(p);
  Elem<T>::unpack(r, f);
}
template <typename T, int VEC>
__device__ __forceinline__ void store_elems(T* p, const float* f) {
  if (VEC == 1) { Elem<T>::store1(p, f[0]); return; }
  stg_stream(p, Elem<T>::pack(f));
}

// per-thread pixel context: bit index per element (0..6), uniform flag, mask values
template <int VEC>
struct PixCtx {
  int bidx[VEC];
  float m[VEC];
  bool uniform;
};

template <int VEC, bool HAS_MASK>
__device__ __forceinline__ void make_ctx(const QGeom& g, int b, int pix, const float* __restrict__ bit_map,
                                         const float* __restrict__ mask, PixCtx<VEC>& ctx) {
  const int h = pix / g.W;
  const int w0 = pix - h * g.W;
  const int ty = nearest_src(h, g.sy, g.Ht);
  const float* brow = bit_map + ((long long)b * g.Ht + ty) * g.Wt;
  ctx.uniform = true;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    // VEC > 1 requires W % VEC == 0, so the vector never leaves row h
    const int tx = nearest_src(w0 + e, g.sx, g.Wt);
    float bf = rintf(__ldg(brow + tx));
    bf = fminf(fmaxf(bf, 2.f), 8.f);
    ctx.bidx[e] = (int)bf - 2;
    if (ctx.bidx[e] != ctx.bidx[0]) ctx.uniform = false;
    ctx.m[e] = 1.f;
  }
  if (HAS_MASK) {
    const float* mp = mask + (long long)b * g.HW + pix;
    if constexpr (VEC == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(mp));
      ctx.m[0] = v.x; ctx.m[1] = v.y; ctx.m[2] = v.z; ctx.m[3] = v.w;
    } else if constexpr (VEC == 8) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(mp));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(mp) + 1);
      ctx.m[0] = v0.x; ctx.m[1] = v0.y; ctx.m[2] = v0.z; ctx.m[3] = v0.w;
      ctx.m[4] = v1.x; ctx.m[5] = v1.y; ctx.m[6] = v1.z; ctx.m[7] = v1.w;
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) ctx.m[e] = __ldg(mp + e);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inference
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, bool HAS_MASK, bool CODES>
__global__ void __launch_bounds__(QTHREADS)
tile_quantize_kernel(const T* __restrict__ x, T* __restrict__ y, QGeom g,
                     const float* __restrict__ bit_map, const float2* __restrict__ qtable,
                     const float* __restrict__ mask, int8_t* __restrict__ codes, bool inplace) {
  const long long gv = (long long)blockIdx.x * QTHREADS + threadIdx.x;
  if (gv >= g.nvec_total) return;
  const int b = (int)(gv / g.nvec);
  const int v = (int)(gv - (long long)b * g.nvec);
  const int pix = v * VEC;
  PixCtx<VEC> ctx;
  make_ctx<VEC, HAS_MASK>(g, b, pix, bit_map, mask, ctx);
  float qmin0, qmax0;
  bit_limits(ctx.bidx[0], qmin0, qmax0);

  const int c_begin = blockIdx.y * QCHUNK;
  const int c_end = min(c_begin + QCHUNK, g.C);
  const long long base = ((long long)b * g.C) * g.HW + pix;
  const float2* trow0 = qtable + (long long)ctx.bidx[0] * g.C;

  for (int c0 = c_begin; c0 < c_end; c0 += QUNROLL) {
    float xv[QUNROLL][VEC];
    float2 sz[QUNROLL];
#pragma unroll
    for (int u = 0; u < QUNROLL; ++u) {
      const int c = c0 + u;
      if (c < c_end) {
        load_elems<T, VEC>(x + base + (long long)c * g.HW, xv[u], inplace);
        sz[u] = __ldg(trow0 + c);
      }
    }
#pragma unroll
    for (int u = 0; u < QUNROLL; ++u) {
      const int c = c0 + u;
      if (c >= c_end) break;
      float out[VEC];
      int8_t cd[VEC];
      if (ctx.uniform) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float q = quant_code(xv[u][e], sz[u].x, sz[u].y, qmin0, qmax0);
          float d = dequant(q, sz[u].x, sz[u].y);
          if (HAS_MASK) d = __fmul_rn(d, ctx.m[e]);
          out[e] = d;
          if (CODES) cd[e] = (int8_t)(int)q;
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float qmin, qmax;
          bit_limits(ctx.bidx[e], qmin, qmax);
          const float2 p = __ldg(qtable + (long long)ctx.bidx[e] * g.C + c);
          const float q = quant_code(xv[u][e], p.x, p.y, qmin, qmax);
          float d = dequant(q, p.x, p.y);
          if (HAS_MASK) d = __fmul_rn(d, ctx.m[e]);
          out[e] = d;
          if (CODES) cd[e] = (int8_t)(int)q;
        }
      }
      store_elems<T, VEC>(y + base + (long long)c * g.HW, out);
      if (CODES) {
        int8_t* cp = codes + base + (long long)c * g.HW;
#pragma unroll
        for (int e = 0; e < VEC; ++e) cp[e] = cd[e];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inference, vector path.  Requirements (checked by the dispatcher): H*W % VEC == 0, W % 4 == 0,
// W % Wt == 0 and (W / Wt) % 4 == 0, so every aligned group of 4 pixels lies in one row and one
// tile.  A thread owns one 16-byte pixel vector = VEC/4 such segments and walks QV_CHUNK
// channels with QV_UNROLL independent LDG.128 in flight; the CTA's {scale, zero_point} rows for
// its channel chunk sit in shared memory (row stride 17 float2: lanes with different bit-widths
// hit different banks).
// ---------------------------------------------------------------------------------------------
// per-channel ranges given directly (no table kernel): packed = [min_c..., -max_c...] or running stats
struct QRanges {
  const float* packed;
  const float* rmin;
  const float* rmax;
};

// {scale, zero_point} of (channel c, bits bi+2)  (quantization.py:41-66)
__device__ __forceinline__ float2 qparams_from_ranges(const QRanges& rg, int C, int c, int bi) {
  const float mn = rg.packed ? __ldg(rg.packed + c) : __ldg(rg.rmin + c);
  const float mx = rg.packed ? -__ldg(rg.packed + C + c) : __ldg(rg.rmax + c);
  float qmin, qmax;
  bit_limits(bi, qmin, qmax);
  const float rng = fmaxf(__fsub_rn(mx, mn), 1e-8f);
  const float scale = __fdiv_rn(rng, __fsub_rn(qmax, qmin));
  const float zp = __fsub_rn(qmin, __fdiv_rn(mn, scale));
  return make_float2(scale, fminf(fmaxf(zp, qmin), qmax));
}

template <typename T, int VEC, bool HAS_MASK, bool CODES, int CH>
__global__ void __launch_bounds__(QV_THREADS, K3_MINB)
tile_quantize_vec_kernel(const T* __restrict__ x, T* __restrict__ y, QGeom g,
                         const float* __restrict__ bit_map, const float2* __restrict__ qtable,
                         const float* __restrict__ mask, int8_t* __restrict__ codes, bool inplace,
                         QRanges rg) {
  constexpr int NSEG = VEC / 4;
  constexpr int UN = QV_UNROLL < CH ? QV_UNROLL : CH;
  __shared__ float4 tab[7 * (CH + 1)];               // {scale, zero_point, RN(1/scale), -}
  const int c_begin = blockIdx.y * CH;
  const int nch = min(CH, g.C - c_begin);
  for (int i = threadIdx.x; i < 7 * CH; i += QV_THREADS) {
    const int bi = i / CH, cl = i - bi * CH;
    if (cl < nch) {
      float2 p;
      if (qtable) p = __ldg(qtable + (long long)bi * g.C + c_begin + cl);
      else p = qparams_from_ranges(rg, g.C, c_begin + cl, bi);
      tab[bi * (CH + 1) + cl] = make_float4(p.x, p.y, __frcp_rn(p.x), 0.f);
    }
  }
  __syncthreads();
  const long long gv = (long long)blockIdx.x * QV_THREADS + threadIdx.x;
  if (gv >= g.nvec_total) return;
  const int b = (int)(gv / g.nvec);
  const int v = (int)(gv - (long long)b * g.nvec);
  const int pix = v * VEC;
  const int h0 = pix / g.W, w0 = pix - h0 * g.W;
  int trow[NSEG];
  float qmin[NSEG], qmax[NSEG], m[VEC];
#pragma unroll
  for (int s = 0; s < NSEG; ++s) {
    int hs = h0, ws = w0 + 4 * s;
    if (ws >= g.W) { ws -= g.W; hs += 1; }
    const int ty = nearest_src(hs, g.sy, g.Ht), tx = nearest_src(ws, g.sx, g.Wt);
    float bf = rintf(__ldg(bit_map + ((long long)b * g.Ht + ty) * g.Wt + tx));
    bf = fminf(fmaxf(bf, 2.f), 8.f);
    const int bidx = (int)bf - 2;
    trow[s] = bidx * (CH + 1);
    bit_limits(bidx, qmin[s], qmax[s]);
    if (HAS_MASK) {
      const float4 mv = __ldg(reinterpret_cast<const float4*>(mask + (long long)b * g.HW + pix) + s);
      m[4 * s + 0] = mv.x; m[4 * s + 1] = mv.y; m[4 * s + 2] = mv.z; m[4 * s + 3] = mv.w;
    }
  }
  const long long base = ((long long)b * g.C + c_begin) * g.HW + pix;
  // running byte pointers: one 64-bit add per channel instead of re-deriving base + c * HW
  const char* xb = reinterpret_cast<const char*>(x + base);
  char* yb = reinterpret_cast<char*>(y + base);
  const long long sb = (long long)g.HW * (long long)sizeof(T);
  (void)inplace;                       // loads are coherent (ld.global, no L1 allocation): y may alias x
  // one channel of the thread's pixel vector: quantize / dequantize / mask, store, optional codes
  auto emit = [&](const uint4& rawv, int c, char* dst) {
    float xv[VEC], out[VEC];
    Elem<T>::unpack(rawv, xv);
    uint32_t cpack[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
      const float4 p = tab[trow[s] + c];
      uint32_t cp = 0;
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const int i = 4 * s + e;
        const float2 q = quant_code_fast2(make_float2(xv[i], xv[i + 1]), p.x, p.y, p.z, qmin[s], qmax[s]);
        float2 d = dequant2(q, p.x, p.y);
        if (HAS_MASK) d = fmul2(d, make_float2(m[i], m[i + 1]));
        out[i] = d.x;
        out[i + 1] = d.y;
        if (CODES) cp |= (((uint32_t)(int)q.x & 0xffu) << (8 * e)) | (((uint32_t)(int)q.y & 0xffu) << (8 * e + 8));
      }
      cpack[s] = cp;
    }
    stg_stream(dst, Elem<T>::pack(out));
    if (CODES) {
      uint32_t* cdst = reinterpret_cast<uint32_t*>(codes + base + (long long)c * g.HW);
#pragma unroll
      for (int s = 0; s < NSEG; ++s) cdst[s] = cpack[s];
    }
  };
#if MCAQ_L2_HINTS
  const unsigned long long l2pol = l2_policy_evict_first();
#endif
  if (nch == CH) {
    // full chunk (the common case): no per-channel predicates, the whole channel walk unrolled so
    // the table offsets are immediates
#pragma unroll
    for (int c0 = 0; c0 < CH; c0 += UN) {
      uint4 raw[UN];
#pragma unroll
#if MCAQ_L2_HINTS
      for (int u = 0; u < UN; ++u) { raw[u] = ldg_noalloc_hint(xb, l2pol); xb += sb; }
#else
      for (int u = 0; u < UN; ++u) { raw[u] = ldg_noalloc(xb); xb += sb; }
#endif
#pragma unroll
      for (int u = 0; u < UN; ++u) { emit(raw[u], c0 + u, yb); yb += sb; }
    }
    return;
  }
#pragma unroll 1
  for (int c0 = 0; c0 < nch; c0 += UN) {
    uint4 raw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u)
      if (c0 + u < nch) raw[u] = ldg_noalloc(xb + u * sb);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (c0 + u >= nch) break;
      emit(raw[u], c0 + u, yb + u * sb);
    }
    xb += UN * sb;
    yb += UN * sb;
  }
}

// ---------------------------------------------------------------------------------------------
// training forward: pre = (1-f)*Q_lo + f*Q_hi ; y = pre*m   (quantization.py:709-727, 742-744)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void frac_quant(float xv, const float2& plo, const float2& phi, const FracCtx& fc,
                                           float& qlo, float& qhi) {
  float mn, mx;
  bit_limits(fc.lo_idx, mn, mx);
  qlo = dequant(quant_code(xv, plo.x, plo.y, mn, mx), plo.x, plo.y);
  bit_limits(fc.hi_idx, mn, mx);
  qhi = dequant(quant_code(xv, phi.x, phi.y, mn, mx), phi.x, phi.y);
}

template <typename T, int VEC, bool HAS_MASK>
__global__ void __launch_bounds__(QTHREADS)
tile_quantize_train_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, QGeom g,
                               const float* __restrict__ bit_map, const float2* __restrict__ qtable,
                               const float* __restrict__ mask) {
  const long long gv = (long long)blockIdx.x * QTHREADS + threadIdx.x;
  if (gv >= g.nvec_total) return;
  const int b = (int)(gv / g.nvec);
  const int v = (int)(gv - (long long)b * g.nvec);
  const int pix = v * VEC;
  const int h = pix / g.W, w0 = pix - h * g.W;
  const int ty = nearest_src(h, g.sy, g.Ht);
  const float* brow = bit_map + ((long long)b * g.Ht + ty) * g.Wt;
  FracCtx fc[VEC];
  float m[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    fc[e] = frac_ctx(__ldg(brow + nearest_src(w0 + e, g.sx, g.Wt)));
    m[e] = HAS_MASK ? __ldg(mask + (long long)b * g.HW + pix + e) : 1.f;
  }
  const int c_begin = blockIdx.y * QCHUNK;
  const int c_end = min(c_begin + QCHUNK, g.C);
  const long long base = ((long long)b * g.C) * g.HW + pix;
  for (int c = c_begin; c < c_end; ++c) {
    float xv[VEC], out[VEC];
    load_elems<T, VEC>(x + base + (long long)c * g.HW, xv, false);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float2 plo = __ldg(qtable + (long long)fc[e].lo_idx * g.C + c);
      const float2 phi = __ldg(qtable + (long long)fc[e].hi_idx * g.C + c);
      float qlo, qhi;
      frac_quant(xv[e], plo, phi, fc[e], qlo, qhi);
      float pre = __fadd_rn(__fmul_rn(fc[e].omf, qlo), __fmul_rn(fc[e].f, qhi));
      out[e] = HAS_MASK ? __fmul_rn(pre, m[e]) : pre;
    }
    store_elems<T, VEC>(y + base + (long long)c * g.HW, out);
  }
}

// ---------------------------------------------------------------------------------------------
// training backward.  One CTA = (image b, tile row ty, tile col tx) x channel chunk so that the
// d(bit_map) reduction is a plain block reduction followed by ONE atomicAdd per CTA, and
// d(mask) is accumulated over the chunk in registers and added atomically per pixel.
// ---------------------------------------------------------------------------------------------
template <typename T, bool HAS_MASK>
__global__ void __launch_bounds__(256)
tile_quantize_train_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x, T* __restrict__ gx, QGeom g,
                               const float* __restrict__ bit_map, const float2* __restrict__ qtable,
                               const float* __restrict__ mask, float* __restrict__ dbit,
                               float* __restrict__ dmask, int cchunk) {
  // grid.x enumerates pixels of image b in blocks of 256; grid.y = channel chunk; grid.z = b
  const int b = blockIdx.z;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const bool ok = pix < g.HW;
  int ty = 0, tx = 0;
  FracCtx fc = frac_ctx(2.f);
  float m = 1.f;
  if (ok) {
    const int h = pix / g.W, w = pix - h * g.W;
    ty = nearest_src(h, g.sy, g.Ht);
    tx = nearest_src(w, g.sx, g.Wt);
    fc = frac_ctx(__ldg(bit_map + ((long long)b * g.Ht + ty) * g.Wt + tx));
    if (HAS_MASK) m = __ldg(mask + (long long)b * g.HW + pix);
  }
  const int c_begin = blockIdx.y * cchunk;
  const int c_end = min(c_begin + cchunk, g.C);
  const long long base = ((long long)b * g.C) * g.HW + pix;
  float acc_bit = 0.f, acc_m = 0.f;
  if (ok) {
    for (int c = c_begin; c < c_end; ++c) {
      const long long off = base + (long long)c * g.HW;
      const float gv = Elem<T>::load1(gy + off);
      const float xv = Elem<T>::load1(x + off);
      const float2 plo = __ldg(qtable + (long long)fc.lo_idx * g.C + c);
      const float2 phi = __ldg(qtable + (long long)fc.hi_idx * g.C + c);
      float qlo, qhi;
      frac_quant(xv, plo, phi, fc, qlo, qhi);
      const float gm = HAS_MASK ? __fmul_rn(gv, m) : gv;
      // dx = g*m*(1-f) + g*m*f  (autograd of the two STE branches, quantization.py:725-727)
      const float dx = __fadd_rn(__fmul_rn(gm, fc.omf), __fmul_rn(gm, fc.f));
      Elem<T>::store1(gx + off, dx);
      acc_bit = fmaf(gm, __fsub_rn(qhi, qlo), acc_bit);
      if (HAS_MASK) {
        const float pre = __fadd_rn(__fmul_rn(fc.omf, qlo), __fmul_rn(fc.f, qhi));
        acc_m = fmaf(gv, pre, acc_m);
      }
    }
    if (HAS_MASK) atomicAdd(dmask + (long long)b * g.HW + pix, acc_m);
  }
  // d(bit_map): warp reduction over contiguous runs of lanes in the same tile, then one atomicAdd
  // per run (tile_w >= 4, so at most 8 runs per warp)
  run_reduce_atomic(acc_bit, ok ? (b * g.Ht + ty) * g.Wt + tx : -1, dbit);
}

// ---------------------------------------------------------------------------------------------
// reference-compatible launcher, odd geometries: builds {scale, zp} per (bit, channel) on the fly from
// min/max (no workspace in that signature), one CTA-level table in shared memory.  The tile of a pixel follows
// F.interpolate(nearest) on (H, n_tiles_h) like the reference's PyTorch path (quantization.py:729-744) -- the
// parity bar -- which coincides with the reference kernel's h / tile_h whenever the grid divides the map.
// Geometries the vector kernel covers never come here.
// ---------------------------------------------------------------------------------------------
template <bool HAS_MASK>
__global__ void __launch_bounds__(256)
spatial_quant_compat_kernel(const float* __restrict__ x, const float* __restrict__ bit_map,
                            const float* __restrict__ mn, const float* __restrict__ mx,
                            const float* __restrict__ mask, float* __restrict__ y, QGeom g) {
  extern __shared__ float2 tab[];                      // [7][cchunk]
  const int cchunk = QCHUNK;
  const int c_begin = blockIdx.y * cchunk;
  const int c_end = min(c_begin + cchunk, g.C);
  for (int i = threadIdx.x; i < 7 * cchunk; i += blockDim.x) {
    const int bi = i / cchunk, c = c_begin + (i - bi * cchunk);
    if (c < c_end) {
      float qmin, qmax;
      bit_limits(bi, qmin, qmax);
      const float rng = fmaxf(__fsub_rn(mx[c], mn[c]), 1e-8f);
      const float scale = __fdiv_rn(rng, __fsub_rn(qmax, qmin));
      float zp = __fsub_rn(qmin, __fdiv_rn(mn[c], scale));
      tab[i] = make_float2(scale, fminf(fmaxf(zp, qmin), qmax));
    }
  }
  __syncthreads();
  const long long gp = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // pixel over B*HW
  if (gp >= (long long)g.B * g.HW) return;
  const int b = (int)(gp / g.HW);
  const int pix = (int)(gp - (long long)b * g.HW);
  const int h = pix / g.W, w = pix - h * g.W;
  const int ty = nearest_src(h, g.sy, g.Ht), tx = nearest_src(w, g.sx, g.Wt);
  float bf = rintf(__ldg(bit_map + ((long long)b * g.Ht + ty) * g.Wt + tx));
  bf = fminf(fmaxf(bf, 2.f), 8.f);
  const int bidx = (int)bf - 2;
  const float m = HAS_MASK ? __ldg(mask + (long long)b * g.HW + pix) : 1.f;
  float qmin, qmax;
  bit_limits(bidx, qmin, qmax);
  const long long base = ((long long)b * g.C) * g.HW + pix;
  for (int c = c_begin; c < c_end; ++c) {
    const float2 p = tab[bidx * cchunk + (c - c_begin)];
    const float q = quant_code(__ldg(x + base + (long long)c * g.HW), p.x, p.y, qmin, qmax);
    float d = dequant(q, p.x, p.y);
    if (HAS_MASK) d = __fmul_rn(d, m);
    y[base + (long long)c * g.HW] = d;
  }
}

static int g_k3_chunk = 0;     // tuning aid (mcaq_debug_k3_chunk): 0 = heuristic, else 8 / 16
static int k3_chunk(unsigned gx, int C) {
  if (g_k3_chunk == 8 || g_k3_chunk == 16) return g_k3_chunk;
  // measured (profiles/r02_k3_chunk.txt): 16 everywhere except where that leaves fewer than two CTAs per SM
  // of 256 threads (C5 bf16 at batch 64), where 8 is 13 % faster; 32 is never ahead
  const long long threads = (long long)gx * QV_THREADS * ((C + QV_CHUNK - 1) / QV_CHUNK);
  return threads < 2LL * 148 * 256 ? 8 : QV_CHUNK;
}

template <typename T, int VEC>
static int launch_quant(const T* x, T* y, int B, int C, int H, int W, const float* bit_map, int Ht, int Wt,
                        const float* qtable, const float* mask, int8_t* codes, cudaStream_t st,
                        QRanges rg = QRanges{nullptr, nullptr, nullptr}) {
  QGeom g = make_geom(B, C, H, W, Ht, Wt, VEC);
  const bool inplace = (const void*)x == (const void*)y;
  const float2* qt = (const float2*)qtable;
  if (VEC > 1) {
    constexpr int V = VEC > 1 ? VEC : 4;
    const unsigned gx = (unsigned)((g.nvec_total + QV_THREADS - 1) / QV_THREADS);
    if (codes) {
      dim3 grid(gx, (unsigned)((C + QV_CHUNK - 1) / QV_CHUNK));
      if (mask) tile_quantize_vec_kernel<T, V, true, true, QV_CHUNK><<<grid, QV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace, rg);
      else tile_quantize_vec_kernel<T, V, false, true, QV_CHUNK><<<grid, QV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace, rg);
    } else {
      // channels per CTA: fewer on the small scales, so that the grid still covers every SM a few times over
      const int ch = k3_chunk(gx, C);
#define MCAQ_K3_LAUNCH(CHV)                                                                                              \
  do {                                                                                                                   \
    dim3 grid(gx, (unsigned)((C + CHV - 1) / CHV));                                                                      \
    if (mask) tile_quantize_vec_kernel<T, V, true, false, CHV><<<grid, QV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace, rg); \
    else tile_quantize_vec_kernel<T, V, false, false, CHV><<<grid, QV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace, rg);   \
  } while (0)
      if (ch == 8) MCAQ_K3_LAUNCH(8);
      else MCAQ_K3_LAUNCH(16);
#undef MCAQ_K3_LAUNCH
    }
    MCAQ_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((unsigned)((g.nvec_total + QTHREADS - 1) / QTHREADS), (unsigned)((C + QCHUNK - 1) / QCHUNK));
  if (mask) {
    if (codes) tile_quantize_kernel<T, 1, true, true><<<grid, QTHREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace);
    else tile_quantize_kernel<T, 1, true, false><<<grid, QTHREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace);
  } else {
    if (codes) tile_quantize_kernel<T, 1, false, true><<<grid, QTHREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace);
    else tile_quantize_kernel<T, 1, false, false><<<grid, QTHREADS, 0, st>>>(x, y, g, bit_map, qt, mask, codes, inplace);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

template <typename T, int VEC>
static int launch_train_fwd(const T* x, T* y, int B, int C, int H, int W, const float* bit_map, int Ht, int Wt,
                            const float* qtable, const float* mask, cudaStream_t st) {
  QGeom g = make_geom(B, C, H, W, Ht, Wt, VEC);
  dim3 grid((unsigned)((g.nvec_total + QTHREADS - 1) / QTHREADS), (unsigned)((C + QCHUNK - 1) / QCHUNK));
  const float2* qt = (const float2*)qtable;
  if (mask) tile_quantize_train_fwd_kernel<T, VEC, true><<<grid, QTHREADS, 0, st>>>(x, y, g, bit_map, qt, mask);
  else tile_quantize_train_fwd_kernel<T, VEC, false><<<grid, QTHREADS, 0, st>>>(x, y, g, bit_map, qt, mask);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_train_bwd(const T* gy, const T* x, T* gx, int B, int C, int H, int W, const float* bit_map,
                            int Ht, int Wt, const float* qtable, const float* mask, float* dbit, float* dmask,
                            cudaStream_t st) {
  QGeom g = make_geom(B, C, H, W, Ht, Wt, 1);
  const int cchunk = 32;
  dim3 grid((unsigned)((g.HW + 255) / 256), (unsigned)((C + cchunk - 1) / cchunk), (unsigned)B);
  const float2* qt = (const float2*)qtable;
  if (mask) tile_quantize_train_bwd_kernel<T, true><<<grid, 256, 0, st>>>(gy, x, gx, g, bit_map, qt, mask, dbit, dmask, cchunk);
  else tile_quantize_train_bwd_kernel<T, false><<<grid, 256, 0, st>>>(gy, x, gx, g, bit_map, qt, mask, dbit, dmask, cchunk);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

static bool vec_ok(const void* a, const void* b, int HW, int W, int VEC, const void* mask) {
  return ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)mask & 15) == 0 &&
         HW % VEC == 0 && W % VEC == 0;
}

}  // namespace mcaq

using namespace mcaq;

static int check_common(const void* x, const void* y, int B, int C, int H, int W, const float* bit_map, int Ht,
                        int Wt, const float* qtable) {
  if (!x || !y || !bit_map || !qtable) return MCAQ_EINVAL;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return MCAQ_EINVAL;
  if ((long long)H * W > 0x7fffffffLL) return MCAQ_EINVAL;
  return 0;
}

extern "C" int mcaq_tile_quantize(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                  const float* bit_map, int Ht, int Wt, const float* qtable,
                                  const float* mask, int8_t* codes, void* stream) {
  int rc = check_common(x, y, B, C, H, W, bit_map, Ht, Wt, qtable);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MCAQ_F32) {
    if (seg_ok(x, y, mask, codes, H * W, W, Wt, 4))
      return launch_quant<float, 4>((const float*)x, (float*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, codes, st);
    return launch_quant<float, 1>((const float*)x, (float*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, codes, st);
  } else if (dtype == MCAQ_BF16 || dtype == MCAQ_F16) {
    const bool v8 = seg_ok(x, y, mask, codes, H * W, W, Wt, 8);
    MCAQ_DISPATCH_16(dtype, h16,
      if (v8) return launch_quant<h16, 8>((const h16*)x, (h16*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, codes, st);
      return launch_quant<h16, 1>((const h16*)x, (h16*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, codes, st));
  }
  return MCAQ_EDTYPE;
}

extern "C" int mcaq_tile_quantize_ranges(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                         const float* bit_map, int Ht, int Wt, const float* packed,
                                         const float* running_min, const float* running_max,
                                         float* qtable_ws, const float* mask, void* stream) {
  if (!packed && (!running_min || !running_max)) return MCAQ_EINVAL;
  float dummy = 0.f;
  int rc = check_common(x, y, B, C, H, W, bit_map, Ht, Wt, &dummy);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  QRanges rg{packed, running_min, running_max};
  if (dtype != MCAQ_F32 && dtype != MCAQ_BF16 && dtype != MCAQ_F16) return MCAQ_EDTYPE;
  const bool vec = dtype == MCAQ_F32 ? seg_ok(x, y, mask, nullptr, H * W, W, Wt, 4)
                                     : seg_ok(x, y, mask, nullptr, H * W, W, Wt, 8);
  if (vec) {
    if (dtype == MCAQ_F32)
      return launch_quant<float, 4>((const float*)x, (float*)y, B, C, H, W, bit_map, Ht, Wt, nullptr, mask, nullptr, st, rg);
    if (dtype == MCAQ_BF16 || dtype == MCAQ_F16)
      MCAQ_DISPATCH_16(dtype, h16,
        return launch_quant<h16, 8>((const h16*)x, (h16*)y, B, C, H, W, bit_map, Ht, Wt, nullptr, mask, nullptr, st, rg));
    return MCAQ_EDTYPE;
  }
  // odd geometry: materialise the table in the caller's workspace, then the scalar kernel
  if (!qtable_ws) return MCAQ_EINVAL;
  rc = mcaq_build_qtable(packed, running_min, running_max, C, qtable_ws, stream);
  if (rc) return rc;
  return mcaq_tile_quantize(x, y, dtype, B, C, H, W, bit_map, Ht, Wt, qtable_ws, mask, nullptr, stream);
}

// Multi-GPU, host-driven form of mcaq_tile_quantize_ranges: a one-CTA kernel waits for the `world`
// ranks' published ranges of the current step and writes their minimum to packed_ws
// (peer_exchange.cuh), then the regular K3 runs on it.  (The fused path does not need this: the
// morphology kernel's first CTA merges at its end.)  The wait is confined to that single CTA on
// purpose: CTAs of a bandwidth kernel spinning on a peer could starve the very kernel they wait for.
extern "C" int mcaq_tile_quantize_xchg(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                       const float* bit_map, int Ht, int Wt, const void* xchg_local, int world,
                                       float* qtable_ws, float* packed_ws, const float* mask, void* stream) {
  if (!xchg_local || !packed_ws || world <= 0 || world > XCHG_MAX_RANKS) return MCAQ_EINVAL;
  int rc = mcaq_xchg_merge(xchg_local, world, C, packed_ws, stream);
  if (rc) return rc;
  return mcaq_tile_quantize_ranges(x, y, dtype, B, C, H, W, bit_map, Ht, Wt, packed_ws, nullptr, nullptr,
                                   qtable_ws, mask, stream);
}

extern "C" int mcaq_tile_quantize_train_fwd(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                            const float* bit_map, int Ht, int Wt, const float* qtable,
                                            const float* mask, void* stream) {
  int rc = check_common(x, y, B, C, H, W, bit_map, Ht, Wt, qtable);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((dtype == MCAQ_F32 || dtype == MCAQ_BF16 || dtype == MCAQ_F16) &&
      train_vec_ok(x, y, nullptr, mask, nullptr, nullptr, dtype, H, W, Wt))
    return train_fwd_vec(x, y, dtype, B, C, H, W, bit_map, Ht, Wt, qtable, mask, nullptr, nullptr, st);
  if (dtype == MCAQ_F32) {
    if (vec_ok(x, y, H * W, W, 4, nullptr))
      return launch_train_fwd<float, 4>((const float*)x, (float*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, st);
    return launch_train_fwd<float, 1>((const float*)x, (float*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, st);
  } else if (dtype == MCAQ_BF16 || dtype == MCAQ_F16) {
    const bool v8 = vec_ok(x, y, H * W, W, 8, nullptr);
    MCAQ_DISPATCH_16(dtype, h16,
      if (v8) return launch_train_fwd<h16, 8>((const h16*)x, (h16*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, st);
      return launch_train_fwd<h16, 1>((const h16*)x, (h16*)y, B, C, H, W, bit_map, Ht, Wt, qtable, mask, st));
  }
  return MCAQ_EDTYPE;
}

extern "C" int mcaq_tile_quantize_train_bwd(const void* grad_y, const void* x, void* grad_x, int dtype,
                                            int B, int C, int H, int W, const float* bit_map, int Ht, int Wt,
                                            const float* qtable, const float* mask, float* dbit, float* dmask,
                                            void* stream) {
  int rc = check_common(x, grad_x, B, C, H, W, bit_map, Ht, Wt, qtable);
  if (rc) return rc;
  if (!grad_y || !dbit || (mask && !dmask)) return MCAQ_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if ((dtype == MCAQ_F32 || dtype == MCAQ_BF16 || dtype == MCAQ_F16) &&
      train_vec_ok(grad_y, grad_x, x, mask, dmask, nullptr, dtype, H, W, Wt))
    return train_bwd_vec(grad_y, x, grad_x, dtype, B, C, H, W, bit_map, Ht, Wt, qtable, mask, nullptr, nullptr,
                         dbit, dmask, st);
  if (dtype == MCAQ_F32)
    return launch_train_bwd<float>((const float*)grad_y, (const float*)x, (float*)grad_x, B, C, H, W, bit_map,
                                   Ht, Wt, qtable, mask, dbit, dmask, st);
  if (dtype == MCAQ_BF16 || dtype == MCAQ_F16)
    MCAQ_DISPATCH_16(dtype, h16,
      return launch_train_bwd<h16>((const h16*)grad_y, (const h16*)x, (h16*)grad_x, B, C, H, W, bit_map, Ht, Wt,
                                   qtable, mask, dbit, dmask, st));
  return MCAQ_EDTYPE;
}

// Level 0 of the drop-in boundary: the reference's launcher (ops/src/mcaq_kernel.cu:102-111, shared with
// engine/MCAQPlugin.cpp:15-24) with an error channel.  tile_h / tile_w must be what the reference's caller
// passes (H / n_tiles_h, W / n_tiles_w, quantization.py:641-642; anything else is MCAQ_EINVAL).  When the
// geometry is 16-byte friendly (every YOLO feature map) it IS the vector kernel of the fused path (per-channel
// ranges given directly, no table kernel); other geometries take the scalar kernel.  Tile rule and rounding
// follow the reference's PyTorch path (nearest, half-to-even), the parity bar.
extern "C" int mcaq_spatial_quantization(const float* input, const float* bit_map, const float* min_vals,
                                         const float* max_vals, const float* mask, float* output, int N, int C,
                                         int H, int W, int tile_h, int tile_w, int n_tiles_h, int n_tiles_w,
                                         void* stream) {
  if (!input || !output || !bit_map || !min_vals || !max_vals) return MCAQ_EINVAL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || n_tiles_h <= 0 || n_tiles_w <= 0 || tile_h <= 0 || tile_w <= 0)
    return MCAQ_EINVAL;
  if ((long long)H * W > 0x7fffffffLL) return MCAQ_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (tile_h != H / n_tiles_h || tile_w != W / n_tiles_w) return MCAQ_EINVAL;
  if (seg_ok(input, output, mask, nullptr, H * W, W, n_tiles_w, 4))
    return launch_quant<float, 4>(input, output, N, C, H, W, bit_map, n_tiles_h, n_tiles_w, nullptr, mask, nullptr, st,
                                  QRanges{nullptr, min_vals, max_vals});
  QGeom g = make_geom(N, C, H, W, n_tiles_h, n_tiles_w, 1);
  dim3 grid((unsigned)(((long long)N * H * W + 255) / 256), (unsigned)((C + QCHUNK - 1) / QCHUNK));
  const size_t smem = 7 * QCHUNK * sizeof(float2);
  if (mask) spatial_quant_compat_kernel<true><<<grid, 256, smem, st>>>(input, bit_map, min_vals, max_vals, mask, output, g);
  else spatial_quant_compat_kernel<false><<<grid, 256, smem, st>>>(input, bit_map, min_vals, max_vals, mask, output, g);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

static thread_local int g_level0_status = 0;
// status of this thread's last launch_spatial_quantization call (0 ok / cudaError_t / negative MCAQ_E*)
extern "C" int mcaq_level0_status() { return g_level0_status; }

// the reference's exact symbol and argument list: returns void there, so the outcome is kept per thread
extern "C" void launch_spatial_quantization(const float* input, const float* bit_map, const float* min_vals,
                                            const float* max_vals, const float* mask, float* output, int N,
                                            int C, int H, int W, int tile_h, int tile_w, int n_tiles_h,
                                            int n_tiles_w, void* stream) {
  g_level0_status = mcaq_spatial_quantization(input, bit_map, min_vals, max_vals, mask, output, N, C, H, W, tile_h,
                                              tile_w, n_tiles_h, n_tiles_w, stream);
}

extern "C" void mcaq_debug_k3_chunk(int ch) { mcaq::g_k3_chunk = ch; }
