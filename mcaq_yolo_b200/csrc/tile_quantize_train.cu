// K3, training forms on the vector path: the fractional-bit compose (quantization.py:699-727,
// 742-744), its straight-through backward (quantization.py:69-118 + autograd of the compose), and
// the feature-level distillation term (train.py:599-610: F.mse_loss(features_q, teacher)) folded
// into both so that y is never re-read for the loss.
//
// Same geometry as the inference vector kernel (tile_quantize.cu): a thread owns one 16-byte
// pixel vector (1 or 2 four-pixel segments, each inside one tile) and walks a 16-channel chunk
// with several 128-bit loads in flight; the CTA's {scale, zero_point, RN(1/scale)} rows of the
// chunk sit in shared memory; x/scale is Markstein's correction (== div.rn, test_gpu_division).
//
// HBM traffic per element of size s: forward 2*s (+4 for the fp32 teacher), backward 3*s
// (read g, read x, write dx; +4 for the teacher).  d(bit_map) is reduced over runs of lanes in the
// same tile and added with one atomic per run; d(mask) with one 16-byte vector reduction per
// 4-pixel segment and channel chunk.
#include <stdlib.h>

#include "tile_quantize.cuh"

namespace mcaq {

constexpr int TV_THREADS = 256;
constexpr int TV_CHUNK = 16;
constexpr int TV_ROW = TV_CHUNK + 1;

static int g_train_force_scalar = 0;

template <int NSEG>
struct TrainCtx {
  int lo_row[NSEG], hi_row[NSEG];
  float f[NSEG], omf[NSEG];
  float mn_lo[NSEG], mx_lo[NSEG], mn_hi[NSEG], mx_hi[NSEG];
  int tile[NSEG];                 // flat (b, ty, tx) index into d(bit_map)
};

__device__ __forceinline__ void fill_table(float4* tab, const float2* __restrict__ qtable, int C, int c_begin,
                                           int nch) {
  for (int i = threadIdx.x; i < 7 * TV_CHUNK; i += TV_THREADS) {
    const int bi = i / TV_CHUNK, cl = i - bi * TV_CHUNK;
    if (cl < nch) {
      const float2 p = __ldg(qtable + (long long)bi * C + c_begin + cl);
      tab[bi * TV_ROW + cl] = make_float4(p.x, p.y, __frcp_rn(p.x), 0.f);
    }
  }
}

template <int VEC, bool HAS_MASK>
__device__ __forceinline__ void make_train_ctx(const QGeom& g, int b, int pix, const float* __restrict__ bit_map,
                                               const float* __restrict__ mask, TrainCtx<VEC / 4>& ctx,
                                               float* m) {
  constexpr int NSEG = VEC / 4;
  const int h0 = pix / g.W, w0 = pix - h0 * g.W;
#pragma unroll
  for (int s = 0; s < NSEG; ++s) {
    int hs = h0, ws = w0 + 4 * s;
    if (ws >= g.W) { ws -= g.W; hs += 1; }
    const int ty = nearest_src(hs, g.sy, g.Ht), tx = nearest_src(ws, g.sx, g.Wt);
    const int t = (b * g.Ht + ty) * g.Wt + tx;
    const FracCtx fc = frac_ctx(__ldg(bit_map + t));
    ctx.tile[s] = t;
    ctx.lo_row[s] = fc.lo_idx * TV_ROW;
    ctx.hi_row[s] = fc.hi_idx * TV_ROW;
    ctx.f[s] = fc.f;
    ctx.omf[s] = fc.omf;
    bit_limits(fc.lo_idx, ctx.mn_lo[s], ctx.mx_lo[s]);
    bit_limits(fc.hi_idx, ctx.mn_hi[s], ctx.mx_hi[s]);
    if (HAS_MASK) {
      const float4 mv = __ldg(reinterpret_cast<const float4*>(mask + (long long)b * g.HW + pix) + s);
      m[4 * s + 0] = mv.x; m[4 * s + 1] = mv.y; m[4 * s + 2] = mv.z; m[4 * s + 3] = mv.w;
    }
  }
}

// Q_lo(x), Q_hi(x) of one element: de-quantised values at floor(b) and floor(b)+1 bits
__device__ __forceinline__ void frac_pair(float xv, const float4& plo, const float4& phi, float mn_lo, float mx_lo,
                                          float mn_hi, float mx_hi, float& qlo, float& qhi) {
  qlo = dequant(quant_code_fast(xv, plo.x, plo.y, plo.z, mn_lo, mx_lo), plo.x, plo.y);
  qhi = dequant(quant_code_fast(xv, phi.x, phi.y, phi.z, mn_hi, mx_hi), phi.x, phi.y);
}

// the same for two elements, packed fp32 pairs (identical roundings)
__device__ __forceinline__ void frac_pair2(float2 xv, const float4& plo, const float4& phi, float mn_lo, float mx_lo,
                                           float mn_hi, float mx_hi, float2& qlo, float2& qhi) {
  qlo = dequant2(quant_code_fast2(xv, plo.x, plo.y, plo.z, mn_lo, mx_lo), plo.x, plo.y);
  qhi = dequant2(quant_code_fast2(xv, phi.x, phi.y, phi.z, mn_hi, mx_hi), phi.x, phi.y);
}

__device__ __forceinline__ void unpack4(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
  f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}

// 16-byte vector reduction into global memory (sm_90+): one L2 operation for four floats
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// forward:  pre = (1-f) Q_lo(x) + f Q_hi(x),  y = pre * m,  [kd_sum += sum (y - teacher)^2]
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, bool HAS_MASK, bool KD>
__global__ void __launch_bounds__(TV_THREADS, 2)
train_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, QGeom g, const float* __restrict__ bit_map,
                     const float2* __restrict__ qtable, const float* __restrict__ mask,
                     const float* __restrict__ teacher, double* __restrict__ kd_sum, int cchunk) {
  constexpr int NSEG = VEC / 4;
  constexpr int UNROLL = KD ? (VEC == 8 ? 2 : 4) : (VEC == 8 ? 4 : 8);
  __shared__ float4 tab[7 * TV_ROW];
  __shared__ float red[TV_THREADS / 32];
  const int c_begin = blockIdx.y * cchunk;
  const int nch = min(cchunk, g.C - c_begin);
  fill_table(tab, qtable, g.C, c_begin, nch);
  __syncthreads();
  const long long gv = (long long)blockIdx.x * TV_THREADS + threadIdx.x;
  float kd_acc = 0.f;
  if (gv < g.nvec_total) {
    const int b = (int)(gv / g.nvec);
    const int pix = (int)(gv - (long long)b * g.nvec) * VEC;
    TrainCtx<NSEG> ctx;
    float m[VEC];
    make_train_ctx<VEC, HAS_MASK>(g, b, pix, bit_map, mask, ctx, m);
    const long long base = ((long long)b * g.C + c_begin) * g.HW + pix;
    const char* xb = reinterpret_cast<const char*>(x + base);
    char* yb = reinterpret_cast<char*>(y + base);
    const char* tb = reinterpret_cast<const char*>(teacher + (KD ? base : 0));
    const long long sb = (long long)g.HW * (long long)sizeof(T);
    const long long st = (long long)g.HW * 4;
    auto emit = [&](const uint4& rawv, const uint4* trawv, int c, char* dst) {
      float xv[VEC], out[VEC];
      Elem<T>::unpack(rawv, xv);
#pragma unroll
      for (int s = 0; s < NSEG; ++s) {
        const float4 plo = tab[ctx.lo_row[s] + c];
        const float4 phi = tab[ctx.hi_row[s] + c];
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const int i = 4 * s + e;
          float2 qlo, qhi;
          frac_pair2(make_float2(xv[i], xv[i + 1]), plo, phi, ctx.mn_lo[s], ctx.mx_lo[s], ctx.mn_hi[s], ctx.mx_hi[s],
                     qlo, qhi);
          float2 pre = fadd2_sep(fmul2(splat2(ctx.omf[s]), qlo), fmul2(splat2(ctx.f[s]), qhi));
          if (HAS_MASK) pre = fmul2(pre, make_float2(m[i], m[i + 1]));
          out[i] = pre.x;
          out[i + 1] = pre.y;
        }
      }
      const uint4 packed = Elem<T>::pack(out);
      stg_stream(dst, packed);
      if (KD) {
        float yr[VEC];
        Elem<T>::unpack(packed, yr);                 // the value as stored (bf16: rounded)
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
          float tv[4];
          unpack4(trawv[s], tv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d = __fsub_rn(yr[4 * s + e], tv[e]);
            kd_acc = fmaf(d, d, kd_acc);
          }
        }
      }
    };
#pragma unroll 1
    for (int c0 = 0; c0 < nch; c0 += UNROLL) {
      uint4 raw[UNROLL];
      uint4 traw[UNROLL][NSEG];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (c0 + u < nch) {
          raw[u] = ldg_stream(xb + u * sb);
          if (KD) {
#pragma unroll
            for (int s = 0; s < NSEG; ++s) traw[u][s] = ldg_stream(tb + u * st + 16 * s);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (c0 + u >= nch) break;
        emit(raw[u], traw[u], c0 + u, yb + u * sb);
      }
      xb += UNROLL * sb;
      yb += UNROLL * sb;
      tb += UNROLL * st;
    }
  }
  if (KD) {
    // CTA sum in fp32 (<= 256 * 16 * 8 terms), then one fp64 atomic per CTA
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) kd_acc += __shfl_xor_sync(0xffffffffu, kd_acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kd_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < TV_THREADS / 32; ++w) s += (double)red[w];
      atomicAdd(kd_sum, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  g_t = g + kd_coef * (y - teacher)   (kd_coef = dL/d(mse) * 2 / numel, device scalar)
//   dx = g_t*m*(1-f) + g_t*m*f ;  dbit[tile] += sum g_t*m*(Q_hi - Q_lo) ;  dmask[pix] += sum_c g_t*pre
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, bool HAS_MASK, bool KD>
__global__ void __launch_bounds__(TV_THREADS, 2)
train_bwd_vec_kernel(const T* __restrict__ gy, const T* __restrict__ x, T* __restrict__ gx, QGeom g,
                     const float* __restrict__ bit_map, const float2* __restrict__ qtable,
                     const float* __restrict__ mask, const float* __restrict__ teacher,
                     const float* __restrict__ kd_coef, float* __restrict__ dbit, float* __restrict__ dmask,
                     int cchunk) {
  constexpr int NSEG = VEC / 4;
  constexpr int UNROLL = KD ? 2 : 4;
  __shared__ float4 tab[7 * TV_ROW];
  const int c_begin = blockIdx.y * cchunk;
  const int nch = min(cchunk, g.C - c_begin);
  fill_table(tab, qtable, g.C, c_begin, nch);
  __syncthreads();
  const long long gv = (long long)blockIdx.x * TV_THREADS + threadIdx.x;
  const bool ok = gv < g.nvec_total;
  float acc_bit[NSEG];
  int tile[NSEG];
#pragma unroll
  for (int s = 0; s < NSEG; ++s) { acc_bit[s] = 0.f; tile[s] = -1; }
  if (ok) {
    const int b = (int)(gv / g.nvec);
    const int pix = (int)(gv - (long long)b * g.nvec) * VEC;
    TrainCtx<NSEG> ctx;
    float m[VEC], acc_m[VEC];
    make_train_ctx<VEC, HAS_MASK>(g, b, pix, bit_map, mask, ctx, m);
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc_m[e] = 0.f;
#pragma unroll
    for (int s = 0; s < NSEG; ++s) tile[s] = ctx.tile[s];
    const float coef = KD ? __ldg(kd_coef) : 0.f;
    const long long base = ((long long)b * g.C + c_begin) * g.HW + pix;
    const char* gb = reinterpret_cast<const char*>(gy + base);
    const char* xb = reinterpret_cast<const char*>(x + base);
    char* ob = reinterpret_cast<char*>(gx + base);
    const char* tb = reinterpret_cast<const char*>(teacher + (KD ? base : 0));
    const long long sb = (long long)g.HW * (long long)sizeof(T);
    const long long st = (long long)g.HW * 4;
    auto emit = [&](const uint4& graw, const uint4& xraw, const uint4* trawv, int c, char* dst) {
      float gvv[VEC], xv[VEC], out[VEC];
      Elem<T>::unpack(graw, gvv);
      Elem<T>::unpack(xraw, xv);
#pragma unroll
      for (int s = 0; s < NSEG; ++s) {
        const float4 plo = tab[ctx.lo_row[s] + c];
        const float4 phi = tab[ctx.hi_row[s] + c];
        float tv[4];
        if (KD) unpack4(trawv[s], tv);
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const int i = 4 * s + e;
          float2 qlo, qhi;
          frac_pair2(make_float2(xv[i], xv[i + 1]), plo, phi, ctx.mn_lo[s], ctx.mx_lo[s], ctx.mn_hi[s], ctx.mx_hi[s],
                     qlo, qhi);
          const float2 omf2 = splat2(ctx.omf[s]), f2 = splat2(ctx.f[s]);
          const float2 pre = fadd2_sep(fmul2(omf2, qlo), fmul2(f2, qhi));
          const float2 m2 = make_float2(m[i], m[i + 1]);
          float2 gt = make_float2(gvv[i], gvv[i + 1]);
          if (KD) {
            const float2 yv = HAS_MASK ? fmul2(pre, m2) : pre;
            const float2 yr = make_float2(Elem<T>::round1(yv.x), Elem<T>::round1(yv.y));
            gt = fadd2_sep(gt, fmul2(splat2(coef), fadd2_sep(yr, make_float2(-tv[e], -tv[e + 1]))));   // yr may be a product
          }
          const float2 gm = HAS_MASK ? fmul2(gt, m2) : gt;
          // dx = g*m*(1-f) + g*m*f  (autograd of the two STE branches, quantization.py:725-727)
          const float2 dxv = fadd2_sep(fmul2(gm, omf2), fmul2(gm, f2));
          out[i] = dxv.x;
          out[i + 1] = dxv.y;
          const float2 dq = fadd2_sep(qhi, make_float2(-qlo.x, -qlo.y));      // q_hi, q_lo are packed products
          acc_bit[s] = fmaf(gm.x, dq.x, acc_bit[s]);
          acc_bit[s] = fmaf(gm.y, dq.y, acc_bit[s]);
          if (HAS_MASK) {
            const float2 am = ffma2(gt, pre, make_float2(acc_m[i], acc_m[i + 1]));
            acc_m[i] = am.x;
            acc_m[i + 1] = am.y;
          }
        }
      }
      stg_stream(dst, Elem<T>::pack(out));
    };
#pragma unroll 1
    for (int c0 = 0; c0 < nch; c0 += UNROLL) {
      uint4 graw[UNROLL], xraw[UNROLL];
      uint4 traw[UNROLL][NSEG];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (c0 + u < nch) {
          graw[u] = ldg_stream(gb + u * sb);
          xraw[u] = ldg_stream(xb + u * sb);
          if (KD) {
#pragma unroll
            for (int s = 0; s < NSEG; ++s) traw[u][s] = ldg_stream(tb + u * st + 16 * s);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (c0 + u >= nch) break;
        emit(graw[u], xraw[u], traw[u], c0 + u, ob + u * sb);
      }
      gb += UNROLL * sb;
      xb += UNROLL * sb;
      ob += UNROLL * sb;
      tb += UNROLL * st;
    }
    if (HAS_MASK) {
      float* dm = dmask + (long long)b * g.HW + pix;
#pragma unroll
      for (int s = 0; s < NSEG; ++s)
        red_add_v4(dm + 4 * s, acc_m[4 * s], acc_m[4 * s + 1], acc_m[4 * s + 2], acc_m[4 * s + 3]);
    }
  }
  // d(bit_map): the two segments of a bf16 vector usually share a tile; then runs of lanes
  if (NSEG == 2 && tile[0] == tile[NSEG - 1] && tile[0] >= 0) {
    acc_bit[0] += acc_bit[NSEG - 1];
    tile[NSEG - 1] = -1;
  }
#pragma unroll
  for (int s = 0; s < NSEG; ++s) {
    // skip the pass when no lane of the warp has a live second segment
    if (s == 0 || __any_sync(0xffffffffu, tile[s] >= 0)) run_reduce_atomic(acc_bit[s], tile[s], dbit);
  }
}

// channels per CTA: 16, or 8 when 16 would leave the GPU under two CTAs per SM (training batches are
// small: 16 images per GPU in BASELINE configs[3]); MCAQ_TRAIN_CHUNK overrides (tuning)
static int pick_chunk(const QGeom& g) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MCAQ_TRAIN_CHUNK");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 4 || forced == 8 || forced == 16) return forced;
  const long long gx = (g.nvec_total + TV_THREADS - 1) / TV_THREADS;
  return gx * ((g.C + 15) / 16) < 2 * 148 ? 8 : 16;
}

template <typename T, int VEC>
static int launch_fwd_vec(const T* x, T* y, const QGeom& g, const float* bit_map, const float* qtable,
                          const float* mask, const float* teacher, double* kd_sum, cudaStream_t st) {
  const int ck = pick_chunk(g);
  dim3 grid((unsigned)((g.nvec_total + TV_THREADS - 1) / TV_THREADS), (unsigned)((g.C + ck - 1) / ck));
  const float2* qt = (const float2*)qtable;
  if (teacher) {
    if (mask) train_fwd_vec_kernel<T, VEC, true, true><<<grid, TV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, teacher, kd_sum, ck);
    else train_fwd_vec_kernel<T, VEC, false, true><<<grid, TV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, teacher, kd_sum, ck);
  } else {
    if (mask) train_fwd_vec_kernel<T, VEC, true, false><<<grid, TV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, teacher, kd_sum, ck);
    else train_fwd_vec_kernel<T, VEC, false, false><<<grid, TV_THREADS, 0, st>>>(x, y, g, bit_map, qt, mask, teacher, kd_sum, ck);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

template <typename T, int VEC>
static int launch_bwd_vec(const T* gy, const T* x, T* gx, const QGeom& g, const float* bit_map, const float* qtable,
                          const float* mask, const float* teacher, const float* kd_coef, float* dbit,
                          float* dmask, cudaStream_t st) {
  const int ck = pick_chunk(g);
  dim3 grid((unsigned)((g.nvec_total + TV_THREADS - 1) / TV_THREADS), (unsigned)((g.C + ck - 1) / ck));
  const float2* qt = (const float2*)qtable;
  if (teacher) {
    if (mask) train_bwd_vec_kernel<T, VEC, true, true><<<grid, TV_THREADS, 0, st>>>(gy, x, gx, g, bit_map, qt, mask, teacher, kd_coef, dbit, dmask, ck);
    else train_bwd_vec_kernel<T, VEC, false, true><<<grid, TV_THREADS, 0, st>>>(gy, x, gx, g, bit_map, qt, mask, teacher, kd_coef, dbit, dmask, ck);
  } else {
    if (mask) train_bwd_vec_kernel<T, VEC, true, false><<<grid, TV_THREADS, 0, st>>>(gy, x, gx, g, bit_map, qt, mask, teacher, kd_coef, dbit, dmask, ck);
    else train_bwd_vec_kernel<T, VEC, false, false><<<grid, TV_THREADS, 0, st>>>(gy, x, gx, g, bit_map, qt, mask, teacher, kd_coef, dbit, dmask, ck);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// geometry + alignment test of the training vector path
bool train_vec_ok(const void* a, const void* b, const void* c, const void* mask, const void* dmask,
                  const void* teacher, int dtype, int H, int W, int Wt) {
  if (g_train_force_scalar) return false;
  const int VEC = dtype == MCAQ_F32 ? 4 : 8;
  return seg_ok(a, b, mask, nullptr, H * W, W, Wt, VEC) && aligned16(c) && aligned16(dmask) && aligned16(teacher);
}

int train_fwd_vec(const void* x, void* y, int dtype, int B, int C, int H, int W, const float* bit_map, int Ht,
                  int Wt, const float* qtable, const float* mask, const float* teacher, double* kd_sum,
                  cudaStream_t st) {
  if (dtype == MCAQ_F32) {
    const QGeom g = make_geom(B, C, H, W, Ht, Wt, 4);
    return launch_fwd_vec<float, 4>((const float*)x, (float*)y, g, bit_map, qtable, mask, teacher, kd_sum, st);
  }
  const QGeom g = make_geom(B, C, H, W, Ht, Wt, 8);
  MCAQ_DISPATCH_16(dtype, h16,
    return launch_fwd_vec<h16, 8>((const h16*)x, (h16*)y, g, bit_map, qtable, mask, teacher, kd_sum, st));
}

int train_bwd_vec(const void* gy, const void* x, void* gx, int dtype, int B, int C, int H, int W,
                  const float* bit_map, int Ht, int Wt, const float* qtable, const float* mask,
                  const float* teacher, const float* kd_coef, float* dbit, float* dmask, cudaStream_t st) {
  if (dtype == MCAQ_F32) {
    const QGeom g = make_geom(B, C, H, W, Ht, Wt, 4);
    return launch_bwd_vec<float, 4>((const float*)gy, (const float*)x, (float*)gx, g, bit_map, qtable, mask,
                                    teacher, kd_coef, dbit, dmask, st);
  }
  const QGeom g = make_geom(B, C, H, W, Ht, Wt, 8);
  MCAQ_DISPATCH_16(dtype, h16,
    return launch_bwd_vec<h16, 8>((const h16*)gy, (const h16*)x, (h16*)gx, g, bit_map, qtable, mask, teacher, kd_coef,
                                  dbit, dmask, st));
}

}  // namespace mcaq

using namespace mcaq;

extern "C" void mcaq_debug_train_scalar(int on) { g_train_force_scalar = on; }

static int kd_args_ok(const void* x, const void* y, int dtype, int B, int C, int H, int W, const float* bit_map,
                      int Ht, int Wt, const float* qtable, const float* teacher) {
  if (!x || !y || !bit_map || !qtable || !teacher) return MCAQ_EINVAL;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return MCAQ_EINVAL;
  if ((long long)H * W > 0x7fffffffLL) return MCAQ_EINVAL;
  if (dtype != MCAQ_F32 && dtype != MCAQ_BF16 && dtype != MCAQ_F16) return MCAQ_EDTYPE;
  return 0;
}

extern "C" int mcaq_tile_quantize_train_fwd_kd(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                               const float* bit_map, int Ht, int Wt, const float* qtable,
                                               const float* mask, const float* teacher, double* kd_sum,
                                               void* stream) {
  int rc = kd_args_ok(x, y, dtype, B, C, H, W, bit_map, Ht, Wt, qtable, teacher);
  if (rc) return rc;
  if (!kd_sum) return MCAQ_EINVAL;
  if (!train_vec_ok(x, y, nullptr, mask, nullptr, teacher, dtype, H, W, Wt)) return MCAQ_EGEOM;
  return train_fwd_vec(x, y, dtype, B, C, H, W, bit_map, Ht, Wt, qtable, mask, teacher, kd_sum,
                       (cudaStream_t)stream);
}

extern "C" int mcaq_tile_quantize_train_bwd_kd(const void* grad_y, const void* x, void* grad_x, int dtype, int B,
                                               int C, int H, int W, const float* bit_map, int Ht, int Wt,
                                               const float* qtable, const float* mask, const float* teacher,
                                               const float* kd_coef, float* dbit, float* dmask, void* stream) {
  int rc = kd_args_ok(x, grad_x, dtype, B, C, H, W, bit_map, Ht, Wt, qtable, teacher);
  if (rc) return rc;
  if (!grad_y || !kd_coef || !dbit || (mask && !dmask)) return MCAQ_EINVAL;
  if (!train_vec_ok(grad_y, grad_x, x, mask, dmask, teacher, dtype, H, W, Wt)) return MCAQ_EGEOM;
  return train_bwd_vec(grad_y, x, grad_x, dtype, B, C, H, W, bit_map, Ht, Wt, qtable, mask, teacher, kd_coef,
                       dbit, dmask, (cudaStream_t)stream);
}
