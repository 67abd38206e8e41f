// channels_last (NHWC) variants of the two HBM sweeps.  The memory of a channels_last tensor is
// (B, H, W, C) with the C channels of a pixel contiguous; the results are defined to be exactly
// those of the NCHW kernels on the same logical tensor (same summation order, same codes), so the
// oracle and the parity tests are layout independent.
//
// K1-NHWC  reads x once with fully coalesced 16-byte loads (thread <-> fixed group of VEC channels,
//          so per-channel min / max is a running 2-wide HMNMX2 / FMNMX in registers), parks the tile in
//          shared memory and re-reads it pixel-wise: one thread sums one 16-channel chunk of one pixel
//          sequentially, then one thread per pixel folds the chunk partials in ATen's cascade order.
// K3-NHWC  thread = one 16-byte vector (VEC channels of one pixel): the pixel's bit-width selects a row
//          of the CTA's {scale, zero_point, 1/scale} table (all channels, shared memory), one
//          LDG.128 / STG.128 per vector.
#include "common.cuh"

namespace mcaq {

constexpr int NH_THREADS = 256;

// ------------------------------------------------------------------------------------------- K1
template <typename T, int VEC, bool RANGES>
__global__ void __launch_bounds__(NH_THREADS)
reduce_planes_nhwc_kernel(const T* __restrict__ x, long long npix, int C, int PT, float* __restrict__ sum_plane,
                          float* __restrict__ abs_plane, int* __restrict__ keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int vpp = C / VEC;                                  // vectors per pixel (power of two <= 256)
  const int vshift = 31 - __clz(vpp);
  const int row = vpp + 1;                                  // padded row (uint4 units): conflict-free pixel-wise reads
  uint4* tile = reinterpret_cast<uint4*>(smem_raw);         // [PT][row]
  float* part = reinterpret_cast<float*>(tile + PT * row);  // [2][PT][C/16] chunk partials
  int* skeys = reinterpret_cast<int*>(part + 2 * PT * (C >> 4));   // [2C] CTA-wide range keys
  const int tid = threadIdx.x;
  const int nchunk = C >> 4;                                // C % 16 == 0, power of two like vpp
  const int cshift = 31 - __clz(nchunk);
  const int vpc = 16 / VEC;                                 // vectors per 16-channel chunk (2 bf16, 4 fp32)
  const long long ntiles = (npix + PT - 1) / PT;

  // running min / max of this thread's VEC channels (its channel group never changes: 256 % vpp == 0)
  float lo[VEC], hi[VEC];
  __nv_bfloat162 lo2[4], hi2[4];
  if (RANGES) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) { lo[e] = INFINITY; hi[e] = -INFINITY; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t pinf = 0x7f807f80u, ninf = 0xff80ff80u;
      lo2[i] = *reinterpret_cast<const __nv_bfloat162*>(&pinf);
      hi2[i] = *reinterpret_cast<const __nv_bfloat162*>(&ninf);
    }
    for (int i = tid; i < 2 * C; i += NH_THREADS) skeys[i] = i < C ? MCAQ_KEY_POS_INF : MCAQ_KEY_NEG_INF;
  }

  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long p0 = t * PT;
    const int np = (int)min((long long)PT, npix - p0);
    const int nv = np * vpp;
    const uint4* src = reinterpret_cast<const uint4*>(x + p0 * C);
    // phase A: coalesced loads, range update, park in shared memory
    for (int v0 = tid; v0 < nv; v0 += 4 * NH_THREADS) {
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v0 + u * NH_THREADS < nv) r[u] = ldg_stream(src + v0 + u * NH_THREADS);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * NH_THREADS;
        if (v < nv) {
          const int p = v >> vshift, k = v & (vpp - 1);
          tile[p * row + k] = r[u];
          if (RANGES) {
            if constexpr (VEC == 8) {
              const __nv_bfloat162* w = reinterpret_cast<const __nv_bfloat162*>(&r[u]);
#pragma unroll
              for (int i = 0; i < 4; ++i) { lo2[i] = __hmin2(lo2[i], w[i]); hi2[i] = __hmax2(hi2[i], w[i]); }
            } else {
              float d[VEC];
              Elem<T>::unpack(r[u], d);
#pragma unroll
              for (int e = 0; e < VEC; ++e) { lo[e] = fminf(lo[e], d[e]); hi[e] = fmaxf(hi[e], d[e]); }
            }
          }
        }
      }
    }
    __syncthreads();
    // phase B1: one thread per (pixel, 16-channel chunk): sequential sums of x and |x|
    for (int task = tid; task < np * nchunk; task += NH_THREADS) {
      const int p = task >> cshift, gch = task & (nchunk - 1);
      float s = 0.f, a = 0.f;
#pragma unroll
      for (int j = 0; j < vpc; ++j) {
        float d[VEC];
        Elem<T>::unpack(tile[p * row + gch * vpc + j], d);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { s = __fadd_rn(s, d[e]); a = __fadd_rn(a, fabsf(d[e])); }
      }
      part[p * nchunk + gch] = s;
      part[(PT + p) * nchunk + gch] = a;
    }
    __syncthreads();
    // phase B2: fold the chunk partials in ATen's cascade order (acc1 += P_g, acc2 += acc1 every 16)
    for (int o = tid; o < 2 * np; o += NH_THREADS) {
      const int plane = o >= np, p = plane ? o - np : o;
      const float* pp = part + (plane * PT + p) * nchunk;
      float acc1 = 0.f, acc2 = 0.f;
      for (int gch = 0; gch < nchunk; ++gch) {
        acc1 = __fadd_rn(acc1, pp[gch]);
        if (((gch + 1) & 15) == 0) { acc2 = __fadd_rn(acc2, acc1); acc1 = 0.f; }
      }
      const float r = __fadd_rn(__fadd_rn(0.f, acc1), acc2);
      (plane ? abs_plane : sum_plane)[p0 + p] = r;
    }
    __syncthreads();
  }

  if (RANGES) {
    const int k = tid & (vpp - 1);                          // this thread's vector slot inside a pixel
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float mn, mx;
      if constexpr (VEC == 8) {
        mn = (e & 1) ? __high2float(lo2[e >> 1]) : __low2float(lo2[e >> 1]);
        mx = (e & 1) ? __high2float(hi2[e >> 1]) : __low2float(hi2[e >> 1]);
      } else {
        mn = lo[e];
        mx = hi[e];
      }
      atomicMin(&skeys[k * VEC + e], float_key(mn));
      atomicMax(&skeys[C + k * VEC + e], float_key(mx));
    }
    __syncthreads();
    for (int i = tid; i < C; i += NH_THREADS) {
      atomicMin(keys + i, skeys[i]);
      atomicMax(keys + C + i, skeys[C + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------- K3
struct NhRanges {
  const float* packed;
  const float* rmin;
  const float* rmax;
};

template <typename T, int VEC, bool HAS_MASK>
__global__ void __launch_bounds__(NH_THREADS)
tile_quantize_nhwc_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int C, int H, int W, int Ht, int Wt,
                          float sy, float sx, const float* __restrict__ bit_map, NhRanges rg,
                          const float* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // {scale, zero_point, RN(1/scale), -} laid out [7][VEC][C / VEC]: for a fixed element index the
  // lanes of a warp (consecutive vector slots k) read consecutive 16-byte entries -- conflict free
  float4* tab = reinterpret_cast<float4*>(smem_raw);
  const int vpp_t = C / VEC;
  for (int i = threadIdx.x; i < 7 * C; i += NH_THREADS) {
    const int bi = i / C, r = i - bi * C;
    const int e_ = r / vpp_t, k_ = r - e_ * vpp_t;
    const int c = k_ * VEC + e_;
    const float mn = rg.packed ? __ldg(rg.packed + c) : __ldg(rg.rmin + c);
    const float mx = rg.packed ? -__ldg(rg.packed + C + c) : __ldg(rg.rmax + c);
    const int half = 1 << (bi + 1);
    const float qmin = -(float)half, qmax = (float)(half - 1);
    const float rng = fmaxf(__fsub_rn(mx, mn), 1e-8f);
    const float scale = __fdiv_rn(rng, __fsub_rn(qmax, qmin));
    const float zp = fminf(fmaxf(__fsub_rn(qmin, __fdiv_rn(mn, scale)), qmin), qmax);
    tab[i] = make_float4(scale, zp, __frcp_rn(scale), 0.f);
  }
  __syncthreads();
  // a CTA walks a contiguous range of pixels; a thread keeps its vector slot k and steps through the
  // pixels ppi at a time, carrying (b, h, w) along instead of dividing
  const int vpp = C / VEC;
  const int vshift = 31 - __clz(vpp);
  const int ppi = NH_THREADS >> vshift;                       // pixels per CTA iteration
  const long long npix = (long long)B * H * W;
  const long long per_cta = ((npix + gridDim.x - 1) / gridDim.x + ppi - 1) / ppi * ppi;
  const long long P0 = (long long)blockIdx.x * per_cta;
  const long long P1 = min(P0 + per_cta, npix);
  const int k = threadIdx.x & (vpp - 1);
  long long pg = P0 + (threadIdx.x >> vshift);
  if (pg >= P1) return;
  int b = (int)(pg / ((long long)H * W));
  int pix = (int)(pg - (long long)b * H * W);
  int h = pix / W, w = pix - h * W;
  for (; pg < P1; pg += ppi) {
    const int ty = nearest_src(h, sy, Ht), tx = nearest_src(w, sx, Wt);
    float bf = rintf(__ldg(bit_map + ((long long)b * Ht + ty) * Wt + tx));
    bf = fminf(fmaxf(bf, 2.f), 8.f);
    const int bidx = (int)bf - 2;
    const int half = 1 << (bidx + 1);
    const float qmin = -(float)half, qmax = (float)(half - 1);
    const float m = HAS_MASK ? __ldg(mask + pg) : 1.f;
    const long long v = (pg << vshift) + k;
    const uint4 raw = ldg_noalloc(reinterpret_cast<const uint4*>(x) + v);
    float xv[VEC], out[VEC];
    Elem<T>::unpack(raw, xv);
    const float4* trow = tab + bidx * C + k;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float4 p = trow[e * vpp];
      const float q = quant_code_fast(xv[e], p.x, p.y, p.z, qmin, qmax);
      float d = dequant(q, p.x, p.y);
      if (HAS_MASK) d = __fmul_rn(d, m);
      out[e] = d;
    }
    stg_stream(reinterpret_cast<uint4*>(y) + v, Elem<T>::pack(out));
    w += ppi;
    while (w >= W) {
      w -= W;
      if (++h == H) { h = 0; ++b; }
    }
  }
}

static int g_nh_sms = 0;
static int nh_sms() {
  if (!g_nh_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_nh_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_nh_sms <= 0) g_nh_sms = 148;
  }
  return g_nh_sms;
}

template <typename T, int VEC>
static int launch_reduce_nhwc(const T* x, int B, int C, int H, int W, float* sp, float* ap, int* keys,
                              cudaStream_t st) {
  const long long npix = (long long)B * H * W;
  const int vpp = C / VEC;
  // pixels per tile: about 32 KB of data, at most one pixel per thread
  int PT = (int)(32768 / ((long long)C * sizeof(T)));
  if (PT > NH_THREADS) PT = NH_THREADS;
  if (PT < 8) PT = 8;
  const size_t smem = (size_t)PT * (vpp + 1) * 16 + (size_t)2 * PT * (C >> 4) * 4 + (size_t)2 * C * 4;
  long long grid = (long long)nh_sms() * 4;
  const long long ntiles = (npix + PT - 1) / PT;
  if (grid > ntiles) grid = ntiles;
  if (keys) {
    auto k = reduce_planes_nhwc_kernel<T, VEC, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)grid, NH_THREADS, smem, st>>>(x, npix, C, PT, sp, ap, keys);
  } else {
    auto k = reduce_planes_nhwc_kernel<T, VEC, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)grid, NH_THREADS, smem, st>>>(x, npix, C, PT, sp, ap, keys);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

template <typename T, int VEC>
static int launch_quant_nhwc(const T* x, T* y, int B, int C, int H, int W, const float* bit_map, int Ht, int Wt,
                             NhRanges rg, const float* mask, cudaStream_t st) {
  const long long nvec = (long long)B * H * W * (C / VEC);
  const size_t smem = (size_t)7 * C * 16;
  // persistent CTAs (the per-CTA table costs 7*C entries): at least ~16 vectors per thread
  long long grid = (nvec + NH_THREADS * 16 - 1) / (NH_THREADS * 16);
  const long long cap = (long long)nh_sms() * (smem > 40 * 1024 ? 3 : 5);
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const float sy = (float)Ht / (float)H, sx = (float)Wt / (float)W;
  if (mask) {
    auto k = tile_quantize_nhwc_kernel<T, VEC, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)grid, NH_THREADS, smem, st>>>(x, y, B, C, H, W, Ht, Wt, sy, sx, bit_map, rg, mask);
  } else {
    auto k = tile_quantize_nhwc_kernel<T, VEC, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)grid, NH_THREADS, smem, st>>>(x, y, B, C, H, W, Ht, Wt, sy, sx, bit_map, rg, mask);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

// geometry the NHWC kernels cover: C a multiple of 16 with C / VEC a power of two <= 256 threads
static bool nhwc_ok(const void* a, const void* b, int C, int vec) {
  if (((uintptr_t)a & 15) || ((uintptr_t)b & 15) || C % 16) return false;
  const int vpp = C / vec;
  return vpp >= 1 && vpp <= NH_THREADS && (vpp & (vpp - 1)) == 0;
}

}  // namespace mcaq

using namespace mcaq;

extern "C" int mcaq_reduce_planes_nhwc(const void* x, int dtype, int B, int C, int H, int W, float* sum_plane,
                                       float* abs_plane, int32_t* keys, void* stream) {
  if (!x || !sum_plane || !abs_plane || B <= 0 || C <= 0 || H <= 0 || W <= 0) return MCAQ_EINVAL;
  if (C >= 4096) return MCAQ_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MCAQ_F32) {
    if (!nhwc_ok(x, x, C, 4)) return MCAQ_EALIGN;
    return launch_reduce_nhwc<float, 4>((const float*)x, B, C, H, W, sum_plane, abs_plane, keys, st);
  }
  if (dtype == MCAQ_BF16) {
    if (!nhwc_ok(x, x, C, 8)) return MCAQ_EALIGN;
    return launch_reduce_nhwc<__nv_bfloat16, 8>((const __nv_bfloat16*)x, B, C, H, W, sum_plane, abs_plane, keys, st);
  }
  return MCAQ_EDTYPE;
}

extern "C" int mcaq_tile_quantize_ranges_nhwc(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                              const float* bit_map, int Ht, int Wt, const float* packed,
                                              const float* running_min, const float* running_max,
                                              const float* mask, void* stream) {
  if (!x || !y || !bit_map || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return MCAQ_EINVAL;
  if (!packed && (!running_min || !running_max)) return MCAQ_EINVAL;
  if ((size_t)7 * C * 16 > 200 * 1024) return MCAQ_ETOOBIG;
  cudaStream_t st = (cudaStream_t)stream;
  NhRanges rg{packed, running_min, running_max};
  if (dtype == MCAQ_F32) {
    if (!nhwc_ok(x, y, C, 4)) return MCAQ_EALIGN;
    return launch_quant_nhwc<float, 4>((const float*)x, (float*)y, B, C, H, W, bit_map, Ht, Wt, rg, mask, st);
  }
  if (dtype == MCAQ_BF16) {
    if (!nhwc_ok(x, y, C, 8)) return MCAQ_EALIGN;
    typedef __nv_bfloat16 bf;
    return launch_quant_nhwc<bf, 8>((const bf*)x, (bf*)y, B, C, H, W, bit_map, Ht, Wt, rg, mask, st);
  }
  return MCAQ_EDTYPE;
}
