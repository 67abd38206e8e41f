// K1: one HBM sweep of an NCHW feature map producing
//   sum_c x, sum_c |x| per pixel (fp32, torch-CPU cascade order) and per-channel min/max.
//
// Layout of the work: a "strip" is 32 consecutive 16-byte pixel vectors of the flattened [B, HW] batch (it may
// straddle two images: every lane derives its own image / pixel, so only the very last strip is ragged); warp g
// of a CTA owns one 16-channel chunk of the strip per pass (16 independent LDG.128 in flight
// per thread), sums it sequentially (chunk partial P_g), reduces the chunk's per-channel
// min/max across the warp (REDUX on order-preserving integer keys per vector, or -- when a warp always owns the
// same 16 channels -- register accumulators joined once per CTA by a transposing butterfly), and parks P_g in shared
// memory.  The CTA then folds the partials in chunk order exactly like ATen's multi_row_sum
// (acc1 += P_g, acc2 += acc1 every 16 chunks) so the planes are bit-identical to
// x.mean(1) * C on the reference's CPU path.  CTAs are persistent over strips; per-channel
// ranges are kept in shared memory and flushed with 2*C global atomics per CTA at the end.
//
// HBM traffic: reads x once (B*C*H*W*s bytes), writes 8 bytes per pixel.
#include "common.cuh"

namespace mcaq {

// raw (still packed) vector kept in registers between the load burst and its use
template <typename T, int VEC>
struct VecIO {
  static_assert(VEC == Elem<T>::VEC, "vector width");
  typedef uint4 Raw;
#if MCAQ_L2_HINTS
  __device__ __forceinline__ static Raw load(const T* p) { return ldg_stream_hint(p, l2_policy_evict_last()); }
#else
  __device__ __forceinline__ static Raw load(const T* p) { return ldg_stream(p); }
#endif
  __device__ __forceinline__ static Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ static void unpack(const Raw& r, float* f) { Elem<T>::unpack(r, f); }
};
template <typename T>
struct VecIO<T, 1> {
  typedef float Raw;
  __device__ __forceinline__ static Raw load(const T* p) { return Elem<T>::load1(p); }
  __device__ __forceinline__ static Raw zero() { return 0.f; }
  __device__ __forceinline__ static void unpack(const Raw& r, float* f) { f[0] = r; }
};

// running min / max of everything a thread has seen of one channel.  bf16 keeps two packed partials
// (even / odd elements of the vectors) and updates them with 2-wide HMNMX2 straight on the raw words;
// the extremes of bf16 data are bf16 values, so nothing is lost.  fp32 keeps plain floats.
template <typename T, int VEC>
struct MinMaxAcc {
  float lo, hi;
  __device__ __forceinline__ void init() { lo = INFINITY; hi = -INFINITY; }
  __device__ __forceinline__ void update(const typename VecIO<T, VEC>::Raw& r) {
    float d[VEC];
    VecIO<T, VEC>::unpack(r, d);
#pragma unroll
    for (int e = 0; e < VEC; ++e) { lo = fminf(lo, d[e]); hi = fmaxf(hi, d[e]); }
  }
  __device__ __forceinline__ float vmin() const { return lo; }
  __device__ __forceinline__ float vmax() const { return hi; }
  __device__ __forceinline__ void merge(const MinMaxAcc& o) { lo = fminf(lo, o.lo); hi = fmaxf(hi, o.hi); }
  __device__ __forceinline__ static MinMaxAcc select(bool c, const MinMaxAcc& a, const MinMaxAcc& b) {
    MinMaxAcc r;
    r.lo = c ? a.lo : b.lo;
    r.hi = c ? a.hi : b.hi;
    return r;
  }
  __device__ __forceinline__ MinMaxAcc shfl_xor(int mask) const {
    MinMaxAcc r;
    r.lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    r.hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return r;
  }
};
// 16-bit types: the partials live as raw 32-bit words (two packed values), so that selects and shuffles of the
// butterfly are plain register moves
template <typename P2, uint32_t POS_INF2, uint32_t NEG_INF2>
struct MinMaxAcc16 {
  uint32_t lo, hi;
  __device__ __forceinline__ static P2 as2(uint32_t w) { return *reinterpret_cast<P2*>(&w); }
  __device__ __forceinline__ static uint32_t raw(P2 v) { return *reinterpret_cast<uint32_t*>(&v); }
  __device__ __forceinline__ void init() { lo = POS_INF2; hi = NEG_INF2; }
  __device__ __forceinline__ void update(const uint4& r) {
    const P2 a = as2(r.x), b = as2(r.y), c = as2(r.z), d = as2(r.w);
    lo = raw(__hmin2(as2(lo), __hmin2(__hmin2(a, b), __hmin2(c, d))));
    hi = raw(__hmax2(as2(hi), __hmax2(__hmax2(a, b), __hmax2(c, d))));
  }
  __device__ __forceinline__ float vmin() const { return fminf(__low2float(as2(lo)), __high2float(as2(lo))); }
  __device__ __forceinline__ float vmax() const { return fmaxf(__low2float(as2(hi)), __high2float(as2(hi))); }
  __device__ __forceinline__ void merge(const MinMaxAcc16& o) {
    lo = raw(__hmin2(as2(lo), as2(o.lo)));
    hi = raw(__hmax2(as2(hi), as2(o.hi)));
  }
  __device__ __forceinline__ static MinMaxAcc16 select(bool c, const MinMaxAcc16& a, const MinMaxAcc16& b) {
    MinMaxAcc16 r;
    r.lo = c ? a.lo : b.lo;
    r.hi = c ? a.hi : b.hi;
    return r;
  }
  __device__ __forceinline__ MinMaxAcc16 shfl_xor(int mask) const {
    MinMaxAcc16 r;
    r.lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    r.hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return r;
  }
};
template <>
struct MinMaxAcc<__nv_bfloat16, 8> : MinMaxAcc16<__nv_bfloat162, 0x7f807f80u, 0xff80ff80u> {};
template <>
struct MinMaxAcc<__half, 8> : MinMaxAcc16<__half2, 0x7c007c00u, 0xfc00fc00u> {};

// One round of the transposing butterfly: the 2*N channel accumulators of every lane become N, lane pairs
// (l, l ^ MASK) splitting the channels between them (bit MASK of the lane set: keeps the upper half)
template <int N, int MASK, typename A>
__device__ __forceinline__ void butterfly_round(A (&v)[16], int lane) {
  const bool up = (lane & MASK) != 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const auto send = A::select(up, v[i], v[i + N]);
    auto keep = A::select(up, v[i + N], v[i]);
    keep.merge(send.shfl_xor(MASK));
    v[i].lo = keep.lo;
    v[i].hi = keep.hi;
  }
}

// RMODE: 0 no ranges, 1 per-vector warp reduction (any C), 2 per-thread running min / max in
// registers when the CTA's G warps cover all channel chunks in one pass (a warp then always owns
// the same 16 channels), reduced across the warp once per CTA
// 512 resident threads per SM at <= 128 registers for every variant (G = 4: four CTAs, G = 8: two,
// G = 16: one): more CTAs with 16 loads in flight each beat fewer CTAs with more registers
template <typename T, int VEC, int G, int RMODE>
__global__ void __launch_bounds__(32 * G, (16 / G) > 0 ? (16 / G) : 1)
reduce_planes_kernel(const T* __restrict__ x, int B, int C, int HW,
                     float* __restrict__ sum_plane, float* __restrict__ abs_plane,
                     int* __restrict__ keys, long long total_vec, long long total_strips) {
  constexpr int NT = 32 * G;
  constexpr int STRIP = 32 * VEC;                 // pixels per strip
  constexpr int NOUT = (2 * STRIP + NT - 1) / NT; // outputs owned per thread in the fold
  constexpr bool RANGES = RMODE == 1;             // shared-memory range table + per-vector REDUX
  constexpr bool ACC = RMODE == 2;                // register accumulators, npass == 1
  MinMaxAcc<T, VEC> mm[16];
  if (ACC) {
#pragma unroll
    for (int j = 0; j < 16; ++j) mm[j].init();
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* part = reinterpret_cast<float*>(smem_raw);           // [2][G][STRIP]
  int* smin = reinterpret_cast<int*>(part + 2 * G * STRIP);   // [C]
  int* smax = smin + C;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nfull = C >> 4;                       // full 16-channel chunks
  const int tail = C & 15;
  const int ngroups = nfull + (tail ? 1 : 0);
  const int npass = (ngroups + G - 1) / G;
  const int nvec = (HW + VEC - 1) / VEC;          // VEC==1 or HW % VEC == 0
  const long long HWl = HW;

  if (RANGES) {
    for (int c = threadIdx.x; c < C; c += NT) { smin[c] = MCAQ_KEY_POS_INF; smax[c] = MCAQ_KEY_NEG_INF; }
    __syncthreads();
  }

  for (long long strip = blockIdx.x; strip < total_strips; strip += gridDim.x) {
    // strips run over the flattened (image, pixel vector) index: HW % VEC == 0, so vector gv of the batch is
    // pixels [gv * VEC, gv * VEC + VEC) of the contiguous [B, HW] planes and only the last strip is ragged
    const long long gv = strip * 32 + lane;
    const bool active = gv < total_vec;
    const long long gvc = active ? gv : total_vec - 1;
    int b, v;
    if (total_vec <= 0x7fffffffLL) {
      b = (int)((unsigned)gvc / (unsigned)nvec);
      v = (int)((unsigned)gvc - (unsigned)b * (unsigned)nvec);
    } else {
      b = (int)(gvc / nvec);
      v = (int)(gvc - (long long)b * nvec);
    }
    const T* xb = x + ((long long)b * C) * HW + (long long)v * VEC;

    // CTA-uniform: all of C in one pass of full chunks and no ragged lane in the strip
    const bool fastfold = VEC > 1 && npass == 1 && tail == 0 && nfull == G && (strip + 1) * 32 <= total_vec;
    float acc0[NOUT], acc1[NOUT], acc2[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; acc2[k] = 0.f; }

#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
      const int gi = pass * G + warp;             // chunk index of this warp
      const int c0 = gi << 4;
      float s[VEC], a[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) { s[e] = 0.f; a[e] = 0.f; }
      if (RMODE != 1 && gi < nfull && __all_sync(0xffffffffu, active)) {
        // common case (full 16-channel chunk, whole strip inside the image): no predicates, no zero
        // fill, running pointer
        typename VecIO<T, VEC>::Raw raw[16];
        const T* pj = xb + (long long)c0 * HWl;
#pragma unroll
        for (int j = 0; j < 16; ++j) { raw[j] = VecIO<T, VEC>::load(pj); pj += HWl; }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float d[VEC];
          VecIO<T, VEC>::unpack(raw[j], d);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            s[e] = __fadd_rn(s[e], d[e]);
            a[e] = __fadd_rn(a[e], fabsf(d[e]));
          }
          if (ACC) mm[j].update(raw[j]);
        }
      } else if (gi < ngroups) {
        const int nch = (gi < nfull) ? 16 : tail;
        typename VecIO<T, VEC>::Raw raw[16];
        if (nch == 16) {
          // running pointer: one 64-bit add per channel instead of re-deriving xb + (c0 + j) * HW
          const T* pj = xb + (long long)c0 * HWl;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            raw[j] = active ? VecIO<T, VEC>::load(pj) : VecIO<T, VEC>::zero();
            pj += HWl;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            raw[j] = (active && j < nch) ? VecIO<T, VEC>::load(xb + (long long)(c0 + j) * HW)
                                         : VecIO<T, VEC>::zero();
        }
        int mymin = MCAQ_KEY_POS_INF, mymax = MCAQ_KEY_NEG_INF;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float d[VEC];
          VecIO<T, VEC>::unpack(raw[j], d);
          float lo = d[0], hi = d[0];
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            // 0 + x0 == x0, so starting from 0 reproduces the sequential chunk sum
            s[e] = __fadd_rn(s[e], d[e]);
            a[e] = __fadd_rn(a[e], fabsf(d[e]));
            lo = fminf(lo, d[e]);
            hi = fmaxf(hi, d[e]);
          }
          if (ACC) {
            if (active && j < nch) mm[j].update(raw[j]);
          }
          if (RANGES) {
            const bool ok = active && j < nch;
            int kmin = __reduce_min_sync(0xffffffffu, ok ? float_key(lo) : MCAQ_KEY_POS_INF);
            int kmax = __reduce_max_sync(0xffffffffu, ok ? float_key(hi) : MCAQ_KEY_NEG_INF);
            if (lane == j) { mymin = kmin; mymax = kmax; }
          }
        }
        if (RANGES && lane < nch) {               // this warp is the only owner of channel c0+lane
          const int c = c0 + lane;
          smin[c] = min(smin[c], mymin);
          smax[c] = max(smax[c], mymax);
        }
      }
      // park the chunk partial: part[plane][warp][lane*VEC + e]
      float* ps = part + (0 * G + warp) * STRIP + lane * VEC;
      float* pa = part + (1 * G + warp) * STRIP + lane * VEC;
#pragma unroll
      for (int e = 0; e < VEC; ++e) { ps[e] = s[e]; pa[e] = a[e]; }
      __syncthreads();
      if (fastfold) {
        // every chunk is full, one pass covers C and the strip lies inside the image (every YOLO
        // width): thread t folds FV consecutive outputs, P_0 + P_1 + ... in chunk order starting
        // from 0 -- the value of the generic fold below ((0 + acc1) + 0, or 0 + acc2 after 16 chunks)
        constexpr int NPT = 2 * STRIP / NT;                    // outputs per thread: 2 * VEC / G
        constexpr int FV = NPT >= 4 ? 4 : (NPT >= 2 ? 2 : 1);
        constexpr int NIT = NPT >= 4 ? NPT / 4 : 1;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int o = (it * NT + threadIdx.x) * FV;
          if (o < 2 * STRIP) {
            const int plane = o / STRIP, q = o - plane * STRIP;
            float r[FV];
#pragma unroll
            for (int e = 0; e < FV; ++e) r[e] = 0.f;
#pragma unroll
            for (int w = 0; w < G; ++w) {
              const float* src = part + (plane * G + w) * STRIP + q;
              float p[FV];
              if (FV == 4) *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(src);
              else if (FV == 2) *reinterpret_cast<float2*>(p) = *reinterpret_cast<const float2*>(src);
              else p[0] = src[0];
#pragma unroll
              for (int e = 0; e < FV; ++e) r[e] = __fadd_rn(r[e], p[e]);
            }
            float* dst = (plane ? abs_plane : sum_plane) + strip * STRIP + q;
            if (FV == 4) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<float4*>(r);
            else if (FV == 2) *reinterpret_cast<float2*>(dst) = *reinterpret_cast<float2*>(r);
            else dst[0] = r[0];
          }
        }
        __syncthreads();
        break;
      }
      // fold in chunk order (ATen multi_row_sum, level_step 16)
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        const int o = threadIdx.x + k * NT;
        if (o < 2 * STRIP) {
          const int plane = o / STRIP, q = o - plane * STRIP;
          for (int w = 0; w < G; ++w) {
            const int g2 = pass * G + w;
            if (g2 >= ngroups) break;
            const float p = part[(plane * G + w) * STRIP + q];
            if (g2 < nfull) {
              acc1[k] = __fadd_rn(acc1[k], p);
              if (((g2 + 1) & 15) == 0) { acc2[k] = __fadd_rn(acc2[k], acc1[k]); acc1[k] = 0.f; }
            } else {
              acc0[k] = p;                        // tail chunk (C % 16 channels)
            }
          }
        }
      }
      __syncthreads();
    }
    // final = ((acc0 + acc1) + acc2) [+ acc3 == 0 for C < 4096]
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      const int o = threadIdx.x + k * NT;
      if (!fastfold && o < 2 * STRIP) {
        const int plane = o / STRIP, q = o - plane * STRIP;
        const long long pix = strip * STRIP + q;
        if (pix < total_vec * VEC) {
          const float r = __fadd_rn(__fadd_rn(acc0[k], acc1[k]), acc2[k]);
          float* dst = plane ? abs_plane : sum_plane;
          dst[pix] = r;
        }
      }
    }
  }

  if (RANGES) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += NT) {
      atomicMin(keys + c, smin[c]);
      atomicMax(keys + C + c, smax[c]);
    }
  }
  if (ACC) {
    // one warp reduction per owned channel for the whole CTA; lane j publishes channel c0 + j
    const int gi = warp;
    if (gi < ngroups) {
      const int nch = (gi < nfull) ? 16 : tail;
      // transposing butterfly (31 shuffle pairs instead of 32 REDUX + key conversions): after the rounds
      // 16 / 8 / 4 / 2 lane l holds channel l >> 1 of half the warp, the last round joins the two halves
      butterfly_round<8, 16>(mm, lane);
      butterfly_round<4, 8>(mm, lane);
      butterfly_round<2, 4>(mm, lane);
      butterfly_round<1, 2>(mm, lane);
      mm[0].merge(mm[0].shfl_xor(1));
      const int j = lane >> 1;
      if (!(lane & 1) && j < nch) {
        atomicMin(keys + (gi << 4) + j, float_key(mm[0].vmin()));
        atomicMax(keys + C + (gi << 4) + j, float_key(mm[0].vmax()));
      }
    }
  }
}

__global__ void ranges_reset_kernel(int* keys, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) { keys[i] = MCAQ_KEY_POS_INF; keys[C + i] = MCAQ_KEY_NEG_INF; }
}

__global__ void ranges_decode_kernel(const int* keys, int C, float* packed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) { packed[i] = key_float(keys[i]); packed[C + i] = -key_float(keys[C + i]); }
}

// running <- momentum*running + (1-momentum)*new, separate mul/add roundings like torch
__global__ void ranges_ema_kernel(const float* packed, int C, float momentum, float one_minus, int first,
                                  float* rmin, float* rmax) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const float mn = packed[i], mx = -packed[C + i];
  if (first) { rmin[i] = mn; rmax[i] = mx; return; }
  rmin[i] = __fadd_rn(__fmul_rn(momentum, rmin[i]), __fmul_rn(one_minus, mn));
  rmax[i] = __fadd_rn(__fmul_rn(momentum, rmax[i]), __fmul_rn(one_minus, mx));
}

// K1's epilogue for training / calibration in ONE launch: decode the range keys into packed = [min, -max] and
// apply the EMA to running_min / running_max (quantization.py:319-353)
__global__ void ranges_finish_kernel(const int* keys, int C, float momentum, float one_minus, int first, float* rmin,
                                     float* rmax, float* packed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const float mn = key_float(keys[i]), mx = key_float(keys[C + i]);
  if (packed) { packed[i] = mn; packed[C + i] = -mx; }
  if (first) { rmin[i] = mn; rmax[i] = mx; return; }
  rmin[i] = __fadd_rn(__fmul_rn(momentum, rmin[i]), __fmul_rn(one_minus, mn));
  rmax[i] = __fadd_rn(__fmul_rn(momentum, rmax[i]), __fmul_rn(one_minus, mx));
}

// qtable[(b-2)*C + c] = {scale, zp}   (quantization.py:41-66)
__global__ void build_qtable_kernel(const float* packed, const float* rmin, const float* rmax, int C,
                                    float2* qtable) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 7 * C) return;
  const int bi = i / C, c = i - bi * C, bits = bi + 2;
  const float mn = packed ? packed[c] : rmin[c];
  const float mx = packed ? -packed[C + c] : rmax[c];
  const float qmin = -(float)(1 << (bits - 1));
  const float qmax = (float)((1 << (bits - 1)) - 1);
  float rng = fmaxf(__fsub_rn(mx, mn), 1e-8f);
  const float scale = __fdiv_rn(rng, __fsub_rn(qmax, qmin));
  float zp = __fsub_rn(qmin, __fdiv_rn(mn, scale));
  zp = fminf(fmaxf(zp, qmin), qmax);
  qtable[i] = make_float2(scale, zp);
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <typename T, int VEC, int G>
static int launch_reduce(const T* x, int B, int C, int HW, float* sp, float* ap, int* keys, cudaStream_t st) {
  const int nvec = (HW + VEC - 1) / VEC;
  const long long total_vec = (long long)B * nvec;
  const long long total = (total_vec + 31) / 32;
  const size_t smem = (size_t)2 * G * 32 * VEC * sizeof(float) + (size_t)2 * C * sizeof(int);
  // persistent CTAs, exactly the resident wave: 512 threads per SM at <= 128 registers (16 / G CTAs).  A
  // second wave only repeats the per-CTA prologue / epilogue (range join + 2C atomics): one wave measured
  // 7 % (C3) to 14 % (C4) faster
  const int per_sm = 16 / G > 0 ? 16 / G : 1;
  long long grid = (long long)num_sms() * per_sm;
  if (grid > total) grid = total;
  if (grid < 1) grid = 1;
  const int npass = ((C + 15) / 16 + G - 1) / G;
  void (*k)(const T*, int, int, int, float*, float*, int*, long long, long long);
  if (!keys) k = reduce_planes_kernel<T, VEC, G, 0>;
  else if (VEC > 1 && npass == 1) k = reduce_planes_kernel<T, VEC, G, 2>;
  else k = reduce_planes_kernel<T, VEC, G, 1>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<(unsigned)grid, 32 * G, smem, st>>>(x, B, C, HW, sp, ap, keys, total_vec, total);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

template <typename T, int VEC>
static int dispatch_groups(const T* x, int B, int C, int HW, float* sp, float* ap, int* keys, cudaStream_t st) {
  const int ngroups = (C + 15) / 16;
  // 16-warp CTAs (one pass over C >= 256, register range accumulators) take a whole SM's register file: nothing of the
  // morphology kernel can be resident beside them.  They pay only when a CTA sees a strip or two (C5 of v8n@640);
  // on larger maps 8-warp CTAs making two or more passes co-reside and win (v8s@1280 fp32: whole step 0.462 ->
  // 0.445 ms, profiles/r02_coreside_sweep.txt)
  const long long strips = ((long long)B * ((HW + VEC - 1) / VEC) + 31) / 32;
  if (ngroups >= 16 && strips < 2LL * num_sms()) return launch_reduce<T, VEC, 16>(x, B, C, HW, sp, ap, keys, st);
  if (ngroups >= 8) return launch_reduce<T, VEC, 8>(x, B, C, HW, sp, ap, keys, st);
  if (ngroups >= 3) return launch_reduce<T, VEC, 4>(x, B, C, HW, sp, ap, keys, st);
  if (ngroups == 2) return launch_reduce<T, VEC, 2>(x, B, C, HW, sp, ap, keys, st);
  return launch_reduce<T, VEC, 1>(x, B, C, HW, sp, ap, keys, st);
}

}  // namespace mcaq

using namespace mcaq;

extern "C" int mcaq_ranges_reset(int32_t* keys, int C, void* stream) {
  if (!keys || C <= 0) return MCAQ_EINVAL;
  ranges_reset_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(keys, C);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_reduce_planes(const void* x, int dtype, int B, int C, int H, int W,
                                  float* sum_plane, float* abs_plane, int32_t* keys, void* stream) {
  if (!x || !sum_plane || !abs_plane || B <= 0 || C <= 0 || H <= 0 || W <= 0) return MCAQ_EINVAL;
  if (C >= 4096) return MCAQ_EINVAL;             // third cascade level not implemented
  const long long hw = (long long)H * W;
  if (hw > 0x7fffffffLL) return MCAQ_EINVAL;
  const int HW = (int)hw;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = ((uintptr_t)x & 15) == 0;
  if (dtype == MCAQ_F32) {
    const float* p = (const float*)x;
    if (aligned && HW % 4 == 0) return dispatch_groups<float, 4>(p, B, C, HW, sum_plane, abs_plane, keys, st);
    return dispatch_groups<float, 1>(p, B, C, HW, sum_plane, abs_plane, keys, st);
  } else if (dtype == MCAQ_BF16) {
    const __nv_bfloat16* p = (const __nv_bfloat16*)x;
    if (aligned && HW % 8 == 0) return dispatch_groups<__nv_bfloat16, 8>(p, B, C, HW, sum_plane, abs_plane, keys, st);
    return dispatch_groups<__nv_bfloat16, 1>(p, B, C, HW, sum_plane, abs_plane, keys, st);
  } else if (dtype == MCAQ_F16) {
    const __half* p = (const __half*)x;
    if (aligned && HW % 8 == 0) return dispatch_groups<__half, 8>(p, B, C, HW, sum_plane, abs_plane, keys, st);
    return dispatch_groups<__half, 1>(p, B, C, HW, sum_plane, abs_plane, keys, st);
  }
  return MCAQ_EDTYPE;
}

extern "C" int mcaq_ranges_decode(const int32_t* keys, int C, float* packed, void* stream) {
  if (!keys || !packed || C <= 0) return MCAQ_EINVAL;
  ranges_decode_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(keys, C, packed);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_ranges_ema(const float* packed, int C, double momentum, int first,
                               float* running_min, float* running_max, void* stream) {
  if (!packed || !running_min || !running_max || C <= 0) return MCAQ_EINVAL;
  // momentum and (1 - momentum) are Python doubles in the reference; each is rounded to fp32
  // when it meets the fp32 tensor (quantization.py:346)
  const float one_minus = (float)(1.0 - momentum);
  ranges_ema_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(packed, C, (float)momentum, one_minus,
                                                                      first, running_min, running_max);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_ranges_finish(const int32_t* keys, int C, double momentum, int first, float* running_min,
                                  float* running_max, float* packed, void* stream) {
  if (!keys || !running_min || !running_max || C <= 0) return MCAQ_EINVAL;
  ranges_finish_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(keys, C, (float)momentum, (float)(1.0 - momentum),
                                                                         first, running_min, running_max, packed);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcaq_build_qtable(const float* packed, const float* running_min, const float* running_max,
                                 int C, float* qtable, void* stream) {
  if (!qtable || C <= 0 || (!packed && (!running_min || !running_max))) return MCAQ_EINVAL;
  build_qtable_kernel<<<(7 * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      packed, running_min, running_max, C, (float2*)qtable);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
