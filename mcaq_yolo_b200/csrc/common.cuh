// Shared device helpers for the MCAQ sm_100a kernels.
#pragma once
#ifndef K3_MAGIC
#define K3_MAGIC 0
#endif
#ifndef K3_MINB
#define K3_MINB 6
#endif

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "mcaq_b200.h"

#define MCAQ_LAUNCH_CHECK()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return (int)e__;                  \
  } while (0)

namespace mcaq {

// ---- order-preserving float <-> int key (for integer atomicMin/Max and REDUX) -------------
__device__ __forceinline__ int float_key(float f) {
  int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key_float(int k) {
  return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}
#define MCAQ_KEY_POS_INF 0x7f800000            /* float_key(+inf) */
#define MCAQ_KEY_NEG_INF ((int)0x807fffff)     /* float_key(-inf) */

// ---- 128-bit streaming loads / stores -----------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// coherent (in-place safe) streaming load: no L1 allocation, but not the read-only path
__device__ __forceinline__ uint4 ldg_noalloc(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_plain(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// L2 residency hints (experiment switch MCAQ_L2_HINTS): the channel sweep marks x evict_last so that the quantise
// sweep of the same scale can still find it in the 126 MB L2, which then reads it evict_first
#ifndef MCAQ_L2_HINTS
#define MCAQ_L2_HINTS 0
#endif
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ldg_stream_hint(const void* p, unsigned long long pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint4 ldg_noalloc_hint(const void* p, unsigned long long pol) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- element traits: VEC elements per 16-byte vector --------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int VEC = 4;
  __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]),
                      __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  __device__ __forceinline__ static float load1(const float* p) { return *p; }
  __device__ __forceinline__ static float round1(float v) { return v; }      // value as stored
  __device__ __forceinline__ static void store1(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int VEC = 8;
  // bf16 -> fp32 is exact: the 16 bits are the high half of the fp32 pattern
  __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
  }
  __device__ __forceinline__ static uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);   // round-to-nearest-even
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
  __device__ __forceinline__ static float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ __forceinline__ static float round1(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  __device__ __forceinline__ static void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct Elem<__half> {
  static constexpr int VEC = 8;
  // fp16 -> fp32 is exact (one HADD2.F32 per pair)
  __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  __device__ __forceinline__ static uint32_t pack2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);              // round-to-nearest-even
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
  __device__ __forceinline__ static float load1(const __half* p) { return __half2float(*p); }
  __device__ __forceinline__ static float round1(float v) { return __half2float(__float2half_rn(v)); }
  __device__ __forceinline__ static void store1(__half* p, float v) { *p = __float2half_rn(v); }
};

// element type of a dtype code with 16-bit storage (MCAQ_BF16 / MCAQ_F16): dispatch helper
#define MCAQ_DISPATCH_16(dtype, T16, ...)                              \
  do {                                                                 \
    if ((dtype) == MCAQ_BF16) { typedef __nv_bfloat16 T16; __VA_ARGS__; } \
    else { typedef __half T16; __VA_ARGS__; }                          \
  } while (0)

// ---- F.interpolate(mode='nearest') source index: min(floor(dst * (float)in/out), in-1) -----
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  int i = (int)floorf(__fmul_rn((float)dst, scale));
  return i < in_size - 1 ? i : in_size - 1;
}

// quantiser for one element; every operation is a separate IEEE fp32 rounding
// (quantization.py:597-600): q = clamp(rint(x/scale + zp)), deq = (q - zp) * scale
__device__ __forceinline__ float quant_code(float x, float scale, float zp, float qmin, float qmax) {
  float t = __fadd_rn(__fdiv_rn(x, scale), zp);
  float q = rintf(t);
  return fminf(fmaxf(q, qmin), qmax);
}
// Same code with the division done as Markstein's correction step: with rinv = RN(1/scale),
//   q0 = RN(x*rinv), r = x - q0*scale (exact, FMA), q = RN(q0 + r*rinv) == RN(x/scale)
// for finite operands in the normal range (|x| < 2^90 here; tests/test_gpu_division.py sweeps it
// against div.rn).  Three FMA-pipe instructions instead of div.rn's check-and-branch sequence.
__device__ __forceinline__ float div_markstein(float x, float scale, float rinv) {
  const float q0 = __fmul_rn(x, rinv);
  const float r = fmaf(-q0, scale, x);
  return fmaf(r, rinv, q0);
}
__device__ __forceinline__ float quant_code_fast(float x, float scale, float zp, float rinv, float qmin,
                                                 float qmax) {
  const float t = __fadd_rn(div_markstein(x, scale, rinv), zp);
#if K3_MAGIC
  // rint on the FMA pipe (the conversion unit runs at 1/8 rate): (t + 1.5*2^23) - 1.5*2^23 is the
  // round-half-even integer for |t| < 2^22; beyond that the result only has to stay outside
  // [qmin, qmax] (|q| <= 128), which it does; copysign restores rintf's signed zero
  const float r = copysignf(__fsub_rn(__fadd_rn(t, 12582912.f), 12582912.f), t);
  return fminf(fmaxf(r, qmin), qmax);
#else
  return fminf(fmaxf(rintf(t), qmin), qmax);
#endif
}
__device__ __forceinline__ float dequant(float q, float scale, float zp) {
  return __fmul_rn(__fsub_rn(q, zp), scale);
}

// ---- packed fp32 pairs (sm_100: FMUL2 / FADD2 / FFMA2) ---------------------------------------
// One instruction performs two independent IEEE round-to-nearest fp32 operations on a 64-bit
// register pair, so results are bit-identical to the scalar forms at half the issue slots.  The
// compiler does not form these on its own; ptxas maps the {lo, hi} moves onto adjacent registers.
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mul.rn.f32x2 rd, ra, rb;\n"
      " mov.b64 {%0,%1}, rd;}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n add.rn.f32x2 rd, ra, rb;\n"
      " mov.b64 {%0,%1}, rd;}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
// a + b where a or b is itself a packed product: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into
// one FFMA2 even under -fmad=false (seen in the SASS: one rounding instead of two; it also folds
// a * 1 + b back).  Scalar add.rn after a packed multiply is left alone, so these adds stay scalar.
__device__ __forceinline__ float2 fadd2_sep(float2 a, float2 b) {
  return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}

// quant_code_fast for two elements: the same seven roundings per element, issued as pairs
__device__ __forceinline__ float2 quant_code_fast2(float2 x, float scale, float zp, float rinv, float qmin,
                                                   float qmax) {
  const float2 s2 = splat2(scale), r2 = splat2(rinv);
  const float2 q0 = fmul2(x, r2);
  const float2 r = ffma2(make_float2(-q0.x, -q0.y), s2, x);
  const float2 t = fadd2(ffma2(r, r2, q0), splat2(zp));
  // (rint as two packed adds of 1.5 * 2^23 instead of two FRND measured no faster: profiles/r02_k3_variants.txt)
  return make_float2(fminf(fmaxf(rintf(t.x), qmin), qmax), fminf(fmaxf(rintf(t.y), qmin), qmax));
}
__device__ __forceinline__ float2 dequant2(float2 q, float scale, float zp) {
  return fmul2(fadd2(q, splat2(-zp)), splat2(scale));
}

}  // namespace mcaq
