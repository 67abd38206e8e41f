// Device helpers shared by the two morphology paths: the fused per-image kernel (morph_fused.cu: planes up
// to 160 columns, tiles up to 32) and the plane pipeline for image-sized inputs (morph_planes.cu).
// Arithmetic contract = oracle/mcaq_oracle.py (see morph_fused.cu).
#pragma once
#include "common.cuh"
#include "mcaq_consts.cuh"

namespace mcaq {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// RN(x / d) with a precomputed rinv = RN(1/d): Markstein's correction in the normal range (swept
// against div.rn in tests/test_gpu_division.py), plain division for zero / tiny / huge numerators.
__device__ __forceinline__ float div_exact(float x, float d, float rinv) {
  const float ax = fabsf(x);
  if (ax >= 1e-30f && ax < 1e27f) return div_markstein(x, d, rinv);
  return __fdiv_rn(x, d);
}

// x / C for the channel mean: an exact scaling when C is a power of two (every YOLOv8 width)
__device__ __forceinline__ float div_channels(float x, float fC, float rC, bool pow2) {
  return pow2 ? __fmul_rn(x, rC) : div_exact(x, fC, rC);
}

// literal atan2 binning for the rare pixels within the guard band of a bin boundary (not inlined: rare)
static __device__ __noinline__ int nms_bin_exact(float gx, float gy) {
  float ang = __fmul_rn(atan2f(gy, gx), kc::RAD2DEG);
  if (ang < 0.f) ang = __fadd_rn(ang, 180.f);
  if (ang < 22.5f || ang >= 157.5f) return 0;
  if (ang < 67.5f) return 1;
  if (ang < 112.5f) return 2;
  return 3;
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
  return nms_bin_exact(gx, gy);
}

// Sobel responses from a 3x3 window of zero-padded values: FMA chain over the taps in row-major
// order (morphology.py:385-395); the centre tap is zero in both kernels.
__device__ __forceinline__ void sobel3(float z00, float z01, float z02, float z10, float z12, float z20, float z21,
                                       float z22, float& gx, float& gy) {
  float a = __fmul_rn(z00, -1.f);
  a = fmaf(z02, 1.f, a); a = fmaf(z10, -2.f, a); a = fmaf(z12, 2.f, a); a = fmaf(z20, -1.f, a);
  gx = fmaf(z22, 1.f, a);
  float c = __fmul_rn(z00, -1.f);
  c = fmaf(z01, -2.f, c); c = fmaf(z02, -1.f, c); c = fmaf(z20, 1.f, c); c = fmaf(z21, 2.f, c);
  gy = fmaf(z22, 1.f, c);
}

__device__ constexpr float ADAPT1[11] = {0x1.20c256p-7f, 0x1.bcb868p-6f, 0x1.0ab508p-4f, 0x1.f2464cp-4f, 0x1.6a7e1cp-3f,
                              0x1.9ac20ap-3f, 0x1.6a7e1cp-3f, 0x1.f2464cp-4f, 0x1.0ab508p-4f, 0x1.bcb868p-6f,
                              0x1.20c256p-7f};
constexpr float ADAPT_GUARD = 4e-3f;


// Otsu threshold from the whole-image histogram (morphology.py:397-418), computed by one warp; every
// lane returns the same values.  fp64 prefix sums of fp32 terms are exact.
__device__ __forceinline__ void otsu_warp(const int* hist, int lane, float& thr255, int& otsu_bin) {
  int cnt[8];
  int tot_i = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { cnt[j] = hist[lane * 8 + j]; tot_i += cnt[j]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot_i += __shfl_xor_sync(0xffffffffu, tot_i, o);
  const float tot = fmaxf((float)tot_i, 1.0f);
  float p[8], pc[8];
  double so = 0.0, sm = 0.0;
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
  const float thr = (float)(2 * best_i + 1) * 0.001953125f;
  thr255 = __fmul_rn(thr, 255.f);
  otsu_bin = best_i;
}


}  // namespace mcaq
