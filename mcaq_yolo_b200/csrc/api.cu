// Misc C-ABI entry points (version, error strings, geometry helpers).
#include "common.cuh"

extern "C" int mcaq_abi_version(void) { return MCAQ_ABI_VERSION; }

extern "C" const char* mcaq_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case MCAQ_EINVAL: return "mcaq: invalid argument (null pointer or non-positive size)";
    case MCAQ_EALIGN: return "mcaq: pointer is not 16-byte aligned";
    case MCAQ_ETOOBIG: return "mcaq: feature plane exceeds the on-chip budget of the morphology kernel";
    case MCAQ_EDTYPE: return "mcaq: unsupported dtype (expected MCAQ_F32, MCAQ_BF16 or MCAQ_F16)";
    case MCAQ_EGEOM: return "mcaq: geometry or alignment outside the vector path (no scalar form of this entry point)";
    default: return "mcaq: unknown error";
  }
}

// largest power of two <= max(4, H / grid_size)   (morphology.py:359-376)
extern "C" int mcaq_tile_size(int H, int grid_size) {
  if (H <= 0 || grid_size <= 0) return MCAQ_EINVAL;
  int raw = H / grid_size;
  if (raw < 4) raw = 4;
  int t = 1;
  while ((t << 1) <= raw) t <<= 1;
  return t;
}

// ---- self test: Markstein division vs div.rn over a sweep of numerators for each scale -------
// For every scale s[j] and every numerator bit pattern x = first + i*stride (i < count), counts the
// cases where div_markstein(x, s, RN(1/s)) != x / s bitwise (NaN == NaN).
__global__ void div_selftest_kernel(const float* scales, int nscales, unsigned first, unsigned stride,
                                    unsigned long long count, unsigned long long* mismatches) {
  const int j = blockIdx.y;
  const float s = scales[j];
  const float rinv = __frcp_rn(s);
  unsigned long long bad = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float(first + (unsigned)(i * stride));
    if (!(fabsf(x) < 1.2379400392853803e27f)) continue;       // contract: finite, |x| < 2^90
    bool neq = false;
    if (fabsf(x) >= 7.888609052210118e-31f) {                 // |x| >= 2^-100: the quotient itself
      const float a = mcaq::div_markstein(x, s, rinv);
      const float b = __fdiv_rn(x, s);
      neq = __float_as_uint(a) != __float_as_uint(b);
    }
    // the integer code for every x (tiny quotients cannot move rint(q + zp)), several zero points
    const float zps[5] = {0.f, -0.5f, 3.25f, -128.f, 126.75f};
#pragma unroll
    for (int z = 0; z < 5; ++z)
      neq |= mcaq::quant_code_fast(x, s, zps[z], rinv, -128.f, 127.f) != mcaq::quant_code(x, s, zps[z], -128.f, 127.f);
    if (neq) {
      ++bad;
      mismatches[1] = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(s);   // one example
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}

extern "C" MCAQ_API int mcaq_selftest_division(const float* scales, int nscales, unsigned first, unsigned stride,
                                               unsigned long long count, unsigned long long* mismatches,
                                               void* stream) {
  if (!scales || !mismatches || nscales <= 0) return MCAQ_EINVAL;
  dim3 grid(1184, nscales);
  div_selftest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scales, nscales, first, stride, count, mismatches);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
