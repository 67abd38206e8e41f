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
    case MCAQ_EDTYPE: return "mcaq: unsupported dtype (expected MCAQ_F32 or MCAQ_BF16)";
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
