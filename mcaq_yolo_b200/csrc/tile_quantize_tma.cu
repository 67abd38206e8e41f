// K3, bulk-copy staged variant: the same per-element arithmetic as tile_quantize_vec_kernel (tile_quantize.cu), but
// the input tile travels global -> shared memory with cp.async.bulk (the TMA engine, 1-D bulk copies completing on an
// mbarrier) instead of LDG.128 into registers.  Bytes in flight then cost shared memory, not registers: a CTA keeps
// TQ_STAGES tiles of 32 channels x 512 (16-bit) / 256 (fp32) pixels = 32 KB each in flight whatever its register
// budget, and the consumer warps only ever wait on data that has landed.
//
//   item      = (image, block of TQ_PIX pixel vectors, chunk of 32 channels): 32 rows of up to 1 KB
//   producer  = one thread: arrive.expect_tx on the stage's mbarrier + 32 bulk copies (one per channel row)
//   consumers = all 256 threads: thread = (pixel vector, group of 8 channels); LDS.128 -> quantise -> STG.128.cs
//   ring      = TQ_STAGES slots; a slot is refilled after the CTA-wide barrier that ends its consumption
//
// Persistent CTAs stride over the items.  Geometry: the vector path's (16-byte aligned rows, H*W % VEC == 0, aligned
// 4-pixel segments inside one tile) plus C % 32 == 0; everything else stays on tile_quantize_vec_kernel.
#include "common.cuh"
#include "tile_quantize.cuh"

namespace mcaq {

constexpr int TQ_THREADS = 256;
#ifndef MCAQ_TQ_CH
#define MCAQ_TQ_CH 32
#endif
#ifndef MCAQ_TQ_STAGES
#define MCAQ_TQ_STAGES 3
#endif
#ifndef MCAQ_TQ_MINB
#define MCAQ_TQ_MINB 2
#endif
constexpr int TQ_CH = MCAQ_TQ_CH;    // channels per item (32 or 16)
constexpr int TQ_CPT = TQ_CH / 4;    // channels per thread (four channel groups x 64 pixel vectors = 256 threads)
constexpr int TQ_VECS = 64;          // pixel vectors per item (512 bf16 / 256 fp32 pixels: 1 KB rows)
constexpr int TQ_STAGES = MCAQ_TQ_STAGES;
constexpr int TQ_ROWB = TQ_VECS * 16;                       // bytes per channel row of a stage
constexpr int TQ_STAGEB = TQ_CH * TQ_ROWB;                  // 32 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TmaRanges { const float* packed; const float* rmin; const float* rmax; };

template <typename T, int VEC, bool HAS_MASK>
__global__ void __launch_bounds__(TQ_THREADS, MCAQ_TQ_MINB)
tile_quantize_tma_kernel(const T* __restrict__ x, T* __restrict__ y, QGeom g, const float* __restrict__ bit_map,
                         const float* __restrict__ mask, TmaRanges rg, int blocks_per_image, int chunks, long long nitems) {
  constexpr int NSEG = VEC / 4;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* stage = smem;                                              // [TQ_STAGES][TQ_CH][TQ_ROWB]
  float4* tab = reinterpret_cast<float4*>(smem + TQ_STAGES * TQ_STAGEB);    // [chunks][7][TQ_CH + 1]  {scale, zp, 1/scale, -}
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + chunks * 7 * (TQ_CH + 1));
  const int tid = threadIdx.x;
  // quantiser rows of ALL channel chunks (C <= 512: at most 16 chunks x 7 x 33 x 16 B = 59 KB; typical 64..256: 7..30 KB)
  for (int i = tid; i < chunks * 7 * TQ_CH; i += TQ_THREADS) {
    const int ck = i / (7 * TQ_CH), r = i - ck * 7 * TQ_CH, bi = r / TQ_CH, cl = r - bi * TQ_CH;
    const int c = ck * TQ_CH + cl;
    const float mn = rg.packed ? __ldg(rg.packed + c) : __ldg(rg.rmin + c);
    const float mx = rg.packed ? -__ldg(rg.packed + g.C + c) : __ldg(rg.rmax + c);
    float qmin, qmax;
    bit_limits(bi, qmin, qmax);
    const float rng = fmaxf(__fsub_rn(mx, mn), 1e-8f);
    const float scale = __fdiv_rn(rng, __fsub_rn(qmax, qmin));
    const float zp = fminf(fmaxf(__fsub_rn(qmin, __fdiv_rn(mn, scale)), qmin), qmax);
    tab[(ck * 7 + bi) * (TQ_CH + 1) + cl] = make_float4(scale, zp, __frcp_rn(scale), 0.f);
  }
  if (tid == 0) {
    for (int s = 0; s < TQ_STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long sb = (long long)g.HW * (long long)sizeof(T);              // channel stride in bytes
  // item -> (image b, channel chunk ck, pixel block pb); chunk fastest so that neighbouring CTAs share bit maps / masks
  auto issue = [&](long long item, int slot) {
    const int ck = (int)(item % chunks);
    const long long t = item / chunks;
    const int pb = (int)(t % blocks_per_image), b = (int)(t / blocks_per_image);
    const int v0 = pb * TQ_VECS, nv = min(TQ_VECS, g.nvec - v0);
    const uint32_t rowb = (uint32_t)nv * 16u;
    const char* src = reinterpret_cast<const char*>(x) + ((long long)b * g.C + (long long)ck * TQ_CH) * sb + (long long)v0 * 16;
    unsigned char* dst = stage + slot * TQ_STAGEB;
    mbar_expect_tx(bars + slot, rowb * TQ_CH);
#pragma unroll 4
    for (int c = 0; c < TQ_CH; ++c) bulk_g2s(dst + c * TQ_ROWB, src + (long long)c * sb, rowb, bars + slot);
  };
  const long long first = blockIdx.x, stride = gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < TQ_STAGES; ++s)
      if (first + s * stride < nitems) issue(first + s * stride, s);
  }
  const int v = tid & (TQ_VECS - 1), cg8 = tid >> 6;                         // pixel vector, group of 8 channels
  int it = 0;
  for (long long item = first; item < nitems; item += stride, ++it) {
    const int slot = it % TQ_STAGES;
    const uint32_t parity = (uint32_t)((it / TQ_STAGES) & 1);
    const int ck = (int)(item % chunks);
    const long long t = item / chunks;
    const int pb = (int)(t % blocks_per_image), b = (int)(t / blocks_per_image);
    const int v0 = pb * TQ_VECS, nv = min(TQ_VECS, g.nvec - v0);
    // per-vector context while the tile is in flight
    int trow[NSEG];
    float qmin[NSEG], qmax[NSEG], m[VEC];
    const int pix = (v0 + v) * VEC;
    if (v < nv) {
      const int h0 = pix / g.W, w0 = pix - h0 * g.W;
#pragma unroll
      for (int s = 0; s < NSEG; ++s) {
        int hs = h0, ws = w0 + 4 * s;
        if (ws >= g.W) { ws -= g.W; hs += 1; }
        const int ty = nearest_src(hs, g.sy, g.Ht), tx = nearest_src(ws, g.sx, g.Wt);
        float bf = rintf(__ldg(bit_map + ((long long)b * g.Ht + ty) * g.Wt + tx));
        bf = fminf(fmaxf(bf, 2.f), 8.f);
        const int bidx = (int)bf - 2;
        trow[s] = (ck * 7 + bidx) * (TQ_CH + 1) + cg8 * TQ_CPT;
        bit_limits(bidx, qmin[s], qmax[s]);
        if (HAS_MASK) {
          const float4 mv = __ldg(reinterpret_cast<const float4*>(mask + (long long)b * g.HW + pix) + s);
          m[4 * s + 0] = mv.x; m[4 * s + 1] = mv.y; m[4 * s + 2] = mv.z; m[4 * s + 3] = mv.w;
        }
      }
    }
    mbar_wait(bars + slot, parity);
    if (v < nv) {
      const unsigned char* src = stage + slot * TQ_STAGEB + (cg8 * TQ_CPT) * TQ_ROWB + v * 16;
      char* yb = reinterpret_cast<char*>(y) + ((long long)b * g.C + (long long)ck * TQ_CH + cg8 * TQ_CPT) * sb + (long long)(v0 + v) * 16;
#pragma unroll
      for (int c = 0; c < TQ_CPT; ++c) {
        const uint4 rawv = *reinterpret_cast<const uint4*>(src + c * TQ_ROWB);
        float xv[VEC], out[VEC];
        Elem<T>::unpack(rawv, xv);
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
          const float4 p = tab[trow[s] + c];
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const int i = 4 * s + e;
            const float2 q = quant_code_fast2(make_float2(xv[i], xv[i + 1]), p.x, p.y, p.z, qmin[s], qmax[s]);
            float2 d = dequant2(q, p.x, p.y);
            if (HAS_MASK) d = fmul2(d, make_float2(m[i], m[i + 1]));
            out[i] = d.x;
            out[i + 1] = d.y;
          }
        }
        stg_stream(yb + (long long)c * sb, Elem<T>::pack(out));
      }
    }
    __syncthreads();                                                          // slot consumed by every thread
    if (tid == 0) {
      const long long nxt = item + (long long)TQ_STAGES * stride;
      if (nxt < nitems) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy reads before the async-proxy refill
        issue(nxt, slot);
      }
    }
  }
}

template <typename T, int VEC>
static int launch_tma(const T* x, T* y, int B, int C, int H, int W, const float* bit_map, int Ht, int Wt, const float* mask,
                      TmaRanges rg, cudaStream_t st) {
  QGeom g = make_geom(B, C, H, W, Ht, Wt, VEC);
  const int chunks = C / TQ_CH;
  const int bpi = (g.nvec + TQ_VECS - 1) / TQ_VECS;
  const long long nitems = (long long)B * bpi * chunks;
  const size_t smem = (size_t)TQ_STAGES * TQ_STAGEB + (size_t)chunks * 7 * (TQ_CH + 1) * sizeof(float4) + TQ_STAGES * sizeof(uint64_t);
  if (smem > 227 * 1024) return MCAQ_ETOOBIG;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  long long grid = (long long)MCAQ_TQ_MINB * sms;
  if (grid > nitems) grid = nitems;
  auto k = mask ? tile_quantize_tma_kernel<T, VEC, true> : tile_quantize_tma_kernel<T, VEC, false>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<(unsigned)grid, TQ_THREADS, smem, st>>>(x, y, g, bit_map, mask, rg, bpi, chunks, nitems);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace mcaq

using namespace mcaq;

// Same contract as mcaq_tile_quantize_ranges for the geometries this variant covers (MCAQ_EGEOM otherwise): bulk-copy
// (TMA) staged input.  y must not alias x (the copy engine reads x while earlier items are being stored).
extern "C" int mcaq_tile_quantize_ranges_tma(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                             const float* bit_map, int Ht, int Wt, const float* packed,
                                             const float* running_min, const float* running_max, const float* mask,
                                             void* stream) {
  if (!x || !y || !bit_map || (!packed && (!running_min || !running_max))) return MCAQ_EINVAL;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0 || (long long)H * W > 0x7fffffffLL) return MCAQ_EINVAL;
  if (x == y) return MCAQ_EINVAL;
  const int VEC = dtype == MCAQ_F32 ? 4 : 8;
  if (C % TQ_CH != 0 || !seg_ok(x, y, mask, nullptr, H * W, W, Wt, VEC)) return MCAQ_EGEOM;
  TmaRanges rg{packed, running_min, running_max};
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MCAQ_F32) return launch_tma<float, 4>((const float*)x, (float*)y, B, C, H, W, bit_map, Ht, Wt, mask, rg, st);
  if (dtype == MCAQ_BF16) {
    typedef __nv_bfloat16 bf;
    return launch_tma<bf, 8>((const bf*)x, (bf*)y, B, C, H, W, bit_map, Ht, Wt, mask, rg, st);
  }
  if (dtype == MCAQ_F16) return launch_tma<__half, 8>((const __half*)x, (__half*)y, B, C, H, W, bit_map, Ht, Wt, mask, rg, st);
  return MCAQ_EDTYPE;
}
