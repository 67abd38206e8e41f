// Training forms of the three tile-level networks (forward with saved statistics, backward):
//   * complexity MLP 8 -> 64 (LN, ReLU) -> 32 (LN, ReLU) -> 1 (sigmoid)  + 5x5 bilateral filter + clamp
//     (core/morphology.py:81-97, 309-354, 962-968): backward only -- the forward IS the inference kernel
//   * complexity-to-bit mapper 3 -> 32 -> 64 -> 32 -> 1 with TRAIN-MODE BatchNorm1d (batch statistics over all
//     B*ht*wt tiles, running statistics updated; core/bit_allocation.py:126, 218-280): forward and backward
//   * learned soft mask (core/quantization.py:213-239): backward -- the forward IS the inference kernel
// and the two scalar reductions the loss needs from a bit map (models/mcaq_yolo.py:86-118, 575).
//
// The problems are tiny (a few thousand rows of 8..64 features; 2.9k / 4.6k / 170 parameters): what costs in
// the reference is ~10^3 eager autograd kernels per step.  Here a network is one or two launches per direction.
// Rows are independent except for the mapper's BatchNorm: its kernels run as ONE thread-block cluster (16 CTAs where the device co-schedules them, else 8,
// rows split over the CTAs) that reduces the per-feature statistics through distributed shared memory and, when
// the batch is sharded over the GPUs of a node, merges them with the other ranks over NVLink peer memory inside
// the same kernel (peer_exchange.cuh protocol: rank-ordered Chan merge of (count, mean, M2), so every rank holds
// bit-identical statistics = the unsharded batch's).  Parameter gradients: every CTA sums
// G^T A over its rows (thread per parameter) and adds the partial into the global block (one atomic per CTA).
//
// Parameter / gradient block layouts = the flat concatenation of the torch parameters in module order:
//   complexity MLP (2881): W0[64][8] b0[64] g1[64] be1[64] W3[32][64] b3[32] g4[32] be4[32] W6[32] b6
//   mapper (4609): W0[32][3] b0[32] g0[32] be0[32] W3[64][32] b3[64] g3[64] be3[64] W6[32][64] b6[32] g6[32] be6[32]
//                  W9[32] b9
//   soft mask (170): W0[8][2][3][3] b0[8] W2[2][8] b2[2]
#include <cooperative_groups.h>

#include "peer_exchange.cuh"
#include "tile_nets.cuh"

namespace cg = cooperative_groups;

namespace mcaq {

// ---- warp helpers: a warp carries one row (tile) at a time, lane = unit ------------------------------------------
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// stage a row-major [OUT][IN] matrix into shared memory with row stride IN + 1 (conflict-free by row AND by column)
__device__ __forceinline__ void stage_matrix(const float* __restrict__ Wg, float* Ws, int OUT, int IN) {
  for (int i = threadIdx.x; i < OUT * IN; i += blockDim.x) Ws[(i / IN) * (IN + 1) + (i % IN)] = __ldg(Wg + i);
}
// add a CTA's per-warp register accumulators into a global gradient block: smem staging (one atomic per warp and
// value into shared memory, then one global atomic per CTA and value)
__device__ __forceinline__ void flush_begin(float* buf, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = 0.f;
  __syncthreads();
}
__device__ __forceinline__ void flush_end(const float* buf, int n, float* __restrict__ g) {
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (buf[i] != 0.f) atomicAdd(g + i, buf[i]);
  __syncthreads();
}

// =================================================================================================
// complexity MLP + bilateral + clamp: backward
// =================================================================================================
// (1) bilateral + clamp: one CTA per image, g_craw[t] accumulated in shared memory
__global__ void __launch_bounds__(256)
bilateral_bwd_kernel(const float* __restrict__ craw, const float* __restrict__ gout, int ht, int wt, float* __restrict__ gcraw) {
  extern __shared__ float acc[];                              // [ht*wt]
  const int b = blockIdx.x, nt = ht * wt;
  const float* c = craw + (long long)b * nt;
  const float* go = gout + (long long)b * nt;
  for (int i = threadIdx.x; i < nt; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const int y = t / wt, x = t - y * wt;
    const float vt = c[t];
    float w[25], v[25];
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int k = 0; k < 25; ++k) {
      const int yy = min(max(y + k / 5 - 2, 0), ht - 1), xx = min(max(x + k % 5 - 2, 0), wt - 1);
      v[k] = c[yy * wt + xx];
      const float d = __fsub_rn(v[k], vt);
      w[k] = __fmul_rn(kc::BILAT[k], exp_f64(__fdiv_rn(-__fmul_rn(d, d), kc::BILAT_DEN)));
      num = fmaf(w[k], v[k], num);
      den = __fadd_rn(den, w[k]);
    }
    const float dinv = __fdiv_rn(1.0f, __fadd_rn(den, 1e-8f));
    const float r = __fmul_rn(num, dinv);
    const float G = (r >= 0.f && r <= 1.f) ? go[t] : 0.f;       // clamp(0, 1) passes the gradient inside the range
    if (G == 0.f) continue;
    float gt = 0.f;
#pragma unroll
    for (int k = 0; k < 25; ++k) {
      const int yy = min(max(y + k / 5 - 2, 0), ht - 1), xx = min(max(x + k % 5 - 2, 0), wt - 1);
      const float d = __fsub_rn(v[k], vt);
      // dr/dw_k = (v_k - r) / (den + eps);  dw_k/dv_k = -w_k 2 d / D,  dw_k/dv_t = +w_k 2 d / D
      const float e = __fmul_rn(__fmul_rn(__fsub_rn(v[k], r), dinv), __fmul_rn(w[k], __fdiv_rn(__fmul_rn(2.f, d), kc::BILAT_DEN)));
      atomicAdd(acc + yy * wt + xx, __fmul_rn(G, __fsub_rn(__fmul_rn(w[k], dinv), e)));
      gt = __fadd_rn(gt, e);
    }
    atomicAdd(acc + t, __fmul_rn(G, gt));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nt; i += blockDim.x) gcraw[(long long)b * nt + i] = acc[i];
}

// (2) MLP: a warp carries a row through the recomputed forward and the backward; lane = unit (two units per lane in
//     the 64-wide layer), LayerNorm statistics by warp shuffles, the 64 x 32 matrix in shared memory (stride 65),
//     parameter gradients in per-lane registers over the warp's rows, flushed once per CTA.
constexpr int CM_NT = 128;
__global__ void __launch_bounds__(CM_NT)
cmlp_bwd_kernel(const float* __restrict__ phi, const float* __restrict__ gcraw, const float* __restrict__ P, int N,
                float* __restrict__ gP) {
  __shared__ float W3s[32 * 65];
  __shared__ float rowbuf[CM_NT / 32][64];
  __shared__ float fb[2881];
  const float* W0 = P; const float* b0 = P + 512; const float* g1 = P + 576; const float* be1 = P + 640;
  const float* W3 = P + 704; const float* b3 = P + 2752; const float* g4 = P + 2784; const float* be4 = P + 2816;
  const float* W6 = P + 2848; const float* b6 = P + 2880;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  stage_matrix(W3, W3s, 32, 64);
  flush_begin(fb, 2881);
  // per-lane constants: units u0 = lane, u1 = lane + 32 of layer 1; unit lane of layer 2
  float w0a[8], w0b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { w0a[i] = __ldg(W0 + lane * 8 + i); w0b[i] = __ldg(W0 + (lane + 32) * 8 + i); }
  const float b0a = __ldg(b0 + lane), b0b = __ldg(b0 + lane + 32);
  const float g1a = __ldg(g1 + lane), g1b = __ldg(g1 + lane + 32), e1a = __ldg(be1 + lane), e1b = __ldg(be1 + lane + 32);
  const float b3l = __ldg(b3 + lane), g4l = __ldg(g4 + lane), e4l = __ldg(be4 + lane), w6l = __ldg(W6 + lane), b6v = __ldg(b6);
  // gradient accumulators
  float dW0a[8], dW0b[8], dW3r[64];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dW0a[i] = 0.f; dW0b[i] = 0.f; }
#pragma unroll
  for (int k = 0; k < 64; ++k) dW3r[k] = 0.f;
  float db0a = 0.f, db0b = 0.f, dg1a = 0.f, dg1b = 0.f, de1a = 0.f, de1b = 0.f, db3 = 0.f, dg4 = 0.f, de4 = 0.f, dW6 = 0.f, db6 = 0.f;
  float* rb = rowbuf[warp];
  const int nw = gridDim.x * (CM_NT / 32);
  for (int r = blockIdx.x * (CM_NT / 32) + warp; r < N; r += nw) {
    float x[8];
    {
      const float4 pa = __ldg(reinterpret_cast<const float4*>(phi + (long long)r * 8));
      const float4 pb = __ldg(reinterpret_cast<const float4*>(phi + (long long)r * 8) + 1);
      x[0] = pa.x; x[1] = pa.y; x[2] = pa.z; x[3] = pa.w; x[4] = pb.x; x[5] = pb.y; x[6] = pb.z; x[7] = pb.w;
    }
    float za = 0.f, zb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { za = fmaf(w0a[i], x[i], za); zb = fmaf(w0b[i], x[i], zb); }
    za = __fadd_rn(za, b0a); zb = __fadd_rn(zb, b0b);
    // LayerNorm(64) + ReLU
    const float m1 = __fmul_rn(wsum(__fadd_rn(za, zb)), 1.0f / 64);
    const float da = __fsub_rn(za, m1), dbv = __fsub_rn(zb, m1);
    const float v1 = __fmul_rn(wsum(fmaf(da, da, __fmul_rn(dbv, dbv))), 1.0f / 64);
    const float rs1 = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(v1, 1e-5f)));
    const float xha = __fmul_rn(da, rs1), xhb = __fmul_rn(dbv, rs1);
    const float h1a = fmaxf(fmaf(xha, g1a, e1a), 0.f), h1b = fmaxf(fmaf(xhb, g1b, e1b), 0.f);
    __syncwarp();
    rb[lane] = h1a; rb[lane + 32] = h1b;
    __syncwarp();
    // layer 2: unit = lane
    float z2 = 0.f;
#pragma unroll 16
    for (int k = 0; k < 64; ++k) z2 = fmaf(W3s[lane * 65 + k], rb[k], z2);
    z2 = __fadd_rn(z2, b3l);
    const float m2 = __fmul_rn(wsum(z2), 1.0f / 32);
    const float d2 = __fsub_rn(z2, m2);
    const float v2 = __fmul_rn(wsum(__fmul_rn(d2, d2)), 1.0f / 32);
    const float rs2 = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(v2, 1e-5f)));
    const float xh2 = __fmul_rn(d2, rs2);
    const float h2 = fmaxf(fmaf(xh2, g4l, e4l), 0.f);
    const float z3 = __fadd_rn(wsum(__fmul_rn(w6l, h2)), b6v);
    const float c = sigmoid_exact(z3);
    const float gz3 = __fmul_rn(__ldg(gcraw + r), __fmul_rn(c, __fsub_rn(1.0f, c)));
    dW6 = fmaf(gz3, h2, dW6);
    db6 = __fadd_rn(db6, gz3);
    // LN(32) backward
    const float gy2 = h2 > 0.f ? __fmul_rn(gz3, w6l) : 0.f;
    dg4 = fmaf(gy2, xh2, dg4);
    de4 = __fadd_rn(de4, gy2);
    const float gx2 = __fmul_rn(gy2, g4l);
    const float t1 = __fmul_rn(wsum(gx2), 1.0f / 32), t2 = __fmul_rn(wsum(__fmul_rn(gx2, xh2)), 1.0f / 32);
    const float gz2 = __fmul_rn(rs2, __fsub_rn(__fsub_rn(gx2, t1), __fmul_rn(xh2, t2)));
    db3 = __fadd_rn(db3, gz2);
#pragma unroll
    for (int k = 0; k < 64; ++k) dW3r[k] = fmaf(gz2, rb[k], dW3r[k]);
    // data gradient of layer 2: gh1[k] = sum_j W3[j][k] gz2[j], k = lane and lane + 32
    float gha = 0.f, ghb = 0.f;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const float gj = __shfl_sync(0xffffffffu, gz2, j);
      gha = fmaf(W3s[j * 65 + lane], gj, gha);
      ghb = fmaf(W3s[j * 65 + lane + 32], gj, ghb);
    }
    const float gya = h1a > 0.f ? gha : 0.f, gyb = h1b > 0.f ? ghb : 0.f;
    dg1a = fmaf(gya, xha, dg1a); dg1b = fmaf(gyb, xhb, dg1b);
    de1a = __fadd_rn(de1a, gya); de1b = __fadd_rn(de1b, gyb);
    const float gxa = __fmul_rn(gya, g1a), gxb = __fmul_rn(gyb, g1b);
    const float u1 = __fmul_rn(wsum(__fadd_rn(gxa, gxb)), 1.0f / 64);
    const float u2 = __fmul_rn(wsum(fmaf(gxa, xha, __fmul_rn(gxb, xhb))), 1.0f / 64);
    const float gza = __fmul_rn(rs1, __fsub_rn(__fsub_rn(gxa, u1), __fmul_rn(xha, u2)));
    const float gzb = __fmul_rn(rs1, __fsub_rn(__fsub_rn(gxb, u1), __fmul_rn(xhb, u2)));
    db0a = __fadd_rn(db0a, gza); db0b = __fadd_rn(db0b, gzb);
#pragma unroll
    for (int i = 0; i < 8; ++i) { dW0a[i] = fmaf(gza, x[i], dW0a[i]); dW0b[i] = fmaf(gzb, x[i], dW0b[i]); }
  }
  // flush: per-warp accumulators -> shared block -> global
#pragma unroll
  for (int i = 0; i < 8; ++i) { atomicAdd(fb + lane * 8 + i, dW0a[i]); atomicAdd(fb + (lane + 32) * 8 + i, dW0b[i]); }
  atomicAdd(fb + 512 + lane, db0a); atomicAdd(fb + 512 + lane + 32, db0b);
  atomicAdd(fb + 576 + lane, dg1a); atomicAdd(fb + 576 + lane + 32, dg1b);
  atomicAdd(fb + 640 + lane, de1a); atomicAdd(fb + 640 + lane + 32, de1b);
#pragma unroll
  for (int k = 0; k < 64; ++k) atomicAdd(fb + 704 + k * 32 + lane, dW3r[k]);      // [k][unit] while in shared memory
  atomicAdd(fb + 2752 + lane, db3); atomicAdd(fb + 2784 + lane, dg4); atomicAdd(fb + 2816 + lane, de4);
  atomicAdd(fb + 2848 + lane, dW6);
  if (lane == 0) atomicAdd(fb + 2880, db6);
  __syncthreads();
  for (int i = threadIdx.x; i < 2881; i += blockDim.x) {
    const float v = fb[i];
    if (v == 0.f) continue;
    const int q = i - 704;                                     // the 32 x 64 block is stored [k][unit]
    atomicAdd(gP + ((q >= 0 && q < 2048) ? 704 + (q & 31) * 64 + (q >> 5) : i), v);
  }
}

// =================================================================================================
// mapper with train-mode BatchNorm: one cluster of MAP_CL CTAs, rows split over the CTAs, a warp per row
// =================================================================================================
constexpr int MAP_CL = 16;                                     // largest cluster (non-portable size, opted into); 8 is the fallback
constexpr int MAP_NT = 512;
// scratch per row: Z1[32] Z2[64] Z3[32] (forward, kept for backward) | GY3[32] (later GY1) GY2[64] (backward)
constexpr int MP_SCR = 256;
constexpr int GATH = 132;                                      // gather row: up to 2 * 64 + 1 values per CTA
struct MapArgs {
  const float* c; const float* P; int N; int rpc;
  float temperature; int use_t; float lo, hi;
  float* scratch; float* stats;
  float* rm[3]; float* rv[3]; float momentum; float eps;      // running statistics (NULL: not tracked)
  long long* nbt[3];                                           // num_batches_tracked of the three BatchNorm layers (NULL: skip)
  float* out;
  const float* gout; float* gc; float* gP;                     // backward
  XchgPeers px;                                                // world > 1: statistics merged over the ranks
};
// shared memory of both kernels (floats): work[392] | gather[MAP_CL][GATH] | mean[64] var[64] | W3s[64][33] | W6s[32][65]
// | rowbuf[MAP_NT / 32][64] | fb[2112]
constexpr int MAP_SM_WORK = 0, MAP_SM_GATH = 392, MAP_SM_MEAN = MAP_SM_GATH + MAP_CL * GATH, MAP_SM_VAR = MAP_SM_MEAN + 64;
constexpr int MAP_SM_W3 = MAP_SM_VAR + 64, MAP_SM_W6 = MAP_SM_W3 + 64 * 33, MAP_SM_ROW = MAP_SM_W6 + 32 * 65;
constexpr int MAP_SM_FB = MAP_SM_ROW + (MAP_NT / 32) * 64, MAP_SM_FLOATS = MAP_SM_FB + 2112;
// Row cache: when a CTA's share of the rows fits, the per-row record lives in shared memory for the whole kernel
// (the statistics passes and the next layer re-read it at LDS latency instead of L2's); the forward still stores the
// pre-activations to the global scratch for the backward launch, the backward loads them once and keeps its GY slots
// on chip.  GY1 reuses GY3's slot (dead after the layer-3 phase), so a cached row is 224 floats.
constexpr int MAP_ROWF = 224;
constexpr int MAP_GY1 = 128;                                   // offset of GY1 (aliases GY3) in both layouts
constexpr int MAP_ROWS_MAX = (227 * 1024 / 4 - MAP_SM_FLOATS) / MAP_ROWF;      // rows per CTA that fit (218)

// exchange of a short vector with the other GPU ranks by cluster CTA 0 (peer_exchange.cuh protocol, bounded wait):
// returns in xm[0..n) the per-rank vectors' rank-ordered combination computed by `merge`.  A timed-out exchange
// degrades to this rank's own values (error word set; peer.RangeExchange.check raises).
template <typename Merge>
__device__ __forceinline__ void rank_exchange(const XchgPeers& px, const float* mine, int n, float* xm, Merge merge) {
  __shared__ int step_s;
  const int tid = threadIdx.x;
  float* local = px.base[px.rank];
  if (tid == 0) { int* ep = reinterpret_cast<int*>(local); step_s = *ep + 1; *ep = step_s; }
  __syncthreads();
  const int e = step_s;
  const int slotf = 2 * 128;                                   // slot capacity (floats) of a C = 128 exchange buffer
  for (int p = 0; p < px.world; ++p) {
    float* dst = px.base[p] + XCHG_SLOTS + (long long)((e & 1) * px.world + px.rank) * slotf;
    if (tid < n) dst[tid] = mine[tid];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < px.world) st_release_sys(reinterpret_cast<int*>(px.base[tid]) + XCHG_FLAGS + 8 * (e & 1) + px.rank, e);
  const bool ok = xchg_wait(local, px.world, e, tid, px.timeout_ns);
  const bool all_ok = __syncthreads_and(ok);
  const float* slots = local + XCHG_SLOTS + (long long)((e & 1) * px.world) * slotf;
  merge(slots, slotf, all_ok, xm);
  __syncthreads();
}

// Cluster-wide per-feature statistics of Zl[r][0..F) over ALL rows of the cluster (and of all ranks):
// two-pass mean / M2 inside a CTA (fixed order), Chan merge over the CTAs in rank order through DSMEM, then over
// the GPU ranks through peer memory.  Result in sm_mean / sm_var (biased), total count returned.
__device__ float batch_stats(cg::cluster_group& cl, const float* Zl, int ld, int r0, int r1, int F, float* sm,
                             float* gather, const XchgPeers& px, float* sm_mean, float* sm_var) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int SL = min(NT / F, 256 / F);                         // row slices per feature (part[] is 256 floats)
  float* part = sm;                                            // [SL][F]
  const int f = tid % F, sl = tid / F;
  const int n = r1 - r0;
  float a = 0.f;
  if (sl < SL) {
    for (int r = r0 + sl; r < r1; r += SL) a = __fadd_rn(a, Zl[(long long)r * ld + f]);
    part[sl * F + f] = a;
  }
  __syncthreads();
  if (tid < F) {
    float mean = 0.f;
    for (int s_ = 0; s_ < SL; ++s_) mean = __fadd_rn(mean, part[s_ * F + tid]);
    sm_mean[tid] = n > 0 ? __fdiv_rn(mean, (float)n) : 0.f;
  }
  __syncthreads();
  a = 0.f;
  if (sl < SL) {
    const float m = sm_mean[f];
    for (int r = r0 + sl; r < r1; r += SL) { const float d = __fsub_rn(Zl[(long long)r * ld + f], m); a = fmaf(d, d, a); }
    part[sl * F + f] = a;
  }
  __syncthreads();
  const int rank = (int)cl.block_rank(), ncta = (int)cl.num_blocks();
  if (tid < F) {
    float m2 = 0.f;
    for (int s_ = 0; s_ < SL; ++s_) m2 = __fadd_rn(m2, part[s_ * F + tid]);
    for (int p = 0; p < ncta; ++p) {                           // all-gather (count, mean, M2) into every CTA
      float* g = cl.map_shared_rank(gather, p) + rank * GATH;
      g[tid] = sm_mean[tid];
      g[F + tid] = m2;
      if (tid == 0) g[2 * F] = (float)n;
    }
  }
  cl.sync();
  // Chan merge in CTA order: identical on every CTA
  float cnt = 0.f, mu = 0.f, M2 = 0.f;
  if (tid < F) {
    for (int p = 0; p < ncta; ++p) {
      const float* g = gather + p * GATH;
      const float nb = g[2 * F];
      if (nb <= 0.f) continue;
      const float d = __fsub_rn(g[tid], mu), tot = __fadd_rn(cnt, nb);
      mu = __fadd_rn(mu, __fmul_rn(d, __fdiv_rn(nb, tot)));
      M2 = __fadd_rn(__fadd_rn(M2, g[F + tid]), __fmul_rn(__fmul_rn(d, d), __fdiv_rn(__fmul_rn(cnt, nb), tot)));
      cnt = tot;
    }
  }
  float* xv = sm;                                              // [mean F | M2 F | count] of this CTA's merged values
  float* xm = sm + 192;                                        // CTA 0: merged over the GPU ranks
  if (px.world > 1) {
    __syncthreads();
    if (tid < F) { xv[tid] = mu; xv[F + tid] = M2; if (tid == 0) xv[2 * F] = cnt; }
    __syncthreads();
    if (rank == 0)
      rank_exchange(px, xv, 2 * F + 1, xm, [&](const float* slots, int slotf, bool all_ok, float* out) {
        if (tid < F) {
          float c2 = 0.f, m_ = 0.f, q_ = 0.f;
          for (int q = 0; q < px.world; ++q) {
            const volatile float* s_ = slots + (long long)q * slotf;
            const float nb = all_ok ? s_[2 * F] : (q == px.rank ? xv[2 * F] : 0.f);
            if (nb <= 0.f) continue;
            const float mq = all_ok ? s_[tid] : xv[tid], qq = all_ok ? s_[F + tid] : xv[F + tid];
            const float d = __fsub_rn(mq, m_), tot = __fadd_rn(c2, nb);
            m_ = __fadd_rn(m_, __fmul_rn(d, __fdiv_rn(nb, tot)));
            q_ = __fadd_rn(__fadd_rn(q_, qq), __fmul_rn(__fmul_rn(d, d), __fdiv_rn(__fmul_rn(c2, nb), tot)));
            c2 = tot;
          }
          out[tid] = m_; out[F + tid] = q_;
          if (tid == 0) out[2 * F] = c2;
        }
      });
    cl.sync();
    if (tid < F) {
      const float* x0 = cl.map_shared_rank(sm + 192, 0);
      mu = x0[tid]; M2 = x0[F + tid]; cnt = x0[2 * F];
    }
    cl.sync();                                                 // CTA 0's buffer is free again
  }
  if (tid < F) { sm_mean[tid] = mu; sm_var[tid] = cnt > 0.f ? __fdiv_rn(M2, cnt) : 0.f; }
  if (tid == 0) sm[391] = cnt;
  __syncthreads();
  return sm[391];
}

// cluster-wide (and rank-wide) SUM of a CTA-local shared vector v[0..n), n <= GATH: result back in v on every CTA
__device__ void cluster_sum(cg::cluster_group& cl, float* v, int n, float* sm, float* gather, const XchgPeers& px) {
  const int tid = threadIdx.x;
  const int rank = (int)cl.block_rank(), ncta = (int)cl.num_blocks();
  __syncthreads();
  if (tid < n)
    for (int p = 0; p < ncta; ++p) cl.map_shared_rank(gather, p)[rank * GATH + tid] = v[tid];
  cl.sync();
  float tot = 0.f;
  if (tid < n)
    for (int p = 0; p < ncta; ++p) tot = __fadd_rn(tot, gather[p * GATH + tid]);
  if (px.world > 1) {
    float* xv = sm;
    float* xm = sm + 192;
    if (tid < n) xv[tid] = tot;
    __syncthreads();
    if (rank == 0)
      rank_exchange(px, xv, n, xm, [&](const float* slots, int slotf, bool all_ok, float* out) {
        if (tid < n) {
          float t = 0.f;
          for (int q = 0; q < px.world; ++q) {
            const volatile float* s_ = slots + (long long)q * slotf;
            t = __fadd_rn(t, all_ok ? s_[tid] : (q == px.rank ? xv[tid] : 0.f));
          }
          out[tid] = t;
        }
      });
    cl.sync();
    if (tid < n) tot = cl.map_shared_rank(sm + 192, 0)[tid];
    cl.sync();
  }
  if (tid < n) v[tid] = tot;
  __syncthreads();
}

__device__ __forceinline__ float bn_relu1(float z, float mean, float rstd, float g, float b) {
  return fmaxf(fmaf(__fmul_rn(__fsub_rn(z, mean), rstd), g, b), 0.f);
}
__device__ __forceinline__ float rstd_of(float var, float eps) { return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps))); }

__device__ __forceinline__ void update_running(float* rm, float* rv, const float* mean, const float* var, float n, int F,
                                               float momentum) {
  if (!rm || (int)threadIdx.x >= F) return;
  const int k = threadIdx.x;
  const float unb = n > 1.f ? __fmul_rn(var[k], __fdiv_rn(n, __fsub_rn(n, 1.f))) : var[k];
  rm[k] = __fadd_rn(__fmul_rn(__fsub_rn(1.f, momentum), rm[k]), __fmul_rn(momentum, mean[k]));
  rv[k] = __fadd_rn(__fmul_rn(__fsub_rn(1.f, momentum), rv[k]), __fmul_rn(momentum, unb));
}

__global__ void __launch_bounds__(MAP_NT) mapper_train_fwd_kernel(const MapArgs A) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cl = cg::this_cluster();
  float* gather = sm + MAP_SM_GATH;
  float* mean = sm + MAP_SM_MEAN;
  float* var = sm + MAP_SM_VAR;
  float* W3s = sm + MAP_SM_W3;
  float* W6s = sm + MAP_SM_W6;
  const float* P = A.P;
  const float* W0 = P; const float* b0 = P + 96; const float* g0 = P + 128; const float* be0 = P + 160;
  const float* W3 = P + 192; const float* b3 = P + 2240; const float* g3 = P + 2304; const float* be3 = P + 2368;
  const float* W6 = P + 2432; const float* b6 = P + 4480; const float* g6 = P + 4512; const float* be6 = P + 4544;
  const float* W9 = P + 4576; const float* b9 = P + 4608;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = MAP_NT / 32;
  float* rb = sm + MAP_SM_ROW + warp * 64;
  const int rank = (int)cl.block_rank();
  const int r0 = min(rank * A.rpc, A.N), r1 = min(r0 + A.rpc, A.N);
  const bool cached = A.rpc <= MAP_ROWS_MAX;                   // launch_mapper sized the dynamic smem accordingly
  float* rows = sm + MAP_SM_FLOATS;
  // record of row r as this kernel re-reads it, and its leading dimension
  const int ld = cached ? MAP_ROWF : MP_SCR;
  float* zb = cached ? rows - (long long)r0 * MAP_ROWF : A.scratch;
  stage_matrix(W3, W3s, 64, 32);
  stage_matrix(W6, W6s, 32, 64);
  // layer 1: unit = lane
  {
    const float w0 = __ldg(W0 + lane * 3), w1 = __ldg(W0 + lane * 3 + 1), w2 = __ldg(W0 + lane * 3 + 2), bb = __ldg(b0 + lane);
    for (int r = r0 + warp; r < r1; r += nwarps) {
      const float c = fminf(fmaxf(__ldg(A.c + r), 0.f), 1.f);
      float a = 0.f;
      a = fmaf(w0, c, a); a = fmaf(w1, __fmul_rn(c, c), a); a = fmaf(w2, log1p_f64(c), a);
      const float z = __fadd_rn(a, bb);
      A.scratch[(long long)r * MP_SCR + lane] = z;
      if (cached) zb[(long long)r * ld + lane] = z;
    }
  }
  __syncthreads();
  float ntot = batch_stats(cl, zb, ld, r0, r1, 32, sm, gather, A.px, mean, var);
  if (rank == 0) {
    update_running(A.rm[0], A.rv[0], mean, var, ntot, 32, A.momentum);
    if (threadIdx.x < 3 && A.nbt[threadIdx.x]) *A.nbt[threadIdx.x] += 1;
    if (threadIdx.x < 32) { A.stats[threadIdx.x] = mean[threadIdx.x]; A.stats[128 + threadIdx.x] = var[threadIdx.x]; }
  }
  // layer 2: units lane, lane + 32
  {
    const float mu = mean[lane], rs = rstd_of(var[lane], A.eps), gg = __ldg(g0 + lane), ee = __ldg(be0 + lane);
    const float ba = __ldg(b3 + lane), bb = __ldg(b3 + lane + 32);
    for (int r = r0 + warp; r < r1; r += nwarps) {
      float* s = zb + (long long)r * ld;
      float* sg = A.scratch + (long long)r * MP_SCR;
      __syncwarp();
      rb[lane] = bn_relu1(s[lane], mu, rs, gg, ee);
      __syncwarp();
      float za = 0.f, zc = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) { const float h = rb[k]; za = fmaf(W3s[lane * 33 + k], h, za); zc = fmaf(W3s[(lane + 32) * 33 + k], h, zc); }
      za = __fadd_rn(za, ba); zc = __fadd_rn(zc, bb);
      sg[32 + lane] = za; sg[64 + lane] = zc;
      if (cached) { s[32 + lane] = za; s[64 + lane] = zc; }
    }
  }
  __syncthreads();
  ntot = batch_stats(cl, zb + 32, ld, r0, r1, 64, sm, gather, A.px, mean, var);
  if (rank == 0) {
    update_running(A.rm[1], A.rv[1], mean, var, ntot, 64, A.momentum);
    if (threadIdx.x < 64) { A.stats[32 + threadIdx.x] = mean[threadIdx.x]; A.stats[160 + threadIdx.x] = var[threadIdx.x]; }
  }
  // layer 3: unit = lane
  {
    const float mua = mean[lane], rsa = rstd_of(var[lane], A.eps), ga = __ldg(g3 + lane), ea = __ldg(be3 + lane);
    const float mub = mean[lane + 32], rsb = rstd_of(var[lane + 32], A.eps), gb = __ldg(g3 + lane + 32), eb = __ldg(be3 + lane + 32);
    const float bb = __ldg(b6 + lane);
    for (int r = r0 + warp; r < r1; r += nwarps) {
      float* s = zb + (long long)r * ld;
      __syncwarp();
      rb[lane] = bn_relu1(s[32 + lane], mua, rsa, ga, ea);
      rb[lane + 32] = bn_relu1(s[64 + lane], mub, rsb, gb, eb);
      __syncwarp();
      float z = 0.f;
#pragma unroll 16
      for (int k = 0; k < 64; ++k) z = fmaf(W6s[lane * 65 + k], rb[k], z);
      z = __fadd_rn(z, bb);
      A.scratch[(long long)r * MP_SCR + 96 + lane] = z;
      if (cached) s[96 + lane] = z;
    }
  }
  __syncthreads();
  ntot = batch_stats(cl, zb + 96, ld, r0, r1, 32, sm, gather, A.px, mean, var);
  if (rank == 0) {
    update_running(A.rm[2], A.rv[2], mean, var, ntot, 32, A.momentum);
    if (threadIdx.x < 32) { A.stats[96 + threadIdx.x] = mean[threadIdx.x]; A.stats[224 + threadIdx.x] = var[threadIdx.x]; }
  }
  // head: 32 -> 1, sigmoid, Eq.17, temperature, straight-through clamp (bit_allocation.py:258-273)
  {
    const float mu = mean[lane], rs = rstd_of(var[lane], A.eps), gg = __ldg(g6 + lane), ee = __ldg(be6 + lane);
    const float w9 = __ldg(W9 + lane), bb = __ldg(b9);
    for (int r = r0 + warp; r < r1; r += nwarps) {
      const float h = bn_relu1(zb[(long long)r * ld + 96 + lane], mu, rs, gg, ee);
      const float sg = sigmoid_exact(__fadd_rn(wsum(__fmul_rn(w9, h)), bb));
      float bits = __fadd_rn(A.lo, __fmul_rn(__fsub_rn(A.hi, A.lo), sg));
      if (A.use_t) bits = __fmul_rn(bits, A.temperature);
      if (lane == 0) A.out[r] = fminf(fmaxf(bits, A.lo), A.hi);   // value of bits + (clamp(bits) - bits).detach()
    }
  }
  cl.sync();                                                   // nobody leaves while a peer may still read its gather area
}

__global__ void __launch_bounds__(MAP_NT) mapper_train_bwd_kernel(const MapArgs A) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cl = cg::this_cluster();
  float* gather = sm + MAP_SM_GATH;
  float* sv = sm + MAP_SM_MEAN;                                // [s1 64 | s2 64]: BatchNorm backward sums (+ count at [128]... see below)
  float* W3s = sm + MAP_SM_W3;
  float* W6s = sm + MAP_SM_W6;
  float* fb = sm + MAP_SM_FB;
  const float* P = A.P;
  const float* W0 = P; const float* g0 = P + 128; const float* be0 = P + 160;
  const float* W3 = P + 192; const float* g3 = P + 2304; const float* be3 = P + 2368;
  const float* W6 = P + 2432; const float* g6 = P + 4512; const float* be6 = P + 4544;
  const float* W9 = P + 4576; const float* b9 = P + 4608;
  const float* mean = A.stats; const float* var = A.stats + 128;
  float* gP = A.gP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = MAP_NT / 32;
  float* rb = sm + MAP_SM_ROW + warp * 64;
  const int rank = (int)cl.block_rank();
  const int r0 = min(rank * A.rpc, A.N), r1 = min(r0 + A.rpc, A.N);
  const bool cached = A.rpc <= MAP_ROWS_MAX;
  float* rows = sm + MAP_SM_FLOATS;
  const int ld = cached ? MAP_ROWF : MP_SCR;
  float* zb = cached ? rows - (long long)r0 * MAP_ROWF : A.scratch;
  stage_matrix(W3, W3s, 64, 32);
  stage_matrix(W6, W6s, 32, 64);
  if (cached) {                                                // the forward's pre-activations, once, 16 bytes per load
    const int nrow = r1 - r0;
    for (int i = threadIdx.x; i < nrow * 32; i += MAP_NT) {
      const int rr = i >> 5, q = i & 31;
      *reinterpret_cast<float4*>(rows + rr * MAP_ROWF + 4 * q) =
          __ldg(reinterpret_cast<const float4*>(A.scratch + (long long)(r0 + rr) * MP_SCR) + q);
    }
    __syncthreads();
  }
  // ---- head + BN3 sums (units = lane) ----------------------------------------------------------------------
  const float mu3 = __ldg(mean + 96 + lane), rs3 = rstd_of(__ldg(var + 96 + lane), A.eps), g6l = __ldg(g6 + lane), e6l = __ldg(be6 + lane);
  {
    const float w9 = __ldg(W9 + lane), bb = __ldg(b9);
    float dW9 = 0.f, db9 = 0.f, s1 = 0.f, s2 = 0.f;
    flush_begin(fb, 130);                                      // [dW9 32 | db9 | pad | s1 32 | s2 32 | count]
    for (int r = r0 + warp; r < r1; r += nwarps) {
      float* s = zb + (long long)r * ld;
      const float z3 = s[96 + lane];
      const float h = bn_relu1(z3, mu3, rs3, g6l, e6l);
      const float sg = sigmoid_exact(__fadd_rn(wsum(__fmul_rn(w9, h)), bb));
      float g = __fmul_rn(__ldg(A.gout + r), __fmul_rn(__fsub_rn(A.hi, A.lo), __fmul_rn(sg, __fsub_rn(1.f, sg))));
      if (A.use_t) g = __fmul_rn(g, A.temperature);
      dW9 = fmaf(g, h, dW9);
      db9 = __fadd_rn(db9, g);
      const float gy = h > 0.f ? __fmul_rn(g, w9) : 0.f;
      s[128 + lane] = gy;
      s1 = __fadd_rn(s1, gy);
      s2 = fmaf(gy, __fmul_rn(__fsub_rn(z3, mu3), rs3), s2);
    }
    atomicAdd(fb + lane, dW9);
    if (lane == 0) atomicAdd(fb + 32, db9);
    atomicAdd(fb + 34 + lane, s1);
    atomicAdd(fb + 66 + lane, s2);
    __syncthreads();
    if (threadIdx.x < 33 && fb[threadIdx.x] != 0.f) atomicAdd(gP + 4576 + threadIdx.x, fb[threadIdx.x]);      // W9, b9
    if (threadIdx.x < 32) {                                                                                      // be6, g6: local rows
      atomicAdd(gP + 4544 + threadIdx.x, fb[34 + threadIdx.x]);
      atomicAdd(gP + 4512 + threadIdx.x, fb[66 + threadIdx.x]);
      sv[threadIdx.x] = fb[34 + threadIdx.x];
      sv[32 + threadIdx.x] = fb[66 + threadIdx.x];
    }
    if (threadIdx.x == 0) sv[64] = (float)(r1 - r0);          // the global row count travels with the first reduction
    cluster_sum(cl, sv, 65, sm, gather, A.px);
  }
  const float ninv = __fdiv_rn(1.0f, sv[64]);
  __syncthreads();
  // ---- layer 3 backward (W6: 32 x 64) + BN2 sums (units lane, lane + 32) -----------------------------------------
  const float mu2a = __ldg(mean + 32 + lane), rs2a = rstd_of(__ldg(var + 32 + lane), A.eps), g3a = __ldg(g3 + lane), e3a = __ldg(be3 + lane);
  const float mu2b = __ldg(mean + 64 + lane), rs2b = rstd_of(__ldg(var + 64 + lane), A.eps), g3b = __ldg(g3 + lane + 32), e3b = __ldg(be3 + lane + 32);
  {
    const float c1 = __fmul_rn(sv[lane], ninv), c2 = __fmul_rn(sv[32 + lane], ninv);
    __syncthreads();
    float dW[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) dW[k] = 0.f;
    float db = 0.f, s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
    flush_begin(fb, 2048 + 32);
    for (int r = r0 + warp; r < r1; r += nwarps) {
      float* s = zb + (long long)r * ld;
      const float xh3 = __fmul_rn(__fsub_rn(s[96 + lane], mu3), rs3);
      const float gz = __fmul_rn(__fmul_rn(g6l, rs3), __fsub_rn(__fsub_rn(s[128 + lane], c1), __fmul_rn(xh3, c2)));
      const float z2a = s[32 + lane], z2b = s[64 + lane];
      const float ha = bn_relu1(z2a, mu2a, rs2a, g3a, e3a), hb = bn_relu1(z2b, mu2b, rs2b, g3b, e3b);
      __syncwarp();
      rb[lane] = ha; rb[lane + 32] = hb;
      __syncwarp();
      db = __fadd_rn(db, gz);
#pragma unroll
      for (int k = 0; k < 64; ++k) dW[k] = fmaf(gz, rb[k], dW[k]);
      float gha = 0.f, ghb = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float gj = __shfl_sync(0xffffffffu, gz, j);
        gha = fmaf(W6s[j * 65 + lane], gj, gha);
        ghb = fmaf(W6s[j * 65 + lane + 32], gj, ghb);
      }
      const float gya = ha > 0.f ? gha : 0.f, gyb = hb > 0.f ? ghb : 0.f;
      s[160 + lane] = gya; s[192 + lane] = gyb;
      s1a = __fadd_rn(s1a, gya); s1b = __fadd_rn(s1b, gyb);
      s2a = fmaf(gya, __fmul_rn(__fsub_rn(z2a, mu2a), rs2a), s2a);
      s2b = fmaf(gyb, __fmul_rn(__fsub_rn(z2b, mu2b), rs2b), s2b);
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) atomicAdd(fb + k * 32 + lane, dW[k]);      // [k][unit]: conflict-free across lanes
    atomicAdd(fb + 2048 + lane, db);
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += MAP_NT) atomicAdd(gP + 2432 + (i & 31) * 64 + (i >> 5), fb[i]);   // W6[unit][k]
    if (threadIdx.x < 32) atomicAdd(gP + 4480 + threadIdx.x, fb[2048 + threadIdx.x]);                           // b6
    __syncthreads();
    flush_begin(fb, 128);
    atomicAdd(fb + lane, s1a); atomicAdd(fb + 32 + lane, s1b); atomicAdd(fb + 64 + lane, s2a); atomicAdd(fb + 96 + lane, s2b);
    __syncthreads();
    if (threadIdx.x < 64) {
      atomicAdd(gP + 2368 + threadIdx.x, fb[threadIdx.x]);                                                       // be3
      atomicAdd(gP + 2304 + threadIdx.x, fb[64 + threadIdx.x]);                                                  // g3
    }
    if (threadIdx.x < 128) sv[threadIdx.x] = fb[threadIdx.x];
    cluster_sum(cl, sv, 128, sm, gather, A.px);
  }
  // ---- layer 2 backward (W3: 64 x 32) + BN1 sums (unit lane) ------------------------------------------------------
  const float mu1 = __ldg(mean + lane), rs1 = rstd_of(__ldg(var + lane), A.eps), g0l = __ldg(g0 + lane), e0l = __ldg(be0 + lane);
  {
    const float c1a = __fmul_rn(sv[lane], ninv), c1b = __fmul_rn(sv[32 + lane], ninv);
    const float c2a = __fmul_rn(sv[64 + lane], ninv), c2b = __fmul_rn(sv[96 + lane], ninv);
    __syncthreads();
    float dWa[32], dWb[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { dWa[k] = 0.f; dWb[k] = 0.f; }
    float dba = 0.f, dbb = 0.f, s1 = 0.f, s2 = 0.f;
    flush_begin(fb, 2048 + 64);
    for (int r = r0 + warp; r < r1; r += nwarps) {
      float* s = zb + (long long)r * ld;
      const float xa = __fmul_rn(__fsub_rn(s[32 + lane], mu2a), rs2a), xb = __fmul_rn(__fsub_rn(s[64 + lane], mu2b), rs2b);
      const float gza = __fmul_rn(__fmul_rn(g3a, rs2a), __fsub_rn(__fsub_rn(s[160 + lane], c1a), __fmul_rn(xa, c2a)));
      const float gzb = __fmul_rn(__fmul_rn(g3b, rs2b), __fsub_rn(__fsub_rn(s[192 + lane], c1b), __fmul_rn(xb, c2b)));
      const float z1 = s[lane];
      const float h1 = bn_relu1(z1, mu1, rs1, g0l, e0l);
      __syncwarp();
      rb[lane] = h1;
      __syncwarp();
      dba = __fadd_rn(dba, gza); dbb = __fadd_rn(dbb, gzb);
#pragma unroll
      for (int k = 0; k < 32; ++k) { const float h = rb[k]; dWa[k] = fmaf(gza, h, dWa[k]); dWb[k] = fmaf(gzb, h, dWb[k]); }
      float gh = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        gh = fmaf(W3s[j * 33 + lane], __shfl_sync(0xffffffffu, gza, j), gh);
        gh = fmaf(W3s[(j + 32) * 33 + lane], __shfl_sync(0xffffffffu, gzb, j), gh);
      }
      const float gy = h1 > 0.f ? gh : 0.f;
      s[MAP_GY1 + lane] = gy;
      s1 = __fadd_rn(s1, gy);
      s2 = fmaf(gy, __fmul_rn(__fsub_rn(z1, mu1), rs1), s2);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) { atomicAdd(fb + k * 64 + lane, dWa[k]); atomicAdd(fb + k * 64 + 32 + lane, dWb[k]); }   // [k][unit]
    atomicAdd(fb + 2048 + lane, dba); atomicAdd(fb + 2048 + 32 + lane, dbb);
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += MAP_NT) atomicAdd(gP + 192 + (i & 63) * 32 + (i >> 6), fb[i]);    // W3[unit][k]
    if (threadIdx.x < 64) atomicAdd(gP + 2240 + threadIdx.x, fb[2048 + threadIdx.x]);                           // b3
    __syncthreads();
    flush_begin(fb, 64);
    atomicAdd(fb + lane, s1); atomicAdd(fb + 32 + lane, s2);
    __syncthreads();
    if (threadIdx.x < 32) {
      atomicAdd(gP + 160 + threadIdx.x, fb[threadIdx.x]);                                                        // be0
      atomicAdd(gP + 128 + threadIdx.x, fb[32 + threadIdx.x]);                                                   // g0
    }
    if (threadIdx.x < 64) sv[threadIdx.x] = fb[threadIdx.x];
    cluster_sum(cl, sv, 64, sm, gather, A.px);
  }
  // ---- layer 1 backward (W0: 32 x 3) and the input gradient ------------------------------------------------------------
  {
    const float c1 = __fmul_rn(sv[lane], ninv), c2 = __fmul_rn(sv[32 + lane], ninv);
    const float w0 = __ldg(W0 + lane * 3), w1 = __ldg(W0 + lane * 3 + 1), w2 = __ldg(W0 + lane * 3 + 2);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, dbv = 0.f;
    __syncthreads();
    flush_begin(fb, 128);
    for (int r = r0 + warp; r < r1; r += nwarps) {
      const float* s = zb + (long long)r * ld;
      const float xh = __fmul_rn(__fsub_rn(s[lane], mu1), rs1);
      const float gz = __fmul_rn(__fmul_rn(g0l, rs1), __fsub_rn(__fsub_rn(s[MAP_GY1 + lane], c1), __fmul_rn(xh, c2)));
      const float craw = __ldg(A.c + r);
      const float c = fminf(fmaxf(craw, 0.f), 1.f);
      const float f1 = __fmul_rn(c, c), f2 = log1p_f64(c);
      d0 = fmaf(gz, c, d0); d1 = fmaf(gz, f1, d1); d2 = fmaf(gz, f2, d2);
      dbv = __fadd_rn(dbv, gz);
      const float gf0 = wsum(__fmul_rn(w0, gz)), gf1 = wsum(__fmul_rn(w1, gz)), gf2 = wsum(__fmul_rn(w2, gz));
      // d/dc [c, c^2, log1p c] and the clamp's pass-through inside [0, 1]
      const float g = __fadd_rn(__fadd_rn(gf0, __fmul_rn(gf1, __fmul_rn(2.f, c))), __fdiv_rn(gf2, __fadd_rn(1.f, c)));
      if (A.gc && lane == 0) A.gc[r] = (craw >= 0.f && craw <= 1.f) ? g : 0.f;
    }
    atomicAdd(fb + lane * 3, d0); atomicAdd(fb + lane * 3 + 1, d1); atomicAdd(fb + lane * 3 + 2, d2);
    atomicAdd(fb + 96 + lane, dbv);
    flush_end(fb, 128, gP);                                                                                     // W0, b0
  }
  cl.sync();
}

// =================================================================================================
// soft mask: backward of m = smooth5x5(nearest_up(softmax(net([bits_norm, act]))[0]))  (quantization.py:213-239)
// =================================================================================================
// one CTA per image.  smem: dmt[nt] | rec[nt][36]: gl, hid[8], gh[8], in[2][9] (+1 pad)
constexpr int SM_REC = 36;
__global__ void __launch_bounds__(256)
softmask_bwd_kernel(const float* __restrict__ dm, const float* __restrict__ bit_map, const float* __restrict__ act_n,
                    const float* __restrict__ P, int H, int W, int Ht, int Wt, float* __restrict__ dbit,
                    float* __restrict__ gP) {
  extern __shared__ __align__(16) float sm[];
  const int nt = Ht * Wt, b = blockIdx.x;
  float* dmt = sm;
  float* rec = dmt + nt;
  const float* W0 = P; const float* b0 = P + 144; const float* W2 = P + 152; const float* b2 = P + 168; const float* ks = P + 170;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // (1) smoothing (replicate padding) + nearest upsampling, transposed -- as a GATHER per tile: a warp sums, over the
  //     pixels whose 5x5 window can reach the tile, the pixel gradient times the weight of the taps that land in it
  const float sy = (float)Ht / (float)H, sx = (float)Wt / (float)W;
  const float* g = dm + (long long)b * H * W;
  for (int t = warp; t < nt; t += nwarps) {
    const int ny = t / Wt, nx = t - ny * Wt;
    // pixel range of the tile under the nearest rule: [hs, he) x [ws, we)
    int hs = (int)ceilf((float)ny / sy), he = (int)ceilf((float)(ny + 1) / sy);
    while (hs > 0 && nearest_src(hs - 1, sy, Ht) >= ny) --hs;
    while (hs < H && nearest_src(hs, sy, Ht) < ny) ++hs;
    he = min(max(he, hs), H);
    while (he > hs && nearest_src(he - 1, sy, Ht) > ny) --he;
    while (he < H && nearest_src(he, sy, Ht) <= ny) ++he;
    int ws = (int)ceilf((float)nx / sx), we = (int)ceilf((float)(nx + 1) / sx);
    while (ws > 0 && nearest_src(ws - 1, sx, Wt) >= nx) --ws;
    while (ws < W && nearest_src(ws, sx, Wt) < nx) ++ws;
    we = min(max(we, ws), W);
    while (we > ws && nearest_src(we - 1, sx, Wt) > nx) --we;
    while (we < W && nearest_src(we, sx, Wt) <= nx) ++we;
    const int y0 = max(hs - 2, 0), y1 = min(he + 2, H), x0 = max(ws - 2, 0), x1 = min(we + 2, W);
    const int ww = x1 - x0, np = (y1 - y0) * ww;
    float acc = 0.f;
    for (int i = lane; i < np; i += 32) {
      const int h = y0 + i / ww, w = x0 + i % ww;
      const float gp = __ldg(g + (long long)h * W + w);
      float wt_ = 0.f;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        const int q = min(max(h + ky - 2, 0), H - 1);
        if (q < hs || q >= he) continue;
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const int p = min(max(w + kx - 2, 0), W - 1);
          if (p >= ws && p < we) wt_ = __fadd_rn(wt_, __ldg(ks + ky * 5 + kx));
        }
      }
      acc = fmaf(gp, wt_, acc);
    }
    acc = wsum(acc);
    if (lane == 0) dmt[t] = acc;
  }
  __syncthreads();
  // (2) tile head: recomputed forward, per-tile records for the parameter gradients
  const float* bm = bit_map + (long long)b * nt;
  const float* an = act_n + (long long)b * nt;
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const int i = t / Wt, j = t - i * Wt;
    float* rc = rec + t * SM_REC;
    float in[2][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int yy = i + k / 3 - 1, xx = j + k % 3 - 1;
      const bool ok = yy >= 0 && yy < Ht && xx >= 0 && xx < Wt;
      in[0][k] = ok ? fminf(fmaxf(__fdiv_rn(__fsub_rn(bm[yy * Wt + xx], 2.0f), 6.0f), 0.f), 1.f) : 0.f;
      in[1][k] = ok ? an[yy * Wt + xx] : 0.f;
      rc[17 + k] = in[0][k];
      rc[26 + k] = in[1][k];
    }
    float hid[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < 9; ++k) a = fmaf(in[c][k], __ldg(W0 + (o * 2 + c) * 9 + k), a);
      hid[o] = fmaxf(__fadd_rn(a, __ldg(b0 + o)), 0.f);
    }
    float l0 = __ldg(b2), l1 = __ldg(b2 + 1);
#pragma unroll
    for (int o = 0; o < 8; ++o) { l0 = fmaf(hid[o], __ldg(W2 + o), l0); l1 = fmaf(hid[o], __ldg(W2 + 8 + o), l1); }
    const float mx = fmaxf(l0, l1);
    const float e0 = exp_f64(__fsub_rn(l0, mx)), e1 = exp_f64(__fsub_rn(l1, mx));
    const float m = __fdiv_rn(e0, __fadd_rn(e0, e1));
    const float gl = __fmul_rn(dmt[t], __fmul_rn(m, __fsub_rn(1.f, m)));     // d m / d l0 = m (1 - m) = - d m / d l1
    rc[0] = gl;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      rc[1 + o] = hid[o];
      rc[9 + o] = hid[o] > 0.f ? __fmul_rn(gl, __fsub_rn(__ldg(W2 + o), __ldg(W2 + 8 + o))) : 0.f;
    }
  }
  __syncthreads();
  // (3) parameter gradients: a thread per parameter sums over the tiles; d bit_map: a thread per tile gathers
  for (int p = threadIdx.x; p < 170; p += blockDim.x) {
    float a = 0.f;
    if (p < 144) {
      const int o = p / 18, ck = p - o * 18;                   // W0[o][c][k]: in[c][k] at rec[17 + c * 9 + k]
      for (int t = 0; t < nt; ++t) a = fmaf(rec[t * SM_REC + 9 + o], rec[t * SM_REC + 17 + ck], a);
    } else if (p < 152) {
      for (int t = 0; t < nt; ++t) a = __fadd_rn(a, rec[t * SM_REC + 9 + (p - 144)]);
    } else if (p < 168) {
      const int q = p - 152, i = q & 7;
      for (int t = 0; t < nt; ++t) a = fmaf(rec[t * SM_REC], rec[t * SM_REC + 1 + i], a);
      if (q >= 8) a = -a;
    } else {
      for (int t = 0; t < nt; ++t) a = __fadd_rn(a, rec[t * SM_REC]);
      if (p == 169) a = -a;
    }
    if (a != 0.f) atomicAdd(gP + p, a);
  }
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const int i = t / Wt, j = t - i * Wt;
    float gin = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {                              // tile s = t - (k - centre) reads this tile through tap k
      const int yy = i - (k / 3 - 1), xx = j - (k % 3 - 1);
      if (yy < 0 || yy >= Ht || xx < 0 || xx >= Wt) continue;
      const float* rs_ = rec + (yy * Wt + xx) * SM_REC + 9;
#pragma unroll
      for (int o = 0; o < 8; ++o) gin = fmaf(rs_[o], __ldg(W0 + (o * 2) * 9 + k), gin);
    }
    const float bn = __fdiv_rn(__fsub_rn(bm[t], 2.0f), 6.0f);
    dbit[(long long)b * nt + t] = (bn >= 0.f && bn <= 1.f) ? __fdiv_rn(gin, 6.0f) : 0.f;
  }
}

// normalised tile activity act / (max_img + 1e-8) of the soft mask (quantization.py:224-226), no gradient:
// one CTA per image from the sum_c |x| plane (window = adaptive average pool)
__global__ void __launch_bounds__(256)
softmask_act_kernel(const float* __restrict__ abs_plane, int C, int H, int W, int Ht, int Wt, float* __restrict__ act_n) {
  extern __shared__ float act[];
  __shared__ float amax_s;
  const int b = blockIdx.x, nt = Ht * Wt;
  softmask_act_generic(abs_plane + (long long)b * H * W, C, H, W, Ht, Wt, 0, Ht, act);
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = -INFINITY;
    for (int t = threadIdx.x; t < nt; t += 32) a = fmaxf(a, act[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
    if (threadIdx.x == 0) amax_s = a;
  }
  __syncthreads();
  const float den = __fadd_rn(amax_s, 1e-8f);
  for (int t = threadIdx.x; t < nt; t += blockDim.x) act_n[(long long)b * nt + t] = __fdiv_rn(act[t], den);
}

// sum of a bit map and its total variation (sum |b[i+1,j] - b[i,j]| + |b[i,j+1] - b[i,j]|), one CTA per image,
// atomically into out[0], out[1] (models/mcaq_yolo.py:86-118, 575); gout: d/db of w0 * sum + w1 * tv
__global__ void __launch_bounds__(256)
bit_stats_kernel(const float* __restrict__ bm, int ht, int wt, float* __restrict__ out, const float* __restrict__ wgt,
                 float* __restrict__ gout) {
  const int b = blockIdx.x, nt = ht * wt;
  const float* p = bm + (long long)b * nt;
  float s = 0.f, tv = 0.f;
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const int y = t / wt, x = t - y * wt;
    const float v = p[t];
    s += v;
    float g = wgt ? wgt[0] : 0.f;
    if (y + 1 < ht) { const float d = p[t + wt] - v; tv += fabsf(d); if (wgt) g -= wgt[1] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
    if (x + 1 < wt) { const float d = p[t + 1] - v; tv += fabsf(d); if (wgt) g -= wgt[1] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
    if (wgt) {
      if (y > 0) { const float d = v - p[t - wt]; g += wgt[1] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
      if (x > 0) { const float d = v - p[t - 1]; g += wgt[1] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
      gout[(long long)b * nt + t] = g;
    }
  }
  if (out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); tv += __shfl_xor_sync(0xffffffffu, tv, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, s); atomicAdd(out + 1, tv); }
  }
}

// ---- avg_bits / Lbit / Lsmooth of up to four scales in ONE launch each way (models/mcaq_yolo.py:575, 86-118) ------------
// out[0] = avg_bits = mean_s mean(b_s), out[1] = (avg_bits - target)^2, out[2] = mean_s TV(b_s) / edges_s,
// out[3 + 2 s], out[4 + 2 s]: the per-scale sum and total variation (kept for the backward and for the sharded path)
constexpr int BL_MAX = 4;
struct BitLossArgs {
  const float* bm[BL_MAX]; float* grad[BL_MAX];
  int B[BL_MAX], ht[BL_MAX], wt[BL_MAX];
  int S; float target;
  float* out; const float* gout;
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += red[w];                    // fixed order: deterministic
  return t;
}

__device__ __forceinline__ float edges_of(int B, int ht, int wt) {
  const float e = (float)B * (float)((ht - 1) * wt + ht * (wt - 1));
  return e < 1.f ? 1.f : e;
}

__global__ void __launch_bounds__(256) bit_losses_fwd_kernel(const BitLossArgs A) {
  __shared__ float red[8];
  float avg = 0.f, sm = 0.f;
  for (int s = 0; s < A.S; ++s) {
    const int ht = A.ht[s], wt = A.wt[s], nt = ht * wt, n = A.B[s] * nt;
    const float* p = A.bm[s];
    float sum = 0.f, tv = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
      const int t = i % nt, y = t / wt, x = t - y * wt;
      const float v = p[i];
      sum += v;
      if (y + 1 < ht) tv += fabsf(p[i + wt] - v);
      if (x + 1 < wt) tv += fabsf(p[i + 1] - v);
    }
    sum = block_sum_256(sum, red);
    tv = block_sum_256(tv, red);
    avg += sum / (float)n;
    sm += tv / edges_of(A.B[s], ht, wt);
    if (threadIdx.x == 0) { A.out[3 + 2 * s] = sum; A.out[4 + 2 * s] = tv; }
  }
  if (threadIdx.x == 0) {
    avg /= (float)A.S;
    const float d = avg - A.target;
    A.out[0] = avg; A.out[1] = d * d; A.out[2] = sm / (float)A.S;
  }
}

// one CTA per (image chunk, scale): d/db of g0 * avg_bits + g1 * Lbit + g2 * Lsmooth
__global__ void __launch_bounds__(256) bit_losses_bwd_kernel(const BitLossArgs A) {
  const int s = blockIdx.y;
  const int ht = A.ht[s], wt = A.wt[s], nt = ht * wt, n = A.B[s] * nt;
  const float avg = A.out[0];
  const float w0 = (A.gout[0] + A.gout[1] * 2.f * (avg - A.target)) / ((float)A.S * (float)n);
  const float w1 = A.gout[2] / ((float)A.S * edges_of(A.B[s], ht, wt));
  const float* p = A.bm[s];
  float* g = A.grad[s];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int t = i % nt, y = t / wt, x = t - y * wt;
    const float v = p[i];
    float acc = w0;
    if (y + 1 < ht) { const float d = p[i + wt] - v; acc -= w1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
    if (x + 1 < wt) { const float d = p[i + 1] - v; acc -= w1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
    if (y > 0) { const float d = v - p[i - wt]; acc += w1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
    if (x > 0) { const float d = v - p[i - 1]; acc += w1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
    g[i] = acc;
  }
}

}  // namespace mcaq

using namespace mcaq;

// avg_bits, Lbit = (avg_bits - target)^2 and Lsmooth of S <= 4 bit maps (B_s, ht_s, wt_s) in one launch: out holds
// 3 + 2 S floats (see BitLossArgs).  With grad_out (3 floats, device) and grads (S pointers): the backward launch.
extern "C" int mcaq_bit_losses(const float* const* bit_maps, const int* B, const int* ht, const int* wt, int S, float target,
                               float* out, const float* grad_out, float* const* grads, void* stream) {
  if (!bit_maps || !B || !ht || !wt || S < 1 || S > BL_MAX || !out || (grad_out && !grads)) return MCAQ_EINVAL;
  BitLossArgs A = {};
  int nmax = 0;
  for (int s = 0; s < S; ++s) {
    if (!bit_maps[s] || B[s] <= 0 || ht[s] <= 0 || wt[s] <= 0 || (grad_out && !grads[s])) return MCAQ_EINVAL;
    A.bm[s] = bit_maps[s]; A.B[s] = B[s]; A.ht[s] = ht[s]; A.wt[s] = wt[s];
    A.grad[s] = grad_out ? grads[s] : nullptr;
    const int n = B[s] * ht[s] * wt[s];
    nmax = n > nmax ? n : nmax;
  }
  A.S = S; A.target = target; A.out = out; A.gout = grad_out;
  if (!grad_out) {
    bit_losses_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(A);
  } else {
    int gx = (nmax + 1023) / 1024;
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
    bit_losses_bwd_kernel<<<dim3(gx, S), 256, 0, (cudaStream_t)stream>>>(A);
  }
  MCAQ_LAUNCH_CHECK();
  return 0;
}

extern "C" long long mcaq_cmlp_train_scratch_floats(int N) { return N > 0 ? 4 : MCAQ_EINVAL; }   // none needed any more
extern "C" long long mcaq_mapper_train_scratch_floats(int N) { return N > 0 ? (long long)N * MP_SCR : MCAQ_EINVAL; }

// complexity path backward: g_craw (through bilateral + clamp) and the MLP's parameter gradients (gP: 2881 floats,
// ACCUMULATED into -- zero it first).  craw: the MLP output before the bilateral filter (mcaq_complexity's raw output).
extern "C" int mcaq_complexity_train_bwd(const float* phi, const float* craw, const float* grad_out, int B, int ht, int wt,
                                         const float* params, float* scratch, float* grad_craw_ws, float* grad_params,
                                         void* stream) {
  if (!phi || !craw || !grad_out || !params || !scratch || !grad_craw_ws || !grad_params || B <= 0 || ht <= 0 || wt <= 0)
    return MCAQ_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int nt = ht * wt, N = B * nt;
  bilateral_bwd_kernel<<<B, 256, nt * sizeof(float), st>>>(craw, grad_out, ht, wt, grad_craw_ws);
  (void)scratch;
  int ctas = (N + 4 * (CM_NT / 32) - 1) / (4 * (CM_NT / 32));       // >= 4 rows per warp, at most 64 CTAs (one flush each)
  ctas = ctas < 1 ? 1 : (ctas > 64 ? 64 : ctas);
  cmlp_bwd_kernel<<<ctas, CM_NT, 0, st>>>(phi, grad_craw_ws, params, N, grad_params);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

static int g_map_cluster = 0;                                  // 0: not probed yet; else the cluster size in use (16 or 8)
// debug / tuning: force the mapper kernels' cluster size (8 or 16; 0 = probe)
extern "C" void mcaq_debug_mapper_cluster(int n) { g_map_cluster = (n == 8 || n == 16) ? n : 0; }

static int launch_mapper(void (*k)(const MapArgs), MapArgs& A, cudaStream_t st) {
  const size_t smem_max = (size_t)(MAP_SM_FLOATS + MAP_ROWS_MAX * MAP_ROWF) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mapper_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    cudaFuncSetAttribute(mapper_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    cudaFuncSetAttribute(mapper_train_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(mapper_train_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  auto configure = [&](int ncta) {
    A.rpc = (A.N + ncta - 1) / ncta;
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(MAP_NT);
    cfg.dynamicSmemBytes = (size_t)(MAP_SM_FLOATS + (A.rpc <= MAP_ROWS_MAX ? A.rpc * MAP_ROWF : 0)) * sizeof(float);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  };
  if (g_map_cluster == 0) {
    // the rows are spread over ONE cluster (the batch statistics need every row): 16 CTAs when the device can co-schedule
    // a 16-CTA cluster at the largest footprint (a GPC with 16 free SMs), else the portable 8
    configure(MAP_CL);
    cfg.dynamicSmemBytes = smem_max;
    int nclusters = 0;
    const cudaError_t q = cudaOccupancyMaxActiveClusters(&nclusters, k, &cfg);
    if (q != cudaSuccess) cudaGetLastError();
    g_map_cluster = (q == cudaSuccess && nclusters >= 1) ? MAP_CL : 8;
  }
  configure(g_map_cluster);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, A);
  if (e != cudaSuccess) return (int)e;
  MCAQ_LAUNCH_CHECK();
  return 0;
}

static int fill_px(XchgPeers& px, void* const* peers, int rank, int world) {
  px.world = 1; px.rank = 0; px.timeout_ns = xchg_timeout_ns();
  if (world <= 1) return 0;
  if (!peers || world > XCHG_MAX_RANKS || rank < 0 || rank >= world) return MCAQ_EINVAL;
  for (int i = 0; i < world; ++i) { if (!peers[i]) return MCAQ_EINVAL; px.base[i] = reinterpret_cast<float*>(peers[i]); }
  px.rank = rank; px.world = world;
  return 0;
}

// Train-mode mapper forward: continuous bits (N), pre-activations kept in scratch, batch statistics in stats (256
// floats: mean[128] var[128]); running_mean / running_var of the three BatchNorm layers updated in place when given
// (momentum 0.1, unbiased variance, like nn.BatchNorm1d).  xchg_*: exchange buffers of a C = 128 RangeExchange when
// the batch is sharded over ranks (statistics of the whole batch on every rank), else world = 1.
extern "C" int mcaq_mapper_train_fwd(const float* cmap, int N, const float* params, float temperature, int use_temperature,
                                     float min_bits, float max_bits, float* scratch, float* stats, float* rm0, float* rv0,
                                     float* rm1, float* rv1, float* rm2, float* rv2, long long* nbt0, long long* nbt1,
                                     long long* nbt2, float momentum, float eps, float* bits,
                                     void* const* xchg_peers, int xchg_rank, int xchg_world, void* stream) {
  if (!cmap || !params || !scratch || !stats || !bits || N <= 0) return MCAQ_EINVAL;
  MapArgs A = {};
  A.c = cmap; A.P = params; A.N = N; A.temperature = temperature; A.use_t = use_temperature; A.lo = min_bits; A.hi = max_bits;
  A.scratch = scratch; A.stats = stats; A.rm[0] = rm0; A.rv[0] = rv0; A.rm[1] = rm1; A.rv[1] = rv1; A.rm[2] = rm2; A.rv[2] = rv2;
  A.nbt[0] = nbt0; A.nbt[1] = nbt1; A.nbt[2] = nbt2;
  A.momentum = momentum; A.eps = eps; A.out = bits;
  int rc = fill_px(A.px, xchg_peers, xchg_rank, xchg_world);
  if (rc) return rc;
  return launch_mapper(mapper_train_fwd_kernel, A, (cudaStream_t)stream);
}

// backward: grad_c (N, may be NULL) and the parameter gradients (4609 floats, ACCUMULATED into)
extern "C" int mcaq_mapper_train_bwd(const float* cmap, int N, const float* params, float temperature, int use_temperature,
                                     float min_bits, float max_bits, float* scratch, const float* stats, float eps,
                                     const float* grad_bits, float* grad_c, float* grad_params, void* const* xchg_peers,
                                     int xchg_rank, int xchg_world, void* stream) {
  if (!cmap || !params || !scratch || !stats || !grad_bits || !grad_params || N <= 0) return MCAQ_EINVAL;
  MapArgs A = {};
  A.c = cmap; A.P = params; A.N = N; A.temperature = temperature; A.use_t = use_temperature; A.lo = min_bits; A.hi = max_bits;
  A.scratch = scratch; A.stats = const_cast<float*>(stats); A.eps = eps; A.gout = grad_bits; A.gc = grad_c; A.gP = grad_params;
  int rc = fill_px(A.px, xchg_peers, xchg_rank, xchg_world);
  if (rc) return rc;
  return launch_mapper(mapper_train_bwd_kernel, A, (cudaStream_t)stream);
}

extern "C" int mcaq_softmask_act(const float* abs_plane, int B, int C, int H, int W, int Ht, int Wt, float* act_norm,
                                 void* stream) {
  if (!abs_plane || !act_norm || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return MCAQ_EINVAL;
  softmask_act_kernel<<<B, 256, Ht * Wt * sizeof(float), (cudaStream_t)stream>>>(abs_plane, C, H, W, Ht, Wt, act_norm);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

// soft-mask backward: d bit_map (B,Ht,Wt) and the net's parameter gradients (170 floats, ACCUMULATED into)
extern "C" int mcaq_softmask_train_bwd(const float* grad_mask, const float* bit_map, const float* act_norm,
                                       const float* params, int B, int H, int W, int Ht, int Wt, float* grad_bit_map,
                                       float* grad_params, void* stream) {
  if (!grad_mask || !bit_map || !act_norm || !params || !grad_bit_map || !grad_params || B <= 0 || H <= 0 || W <= 0 ||
      Ht <= 0 || Wt <= 0)
    return MCAQ_EINVAL;
  const size_t smem = ((size_t)Ht * Wt * (1 + SM_REC)) * sizeof(float);
  if (smem > 200 * 1024) return MCAQ_ETOOBIG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(softmask_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  softmask_bwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(grad_mask, bit_map, act_norm, params, H, W, Ht, Wt, grad_bit_map,
                                                            grad_params);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

// out[0] += sum(bit_map), out[1] += TV(bit_map); with weights (device, 2 floats) also grad = d(w0 sum + w1 TV)/d bit_map
extern "C" int mcaq_bit_stats(const float* bit_map, int B, int ht, int wt, float* out2, const float* weights2,
                              float* grad_bit_map, void* stream) {
  if (!bit_map || B <= 0 || ht <= 0 || wt <= 0 || (!out2 && !grad_bit_map) || (grad_bit_map && !weights2)) return MCAQ_EINVAL;
  bit_stats_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(bit_map, ht, wt, out2, grad_bit_map ? weights2 : nullptr, grad_bit_map);
  MCAQ_LAUNCH_CHECK();
  return 0;
}
