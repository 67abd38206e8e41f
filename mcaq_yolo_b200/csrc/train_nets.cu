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
// Rows are independent except for the mapper's BatchNorm: its kernels run as ONE thread-block cluster (8 CTAs,
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

// ---- small dense helpers: thread-private vectors, weights read at warp-uniform addresses (broadcast) ----------
template <int IN, int OUT>
__device__ __forceinline__ void dense_fwd(const float* __restrict__ Wm, const float* __restrict__ b, const float (&x)[IN],
                                          float (&y)[OUT]) {
#pragma unroll 4
  for (int j = 0; j < OUT; ++j) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < IN; ++i) a = fmaf(__ldg(Wm + j * IN + i), x[i], a);
    y[j] = __fadd_rn(a, __ldg(b + j));
  }
}
template <int IN, int OUT>
__device__ __forceinline__ void dense_bwd_data(const float* __restrict__ Wm, const float (&gy)[OUT], float (&gx)[IN]) {
#pragma unroll
  for (int i = 0; i < IN; ++i) gx[i] = 0.f;
#pragma unroll 4
  for (int j = 0; j < OUT; ++j) {
    const float g = gy[j];
#pragma unroll
    for (int i = 0; i < IN; ++i) gx[i] = fmaf(__ldg(Wm + j * IN + i), g, gx[i]);
  }
}

// gW[j][i] += sum_r GY[r][j] * X[r][i], gb[j] += sum_r GY[r][j] over rows [r0, r1): thread per parameter
__device__ __forceinline__ void wgrad_rows(const float* __restrict__ GY, int ldg_, const float* __restrict__ X, int ldx,
                                           int r0, int r1, int IN, int OUT, float* __restrict__ gW, float* __restrict__ gb) {
  for (int p = threadIdx.x; p < OUT * IN + OUT; p += blockDim.x) {
    float acc = 0.f;
    if (p < OUT * IN) {
      const int j = p / IN, i = p - j * IN;
      for (int r = r0; r < r1; ++r) acc = fmaf(GY[(long long)r * ldg_ + j], X[(long long)r * ldx + i], acc);
      atomicAdd(gW + p, acc);
    } else {
      const int j = p - OUT * IN;
      for (int r = r0; r < r1; ++r) acc = __fadd_rn(acc, GY[(long long)r * ldg_ + j]);
      atomicAdd(gb + j, acc);
    }
  }
}
// column sums of P[r][k] (and optionally Q[r][k]) over rows [r0, r1) added into g1 / g2 (LN / BN affine gradients)
__device__ __forceinline__ void colsum_rows(const float* __restrict__ P, const float* __restrict__ Q, int ld, int r0, int r1,
                                            int K, float* __restrict__ g1, float* __restrict__ g2) {
  for (int k = threadIdx.x; k < 2 * K; k += blockDim.x) {
    const float* src = k < K ? P : Q;
    if (!src) continue;
    const int kk = k < K ? k : k - K;
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc = __fadd_rn(acc, src[(long long)r * ld + kk]);
    atomicAdd((k < K ? g1 : g2) + kk, acc);
  }
}

template <int D>
__device__ __forceinline__ void ln_fwd(const float (&z)[D], const float* __restrict__ g, const float* __restrict__ b,
                                       float (&xh)[D], float (&y)[D], float& rstd) {
  float m = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) m = __fadd_rn(m, z[k]);
  m = __fmul_rn(m, 1.0f / D);
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) { const float d = __fsub_rn(z[k], m); v = fmaf(d, d, v); }
  v = __fmul_rn(v, 1.0f / D);
  rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(v, 1e-5f)));
#pragma unroll
  for (int k = 0; k < D; ++k) {
    xh[k] = __fmul_rn(__fsub_rn(z[k], m), rstd);
    y[k] = fmaxf(fmaf(xh[k], __ldg(g + k), __ldg(b + k)), 0.f);
  }
}
// gz from gy (gradient at the LN output BEFORE the ReLU mask is applied by the caller)
template <int D>
__device__ __forceinline__ void ln_bwd(const float (&gy)[D], const float (&xh)[D], const float* __restrict__ g, float rstd,
                                       float (&gz)[D]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const float gx = __fmul_rn(gy[k], __ldg(g + k));
    gz[k] = gx;
    s1 = __fadd_rn(s1, gx);
    s2 = fmaf(gx, xh[k], s2);
  }
  s1 = __fmul_rn(s1, 1.0f / D);
  s2 = __fmul_rn(s2, 1.0f / D);
#pragma unroll
  for (int k = 0; k < D; ++k) gz[k] = __fmul_rn(rstd, __fsub_rn(__fsub_rn(gz[k], s1), __fmul_rn(xh[k], s2)));
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

// (2) MLP: rows [blockIdx.x * rpc, ...): data path (thread per row) into scratch, then the CTA's share of the
//     parameter gradients.  scratch per row (floats): H1[64] GZ1[64] P1[64] Q1[64] | H2[32] GZ2[32] P2[32] Q2[32] | GZ3
constexpr int CM_SCR = 4 * 64 + 4 * 32 + 4;
__global__ void __launch_bounds__(128)
cmlp_bwd_kernel(const float* __restrict__ phi, const float* __restrict__ gcraw, const float* __restrict__ P, int N, int rpc,
                float* __restrict__ scratch, float* __restrict__ gP) {
  const float* W0 = P; const float* b0 = P + 512; const float* g1 = P + 576; const float* be1 = P + 640;
  const float* W3 = P + 704; const float* b3 = P + 2752; const float* g4 = P + 2784; const float* be4 = P + 2816;
  const float* W6 = P + 2848; const float* b6 = P + 2880;
  const int r0 = blockIdx.x * rpc, r1 = min(r0 + rpc, N);
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    float x[8], z1[64], xh1[64], h1[64], z2[32], xh2[32], h2[32];
    float rs1, rs2;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = phi[(long long)r * 8 + i];
    dense_fwd<8, 64>(W0, b0, x, z1);
    ln_fwd<64>(z1, g1, be1, xh1, h1, rs1);
    dense_fwd<64, 32>(W3, b3, h1, z2);
    ln_fwd<32>(z2, g4, be4, xh2, h2, rs2);
    float z3 = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) z3 = fmaf(__ldg(W6 + k), h2[k], z3);
    z3 = __fadd_rn(z3, __ldg(b6));
    const float c = sigmoid_exact(z3);
    const float gz3 = __fmul_rn(gcraw[r], __fmul_rn(c, __fsub_rn(1.0f, c)));
    float* s = scratch + (long long)r * CM_SCR;
    float gy2[32], gz2[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) gy2[k] = h2[k] > 0.f ? __fmul_rn(gz3, __ldg(W6 + k)) : 0.f;
    ln_bwd<32>(gy2, xh2, g4, rs2, gz2);
    float gh1[64], gz1[64];
    dense_bwd_data<64, 32>(W3, gz2, gh1);
#pragma unroll
    for (int k = 0; k < 64; ++k) gh1[k] = h1[k] > 0.f ? gh1[k] : 0.f;
    ln_bwd<64>(gh1, xh1, g1, rs1, gz1);
#pragma unroll
    for (int k = 0; k < 64; ++k) { s[k] = h1[k]; s[64 + k] = gz1[k]; s[128 + k] = __fmul_rn(gh1[k], xh1[k]); s[192 + k] = gh1[k]; }
#pragma unroll
    for (int k = 0; k < 32; ++k) { s[256 + k] = h2[k]; s[288 + k] = gz2[k]; s[320 + k] = __fmul_rn(gy2[k], xh2[k]); s[352 + k] = gy2[k]; }
    s[384] = gz3;
  }
  __syncthreads();
  const float* S = scratch;
  wgrad_rows(S + 64, CM_SCR, phi, 8, r0, r1, 8, 64, gP, gP + 512);                     // W0, b0
  colsum_rows(S + 128, S + 192, CM_SCR, r0, r1, 64, gP + 576, gP + 640);               // g1, be1
  wgrad_rows(S + 288, CM_SCR, S, CM_SCR, r0, r1, 64, 32, gP + 704, gP + 2752);         // W3, b3
  colsum_rows(S + 320, S + 352, CM_SCR, r0, r1, 32, gP + 2784, gP + 2816);             // g4, be4
  wgrad_rows(S + 384, CM_SCR, S + 256, CM_SCR, r0, r1, 32, 1, gP + 2848, gP + 2880);   // W6, b6
}

// =================================================================================================
// mapper with train-mode BatchNorm: one cluster of MAP_CL CTAs
// =================================================================================================
constexpr int MAP_CL = 8;
constexpr int MAP_NT = 256;
// scratch per row: Z1[32] Z2[64] Z3[32] (forward, kept for backward) | F[4] H1[32] H2[64] H3[32] GZ1[32] GZ2[64] GZ3[32] GZ4[4]
constexpr int MP_Z = 128;
constexpr int MP_SCR = MP_Z + 4 + 128 + 128 + 4;
// saved statistics: mean[128] rstd[128] (layer offsets 0 / 32 / 96)
struct MapArgs {
  const float* c; const float* P; int N; int rpc;
  float temperature; int use_t; float lo, hi;
  float* scratch; float* stats;
  float* rm[3]; float* rv[3]; float momentum; float eps;      // running statistics (NULL: not tracked)
  float* out;
  const float* gout; float* gc; float* gP;                     // backward
  XchgPeers px;                                                // world > 1: statistics merged over the ranks
};

// Cluster-wide per-feature statistics of Zl[r][0..F) over ALL rows of the cluster (and of all ranks):
// two-pass mean / M2 inside a CTA (fixed order), Chan merge over the CTAs in rank order through DSMEM, then over
// the GPU ranks through peer memory.  Result in sm_mean / sm_var (biased), count in *n_tot.  sm: >= 6*F + 8 floats.
__device__ void batch_stats(cg::cluster_group& cl, const float* Zl, int ld, int r0, int r1, int F, float* sm,
                            float* gather /*[MAP_CL][2F+1] in EVERY CTA*/, const XchgPeers& px, float* sm_mean, float* sm_var,
                            float& n_tot) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int SL = NT / F;                                       // row slices per feature
  float* part = sm;                                            // [SL][F]
  const int f = tid % F, sl = tid / F;
  const int n = r1 - r0;
  // pass 1: mean
  float a = 0.f;
  if (sl < SL) for (int r = r0 + sl; r < r1; r += SL) a = __fadd_rn(a, Zl[(long long)r * ld + f]);
  if (sl < SL) part[sl * F + f] = a;
  __syncthreads();
  float mean = 0.f;
  if (tid < F) {
    for (int s_ = 0; s_ < SL; ++s_) mean = __fadd_rn(mean, part[s_ * F + tid]);
    mean = n > 0 ? __fdiv_rn(mean, (float)n) : 0.f;
    sm_mean[tid] = mean;
  }
  __syncthreads();
  // pass 2: M2 around the CTA's mean
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
      float* g = cl.map_shared_rank(gather, p) + rank * (2 * F + 1);
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
      const float* g = gather + p * (2 * F + 1);
      const float nb = g[2 * F];
      if (nb <= 0.f) continue;
      const float d = __fsub_rn(g[tid], mu), tot = __fadd_rn(cnt, nb);
      mu = __fadd_rn(mu, __fmul_rn(d, __fdiv_rn(nb, tot)));
      M2 = __fadd_rn(__fadd_rn(M2, g[F + tid]), __fmul_rn(__fmul_rn(d, d), __fdiv_rn(__fmul_rn(cnt, nb), tot)));
      cnt = tot;
    }
  }
  // GPU ranks: CTA 0 publishes (count, mean, M2), waits for all, every CTA then reads the merged values from
  // CTA 0's shared memory.  Same protocol / bounded wait as the range exchange (peer_exchange.cuh).
  if (px.world > 1) {
    float* xm = sm + 3 * F;                                    // CTA 0: merged [mean F | M2 F | count]
    if (rank == 0) {
      __shared__ int step_s;
      float* local = px.base[px.rank];
      if (tid == 0) { int* ep = reinterpret_cast<int*>(local); step_s = *ep + 1; *ep = step_s; }
      __syncthreads();
      const int e = step_s;
      const int slotf = 2 * 128;                               // slot capacity (floats) of a C = 128 exchange buffer
      for (int p = 0; p < px.world; ++p) {
        float* dst = px.base[p] + XCHG_SLOTS + (long long)((e & 1) * px.world + px.rank) * slotf;
        if (tid < F) { dst[tid] = mu; dst[F + tid] = M2; }
        if (tid == 0) dst[2 * F] = cnt;
      }
      __threadfence_system();
      __syncthreads();
      if (tid < px.world) st_release_sys(reinterpret_cast<int*>(px.base[tid]) + XCHG_FLAGS + 8 * (e & 1) + px.rank, e);
      const bool ok = xchg_wait(local, px.world, e, tid, px.timeout_ns);
      const bool all_ok = __syncthreads_and(ok);
      if (tid < F) {
        float c2 = 0.f, m_ = 0.f, q_ = 0.f;
        for (int q = 0; q < px.world; ++q) {
          // a timed-out exchange degrades to this rank's own statistics (error word set, peer.check raises)
          const volatile float* s_ = local + XCHG_SLOTS + (long long)((e & 1) * px.world + q) * slotf;
          const float nb = all_ok ? s_[2 * F] : (q == px.rank ? cnt : 0.f);
          if (nb <= 0.f) continue;
          const float mq = all_ok ? s_[tid] : mu, qq = all_ok ? s_[F + tid] : M2;
          const float d = __fsub_rn(mq, m_), tot = __fadd_rn(c2, nb);
          m_ = __fadd_rn(m_, __fmul_rn(d, __fdiv_rn(nb, tot)));
          q_ = __fadd_rn(__fadd_rn(q_, qq), __fmul_rn(__fmul_rn(d, d), __fdiv_rn(__fmul_rn(c2, nb), tot)));
          c2 = tot;
        }
        xm[tid] = m_; xm[F + tid] = q_;
        if (tid == 0) xm[2 * F] = c2;
      }
    }
    cl.sync();
    if (tid < F) {
      const float* x0 = cl.map_shared_rank(sm + 3 * F, 0);
      mu = x0[tid]; M2 = x0[F + tid]; cnt = x0[2 * F];
    }
    cl.sync();                                                 // CTA 0's buffer is free for the next layer
  }
  if (tid < F) { sm_mean[tid] = mu; sm_var[tid] = cnt > 0.f ? __fdiv_rn(M2, cnt) : 0.f; }
  if (tid == 0) sm[6 * F] = cnt;
  __syncthreads();
  n_tot = sm[6 * F];
}

// cluster-wide sums of two per-feature quantities (BatchNorm backward: sum gy, sum gy * xhat), ranks included
__device__ void batch_sums2(cg::cluster_group& cl, const float* A, const float* Bq, int ld, int r0, int r1, int F, float* sm,
                            float* gather, const XchgPeers& px, float* s1, float* s2) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int SL = NT / (2 * F) > 0 ? NT / (2 * F) : 1;
  float* part = sm;                                            // [SL][2F]
  const int q = tid % (2 * F), sl = tid / (2 * F);
  float a = 0.f;
  if (sl < SL) {
    const float* src = q < F ? A : Bq;
    const int k = q < F ? q : q - F;
    for (int r = r0 + sl; r < r1; r += SL) a = __fadd_rn(a, src[(long long)r * ld + k]);
    part[sl * 2 * F + q] = a;
  }
  __syncthreads();
  const int rank = (int)cl.block_rank(), ncta = (int)cl.num_blocks();
  if (tid < 2 * F) {
    float t = 0.f;
    for (int s_ = 0; s_ < SL; ++s_) t = __fadd_rn(t, part[s_ * 2 * F + tid]);
    for (int p = 0; p < ncta; ++p) cl.map_shared_rank(gather, p)[rank * (2 * F + 1) + tid] = t;
  }
  cl.sync();
  float tot = 0.f;
  if (tid < 2 * F)
    for (int p = 0; p < ncta; ++p) tot = __fadd_rn(tot, gather[p * (2 * F + 1) + tid]);
  if (px.world > 1) {
    float* xm = sm + 3 * F;
    if (rank == 0) {
      __shared__ int step_b;
      float* local = px.base[px.rank];
      if (tid == 0) { int* ep = reinterpret_cast<int*>(local); step_b = *ep + 1; *ep = step_b; }
      __syncthreads();
      const int e = step_b;
      const int slotf = 2 * 128;
      for (int p = 0; p < px.world; ++p) {
        float* dst = px.base[p] + XCHG_SLOTS + (long long)((e & 1) * px.world + px.rank) * slotf;
        if (tid < 2 * F) dst[tid] = tot;
      }
      __threadfence_system();
      __syncthreads();
      if (tid < px.world) st_release_sys(reinterpret_cast<int*>(px.base[tid]) + XCHG_FLAGS + 8 * (e & 1) + px.rank, e);
      const bool ok = xchg_wait(local, px.world, e, tid, px.timeout_ns);
      const bool all_ok = __syncthreads_and(ok);
      if (tid < 2 * F) {
        float t = 0.f;
        for (int w = 0; w < px.world; ++w) {
          const volatile float* s_ = local + XCHG_SLOTS + (long long)((e & 1) * px.world + w) * slotf;
          t = __fadd_rn(t, all_ok ? s_[tid] : (w == px.rank ? tot : 0.f));
        }
        xm[tid] = t;
      }
    }
    cl.sync();
    if (tid < 2 * F) tot = cl.map_shared_rank(sm + 3 * F, 0)[tid];
    cl.sync();
  }
  if (tid < F) s1[tid] = tot;
  else if (tid < 2 * F) s2[tid - F] = tot;
  __syncthreads();
}

// BN (train) + ReLU of a stored pre-activation row
template <int F>
__device__ __forceinline__ void bn_relu_row(const float* __restrict__ z, const float* mean, const float* var, float eps,
                                            const float* __restrict__ g, const float* __restrict__ b, float (&h)[F]) {
#pragma unroll
  for (int k = 0; k < F; ++k) {
    const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[k], eps)));
    h[k] = fmaxf(fmaf(__fmul_rn(__fsub_rn(z[k], mean[k]), rstd), __ldg(g + k), __ldg(b + k)), 0.f);
  }
}

__device__ __forceinline__ void update_running(float* rm, float* rv, const float* mean, const float* var, float n, int F,
                                               float momentum) {
  if (!rm || (int)threadIdx.x >= F) return;
  const int k = threadIdx.x;
  const float unb = n > 1.f ? __fmul_rn(var[k], __fdiv_rn(n, __fsub_rn(n, 1.f))) : var[k];
  rm[k] = __fadd_rn(__fmul_rn(__fsub_rn(1.f, momentum), rm[k]), __fmul_rn(momentum, mean[k]));
  rv[k] = __fadd_rn(__fmul_rn(__fsub_rn(1.f, momentum), rv[k]), __fmul_rn(momentum, unb));
}

__global__ void __launch_bounds__(MAP_NT) mapper_train_fwd_kernel(const MapArgs A) {
  extern __shared__ float sm[];                                // work [6*64+8] | gather [MAP_CL][2*64+1] | mean[64] var[64]
  cg::cluster_group cl = cg::this_cluster();
  float* gather = sm + 6 * 64 + 8;
  float* mean = gather + MAP_CL * (2 * 64 + 1);
  float* var = mean + 64;
  const float* P = A.P;
  const float* W0 = P; const float* b0 = P + 96; const float* g0 = P + 128; const float* be0 = P + 160;
  const float* W3 = P + 192; const float* b3 = P + 2240; const float* g3 = P + 2304; const float* be3 = P + 2368;
  const float* W6 = P + 2432; const float* b6 = P + 4480; const float* g6 = P + 4512; const float* be6 = P + 4544;
  const float* W9 = P + 4576; const float* b9 = P + 4608;
  const int rank = (int)cl.block_rank();
  const int r0 = min(rank * A.rpc, A.N), r1 = min(r0 + A.rpc, A.N);
  float ntot;
  // layer 1
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const float c = fminf(fmaxf(A.c[r], 0.f), 1.f);
    float f[3] = {c, __fmul_rn(c, c), log1p_f64(c)}, z[32];
    dense_fwd<3, 32>(W0, b0, f, z);
    float* s = A.scratch + (long long)r * MP_SCR;
#pragma unroll
    for (int k = 0; k < 32; ++k) s[k] = z[k];
  }
  __syncthreads();
  batch_stats(cl, A.scratch, MP_SCR, r0, r1, 32, sm, gather, A.px, mean, var, ntot);
  if (rank == 0) {
    update_running(A.rm[0], A.rv[0], mean, var, ntot, 32, A.momentum);
    if (threadIdx.x < 32) { A.stats[threadIdx.x] = mean[threadIdx.x]; A.stats[128 + threadIdx.x] = var[threadIdx.x]; }
  }
  // layer 2
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    float* s = A.scratch + (long long)r * MP_SCR;
    float h[32], z[64];
    bn_relu_row<32>(s, mean, var, A.eps, g0, be0, h);
    dense_fwd<32, 64>(W3, b3, h, z);
#pragma unroll
    for (int k = 0; k < 64; ++k) s[32 + k] = z[k];
  }
  __syncthreads();
  batch_stats(cl, A.scratch + 32, MP_SCR, r0, r1, 64, sm, gather, A.px, mean, var, ntot);
  if (rank == 0) {
    update_running(A.rm[1], A.rv[1], mean, var, ntot, 64, A.momentum);
    if (threadIdx.x < 64) { A.stats[32 + threadIdx.x] = mean[threadIdx.x]; A.stats[160 + threadIdx.x] = var[threadIdx.x]; }
  }
  // layer 3
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    float* s = A.scratch + (long long)r * MP_SCR;
    float h[64], z[32];
    bn_relu_row<64>(s + 32, mean, var, A.eps, g3, be3, h);
    dense_fwd<64, 32>(W6, b6, h, z);
#pragma unroll
    for (int k = 0; k < 32; ++k) s[96 + k] = z[k];
  }
  __syncthreads();
  batch_stats(cl, A.scratch + 96, MP_SCR, r0, r1, 32, sm, gather, A.px, mean, var, ntot);
  if (rank == 0) {
    update_running(A.rm[2], A.rv[2], mean, var, ntot, 32, A.momentum);
    if (threadIdx.x < 32) { A.stats[96 + threadIdx.x] = mean[threadIdx.x]; A.stats[224 + threadIdx.x] = var[threadIdx.x]; }
  }
  // head: 32 -> 1, sigmoid, Eq.17, temperature, straight-through clamp (bit_allocation.py:258-273)
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const float* s = A.scratch + (long long)r * MP_SCR;
    float h[32];
    bn_relu_row<32>(s + 96, mean, var, A.eps, g6, be6, h);
    float z = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) z = fmaf(__ldg(W9 + k), h[k], z);
    const float sg = sigmoid_exact(__fadd_rn(z, __ldg(b9)));
    float bits = __fadd_rn(A.lo, __fmul_rn(__fsub_rn(A.hi, A.lo), sg));
    if (A.use_t) bits = __fmul_rn(bits, A.temperature);
    A.out[r] = fminf(fmaxf(bits, A.lo), A.hi);                 // value of bits + (clamp(bits) - bits).detach()
  }
  cl.sync();                                                   // nobody leaves while a peer may still read its gather area
}

__global__ void __launch_bounds__(MAP_NT) mapper_train_bwd_kernel(const MapArgs A) {
  extern __shared__ float sm[];
  cg::cluster_group cl = cg::this_cluster();
  float* gather = sm + 6 * 64 + 8;
  float* s1 = gather + MAP_CL * (2 * 64 + 1);
  float* s2 = s1 + 64;
  const float* P = A.P;
  const float* W0 = P; const float* g0 = P + 128; const float* be0 = P + 160;
  const float* W3 = P + 192; const float* g3 = P + 2304; const float* be3 = P + 2368;
  const float* W6 = P + 2432; const float* g6 = P + 4512; const float* be6 = P + 4544;
  const float* W9 = P + 4576; const float* b9 = P + 4608;
  const float* mean = A.stats; const float* var = A.stats + 128;
  float* gP = A.gP;
  const int rank = (int)cl.block_rank();
  const int r0 = min(rank * A.rpc, A.N), r1 = min(r0 + A.rpc, A.N);
  // total row count (all CTAs / ranks): the BN backward divides by it; counts travel with batch_sums2 as a feature
  // scratch columns (per row): 128 F[4] | 132 H1[32] | 164 H2[64] | 228 H3[32] | 260 GZ1[32] | 292 GZ2[64] | 356 GZ3[32] | 388 GZ4
  // ---- head + BN3: gy3 and gy3 * xhat3 into GZ3 / H3 columns temporarily --------------------------------------
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    float* s = A.scratch + (long long)r * MP_SCR;
    float h[32];
    bn_relu_row<32>(s + 96, mean + 96, var + 96, A.eps, g6, be6, h);
    float z = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) z = fmaf(__ldg(W9 + k), h[k], z);
    const float sg = sigmoid_exact(__fadd_rn(z, __ldg(b9)));
    float g = __fmul_rn(A.gout[r], __fmul_rn(__fsub_rn(A.hi, A.lo), __fmul_rn(sg, __fsub_rn(1.f, sg))));
    if (A.use_t) g = __fmul_rn(g, A.temperature);
    s[388] = g;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      s[228 + k] = h[k];
      const float gy = h[k] > 0.f ? __fmul_rn(g, __ldg(W9 + k)) : 0.f;
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[96 + k], A.eps)));
      const float xh = __fmul_rn(__fsub_rn(s[96 + k], mean[96 + k]), rstd);
      s[356 + k] = gy;                                         // gy3
      s[132 + k] = __fmul_rn(gy, xh);                          // gy3 * xhat3 (H1 columns are free until layer 1)
    }
    s[389] = 1.0f;                                             // row counter
  }
  __syncthreads();
  wgrad_rows(A.scratch + 388, MP_SCR, A.scratch + 228, MP_SCR, r0, r1, 32, 1, gP + 4576, gP + 4608);   // W9, b9
  batch_sums2(cl, A.scratch + 356, A.scratch + 132, MP_SCR, r0, r1, 32, sm, gather, A.px, s1, s2);
  colsum_rows(A.scratch + 132, A.scratch + 356, MP_SCR, r0, r1, 32, gP + 4512, gP + 4544);              // g6, be6 (local rows)
  // global row count: sum of the counter column
  float* cnt_s = sm + 6 * 64;                                  // survives batch_sums2's scratch use (part < 6*64)
  {
    float* c1 = s1 + 32;                                       // s1[32..63], s2[32..63] unused at F = 32
    batch_sums2(cl, A.scratch + 389, A.scratch + 389, MP_SCR, r0, r1, 1, sm, gather, A.px, c1, c1 + 1);
    if (threadIdx.x == 0) cnt_s[0] = c1[0];
    __syncthreads();
  }
  const float ninv = __fdiv_rn(1.0f, cnt_s[0]);
  __syncthreads();
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {   // gz3, then layer-3 data gradient -> gy2
    float* s = A.scratch + (long long)r * MP_SCR;
    float gz[32], gh[64], h2[64];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[96 + k], A.eps)));
      const float xh = __fmul_rn(__fsub_rn(s[96 + k], mean[96 + k]), rstd);
      gz[k] = __fmul_rn(__fmul_rn(__ldg(g6 + k), rstd),
                        __fsub_rn(__fsub_rn(s[356 + k], __fmul_rn(s1[k], ninv)), __fmul_rn(xh, __fmul_rn(s2[k], ninv))));
      s[356 + k] = gz[k];
    }
    dense_bwd_data<64, 32>(W6, gz, gh);
    bn_relu_row<64>(s + 32, mean + 32, var + 32, A.eps, g3, be3, h2);
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      s[164 + k] = h2[k];
      const float gy = h2[k] > 0.f ? gh[k] : 0.f;
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[32 + k], A.eps)));
      const float xh = __fmul_rn(__fsub_rn(s[32 + k], mean[32 + k]), rstd);
      s[292 + k] = gy;                                         // gy2
      gh[k] = __fmul_rn(gy, xh);
    }
    // gy2 * xhat2 parked in the Z3 columns?  no: Z3 is still needed by nobody after gz3 -> reuse 96..127 + 228.. is H3
    // (64 values: first 32 into the dead Z3 columns, last 32 into the dead "gy3 * xhat3" columns 132..163)
#pragma unroll
    for (int k = 0; k < 32; ++k) { s[96 + k] = gh[k]; s[132 + k] = gh[32 + k]; }
  }
  __syncthreads();
  wgrad_rows(A.scratch + 356, MP_SCR, A.scratch + 164, MP_SCR, r0, r1, 64, 32, gP + 2432, gP + 4480);   // W6, b6
  // BN2 sums over 64 features: gy2 at 292..355; gy2 * xhat2 split over two column groups -> two calls of 32
  batch_sums2(cl, A.scratch + 292, A.scratch + 96, MP_SCR, r0, r1, 32, sm, gather, A.px, s1, s2);
  colsum_rows(A.scratch + 96, A.scratch + 292, MP_SCR, r0, r1, 32, gP + 2304, gP + 2368);
  {
    float* s1b = s1 + 32;
    float* s2b = s2 + 32;
    batch_sums2(cl, A.scratch + 324, A.scratch + 132, MP_SCR, r0, r1, 32, sm, gather, A.px, s1b, s2b);
    colsum_rows(A.scratch + 132, A.scratch + 324, MP_SCR, r0, r1, 32, gP + 2304 + 32, gP + 2368 + 32);
  }
  __syncthreads();                                             // column sums done before the columns are rewritten
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {   // gz2, layer-2 data gradient -> gy1
    float* s = A.scratch + (long long)r * MP_SCR;
    float gz[64], gh[32], h1[32];
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[32 + k], A.eps)));
      const float xh = __fmul_rn(__fsub_rn(s[32 + k], mean[32 + k]), rstd);
      gz[k] = __fmul_rn(__fmul_rn(__ldg(g3 + k), rstd),
                        __fsub_rn(__fsub_rn(s[292 + k], __fmul_rn(s1[k], ninv)), __fmul_rn(xh, __fmul_rn(s2[k], ninv))));
      s[292 + k] = gz[k];
    }
    dense_bwd_data<32, 64>(W3, gz, gh);
    bn_relu_row<32>(s, mean, var, A.eps, g0, be0, h1);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      s[132 + k] = h1[k];
      const float gy = h1[k] > 0.f ? gh[k] : 0.f;
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[k], A.eps)));
      const float xh = __fmul_rn(__fsub_rn(s[k], mean[k]), rstd);
      s[260 + k] = gy;                                         // gy1
      s[96 + k] = __fmul_rn(gy, xh);                           // gy1 * xhat1
    }
  }
  __syncthreads();
  wgrad_rows(A.scratch + 292, MP_SCR, A.scratch + 132, MP_SCR, r0, r1, 32, 64, gP + 192, gP + 2240);    // W3, b3
  batch_sums2(cl, A.scratch + 260, A.scratch + 96, MP_SCR, r0, r1, 32, sm, gather, A.px, s1, s2);
  colsum_rows(A.scratch + 96, A.scratch + 260, MP_SCR, r0, r1, 32, gP + 128, gP + 160);
  __syncthreads();
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {   // gz1, input gradient
    float* s = A.scratch + (long long)r * MP_SCR;
    float gz[32], gf[3];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[k], A.eps)));
      const float xh = __fmul_rn(__fsub_rn(s[k], mean[k]), rstd);
      gz[k] = __fmul_rn(__fmul_rn(__ldg(g0 + k), rstd),
                        __fsub_rn(__fsub_rn(s[260 + k], __fmul_rn(s1[k], ninv)), __fmul_rn(xh, __fmul_rn(s2[k], ninv))));
      s[260 + k] = gz[k];
    }
    dense_bwd_data<3, 32>(W0, gz, gf);
    const float craw = A.c[r];
    const float c = fminf(fmaxf(craw, 0.f), 1.f);
    s[128] = c; s[129] = __fmul_rn(c, c); s[130] = log1p_f64(c);
    // d/dc [c, c^2, log1p c] and the clamp's pass-through inside [0, 1]
    const float g = __fadd_rn(__fadd_rn(gf[0], __fmul_rn(gf[1], __fmul_rn(2.f, c))), __fdiv_rn(gf[2], __fadd_rn(1.f, c)));
    if (A.gc) A.gc[r] = (craw >= 0.f && craw <= 1.f) ? g : 0.f;
  }
  __syncthreads();
  wgrad_rows(A.scratch + 260, MP_SCR, A.scratch + 128, MP_SCR, r0, r1, 3, 32, gP, gP + 96);             // W0, b0
  cl.sync();
}

// =================================================================================================
// soft mask: backward of m = smooth5x5(nearest_up(softmax(net([bits_norm, act]))[0]))  (quantization.py:213-239)
// =================================================================================================
// one CTA per image.  smem: dmt[nt] | gin[nt] | gw[170]
__global__ void __launch_bounds__(256)
softmask_bwd_kernel(const float* __restrict__ dm, const float* __restrict__ bit_map, const float* __restrict__ act_n,
                    const float* __restrict__ P, int H, int W, int Ht, int Wt, float* __restrict__ dbit,
                    float* __restrict__ gP) {
  extern __shared__ float sm[];
  const int nt = Ht * Wt, b = blockIdx.x;
  float* dmt = sm;
  float* gin = dmt + nt;
  float* gw = gin + nt;
  const float* W0 = P; const float* b0 = P + 144; const float* W2 = P + 152; const float* b2 = P + 168; const float* ks = P + 170;
  for (int i = threadIdx.x; i < 2 * nt + 170; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  // (1) smoothing (replicate padding) + nearest upsampling, transposed: pixel gradients scattered to tiles
  const float sy = (float)Ht / (float)H, sx = (float)Wt / (float)W;
  const float* g = dm + (long long)b * H * W;
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    const int h = p / W, w = p - h * W;
    const float gp = g[p];
    if (gp == 0.f) continue;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int iy = nearest_src(min(max(h + ky - 2, 0), H - 1), sy, Ht);
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int ix = nearest_src(min(max(w + kx - 2, 0), W - 1), sx, Wt);
        atomicAdd(dmt + iy * Wt + ix, __fmul_rn(gp, __ldg(ks + ky * 5 + kx)));
      }
    }
  }
  __syncthreads();
  // (2) tile head backward
  const float* bm = bit_map + (long long)b * nt;
  const float* an = act_n + (long long)b * nt;
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const int i = t / Wt, j = t - i * Wt;
    float in[2][9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = i + ky - 1, xx = j + kx - 1;
        const bool ok = yy >= 0 && yy < Ht && xx >= 0 && xx < Wt;
        in[0][ky * 3 + kx] = ok ? fminf(fmaxf(__fdiv_rn(__fsub_rn(bm[yy * Wt + xx], 2.0f), 6.0f), 0.f), 1.f) : 0.f;
        in[1][ky * 3 + kx] = ok ? an[yy * Wt + xx] : 0.f;
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
    if (gl == 0.f) continue;
    atomicAdd(gw + 168, gl);
    atomicAdd(gw + 169, -gl);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      atomicAdd(gw + 152 + o, __fmul_rn(gl, hid[o]));
      atomicAdd(gw + 160 + o, -__fmul_rn(gl, hid[o]));
      const float gh = hid[o] > 0.f ? __fmul_rn(gl, __fsub_rn(__ldg(W2 + o), __ldg(W2 + 8 + o))) : 0.f;
      if (gh == 0.f) continue;
      atomicAdd(gw + 144 + o, gh);
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        atomicAdd(gw + (o * 2 + 0) * 9 + k, __fmul_rn(gh, in[0][k]));
        atomicAdd(gw + (o * 2 + 1) * 9 + k, __fmul_rn(gh, in[1][k]));
        const int yy = i + k / 3 - 1, xx = j + k % 3 - 1;
        if (yy >= 0 && yy < Ht && xx >= 0 && xx < Wt) atomicAdd(gin + yy * Wt + xx, __fmul_rn(gh, __ldg(W0 + (o * 2) * 9 + k)));
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const float bn = __fdiv_rn(__fsub_rn(bm[t], 2.0f), 6.0f);
    dbit[(long long)b * nt + t] = (bn >= 0.f && bn <= 1.f) ? __fdiv_rn(gin[t], 6.0f) : 0.f;
  }
  for (int i = threadIdx.x; i < 170; i += blockDim.x)
    if (gw[i] != 0.f) atomicAdd(gP + i, gw[i]);
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

}  // namespace mcaq

using namespace mcaq;

extern "C" long long mcaq_cmlp_train_scratch_floats(int N) { return N > 0 ? (long long)N * CM_SCR : MCAQ_EINVAL; }
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
  const int rpc = 128;
  cmlp_bwd_kernel<<<(N + rpc - 1) / rpc, 128, 0, st>>>(phi, grad_craw_ws, params, N, rpc, scratch, grad_params);
  MCAQ_LAUNCH_CHECK();
  return 0;
}

static int launch_mapper(void (*k)(const MapArgs), MapArgs& A, cudaStream_t st) {
  const size_t smem = (6 * 64 + 8 + MAP_CL * (2 * 64 + 1) + 128) * sizeof(float);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(MAP_CL);
  cfg.blockDim = dim3(MAP_NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MAP_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  A.rpc = (A.N + MAP_CL - 1) / MAP_CL;
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
                                     float* rm1, float* rv1, float* rm2, float* rv2, float momentum, float eps, float* bits,
                                     void* const* xchg_peers, int xchg_rank, int xchg_world, void* stream) {
  if (!cmap || !params || !scratch || !stats || !bits || N <= 0) return MCAQ_EINVAL;
  MapArgs A = {};
  A.c = cmap; A.P = params; A.N = N; A.temperature = temperature; A.use_t = use_temperature; A.lo = min_bits; A.hi = max_bits;
  A.scratch = scratch; A.stats = stats; A.rm[0] = rm0; A.rv[0] = rv0; A.rm[1] = rm1; A.rv[1] = rv1; A.rm[2] = rm2; A.rv[2] = rv2;
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
  const size_t smem = (2 * (size_t)Ht * Wt + 170) * sizeof(float);
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
