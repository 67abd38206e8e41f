/*
 * mcaq_b200.h -- C ABI of libmcaq_b200.so: the B200-native (sm_100a) implementation of the
 * MCAQ-YOLO hot path (tile-wise spatial adaptive quantization + morphological complexity).
 *
 * Every entry point takes plain device pointers and sizes, launches asynchronously on the
 * caller's stream (a cudaStream_t passed as void*), never synchronises and never allocates,
 * so every call is CUDA-graph capturable.  Return value: 0 on success, a positive
 * cudaError_t when a launch failed, a negative MCAQ_E* code for bad arguments.
 *
 * Reference interfaces these replace (paths under the reference's mcaq_yolo/):
 *   ops/src/mcaq_kernel.cu:102-123   launch_spatial_quantization      (kept, same signature)
 *   ops/src/mcaq_ops.cpp:22-77       mcaq_cuda_ops.spatial_quantize   (Python shim on top)
 *   core/quantization.py:319-353,409-434,636-746   ranges / EMA / compose / STE backward
 *   core/morphology.py:826-873,939-973             phi tiles, complexity MLP, bilateral
 *   core/bit_allocation.py:42-80,218-280           linear / MLP bit mappers (eval)
 *   core/quantization.py:213-239                   learned soft mask
 */
#ifndef MCAQ_B200_H
#define MCAQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCAQ_ABI_VERSION 1

#if defined(__GNUC__)
#define MCAQ_API __attribute__((visibility("default")))
#else
#define MCAQ_API
#endif

/* element type of feature maps */
#define MCAQ_F32  0
#define MCAQ_BF16 1
#define MCAQ_F16  2   /* IEEE half: what the hooked backbone outputs are under torch.autocast (train.py:192, 582, 748) */

/* error codes (negative) */
#define MCAQ_EINVAL   (-1)   /* bad size / null pointer */
#define MCAQ_EALIGN   (-2)   /* pointer not 16-byte aligned */
#define MCAQ_ETOOBIG  (-3)   /* plane does not fit the on-chip budget of the morphology kernel */
#define MCAQ_EDTYPE   (-4)
#define MCAQ_EGEOM    (-5)   /* geometry / alignment outside the vector path of an entry point that has no scalar form */

/* number of floats in the constant block expected by the morphology kernels, and the
 * packed-parameter block sizes of the three small networks (see mcaq_b200/constants.py) */
#define MCAQ_CONSTS_FLOATS     192
#define MCAQ_CMLP_FLOATS       2884   /* complexity_mlp 8-64-LN-32-LN-1: W0^T[8][64] b0 g1 be1 W3^T[64][32] b3 g4 be4 W6[32] b6 pad3 */
#define MCAQ_MAPPER_FLOATS     4612   /* mapping_network, BN folded: W0^T[3][32] (b,alpha,beta)[32] W3^T[32][64] (..)[64] W6^T[64][32] (..)[32] W9[32] b9 pad3 */
#define MCAQ_MAPPER_STEPS_FLOATS 12   /* step table that may follow the mapper block: 8 steps, valid, temperature, lo, hi */
#define MCAQ_SOFTMASK_FLOATS   196    /* conv3x3(2->8)+b, conv1x1(8->2)+b, 5x5 smooth, pad1; all three blocks 16-byte aligned */

MCAQ_API int mcaq_abi_version(void);
MCAQ_API const char* mcaq_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * K1  one HBM sweep of x (NCHW, contiguous):
 *     sum_plane[b,h,w] = sum_c x        (torch CPU cascade order, see oracle.cascade_sum)
 *     abs_plane[b,h,w] = sum_c |x|
 *     keys[0..C)  = ordered-int key of min over (b,h,w) per channel   (atomicMin)
 *     keys[C..2C) = ordered-int key of max                            (atomicMax)
 * keys may be NULL (frozen calibration: ranges not needed).  Call mcaq_ranges_reset first.
 * Replaces: x.mean(1) morphology.py:837, x.abs().mean(1) quantization.py:224,
 *           x.amin/amax quantization.py:333-334, 425-426, 653-654.
 */
MCAQ_API int mcaq_ranges_reset(int32_t* keys, int C, void* stream);
MCAQ_API int mcaq_reduce_planes(const void* x, int dtype, int B, int C, int H, int W,
                       float* sum_plane, float* abs_plane, int32_t* keys, void* stream);

/* channels_last (NHWC memory: B,H,W,C) forms of K1 and K3.  Results are defined to be exactly those
 * of the NCHW kernels on the same logical tensor (same summation order, same codes).  Covered
 * geometry: C a multiple of 16 with C / (16 / sizeof(T)) a power of two <= 256 (every YOLOv8
 * width), 16-byte aligned pointers; otherwise MCAQ_EALIGN (make the tensor NCHW-contiguous). */
MCAQ_API int mcaq_reduce_planes_nhwc(const void* x, int dtype, int B, int C, int H, int W,
                                     float* sum_plane, float* abs_plane, int32_t* keys, void* stream);
MCAQ_API int mcaq_tile_quantize_ranges_nhwc(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                            const float* bit_map, int Ht, int Wt, const float* packed,
                                            const float* running_min, const float* running_max,
                                            const float* mask, void* stream);

/* keys -> packed[0..C) = min, packed[C..2C) = -max  (one MIN all-reduce merges ranks) */
MCAQ_API int mcaq_ranges_decode(const int32_t* keys, int C, float* packed, void* stream);

/* EMA update of running_min/max (quantization.py:340-347); first!=0 adopts the batch stats */
MCAQ_API int mcaq_ranges_ema(const float* packed, int C, double momentum, int first,
                    float* running_min, float* running_max, void* stream);
/* K1's epilogue for training / calibration in one launch: keys -> packed (may be NULL) AND the EMA of the running
 * statistics (quantization.py:319-353); the separate decode + ema pair remains for the all-reduced (sharded) case */
MCAQ_API int mcaq_ranges_finish(const int32_t* keys, int C, double momentum, int first, float* running_min,
                                float* running_max, float* packed, void* stream);

/* qtable[(b-2)*C + c] = {scale, zero_point} for b = 2..8 (quantization.py:41-66).
 * Ranges come either from `packed` (min, -max) or from running_min/max (packed == NULL). */
MCAQ_API int mcaq_build_qtable(const float* packed, const float* running_min, const float* running_max,
                      int C, float* qtable /* 7*C*2 floats */, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  y = m * dequant(quant_b(x)) with b = bit of the pixel's tile (nearest rule of
 *     F.interpolate), per-channel scale/zero-point from qtable.  x, y: NCHW, dtype f32/bf16;
 *     y may alias x (in place).  mask (B,H,W) fp32 or NULL.  codes (int8, same shape) or NULL.
 * Replaces: quantization.py:729-744 (_forward_pytorch, eval) and mcaq_kernel.cu:12-99.
 */
MCAQ_API int mcaq_tile_quantize(const void* x, void* y, int dtype, int B, int C, int H, int W,
                       const float* bit_map, int Ht, int Wt, const float* qtable,
                       const float* mask, int8_t* codes, void* stream);

/* Same as mcaq_tile_quantize with the per-channel ranges given directly (packed = [min, -max]
 * as written by mcaq_ranges_decode / mcaq_morph_fused, or running_min/running_max when packed is
 * NULL): each CTA derives its {scale, zero_point} rows itself, so no table kernel is launched.
 * qtable_ws (7*C*2 floats) is only used for geometries the vector kernel does not cover. */
MCAQ_API int mcaq_tile_quantize_ranges(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                       const float* bit_map, int Ht, int Wt, const float* packed,
                                       const float* running_min, const float* running_max,
                                       float* qtable_ws, const float* mask, void* stream);
/* Bulk-copy (TMA) staged variant of mcaq_tile_quantize_ranges: the input tile travels global -> shared memory with
 * cp.async.bulk + mbarrier (csrc/tile_quantize_tma.cu), identical results.  Covers the vector geometries with
 * C % 32 == 0 and y != x; MCAQ_EGEOM otherwise (the caller then uses mcaq_tile_quantize_ranges). */
MCAQ_API int mcaq_tile_quantize_ranges_tma(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                           const float* bit_map, int Ht, int Wt, const float* packed,
                                           const float* running_min, const float* running_max, const float* mask,
                                           void* stream);

/* Training forward (fractional bits, quantization.py:699-727, 742-744):
 *   pre = (1-f) Q_floor(b)(x) + f Q_floor(b)+1(x),  y = pre * m                              */
MCAQ_API int mcaq_tile_quantize_train_fwd(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                 const float* bit_map, int Ht, int Wt, const float* qtable,
                                 const float* mask, void* stream);

/* Training backward (STE, quantization.py:101-118 + autograd of the compose):
 *   dx   = g*m*(1-f) + g*m*f
 *   dbit[b,ty,tx] += sum_{c, pixels of tile} g*m*(Q_hi - Q_lo)     (dbit must be zeroed)
 *   dmask[b,h,w]   = sum_c g*pre                                    (NULL when no mask)     */
MCAQ_API int mcaq_tile_quantize_train_bwd(const void* grad_y, const void* x, void* grad_x, int dtype,
                                 int B, int C, int H, int W,
                                 const float* bit_map, int Ht, int Wt, const float* qtable,
                                 const float* mask, float* dbit, float* dmask, void* stream);

/* Training forward / backward with the feature-level distillation term folded in (SURVEY 8f-3;
 * train.py:599-610: F.mse_loss(features_q.float(), teacher.float()) per hooked layer):
 *   forward : *kd_sum += sum (y - teacher)^2  (one fp64 atomic per CTA; caller zeroes it and divides by numel)
 *   backward: g_t = grad_y + (*kd_coef) * (y - teacher), y recomputed from x, then as the plain backward;
 *             kd_coef is a DEVICE scalar = dL/d(mse) * 2 / numel, so nothing synchronises.
 * teacher is fp32 (B,C,H,W) whatever the dtype of x (the reference's teacher runs in fp32).
 * Vector geometry only (16-byte aligned pointers, H*W % VEC == 0, W % 4 == 0, W % Wt == 0,
 * (W/Wt) % 4 == 0: every YOLO feature map): otherwise MCAQ_EGEOM and the caller composes the plain
 * entry points with its own MSE. */
MCAQ_API int mcaq_tile_quantize_train_fwd_kd(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                 const float* bit_map, int Ht, int Wt, const float* qtable,
                                 const float* mask, const float* teacher, double* kd_sum, void* stream);
MCAQ_API int mcaq_tile_quantize_train_bwd_kd(const void* grad_y, const void* x, void* grad_x, int dtype,
                                 int B, int C, int H, int W,
                                 const float* bit_map, int Ht, int Wt, const float* qtable,
                                 const float* mask, const float* teacher, const float* kd_coef,
                                 float* dbit, float* dmask, void* stream);

/* Level 0: the reference's launcher, same symbol and argument list (ops/src/mcaq_kernel.cu:102-111,
 * engine/MCAQPlugin.cpp:15-24).  It returns void there; here the outcome of a thread's last call is
 * kept in mcaq_level0_status() (0 / cudaError_t / negative MCAQ_E*), and mcaq_spatial_quantization is the
 * same call returning that code.  tile_h / tile_w must be H / n_tiles_h and W / n_tiles_w, what the
 * reference's caller passes (quantization.py:641-642); anything else is MCAQ_EINVAL.  For 16-byte friendly
 * geometries -- every YOLO feature map -- this IS the vector kernel of the fused path (K3, ranges given
 * directly); others take a scalar kernel.  Tile rule and rounding follow the reference's PyTorch path
 * (F.interpolate nearest, half-to-even, IEEE division) -- the parity bar -- which coincides with the reference
 * kernel's h / tile_h whenever the tile grid divides the map. */
MCAQ_API void launch_spatial_quantization(const float* input, const float* bit_map,
                                 const float* min_vals, const float* max_vals,
                                 const float* mask, float* output,
                                 int N, int C, int H, int W, int tile_h, int tile_w,
                                 int n_tiles_h, int n_tiles_w, void* stream);
MCAQ_API int mcaq_spatial_quantization(const float* input, const float* bit_map,
                                 const float* min_vals, const float* max_vals,
                                 const float* mask, float* output,
                                 int N, int C, int H, int W, int tile_h, int tile_w,
                                 int n_tiles_h, int n_tiles_w, void* stream);
MCAQ_API int mcaq_level0_status(void);

/* ------------------------------------------------------------------------------------------
 * K2  per-image morphology on the on-chip gray plane (one CTA per image).
 *   in : sum_plane (B,H,W) from K1, C, grid_size (consts: reserved, may be NULL -- the stencils
 *        are compiled in from csrc/mcaq_consts.cuh)
 *   out: phi (B,ht,wt,8) fp32                                     morphology.py:826-873
 *   optional debug outputs (NULL to skip): gray (B,Hc,Wc) fp32, edge_bits / bin_bits
 *   (B,Hc,ceil(Wc/32)) uint32 bit planes, lbp_hist (B,ht,wt,10) int32, counts
 *   (B,ht,wt,12) int32 = {edge, area, perim, euler_x4, N_2, N_4, N_8, N_16, N_32, otsu_bin,0,0}.
 */
MCAQ_API int mcaq_morph_phi(const float* sum_plane, int B, int C, int H, int W, int grid_size,
                   const float* consts, float* phi,
                   float* gray_dbg, uint32_t* edge_bits_dbg, uint32_t* bin_bits_dbg,
                   int32_t* lbp_hist_dbg, int32_t* counts_dbg, void* stream);

/* Image-sized planes (curriculum scoring on raw 640x640 / 1280x1280 images: utils/dataset.py:345-353 ->
 * core/morphology.py:923-937; tests/test_smoke.py:33-47 parametrises H=640): the fused kernel above keeps a
 * plane of at most 160 columns / tiles of at most 32 pixels on chip.  mcaq_morph_fits tells which path a
 * geometry takes; mcaq_morph_phi_planes computes the same phi with a five-launch pipeline over L2-resident
 * planes (min/max, stage A, Otsu, stage B; csrc/morph_planes.cu), tiles up to 128 pixels, any plane size,
 * in a caller-owned workspace of mcaq_morph_planes_workspace bytes (16-byte aligned).  counts_dbg there is
 * (B,ht,wt,12): edge, area, perimeter, 4*Euler, N_2..N_128 (7 slots), Otsu bin. */
MCAQ_API int mcaq_morph_fits(int B, int C, int H, int W, int grid_size);
MCAQ_API long long mcaq_morph_planes_workspace(int B, int C, int H, int W, int grid_size);
MCAQ_API int mcaq_morph_phi_planes(const float* sum_plane, int B, int C, int H, int W, int grid_size,
                                   void* workspace, long long workspace_bytes, float* phi, float* gray_dbg,
                                   uint32_t* edge_bits_dbg, uint32_t* bin_bits_dbg, int32_t* lbp_hist_dbg,
                                   int32_t* counts_dbg, void* stream);

/* K2 fused: everything between the two HBM sweeps of the inference hook in ONE launch
 * (models/mcaq_yolo.py:426-447): phi -> complexity (MLP, bilateral) -> bit map (MLP mapper, or
 * LinearBitMapper when linear_mapper != 0) -> soft mask m (B,H,W) (skipped when softmask NULL).
 * linear_mapper == 2: `mapper` is the MLP block followed by its step table (mcaq_mapper_steps, built
 * for the same temperature / min_bits / max_bits); integer (continuous == 0) bit maps are then read off
 * the staircase instead of re-evaluating the network per tile.
 * If keys != NULL, CTA 0 also decodes K1's range keys into packed_ranges ([min, -max], 2C floats)
 * and re-arms the keys for the next sweep.  phi may be NULL. */
MCAQ_API int mcaq_morph_fused(const float* sum_plane, const float* abs_plane, int B, int C, int H, int W,
                              int grid_size, int32_t* keys, float* packed_ranges, const float* cmlp,
                              const float* mapper, int linear_mapper, const float* softmask,
                              float temperature, int use_temperature, int continuous, float min_bits,
                              float max_bits, float eps_spread, float* phi, float* complexity,
                              float* bit_map, float* mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU, batch sharded): per-channel range merge over peer memory.
 * The reference has no data parallelism; parity with its single-process batch needs min/max over
 * the whole batch (quantization.py:423-426, 650-654).  Each rank owns one exchange buffer per scale
 * (mcaq_xchg_bytes) that every other rank of the node maps through CUDA IPC.  K2's first CTA
 * (mcaq_morph_fused_xchg) stores this rank's [min, -max] into every rank's buffer and releases the
 * step number when the kernel starts, and waits for all ranks / writes the minimum into
 * packed_ranges when it ends; K3 then runs unchanged on packed_ranges.  No collective launch, no
 * host synchronisation.  mcaq_tile_quantize_xchg is the host-driven form (one-CTA merge kernel +
 * K3) for callers that publish with mcaq_xchg_publish.
 */
MCAQ_API long long mcaq_xchg_bytes(int C, int world);
MCAQ_API int mcaq_xchg_alloc(long long bytes, void** out);           /* cudaMalloc + zero */
MCAQ_API int mcaq_xchg_free(void* p);
MCAQ_API int mcaq_xchg_export(void* p, void* handle64);              /* cudaIpcGetMemHandle */
MCAQ_API int mcaq_xchg_open(const void* handle64, void** out);       /* cudaIpcOpenMemHandle */
MCAQ_API int mcaq_xchg_close(void* p);
/* Waits are bounded (default 2 s per wait): a peer that died or ran a different sequence of exchange
 * steps makes the waiting launch keep THIS rank's own ranges and record the step in the buffer's error
 * word instead of spinning forever.  mcaq_xchg_error synchronises, returns the first such step (0 = none)
 * in *step and clears the word. */
MCAQ_API void mcaq_xchg_set_timeout_ms(int ms);
MCAQ_API int mcaq_xchg_error(void* local, int* step);
/* stand-alone halves of the protocol (tests, host-driven use): publish `packed` as this rank's
 * next step / wait for all ranks and write the merged vector */
MCAQ_API int mcaq_xchg_publish(void* const* peers, int rank, int world, const float* packed, int C, void* stream);
MCAQ_API int mcaq_xchg_merge(const void* local, int world, int C, float* packed, void* stream);

MCAQ_API int mcaq_morph_fused_xchg(const float* sum_plane, const float* abs_plane, int B, int C, int H, int W,
                                   int grid_size, int32_t* keys, float* packed_ranges, const float* cmlp,
                                   const float* mapper, int linear_mapper, const float* softmask,
                                   float temperature, int use_temperature, int continuous, float min_bits,
                                   float max_bits, float eps_spread, float* phi, float* complexity,
                                   float* bit_map, float* mask, void* const* xchg_peers, int xchg_rank,
                                   int xchg_world, void* stream);

MCAQ_API int mcaq_tile_quantize_xchg(const void* x, void* y, int dtype, int B, int C, int H, int W,
                                     const float* bit_map, int Ht, int Wt, const void* xchg_local, int world,
                                     float* qtable_ws, float* packed_ws, const float* mask, void* stream);

/* Step table of the MLP bit mapper for integer outputs (bit_allocation.py:218-280 in eval mode): the
 * mapper is monotone in the scalar complexity c (Eq.18), so its rounded output is a staircase.
 * steps[k], k < 8 = smallest c in [0,1] with bits(c) >= min_bits + 1 + k (0 always, +inf never), found by
 * bisection with the mapper kernel itself; steps[8] = 1 if the table is monotone (else the fused kernel
 * evaluates the network), steps[9..11] = temperature, min_bits, max_bits it was built for.
 * One small launch; rebuild when the weights, the temperature or the bit range change. */
MCAQ_API int mcaq_mapper_steps(const float* mapper, float temperature, int use_temperature, float min_bits,
                               float max_bits, float* steps /* MCAQ_MAPPER_STEPS_FLOATS */, void* stream);

/* phi -> complexity (MLP + LayerNorm + sigmoid, 5x5 bilateral, clamp)  morphology.py:959-968 */
MCAQ_API int mcaq_complexity(const float* phi, int B, int ht, int wt, const float* cmlp, const float* consts,
                    float* complexity_raw /* nullable */, float* complexity, void* stream);

/* complexity -> bit map, eval semantics (integer bits unless continuous != 0).
 * mapper == NULL selects LinearBitMapper (2/98 % quantiles).  bit_allocation.py:42-80, 218-280 */
MCAQ_API int mcaq_bit_mapper(const float* complexity, int B, int ht, int wt, const float* mapper,
                    float temperature, int use_temperature, int continuous,
                    float min_bits, float max_bits, float eps_spread, float* bit_map, void* stream);

/* soft mask m (B,H,W) from bit map + abs_plane  (quantization.py:213-239) */
MCAQ_API int mcaq_soft_mask(const float* bit_map, int Ht, int Wt, const float* abs_plane,
                   int B, int C, int H, int W, const float* softmask, float* mask_tiles /* nullable */,
                   float* mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training forms of the three tile-level networks (csrc/train_nets.cu).  The reference trains them through
 * ~10^3 eager autograd kernels per step (core/morphology.py:81-97, 309-354; core/bit_allocation.py:218-280 with
 * train-mode BatchNorm1d :126; core/quantization.py:213-239); here a network is one or two launches per
 * direction.  Parameter / gradient blocks are the flat concatenation of the torch parameters in module order
 * (complexity MLP 2881, mapper 4609, soft mask 170 floats); gradient blocks are ACCUMULATED into.
 *   complexity: forward = mcaq_complexity (keep its raw output); mcaq_complexity_train_bwd = bilateral + clamp +
 *               MLP backward (scratch: mcaq_cmlp_train_scratch_floats(B*ht*wt) floats, grad_craw_ws: B*ht*wt)
 *   mapper    : mcaq_mapper_train_fwd / _bwd run as ONE thread-block cluster (16 CTAs, or the portable 8); BatchNorm batch statistics
 *               are reduced through distributed shared memory and -- xchg_world > 1, buffers of a C = 128 range
 *               exchange -- merged with the other ranks of the node over NVLink peer memory inside the kernel
 *               (rank-ordered Chan merge: every rank gets the statistics of the unsharded batch).
 *               scratch: mcaq_mapper_train_scratch_floats(N) floats shared by forward and backward; stats: 256.
 *   soft mask : forward = mcaq_soft_mask; mcaq_softmask_act gives the normalised tile activity the backward needs
 *   losses    : mcaq_bit_stats adds sum(bit_map) and its total variation into out2 (Lbit / Lsmooth,
 *               models/mcaq_yolo.py:86-118, 575) and, given weights2, writes d(w0 sum + w1 TV)/d bit_map */
MCAQ_API long long mcaq_cmlp_train_scratch_floats(int N);
MCAQ_API long long mcaq_mapper_train_scratch_floats(int N);
MCAQ_API int mcaq_complexity_train_bwd(const float* phi, const float* craw, const float* grad_out, int B, int ht, int wt,
                                       const float* params, float* scratch, float* grad_craw_ws, float* grad_params,
                                       void* stream);
MCAQ_API int mcaq_mapper_train_fwd(const float* cmap, int N, const float* params, float temperature, int use_temperature,
                                   float min_bits, float max_bits, float* scratch, float* stats, float* rm0, float* rv0,
                                   float* rm1, float* rv1, float* rm2, float* rv2, long long* nbt0, long long* nbt1,
                                   long long* nbt2, float momentum, float eps, float* bits,
                                   void* const* xchg_peers, int xchg_rank, int xchg_world, void* stream);
MCAQ_API int mcaq_mapper_train_bwd(const float* cmap, int N, const float* params, float temperature, int use_temperature,
                                   float min_bits, float max_bits, float* scratch, const float* stats, float eps,
                                   const float* grad_bits, float* grad_c, float* grad_params, void* const* xchg_peers,
                                   int xchg_rank, int xchg_world, void* stream);
MCAQ_API int mcaq_softmask_act(const float* abs_plane, int B, int C, int H, int W, int Ht, int Wt, float* act_norm,
                               void* stream);
MCAQ_API int mcaq_softmask_train_bwd(const float* grad_mask, const float* bit_map, const float* act_norm,
                                     const float* params, int B, int H, int W, int Ht, int Wt, float* grad_bit_map,
                                     float* grad_params, void* stream);
MCAQ_API int mcaq_bit_stats(const float* bit_map, int B, int ht, int wt, float* out2, const float* weights2,
                            float* grad_bit_map, void* stream);

/* avg_bits = mean over scales of the per-scale mean, Lbit = (avg_bits - target)^2, Lsmooth = mean over scales of
 * TV / edge count of S <= 4 bit maps (B_s, ht_s, wt_s), exactly as MCAQYOLO.forward / MCAQYOLOLoss form them
 * (models/mcaq_yolo.py:575, 86-118), in ONE launch: out[0..2] = the three scalars, out[3 + 2 s], out[4 + 2 s] = the
 * per-scale sum and total variation.  With grad_out (3 floats on the device: d/d avg_bits, d/d Lbit, d/d Lsmooth) the
 * call is the backward instead: grads[s] receives d/d bit_map_s (out must hold the forward's values).  The pointer
 * arrays are HOST arrays of device pointers. */
MCAQ_API int mcaq_bit_losses(const float* const* bit_maps, const int* B, const int* ht, const int* wt, int S, float target,
                             float* out, const float* grad_out, float* const* grads, void* stream);

/* Self test of the quantiser's division (RN(x/scale) by Markstein's correction with RN(1/scale))
 * against div.rn: sweeps numerators with bit patterns first + i*stride, i < count, for every
 * scale; *mismatches must be zeroed by the caller. */
MCAQ_API int mcaq_selftest_division(const float* scales, int nscales, unsigned first, unsigned stride, /* mismatches[2]: count, one example */
                                    unsigned long long count, unsigned long long* mismatches, void* stream);

/* debug: device buffer of 16 clock64() stamps per image written by mcaq_morph_phi (NULL disables) */
MCAQ_API void mcaq_debug_stage_clocks(long long* dev_buf);

/* debug / tuning: force the number of CTAs (cluster size 1, 2, 4 or 8) an image is split over in
 * the morphology kernel; 0 = automatic */
MCAQ_API void mcaq_debug_cluster_split(int ns);

/* debug / tuning: force the CTA size (threads, multiple of 32, 64..512) of the morphology kernel; 0 = automatic */
MCAQ_API void mcaq_debug_morph_threads(int n);

/* debug / tests: 1 = route the training forward / backward through the scalar kernels even when the
 * vector path applies (the two must agree: y and dx bit for bit) */
MCAQ_API void mcaq_debug_train_scalar(int on);
/* Tuning aid: channels per CTA of the inference quantise sweep (8 / 16; 0 = the built-in choice). */
MCAQ_API void mcaq_debug_k3_chunk(int ch);
/* Tuning aid: cluster size of the train-mode mapper kernels (8 or 16; 0 = probe the device once). */
MCAQ_API void mcaq_debug_mapper_cluster(int n);

/* split policy of the morphology kernel: 0 (default) throughput -- one CTA per image unless the batch
 * is too small to fill half the GPU (for callers that keep several launches in flight); 1 latency --
 * every image split over a thread-block cluster, two CTAs per SM (serial callers: the forward hook) */
MCAQ_API void mcaq_morph_policy(int latency);

/* tile size rule of the analyzer (morphology.py:359-376) */
MCAQ_API int mcaq_tile_size(int H, int grid_size);

#ifdef __cplusplus
}
#endif
#endif /* MCAQ_B200_H */
