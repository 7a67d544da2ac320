/*
 * rfdnet_b200.h -- C ABI of librfdnet_b200.so: the B200 (sm_100a) implementation of RfD-Net's
 * point-cloud hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all tensors are dense, contiguous, row-major in the shape given in the comment;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call is
 *     asynchronous on that stream and never synchronises the host;
 *   - outputs are fully written by the callee (the reference's wrappers rely on torch::zeros;
 *     here the kernels write every element, or the entry point memsets first where the
 *     reference semantics need a zero background);
 *   - return value: RFD_OK (0) or a negative RFD_ERR_* code; rfd_status_string() explains it,
 *     rfd_last_error() returns the CUDA error text of the last failing call on this thread.
 *     Unlike the reference (cuda_utils.h:30-39: fprintf + exit(-1)) nothing here terminates
 *     the process.
 *
 * Reference interface each entry point replaces (paths relative to
 * /root/reference/external/pointnet2_ops_lib/pointnet2_ops/_ext-src/):
 *   rfd_furthest_point_sampling  <- furthest_point_sampling   include/sampling.h:6,  src/sampling.cpp:66-87
 *   rfd_gather_points            <- gather_points             include/sampling.h:4,  src/sampling.cpp:15-38
 *   rfd_gather_points_grad       <- gather_points_grad        include/sampling.h:5,  src/sampling.cpp:40-65
 *   rfd_ball_query               <- ball_query                include/ball_query.h:4-5, src/ball_query.cpp:8-32
 *   rfd_group_points             <- group_points              include/group_points.h:4, src/group_points.cpp:12-36
 *   rfd_group_points_grad        <- group_points_grad         include/group_points.h:5, src/group_points.cpp:38-62
 *   rfd_three_nn                 <- three_nn                  include/interpolate.h:6,  src/interpolate.cpp:14-40
 *   rfd_three_interpolate        <- three_interpolate         include/interpolate.h:7-8, src/interpolate.cpp:42-70
 *   rfd_three_interpolate_grad   <- three_interpolate_grad    include/interpolate.h:9-10, src/interpolate.cpp:71-100
 *   rfd_query_and_group          <- QueryAndGroup.forward     ../pointnet2_utils.py:302-361 (ball_query + 2x group_points + sub + div + cat)
 *   rfd_pointwise_mlp_f32        <- Conv{1,2}d(1x1)+BatchNorm(eval)+ReLU(+max_pool2d)  ../pointnet2_modules.py:9-19,237-243
 *   rfd_mlp_chain, rfd_sa_mlp_chain <- the same stacks per module on tensor cores (see below)
 *   rfd_three_nn_interpolate     <- PointnetFPModule.forward 3-NN + weights + interpolate  ../pointnet2_modules.py:381-389
 *   rfd_onet_*                   <- DecoderCBatchNorm.forward /root/reference/models/iscnet/modules/occ_decoder.py:110-122
 *                                   (+ layers.py:98-107 CResnetBlockConv1d, :226-242 CBatchNorm1d), eval mode
 *   rfd_make_3d_grid             <- make_3d_grid              /root/reference/external/common.py:157-176 (x box_size, generator.py:92-95)
 */
#ifndef RFDNET_B200_H
#define RFDNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFD_OK 0
#define RFD_ERR_INVALID_ARGUMENT (-1) /* null pointer, negative size, unsupported combination */
#define RFD_ERR_UNSUPPORTED_SIZE (-2) /* size outside what the kernels were built for (see each call) */
#define RFD_ERR_CUDA (-3)             /* a CUDA runtime call or launch failed; see rfd_last_error() */
#define RFD_ERR_NO_DEVICE (-4)        /* no sm_100 device / wrong architecture */

#define RFD_ABI_VERSION 4

int rfd_abi_version(void);
const char *rfd_status_string(int status);
const char *rfd_last_error(void);
/* sm count, compute capability of the current device; RFD_ERR_NO_DEVICE if it is not compute 10.x */
int rfd_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* number of kernels launched by this library in this process since load (bench.py's gpu_launches) */
long long rfd_launch_count(void);

/* ---- (a1) furthest point sampling: xyz (B,N,3) f32 -> idx (B,m) i32.  N <= 196608 per scene.
 * Bit-exact with the reference kernel incl. its tie-break and the |p|^2 <= 1e-3 skip rule. */
int rfd_furthest_point_sampling(const float *xyz, int B, int N, int m, int *idx, void *stream);
/* same, additionally emitting the sampled coordinates new_xyz (B,m,3) f32 (NULL = skip): replaces the
 * transpose + gather_points + transpose of PointnetSAModuleVotes.forward (../pointnet2_modules.py:219-226). */
int rfd_furthest_point_sampling_xyz(const float *xyz, int B, int N, int m, int *idx, float *new_xyz, void *stream);

/* FPS of a cloud that is itself in FPS order (SA(k+1) samples SA(k)'s samples, pointnet2backbone.py:104-113) returns
 * 0..m-1.  rfd_fps_prefix_check PROVES per scene whether rfd_furthest_point_sampling(xyz)[0..m) == (0,1,...,m-1), with
 * the sampler's own arithmetic and tie-break, as one parallel sweep instead of m-1 serial rounds:
 *   flag (B) i32 = 1 iff proved; own_ws: scratch, B*m floats.  m <= 3072 (larger: flag = 0).
 * rfd_furthest_point_sampling_cond is the sampler with that flag (NULL = always sample): scenes whose flag is non-zero
 * get idx = 0..m-1 (and new_xyz = xyz[0..m)) immediately, the others run the full sampler -- results are identical
 * to rfd_furthest_point_sampling_xyz in both cases. */
int rfd_fps_prefix_check(const float *xyz, int B, int N, int m, float *own_ws, int *flag, void *stream);
int rfd_furthest_point_sampling_cond(const float *xyz, int B, int N, int m, const int *prefix_flag, int *idx,
                                     float *new_xyz, void *stream);

/* ---- (a2) gather: points (B,C,N), idx (B,M) -> out (B,C,M);  grad: grad_out (B,C,M) -> grad_points (B,C,N) */
int rfd_gather_points(const float *points, const int *idx, int B, int C, int N, int M, float *out, void *stream);
int rfd_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M, float *grad_points,
                           void *stream);

/* ---- (a3) ball query: new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) i32; nsample <= 1024 */
int rfd_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius, int nsample, int *idx,
                   void *stream);

/* ---- (a4) group: points (B,C,N), idx (B,M,S) -> out (B,C,M,S);  grad: (B,C,M,S) -> (B,C,N) */
int rfd_group_points(const float *points, const int *idx, int B, int C, int N, int M, int S, float *out,
                     void *stream);
int rfd_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M, int S,
                          float *grad_points, void *stream);

/* ---- (a5) fused ball query + grouping.
 * xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or NULL (C=0)
 *   -> new_features (B, (use_xyz?3:0)+C, M, S) ; grouped_xyz (B,3,M,S) or NULL ; idx (B,M,S) or NULL.
 * grouped xyz = (xyz[idx] - new_xyz) and, when normalize_xyz, * (1.0f/radius) -- the CUDA reference
 * (torch `tensor /= python_float` lowers to a multiply by the f32 reciprocal).  nsample <= 1024.
 * Uses a persistent per-(device, stream) workspace (uniform grid for N >= 8192, point-major feature copy) grown with
 * cudaMalloc on first use: run once eagerly before capturing the call in a CUDA graph. */
int rfd_query_and_group(const float *xyz, const float *new_xyz, const float *features, int B, int N, int M, int C,
                        float radius, int nsample, int use_xyz, int normalize_xyz, float *new_features,
                        float *grouped_xyz, int *idx, void *stream);

/* ---- (a7) three_nn: unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32 (squared), idx (B,n,3) i32 */
int rfd_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx,
                 void *stream);
/* ---- (a8) three_interpolate: points (B,C,m), idx (B,n,3), weight (B,n,3) -> out (B,C,n); and its grad */
int rfd_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C, int m, int n,
                          float *out, void *stream);
int rfd_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B, int C, int n,
                               int m, float *grad_points, void *stream);
/* fused FP front end: 3-NN + sqrt + 1/(d+1e-8) normalised weights + interpolation, writing into the first
 * C channels of out (B,Ctot,n) (the skip features are concatenated by the caller into channels C..Ctot). */
int rfd_three_nn_interpolate(const float *unknown, const float *known, const float *known_feats, int B, int n,
                             int m, int C, int Ctot, float *out, void *stream);

/* ---- (a6/a9/a10/a11) pointwise (1x1 conv) layer in fp32, BN folded (eval), optional ReLU and max-pool.
 * x (B,Cin,L), W (Cout,Cin) row-major, scale/shift (Cout) [y = relu?(scale*(W.x)+shift)],
 * pool = 1: y (B,Cout,L); pool = S>1: max over each run of S consecutive positions -> y (B,Cout,L/S).
 * residual (B,Cout,L) or NULL is added after the affine (before ReLU is NOT applied to it: y = act(...) + res). */
int rfd_pointwise_mlp_f32(const float *x, const float *W, const float *scale, const float *shift,
                          const float *residual, int relu, int pool, int B, int Cin, int Cout, int L, float *y,
                          void *stream);

/* ---- (a6/a9/a10/a11) every dense pointwise MLP of the detection pass on tcgen05 tensor cores (eval mode).
 * One kernel per module replaces the Conv(1x1)+BatchNorm+ReLU stacks (and the max_pool2d of an SA layer) of
 * PointnetSAModuleVotes (../pointnet2_modules.py:9-19,237-243), PointnetFPModule (:395-405), VotingModule
 * (vote_module.py:34-61) and the ProposalModule head (proposal_module.py:85-124).
 *   mode: RFD_MLP_MODE_BF16 / _F16 (one MMA per K step) or _F16X3 (split fp16, a_hi.w_hi + a_lo.w_hi + a_hi.w_lo:
 *         fp32-grade results, BASELINE config 2's 1e-4) -- fp32 accumulation and fp32 BatchNorm affine in every mode.
 *   layers: 1..3, widths C1, C2, C3 (0 = layer absent); hidden widths <= 256, last width <= 512; every layer
 *         computes y = scale * (W . x) + shift (the folded BatchNorm / conv bias), hidden layers apply ReLU, the last
 *         one iff relu_last.
 *   pack (once per checkpoint): W_l (C_l, C_{l-1}) f32 row-major; W1 is (C1, xyz + K0): with xyz = 3 its first three
 *         input columns are the relative-xyz channels of an SA layer, which the kernel applies in fp32 from
 *         (xyz[idx] - new_xyz) * (1/radius) instead of feeding them to the tensor cores.
 *         -> packed (rfd_mlp_chain_packed_bytes bytes; 0 = unsupported widths).
 *   rfd_mlp_chain: dense rows.  x (B, K0, L) channel-major -> out_cm (B, C, L/pool) and/or out_pm (B, L/pool, C)
 *         (either may be NULL); pool = 1, or 16/32/64/128 = max over runs of `pool` consecutive rows (the grouped
 *         (B,3+C,M,S) tensor viewed as L = M*S; pool > 32 needs relu_last).
 *   rfd_sa_mlp_chain: full set-abstraction fusion: rows are gathered through idx (B,M,S)
 *         (rfd_ball_query) from xyz (B,N,3) and POINT-MAJOR features feat_pm (B,N,C) (rfd_transpose_features, or the
 *         out_pm of the previous layer), centred on new_xyz (B,M,3) exactly as rfd_query_and_group would -- the grouped
 *         tensor is never materialised -- then MLP + max over S.  S in {16,32,64,128}. */
#define RFD_MLP_MODE_BF16 1
#define RFD_MLP_MODE_F16 2
#define RFD_MLP_MODE_F16X3 3
size_t rfd_mlp_chain_packed_bytes(int mode, int K0, int xyz, int C1, int C2, int C3);
int rfd_mlp_chain_pack(int mode, int K0, int xyz, const float *W1, const float *scale1, const float *shift1, int C1,
                       const float *W2, const float *scale2, const float *shift2, int C2, const float *W3,
                       const float *scale3, const float *shift3, int C3, int relu_last, void *packed, void *stream);
int rfd_mlp_chain(int mode, const float *x, int B, int K0, int L, const void *packed, int C1, int C2, int C3,
                  int relu_last, int pool, float *out_cm, float *out_pm, void *stream);
/* rfd_mlp_chain with the options the wide PointNet-style encoders of SkipPropagation need (pointseg.py, layers.py:340-392):
 *   relu_in     ReLU applied to x while it is loaded (ResnetBlockFC's leading activation)
 *   gbias       (B, L/gbias_rows, n0) f32 or NULL: per-group pre-activation bias of layer 0, y = scale*(acc+gbias)+shift, n0 =
 *               C1 rounded up to 64 (16 when C1 is the only layer) -- the contribution of a per-cloud constant input
 *               (the max-pooled "global" half of a PointNet concat) without ever building the concatenated tensor
 *   out_pool    (B, C_last, L/pool_rows) f32 or NULL, initialised by the caller (-inf): max over groups of pool_rows rows of
 *               the unpooled output (pool must be 1), any sign; out_cm / out_pm may then both be NULL
 * gbias_rows and pool_rows are multiples of the 128-row tile and divide L. */
int rfd_mlp_chain_ex(int mode, const float *x, int B, int K0, int L, const void *packed, int C1, int C2, int C3,
                     int relu_last, int pool, float *out_cm, float *out_pm, int relu_in, const float *gbias,
                     int gbias_rows, float *out_pool, int pool_rows, void *stream);
/* the same options on ROW-MAJOR operands: x_pm (B, L, ldi) f32, the K0 input channels in columns [0, K0) of every row (ldi a
 * multiple of 4, base 16-byte aligned); out_pm (B, L, ldo) receives the C_last outputs in columns [out_col0, out_col0 +
 * C_last) (or NULL); out_pool (B, L/pool_rows, C_last).  A tile reads 128 contiguous rows -- sequential HBM access where the
 * channel-major form touches one DRAM page per channel -- and channel concatenation is a column offset into a wider row. */
int rfd_mlp_chain_rows(int mode, const float *x_pm, int ldi, int B, int K0, int L, const void *packed, int C1, int C2, int C3,
                       int relu_last, float *out_pm, int ldo, int out_col0, int relu_in, const float *gbias,
                       int gbias_rows, float *out_pool, int pool_rows, void *stream);
int rfd_sa_mlp_chain(int mode, const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx, int B, int N,
                     int M, int S, int C, float radius, int normalize_xyz, const void *packed, int C1, int C2, int C3,
                     float *out_cm, float *out_pm, void *stream);
/* features (B,C,N) channel-major -> point-major (B,N,Cp), Cp >= C a multiple of 4, padding zero-filled */
int rfd_transpose_features(const float *features, int B, int C, int N, int Cp, float *out, void *stream);

/* ---- STN_Group (pointnet2_modules.py:468-537), the per-proposal grouping of SkipPropagation.
 * rfd_query_and_group_rotated = rfd_query_and_group + the rotation of the relative coordinates about z by the box heading
 * (rot = [[cos, sin, 0], [-sin, cos, 0], [0, 0, 1]], :513-526), heading (B,M) radians, in the same kernel (RfD-Net:
 * radius 1.0, nsample 1024 over the whole cloud).  rfd_stn_apply = STN3d's last step (:455-462): theta (B,12,M) is the
 * regressed 3x4 matrix (row major, WITHOUT the identity, which the kernel adds), out = theta[:, :3] . g + theta[:, 3].
 * With use_xyz the xyz channels inside new_features are rotated as well; the reference leaves those unrotated (only the
 * returned grouped_xyz is rotated, :497-526) -- rfdnet_b200.stn_group.STN_Group restores them with a second call. */
int rfd_query_and_group_rotated(const float *xyz, const float *new_xyz, const float *features, const float *heading,
                                int B, int N, int M, int C, float radius, int nsample, int use_xyz, int normalize_xyz,
                                float *new_features, float *grouped_xyz, int *idx, void *stream);
int rfd_stn_apply(const float *grouped_xyz /*(B,3,M,S)*/, const float *theta /*(B,12,M)*/, int B, int M, int S,
                  float *out /*(B,3,M,S)*/, void *stream);

/* ---- (a13) occupancy query lattice: out (R^3,3) f32 = box_size * linspace(-0.5,0.5,R) on each axis, z fastest */
int rfd_make_3d_grid(int R, float box_size, float *out, void *stream);

/* ---- occupancy mask (first step): logits (B,T) f32 -> bits (B, ceil(T/32)) u32 with bit t%32 of
 * word t/32 set iff logit >= threshold (the reference thresholds at logit(0.5) = 0, generator.py:160;
 * external/common.py:7-35 compute_iou), counts (B) i32 = occupied points per object (NULL = skip). */
int rfd_occupancy_bits(const float *logits, int B, int T, float threshold, uint32_t *bits, int *counts, void *stream);

/* ---- surface extraction: Generator3D.extract_mesh (models/iscnet/modules/generator.py:145-168) for a
 * batch of objects, on the device: np.pad(occ_hat, 1, -1e6) + mcubes.marching_cubes(., threshold) (PyMCubes 0.1.2,
 * environment.yml:77; Bourke's table, corner bit set iff value <= threshold) + the vertex transform
 * box_size * (((v - 0.5) - 1) / (R-1) - 0.5), all evaluated in fp64 in the reference's operation order.
 *   logits (B, R^3) f32 (x slowest, z fastest: values.reshape(R,R,R), generator.py:97); R <= 32
 *   threshold = log(t) - log(1-t) as the reference computes it (0.0 for t = 0.5); box_size = 1 + padding
 *   vertices: pool of cap_vertices x 3 f64 (vertex_f64 != 0) or f32; triangles: pool of cap_triangles x 3 i32 holding
 *   OBJECT-LOCAL vertex ids; ranges (B,4) i32 = {vertex offset, vertex count, triangle offset, triangle count} of every
 *   object inside the pools (offsets -1 if the object did not fit: counts are still valid, re-run with larger pools);
 *   totals: 3 x u64, ZEROED BY THE CALLER, incremented atomically: vertices reserved, triangles reserved, objects that
 *   did not fit.
 * Order inside an object: vertices by (lower end point of their lattice edge, axis), triangles by cell (x slowest)
 * then table order (= PyMCubes' triangle order; its vertex order, which is creation order, is not reproduced). */
int rfd_extract_mesh(const float *logits, int B, int R, double threshold, double box_size, void *vertices,
                     int vertex_f64, int *triangles, long long cap_vertices, long long cap_triangles, int *ranges,
                     unsigned long long *totals, void *stream);

/* ---- (a12) ONet decoder (DecoderCBatchNorm, eval mode), hidden = 256, n_blocks = 5.
 * Step 1 (once per checkpoint): pack the fp32 fc weights into the device layout the kernel streams:
 *   fc_w (10,256,256) f32 = [blocks.0.fc_0, blocks.0.fc_1, blocks.1.fc_0, ...].weight  ([out][in])
 *   -> packed (rfd_onet_packed_bytes(mode) bytes): 16-bit operands, one 32-KB image per (layer, 64-wide K panel) of
 *      the UMMA K-major 128B-swizzled B operand, in consumption order (MODE_F16X3: hi image then lo image).
 * Step 2 (per batch of objects): conditional-BN tables (tiny fp32 GEMMs)
 *   c (B,c_dim); gamma_w/beta_w (11,256,c_dim) and gamma_b/beta_b (11,256) = conv_gamma/conv_beta of
 *   [blocks.0.bn_0, blocks.0.bn_1, ..., blocks.4.bn_1, bn]; run_mean/run_var (11,256); eps;
 *   fc_bias (10,256) biases of the fc layers above; x_bias (B,256) = fc_p.bias + fc_z(z) (+ fc_z.bias)
 *   -> aff (B, rfd_onet_aff_floats()) f32: per object [11][2][256] scale a / shift c with
 *      relu(CBN_l(x_true)) == relu(a * x_acc + c)  (x_acc = bias-free accumulator held by the kernel),
 *      the 256 x_bias values, and the same (a,c) again in the column-pair-interleaved order
 *      [11][128]{a0,a1,c0,c1} the tensor-core kernel stages in shared memory.
 * Step 3: logits (B,T) = decoder(p).  p is (B,T,3) with p_batch_stride = T*3 floats, or one shared
 *   (T,3) lattice for every object with p_batch_stride = 0.
 * `mode` selects the tensor-core operand format (weights must have been packed with the same mode); accumulation, the
 * residual stream, the conditional-BN affine, fc_p and fc_out are fp32 in every mode:
 *   RFD_ONET_MODE_BF16  (1)  bf16 x bf16, one MMA per K step.  |dlogit| ~ 3e-3 on unit-scale logits.
 *   RFD_ONET_MODE_F16   (2)  fp16 x fp16, one MMA per K step, same speed, 8x smaller rounding error (~4e-4): meets
 *                            BASELINE config 4's 1e-3.  The default of the Python mirror and of bench.py.  Activations
 *                            saturate at +-65504 (never reached after conditional BN in practice).
 *   RFD_ONET_MODE_F16X3 (3)  split fp16: a_hi.w_hi + a_lo.w_hi + a_hi.w_lo, three MMAs per K step (~22 significant
 *                            bits per operand): |dlogit| ~ 1e-6, meets the north star's 1e-4 ("exact" tensor-core mode).
 * Anything else: RFD_ERR_INVALID_ARGUMENT. */
#define RFD_ONET_MODE_BF16 1
#define RFD_ONET_MODE_F16 2
#define RFD_ONET_MODE_F16X3 3
size_t rfd_onet_packed_bytes(int mode); /* 0 for an unknown mode */
size_t rfd_onet_aff_floats(void);
int rfd_onet_pack_weights(const float *fc_w, int mode, void *packed, void *stream);
int rfd_onet_cbn_tables(const float *c, int B, int c_dim, const float *gamma_w, const float *gamma_b,
                        const float *beta_w, const float *beta_b, const float *run_mean, const float *run_var,
                        float eps, const float *fc_bias, const float *x_bias, float *aff, void *stream);
int rfd_onet_decode(const float *p, long long p_batch_stride, int B, int T, const float *fc_p_w /*(256,3)*/,
                    const void *packed, int mode, const float *aff, const float *fc_out_w /*(256)*/,
                    float fc_out_b, float *logits, void *stream);
/* tuning knob (process-wide, atomic): 1 = independent CTAs; 2 = CTA pairs (thread-block clusters) that share every
 * weight stage through cp.async.bulk multicast, halving the L2 -> SM weight traffic.  Default: RFD_ONET_CLUSTER or 2. */
int rfd_onet_decode_set_cluster(int cluster);
/* fp32 CUDA-core implementation of the same decoder (exact path, slow): the 1e-4 parity claim, the
 * on-GPU yardstick for the tensor-core path, and the "fp32" row of the benchmark.
 * workspace: at least 2*256*T*4 bytes (one object); larger = more objects per pass. */
int rfd_onet_decode_f32(const float *p, long long p_batch_stride, int B, int T, const float *fc_p_w,
                        const float *fc_w /*(10,256,256)*/, const float *aff, const float *fc_out_w,
                        float fc_out_b, float *logits, float *workspace, size_t workspace_bytes, void *stream);
/* diagnostics: rfd_onet_decode (mode RFD_ONET_MODE_F16, independent CTAs) through an instrumented kernel whose CTA 0
 * records clock64() at the MMA/epilogue hand-off points of its first two tiles into `trace` (device, 2*10*16 u64). */
int rfd_onet_decode_traced(const float *p, long long p_batch_stride, int B, int T, const float *fc_p_w,
                           const void *packed, int mode, const float *aff, const float *fc_out_w, float fc_out_b,
                           float *logits, unsigned long long *trace, void *stream);
/* tcgen05 plumbing self-test: D (128,256) f32 = bf16(A (128,64)) . bf16(B (256,64))^T */
int rfd_umma_selftest(const float *A, const float *B, float *D, void *stream);
/* same product with the A operand staged in tensor memory (tcgen05.st.16x128b + TS-mode tcgen05.mma) */
int rfd_umma_selftest_ts(const float *A, const float *B, float *D, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RFDNET_B200_H */
