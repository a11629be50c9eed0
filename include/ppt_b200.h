/*
 * ppt_b200.h -- C ABI of libppt_b200.so: the B200 (sm_100a) point-cloud
 * tokenizer behind auniquesun/PPT's Python hot path.
 *
 * The reference has no FFI for this path: its "operator interface" is a set
 * of Python functions / nn.Modules built from ATen ops.  Each entry point
 * below replaces one of them; the reference lines are cited per function
 * (paths relative to the reference root) and INTEGRATION.md shows the ctypes
 * binding a maintainer would add on the reference side.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer into memory owned by the caller; the
 *     library never allocates, frees or synchronises, and keeps no state
 *     except cached cudaFuncSetAttribute calls;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and
 *     the call returns immediately;
 *   - tensors are dense row-major, fp32 coordinates/features, int64 indices
 *     (the dtypes of the reference API);
 *   - return value: 0 on success, a cudaError_t (> 0) for a CUDA failure, or
 *     a negative PPT_E* code for an argument the kernels do not support.
 */
#ifndef PPT_B200_H_
#define PPT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPT_B200_ABI_VERSION 2

#define PPT_EINVAL (-1) /* bad shape / null pointer */
#define PPT_ERANGE (-2) /* size outside what the kernel supports */

/* Precision modes of the patch Encoder. */
#define PPT_ENC_FP16 0   /* fp16 operands, fp32 accumulate (tcgen05 kind::f16) */
#define PPT_ENC_BF16 1   /* bf16 operands, fp32 accumulate */
#define PPT_ENC_FP16X3 2 /* fp16 hi/lo split (power-of-two pre-scaled), 3 MMAs per product: fp32-parity mode */

int ppt_abi_version(void);

/* Human-readable text for a return code (static storage). */
const char *ppt_strerror(int code);

/* Spatial index of a batch of clouds (Morton-cell order + per-row bounding boxes), shared by
 * ppt_fps and ppt_knn / ppt_knn_group: with it they skip the parts of a cloud that cannot change the
 * result -- the outputs are bit-identical with and without it.  Supported for 512 <= N <= 32768
 * (ppt_spatial_index_bytes returns 0 otherwise; pass index = NULL then); ppt_fps uses it up to 8192
 * points (its state lives in shared memory) and runs its plain kernel above, ppt_knn* use it throughout.
 *   index: ppt_spatial_index_bytes(B, N) bytes of caller-owned scratch, valid until xyz changes. */
int64_t ppt_spatial_index_bytes(int B, int N);
int ppt_spatial_index_build(const float *xyz, void *index, int B, int N, void *stream);

/* farthest_point_sample -- models/pointbert/misc.py:44-69,
 * models/pointbert/pointnet2_utils.py:95-116, models/pointnet2/pointnet2_utils.py:63-84,
 * models/pointmlp/pointMLP.py:64-84 (identical results).
 * The caller draws `start` with the reference's own torch.randint call.
 *   xyz [B,N,3] f32; start [B] i64; idx_out [B,G] i64;
 *   centers_out [B,G,3] f32 or NULL (= index_points(xyz, idx), misc.py:12-24 `fps`);
 *   index: NULL or a built spatial index of xyz.
 * N <= 65536. */
int ppt_fps(const float *xyz, const int64_t *start, int64_t *idx_out, float *centers_out,
            const void *index, int B, int N, int G, void *stream);

/* square_distance -- models/pointbert/dvae.py:130-149 (same text in both
 * pointnet2_utils.py copies and pointMLP.py:23-42).
 *   src [B,S,3], dst [B,N,3] -> out [B,S,N] f32, bit pattern of the reference's CPU path. */
int ppt_square_distance(const float *src, const float *dst, float *out, int B, int S, int N, void *stream);

/* knn_point -- models/pointbert/dvae.py:116-127 (pointnet2_utils.py:20-34, pointMLP.py:110-121).
 *   xyz [B,N,3]; query [B,S,3] -> idx_out [B,S,k] i64, the k nearest under
 *   (distance, index), ascending; dist_out [B,S,k] f32 or NULL.  1 <= k <= 32, k <= N.
 *   index: NULL (full scan of the cloud per query) or a built spatial index of xyz (pruned search). */
int ppt_knn(const float *xyz, const float *query, int64_t *idx_out, float *dist_out, const void *index,
            int B, int N, int S, int k, void *stream);

/* Group.forward after FPS -- models/pointbert/dvae.py:159-181: kNN of each
 * centre, flat gather of the neighbours, subtraction of the centre.
 *   xyz [B,N,3]; center [B,G,3] -> neighborhood_out [B,G,k,3] f32;
 *   idx_out [B,G,k] i64 or NULL.  1 <= k <= 32.  index as for ppt_knn. */
int ppt_knn_group(const float *xyz, const float *center, float *neighborhood_out, int64_t *idx_out,
                  const void *index, int B, int N, int G, int k, void *stream);

/* query_ball_point -- models/pointnet2/pointnet2_utils.py:87-107
 * (pointbert/pointnet2_utils.py:119-139, pointMLP.py:87-107).
 * `radius2` is (float)(radius**2): torch compares in fp32.
 *   xyz [B,N,3]; new_xyz [B,S,3] -> idx_out [B,S,nsample] i64. */
int ppt_ball_query(const float *xyz, const float *new_xyz, int64_t *idx_out, float radius2,
                   int B, int N, int S, int nsample, void *stream);

/* index_points -- models/pointbert/misc.py:26-42 (and the three copies):
 *   points [B,N,C]; idx [B,M] i64 (M = product of the trailing idx dims) -> out [B,M,C]. */
int ppt_gather(const float *points, const int64_t *idx, float *out, int B, int N, int C, int M, void *stream);

/* Grouping tail of sample_and_group / PointNetSetAbstractionMsg.forward --
 * models/pointnet2/pointnet2_utils.py:127-134 and :244-254:
 *   out[b,s,j,:] = cat(xyz[b,idx] - new_xyz[b,s], points[b,idx])   (xyz_first = 1, SSG order)
 *                = cat(points[b,idx], xyz[b,idx] - new_xyz[b,s])   (xyz_first = 0, MSG order)
 *   xyz [B,N,3]; new_xyz [B,S,3]; points [B,N,D] or NULL (D = 0); idx [B,S,K] i64;
 *   out [B,S,K,3+D]. */
int ppt_group_concat(const float *xyz, const float *new_xyz, const float *points, const int64_t *idx,
                     float *out, int B, int N, int S, int K, int D, int xyz_first, void *stream);

/* three_nn part of PointNetFeaturePropagation.forward --
 * models/pointnet2/pointnet2_utils.py:300-302 (pointbert/pointnet2_utils.py:340-342):
 * squared distances (not sqrt), three nearest under (distance, index).
 *   unknown [B,N,3]; known [B,S,3], S >= 3 -> dist_out [B,N,3] f32; idx_out [B,N,3] i64. */
int ppt_three_nn(const float *unknown, const float *known, float *dist_out, int64_t *idx_out,
                 int B, int N, int S, void *stream);

/* three_interpolate -- models/pointnet2/pointnet2_utils.py:304-307:
 * weights 1/(d+1e-8) normalised, ((w0*f0 + w1*f1) + w2*f2).
 *   feats [B,S,D]; idx [B,N,3] i64; dist [B,N,3] -> out [B,N,D]. */
int ppt_three_interpolate(const float *feats, const int64_t *idx, const float *dist, float *out,
                          int B, int N, int S, int D, void *stream);

/* Gradient of three_interpolate w.r.t. feats (the part-seg head trains through
 * it, models/pointbert/point_encoder.py:404-413).  grad_feats [B,S,D] must be
 * zero-filled by the caller; contributions are accumulated with red.global.add. */
int ppt_three_interpolate_grad(const float *grad_out, const int64_t *idx, const float *dist,
                               float *grad_feats, int B, int N, int S, int D, void *stream);

/* DGCNN_Propagation.get_graph_feature behind its kNN (models/pointbert/pointnet2_utils.py:392-442, the part-seg
 * head, point_encoder.py:409-411):
 *   out[b, c, q, j] = x_k[b, c, idx[b,q,j]] - x_q[b, c, q]   (c < C),   out[b, C + c, q, j] = x_q[b, c, q].
 * x_q [B,C,Nq], x_k [B,C,Nk] f32 channel-first, idx [B,Nq,k] int64 in [0,Nk) -> out [B,2C,Nq,k] f32.  Bit-exact. */
int ppt_graph_feature(const float *x_q, const float *x_k, const int64_t *idx, float *out,
                      int B, int C, int Nq, int Nk, int k, void *stream);

/* Its gradient w.r.t. x_q [B,C,Nq] (written) and x_k [B,C,Nk] (accumulated with atomics into a buffer the
 * caller has zero-filled); grad_out [B,2C,Nq,k]. */
int ppt_graph_feature_grad(const float *grad_out, const int64_t *idx, float *grad_xq, float *grad_xk,
                           int B, int C, int Nq, int Nk, int k, void *stream);

/* DGCNN_Propagation layer behind its per-point GEMMs (models/pointbert/pointnet2_utils.py:382-390, 444-467; eval):
 * the layer's Conv2d(2C -> Co, 1x1, no bias) on cat(x_k[idx] - x_q, x_q) is linear, so with W = [Wa | Wb] its output
 * is U[:, idx] + V with U = Wa x_k [B,Co,Nk] and V = (Wb - Wa) x_q [B,Co,Nq] (two plain GEMMs the caller runs).
 * This entry does the rest: GroupNorm(G groups, eps) statistics over (Co/G, Nq, k), affine (gamma, beta [Co]),
 * LeakyReLU(slope) and the max over the k neighbours -> out [B,Co,Nq].  idx [B,Nq,k] int64 in [0,Nk), k <= 16.
 *   workspace: ppt_edge_gn_workspace_bytes(B, G) bytes. */
int64_t ppt_edge_gn_workspace_bytes(int B, int G);
int ppt_edge_gn_max_forward(const float *U, const float *V, const int64_t *idx, const float *gamma, const float *beta,
                            void *workspace, float *out, int B, int C, int Nq, int Nk, int k, int G, float eps,
                            float slope, void *stream);

/* ---- PointNet++ set-abstraction shared MLP + max-pool (tcgen05) --------------
 * PointNetSetAbstraction[Msg].forward behind the grouping, eval mode (models/pointnet2/pointnet2_utils.py:196-201,
 * 256-261): 3 x (Conv2d 1x1 + BatchNorm2d + ReLU) over [xyz - centre, features] of every (group, sample), then max
 * over the nsample samples.  The gather / centre / concat is fused into the build of the first layer's operand
 * images; activations between layers are fp16 operand images in `workspace`.
 *   xyz [B,N,3], feats [B,N,D] (channel-last; NULL iff D == 0), new_xyz [B,S,3] (zeros for group_all),
 *   idx [B,S,nsample] int64 (ppt_ball_query; arange for group_all), nsample in {16,32,64,128};
 *   packed: ppt_b200/encoder_pack.py:pack_sa_mlp (BatchNorm folded, layer-1 columns in [features, xyz] order),
 *   ppt_sa_mlp_packed_bytes(D + 3, c1, c2, c3) bytes (PPT_ERANGE if a layer has more than 512 input channels);
 *   workspace: ppt_sa_mlp_workspace_bytes(B*S*nsample, D + 3, c1, c2, c3) bytes;
 *   out [B, c3, S] f32 (channel-first, as the module returns it).  mode: PPT_ENC_FP16 or PPT_ENC_BF16, optionally
 *   | PPT_SA_PER_LAYER: one kernel per layer with the activations as operand images in `workspace` (the fallback
 *   taken by itself when a level's activations do not fit in shared memory) instead of the single fused kernel. */
#define PPT_SA_PER_LAYER 0x100
int64_t ppt_sa_mlp_packed_bytes(int c0, int c1, int c2, int c3);
int64_t ppt_sa_mlp_workspace_bytes(int64_t num_columns, int c0, int c1, int c2, int c3);
int ppt_sa_mlp_forward(const float *xyz, const float *feats, const float *new_xyz, const int64_t *idx,
                       const void *packed, void *workspace, float *out, int B, int N, int S, int nsample, int D,
                       int c1, int c2, int c3, int mode, void *stream);

/* ---- PointNetFeaturePropagation: three_interpolate + concat + two-layer MLP (tcgen05) ----------------------------
 * PointNetFeaturePropagation.forward behind three_nn, eval mode (models/pointnet2/pointnet2_utils.py:304-319;
 * the part-seg head's propagation_{0,1,2}: in = 384 + 3 (+16), mlp = [1536, 384], models/pointbert/point_encoder.py:300-302):
 *     new_points = cat([points1, three_interpolate(points2, idx, dist)]) -> 2 x (Conv1d 1x1 + BatchNorm1d + ReLU)
 * The interpolation and the concatenation happen while the first layer's fp16 operand images are built (neither the
 * interpolated [B,N,D2] tensor nor the concatenated one exists); layer 1 keeps its input tile in shared memory, layer 2
 * (K = c1 up to 4096) streams it in K blocks with all its output units resident in tensor memory.
 *   points1 [B,D1,N] f32 channel-first (NULL iff D1 == 0); feats2 = points2 as [B,S,D2] f32 channel-last;
 *   idx [B,N,3] i64, dist [B,N,3] f32 from ppt_three_nn (S >= 3);
 *   packed: ppt_b200/encoder_pack.py:pack_fp_mlp (BatchNorm folded, layer-1 columns in [interpolated | points1] order),
 *   ppt_fp_mlp_packed_bytes(D1 + D2, c1, c2) bytes (PPT_ERANGE if D1 + D2 > 512 or c2 > 512);
 *   workspace: ppt_fp_mlp_workspace_bytes(B * N, D1 + D2, c1, c2) bytes;
 *   out [B, c2, N] f32 channel-first (as the module returns it).  mode: PPT_ENC_FP16 or PPT_ENC_BF16.  Forward only. */
int64_t ppt_fp_mlp_packed_bytes(int c0, int c1, int c2);
int64_t ppt_fp_mlp_workspace_bytes(int64_t num_points, int c0, int c1, int c2);
int ppt_fp_mlp_forward(const float *points1, const float *feats2, const int64_t *idx, const float *dist,
                       const void *packed, void *workspace, float *out, int B, int N, int S, int D1, int D2,
                       int c1, int c2, int mode, void *stream);

/* ---- mini-PointNet patch Encoder + reduce_dim (tcgen05) ---------------------
 * Encoder.forward in eval mode, models/pointbert/dvae.py:201-215, followed by
 * reduce_dim, models/pointbert/point_encoder.py:133,239.
 *
 * The host folds BatchNorm into the convolutions, splits second_conv.0 into
 * its global-feature and per-point halves and packs everything into one device
 * blob (ppt_b200/encoder_pack.py documents the layout); the blob is opaque here.
 *   packed_bytes = ppt_encoder_packed_bytes(mode);
 *   workspace    = ppt_encoder_workspace_bytes(num_groups, mode) bytes of scratch.
 *   neighborhood [num_groups, 32, 3] f32 -> tokens_out [num_groups, 384] f32 (or NULL),
 *   features_out [num_groups, 256] f32 or NULL (the Encoder's own output); not both NULL.
 * num_groups = B*G; any value >= 1. */
int64_t ppt_encoder_packed_bytes(int mode);
int64_t ppt_encoder_workspace_bytes(int64_t num_groups, int mode);
int ppt_encoder_forward(const float *neighborhood, const void *packed, void *workspace,
                        float *features_out, float *tokens_out, int64_t num_groups, int mode, void *stream);

/* Same pipeline, one or more of its four launches at a time (bit 0 stage1, bit 1 per-group
 * linear -> c, bit 2 stage2, bit 3 per-group linear -> tokens), in order, sharing `workspace`.
 * Lets a caller bracket one launch with events; tokens_out may be NULL if features_out is given
 * (plain Encoder.forward without reduce_dim). */
int ppt_encoder_forward_phases(const float *neighborhood, const void *packed, void *workspace,
                               float *features_out, float *tokens_out, int64_t num_groups, int mode,
                               int phases, void *stream);

/* The general entry point behind the two above.
 *   flags: PPT_TOKENS_F16 -- tokens_out is [num_groups, 384] IEEE fp16 (the fp32 result rounded once more, saturating):
 *          half the bytes for a caller that ships tokens to the host or feeds an fp16/autocast transformer;
 *   clock_acc: NULL, or a device int64[2] the caller zeroed: CTA 0 of the stage-2 launch adds its lifetime in
 *          nanoseconds (globaltimer) to [0] and in SM cycles (clock64) to [1] -- [1]/[0] is the SM clock in GHz inside
 *          that kernel (measurement aid; selects a separately compiled copy of the kernel, the library keeps no state). */
#define PPT_TOKENS_F16 1
int ppt_encoder_forward_ex(const float *neighborhood, const void *packed, void *workspace,
                           float *features_out, void *tokens_out, int64_t num_groups, int mode,
                           int phases, int flags, void *clock_acc, void *stream);

/* ---- Encoder.forward under model.train(): batch-statistics BatchNorm --------
 * PPT trains with model.train() (main_cls.py:169), which puts the FROZEN Encoder's two BatchNorm1d layers
 * (models/pointbert/dvae.py:190,196) in batch-statistics mode: they normalise with the mean / biased variance
 * of the current batch and update their running statistics (momentum, unbiased variance, num_batches_tracked).
 * Forward only (the Encoder's parameters take no gradient in PPT, models/ULIP_models.py:505).
 *   packed_train: MUTABLE device copy of the blob packed with both BatchNorms as identities
 *                 (ppt_b200/encoder_pack.py:pack_encoder_train); sections that depend on the batch are rewritten;
 *   bn: device pointers into the module's own tensors (fp32 contiguous; num_batches_tracked int64 or NULL);
 *       running_mean / running_var / num_batches_tracked are updated in place, in stream order;
 *   workspace: ppt_encoder_train_workspace_bytes(num_groups, mode) bytes. */
typedef struct {
  const float *conv1_weight, *conv1_bias;       /* first_conv.0: [128,3(,1)], [128] */
  const float *bn1_weight, *bn1_bias;           /* first_conv.1 */
  float *bn1_running_mean, *bn1_running_var;    /* [128] */
  int64_t *bn1_num_batches_tracked;
  const float *bn2_weight, *bn2_bias;           /* second_conv.1 */
  float *bn2_running_mean, *bn2_running_var;    /* [512] */
  int64_t *bn2_num_batches_tracked;
  float momentum, eps;                          /* 0.1, 1e-5 in the reference */
} ppt_encoder_bn_t;
int64_t ppt_encoder_train_workspace_bytes(int64_t num_groups, int mode);
int ppt_encoder_forward_train(const float *neighborhood, void *packed_train, const ppt_encoder_bn_t *bn,
                              void *workspace, float *features_out, float *tokens_out, int64_t num_groups,
                              int mode, void *stream);

/* ---- pos_embed + token assembly (the step right after the tokenizer) ----------
 * PointTransformer.forward, models/pointbert/point_encoder.py:239-247:
 *     x   = cat(cls_token, reduce_dim(encoder(neighborhood)))         [clouds, G+1, 384]
 *     pos = cat(cls_pos,   pos_embed(center))                         [clouds, G+1, 384]
 * with pos_embed = Linear(3,128) -> GELU (erf) -> Linear(128,384) (point_encoder.py:138-142).  The
 * tokens are stored straight into rows 1..G of x_out (no separate tokens tensor, no concatenation
 * pass), the 128 -> 384 layer runs on the tensor cores like reduce_dim.
 *   posembed_packed: ppt_b200/encoder_pack.py:pack_pos_embed (opaque), ppt_posembed_packed_bytes(mode) bytes;
 *   workspace: ppt_tokenizer_workspace_bytes(num_groups, mode) bytes;
 *   center [num_groups, 3] f32; num_groups = clouds * groups_per_cloud, groups_per_cloud >= 32;
 *   x_out may be NULL (only pos is computed; neighborhood / encoder_packed are then unused). */
int64_t ppt_posembed_packed_bytes(int mode);
int64_t ppt_tokenizer_workspace_bytes(int64_t num_groups, int mode);
int ppt_tokenizer_forward(const float *neighborhood, const float *center, const void *encoder_packed,
                          const void *posembed_packed, void *workspace, float *x_out, float *pos_out,
                          int64_t num_groups, int groups_per_cloud, int mode, void *stream);

/* Self-test of the tcgen05 building blocks (one 128 x N x K GEMM through the
 * same smem layouts, descriptors and epilogue the Encoder uses).
 *   a [128,K] f32, b [N,K] f32 -> d [128,N] f32 = a * b^T with operands rounded to `mode`'s type. */
int ppt_selftest_umma(const float *a, const float *b, float *d, int N, int K, int mode, void *stream);

/* Measurement aid (bench.py): one device thread records (globaltimer ns, clock64 cycles) pairs every
 * period_ns into out [samples][2] int64, on `stream` -- run it on a side stream next to the kernels under
 * test to see the SM clock inside them (DESIGN.md "clocks under load"). */
int ppt_clock_probe(void *out, int samples, int64_t period_ns, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PPT_B200_H_ */
