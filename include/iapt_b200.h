/*
 * iapt_b200.h - C ABI of libiapt_b200.so: hand-written sm_100a kernels for ImgAdaPoinTr's point-cloud
 * geometry hot path (Chamfer nearest-neighbour search fwd/bwd and fused losses, furthest-point sampling, gather,
 * kNN, the kNN consumers of the grouper and the geometry-aware transformer, pointnet2-style grouping, auction EMD).
 *
 * This header is the drop-in boundary.  Each entry point names the reference interface it replaces
 * (paths relative to the reference repo root); INTEGRATION.md shows the reference-side bindings.
 *
 * Conventions (all entry points)
 *   - plain C types only: device pointers, sizes and a stream; no torch / C++ types cross the boundary;
 *   - every pointer is a DEVICE pointer to contiguous row-major memory on the current CUDA device unless the
 *     name ends in `_host`; float = IEEE binary32, indices int32 (int64 for kNN, as the reference returns);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is enqueued, not synchronised;
 *   - return value: 0 on success, a positive cudaError_t value for CUDA failures, a negative IAPT_E* code for
 *     argument errors (the reference printf()s CUDA errors and continues, chamfer.cu:166-169 - we return them);
 *     iapt_error_string() describes any of them;
 *   - workspaces are caller-allocated (so a framework's caching allocator owns all memory); the *_workspace_bytes
 *     functions are pure host arithmetic and never touch the GPU.
 */
#ifndef IAPT_B200_H_
#define IAPT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IAPT_ABI_VERSION 2

#define IAPT_EINVAL (-1)     /* bad size / null pointer */
#define IAPT_EWORKSPACE (-2) /* workspace too small */
#define IAPT_EUNSUPPORTED (-3)

typedef void* iapt_stream_t; /* cudaStream_t */

int iapt_abi_version(void);
const char* iapt_error_string(int code);
/* Number of kernels launched through this library by the calling process so far (bench.py's gpu_launches). */
unsigned long long iapt_launch_count(void);

/* Measurement hook (bench.py roofline): when enabled, every launch that runs the nearest-neighbour search records a
 * CUDA-event pair on its launch stream tightly around the dominant kernel (chamfer_step_kernel).
 * iapt_profile_collect() synchronises those events and returns the summed device time in ms and the number of
 * launches; enabling (or disabling) clears the records.  Not thread-safe; for benchmarking only. */
int iapt_profile_enable(int on);
int iapt_profile_collect(double* total_ms, int* launches);

/* ---- Chamfer nearest-neighbour search --------------------------------------------------------------------
 * Replaces chamfer.forward -> chamfer_cuda_forward (extensions/chamfer_dist/chamfer_cuda.cpp:22-25,
 * chamfer.cu:147-171; kernel chamfer.cu:15-145).
 *   xyz1 [B,N,3], xyz2 [B,M,3]  ->  dist1 [B,N] = min_k |xyz1[b,j]-xyz2[b,k]|^2, idx1 [B,N] = lowest argmin k,
 *                                   dist2 [B,M], idx2 [B,M] with the roles swapped.
 * Arithmetic: d = fma(dz,dz, fma(dx,dx, dy*dy)) in FP32 - bit-identical to the reference kernel as compiled by
 * nvcc 12.9 - so dist AND idx match the reference bit for bit (lowest index on exact ties, chamfer.cu:47-50,137-140).
 * N == 0 or M == 0 leaves zero-filled outputs like the reference's torch::zeros.  Coordinates must be finite
 * with |x| < 1e15 (NaN/Inf are outside the reference's contract as well).
 * Either output pair may be NULL (dist1/idx1 or dist2/idx2) to skip that direction's final write-out. */
size_t iapt_chamfer_forward_workspace_bytes(int B, int N, int M);
int iapt_chamfer_forward(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2,
                         int32_t* idx1, int32_t* idx2, void* workspace, size_t workspace_bytes,
                         iapt_stream_t stream);

/* Replaces chamfer.backward -> chamfer_cuda_backward (chamfer_cuda.cpp:27-34, chamfer.cu:203-229; kernel :173-201).
 *   grad_xyz1[b,j] = 2 g1[b,j] (x1_j - x2_idx1[j]) - sum_{k: idx2[k]==j} 2 g2[b,k] (x2_k - x1_j), and symmetrically
 *   for grad_xyz2.  Deterministic and free of floating-point atomics: the points that chose j as nearest neighbour are
 *   grouped per j (counting sort with integer atomics) and summed by j's owner - up to 8 of them in ascending source
 *   index, more of them exactly (64-bit fixed point, scale from the segment's largest term), so the result depends
 *   neither on scheduling nor on the launch geometry and a collapsed prediction is a parallel reduction.
 *   Outputs are fully written (no pre-zeroing needed); either gradient may be NULL (not wanted), a NULL grad_dist
 *   means zeros.  The workspace is REQUIRED (iapt_chamfer_backward_workspace_bytes: 12 bytes per point + O(B)).
 *   Indices outside [0, other side's size) are ignored (the reference would read out of bounds). */
size_t iapt_chamfer_backward_workspace_bytes(int B, int N, int M);
int iapt_chamfer_backward(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                          const float* grad_dist1, const float* grad_dist2, int B, int N, int M, float* grad_xyz1,
                          float* grad_xyz2, void* workspace, size_t workspace_bytes, iapt_stream_t stream);

/* Fused loss epilogue for ChamferDistanceL1 / L2 / L2_split / L1_PM (extensions/chamfer_dist/__init__.py:28-103):
 * from dist1/dist2 computes, per cloud b, sums[b] = { sum_j dist1, sum_k dist2, sum_j sqrt(dist1), sum_k sqrt(dist2) }
 * (FP32 sums in a fixed order -> deterministic) so the host forms mean(.) without extra elementwise launches.
 * sums is [B,4] float. */
int iapt_chamfer_reduce(const float* dist1, const float* dist2, int B, int N, int M, float* sums,
                        iapt_stream_t stream);

/* Fused loss modules (extensions/chamfer_dist/__init__.py:28-103) - the "next" row of SURVEY.md section 8f.
 * iapt_chamfer_loss: from dist1/dist2 writes the per-cloud sums [B,4] (as iapt_chamfer_reduce) AND the final loss
 * scalars loss[2] with the reference's reductions over ALL B*N / B*M elements:
 *   IAPT_LOSS_L1       loss[0] = (mean sqrt d1 + mean sqrt d2) / 2          (ChamferDistanceL1,       :64-84)
 *   IAPT_LOSS_L2       loss[0] =  mean d1 + mean d2                         (ChamferDistanceL2,       :28-44)
 *   IAPT_LOSS_L2_SPLIT loss[0] =  mean d1, loss[1] = mean d2                (ChamferDistanceL2_split, :46-62)
 *   IAPT_LOSS_L1_PM    loss[0] =  mean sqrt d1                              (ChamferDistanceL1_PM,    :86-103)
 * iapt_chamfer_backward_fused: the gradient of that loss w.r.t. xyz1/xyz2 in ONE kernel - iapt_chamfer_backward with
 * grad_dist formed on the fly (upstream * 0.5/sqrt(d)/count resp. upstream/count), replacing the sqrt/mean autograd
 * chain; same workspace as iapt_chamfer_backward (required), either gradient may be NULL.  grad_loss points to the upstream gradient scalar(s) on the device (2 floats for L2_SPLIT, 1 otherwise;
 * NULL = 1).  A zero distance under L1 yields inf*0 = NaN exactly like the reference's sqrt backward. */
#define IAPT_LOSS_L1 1
#define IAPT_LOSS_L2 2
#define IAPT_LOSS_L2_SPLIT 3
#define IAPT_LOSS_L1_PM 4
int iapt_chamfer_loss(const float* dist1, const float* dist2, int B, int N, int M, int kind, float* sums, float* loss,
                      iapt_stream_t stream);
int iapt_chamfer_backward_fused(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                                const float* dist1, const float* dist2, int kind, const float* grad_loss, int B, int N,
                                int M, float* grad_xyz1, float* grad_xyz2, void* workspace, size_t workspace_bytes,
                                iapt_stream_t stream);

/* ---- the whole Chamfer step in ONE launch ------------------------------------------------------------------------
 * iapt_chamfer_step = iapt_chamfer_forward + iapt_chamfer_loss + iapt_fscore_counts + iapt_chamfer_backward_fused as a
 * single persistent kernel (plus one memset): what ChamferDistanceL1(pred, gt) followed by loss.backward() does in
 * the reference with 2+2 kernels, 6 memsets and the sqrt/mean autograd chain (extensions/chamfer_dist/__init__.py:
 * 13-25,64-84; chamfer.cu:147-229), and what Metrics.get does with two searches and two KD-tree passes
 * (utils/metrics.py:45-109).
 *   dist1/dist2/idx1/idx2  required outputs (as iapt_chamfer_forward);
 *   kind                   0 or IAPT_LOSS_*: which reduction `loss` [2] receives (loss may be NULL);
 *   sums [B,4], counts [B,2] (threshold fscore_th), valid [B,2]   optional per-cloud outputs (valid = number of
 *                          points whose coordinate sum is not zero, i.e. that survive `ignore_zeros`);
 *   grad_xyz1 / grad_xyz2  optional: gradient of the loss `kind` w.r.t. the clouds.  grad_loss = device pointer to the
 *                          upstream scalar(s) (2 floats for L2_SPLIT), NULL = 1: a caller that learns the upstream
 *                          gradient later (autograd) takes the unit gradient here and applies iapt_scale then.
 *   flags                  IAPT_STEP_IGNORE_ZERO_SUM: points with x+y+z == 0 neither query nor answer and are left out
 *                          of every sum and count - the `ignore_zeros` filter of chamfer_dist/__init__.py:37-41 without
 *                          compacting the clouds (dist/idx of such points are meaningless; not combinable with
 *                          gradients);  IAPT_STEP_GATE_ON_VALID: `valid` is an INPUT written by an earlier launch on the
 *                          same clouds - if it equals (N, M) for every cloud (nothing to ignore) the launch does
 *                          nothing and leaves its outputs untouched, so a caller can issue the `ignore_zeros` pass
 *                          unconditionally, without reading anything back;
 *                          IAPT_STEP_STATS: record device timestamps for iapt_chamfer_step_stats.
 * Per-cloud sums are folded in a fixed order that does not depend on B or the GPU: bit-reproducible.
 * iapt_chamfer_step_stats (synchronous, measurement only): stats_host[9] = globaltimer ns {first start of a search
 * item, kernel start, last end of a search item, kernel end, last end of the search / index-resolution / prefix / fill /
 * gather items} of the last IAPT_STEP_STATS launch on that workspace.
 * iapt_scale: dst[i] = src[i] * factor[0] (factor on the device) - applies the upstream gradient to a unit gradient. */
#define IAPT_STEP_IGNORE_ZERO_SUM 1
#define IAPT_STEP_STATS 2
#define IAPT_STEP_GATE_ON_VALID 4
size_t iapt_chamfer_step_workspace_bytes(int B, int N, int M);
int iapt_chamfer_step(const float* xyz1, const float* xyz2, int B, int N, int M, int kind, int flags, float fscore_th,
                      float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* sums, int32_t* counts,
                      int32_t* valid, float* loss, const float* grad_loss, float* grad_xyz1, float* grad_xyz2,
                      void* workspace, size_t workspace_bytes, iapt_stream_t stream);
int iapt_chamfer_step_stats(const void* workspace, unsigned long long* stats_host);
int iapt_scale(const float* src, const float* factor, long long n, float* dst, iapt_stream_t stream);

/* F-Score pieces (utils/metrics.py:67-90) from the same search: counts[b] = { #{j: sqrt(dist1[b,j]) < th},
 * #{k: sqrt(dist2[b,k]) < th} } as int32 [B,2]. */
int iapt_fscore_counts(const float* dist1, const float* dist2, int B, int N, int M, float th, int32_t* counts,
                       iapt_stream_t stream);

/* ---- Furthest point sampling -----------------------------------------------------------------------------
 * Replaces pointnet2_ops.pointnet2_utils.furthest_point_sample -> _ext.furthest_point_sampling
 * (third-party pointnet2_ops_lib, not vendored in the reference; call sites utils/misc.py:37,
 * models/base_blocks.py:550).   xyz [B,N,3] -> idx [B,npoint] int32.
 * Semantics: start at index 0, running min distance initialised to 1e10, d = fma(dz,dz,fma(dx,dx,dy*dy)),
 * next = argmax with the LOWEST index on ties; flag IAPT_FPS_SKIP_NEAR_ORIGIN reproduces upstream's
 * `if (|p|^2 <= 1e-3) continue;` (points that are never updated nor selected - PCN zero padding). */
#define IAPT_FPS_SKIP_NEAR_ORIGIN 1
/* Tie order among equal maxima of the running minimum.  Default: the LOWEST index.  IAPT_FPS_TIE_UPSTREAM selects
 * upstream's order as published (furthest_point_sampling_kernel<block>: thread-strided scan with strict `>`, shared-memory
 * tree keeping the lower thread) = the smallest (k mod block, k), block = largest power of two <= N, at most 512.  The two
 * differ only when points with DIFFERENT coordinates tie exactly - which happens on 4 of the reference's 8 demo clouds
 * (tests/test_oracle.py); upstream's binary is absent from the image, so either order is unpinned against it. */
#define IAPT_FPS_TIE_UPSTREAM 2
size_t iapt_fps_workspace_bytes(int B, int N, int npoint);
int iapt_fps(const float* xyz, int B, int N, int npoint, int flags, int32_t* idx, void* workspace,
             size_t workspace_bytes, iapt_stream_t stream);

/* Replaces pointnet2_utils.gather_operation -> _ext.gather_points / gather_points_grad
 * (call sites utils/misc.py:38, models/base_blocks.py:555-558).
 *   features [B,C,N], idx [B,npoint] -> out [B,C,npoint];  grad: grad_out [B,C,npoint] -> grad_features [B,C,N]
 *   (fully written, deterministic: one owner per output element, ascending source order). */
int iapt_gather(const float* features, const int32_t* idx, int B, int C, int N, int npoint, float* out,
                iapt_stream_t stream);
int iapt_gather_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int npoint,
                     float* grad_features, iapt_stream_t stream);
/* utils.misc.fps fused (utils/misc.py:32-40): xyz [B,N,3], idx [B,npoint] -> out [B,npoint,3] without the two
 * transpose().contiguous() copies the reference makes around gather_operation. */
int iapt_gather_xyz(const float* xyz, const int32_t* idx, int B, int N, int npoint, float* out,
                    iapt_stream_t stream);

/* ---- kNN grouping ----------------------------------------------------------------------------------------
 * Replaces models/Transformer_utils.py:19-30 knn_point(nsample, xyz, new_xyz) (= square_distance :32-51 +
 * torch.topk(largest=False, sorted=False)) without materialising the [B,Q,R] matrix.
 *   ref [B,R,3], query [B,Q,3] -> idx [B,Q,k] int64 ascending by (distance, index); dist [B,Q,k] optional (NULL).
 * 1 <= k <= 64 and k <= R. */
#define IAPT_KNN_MAX_K 64
/* Workspace: per-slice lists (few queries, many references: the loss kNN), or the curve-ordered copies of both clouds and
 * their tile boxes (k <= 16, 1024..4096 references, dense queries in enough warps: the encoder's 2048 x 2048 call), else 0.
 * Ask for every call; the answer follows the same dispatch as iapt_knn. */
size_t iapt_knn_workspace_bytes(int B, int R, int Q, int k);
int iapt_knn(const float* ref, const float* query, int B, int R, int Q, int k, int64_t* idx, float* dist,
             void* workspace, size_t workspace_bytes, iapt_stream_t stream);

/* Replaces models/Transformer_utils.py:32-51 square_distance(src, dst) for callers that want the matrix itself:
 *   src [B,N,C], dst [B,M,C] -> out [B,N,M] = -2 src.dst^T + |src|^2 + |dst|^2, the reference's expansion form in its
 * op order (so values agree with the reference's to FP32 rounding, NOT with the direct-form distances of iapt_knn /
 * iapt_chamfer_forward).  1 <= C <= 16. */
int iapt_square_distance(const float* src, const float* dst, int B, int N, int M, int C, float* out,
                         iapt_stream_t stream);

/* ---- kNN consumers -----------------------------------------------------------------
 * iapt_graph_feature replaces the gather / expand / cat after knn_point in DGCNN_Grouper.get_graph_feature
 * (models/base_blocks.py:582-588):   x_q [B,C,Nq], x_k [B,C,Nk], idx [B,Nq,k] int64 (as knn_point returns)  ->
 *   out [B,2C,Nq,k]:  out[b,c,q,j] = x_k[b,c,idx[b,q,j]] - x_q[b,c,q],   out[b,C+c,q,j] = x_q[b,c,q].
 * iapt_graph_feature_grad: grad_out [B,2C,Nq,k] -> grad_x_q [B,C,Nq] and grad_x_k [B,C,Nk] (either may be NULL),
 * deterministic (inverse index by a stable counting sort of idx, shared by all channels; no float atomics).
 * iapt_index_points replaces index_points (models/Transformer_utils.py:53-69): points [B,N,C], idx [B,S] int64 (S may be
 * the flattened Q*k) -> out [B,S,C].
 * Preconditions of every idx-consuming entry point: 0 <= idx < the indexed extent (the reference's advanced indexing
 * raises a device assert otherwise; here an out-of-range index is undefined behaviour), and the number of positions per
 * cloud (Nq*k, P*S, S) fits in an int (checked: IAPT_EINVAL). */
int iapt_graph_feature(const float* x_q, const float* x_k, const int64_t* idx, int B, int C, int Nq, int Nk, int k,
                       float* out, iapt_stream_t stream);
size_t iapt_graph_feature_grad_workspace_bytes(int B, int Nq, int Nk, int k);
int iapt_graph_feature_grad(const float* grad_out, const int64_t* idx, int B, int C, int Nq, int Nk, int k,
                            float* grad_x_q, float* grad_x_k, void* workspace, size_t workspace_bytes,
                            iapt_stream_t stream);
int iapt_index_points(const float* points, const int64_t* idx, int B, int N, int C, long long S, float* out,
                      iapt_stream_t stream);
/* Gradient of iapt_index_points w.r.t. points: grad_out [B,S,C] -> grad_points [B,N,C], each row the sum, in ascending
 * position, of the grad_out rows that gathered it (deterministic; torch's index backward is an atomicAdd scatter).
 * workspace: iapt_graph_feature_grad_workspace_bytes(B, S, N, 1). */
int iapt_index_points_grad(const float* grad_out, const int64_t* idx, int B, int N, int C, long long S,
                           float* grad_points, void* workspace, size_t workspace_bytes, iapt_stream_t stream);
/* Channels-last graph feature of DynamicGraphAttention (models/Transformer_utils.py:814-816 and :842-855):
 *   q [B,N,C], v [B,M,C], idx [B,N,k] int64  ->  out [B,N,k,2C] = cat(v[b,idx[b,n,j],:] - q[b,n,:], q[b,n,:])
 * replacing index_points + unsqueeze/expand + sub + cat.  _grad: grad_out [B,N,k,2C] -> grad_q [B,N,C], grad_v [B,M,C]
 * (either may be NULL); workspace (grad_v only): iapt_graph_feature_grad_workspace_bytes(B, N, M, k). */
int iapt_graph_feature_cl(const float* q, const float* v, const int64_t* idx, int B, int N, int M, int C, int k,
                          float* out, iapt_stream_t stream);
int iapt_graph_feature_cl_grad(const float* grad_out, const int64_t* idx, int B, int N, int M, int C, int k,
                               float* grad_q, float* grad_v, void* workspace, size_t workspace_bytes,
                               iapt_stream_t stream);

/* ---- the remaining pointnet2_ops entry points the reference imports (cold paths:
 * Deformable* blocks, models/Transformer_utils.py:228-232; SnowFlakeNet, models/external/SnowFlakeNet_utils.py:4,115-120) --
 * iapt_group_gather: features [B,C,M], idx [B,P,S] int64, weight [B,P,S] (may be NULL for reduce=0)
 *   reduce = 0 -> out [B,C,P,S] = features[b,c,idx]   (pointnet2_utils.grouping_operation)
 *   reduce = 1 -> out [B,C,P]   = sum_s weight * features[b,c,idx]   (pointnet2_utils.three_interpolate, S = 3)
 * iapt_group_gather_grad: gradient w.r.t. features, deterministic (same inverse-index scheme and workspace size as
 *   iapt_graph_feature_grad: iapt_graph_feature_grad_workspace_bytes(B, P, M, S)).
 * iapt_ball_query (pointnet2_utils.ball_query): xyz [B,N,3], new_xyz [B,P,3] -> idx [B,P,nsample] int32, the first
 *   nsample points in ascending index with squared distance < radius^2, padded with the first hit (zeros if none).
 * three_nn needs no entry point of its own: it is iapt_knn with k = 3 plus a square root (host mirror). */
int iapt_group_gather(const float* features, const int64_t* idx, const float* weight, int B, int C, int M, int P, int S,
                      int reduce, float* out, iapt_stream_t stream);
int iapt_group_gather_grad(const float* grad_out, const int64_t* idx, const float* weight, int B, int C, int M, int P,
                           int S, int reduce, float* grad_features, void* workspace, size_t workspace_bytes,
                           iapt_stream_t stream);
int iapt_ball_query(const float* xyz, const float* new_xyz, int B, int N, int P, float radius, int nsample, int32_t* idx,
                    iapt_stream_t stream);

/* ---- EMD approximation by auction ---------------------------------------------------
 * Replaces extensions/emd: emd.forward / emd.backward (emd.cpp:14-30, emd_cuda.cu:230-308) behind emdFunction
 * (emd_module.py:31-72), the test-time EMDistance metric (utils/metrics.py:111-116, eps = 0.005, 100 iterations).
 *   xyz1 [B,n,3] predictions, xyz2 [B,n,3] targets (same size, coordinates in [0,1] as upstream assumes)
 *   -> dist [B,n] squared distance of every prediction to its assigned target, assignment [B,n] int32.
 * The result is an approximation and the assignment need not be a bijection (as upstream).  Where the reference leaves
 * a choice to thread scheduling it is fixed: equal best values keep the lowest target index, of several bidders within
 * 1e-6 of a target's best increment the highest prediction index wins - so the output is deterministic.
 * No restriction of n to multiples of 1024 or of B to 512.  backward: grad_xyz1 [B,n,3] (xyz2 receives none upstream). */
size_t iapt_emd_workspace_bytes(int B, int n);
int iapt_emd_forward(const float* xyz1, const float* xyz2, int B, int n, float eps, int iters, float* dist,
                     int32_t* assignment, void* workspace, size_t workspace_bytes, iapt_stream_t stream);
int iapt_emd_backward(const float* xyz1, const float* xyz2, const float* grad_dist, const int32_t* assignment, int B,
                      int n, float* grad_xyz1, iapt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IAPT_B200_H_ */
