// c_api.cu - the extern "C" surface declared in include/iapt_b200.h: argument validation + launch.
#include <atomic>
#include <cstdio>

#include "../../include/iapt_b200.h"
#include "chamfer_step.h"
#include "common.cuh"

namespace iapt {
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int chamfer_metrics_launch(const float*, const float*, int, int, int, float, float*, int32_t*, cudaStream_t);
int chamfer_loss_launch(const float*, const float*, int, int, int, int, float*, float*, cudaStream_t);
size_t fps_workspace_bytes(int B, int N, int npoint);
int fps_launch(const float*, int, int, int, int, int32_t*, void*, cudaStream_t);
int gather_launch(const float*, const int32_t*, int, int, int, int, float*, cudaStream_t);
int gather_xyz_launch(const float*, const int32_t*, int, int, int, float*, cudaStream_t);
int gather_grad_launch(const float*, const int32_t*, int, int, int, int, float*, cudaStream_t);
size_t knn_workspace_bytes(int B, int R, int Q, int k);
int knn_launch(const float*, const float*, int, int, int, int, int64_t*, float*, void*, cudaStream_t);
int square_distance_launch(const float*, const float*, int, int, int, int, float*, cudaStream_t);
int graph_feature_launch(const float*, const float*, const long long*, int, int, int, int, int, float*, cudaStream_t);
int index_points_launch(const float*, const long long*, int, int, int, long long, float*, cudaStream_t);
size_t graph_feature_grad_workspace_bytes(int B, int Nq, int Nk, int k);
int group_gather_launch(const float*, const long long*, const float*, int, int, int, int, int, int, float*, cudaStream_t);
int group_gather_grad_launch(const float*, const long long*, const float*, int, int, int, int, int, int, float*, void*,
                             cudaStream_t);
int ball_query_launch(const float*, const float*, int, int, int, float, int, int*, cudaStream_t);
int graph_feature_grad_launch(const float*, const long long*, int, int, int, int, int, float*, float*, void*, cudaStream_t);
int graph_feature_cl_launch(const float*, const float*, const long long*, int, int, int, int, int, float*, cudaStream_t);
int graph_feature_cl_grad_launch(const float*, const long long*, int, int, int, int, int, float*, float*, void*,
                                 cudaStream_t);
int index_points_grad_launch(const float*, const long long*, int, int, int, long long, float*, void*, cudaStream_t);
size_t emd_workspace_bytes(int B, int n);
int emd_forward_launch(const float*, const float*, int, int, float, int, float*, int*, void*, cudaStream_t);
int emd_backward_launch(const float*, const float*, const float*, const int*, int, int, float*, cudaStream_t);
}  // namespace iapt

using namespace iapt;

#define IAPT_REQUIRE(cond) \
  do {                     \
    if (!(cond)) return IAPT_EINVAL; \
  } while (0)

extern "C" {

int iapt_abi_version(void) { return IAPT_ABI_VERSION; }

const char* iapt_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case IAPT_EINVAL: return "iapt: invalid argument (null pointer or bad size)";
    case IAPT_EWORKSPACE: return "iapt: workspace too small (see *_workspace_bytes)";
    case IAPT_EUNSUPPORTED: return "iapt: unsupported configuration";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "iapt: unknown error";
  }
}

unsigned long long iapt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int iapt_profile_enable(int on) {
  profile_enable(on != 0);
  return 0;
}

int iapt_profile_collect(double* total_ms, int* launches) { return profile_collect(total_ms, launches); }

size_t iapt_chamfer_forward_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  return chamfer_forward_workspace_bytes(B, N, M);
}

int iapt_chamfer_forward(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2,
                         int32_t* idx1, int32_t* idx2, void* workspace, size_t workspace_bytes,
                         iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 0 && M >= 0);
  IAPT_REQUIRE((dist1 == nullptr) == (idx1 == nullptr) && (dist2 == nullptr) == (idx2 == nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return 0;
  if (N == 0 || M == 0) {  // reference: torch::zeros outputs, kernel writes nothing (chamfer.cu:152-157)
    cudaError_t e = cudaSuccess;
    if (dist1 && N > 0) {
      e = cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)B * N, st);
      if (e == cudaSuccess) e = cudaMemsetAsync(idx1, 0, sizeof(int32_t) * (size_t)B * N, st);
    }
    if (e == cudaSuccess && dist2 && M > 0) {
      e = cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)B * M, st);
      if (e == cudaSuccess) e = cudaMemsetAsync(idx2, 0, sizeof(int32_t) * (size_t)B * M, st);
    }
    return (int)e;
  }
  IAPT_REQUIRE(xyz1 && xyz2 && workspace);
  IAPT_REQUIRE((long long)B * N < (1LL << 31) && (long long)B * M < (1LL << 31));
  if (workspace_bytes < chamfer_forward_workspace_bytes(B, N, M)) return IAPT_EWORKSPACE;
  StepArgs a;
  a.xyz1 = xyz1;
  a.xyz2 = xyz2;
  a.B = B;
  a.N = N;
  a.M = M;
  a.dist1 = dist1;
  a.dist2 = dist2;
  a.idx1 = idx1;
  a.idx2 = idx2;
  a.workspace = workspace;
  return chamfer_step_launch(a, st);
}

// (mode, scale1, scale2, stride of the upstream scalars) of the analytic loss gradient, extensions/chamfer_dist/
// __init__.py:43-44,61-62,82-84,101-103: d mean(x) = 1/count, d sqrt(d) = 0.5/sqrt(d).
static void loss_gradient_weights(int kind, int B, int N, int M, int* mode, float* s1, float* s2, int* stride) {
  const float n1 = (float)B * (float)N, n2 = (float)B * (float)M;
  *stride = 0;
  switch (kind) {
    case IAPT_LOSS_L1: *mode = 1; *s1 = 0.25f / n1; *s2 = 0.25f / n2; break;  // (q1+q2)/2
    case IAPT_LOSS_L2: *mode = 2; *s1 = 1.f / n1; *s2 = 1.f / n2; break;
    case IAPT_LOSS_L2_SPLIT: *mode = 2; *s1 = 1.f / n1; *s2 = 1.f / n2; *stride = 1; break;
    default: *mode = 1; *s1 = 0.5f / n1; *s2 = 0.f; break;                    // L1_PM: mean sqrt(dist1) only
  }
}

size_t iapt_chamfer_step_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  return chamfer_step_workspace_bytes(B, N, M);
}

int iapt_chamfer_step(const float* xyz1, const float* xyz2, int B, int N, int M, int kind, int flags, float fscore_th,
                      float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* sums, int32_t* counts,
                      int32_t* valid, float* loss, const float* grad_loss, float* grad_xyz1, float* grad_xyz2,
                      void* workspace, size_t workspace_bytes, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 1 && M >= 1 && kind >= 0 && kind <= IAPT_LOSS_L1_PM);
  if (B == 0) return 0;
  IAPT_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2 && workspace);
  IAPT_REQUIRE((long long)B * N < (1LL << 31) && (long long)B * M < (1LL << 31));
  IAPT_REQUIRE(loss == nullptr || kind != 0);
  IAPT_REQUIRE(!(flags & IAPT_STEP_GATE_ON_VALID) || valid != nullptr);
  const bool grads = grad_xyz1 != nullptr || grad_xyz2 != nullptr;
  if (grads) {
    IAPT_REQUIRE(kind != 0);
    // the two outputs of L2_split carry independent upstream gradients: they must be given
    if (kind == IAPT_LOSS_L2_SPLIT && grad_loss == nullptr) return IAPT_EUNSUPPORTED;
    // the mean over the unmasked points is only known on the device
    if (flags & IAPT_STEP_IGNORE_ZERO_SUM) return IAPT_EUNSUPPORTED;
  }
  if (workspace_bytes < chamfer_step_workspace_bytes(B, N, M)) return IAPT_EWORKSPACE;
  StepArgs a;
  a.xyz1 = xyz1;
  a.xyz2 = xyz2;
  a.B = B;
  a.N = N;
  a.M = M;
  a.mask_zero_sum = (flags & IAPT_STEP_IGNORE_ZERO_SUM) != 0;
  a.stats = (flags & IAPT_STEP_STATS) != 0;
  a.dist1 = dist1;
  a.dist2 = dist2;
  a.idx1 = idx1;
  a.idx2 = idx2;
  a.th = fscore_th;
  a.kind = kind;
  a.sums = sums;
  a.counts = counts;
  a.valid = valid;
  a.valid_is_input = (flags & IAPT_STEP_GATE_ON_VALID) != 0;
  a.loss = loss;
  if (grads) {
    loss_gradient_weights(kind, B, N, M, &a.mode, &a.scale1, &a.scale2, &a.grad_out_stride);
    a.grad_out = grad_loss;
    a.gw1 = a.scale1 != 0.f ? dist1 : nullptr;
    a.gw2 = a.scale2 != 0.f ? dist2 : nullptr;
    a.grad_xyz1 = grad_xyz1;
    a.grad_xyz2 = grad_xyz2;
  }
  a.workspace = workspace;
  return chamfer_step_launch(a, (cudaStream_t)stream);
}

int iapt_chamfer_step_stats(const void* workspace, unsigned long long* stats_host) {
  IAPT_REQUIRE(workspace && stats_host);
  return (int)cudaMemcpy(stats_host, workspace, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
}

int iapt_scale(const float* src, const float* factor, long long n, float* dst, iapt_stream_t stream) {
  IAPT_REQUIRE(n >= 0);
  if (n == 0) return 0;
  IAPT_REQUIRE(src && factor && dst);
  return scale_launch(src, factor, n, dst, (cudaStream_t)stream);
}

size_t iapt_chamfer_backward_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  return chamfer_backward_workspace_bytes(B, N, M);
}

static int backward_common(StepArgs& a, const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                           int B, int N, int M, float* grad_xyz1, float* grad_xyz2, void* workspace,
                           size_t workspace_bytes, cudaStream_t st) {
  if (workspace == nullptr || workspace_bytes < chamfer_backward_workspace_bytes(B, N, M)) return IAPT_EWORKSPACE;
  a.xyz1 = xyz1;
  a.xyz2 = xyz2;
  a.B = B;
  a.N = N;
  a.M = M;
  a.search = false;
  a.idx1 = const_cast<int32_t*>(idx1);  // inputs in a backward-only launch
  a.idx2 = const_cast<int32_t*>(idx2);
  a.grad_xyz1 = grad_xyz1;
  a.grad_xyz2 = grad_xyz2;
  a.workspace = workspace;
  return chamfer_step_launch(a, st);
}

int iapt_chamfer_backward(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                          const float* grad_dist1, const float* grad_dist2, int B, int N, int M, float* grad_xyz1,
                          float* grad_xyz2, void* workspace, size_t workspace_bytes, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 0 && M >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return 0;
  if (N == 0 || M == 0) {
    cudaError_t e = cudaSuccess;
    if (N > 0 && grad_xyz1) e = cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * 3 * (size_t)B * N, st);
    if (e == cudaSuccess && M > 0 && grad_xyz2) e = cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * 3 * (size_t)B * M, st);
    return (int)e;
  }
  IAPT_REQUIRE(xyz1 && xyz2 && idx1 && idx2 && (grad_xyz1 || grad_xyz2));
  IAPT_REQUIRE((long long)B * N < (1LL << 31) && (long long)B * M < (1LL << 31));
  StepArgs a;
  a.mode = 0;
  a.gw1 = grad_dist1;
  a.gw2 = grad_dist2;
  return backward_common(a, xyz1, xyz2, idx1, idx2, B, N, M, grad_xyz1, grad_xyz2, workspace, workspace_bytes, st);
}

int iapt_chamfer_loss(const float* dist1, const float* dist2, int B, int N, int M, int kind, float* sums, float* loss,
                      iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 1 && N >= 1 && M >= 1 && kind >= IAPT_LOSS_L1 && kind <= IAPT_LOSS_L1_PM);
  IAPT_REQUIRE(dist1 && dist2 && sums && loss);
  return chamfer_loss_launch(dist1, dist2, B, N, M, kind, sums, loss, (cudaStream_t)stream);
}

int iapt_chamfer_backward_fused(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                                const float* dist1, const float* dist2, int kind, const float* grad_loss, int B, int N,
                                int M, float* grad_xyz1, float* grad_xyz2, void* workspace, size_t workspace_bytes,
                                iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 1 && N >= 1 && M >= 1 && kind >= IAPT_LOSS_L1 && kind <= IAPT_LOSS_L1_PM);
  IAPT_REQUIRE(xyz1 && xyz2 && idx1 && idx2 && dist1 && dist2 && (grad_xyz1 || grad_xyz2));
  IAPT_REQUIRE((long long)B * N < (1LL << 31) && (long long)B * M < (1LL << 31));
  StepArgs a;
  loss_gradient_weights(kind, B, N, M, &a.mode, &a.scale1, &a.scale2, &a.grad_out_stride);
  a.grad_out = grad_loss;
  a.gw1 = a.scale1 != 0.f ? dist1 : nullptr;
  a.gw2 = a.scale2 != 0.f ? dist2 : nullptr;
  return backward_common(a, xyz1, xyz2, idx1, idx2, B, N, M, grad_xyz1, grad_xyz2, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

int iapt_chamfer_reduce(const float* dist1, const float* dist2, int B, int N, int M, float* sums,
                        iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 0 && M >= 0);
  if (B == 0) return 0;
  IAPT_REQUIRE(dist1 && dist2 && sums);
  return chamfer_metrics_launch(dist1, dist2, B, N, M, 0.f, sums, nullptr, (cudaStream_t)stream);
}

int iapt_fscore_counts(const float* dist1, const float* dist2, int B, int N, int M, float th, int32_t* counts,
                       iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 0 && M >= 0);
  if (B == 0) return 0;
  IAPT_REQUIRE(dist1 && dist2 && counts);
  return chamfer_metrics_launch(dist1, dist2, B, N, M, th, nullptr, counts, (cudaStream_t)stream);
}

size_t iapt_fps_workspace_bytes(int B, int N, int npoint) {
  if (B <= 0 || N <= 0) return 0;
  return fps_workspace_bytes(B, N, npoint);
}

int iapt_fps(const float* xyz, int B, int N, int npoint, int flags, int32_t* idx, void* workspace,
             size_t workspace_bytes, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 1 && npoint >= 0);
  if (B == 0 || npoint == 0) return 0;
  IAPT_REQUIRE(xyz && idx);
  const size_t need = fps_workspace_bytes(B, N, npoint);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) return IAPT_EWORKSPACE;
  return fps_launch(xyz, B, N, npoint, flags, idx, workspace, (cudaStream_t)stream);
}

int iapt_gather(const float* features, const int32_t* idx, int B, int C, int N, int npoint, float* out,
                iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && C >= 0 && N >= 1 && npoint >= 0);
  if (B == 0 || C == 0 || npoint == 0) return 0;
  IAPT_REQUIRE(features && idx && out);
  return gather_launch(features, idx, B, C, N, npoint, out, (cudaStream_t)stream);
}

int iapt_gather_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int npoint, float* grad_features,
                     iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && C >= 0 && N >= 1 && npoint >= 0);
  if (B == 0 || C == 0) return 0;
  IAPT_REQUIRE(grad_features);
  if (npoint == 0) return (int)cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * N, (cudaStream_t)stream);
  IAPT_REQUIRE(grad_out && idx);
  IAPT_REQUIRE(B <= 65535);
  return gather_grad_launch(grad_out, idx, B, C, N, npoint, grad_features, (cudaStream_t)stream);
}

int iapt_gather_xyz(const float* xyz, const int32_t* idx, int B, int N, int npoint, float* out,
                    iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 1 && npoint >= 0);
  if (B == 0 || npoint == 0) return 0;
  IAPT_REQUIRE(xyz && idx && out);
  return gather_xyz_launch(xyz, idx, B, N, npoint, out, (cudaStream_t)stream);
}

size_t iapt_knn_workspace_bytes(int B, int R, int Q, int k) {
  if (B <= 0 || R <= 0 || Q <= 0 || k <= 0) return 0;
  return knn_workspace_bytes(B, R, Q, k);
}

int iapt_knn(const float* ref, const float* query, int B, int R, int Q, int k, int64_t* idx, float* dist,
             void* workspace, size_t workspace_bytes, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && R >= 1 && Q >= 0 && k >= 1);
  if (k > IAPT_KNN_MAX_K) return IAPT_EUNSUPPORTED;
  IAPT_REQUIRE(k <= R);
  if (B == 0 || Q == 0) return 0;
  IAPT_REQUIRE(ref && query && idx);
  const size_t need = knn_workspace_bytes(B, R, Q, k);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) return IAPT_EWORKSPACE;
  return knn_launch(ref, query, B, R, Q, k, idx, dist, need ? workspace : nullptr, (cudaStream_t)stream);
}

int iapt_square_distance(const float* src, const float* dst, int B, int N, int M, int C, float* out,
                         iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 0 && M >= 0 && C >= 1 && B <= 65535 && (N + 7) / 8 <= 65535);
  if (B == 0 || N == 0 || M == 0) return 0;
  IAPT_REQUIRE(src && dst && out);
  return square_distance_launch(src, dst, B, N, M, C, out, (cudaStream_t)stream);
}

int iapt_graph_feature(const float* x_q, const float* x_k, const int64_t* idx, int B, int C, int Nq, int Nk, int k,
                       float* out, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && C >= 0 && Nq >= 0 && Nk >= 1 && k >= 0 && B <= 65535);
  if (B == 0 || C == 0 || Nq == 0 || k == 0) return 0;
  IAPT_REQUIRE(x_q && x_k && idx && out);
  return graph_feature_launch(x_q, x_k, (const long long*)idx, B, C, Nq, Nk, k, out, (cudaStream_t)stream);
}

size_t iapt_graph_feature_grad_workspace_bytes(int B, int Nq, int Nk, int k) {
  if (B <= 0 || Nq <= 0 || Nk <= 0 || k <= 0) return 0;
  return graph_feature_grad_workspace_bytes(B, Nq, Nk, k);
}

int iapt_graph_feature_grad(const float* grad_out, const int64_t* idx, int B, int C, int Nq, int Nk, int k,
                            float* grad_x_q, float* grad_x_k, void* workspace, size_t workspace_bytes,
                            iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && C >= 0 && Nq >= 0 && Nk >= 1 && k >= 0 && B <= 65535 && C <= 65535 &&
               (long long)Nq * k < (1LL << 31));  // positions are ints inside the inverse index
  if (B == 0 || C == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (Nq == 0 || k == 0) {
    cudaError_t e = cudaSuccess;
    if (grad_x_k) e = cudaMemsetAsync(grad_x_k, 0, sizeof(float) * (size_t)B * C * Nk, st);
    return (int)e;
  }
  IAPT_REQUIRE(grad_out && idx);
  if (grad_x_k && (workspace == nullptr || workspace_bytes < graph_feature_grad_workspace_bytes(B, Nq, Nk, k)))
    return IAPT_EWORKSPACE;
  return graph_feature_grad_launch(grad_out, (const long long*)idx, B, C, Nq, Nk, k, grad_x_q, grad_x_k, workspace, st);
}

int iapt_index_points(const float* points, const int64_t* idx, int B, int N, int C, long long S, float* out,
                      iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 1 && C >= 0 && S >= 0);
  if (B == 0 || C == 0 || S == 0) return 0;
  IAPT_REQUIRE(points && idx && out);
  return index_points_launch(points, (const long long*)idx, B, N, C, S, out, (cudaStream_t)stream);
}

int iapt_index_points_grad(const float* grad_out, const int64_t* idx, int B, int N, int C, long long S,
                           float* grad_points, void* workspace, size_t workspace_bytes, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 1 && C >= 0 && S >= 0 && S < (1LL << 31));
  if (B == 0 || C == 0) return 0;
  IAPT_REQUIRE(grad_points);
  if (S == 0) return (int)cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * N * C, (cudaStream_t)stream);
  IAPT_REQUIRE(grad_out && idx);
  if (workspace == nullptr || workspace_bytes < graph_feature_grad_workspace_bytes(B, (int)S, N, 1)) return IAPT_EWORKSPACE;
  return index_points_grad_launch(grad_out, (const long long*)idx, B, N, C, S, grad_points, workspace,
                                  (cudaStream_t)stream);
}

int iapt_graph_feature_cl(const float* q, const float* v, const int64_t* idx, int B, int N, int M, int C, int k,
                          float* out, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 0 && M >= 1 && C >= 0 && k >= 0);
  if (B == 0 || N == 0 || C == 0 || k == 0) return 0;
  IAPT_REQUIRE(q && v && idx && out);
  return graph_feature_cl_launch(q, v, (const long long*)idx, B, N, M, C, k, out, (cudaStream_t)stream);
}

int iapt_graph_feature_cl_grad(const float* grad_out, const int64_t* idx, int B, int N, int M, int C, int k,
                               float* grad_q, float* grad_v, void* workspace, size_t workspace_bytes,
                               iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 0 && M >= 1 && C >= 0 && k >= 0 && (long long)N * k < (1LL << 31));
  if (B == 0 || C == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0 || k == 0) return grad_v ? (int)cudaMemsetAsync(grad_v, 0, sizeof(float) * (size_t)B * M * C, st) : 0;
  IAPT_REQUIRE(grad_out && idx);
  if (grad_v && (workspace == nullptr || workspace_bytes < graph_feature_grad_workspace_bytes(B, N, M, k)))
    return IAPT_EWORKSPACE;
  return graph_feature_cl_grad_launch(grad_out, (const long long*)idx, B, N, M, C, k, grad_q, grad_v, workspace, st);
}

int iapt_group_gather(const float* features, const int64_t* idx, const float* weight, int B, int C, int M, int P, int S,
                      int reduce, float* out, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && C >= 0 && M >= 1 && P >= 0 && S >= 1 && (reduce == 0 || reduce == 1));
  if (B == 0 || C == 0 || P == 0) return 0;
  IAPT_REQUIRE(features && idx && out && (reduce == 0 || weight));
  return group_gather_launch(features, (const long long*)idx, weight, B, C, M, P, S, reduce, out, (cudaStream_t)stream);
}

int iapt_group_gather_grad(const float* grad_out, const int64_t* idx, const float* weight, int B, int C, int M, int P,
                           int S, int reduce, float* grad_features, void* workspace, size_t workspace_bytes,
                           iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && C >= 0 && M >= 1 && P >= 0 && S >= 1 && (reduce == 0 || reduce == 1) && B <= 65535 && C <= 65535 &&
               (long long)P * S < (1LL << 31));
  if (B == 0 || C == 0) return 0;
  IAPT_REQUIRE(grad_features);
  if (P == 0)
    return (int)cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * M, (cudaStream_t)stream);
  IAPT_REQUIRE(grad_out && idx && (reduce == 0 || weight));
  if (workspace == nullptr || workspace_bytes < graph_feature_grad_workspace_bytes(B, P, M, S)) return IAPT_EWORKSPACE;
  return group_gather_grad_launch(grad_out, (const long long*)idx, weight, B, C, M, P, S, reduce, grad_features,
                                  workspace, (cudaStream_t)stream);
}

int iapt_ball_query(const float* xyz, const float* new_xyz, int B, int N, int P, float radius, int nsample, int32_t* idx,
                    iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && N >= 1 && P >= 0 && nsample >= 1 && radius >= 0.f);
  if (B == 0 || P == 0) return 0;
  IAPT_REQUIRE(xyz && new_xyz && idx);
  return ball_query_launch(xyz, new_xyz, B, N, P, radius, nsample, idx, (cudaStream_t)stream);
}

size_t iapt_emd_workspace_bytes(int B, int n) {
  if (B <= 0 || n <= 0) return 0;
  return emd_workspace_bytes(B, n);
}

int iapt_emd_forward(const float* xyz1, const float* xyz2, int B, int n, float eps, int iters, float* dist,
                     int32_t* assignment, void* workspace, size_t workspace_bytes, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && n >= 0 && iters >= 1 && B <= 65535 && eps >= 0.f);
  if (B == 0 || n == 0) return 0;
  IAPT_REQUIRE(xyz1 && xyz2 && dist && assignment);
  if (workspace == nullptr || workspace_bytes < emd_workspace_bytes(B, n)) return IAPT_EWORKSPACE;
  return emd_forward_launch(xyz1, xyz2, B, n, eps, iters, dist, assignment, workspace, (cudaStream_t)stream);
}

int iapt_emd_backward(const float* xyz1, const float* xyz2, const float* grad_dist, const int32_t* assignment, int B,
                      int n, float* grad_xyz1, iapt_stream_t stream) {
  IAPT_REQUIRE(B >= 0 && n >= 0);
  if (B == 0 || n == 0) return 0;
  IAPT_REQUIRE(xyz1 && xyz2 && grad_dist && assignment && grad_xyz1);
  return emd_backward_launch(xyz1, xyz2, grad_dist, assignment, B, n, grad_xyz1, (cudaStream_t)stream);
}

}  // extern "C"
