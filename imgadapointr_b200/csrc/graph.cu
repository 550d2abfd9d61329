// graph.cu - kNN consumers for sm_100a.
//
// Replaces the tensor gymnastics after knn_point in the reference:
//   * DGCNN_Grouper.get_graph_feature (models/base_blocks.py:582-588): transpose x_k, flat fancy-index by idx + batch
//     base, view/permute/contiguous, expand x_q, cat((feature - x_q, x_q)) - five passes over a [B,2C,Q,k] tensor;
//   * index_points (models/Transformer_utils.py:53-69): batch-index arange/repeat + advanced indexing.
// Here: ONE kernel writes [B,2C,Q,k] directly (HBM-bound on the output: 8*B*C*Q*k bytes), and a deterministic
// backward: grad_x_q is a fixed-order sum over k; grad_x_k is owner-computes through a per-cloud inverse index
// (stable counting sort of idx built once and shared by all C channels) - no floating-point atomics, whereas the
// reference's index backward accumulates with atomicAdd in run-dependent order.
#include <cstdint>

#include "common.cuh"

namespace iapt {

// ---- forward: out[b,c,q,j] = x_k[b,c,idx[b,q,j]] - x_q[b,c,q];  out[b,C+c,q,j] = x_q[b,c,q] ---------------------
// A thread owns kVec consecutive (q, j) positions (kVec = 4 when k % 4 == 0: one 128-bit store per channel half) and
// kGfChannelsPerCta channels; all gathers of a thread are independent, so 8 * kVec loads are in flight per thread.
constexpr int kGfThreads = 256;
constexpr int kGfChannelsPerCta = 4;

template <int kVec>
__global__ void __launch_bounds__(kGfThreads) graph_feature_kernel(const float* __restrict__ xq,
                                                                   const float* __restrict__ xk,
                                                                   const long long* __restrict__ idx, int C, int Nq,
                                                                   int Nk, int k, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kGfChannelsPerCta;
  const long long L = (long long)Nq * k;
  const long long pos = ((long long)blockIdx.x * kGfThreads + threadIdx.x) * kVec;  // flattened (q, j)
  if (pos >= L) return;
  const int q = (int)(pos / k);  // kVec divides k: all kVec positions share q
  int n[kVec];
#pragma unroll
  for (int v = 0; v < kVec; ++v) n[v] = (int)__ldg(idx + (size_t)b * L + pos + v);
#pragma unroll
  for (int cc = 0; cc < kGfChannelsPerCta; ++cc) {
    const int c = c0 + cc;
    if (c >= C) break;
    const float vq = __ldg(xq + ((size_t)b * C + c) * Nq + q);
    const float* row = xk + ((size_t)b * C + c) * Nk;
    float d[kVec];
#pragma unroll
    for (int v = 0; v < kVec; ++v) d[v] = __fsub_rn(__ldg(row + n[v]), vq);
    float* lo = out + ((size_t)b * 2 * C + c) * L + pos;
    float* hi = out + ((size_t)b * 2 * C + C + c) * L + pos;
    if (kVec == 4) {
      *reinterpret_cast<float4*>(lo) = make_float4(d[0], d[1], d[2], d[3]);
      *reinterpret_cast<float4*>(hi) = make_float4(vq, vq, vq, vq);
    } else {
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        lo[v] = d[v];
        hi[v] = vq;
      }
    }
  }
}

// ---- index_points: out[b,s,:] = points[b,idx[b,s],:]  (points [B,N,C], idx [B,S] int64, any S incl. S = Q*k) -----
__global__ void __launch_bounds__(256) index_points_kernel(const float* __restrict__ pts,
                                                           const long long* __restrict__ idx, int N, int C,
                                                           long long S, float* __restrict__ out, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const long long bs = t / C;
    const long long b = bs / S;
    out[t] = __ldg(pts + ((size_t)b * N + __ldg(idx + bs)) * C + c);
  }
}

// ---- inverse index (per cloud): start[b, 0..Nk], perm[b, 0..L) = positions p sorted by (idx[p], p) -----------------
// Step 1: counts by integer atomics (order independent -> deterministic).
__global__ void __launch_bounds__(256) inv_count_kernel(const long long* __restrict__ idx, long long L, int Nk,
                                                        int* __restrict__ counts) {
  const int b = blockIdx.y;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < L; p += (long long)gridDim.x * blockDim.x)
    atomicAdd(counts + (size_t)b * (Nk + 1) + (int)__ldg(idx + (size_t)b * L + p), 1);
}

// Step 2: exclusive scan of counts[b, :] in place (one CTA per cloud, fixed order).
__global__ void __launch_bounds__(1024) inv_scan_kernel(int* __restrict__ counts, int Nk) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  int* c = counts + (size_t)blockIdx.x * (Nk + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base <= Nk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < Nk) ? c[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;  // inclusive
    }
    __syncthreads();
    const int before = carry + (warp ? warp_tot[warp - 1] : 0) + s - v;
    if (i <= Nk) c[i] = before;  // c[Nk] receives the grand total on the last chunk
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
}

// Step 3: stable fill.  A warp owns kInvBins consecutive bins with their write cursors in shared memory and scans
// all L keys in ascending position; same-bin hits inside a 32-key step take consecutive slots in lane order.
constexpr int kInvBins = 64;
__global__ void __launch_bounds__(256) inv_fill_kernel(const long long* __restrict__ idx, long long L, int Nk,
                                                       const int* __restrict__ start, int* __restrict__ perm) {
  __shared__ int cursor[8][kInvBins];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int j0 = (blockIdx.x * 8 + warp) * kInvBins;
  if (j0 >= Nk) return;
  const int* st = start + (size_t)b * (Nk + 1);
  int* cur = cursor[warp];
  for (int i = lane; i < kInvBins; i += 32) cur[i] = (j0 + i < Nk) ? st[j0 + i] : 0;
  __syncwarp();
  const long long* key = idx + (size_t)b * L;
  int* out = perm + (size_t)b * L;
  const unsigned lt = (1u << lane) - 1u;
  for (long long p0 = 0; p0 < L; p0 += 256) {
    int kk[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long p = p0 + e * 32 + lane;
      kk[e] = (p < L) ? (int)__ldg(key + p) : -1;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned bin = (unsigned)(kk[e] - j0);
      const bool hit = kk[e] >= 0 && bin < (unsigned)kInvBins;
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask == 0u) continue;
      const int first = hit ? cur[bin] : 0;
      __syncwarp();  // every peer has read the cursor before one of them advances it
      if (hit) {
        const unsigned peers = __match_any_sync(mask, bin);
        const int rank = __popc(peers & lt);
        out[first + rank] = (int)(p0 + e * 32 + lane);
        if (rank == __popc(peers) - 1) cur[bin] = first + rank + 1;  // the last peer advances the cursor
      }
      __syncwarp();
    }
  }
}

// ---- backward ------------------------------------------------------------------------------------------------
// grad_x_q[b,c,q] = sum_j (go[b,C+c,q,j] - go[b,c,q,j])       (fixed order over j)
__global__ void __launch_bounds__(256) graph_feature_grad_q_kernel(const float* __restrict__ go, int C, int Nq, int k,
                                                                   float* __restrict__ gq, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % Nq);
    const long long bc = t / Nq;
    const long long b = bc / C;
    const int c = (int)(bc - b * C);
    const float* lo = go + (((size_t)b * 2 * C + c) * Nq + q) * k;
    const float* hi = go + (((size_t)b * 2 * C + C + c) * Nq + q) * k;
    float s = 0.f;
    for (int j = 0; j < k; ++j) s = __fadd_rn(s, __fsub_rn(__ldg(hi + j), __ldg(lo + j)));
    gq[t] = s;
  }
}

// grad_x_k[b,c,n] = sum_{p in inverse list of n, ascending} go[b,c,p].  Owner thread per (b, c, n); the list is
// walked 8 entries at a time so 8 independent gathers of go are in flight, and summed in list order.
__global__ void __launch_bounds__(256) graph_feature_grad_k_kernel(const float* __restrict__ go,
                                                                   const int* __restrict__ start,
                                                                   const int* __restrict__ perm, int C, long long L,
                                                                   int Nk, float* __restrict__ gk) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Nk) return;
  const float* plane = go + ((size_t)b * 2 * C + c) * L;
  const int* st = start + (size_t)b * (Nk + 1);
  const int* pm = perm + (size_t)b * L;
  const int s0 = st[n], s1 = st[n + 1];
  float s = 0.f;
  for (int i0 = s0; i0 < s1; i0 += 8) {
    int pp[8];
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) pp[e] = (i0 + e < s1) ? __ldg(pm + i0 + e) : -1;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (pp[e] >= 0) ? __ldg(plane + pp[e]) : 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (pp[e] >= 0) s = __fadd_rn(s, v[e]);
  }
  gk[((size_t)b * C + c) * Nk + n] = s;
}

// ---- grouped gathers for the remaining pointnet2_ops entry points ---------------------
// reduce == 0 (grouping_operation): out[b,c,p,s] = f[b,c,idx[b,p,s]]            (* w[b,p,s] when w != null)
// reduce == 1 (three_interpolate):  out[b,c,p]   = sum_s w[b,p,s] * f[b,c,idx[b,p,s]]   (fixed order over s)
__global__ void __launch_bounds__(256) group_gather_kernel(const float* __restrict__ f, const long long* __restrict__ idx,
                                                           const float* __restrict__ w, int C, int M, int P, int S,
                                                           int reduce, float* __restrict__ out, long long total) {
  const long long L = (long long)P * S;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    if (reduce) {
      const int pp = (int)(t % P);
      const long long bc = t / P;
      const long long b = bc / C;
      const float* row = f + bc * M;
      float acc = 0.f;
      for (int j = 0; j < S; ++j) {
        const long long pos = b * L + (long long)pp * S + j;
        acc = __fadd_rn(acc, __fmul_rn(__ldg(w + pos), __ldg(row + __ldg(idx + pos))));
      }
      out[t] = acc;
    } else {
      const long long pos = t % L;
      const long long bc = t / L;
      const long long b = bc / C;
      float v = __ldg(f + bc * M + __ldg(idx + b * L + pos));
      if (w) v = __fmul_rn(v, __ldg(w + b * L + pos));
      out[t] = v;
    }
  }
}

// grad_f[b,c,m] = sum over the inverse list of m (ascending position) of w[pos] * go[b,c, reduce ? pos / S : pos]
__global__ void __launch_bounds__(256) group_gather_grad_kernel(const float* __restrict__ go,
                                                                const int* __restrict__ start,
                                                                const int* __restrict__ perm,
                                                                const float* __restrict__ w, int C, long long L, int S,
                                                                int reduce, int M, float* __restrict__ gf) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long plane_len = reduce ? L / S : L;
  const float* plane = go + ((size_t)b * C + c) * plane_len;
  const int* st = start + (size_t)b * (M + 1);
  const int* pm = perm + (size_t)b * L;
  const float* wb = w ? w + (size_t)b * L : nullptr;
  float acc = 0.f;
  for (int i = st[m]; i < st[m + 1]; ++i) {
    const int pos = __ldg(pm + i);
    float v = __ldg(plane + (reduce ? pos / S : pos));
    if (wb) v = __fmul_rn(v, __ldg(wb + pos));
    acc = __fadd_rn(acc, v);
  }
  gf[((size_t)b * C + c) * M + m] = acc;
}

// ball_query (pointnet2_ops ball_query_kernel): the first nsample references, in ascending index, whose squared
// distance to the query is < radius^2; unused slots repeat the first hit; no hit leaves zeros.
__global__ void __launch_bounds__(256) ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ qry,
                                                         int N, int P, float radius2, int nsample,
                                                         int* __restrict__ idx, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / P;
    const float* q = qry + 3 * t;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    const float* X = xyz + (size_t)b * N * 3;
    int* o = idx + t * nsample;
    int cnt = 0;
    for (int k = 0; k < N && cnt < nsample; ++k) {
      const float d2 = sqdist(qx, qy, qz, __ldg(X + 3 * k), __ldg(X + 3 * k + 1), __ldg(X + 3 * k + 2));
      if (d2 < radius2) {
        if (cnt == 0)
          for (int l = 0; l < nsample; ++l) o[l] = k;
        o[cnt] = k;
        ++cnt;
      }
    }
    if (cnt == 0)
      for (int l = 0; l < nsample; ++l) o[l] = 0;
  }
}

static inline int grid1d(long long total, int threads, long long cap = 148LL * 32) {
  long long g = (total + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int graph_feature_launch(const float* xq, const float* xk, const long long* idx, int B, int C, int Nq, int Nk, int k,
                         float* out, cudaStream_t stream) {
  const long long L = (long long)Nq * k;
  const unsigned gy = (C + kGfChannelsPerCta - 1) / kGfChannelsPerCta;
  if (k % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    dim3 grid((unsigned)((L / 4 + kGfThreads - 1) / kGfThreads), gy, B);
    graph_feature_kernel<4><<<grid, kGfThreads, 0, stream>>>(xq, xk, idx, C, Nq, Nk, k, out);
  } else {
    dim3 grid((unsigned)((L + kGfThreads - 1) / kGfThreads), gy, B);
    graph_feature_kernel<1><<<grid, kGfThreads, 0, stream>>>(xq, xk, idx, C, Nq, Nk, k, out);
  }
  count_launch();
  return (int)cudaGetLastError();
}

int index_points_launch(const float* pts, const long long* idx, int B, int N, int C, long long S, float* out,
                        cudaStream_t stream) {
  const long long total = (long long)B * S * C;
  index_points_kernel<<<grid1d(total, 256), 256, 0, stream>>>(pts, idx, N, C, S, out, total);
  count_launch();
  return (int)cudaGetLastError();
}

size_t graph_feature_grad_workspace_bytes(int B, int Nq, int Nk, int k) {
  return sizeof(int) * ((size_t)B * (Nk + 1) + (size_t)B * Nq * k);
}

int graph_feature_grad_launch(const float* go, const long long* idx, int B, int C, int Nq, int Nk, int k, float* gq,
                              float* gk, void* workspace, cudaStream_t stream) {
  const long long L = (long long)Nq * k;
  int* start = static_cast<int*>(workspace);
  int* perm = start + (size_t)B * (Nk + 1);
  cudaError_t err;
  if (gq) {
    const long long total = (long long)B * C * Nq;
    graph_feature_grad_q_kernel<<<grid1d(total, 256), 256, 0, stream>>>(go, C, Nq, k, gq, total);
    count_launch();
  }
  if (gk) {
    if ((err = cudaMemsetAsync(start, 0, sizeof(int) * (size_t)B * (Nk + 1), stream)) != cudaSuccess) return (int)err;
    dim3 gc((unsigned)grid1d(L, 256, 64), B);
    inv_count_kernel<<<gc, 256, 0, stream>>>(idx, L, Nk, start);
    inv_scan_kernel<<<B, 1024, 0, stream>>>(start, Nk);
    dim3 gf((Nk + 8 * kInvBins - 1) / (8 * kInvBins), B);
    inv_fill_kernel<<<gf, 256, 0, stream>>>(idx, L, Nk, start, perm);
    dim3 gg((Nk + 255) / 256, C, B);
    graph_feature_grad_k_kernel<<<gg, 256, 0, stream>>>(go, start, perm, C, L, Nk, gk);
    count_launch(4);
  }
  return (int)cudaGetLastError();
}

int group_gather_launch(const float* f, const long long* idx, const float* w, int B, int C, int M, int P, int S,
                        int reduce, float* out, cudaStream_t stream) {
  const long long total = (long long)B * C * P * (reduce ? 1 : S);
  group_gather_kernel<<<grid1d(total, 256), 256, 0, stream>>>(f, idx, w, C, M, P, S, reduce, out, total);
  count_launch();
  return (int)cudaGetLastError();
}

int group_gather_grad_launch(const float* go, const long long* idx, const float* w, int B, int C, int M, int P, int S,
                             int reduce, float* gf, void* workspace, cudaStream_t stream) {
  const long long L = (long long)P * S;
  int* start = static_cast<int*>(workspace);
  int* perm = start + (size_t)B * (M + 1);
  cudaError_t err;
  if ((err = cudaMemsetAsync(start, 0, sizeof(int) * (size_t)B * (M + 1), stream)) != cudaSuccess) return (int)err;
  dim3 gc((unsigned)grid1d(L, 256, 64), B);
  inv_count_kernel<<<gc, 256, 0, stream>>>(idx, L, M, start);
  inv_scan_kernel<<<B, 1024, 0, stream>>>(start, M);
  dim3 gfill((M + 8 * kInvBins - 1) / (8 * kInvBins), B);
  inv_fill_kernel<<<gfill, 256, 0, stream>>>(idx, L, M, start, perm);
  dim3 gg((M + 255) / 256, C, B);
  group_gather_grad_kernel<<<gg, 256, 0, stream>>>(go, start, perm, w, C, L, S, reduce, M, gf);
  count_launch(4);
  return (int)cudaGetLastError();
}

int ball_query_launch(const float* xyz, const float* qry, int B, int N, int P, float radius, int nsample, int* idx,
                      cudaStream_t stream) {
  const long long total = (long long)B * P;
  ball_query_kernel<<<grid1d(total, 256), 256, 0, stream>>>(xyz, qry, N, P, radius * radius, nsample, idx, total);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
