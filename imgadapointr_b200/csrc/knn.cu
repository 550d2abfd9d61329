// knn.cu - fused distance + top-k selection (kNN grouping) for sm_100a.
//
// Replaces knn_point = square_distance + torch.topk(largest=False, sorted=False)
// (reference models/Transformer_utils.py:19-51), which materialises the [B,Q,R] FP32 matrix (805 MB at
// B=48, Q=R=2048) through a K=3 SGEMM and re-reads it in topk.  Here nothing is materialised: reference
// points are staged once per CTA in shared memory (SoA) and every warp owns one query at a time, keeping the
// current k best as a sorted list spread over its lanes (lane i = i-th smallest).  A 32-reference step costs one
// direct-form distance per lane, a compare against the current k-th distance and one ballot; the (rare)
// candidates are inserted with a ballot-computed position and one shuffle-up.
// Output is ascending by (distance, index); on exact distance ties the LOWEST index wins (references are scanned
// in ascending index and only a strictly smaller distance displaces a list entry).
#include "common.cuh"

namespace iapt {

constexpr int kKnnThreads = 256;
constexpr int kKnnWarps = kKnnThreads / 32;
constexpr int kKnnTile = 2048;         // reference points per smem tile (24 KB SoA)
constexpr int kKnnQueriesPerCta = 32;  // queries per CTA (4 per warp), amortises the tile fill

template <int kSlots>  // list entries per lane: k <= 32 * kSlots
__global__ void __launch_bounds__(kKnnThreads) knn_kernel(const float* __restrict__ ref, const float* __restrict__ qry,
                                                          int R, int Q, int k, int q_blocks,
                                                          int64_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  __shared__ float sx[kKnnTile], sy[kKnnTile], sz[kKnnTile];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x / q_blocks;
  const int q0 = (blockIdx.x - b * q_blocks) * kKnnQueriesPerCta;
  const float* __restrict__ Rb = ref + (size_t)b * R * 3;
  const float* __restrict__ Qb = qry + (size_t)b * Q * 3;
  constexpr int kQPW = kKnnQueriesPerCta / kKnnWarps;
  const float kInf = __int_as_float(0x7f800000);

  float qx[kQPW], qy[kQPW], qz[kQPW];
  float ld[kQPW][kSlots];  // sorted list: entry (s*32 + lane)
  int li[kQPW][kSlots];
  float tau[kQPW];
#pragma unroll
  for (int u = 0; u < kQPW; ++u) {
    const int q = q0 + warp * kQPW + u;
    const int qq = q < Q ? q : 0;
    qx[u] = __ldg(Qb + 3 * (size_t)qq + 0);
    qy[u] = __ldg(Qb + 3 * (size_t)qq + 1);
    qz[u] = __ldg(Qb + 3 * (size_t)qq + 2);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      ld[u][s] = kInf;
      li[u][s] = 0x7fffffff;
    }
    tau[u] = kInf;
  }

  for (int t0 = 0; t0 < R; t0 += kKnnTile) {
    const int tn = min(kKnnTile, R - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * 3; i += kKnnThreads) {
      const float v = __ldg(Rb + 3 * (size_t)t0 + i);
      const int r = i / 3, c = i - 3 * r;
      (c == 0 ? sx : (c == 1 ? sy : sz))[r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kQPW; ++u) {
      if (q0 + warp * kQPW + u >= Q) continue;  // warp-uniform
      for (int r0 = 0; r0 < tn; r0 += 32) {
        const int r = r0 + lane;
        float d = kInf;
        if (r < tn) d = sqdist(qx[u], qy[u], qz[u], sx[r], sy[r], sz[r]);
        unsigned mask = __ballot_sync(0xffffffffu, d < tau[u]);
        while (mask) {
          const int src = __ffs(mask) - 1;
          const float dc = __shfl_sync(0xffffffffu, d, src);
          const int ic = t0 + r0 + src;
          // position = number of list entries <= dc (equal distances already listed have lower indices)
          int pos = 0;
#pragma unroll
          for (int s = 0; s < kSlots; ++s) pos += __popc(__ballot_sync(0xffffffffu, ld[u][s] <= dc));
          // shift entries at >= pos up by one (highest slot first), then drop the candidate in
#pragma unroll
          for (int s = kSlots - 1; s >= 0; --s) {
            float up_d = __shfl_up_sync(0xffffffffu, ld[u][s], 1);
            int up_i = __shfl_up_sync(0xffffffffu, li[u][s], 1);
            if (s > 0) {
              const float carry_d = __shfl_sync(0xffffffffu, ld[u][s - 1], 31);
              const int carry_i = __shfl_sync(0xffffffffu, li[u][s - 1], 31);
              if (lane == 0) {
                up_d = carry_d;
                up_i = carry_i;
              }
            }
            const int e = s * 32 + lane;
            if (e > pos) {
              ld[u][s] = up_d;
              li[u][s] = up_i;
            } else if (e == pos) {
              ld[u][s] = dc;
              li[u][s] = ic;
            }
          }
          float kth = ld[u][0];
          if (kSlots > 1 && k > 32) kth = ld[u][kSlots - 1];
          tau[u] = __shfl_sync(0xffffffffu, kth, (k - 1) & 31);
          mask &= mask - 1;
          mask &= __ballot_sync(0xffffffffu, d < tau[u]);
        }
      }
    }
  }

#pragma unroll
  for (int u = 0; u < kQPW; ++u) {
    const int q = q0 + warp * kQPW + u;
    if (q >= Q) continue;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int e = s * 32 + lane;
      if (e < k) {
        const size_t o = ((size_t)b * Q + q) * k + e;
        idx_out[o] = (int64_t)li[u][s];
        if (dist_out) dist_out[o] = ld[u][s];
      }
    }
  }
}

int knn_launch(const float* ref, const float* qry, int B, int R, int Q, int k, int64_t* idx, float* dist,
               cudaStream_t stream) {
  const int q_blocks = (Q + kKnnQueriesPerCta - 1) / kKnnQueriesPerCta;
  const long long blocks = (long long)B * q_blocks;
  if (k <= 32)
    knn_kernel<1><<<(unsigned)blocks, kKnnThreads, 0, stream>>>(ref, qry, R, Q, k, q_blocks, idx, dist);
  else
    knn_kernel<2><<<(unsigned)blocks, kKnnThreads, 0, stream>>>(ref, qry, R, Q, k, q_blocks, idx, dist);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
