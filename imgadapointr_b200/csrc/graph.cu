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
#include <cstdlib>

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

// Clouds with at most kInvCtaBins bins take ONE launch instead of the three above: a CTA owns the cloud, its warps
// own consecutive slices of the positions.  Pass 1 counts into warp-private shared-memory histograms (order free);
// a block scan turns the bin totals into start[] and each histogram into that warp's write cursors (slices in
// position order, so the sort stays stable); pass 2 re-reads the keys in order and places them - same-bin hits inside
// a 32-key step take consecutive slots in lane order.  Every key hits, so the match/ballot work that the 64-bin
// kernel above repeats in 32 warps per cloud is done once: 142 us -> 9 us at 48 clouds x 32768 positions x 2048 bins.
constexpr int kInvCtaBins = 2048;
constexpr int kInvCtaBits = 11;  // log2(kInvCtaBins)
constexpr int kInvCtaWarps = 16;
constexpr int kInvCtaThreads = 32 * kInvCtaWarps;

__global__ void __launch_bounds__(kInvCtaThreads) inv_index_cta_kernel(const long long* __restrict__ idx, long long L,
                                                                       int Nk, int* __restrict__ start,
                                                                       int* __restrict__ perm) {
  extern __shared__ int inv_hist[];  // [kInvCtaWarps][Nk]
  __shared__ int warp_tot[kInvCtaWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const long long* key = idx + (size_t)b * L;
  int* out = perm + (size_t)b * L;
  int* st = start + (size_t)b * (Nk + 1);
  int* mine = inv_hist + warp * Nk;
  for (int i = threadIdx.x; i < kInvCtaWarps * Nk; i += kInvCtaThreads) inv_hist[i] = 0;
  __syncthreads();
  const long long chunk = ((L + kInvCtaWarps - 1) / kInvCtaWarps + 31) / 32 * 32;
  const long long p_lo = min(L, warp * chunk), p_hi = min(L, p_lo + chunk);
  const unsigned lt = (1u << lane) - 1u;
  constexpr int kAhead = 16;  // keys in flight per lane: both passes are bound by the L2 round trip of the key loads
  for (long long p0 = p_lo; p0 < p_hi; p0 += 32 * kAhead) {
    int kk[kAhead];
#pragma unroll
    for (int e = 0; e < kAhead; ++e) {
      const long long p = p0 + e * 32 + lane;
      kk[e] = p < p_hi ? (int)__ldg(key + p) : -1;
    }
#pragma unroll
    for (int e = 0; e < kAhead; ++e)
      if (kk[e] >= 0) atomicAdd(mine + kk[e], 1);  // (ballot-aggregated plain adds measured slower: 64 vs 44 us)
  }
  __syncthreads();

  // block exclusive scan over the bins; thread t owns bins [t*per, (t+1)*per)
  const int per = (Nk + kInvCtaThreads - 1) / kInvCtaThreads;
  const int j0 = min(Nk, (int)threadIdx.x * per), j1 = min(Nk, j0 + per);
  int local = 0;
  for (int j = j0; j < j1; ++j)
    for (int w = 0; w < kInvCtaWarps; ++w) local += inv_hist[w * Nk + j];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int base = incl - local;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  for (int j = j0; j < j1; ++j) {
    st[j] = base;
    for (int w = 0; w < kInvCtaWarps; ++w) {  // histogram -> this warp's first slot for the bin
      const int t = inv_hist[w * Nk + j];
      inv_hist[w * Nk + j] = base;
      base += t;
    }
  }
  if (threadIdx.x == 0) st[Nk] = (int)L;
  __syncthreads();

  for (long long p0 = p_lo; p0 < p_hi; p0 += 32 * kAhead) {
    int kk[kAhead];
#pragma unroll
    for (int e = 0; e < kAhead; ++e) {
      const long long p = p0 + e * 32 + lane;
      kk[e] = p < p_hi ? (int)__ldg(key + p) : -1;
    }
#pragma unroll
    for (int e = 0; e < kAhead; ++e) {
      const bool valid = kk[e] >= 0;
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      if (vmask == 0u) break;  // warp-uniform: the slice ended
      const int bin = valid ? kk[e] : 0;
      const int first = valid ? mine[bin] : 0;
      __syncwarp();  // every peer has read the cursor before one of them advances it
      // lanes holding the same bin: AND of one ballot per key bit (independent VOTEs; MATCH.ANY measured ~100 cycles
      // per warp instruction and serialised the whole pass)
      unsigned peers = vmask;
#pragma unroll
      for (int bit = 0; bit < kInvCtaBits; ++bit) {
        const unsigned set = __ballot_sync(0xffffffffu, (bin >> bit) & 1);
        peers &= ((bin >> bit) & 1) ? set : ~set;
      }
      if (valid) {
        const int rank = __popc(peers & lt);
        out[first + rank] = (int)(p0 + e * 32 + lane);
        if (rank == __popc(peers) - 1) mine[bin] = first + rank + 1;  // the last peer advances the cursor
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

// ---- both gradients of one (cloud, channel) plane from ONE pass over grad_out's lower half ------------------------
// When the plane go[b,c,:,:] fits in shared memory (L <= 49152 floats: every DGCNN grouper shape), thread 0 issues
// cp.async.bulk (TMA) copies of it against an mbarrier and the CTA spends the flight time fetching the heads of its
// inverse lists.  After the wait grad_x_k's owner threads gather from shared memory - instead of one 32-byte sector
// read per 4-byte element, the kernel above - and grad_x_q subtracts the staged rows from the upper half streamed
// from global memory.  Same summation orders as the two kernels above, so the results are bit-identical to them.
// (A register-staged LDG -> STS version with a chunk loop measured 2x slower than this: every dependent round trip
// to L2 costs ~1 us here.)
constexpr int kPlaneBins = 2;          // bins per thread: Nk <= 2 * threads, threads in {256, 512, 1024}
constexpr int kPlaneHead = 12;         // inverse-list entries fetched during the copy (mean list length = k)
constexpr int kPlaneMaxFloats = 49152; // 192 KB

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int kPlaneThreads>
__global__ void __launch_bounds__(kPlaneThreads, 1024 / kPlaneThreads) graph_feature_grad_plane_kernel(
    const float* __restrict__ go, const int* __restrict__ start, const int* __restrict__ perm, int C, int Nq, int Nk,
    int k, float* __restrict__ gq, float* __restrict__ gk) {
  extern __shared__ __align__(128) float plane[];
  __shared__ __align__(8) unsigned long long bar;
  const int c = blockIdx.x, b = blockIdx.y;
  const int L = Nq * k;
  const float* lo = go + ((size_t)b * 2 * C + c) * L;
  const float* hi = go + ((size_t)b * 2 * C + C + c) * L;
  const int* st = start + (size_t)b * (Nk + 1);
  const int* pm = perm + (size_t)b * L;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)L * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    for (unsigned off = 0; off < bytes; off += 32768u) {
      const unsigned n = min(32768u, bytes - off);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(reinterpret_cast<const char*>(plane) + off)),
          "l"(reinterpret_cast<const char*>(lo) + off), "r"(n), "r"(smem_u32(&bar))
          : "memory");
    }
  }

  // while the plane travels: the heads of this thread's inverse lists
  int s0[kPlaneBins], s1[kPlaneBins], head[kPlaneBins][kPlaneHead];
#pragma unroll
  for (int u = 0; u < kPlaneBins; ++u) {
    const int n = u * kPlaneThreads + threadIdx.x;
    const bool on = gk != nullptr && n < Nk;
    s0[u] = on ? __ldg(st + n) : 0;
    s1[u] = on ? __ldg(st + n + 1) : 0;
  }
#pragma unroll
  for (int u = 0; u < kPlaneBins; ++u)
#pragma unroll
    for (int e = 0; e < kPlaneHead; ++e) head[u][e] = (s0[u] + e < s1[u]) ? __ldg(pm + s0[u] + e) : -1;

  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(&bar))
        : "memory");
  }

  if (gk != nullptr) {
#pragma unroll
    for (int u = 0; u < kPlaneBins; ++u) {
      const int n = u * kPlaneThreads + threadIdx.x;
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < kPlaneHead; ++e)
        if (head[u][e] >= 0) acc = __fadd_rn(acc, plane[head[u][e]]);
      for (int i0 = s0[u] + kPlaneHead; i0 < s1[u]; i0 += 8) {  // the tail of a longer list, 8 entries per round
        int pp[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) pp[e] = (i0 + e < s1[u]) ? __ldg(pm + i0 + e) : -1;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (pp[e] >= 0) acc = __fadd_rn(acc, plane[pp[e]]);
      }
      if (n < Nk) gk[((size_t)b * C + c) * Nk + n] = acc;
    }
  }
  if (gq != nullptr) {
    for (int r0 = 0; r0 < Nq; r0 += 2 * kPlaneThreads) {
      if (k == 16) {  // the grouper's k: both rows of this thread requested before either is consumed
        float4 hv[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = r0 + u * kPlaneThreads + threadIdx.x;
          if (r < Nq) {
#pragma unroll
            for (int j = 0; j < 4; ++j) hv[u][j] = __ldg(reinterpret_cast<const float4*>(hi + (size_t)r * 16) + j);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = r0 + u * kPlaneThreads + threadIdx.x;
          if (r < Nq) {
            const float4* l = reinterpret_cast<const float4*>(plane + (size_t)r * 16);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 lv = l[j];
              s = __fadd_rn(s, __fsub_rn(hv[u][j].x, lv.x));
              s = __fadd_rn(s, __fsub_rn(hv[u][j].y, lv.y));
              s = __fadd_rn(s, __fsub_rn(hv[u][j].z, lv.z));
              s = __fadd_rn(s, __fsub_rn(hv[u][j].w, lv.w));
            }
            gq[((size_t)b * C + c) * Nq + r] = s;
          }
        }
      } else {
        for (int u = 0; u < 2; ++u) {
          const int r = r0 + u * kPlaneThreads + threadIdx.x;
          if (r < Nq) {
            const float* h = hi + (size_t)r * k;
            const float* l = plane + (size_t)r * k;
            float s = 0.f;
            for (int j = 0; j < k; j += 4) {
              const float4 hv = __ldg(reinterpret_cast<const float4*>(h + j));
              const float4 lv = *reinterpret_cast<const float4*>(l + j);
              s = __fadd_rn(s, __fsub_rn(hv.x, lv.x));
              s = __fadd_rn(s, __fsub_rn(hv.y, lv.y));
              s = __fadd_rn(s, __fsub_rn(hv.z, lv.z));
              s = __fadd_rn(s, __fsub_rn(hv.w, lv.w));
            }
            gq[((size_t)b * C + c) * Nq + r] = s;
          }
        }
      }
    }
  }
}

// ---- channels-last graph feature (DynamicGraphAttention, models/Transformer_utils.py:814-816,842-855) ------------
// The geometry-aware transformer gathers token features [B,M,C] (C = 384) with the kNN indices and builds
//   feature[b,n,j,:] = cat(v[b,idx[b,n,j],:] - q[b,n,:], q[b,n,:])            ([B,N,k,2C])
// through index_points + expand + sub + cat: four passes over a 600 MB tensor at B=48, N=512, k=8.  Here one warp per
// (b,n) writes its k rows directly (128-bit coalesced), and the backward is deterministic:
//   grad_q[b,n,:] = sum_j (go[b,n,j,C:] - go[b,n,j,:C])           fixed order over j
//   grad_v[b,m,:] = sum over the inverse list of m (ascending position) of go[b,p,:C]     - whole rows, coalesced
// (the reference's index backward is an atomicAdd scatter in run-dependent order).
constexpr int kClThreads = 256;

__global__ void __launch_bounds__(kClThreads) graph_feature_cl_kernel(const float* __restrict__ q,
                                                                      const float* __restrict__ v,
                                                                      const long long* __restrict__ idx, int N, int M,
                                                                      int C, int k, float* __restrict__ out,
                                                                      long long rows, int vec) {
  const int lane = threadIdx.x & 31;
  const long long bn = ((long long)blockIdx.x * kClThreads + threadIdx.x) >> 5;  // warp = one (b, n)
  if (bn >= rows) return;
  const long long b = bn / N;
  const float* qr = q + (size_t)bn * C;
  const long long* ir = idx + (size_t)bn * k;
  float* o = out + (size_t)bn * k * 2 * C;
  if (vec) {
    const int C4 = C >> 2;
    for (int c0 = 0; c0 < C4; c0 += 32 * 4) {  // up to 4 float4 columns per lane stay in registers across the k rows
      float4 qv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        qv[u] = c < C4 ? __ldg(reinterpret_cast<const float4*>(qr) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int j = 0; j < k; ++j) {
        const float4* vr = reinterpret_cast<const float4*>(v + ((size_t)b * M + (size_t)__ldg(ir + j)) * C);
        float4* lo = reinterpret_cast<float4*>(o + (size_t)j * 2 * C);
        float4* hi = lo + C4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          if (c < C4) {
            const float4 x = __ldg(vr + c);
            lo[c] = make_float4(__fsub_rn(x.x, qv[u].x), __fsub_rn(x.y, qv[u].y), __fsub_rn(x.z, qv[u].z),
                                __fsub_rn(x.w, qv[u].w));
            hi[c] = qv[u];
          }
        }
      }
    }
  } else {
    for (int j = 0; j < k; ++j) {
      const float* vr = v + ((size_t)b * M + (size_t)__ldg(ir + j)) * C;
      float* lo = o + (size_t)j * 2 * C;
      for (int c = lane; c < C; c += 32) {
        const float x = __ldg(qr + c);
        lo[c] = __fsub_rn(__ldg(vr + c), x);
        lo[C + c] = x;
      }
    }
  }
}

// grad_q: warp per (b, n)
__global__ void __launch_bounds__(kClThreads) graph_feature_cl_grad_q_kernel(const float* __restrict__ go, int C, int k,
                                                                             float* __restrict__ gq, long long rows,
                                                                             int vec) {
  const int lane = threadIdx.x & 31;
  const long long bn = ((long long)blockIdx.x * kClThreads + threadIdx.x) >> 5;
  if (bn >= rows) return;
  const float* g = go + (size_t)bn * k * 2 * C;
  float* o = gq + (size_t)bn * C;
  if (vec) {
    const int C4 = C >> 2;
    for (int c = lane; c < C4; c += 32) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < k; ++j) {
        const float4* row = reinterpret_cast<const float4*>(g + (size_t)j * 2 * C);
        const float4 l = __ldg(row + c), h = __ldg(row + C4 + c);
        s.x = __fadd_rn(s.x, __fsub_rn(h.x, l.x));
        s.y = __fadd_rn(s.y, __fsub_rn(h.y, l.y));
        s.z = __fadd_rn(s.z, __fsub_rn(h.z, l.z));
        s.w = __fadd_rn(s.w, __fsub_rn(h.w, l.w));
      }
      reinterpret_cast<float4*>(o)[c] = s;
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float s = 0.f;
      for (int j = 0; j < k; ++j) s = __fadd_rn(s, __fsub_rn(__ldg(g + (size_t)j * 2 * C + C + c), __ldg(g + (size_t)j * 2 * C + c)));
      o[c] = s;
    }
  }
}

// out[b,m,0:C] = sum over the inverse list of m (ascending position p) of src[b,p,0:C], rows of src are `ld` floats
// apart (ld = 2C: lower half of the graph feature gradient; ld = C: index_points gradient).  Warp per (b, m).
__global__ void __launch_bounds__(kClThreads) rows_gather_sum_kernel(const float* __restrict__ src,
                                                                     const int* __restrict__ start,
                                                                     const int* __restrict__ perm, int M, int C,
                                                                     long long L, long long ld, float* __restrict__ out,
                                                                     long long rows, int vec) {
  const int lane = threadIdx.x & 31;
  const long long bm = ((long long)blockIdx.x * kClThreads + threadIdx.x) >> 5;
  if (bm >= rows) return;
  const long long b = bm / M;
  const int m = (int)(bm - b * M);
  const int* st = start + (size_t)b * (M + 1);
  const int s0 = __ldg(st + m), s1 = __ldg(st + m + 1);
  const int* pm = perm + (size_t)b * L;
  const float* base = src + (size_t)b * L * ld;
  float* o = out + (size_t)bm * C;
  if (vec) {
    const int C4 = C >> 2;
    for (int c0 = 0; c0 < C4; c0 += 32 * 4) {
      float4 acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i0 = s0; i0 < s1; i0 += 32) {
        const int mine = (i0 + lane < s1) ? __ldg(pm + i0 + lane) : 0;  // 32 list entries per coalesced read
        const int cnt = min(32, s1 - i0);
        for (int e = 0; e < cnt; ++e) {
          const int p = __shfl_sync(0xffffffffu, mine, e);
          const float4* row = reinterpret_cast<const float4*>(base + (size_t)p * ld);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * 32 + lane;
            if (c < C4) {
              const float4 x = __ldg(row + c);
              acc[u].x = __fadd_rn(acc[u].x, x.x);
              acc[u].y = __fadd_rn(acc[u].y, x.y);
              acc[u].z = __fadd_rn(acc[u].z, x.z);
              acc[u].w = __fadd_rn(acc[u].w, x.w);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        if (c < C4) reinterpret_cast<float4*>(o)[c] = acc[u];
      }
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      for (int i = s0; i < s1; ++i) acc = __fadd_rn(acc, __ldg(base + (size_t)__ldg(pm + i) * ld + c));
      o[c] = acc;
    }
  }
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

// start[b, 0..Nk] and perm[b, 0..L): one launch when a CTA can hold the cloud's histograms, else count / scan / fill.
static int inverse_index_launch(const long long* idx, int B, long long L, int Nk, int* start, int* perm,
                                cudaStream_t stream) {
  static const bool force3 = [] { const char* e = std::getenv("IAPT_INV_INDEX_3PASS"); return e && e[0] == '1'; }();
  if (Nk <= kInvCtaBins && !force3) {
    const size_t smem = sizeof(int) * (size_t)kInvCtaWarps * Nk;
    static PerDeviceOnce once;
    if (once.pending()) {
      cudaFuncSetAttribute(inv_index_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(int) * kInvCtaWarps * kInvCtaBins));
      once.done();
    }
    inv_index_cta_kernel<<<B, kInvCtaThreads, smem, stream>>>(idx, L, Nk, start, perm);
    count_launch();
    return (int)cudaGetLastError();
  }
  cudaError_t err;
  if ((err = cudaMemsetAsync(start, 0, sizeof(int) * (size_t)B * (Nk + 1), stream)) != cudaSuccess) return (int)err;
  dim3 gc((unsigned)grid1d(L, 256, 64), B);
  inv_count_kernel<<<gc, 256, 0, stream>>>(idx, L, Nk, start);
  inv_scan_kernel<<<B, 1024, 0, stream>>>(start, Nk);
  dim3 gf((Nk + 8 * kInvBins - 1) / (8 * kInvBins), B);
  inv_fill_kernel<<<gf, 256, 0, stream>>>(idx, L, Nk, start, perm);
  count_launch(3);
  return (int)cudaGetLastError();
}

int graph_feature_grad_launch(const float* go, const long long* idx, int B, int C, int Nq, int Nk, int k, float* gq,
                              float* gk, void* workspace, cudaStream_t stream) {
  const long long L = (long long)Nq * k;
  int* start = static_cast<int*>(workspace);
  int* perm = start + (size_t)B * (Nk + 1);
  if (gk) {
    const int rc = inverse_index_launch(idx, B, L, Nk, start, perm, stream);
    if (rc != 0) return rc;
  }
  static const bool no_plane = [] { const char* e = std::getenv("IAPT_GRAPH_GRAD_PLANE"); return e && e[0] == '0'; }();
  if (!no_plane && Nk <= 1024 * kPlaneBins && L <= kPlaneMaxFloats && k % 4 == 0 &&
      reinterpret_cast<uintptr_t>(go) % 16 == 0) {
    static PerDeviceOnce once;
    if (once.pending()) {
      const int max_smem = (int)(sizeof(float) * kPlaneMaxFloats);
      cudaFuncSetAttribute(graph_feature_grad_plane_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      cudaFuncSetAttribute(graph_feature_grad_plane_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      cudaFuncSetAttribute(graph_feature_grad_plane_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      once.done();
    }
    // smallest CTA that still gives every thread at most 2 bins and 2 query rows: small planes then run several
    // CTAs per SM, which is what hides their copy and list latencies
    const int need = max((Nk + kPlaneBins - 1) / kPlaneBins, (Nq + 1) / 2);
    const dim3 grid(C, B);
    const size_t smem = sizeof(float) * (size_t)L;
    if (need <= 256)
      graph_feature_grad_plane_kernel<256><<<grid, 256, smem, stream>>>(go, start, perm, C, Nq, Nk, k, gq, gk);
    else if (need <= 512)
      graph_feature_grad_plane_kernel<512><<<grid, 512, smem, stream>>>(go, start, perm, C, Nq, Nk, k, gq, gk);
    else
      graph_feature_grad_plane_kernel<1024><<<grid, 1024, smem, stream>>>(go, start, perm, C, Nq, Nk, k, gq, gk);
    count_launch();
    return (int)cudaGetLastError();
  }
  if (gq) {
    const long long total = (long long)B * C * Nq;
    graph_feature_grad_q_kernel<<<grid1d(total, 256), 256, 0, stream>>>(go, C, Nq, k, gq, total);
    count_launch();
  }
  if (gk) {
    dim3 gg((Nk + 255) / 256, C, B);
    graph_feature_grad_k_kernel<<<gg, 256, 0, stream>>>(go, start, perm, C, L, Nk, gk);
    count_launch();
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
  const int rc = inverse_index_launch(idx, B, L, M, start, perm, stream);
  if (rc != 0) return rc;
  dim3 gg((M + 255) / 256, C, B);
  group_gather_grad_kernel<<<gg, 256, 0, stream>>>(go, start, perm, w, C, L, S, reduce, M, gf);
  count_launch();
  return (int)cudaGetLastError();
}

static inline bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

int graph_feature_cl_launch(const float* q, const float* v, const long long* idx, int B, int N, int M, int C, int k,
                            float* out, cudaStream_t stream) {
  const long long rows = (long long)B * N;
  const int vec = (C % 4 == 0 && aligned16(q) && aligned16(v) && aligned16(out)) ? 1 : 0;
  const long long blocks = (rows * 32 + kClThreads - 1) / kClThreads;
  graph_feature_cl_kernel<<<(unsigned)blocks, kClThreads, 0, stream>>>(q, v, idx, N, M, C, k, out, rows, vec);
  count_launch();
  return (int)cudaGetLastError();
}

int graph_feature_cl_grad_launch(const float* go, const long long* idx, int B, int N, int M, int C, int k, float* gq,
                                 float* gv, void* workspace, cudaStream_t stream) {
  const long long L = (long long)N * k;
  if (gq) {
    const long long rows = (long long)B * N;
    const int vec = (C % 4 == 0 && aligned16(go) && aligned16(gq)) ? 1 : 0;
    graph_feature_cl_grad_q_kernel<<<(unsigned)((rows * 32 + kClThreads - 1) / kClThreads), kClThreads, 0, stream>>>(
        go, C, k, gq, rows, vec);
    count_launch();
  }
  if (gv) {
    int* start = static_cast<int*>(workspace);
    int* perm = start + (size_t)B * (M + 1);
    const int rc = inverse_index_launch(idx, B, L, M, start, perm, stream);
    if (rc != 0) return rc;
    const long long rows = (long long)B * M;
    const int vec = (C % 4 == 0 && aligned16(go) && aligned16(gv)) ? 1 : 0;
    rows_gather_sum_kernel<<<(unsigned)((rows * 32 + kClThreads - 1) / kClThreads), kClThreads, 0, stream>>>(
        go, start, perm, M, C, L, 2LL * C, gv, rows, vec);
    count_launch();
  }
  return (int)cudaGetLastError();
}

int index_points_grad_launch(const float* go, const long long* idx, int B, int N, int C, long long S, float* gp,
                             void* workspace, cudaStream_t stream) {
  int* start = static_cast<int*>(workspace);
  int* perm = start + (size_t)B * (N + 1);
  const int rc = inverse_index_launch(idx, B, S, N, start, perm, stream);
  if (rc != 0) return rc;
  const long long rows = (long long)B * N;
  const int vec = (C % 4 == 0 && aligned16(go) && aligned16(gp)) ? 1 : 0;
  rows_gather_sum_kernel<<<(unsigned)((rows * 32 + kClThreads - 1) / kClThreads), kClThreads, 0, stream>>>(
      go, start, perm, N, C, S, (long long)C, gp, rows, vec);
  count_launch();
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
