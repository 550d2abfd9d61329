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
#include <cstdlib>

#include "common.cuh"

namespace iapt {

constexpr int kKnnThreads = 256;
constexpr int kKnnWarps = kKnnThreads / 32;
constexpr int kKnnTile = 2048;         // reference points per smem tile (24 KB SoA)
constexpr int kKnnQueriesPerCta = 32;  // queries per CTA (4 per warp), amortises the tile fill

// Insert this step's candidates (one per lane: distance d, reference index my_idx; d = +inf for "none") into the
// warp-distributed sorted list (entry s*32+lane).  Candidates are taken in lane order and only a strictly smaller
// distance displaces an entry, so with ascending indices across lanes/steps the lowest index wins exact ties.
template <int kSlots>
__device__ __forceinline__ void knn_insert(float d, int my_idx, int k, float (&ld)[kSlots], int (&li)[kSlots],
                                           float& tau) {
  const int lane = threadIdx.x & 31;
  unsigned mask = __ballot_sync(0xffffffffu, d < tau);
  while (mask) {
    const int src = __ffs(mask) - 1;
    const float dc = __shfl_sync(0xffffffffu, d, src);
    const int ic = __shfl_sync(0xffffffffu, my_idx, src);
    // position = number of list entries <= dc (equal distances already listed have lower indices)
    int pos = 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) pos += __popc(__ballot_sync(0xffffffffu, ld[s] <= dc));
    // shift entries at >= pos up by one (highest slot first), then drop the candidate in
#pragma unroll
    for (int s = kSlots - 1; s >= 0; --s) {
      float up_d = __shfl_up_sync(0xffffffffu, ld[s], 1);
      int up_i = __shfl_up_sync(0xffffffffu, li[s], 1);
      if (s > 0) {
        const float carry_d = __shfl_sync(0xffffffffu, ld[s - 1], 31);
        const int carry_i = __shfl_sync(0xffffffffu, li[s - 1], 31);
        if (lane == 0) {
          up_d = carry_d;
          up_i = carry_i;
        }
      }
      const int e = s * 32 + lane;
      if (e > pos) {
        ld[s] = up_d;
        li[s] = up_i;
      } else if (e == pos) {
        ld[s] = dc;
        li[s] = ic;
      }
    }
    float kth = ld[0];
    if (kSlots > 1 && k > 32) kth = ld[kSlots - 1];
    tau = __shfl_sync(0xffffffffu, kth, (k - 1) & 31);
    mask &= mask - 1;
    mask &= __ballot_sync(0xffffffffu, d < tau);
  }
}

// Sort one (distance, index) pair per lane ascending across the warp (bitonic network, 15 exchange stages): the list
// of an empty query is seeded with its first 32 references in ~120 instructions instead of 32 serial insertions.
__device__ __forceinline__ void warp_sort_pairs(float& d, int& i) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, stride);
      const int oi = __shfl_xor_sync(0xffffffffu, i, stride);
      const bool up = (lane & size) == 0;        // ascending block
      const bool lower = (lane & stride) == 0;   // this lane is the lower position of its pair
      const bool other_less = (od < d) || (od == d && oi < i);  // indices are distinct: never equal
      if ((up == lower) ? other_less : !other_less) {
        d = od;
        i = oi;
      }
    }
  }
}

// r_splits > 1: blockIdx.y selects a slice of the reference range and the per-slice lists go to part_d / part_i
// ([B*Q, r_splits, k]) for knn_merge_kernel - used when there are few queries and many references (the loss kNN:
// Q = 64, R = 16384, k = 32, models/external/AdaPoinTr.py:201), where one warp per query cannot fill the machine.
template <int kSlots, int kQPW>  // list entries per lane: k <= 32 * kSlots; kQPW queries per warp (processed in turn)
__global__ void __launch_bounds__(kKnnThreads) knn_kernel(const float* __restrict__ ref, const float* __restrict__ qry,
                                                          int R, int Q, int k, int q_blocks,
                                                          int64_t* __restrict__ idx_out, float* __restrict__ dist_out,
                                                          const int* __restrict__ only_flagged, int r_splits,
                                                          float* __restrict__ part_d, int* __restrict__ part_i) {
  __shared__ float4 sp[kKnnTile];  // one broadcast-free LDS.128 per reference
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x / q_blocks;
  const int q0 = (blockIdx.x - b * q_blocks) * (kKnnWarps * kQPW);
  if (only_flagged != nullptr) {  // second pass of the select path: only queries whose candidate buffer overflowed
    int any = 0;
    if (threadIdx.x < (kKnnWarps * kQPW) && q0 + (int)threadIdx.x < Q) any = only_flagged[(size_t)b * Q + q0 + threadIdx.x];
    if (!__syncthreads_or(any)) return;
  }
  const int split = blockIdx.y;
  const int per = (R + r_splits - 1) / r_splits;
  const int r_lo = split * per, r_hi = min(R, r_lo + per);
  const float* __restrict__ Rb = ref + (size_t)b * R * 3;
  const float* __restrict__ Qb = qry + (size_t)b * Q * 3;
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

  for (int t0 = r_lo; t0 < r_hi; t0 += kKnnTile) {
    const int tn = min(kKnnTile, r_hi - t0);
    __syncthreads();
    for (int r = threadIdx.x; r < tn; r += kKnnThreads) {
      const float* p = Rb + 3 * (size_t)(t0 + r);
      sp[r] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kQPW; ++u) {
      if (q0 + warp * kQPW + u >= Q) continue;  // warp-uniform
      if (only_flagged != nullptr && only_flagged[(size_t)b * Q + q0 + warp * kQPW + u] == 0) continue;
      int r_first = 0;
      if (t0 == r_lo) {  // empty list: seed it with the first 32 references by one warp sort
        float d = kInf;
        int i = t0 + lane;
        if (lane < tn) {
          const float4 p = sp[lane];
          d = sqdist(qx[u], qy[u], qz[u], p.x, p.y, p.z);
        }
        warp_sort_pairs(d, i);
        ld[u][0] = d;
        li[u][0] = (d < kInf) ? i : 0x7fffffff;
        float kth = ld[u][0];
        if (kSlots > 1 && k > 32) kth = ld[u][kSlots - 1];
        tau[u] = __shfl_sync(0xffffffffu, kth, (k - 1) & 31);
        r_first = 32;
      }
      for (int r0 = r_first; r0 < tn; r0 += 32) {
        const int r = r0 + lane;
        float d = kInf;
        if (r < tn) {
          const float4 p = sp[r];
          d = sqdist(qx[u], qy[u], qz[u], p.x, p.y, p.z);
        }
        knn_insert<kSlots>(d, t0 + r, k, ld[u], li[u], tau[u]);
      }
    }
  }

#pragma unroll
  for (int u = 0; u < kQPW; ++u) {
    const int q = q0 + warp * kQPW + u;
    if (q >= Q) continue;
    if (only_flagged != nullptr && only_flagged[(size_t)b * Q + q] == 0) continue;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int e = s * 32 + lane;
      if (e < k) {
        if (r_splits > 1) {
          const size_t o = (((size_t)b * Q + q) * r_splits + split) * k + e;
          part_d[o] = ld[u][s];
          part_i[o] = li[u][s];
        } else {
          const size_t o = ((size_t)b * Q + q) * k + e;
          // an unfilled slot (only possible when distances were NaN - outside the contract) must stay a valid index
          idx_out[o] = (int64_t)(li[u][s] == 0x7fffffff ? 0 : li[u][s]);
          if (dist_out) dist_out[o] = ld[u][s];
        }
      }
    }
  }
}

// Merge the r_splits per-slice lists of each query (each ascending by (distance, index), slices in ascending index
// order) into the final k: the same insertion scheme over r_splits*k candidates, one warp per query.
template <int kSlots>
__global__ void __launch_bounds__(kKnnThreads) knn_merge_kernel(const float* __restrict__ part_d,
                                                                const int* __restrict__ part_i, long long n_queries,
                                                                int k, int r_splits, int64_t* __restrict__ idx_out,
                                                                float* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * kKnnWarps + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const float kInf = __int_as_float(0x7f800000);
  float ld[kSlots];
  int li[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    ld[s] = kInf;
    li[s] = 0x7fffffff;
  }
  float tau = kInf;
  const int total = r_splits * k;
  const float* pd = part_d + (size_t)q * total;
  const int* pi = part_i + (size_t)q * total;
  for (int c0 = 0; c0 < total; c0 += 32) {
    const int c = c0 + lane;
    float d = kInf;
    int i = 0x7fffffff;
    if (c < total) {
      i = __ldg(pi + c);
      if (i != 0x7fffffff) d = __ldg(pd + c);
    }
    knn_insert<kSlots>(d, i, k, ld, li, tau);
  }
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int e = s * 32 + lane;
    if (e < k) {
      idx_out[(size_t)q * k + e] = (int64_t)(li[s] == 0x7fffffff ? 0 : li[s]);
      if (dist_out) dist_out[(size_t)q * k + e] = ld[s];
    }
  }
}

// ---- k <= 16: thread-per-query threshold selection ------------------------------------------------------------
// The warp-per-query kernel above spends most of its issue slots on list insertions (ncu: issue-bound, 81 %).
// For the encoder's k = 8 / 16 calls a cheaper exact scheme: every THREAD owns one query and sweeps the references
// (broadcast LDS.128 from a shared-memory tile) twice:
//   sweep 1: 32 running minima, one per residue class of the reference index (mod 32) - 1 FMNMX per pair.  The k-th
//            smallest of those 32 minima, tau0, is an upper bound on the k-th nearest distance (they belong to 32
//            distinct references) and a tight one: ~22 references lie below it for k = 16, R = 2048.
//   sweep 2: references with d <= tau0 are appended in ascending index to a per-thread candidate list (u16 indices in
//            shared memory) - a bit mask per 32 references, expanded once per chunk, keeps the inner loop branch-free;
//   select : a stable insertion sort by distance of the <= kSelCap candidates gives the k nearest ascending by
//            (distance, index) - the same lowest-index tie rule.
// A query whose candidate list overflows (adversarial clustering) is flagged and redone by knn_kernel<1>.
constexpr int kSelThreads = 256;
constexpr int kSelTile = 2048;  // references per smem tile (float4: 32 KB)
constexpr int kSelCap = 64;     // candidates per query
constexpr int kSelMaxK = 16;

__device__ __forceinline__ void cmpswap(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}

// in-register bitonic sort of 32 floats (ascending); indices are compile-time after unrolling
__device__ __forceinline__ void sort32(float (&v)[32]) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = i ^ stride;
        if (j > i) {
          if ((i & size) == 0)
            cmpswap(v[i], v[j]);
          else
            cmpswap(v[j], v[i]);
        }
      }
    }
  }
}

struct __align__(16) SelSmem {
  union {
    float4 ref[kSelTile];  // sweeps: the reference tile (32 KB)
    struct {               // select phase (after a block barrier): per-thread sorted top-k
      float topd[kSelMaxK][kSelThreads];
      int topi[kSelMaxK][kSelThreads];
    } top;
  };
  unsigned short cand[kSelCap][kSelThreads];  // candidate reference indices (R <= 65536 on this path)
};

__global__ void __launch_bounds__(kSelThreads) knn_select_kernel(const float* __restrict__ ref,
                                                                 const float* __restrict__ qry, int R, int Q, int k,
                                                                 int q_blocks, int64_t* __restrict__ idx_out,
                                                                 float* __restrict__ dist_out, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char sel_raw[];
  SelSmem& sm = *reinterpret_cast<SelSmem*>(sel_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.x / q_blocks;
  const int q = (blockIdx.x - b * q_blocks) * kSelThreads + tid;
  const bool live = q < Q;
  const float* __restrict__ Rb = ref + (size_t)b * R * 3;
  const float kInf = __int_as_float(0x7f800000);
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) {
    const float* qp = qry + ((size_t)b * Q + q) * 3;
    qx = __ldg(qp);
    qy = __ldg(qp + 1);
    qz = __ldg(qp + 2);
  }

  auto stage = [&](int t0, int tn) {
    __syncthreads();
    for (int r = tid; r < kSelTile; r += kSelThreads) {
      float4 v = make_float4(kPadCoord(), kPadCoord(), kPadCoord(), 0.f);
      if (r < tn) {
        const float* p = Rb + 3 * (size_t)(t0 + r);
        v = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
      }
      sm.ref[r] = v;
    }
    __syncthreads();
  };

  // ---- sweep 1: residue-class minima ---------------------------------------------------------------------------
  float mn[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) mn[i] = kInf;
  for (int t0 = 0; t0 < R; t0 += kSelTile) {
    const int tn = min(kSelTile, R - t0);
    stage(t0, tn);
    const int steps = (tn + 31) / 32;  // padded entries are 1e18 sentinels: distances ~3e36, never below a real one
    for (int s = 0; s < steps; ++s) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float4 p = sm.ref[s * 32 + i];
        mn[i] = fminf(mn[i], sqdist(qx, qy, qz, p.x, p.y, p.z));
      }
    }
  }
  sort32(mn);
  float tau = mn[0];
#pragma unroll
  for (int i = 1; i < kSelMaxK; ++i)
    if (i < k) tau = mn[i];  // k-th smallest class minimum (k <= 16)

  // ---- sweep 2: collect candidates -----------------------------------------------------------------------------
  // Branch-free inner loop: 32 references set bits of a per-thread mask (FSETP + SEL), so the loads and the FP chains
  // of neighbouring references overlap; the set bits are appended to the candidate list once per 32-reference chunk.
  // Of the references exactly AT the threshold only the first k (ascending index) are kept - zero-padded clouds hold
  // hundreds of them; a query with more than kSelCap candidates goes to the list kernel.
  int count = 0, eq_taken = 0;
  float thr = tau;  // inclusive threshold of the sweep
  for (int t0 = 0; t0 < R; t0 += kSelTile) {
    const int tn = min(kSelTile, R - t0);
    if (R > kSelTile || t0 > 0) stage(t0, tn);  // single-tile clouds are still resident from sweep 1
    const int steps = (tn + 31) / 32;           // sentinel padding: d ~ 3e36 > tau
    for (int s = 0; s < steps; ++s) {
      unsigned m = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float4 p = sm.ref[s * 32 + i];
        if (sqdist(qx, qy, qz, p.x, p.y, p.z) <= thr) m |= 1u << i;
      }
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1u;
        const float4 p = sm.ref[s * 32 + i];
        if (sqdist(qx, qy, qz, p.x, p.y, p.z) == tau) {
          if (++eq_taken > k) continue;
          if (eq_taken == k) thr = tau > 0.f ? __uint_as_float(__float_as_uint(tau) - 1u) : -1.f;  // now strictly below
        }
        if (count < kSelCap) sm.cand[count][tid] = (unsigned short)(t0 + s * 32 + i);
        ++count;
      }
    }
  }
  const bool overflow = count > kSelCap;
  if (live) flags[(size_t)b * Q + q] = overflow ? 1 : 0;
  __syncthreads();  // every thread is done with the reference tile: its storage becomes the top-k arrays
  if (!live || overflow) return;

  // ---- select: stable insertion by distance (candidates arrive in ascending index) ---------------------------------
  // The next candidate's coordinates are fetched while the current one is inserted.
  int have = 0;
  float nx = 0.f, ny = 0.f, nz = 0.f;
  int nr = 0;
  if (count > 0) {
    nr = sm.cand[0][tid];
    const float* p = Rb + 3 * (size_t)nr;
    nx = __ldg(p);
    ny = __ldg(p + 1);
    nz = __ldg(p + 2);
  }
  for (int c = 0; c < count; ++c) {
    const int r = nr;
    const float d = sqdist(qx, qy, qz, nx, ny, nz);
    if (c + 1 < count) {
      nr = sm.cand[c + 1][tid];
      const float* p = Rb + 3 * (size_t)nr;
      nx = __ldg(p);
      ny = __ldg(p + 1);
      nz = __ldg(p + 2);
    }
    if (have == k && !(d < sm.top.topd[k - 1][tid])) continue;
    int pos = have < k ? have : k - 1;
    while (pos > 0 && d < sm.top.topd[pos - 1][tid]) {
      sm.top.topd[pos][tid] = sm.top.topd[pos - 1][tid];
      sm.top.topi[pos][tid] = sm.top.topi[pos - 1][tid];
      --pos;
    }
    sm.top.topd[pos][tid] = d;
    sm.top.topi[pos][tid] = r;
    if (have < k) ++have;
  }
  const size_t o = ((size_t)b * Q + q) * k;
  for (int j = 0; j < k; ++j) {
    const bool filled = j < have;  // fewer than k only if distances were NaN (outside the contract): stay in range
    idx_out[o + j] = filled ? (int64_t)sm.top.topi[j][tid] : 0;
    if (dist_out) dist_out[o + j] = filled ? sm.top.topd[j][tid] : kInf;
  }
}

// Reference-range splits for the list kernel: only when the query count cannot fill the machine on its own.  In that
// mode a warp owns ONE query (8 per CTA) instead of four in turn.
static int knn_r_splits(int B, int R, int Q) {
  const long long ctas4 = (long long)B * ((Q + kKnnQueriesPerCta - 1) / kKnnQueriesPerCta);
  if (ctas4 >= 296 || R < 2 * kKnnTile) return 1;
  const long long ctas = (long long)B * ((Q + kKnnWarps - 1) / kKnnWarps);
  static long long target = 0;
  if (target == 0) {
    const char* e = std::getenv("IAPT_KNN_SPLIT_TARGET");  // development knob: total CTAs to aim for
    target = (e && std::atoi(e) > 0) ? std::atoi(e) : 600;
  }
  long long s = (target + ctas - 1) / ctas;       // default: ~4 CTAs (32 warps) per SM; more slices only add merge work
  const long long max_s = R / 1024;               // keep >= 1024 references per slice
  if (s > max_s) s = max_s;
  if (s > 32) s = 32;
  return s < 2 ? 1 : (int)s;
}

size_t knn_workspace_bytes(int B, int R, int Q, int k) {
  size_t bytes = k <= kSelMaxK ? sizeof(int) * (size_t)B * Q : 0;
  const int rs = knn_r_splits(B, R, Q);
  if (rs > 1) bytes += (sizeof(float) + sizeof(int)) * (size_t)B * Q * rs * k;
  return bytes;
}

template <int kSlots>
static void knn_list_launch(const float* ref, const float* qry, int B, int R, int Q, int k, int64_t* idx, float* dist,
                            const int* flags, void* split_ws, cudaStream_t stream) {
  const int rs = (flags == nullptr && split_ws != nullptr) ? knn_r_splits(B, R, Q) : 1;
  if (rs > 1) {
    const int q_blocks = (Q + kKnnWarps - 1) / kKnnWarps;
    dim3 grid((unsigned)((long long)B * q_blocks), rs);
    float* pd = static_cast<float*>(split_ws);
    int* pi = reinterpret_cast<int*>(pd + (size_t)B * Q * rs * k);
    knn_kernel<kSlots, 1><<<grid, kKnnThreads, 0, stream>>>(ref, qry, R, Q, k, q_blocks, idx, dist, nullptr, rs, pd, pi);
    const long long nq = (long long)B * Q;
    knn_merge_kernel<kSlots><<<(unsigned)((nq + kKnnWarps - 1) / kKnnWarps), kKnnThreads, 0, stream>>>(pd, pi, nq, k, rs,
                                                                                                     idx, dist);
    count_launch(2);
  } else if (flags == nullptr) {
    // one query per warp (measured 10-15 % faster than four queries per warp in turn on every encoder shape)
    const int q_blocks = (Q + kKnnWarps - 1) / kKnnWarps;
    dim3 grid((unsigned)((long long)B * q_blocks), 1);
    knn_kernel<kSlots, 1><<<grid, kKnnThreads, 0, stream>>>(ref, qry, R, Q, k, q_blocks, idx, dist, nullptr, 1, nullptr,
                                                            nullptr);
    count_launch();
  } else {
    const int q_blocks = (Q + kKnnQueriesPerCta - 1) / kKnnQueriesPerCta;
    dim3 grid((unsigned)((long long)B * q_blocks), 1);
    knn_kernel<kSlots, 4><<<grid, kKnnThreads, 0, stream>>>(ref, qry, R, Q, k, q_blocks, idx, dist, flags, 1, nullptr,
                                                            nullptr);
    count_launch();
  }
}

int knn_launch(const float* ref, const float* qry, int B, int R, int Q, int k, int64_t* idx, float* dist,
               void* workspace, cudaStream_t stream) {
  static int force_list = -1;
  if (force_list < 0) {
    const char* e = std::getenv("IAPT_KNN_FORCE_LIST");
    force_list = (e && e[0] == '1') ? 1 : 0;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  int* flags = (k <= kSelMaxK && ws != nullptr) ? reinterpret_cast<int*>(ws) : nullptr;
  void* split_ws = (ws != nullptr && knn_r_splits(B, R, Q) > 1) ? ws + (k <= kSelMaxK ? sizeof(int) * (size_t)B * Q : 0)
                                                                : nullptr;
  // The select path needs enough queries to fill the machine with one thread each; small calls keep the
  // warp-per-query list kernel (32 lanes per query).
  if (k <= kSelMaxK && R <= 65536 && flags != nullptr && !force_list && (long long)B * Q >= 65536) {
    static PerDeviceOnce once;
    if (once.pending()) {
      cudaFuncSetAttribute(knn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelSmem));
      once.done();
    }
    const int sb = (Q + kSelThreads - 1) / kSelThreads;
    knn_select_kernel<<<(unsigned)((long long)B * sb), kSelThreads, sizeof(SelSmem), stream>>>(
        ref, qry, R, Q, k, sb, idx, dist, flags);
    count_launch();
    knn_list_launch<1>(ref, qry, B, R, Q, k, idx, dist, flags, nullptr, stream);
    return (int)cudaGetLastError();
  }
  if (k <= 32)
    knn_list_launch<1>(ref, qry, B, R, Q, k, idx, dist, nullptr, split_ws, stream);
  else
    knn_list_launch<2>(ref, qry, B, R, Q, k, idx, dist, nullptr, split_ws, stream);
  return (int)cudaGetLastError();
}

}  // namespace iapt
