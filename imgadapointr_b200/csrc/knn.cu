// knn.cu - fused distance + top-k selection (kNN grouping) for sm_100a.
//
// Replaces knn_point = square_distance + torch.topk(largest=False, sorted=False)
// (reference models/Transformer_utils.py:19-51), which materialises the [B,Q,R] FP32 matrix (805 MB at
// B=48, Q=R=2048) through a K=3 SGEMM and re-reads it in topk.  Here nothing is materialised.  Three paths:
//  * knn_kernel (any k <= 64, any batch): reference points are staged once per CTA in shared memory (float4) and every
//    warp owns one query, keeping the current k best as a sorted list spread over its lanes (lane i = i-th smallest).
//    A 32-reference step costs one direct-form distance per lane, a compare against the current k-th distance and one
//    ballot; the (rare) candidates are inserted with a ballot-computed position and one shuffle-up;
//  * knn_select_kernel (k <= 16 and at least 256 warps of queries - the encoder's calls): one THREAD per query, two
//    sweeps and a threshold (see below); 2.9x faster than the list kernel at 48 x 2048 x 2048, k = 16;
//  * knn_bin_kernel + knn_culled_kernel (k <= 16, 1024..4096 references, from 1024 warps of queries): both clouds binned
//    along a Hilbert curve, one bound per query, and only the tiles of 32 references whose box can hold a neighbour are
//    swept (a quarter of the pairs at 2048 points): another 1.4-1.8x at 2048 x 2048, 2.6x at 4096 x 4096.
// Output is ascending by (distance, index); on exact distance ties the LOWEST index wins (references are scanned
// in ascending index and only a strictly smaller distance displaces a list entry).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace iapt {

constexpr int kKnnThreads = 256;
constexpr int kKnnWarps = kKnnThreads / 32;
constexpr int kKnnTile = 2048;         // reference points per smem tile (32 KB of float4)

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
template <int kSlots>  // list entries per lane: k <= 32 * kSlots
__global__ void __launch_bounds__(kKnnThreads) knn_kernel(const float* __restrict__ ref, const float* __restrict__ qry,
                                                          int R, int Q, int k, int q_blocks,
                                                          int64_t* __restrict__ idx_out, float* __restrict__ dist_out,
                                                          int r_splits, float* __restrict__ part_d,
                                                          int* __restrict__ part_i) {
  __shared__ float4 sp[kKnnTile];  // one broadcast-free LDS.128 per reference
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x / q_blocks;
  const int q = (blockIdx.x - b * q_blocks) * kKnnWarps + warp;  // one query per warp (measured 10-15 % faster than
  const int split = blockIdx.y;                                   // four in turn on every encoder shape)
  const int per = (R + r_splits - 1) / r_splits;
  const int r_lo = split * per, r_hi = min(R, r_lo + per);
  const float* __restrict__ Rb = ref + (size_t)b * R * 3;
  const float* __restrict__ Qb = qry + (size_t)b * Q * 3;
  const float kInf = __int_as_float(0x7f800000);

  const int qq = q < Q ? q : 0;
  const float qx = __ldg(Qb + 3 * (size_t)qq + 0);
  const float qy = __ldg(Qb + 3 * (size_t)qq + 1);
  const float qz = __ldg(Qb + 3 * (size_t)qq + 2);
  float ld[kSlots];  // sorted list: entry (s*32 + lane)
  int li[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    ld[s] = kInf;
    li[s] = 0x7fffffff;
  }
  float tau = kInf;

  for (int t0 = r_lo; t0 < r_hi; t0 += kKnnTile) {
    const int tn = min(kKnnTile, r_hi - t0);
    __syncthreads();
    for (int r = threadIdx.x; r < tn; r += kKnnThreads) {
      const float* p = Rb + 3 * (size_t)(t0 + r);
      sp[r] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
    }
    __syncthreads();
    if (q >= Q) continue;  // warp-uniform
    int r_first = 0;
    if (t0 == r_lo) {  // empty list: seed it with the first 32 references by one warp sort
      float d = kInf;
      int i = t0 + lane;
      if (lane < tn) {
        const float4 p = sp[lane];
        d = sqdist(qx, qy, qz, p.x, p.y, p.z);
      }
      warp_sort_pairs(d, i);
      ld[0] = d;
      li[0] = (d < kInf) ? i : 0x7fffffff;
      float kth = ld[0];
      if (kSlots > 1 && k > 32) kth = ld[kSlots - 1];
      tau = __shfl_sync(0xffffffffu, kth, (k - 1) & 31);
      r_first = 32;
    }
    for (int r0 = r_first; r0 < tn; r0 += 32) {
      const int r = r0 + lane;
      float d = kInf;
      if (r < tn) {
        const float4 p = sp[r];
        d = sqdist(qx, qy, qz, p.x, p.y, p.z);
      }
      knn_insert<kSlots>(d, t0 + r, k, ld, li, tau);
    }
  }

  if (q >= Q) return;
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int e = s * 32 + lane;
    if (e < k) {
      if (r_splits > 1) {
        const size_t o = (((size_t)b * Q + q) * r_splits + split) * k + e;
        part_d[o] = ld[s];
        part_i[o] = li[s];
      } else {
        const size_t o = ((size_t)b * Q + q) * k + e;
        // an unfilled slot (only possible when distances were NaN - outside the contract) must stay a valid index
        idx_out[o] = (int64_t)(li[s] == 0x7fffffff ? 0 : li[s]);
        if (dist_out) dist_out[o] = ld[s];
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
// (warp-uniform LDS.128 from a shared-memory tile) twice:
//   sweep 1: 32 running minima over 32 classes of distinct references - 1 FMNMX per pair.  The k-th smallest of the
//            32 minima, tau, is an upper bound on the k-th nearest distance.  Only every kStride-th group of four
//            references is looked at: any distinct references bound the k-th distance, half the sweep is saved and the
//            candidate list grows from ~22 to ~44 entries (k = 16);
//   sweep 2: references with d <= tau set bits of a per-thread mask (FSETP + SEL, branch-free so loads and FP chains
//            overlap); once per 32 references the set bits are appended, in ascending index, to a per-thread candidate
//            list (u16 indices in shared memory);
//   select : a branch-free stable insert of the candidates into a register-resident sorted list gives the k nearest
//            ascending by (distance, index) - the same lowest-index tie rule.
// Layout choices, each from a measurement (profiles/r01f_*, scripts/microbench/lds_bcast.cu):
//  * a warp-uniform LDS costs ~3.3 issue cycles whatever its width, so the tile is a RAW copy of the [R,3] rows and
//    four references come out of three LDS.128 (0.75 loads per pair instead of 1);
//  * a warp is the unit of work (32 queries, its own tile, __syncwarp only): 48 x 2048 queries are 3072 warps = 20.8 per
//    SM, so every SM gets 20 or 21 of them (256-thread CTAs gave 2 or 3 per SM: 14 % of the machine idle);
//  * calls with few queries put kSplit warps on the same 32 queries, each sweeping a slice of the references: the
//    slices' class minima are combined through shared memory after sweep 1 (one tau per query, as without slices) and
//    warp 0 selects from the slices' candidate lists in slice order.
// A query whose candidate list overflows is redone inside the kernel by k exact minimum sweeps (see below).
constexpr int kSelMaxK = 16;
constexpr int kS2Tile = 128;                  // references per tile: 96 float4 = the rows' 384 floats verbatim
constexpr int kS2Vec = kS2Tile * 3 / 4;       // float4 per tile
constexpr int kS2Cap = 112;                   // candidates per query (and slice)
constexpr int kS2TieCheckFrom = 80;           // candidates appended before ties at the threshold start to be counted

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

// ascending compare-exchange / bitonic networks on 64-bit keys held in registers (indices are compile-time after unrolling)
__device__ __forceinline__ void cswap_u64(u64& a, u64& b) {
  const bool s = b < a;
  const u64 lo = s ? b : a, hi = s ? a : b;
  a = lo;
  b = hi;
}
template <int N>
__device__ __forceinline__ void bitonic_merge_u64(u64 (&v)[N]) {  // bitonic sequence -> ascending
#pragma unroll
  for (int stride = N / 2; stride > 0; stride >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if ((i & stride) == 0) cswap_u64(v[i], v[i + stride]);
  }
}
template <int N>
__device__ __forceinline__ void bitonic_sort_u64(u64 (&v)[N]) {
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int j = i ^ stride;
        if (j > i) {
          if ((i & size) == 0)
            cswap_u64(v[i], v[j]);
          else
            cswap_u64(v[j], v[i]);
        }
      }
    }
  }
}

constexpr int kCuBatch = 8;  // candidates folded at once (select and culled kernels)

struct __align__(16) S2Smem {
  float4 tile[kS2Vec];
  union {
    unsigned short cand[kS2Cap][32];  // sweep 2 -> select
    float minima[32][32];             // kSplit > 1: this slice's class minima, exchanged after sweep 1
    u64 keys[kSelMaxK][32];           // kSplit > 1: this slice's folded top-k, merged by warp 0
  };
};

template <int kStride, int kK, int kSplit>
__global__ void __launch_bounds__(32 * kSplit, kSplit == 1 ? 21 : 24 / kSplit)
    knn_select_kernel(const float* __restrict__ ref, const float* __restrict__ qry, int R, int Q, int k, int q_blocks,
                      int aligned, int64_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  __shared__ S2Smem smem[kSplit];
  __shared__ int counts[kSplit][32];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;  // slice of the references this warp sweeps
  S2Smem& sm = smem[w];
  const int b = blockIdx.x / q_blocks;
  const int q = (blockIdx.x - b * q_blocks) * 32 + lane;
  const bool live = q < Q;
  const float* __restrict__ Rb = ref + (size_t)b * R * 3;
  const float kInf = __int_as_float(0x7f800000);
  // slice [lo, hi): whole tiles, so slice starts stay 16-byte aligned whenever the cloud is
  const int per = ((R + kSplit - 1) / kSplit + kS2Tile - 1) / kS2Tile * kS2Tile;
  const int lo = min(R, w * per), hi = min(R, lo + per);
  const int fl0 = 3 * lo, nfl = 3 * hi;  // float range of the slice inside the cloud
  const int n_tiles = (hi - lo + kS2Tile - 1) / kS2Tile;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) {
    const float* qp = qry + ((size_t)b * Q + q) * 3;
    qx = __ldg(qp);
    qy = __ldg(qp + 1);
    qz = __ldg(qp + 2);
  }

  // Tile t -> three float4 per lane, fetched one tile ahead of its use.  The loads are unconditional (addresses
  // clamped into the slice); publish() swaps in the sentinels for floats past its end.
  auto fetch_range = [&](int t, int first, int end, float4 (&v)[3]) {
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const int f0 = first + (t * kS2Vec + e * 32 + lane) * 4;
      if (aligned) {
        v[e] = __ldg(reinterpret_cast<const float4*>(Rb + min(f0, end - 4)));
      } else {
        v[e].x = __ldg(Rb + min(f0 + 0, end - 1));
        v[e].y = __ldg(Rb + min(f0 + 1, end - 1));
        v[e].z = __ldg(Rb + min(f0 + 2, end - 1));
        v[e].w = __ldg(Rb + min(f0 + 3, end - 1));
      }
    }
  };
  auto publish_range = [&](int t, int first, int end, const float4 (&v)[3]) {
    __syncwarp();  // every lane is done with the previous tile
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const int f0 = first + (t * kS2Vec + e * 32 + lane) * 4;
      float4 wv = v[e];
      if (f0 + 0 >= end) wv.x = kPadCoord();
      if (f0 + 1 >= end) wv.y = kPadCoord();
      if (f0 + 2 >= end) wv.z = kPadCoord();
      if (f0 + 3 >= end) wv.w = kPadCoord();
      sm.tile[e * 32 + lane] = wv;
    }
    __syncwarp();
  };
  auto fetch = [&](int t, float4 (&v)[3]) { fetch_range(t, fl0, nfl, v); };
  auto publish = [&](int t, const float4 (&v)[3]) { publish_range(t, fl0, nfl, v); };

  // ---- sweep 1: 32 class minima over the sampled references ---------------------------------------------------
  float mn[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) mn[i] = kInf;
  float4 nx[3];
  if (n_tiles > 0) fetch(0, nx);
  for (int t = 0; t < n_tiles; ++t) {
    publish(t, nx);
    if (t + 1 < n_tiles) fetch(t + 1, nx);
#pragma unroll
    for (int g0 = 0; g0 < kS2Tile / 4; g0 += 8 * kStride) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* G = &sm.tile[3 * (g0 + j * kStride)];
        const float4 A = G[0], Bv = G[1], C = G[2];
        mn[4 * j + 0] = fminf(mn[4 * j + 0], sqdist(qx, qy, qz, A.x, A.y, A.z));
        mn[4 * j + 1] = fminf(mn[4 * j + 1], sqdist(qx, qy, qz, A.w, Bv.x, Bv.y));
        mn[4 * j + 2] = fminf(mn[4 * j + 2], sqdist(qx, qy, qz, Bv.z, Bv.w, C.x));
        mn[4 * j + 3] = fminf(mn[4 * j + 3], sqdist(qx, qy, qz, C.y, C.z, C.w));
      }
    }
  }
  if (kSplit > 1) {
    // the slices' minima of a class belong to different references: their minimum is the class minimum of the cloud
#pragma unroll
    for (int i = 0; i < 32; ++i) sm.minima[i][lane] = mn[i];
    __syncthreads();
#pragma unroll
    for (int o = 1; o < kSplit; ++o) {
      const S2Smem& other = smem[(w + o) % kSplit];
#pragma unroll
      for (int i = 0; i < 32; ++i) mn[i] = fminf(mn[i], other.minima[i][lane]);
    }
    __syncthreads();  // minima[] aliases cand[]
  }
  sort32(mn);
  float tau = mn[0];
#pragma unroll
  for (int i = 1; i < kK; ++i)
    if (i < k) tau = mn[i];

  // ---- sweep 2: candidates (d <= thr) ------------------------------------------------------------------------------
  // References exactly AT tau (hundreds in a zero-padded cloud) are all taken until the list holds kS2TieCheckFrom
  // entries - the lowest-index ones are among them - then counted, and after k more the threshold drops just below
  // tau.  Typical queries (~44 candidates) never pay for the check.
  int count = 0, eq_taken = 0;
  float thr = tau;
  const float* tf = reinterpret_cast<const float*>(sm.tile);
  // kSplit == 1: the running top-k lives in registers during the sweep.  A candidate list with fewer than 32 free slots
  // is FOLDED into it (every lane of the warp at once) and the k-th entry becomes the new, exact threshold - so a loose
  // class-minima bound (index layouts correlated with position: ordered surface samplings) or hundreds of exact ties
  // cost extra folds instead of the exact-fallback sweeps below.  Candidates arrive in ascending index and the insert
  // is stable, so equal distances keep the lowest indices.
  u64 tk[kK];  // ascending (distance bits, index) keys: the k nearest seen so far
#pragma unroll
  for (int j = 0; j < kK; ++j) tk[j] = ~0ull;
  // kCuBatch candidates of a list at a time are re-evaluated (their loads in flight together), sorted by a register
  // network and merged into the sorted top-k by one bitonic split + merge: ~45 instructions per candidate where an
  // insertion costs ~100.  Every lane of the warp takes part (n = 0 for a lane with nothing to fold).
  auto fold_list = [&](const S2Smem& src, int n) {
    const int n_max = redux_max_s32(n);
    for (int c0 = 0; c0 < n_max; c0 += kCuBatch) {
      u64 nk[kCuBatch];
#pragma unroll
      for (int j = 0; j < kCuBatch; ++j) {
        nk[j] = ~0ull;
        if (c0 + j < n) {
          const unsigned cr = src.cand[c0 + j][lane];
          const float* p = Rb + 3 * (size_t)cr;
          nk[j] = ((u64)__float_as_uint(sqdist(qx, qy, qz, __ldg(p), __ldg(p + 1), __ldg(p + 2))) << 32) | cr;
        }
      }
      bitonic_sort_u64<kCuBatch>(nk);
#pragma unroll
      for (int j = 0; j < kK; ++j) {  // the kK smallest of (tk ascending, nk descending): a bitonic sequence
        const int rj = kK - 1 - j;
        if (rj < kCuBatch) tk[j] = nk[rj] < tk[j] ? nk[rj] : tk[j];
      }
      bitonic_merge_u64<kK>(tk);
    }
  };
  auto fold = [&]() {
    fold_list(sm, count);
    count = 0;
    u64 kth = tk[0];
#pragma unroll
    for (int j = 1; j < kK; ++j)
      if (j < k) kth = tk[j];
    // k references are in the list: the sweep runs in ascending index and ties keep the LOWEST indices, so from now on
    // only a strictly smaller distance can enter (kth stays empty while fewer than k references were folded)
    if (kth != ~0ull) {
      const float kd = __uint_as_float((unsigned)(kth >> 32));
      thr = fminf(thr, kd > 0.f ? __uint_as_float(__float_as_uint(kd) - 1u) : -1.f);
    }
  };
  if (n_tiles > 0) fetch(0, nx);
  for (int t = 0; t < n_tiles; ++t) {
    publish(t, nx);
    if (t + 1 < n_tiles) fetch(t + 1, nx);
#pragma unroll 1
    for (int c = 0; c < kS2Tile / 32; ++c) {
      unsigned m = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* G = &sm.tile[3 * (c * 8 + j)];
        const float4 A = G[0], Bv = G[1], C = G[2];
        if (sqdist(qx, qy, qz, A.x, A.y, A.z) <= thr) m |= 1u << (4 * j + 0);
        if (sqdist(qx, qy, qz, A.w, Bv.x, Bv.y) <= thr) m |= 1u << (4 * j + 1);
        if (sqdist(qx, qy, qz, Bv.z, Bv.w, C.x) <= thr) m |= 1u << (4 * j + 2);
        if (sqdist(qx, qy, qz, C.y, C.z, C.w) <= thr) m |= 1u << (4 * j + 3);
      }
      if (kSplit == 1 && !live) m = 0u;  // a dead lane must not grow a list nobody folds
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1u;
        const int r = lo + t * kS2Tile + c * 32 + i;
        if (r >= hi) continue;  // sentinel (only reachable when tau itself is a sentinel distance or +inf)
        if (kSplit == 1) {  // (the list is folded before a chunk could overfill it)
          sm.cand[count][lane] = (unsigned short)r;
          ++count;
          continue;
        }
        if (count >= kS2TieCheckFrom) {
          const float* p = tf + 3 * (c * 32 + i);
          if (sqdist(qx, qy, qz, p[0], p[1], p[2]) == tau) {
            if (++eq_taken > k) continue;
            if (eq_taken == k) thr = tau > 0.f ? __uint_as_float(__float_as_uint(tau) - 1u) : -1.f;
          }
        }
        if (count < kS2Cap) sm.cand[count][lane] = (unsigned short)r;
        ++count;
      }
      if (kSplit == 1 && __any_sync(0xffffffffu, live && count > kS2Cap - 32)) fold();
    }
  }
  bool overflow = kSplit > 1 && count > kS2Cap;
  if (kSplit > 1) {  // every warp folds its own slice's list; warp 0 merges the sorted lists (keys carry the index: any order)
    fold_list(sm, live ? min(count, kS2Cap) : 0);  // (an overflowed list's result is discarded below)
    counts[w][lane] = count;
    if (w != 0) {
#pragma unroll
      for (int j = 0; j < kK; ++j) sm.keys[j][lane] = tk[j];  // (cand[] is done with)
    }
    __syncthreads();
    if (w != 0) return;
#pragma unroll
    for (int o = 1; o < kSplit; ++o) overflow |= counts[o][lane] > kS2Cap;
  }
  const size_t o = ((size_t)b * Q + q) * k;

  // ---- overflowed lists (a pure rank-statistics tail once ties are capped: practically never) -----------------------
  // Exact but slow: the warp sweeps the WHOLE cloud k more times and each overflowed lane picks, per sweep, the
  // smallest (distance, index) above its previous pick.  All lanes keep staging tiles; only those lanes compute.
  if (__any_sync(0xffffffffu, overflow)) {
    const int all_tiles = (R + kS2Tile - 1) / kS2Tile;
    float pd = -1.f;  // previous pick; distances are >= 0
    int pi = -1;
    for (int j = 0; j < k; ++j) {
      float bd = kInf;
      int bi = 0x7fffffff;
      fetch_range(0, 0, 3 * R, nx);
      for (int t = 0; t < all_tiles; ++t) {
        publish_range(t, 0, 3 * R, nx);
        if (t + 1 < all_tiles) fetch_range(t + 1, 0, 3 * R, nx);
        if (overflow) {
          const int tn = min(kS2Tile, R - t * kS2Tile);
          for (int i = 0; i < tn; ++i) {
            const float d = sqdist(qx, qy, qz, tf[3 * i], tf[3 * i + 1], tf[3 * i + 2]);
            const int r = t * kS2Tile + i;
            const bool after = d > pd || (d == pd && r > pi);
            if (after && d < bd) {  // strict: the first (lowest-index) reference at the minimal distance stays
              bd = d;
              bi = r;
            }
          }
        }
      }
      if (overflow && live) {
        idx_out[o + j] = (int64_t)(bi == 0x7fffffff ? 0 : bi);
        if (dist_out) dist_out[o + j] = bd;
      }
      pd = bd;
      pi = bi;
    }
  }
  // ---- select: the k nearest of the listed candidates (every lane of the warp folds; only live, non-overflowed ones list) ---
  if (kSplit == 1) {
    fold();
  } else {
#pragma unroll
    for (int sl = 1; sl < kSplit; ++sl) {  // one bitonic split + merge per slice
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        const u64 ok = smem[sl].keys[kK - 1 - j][lane];
        tk[j] = ok < tk[j] ? ok : tk[j];
      }
      bitonic_merge_u64<kK>(tk);
    }
  }
  if (!live || overflow) return;
#pragma unroll
  for (int j = 0; j < kK; ++j) {
    if (j < k) {
      // fewer than k entries only if distances were NaN (outside the contract): index 0 keeps consumers in range
      idx_out[o + j] = tk[j] == ~0ull ? 0 : (int64_t)(unsigned)tk[j];
      if (dist_out) dist_out[o + j] = tk[j] == ~0ull ? kInf : __uint_as_float((unsigned)(tk[j] >> 32));
    }
  }
}

// =====================================================================================================================
// Spatially culled selection (k <= 16, 1024 <= R <= 4096): bin, bound, cull, select - still exact
// =====================================================================================================================
// The select kernel above evaluates every (query, reference) pair 1.5 times; this path evaluates only the pairs that
// can matter (a quarter of them at 2048 points) and keeps the selection machinery cheap.
//  * knn_bin_kernel orders each cloud along a 3-D HILBERT curve over a 16^3 grid by ONE counting sort (shared-memory
//    histogram, scan, scatter; the order inside a cell is whatever the atomics give - it only decides which tile a point
//    lands in, never a result).  Consecutive Hilbert cells are face neighbours, so a tile of 32 consecutive points has a
//    box about a third smaller than a Z-order tile: a warp of 32 consecutive queries needs 15-17 of the 64 tiles of a
//    2048-point cloud (Z-order: 21-26; a k-d partition would need 12-15; simulated on the bench clouds).  Every tile
//    gets its bounding box, its smallest original index and a flag "all 32 points coincide".
//  * knn_culled_kernel computes ONE bound per query - the k-th smallest of 32 class minima over the n_seeds tiles
//    nearest to the warp's queries - and then the list of tiles ANY of its queries needs (box distance <= bound) is
//    fixed: the sweep over it has no data-dependent control flow between tiles and prefetches freely.  (The first cut
//    of this path visited tiles nearest-first with a moving bound: a dependent chain cull test -> load -> stage -> sweep
//    per tile that left the kernel latency-bound at 0.16 ms; it also used a Z-order bitonic sort of 28 us.)
//  * the sweep's test is the sign of f = T - d evaluated as three FMAs onto T (T = the bound inflated by 2^-20, far
//    above the rounding of either side: a prefilter that can only let too MANY through), collected with one funnel
//    shift per reference: 7 instructions per pair instead of 9.  Candidates are re-evaluated with the exact direct-form
//    distance when they are folded into the (distance bits, ORIGINAL index) ordered top-k, so values, order and the
//    lowest-index tie rule are those of every other path although the sweep runs in curve order.
//  * the fold takes 8 candidates at a time: sorted by a register network and merged into the sorted top-k by one
//    bitonic split + merge (~45 instructions per candidate; an insertion costs ~100).
//  * duplicates (the zero padding of a PCN partial: hundreds of references at the origin tie at every bound): a cell
//    spanning several tiles is put in ascending original index by the bin kernel; once a query holds k references, a
//    candidate is checked exactly against its k-th (distance, index) before it is listed, and a tile that is one point
//    is skipped when that point ties or loses even with the tile's smallest index.
struct __align__(16) PrTileBox {
  float4 lo, hi;  // lo.w: smallest original index in the tile (int bits); hi.w: 1 when all its points coincide
};
constexpr int kPrMaxN = 4096;  // points per cloud on this path (4 per thread of the bin kernel; tile ids fit 7 bits)
constexpr int kPrMinR = 1024;

constexpr int kBinThreads = 1024;
constexpr int kBinPer = kPrMaxN / kBinThreads;  // points per thread, kept in registers
constexpr int kBinCells = 4096;                 // 16^3 Hilbert cells
constexpr int kBinBig = 8;                      // cells of more than one tile that get re-ordered by index
constexpr int kCuCap = 80;                      // candidates per query
constexpr int kCuMaxTiles = kPrMaxN / 32;

// float <-> unsigned whose order is the float order (for REDUX min / max over signed values)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Hilbert index of a cell (4 bits per axis -> 12 bits), Skilling's transposition algorithm.
__device__ __forceinline__ unsigned hilbert12(unsigned X0, unsigned X1, unsigned X2) {
#pragma unroll
  for (unsigned Qb = 8; Qb > 1; Qb >>= 1) {
    const unsigned P = Qb - 1;
    if (X0 & Qb) X0 ^= P;
    if (X1 & Qb) {
      X0 ^= P;
    } else {
      const unsigned t = (X0 ^ X1) & P;
      X0 ^= t;
      X1 ^= t;
    }
    if (X2 & Qb) {
      X0 ^= P;
    } else {
      const unsigned t = (X0 ^ X2) & P;
      X0 ^= t;
      X2 ^= t;
    }
  }
  X1 ^= X0;
  X2 ^= X1;
  unsigned t = 0;
#pragma unroll
  for (unsigned Qb = 8; Qb > 1; Qb >>= 1)
    if (X2 & Qb) t ^= Qb - 1;
  X0 ^= t;
  X1 ^= t;
  X2 ^= t;
  unsigned key = 0;
#pragma unroll
  for (int bit = 3; bit >= 0; --bit)
    key = (key << 3) | (((X0 >> bit) & 1u) << 2) | (((X1 >> bit) & 1u) << 1) | ((X2 >> bit) & 1u);
  return key;
}

// grid (B, 1 or 2): blockIdx.y == 0 bins the references (and writes their tile boxes), 1 the queries.
__global__ void __launch_bounds__(kBinThreads) knn_bin_kernel(const float* __restrict__ ref, int R, int Rp,
                                                              float4* sref, PrTileBox* __restrict__ boxes,
                                                              const float* __restrict__ qry, int Q, int Qp, float4* sqry) {
  __shared__ __align__(16) unsigned hist[kBinCells];
  __shared__ unsigned red[6][32];
  __shared__ unsigned wsum[32];
  __shared__ unsigned big_n, big_start[kBinBig], big_count[kBinBig];
  const bool is_q = blockIdx.y == 1;
  const int n = is_q ? Q : R, n_pad = is_q ? Qp : Rp;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ P = (is_q ? qry : ref) + (size_t)b * n * 3;
  float4* out = (is_q ? sqry : sref) + (size_t)b * n_pad;
  const float kInf = __int_as_float(0x7f800000);

  // ---- the thread's points, the cloud's bounding box ------------------------------------------------------------------
  float px[kBinPer], py[kBinPer], pz[kBinPer];
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
#pragma unroll
  for (int j = 0; j < kBinPer; ++j) {
    const int i = tid + j * kBinThreads;
    px[j] = py[j] = pz[j] = 0.f;
    if (i < n) {
      px[j] = __ldg(P + 3 * (size_t)i);
      py[j] = __ldg(P + 3 * (size_t)i + 1);
      pz[j] = __ldg(P + 3 * (size_t)i + 2);
      const unsigned ux = f2ord(px[j]), uy = f2ord(py[j]), uz = f2ord(pz[j]);
      lo[0] = min(lo[0], ux), hi[0] = max(hi[0], ux);
      lo[1] = min(lo[1], uy), hi[1] = max(hi[1], uy);
      lo[2] = min(lo[2], uz), hi[2] = max(hi[2], uz);
    }
  }
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const unsigned l = redux_min_u32(lo[e]), h = redux_max_u32(hi[e]);
    if (lane == 0) {
      red[e][warp] = l;
      red[3 + e][warp] = h;
    }
  }
#pragma unroll
  for (int j = 0; j < kBinCells / kBinThreads; ++j) hist[tid + j * kBinThreads] = 0u;
  if (tid == 0) big_n = 0u;
  __syncthreads();
  float blo[3], scale[3];
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const float l = ord2f(redux_min_u32(red[e][lane])), h = ord2f(redux_max_u32(red[3 + e][lane]));
    blo[e] = l;
    scale[e] = (h > l) ? 16.f / (h - l) : 0.f;  // (NaN / Inf coordinates: any cell is a valid cell)
  }

  // ---- cell of every point, its rank inside the cell --------------------------------------------------------------------
  unsigned cell[kBinPer], rank[kBinPer];
#pragma unroll
  for (int j = 0; j < kBinPer; ++j) {
    const int i = tid + j * kBinThreads;
    cell[j] = 0u;
    rank[j] = 0u;
    if (i < n) {
      const int cx = min(15, max(0, (int)((px[j] - blo[0]) * scale[0])));
      const int cy = min(15, max(0, (int)((py[j] - blo[1]) * scale[1])));
      const int cz = min(15, max(0, (int)((pz[j] - blo[2]) * scale[2])));
      cell[j] = hilbert12((unsigned)cx, (unsigned)cy, (unsigned)cz);
      rank[j] = atomicAdd(&hist[cell[j]], 1u);
    }
  }
  __syncthreads();

  // ---- exclusive scan of the 4096 counts (4 consecutive cells per thread) ------------------------------------------------
  {
    constexpr int kPer = kBinCells / kBinThreads;
    unsigned c[kPer], s = 0u;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      c[j] = hist[tid * kPer + j];
      s += c[j];
    }
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned wv = wsum[lane], winc = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    const unsigned wbase = __shfl_sync(0xffffffffu, winc - wv, warp);
    unsigned base = wbase + inc - s;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      hist[tid * kPer + j] = base;
      if (c[j] > 32u) {  // a cell spanning tiles: re-ordered by original index below
        const unsigned slot = atomicAdd(&big_n, 1u);
        if (slot < (unsigned)kBinBig) {
          big_start[slot] = base;
          big_count[slot] = c[j];
        }
      }
      base += c[j];
    }
  }
  __syncthreads();

  // ---- binned points (x, y, z, original index); sentinels behind them ------------------------------------------------------
#pragma unroll
  for (int j = 0; j < kBinPer; ++j) {
    const int i = tid + j * kBinThreads;
    if (i < n) out[hist[cell[j]] + rank[j]] = make_float4(px[j], py[j], pz[j], __int_as_float(i));
  }
  for (int i = n + tid; i < n_pad; i += kBinThreads)
    out[i] = make_float4(kPadCoord(), kPadCoord(), kPadCoord(), __int_as_float(-1));
  if (is_q) return;

  // ---- cells of more than one tile (duplicated points: the zero padding of a PCN partial, hundreds of references at the
  // origin) are put in ascending ORIGINAL index: the first tiles of such a cell then hold its lowest indices, and once the
  // top-k of a query is full of them the later tiles are skipped by their minimum index (knn_culled_kernel).  Rank by
  // counting, the indices staged in hist[] (free now); only the order inside the cell changes - never a result. -----------
  __syncthreads();
  {
    int* sidx = reinterpret_cast<int*>(hist);
    const int nb = (int)min(big_n, (unsigned)kBinBig);
    for (int g = 0; g < nb; ++g) {
      const int start = (int)big_start[g], m = (int)big_count[g];
      float4 v[kBinPer];
#pragma unroll
      for (int j = 0; j < kBinPer; ++j) {
        const int e = tid + j * kBinThreads;
        if (e < m) {
          v[j] = __ldcg(out + start + e);
          sidx[e] = __float_as_int(v[j].w);
        }
      }
      if (tid < 3 && m + tid < kBinCells) sidx[m + tid] = 0x7fffffff;  // (the last int4 read runs up to 3 entries past m)
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kBinPer; ++j) {
        const int e = tid + j * kBinThreads;
        if (e < m) {
          const int mine = __float_as_int(v[j].w);
          int rk = 0;
          const int4* s4 = reinterpret_cast<const int4*>(sidx);
          for (int f = 0; f < (m + 3) / 4; ++f) {
            const int4 q4 = s4[f];
            rk += (q4.x < mine ? 1 : 0) + (q4.y < mine ? 1 : 0) + (q4.z < mine ? 1 : 0) + (q4.w < mine ? 1 : 0);
          }
          out[start + rk] = v[j];
        }
      }
      __syncthreads();
    }
  }

  // ---- tile boxes (this CTA's own writes, visible after the barrier; read around L1).  lo.w = the smallest original
  // index in the tile, hi.w = 1 when all its points coincide ---------------------------------------------------------------
  for (int t = warp; t < n_pad / 32; t += kBinThreads / 32) {
    const float4 v = __ldcg(out + t * 32 + lane);
    const bool real = __float_as_int(v.w) >= 0;
    const int min_idx = redux_min_s32(real ? __float_as_int(v.w) : 0x7fffffff);
    const unsigned lx = redux_min_u32(real ? f2ord(v.x) : 0xffffffffu), hx = redux_max_u32(real ? f2ord(v.x) : 0u);
    const unsigned ly = redux_min_u32(real ? f2ord(v.y) : 0xffffffffu), hy = redux_max_u32(real ? f2ord(v.y) : 0u);
    const unsigned lz = redux_min_u32(real ? f2ord(v.z) : 0xffffffffu), hz = redux_max_u32(real ? f2ord(v.z) : 0u);
    if (lane == 0) {
      PrTileBox bx;
      bx.lo = make_float4(ord2f(lx), ord2f(ly), ord2f(lz), __int_as_float(min_idx));
      bx.hi = make_float4(ord2f(hx), ord2f(hy), ord2f(hz), __int_as_float((lx == hx && ly == hy && lz == hz) ? 1 : 0));
      boxes[(size_t)b * (n_pad / 32) + t] = bx;
    }
  }
  (void)kInf;
}

// every warp's last claim fails and still moves its end of the deque by one: the bias keeps the tail of a short list
// from wrapping below zero (at most kG - 1 helpers, one failed claim each)
constexpr int kCuDequeBias = 8;

// One CTA = kG warps = kG TASKS (consecutive tiles of 32 queries of one cloud, sharing its tile boxes).  The work of a
// task varies 3x with where the curve cuts the cloud (17 tiles on average, 35-50 for a warp whose queries straddle a jump),
// all tasks are resident from the start, and a warp alone on its scheduler runs at a third of the shared rate - so the
// kernel lasts as long as its heaviest task unless the others help.  Each task's tile list is a deque in shared memory
// (head in the low half of one word, tail in the high half, one atomicAdd per claim: the owner takes tiles from the
// front, a helper from the back, and whoever sees head >= tail in the value the atomic returns stops).  A warp that
// finishes its own task un-helped writes its result, then helps the task of its CTA with the most tiles left: same
// queries (from shared memory), the bound the owner last published, a top-k of its own - which the owner merges after the
// CTA's final barrier (one bitonic split + merge per helper).
struct __align__(16) CuTaskSmem {
  float4 stage[2][32];  // a tile of 32 references (x, y, z, original index), double-buffered: cp.async lands here
  union {
    unsigned short cand[kCuCap][32];  // sweep -> fold
    float minima[32][32];             // solo: this warp's class minima, exchanged after the bound pass
    u64 keys[kSelMaxK][32];           // a warp's top-k for another warp to merge (a helper's; in solo mode every warp's)
  };
  unsigned char tlist[kCuMaxTiles];
  float4 qv[32];        // the task's queries (x, y, z, original index)
  float tpub[32];       // the bounds its owner last published
  unsigned deque;       // tlist[head, tail): (head + bias) | (tail + bias) << 16; 0 = not open yet
  unsigned helped_by;   // bit h: warp h took tiles of this task
};

template <int kK, int kG>
__global__ void __launch_bounds__(32 * kG, kG == 1 ? 21 : kG == 2 ? 11 : 6) knn_culled_kernel(const float4* __restrict__ sref,
                                                                              const PrTileBox* __restrict__ boxes,
                                                                              const float4* __restrict__ sqry, int R, int Rp,
                                                                              int Q, int Qp, int k, int q_blocks,
                                                                              int n_seeds, int probes, int solo,
                                                                              int64_t* __restrict__ idx_out,
                                                                              float* __restrict__ dist_out) {
  extern __shared__ float4 sbox[];  // [2 * nt]: lo, hi of every tile of the cloud
  __shared__ CuTaskSmem tsm[kG];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  CuTaskSmem& sm = tsm[w];
  // solo: ONE tile of 32 queries per CTA, shared by its kG warps - for calls whose tasks do not fill the machine (small
  // batches).  Every warp holds the same queries, takes every kG-th seed tile of the common ranking (the class minima are
  // exchanged through shared memory) and every kG-th tile of the cloud into a deque of its own; a warp that runs out takes
  // tiles from the fullest deque and keeps folding into ITS OWN top-k (same queries: no hand-over), and warp 0 merges the kG
  // sorted lists at the end.
  const int groups = solo ? q_blocks : (q_blocks + kG - 1) / kG;
  const int b = blockIdx.x / groups;
  const int qb = solo ? blockIdx.x - b * groups : (blockIdx.x - b * groups) * kG + w;
  // (without solo the last CTA of a cloud may hold fewer than kG tasks: those warps only help)
  const bool has_task = solo || qb < q_blocks;
  const int tw = solo ? kG : 1, tr = solo ? w : 0;  // this warp takes the seeds / tiles with number = tr (mod tw)
  const float kInf = __int_as_float(0x7f800000);
  const float4* __restrict__ Rb = sref + (size_t)b * Rp;
  const int nt = Rp / 32;
  float4 qv = make_float4(kPadCoord(), kPadCoord(), kPadCoord(), __int_as_float(-1));
  if (has_task) qv = __ldg(sqry + (size_t)b * Qp + qb * 32 + lane);
  {
    const float4* __restrict__ Bg = reinterpret_cast<const float4*>(boxes + (size_t)b * nt);
    for (int i = threadIdx.x; i < 2 * nt; i += 32 * kG) sbox[i] = __ldg(Bg + i);
  }
  if (kG > 1) {
    sm.qv[lane] = qv;
    if (lane == 0) {
      sm.deque = 0u;
      sm.helped_by = 0u;
    }
    __syncthreads();
  } else {
    __syncwarp();
  }

  // A tile travels from L2 to shared memory by cp.async (LDGSTS), 16 bytes per lane, issued one tile ahead: no register
  // holds it during the sweep of the tile before (a register prefetch was spilled there, and the spill store waited for
  // the load - ncu: the exposed L2 latency of every tile).
  auto fetch_tile = [&](int buf, int t) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.stage[buf][lane]);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n cp.async.commit_group;" ::"r"(dst), "l"(Rb + t * 32 + lane)
                 : "memory");
  };
  auto tile_landed = [&]() {  // every lane's part of the newest tile is in shared memory; every lane is done with the older one
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  };
  // prefilter threshold: d <= tau (both in float) implies T - d >= 0 in the fused evaluation of the sweep - three roundings
  // of at most 2^-24 relative to max(d, T) each against a margin of 2^-20 (+ an absolute one for the subnormal range)
  auto inflate = [](float v) { return __fmaf_rn(v, 1.0f + 9.5367431640625e-7f, 1e-37f); };

  // state of the task being swept (the warp's own first, then possibly one it helps)
  float qx = qv.x, qy = qv.y, qz = qv.z;
  int qi = __float_as_int(qv.w);
  bool live = qi >= 0;
  float T = kInf;
  int n_list = 0;

  if (has_task) {
    // box of the warp's queries
    float cl[3], ch[3];
    {
      const float q3[3] = {qx, qy, qz};
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        cl[e] = ord2f(redux_min_u32(live ? f2ord(q3[e]) : 0xffffffffu));
        ch[e] = ord2f(redux_max_u32(live ? f2ord(q3[e]) : 0u));
      }
    }
    // ---- bound: class minima over the n_seeds tiles nearest to the warp (distance from the centre of the query box to
    // the FARTHEST corner of the tile box: compact close tiles first).  Class = position inside the tile; different tiles
    // hold different references, so the k-th smallest of the 32 minima bounds the k-th neighbour distance. -------------------
    float tau;
    {
      unsigned key[kCuMaxTiles / 32];
      unsigned deg = 0u;  // bit i: this lane's i-th tile is one point 32 times (duplicates)
      const float cx = 0.5f * (cl[0] + ch[0]), cy = 0.5f * (cl[1] + ch[1]), cz = 0.5f * (cl[2] + ch[2]);
#pragma unroll
      for (int i = 0; i < kCuMaxTiles / 32; ++i) {
        const int t = lane + 32 * i;
        key[i] = 0xffffffffu;
        if (t < nt) {
          const float4 bl = sbox[2 * t], bh = sbox[2 * t + 1];
          const float fx = fmaxf(fabsf(cx - bl.x), fabsf(cx - bh.x));
          const float fy = fmaxf(fabsf(cy - bl.y), fabsf(cy - bh.y));
          const float fz = fmaxf(fabsf(cz - bl.z), fabsf(cz - bh.z));
          key[i] = (__float_as_uint(fx * fx + fy * fy + fz * fz) & 0xffffff80u) | (unsigned)t;  // >= +0: bits order like uints
          if (__float_as_int(bh.w) == 1) deg |= 1u << i;
        }
      }
      // Seeds alternate between the centre ranking and PROBES: the unused tile nearest to the query of one lane (0, the
      // last live one, 16, 8, 24, ...).  A warp whose 32 consecutive queries straddle a jump of the curve (or the block of
      // zero padding and the real points) has nothing near its centre; the probes give both ends their neighbourhood.
      const int n_live = __popc(__ballot_sync(0xffffffffu, live));  // live lanes are a prefix
      // a tile that is ONE point gives every class the same distance: a second one (the zero padding of a PCN partial
      // fills dozens) would add nothing, so taking one retires them all
      auto mark = [&](int id) {
        bool was_deg = false;
#pragma unroll
        for (int i = 0; i < kCuMaxTiles / 32; ++i)
          if (lane + 32 * i == id) {
            key[i] = 0xffffffffu;
            was_deg = (deg >> i) & 1u;
          }
        if (__any_sync(0xffffffffu, was_deg)) {
#pragma unroll
          for (int i = 0; i < kCuMaxTiles / 32; ++i)
            if ((deg >> i) & 1u) key[i] = 0xffffffffu;
        }
      };
      auto pick = [&](unsigned m) {  // -1: no tile left
        const unsigned wk = redux_min_u32(m);
        if (wk == 0xffffffffu) return -1;
        const int id = (int)(wk & 0x7fu);
        mark(id);
        return id;
      };
      auto seed_centre = [&]() {
        unsigned m = key[0];
#pragma unroll
        for (int i = 1; i < kCuMaxTiles / 32; ++i) m = min(m, key[i]);
        return pick(m);
      };
      auto seed_probe = [&](int p) {
        int L = p == 0 ? 0 : p == 1 ? 31 : (int)(__brev((unsigned)(p - 1)) >> 27);
        L = min(L, max(n_live - 1, 0));
        const float px = __shfl_sync(0xffffffffu, qx, L), py = __shfl_sync(0xffffffffu, qy, L),
                    pz = __shfl_sync(0xffffffffu, qz, L);
        unsigned m = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < kCuMaxTiles / 32; ++i) {
          const int t = lane + 32 * i;
          if (key[i] != 0xffffffffu) {  // (unused, hence < nt)
            const float4 bl = sbox[2 * t], bh = sbox[2 * t + 1];
            const float fx = fmaxf(fabsf(px - bl.x), fabsf(px - bh.x));
            const float fy = fmaxf(fabsf(py - bl.y), fabsf(py - bh.y));
            const float fz = fmaxf(fabsf(pz - bl.z), fabsf(pz - bh.z));
            m = min(m, (__float_as_uint(fx * fx + fy * fy + fz * fz) & 0xffffff80u) | (unsigned)t);
          }
        }
        return pick(m);
      };
      int seed_no = 0;
      auto next_seed = [&]() {  // this warp's next tile of the common ranking (solo: every warp walks it, keeps every tw-th)
        int t = -1;
        for (int o = 0; o < tw; ++o) {
          const int u = seed_no++;
          const int id = (probes && (u & 1)) ? seed_probe(u >> 1) : seed_centre();
          if (o == tr) t = id;
        }
        return t;
      };
      float mn[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) mn[i] = kInf;
      int t_seed = next_seed();
      if (t_seed >= 0) fetch_tile(0, t_seed);
      const int my_seeds = max(1, n_seeds / tw);
      for (int u = 0; u < my_seeds && t_seed >= 0; ++u) {
        const int buf = u & 1;
        tile_landed();
        t_seed = u + 1 < my_seeds ? next_seed() : -1;
        if (t_seed >= 0) fetch_tile(buf ^ 1, t_seed);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float4 p = sm.stage[buf][c];
          mn[c] = fminf(mn[c], sqdist(qx, qy, qz, p.x, p.y, p.z));
        }
      }
      if (kG > 1 && solo) {  // the warps' seed tiles hold different references: the minimum over the warps is the class minimum
#pragma unroll
        for (int i = 0; i < 32; ++i) sm.minima[i][lane] = mn[i];
        __syncthreads();
#pragma unroll
        for (int o = 1; o < kG; ++o) {
          const CuTaskSmem& other = tsm[(w + o) % kG];
#pragma unroll
          for (int i = 0; i < 32; ++i) mn[i] = fminf(mn[i], other.minima[i][lane]);
        }
        __syncthreads();  // minima[] aliases cand[]
      }
      sort32(mn);
      tau = mn[0];
#pragma unroll
      for (int i = 1; i < kK; ++i)
        if (i < k) tau = mn[i];
      __syncwarp();  // the stage buffers are reused by the sweep
    }
    T = inflate(tau);

    // ---- the tiles any query of the warp needs: box-to-box prefilter (a lane per tile), then the exact per-query test ----
    {
      const float tmax = __uint_as_float(redux_max_u32(live ? __float_as_uint(T) : 0u));  // T >= +0
      for (int i = 0; i * 32 * tw < nt; ++i) {
        const int t = (lane + 32 * i) * tw + tr;
        bool maybe = false;
        if (t < nt) {
          const float4 bl = sbox[2 * t], bh = sbox[2 * t + 1];
          const float gx = fmaxf(0.f, fmaxf(bl.x - ch[0], cl[0] - bh.x));
          const float gy = fmaxf(0.f, fmaxf(bl.y - ch[1], cl[1] - bh.y));
          const float gz = fmaxf(0.f, fmaxf(bl.z - ch[2], cl[2] - bh.z));
          maybe = (gx * gx + gy * gy + gz * gz) * 0.99999619f <= tmax;
        }
        unsigned mask = __ballot_sync(0xffffffffu, maybe);
        while (mask) {
          const int t2 = (32 * i + __ffs(mask) - 1) * tw + tr;
          mask &= mask - 1u;
          const float4 bl = sbox[2 * t2], bh = sbox[2 * t2 + 1];  // warp-uniform: broadcast
          const float gx = fmaxf(0.f, fmaxf(bl.x - qx, qx - bh.x));
          const float gy = fmaxf(0.f, fmaxf(bl.y - qy, qy - bh.y));
          const float gz = fmaxf(0.f, fmaxf(bl.z - qz, qz - bh.z));
          // (1 - 2^-18): conservative against the rounding of both sides
          if (__any_sync(0xffffffffu, live && (gx * gx + gy * gy + gz * gz) * 0.99999619f <= T)) {
            if (lane == 0) sm.tlist[n_list] = (unsigned char)t2;
            ++n_list;
          }
        }
      }
    }
    if (kG > 1) {  // publish the bounds and open the deque (after the list: release order for the helpers)
      sm.tpub[lane] = T;
      __syncwarp();
      __threadfence_block();
      if (lane == 0) atomicExch(&sm.deque, (unsigned)(n_list + kCuDequeBias) << 16 | (unsigned)kCuDequeBias);
    }
    __syncwarp();
  }

  // ---- sweep: role 0 = the warp's own task (tiles from the front of its deque), role 1 = the task it helps (from the back)
  u64 tk[kK];  // ascending (distance bits, original index) keys of the best references seen so far
  int count = 0;
  bool folded = false;  // warp-uniform: a fold has happened (lanes that hold k references have an exact kth)
  u64 kth = ~0ull;
  int victim = -1;
  for (int role = 0; role < (kG > 1 ? (solo ? 1 + 2 * kG : 2) : 1); ++role) {  // (solo: a warp may help several times)
    CuTaskSmem* task = &sm;  // whose tile list
    if (role == 0) {
      if (!has_task) continue;
    } else {
      // the task of this CTA with the most tiles left (at least 2); announce the help BEFORE the first claim
      victim = -1;
      int best = 1;
#pragma unroll
      for (int o = 1; o < kG; ++o) {
        const int v = (w + o) % kG;
        const unsigned dq = *reinterpret_cast<volatile unsigned*>(&tsm[v].deque);
        const int left = (int)(dq >> 16) - (int)(dq & 0xffffu);
        if (left > best) {
          best = left;
          victim = v;
        }
      }
      if (victim < 0) break;
      task = &tsm[victim];
      if (!solo) {
        if (lane == 0) atomicOr(&task->helped_by, 1u << w);
        __threadfence_block();
        __syncwarp();
        const float4 vq = task->qv[lane];
        qx = vq.x, qy = vq.y, qz = vq.z;
        qi = __float_as_int(vq.w);
        live = qi >= 0;
        T = *reinterpret_cast<volatile float*>(&task->tpub[lane]);  // (the owner may have tightened it since: any is valid)
      }
    }
    if (!(solo && role >= 1)) {  // (solo: the same queries - the warp goes on with the top-k and the bound it has)
#pragma unroll
      for (int j = 0; j < kK; ++j) tk[j] = ~0ull;
      count = 0;
      folded = false;
      kth = ~0ull;
    }
    // fold: kCuBatch candidates at a time are re-evaluated exactly (their loads in flight together), sorted by a register
    // network and merged into the sorted top-k by one bitonic split + merge: ~45 instructions per candidate where an
    // insertion costs ~100
    auto fold = [&]() {
      const int n_max = redux_max_s32(count);
      for (int c0 = 0; c0 < n_max; c0 += kCuBatch) {
        u64 nk[kCuBatch];
#pragma unroll
        for (int j = 0; j < kCuBatch; ++j) {
          nk[j] = ~0ull;
          if (c0 + j < count) {
            const float4 cp = __ldg(Rb + sm.cand[c0 + j][lane]);
            nk[j] = ((u64)__float_as_uint(sqdist(qx, qy, qz, cp.x, cp.y, cp.z)) << 32) | (unsigned)__float_as_int(cp.w);
          }
        }
        bitonic_sort_u64<kCuBatch>(nk);
#pragma unroll
        for (int j = 0; j < kK; ++j) {  // the kK smallest of (tk ascending, nk descending): a bitonic sequence
          const int rj = kK - 1 - j;
          if (rj < kCuBatch) tk[j] = nk[rj] < tk[j] ? nk[rj] : tk[j];
        }
        bitonic_merge_u64<kK>(tk);
      }
      count = 0;
      kth = tk[0];
#pragma unroll
      for (int j = 1; j < kK; ++j)
        if (j < k) kth = tk[j];
      if (kth != ~0ull) T = fminf(T, inflate(__uint_as_float((unsigned)(kth >> 32))));
      if (kG > 1 && role == 0) sm.tpub[lane] = T;  // (helpers read it once, when they start)
    };
    // next tile of the task: the owner claims from the front, a helper from the back; -1 = the deque is empty
    int next_own = 0;
    auto claim = [&]() {
      if (kG == 1) return next_own < n_list ? (int)sm.tlist[next_own++] : -1;  // (no deque)
      unsigned old = 0u;
      if (lane == 0) old = role == 0 ? atomicAdd(&task->deque, 1u) : atomicSub(&task->deque, 1u << 16);
      old = __shfl_sync(0xffffffffu, old, 0);
      const int head = (int)(old & 0xffffu) - kCuDequeBias, tail = (int)(old >> 16) - kCuDequeBias;
      if (head >= tail) return -1;
      return (int)task->tlist[role == 0 ? head : tail - 1];
    };
    int buf = 0;
    int t = claim();
    if (t >= 0) fetch_tile(0, t);
    while (t >= 0) {
      const int t_next = claim();  // (claimed one tile ahead: its copy travels during this tile's sweep)
      tile_landed();
      if (t_next >= 0) fetch_tile(buf ^ 1, t_next);
      bool skip = false;
      if (folded) {
        // bounds have moved since the list was made (a query that holds k references knows its exact k-th (distance,
        // index)): a tile is skipped when no query can still gain from it - its box is farther than the query's bound, or
        // it is ONE point (duplicates: zero padding) that ties or loses even with its smallest index
        const float4 bl = sbox[2 * t], bh = sbox[2 * t + 1];
        const float gx = fmaxf(0.f, fmaxf(bl.x - qx, qx - bh.x));
        const float gy = fmaxf(0.f, fmaxf(bl.y - qy, qy - bh.y));
        const float gz = fmaxf(0.f, fmaxf(bl.z - qz, qz - bh.z));
        bool need = (gx * gx + gy * gy + gz * gz) * 0.99999619f <= T;  // (T >= the k-th distance, inflated)
        if (__float_as_int(bh.w) == 1 && kth != ~0ull)
          need = (((u64)__float_as_uint(sqdist(qx, qy, qz, bl.x, bl.y, bl.z)) << 32) | (unsigned)__float_as_int(bl.w)) < kth;
        skip = !__any_sync(0xffffffffu, live && need);
      }
      if (!skip) {
        unsigned rej = 0u;  // bit (31 - c): reference c of the tile is farther than T
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float4 p = sm.stage[buf][c];
          const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
          const float f = __fmaf_rn(-dz, dz, __fmaf_rn(-dx, dx, __fmaf_rn(-dy, dy, T)));
          rej = __funnelshift_l(__float_as_uint(f), rej, 1);
        }
        unsigned m = live ? __brev(~rej) : 0u;  // bit c: reference c is a candidate
        if (t == nt - 1 && (R & 31)) m &= (1u << (R & 31)) - 1u;  // padding sentinels (only reachable with a non-finite bound)
        while (m) {
          const int c = __ffs(m) - 1;
          m &= m - 1u;
          if (kth != ~0ull) {  // exact test against the current k-th (distance, index): ties at the bound do not flood the list
            const float4 p = sm.stage[buf][c];
            const u64 ck = ((u64)__float_as_uint(sqdist(qx, qy, qz, p.x, p.y, p.z)) << 32) | (unsigned)__float_as_int(p.w);
            if (ck >= kth) continue;
          }
          sm.cand[count][lane] = (unsigned short)(t * 32 + c);  // count < kCuCap: a list is folded before a tile could overfill it
          ++count;
        }
      }
      // the one fold site: a list with fewer than 32 free slots, and the end of the task
      if (t_next < 0 || __any_sync(0xffffffffu, count > kCuCap - 32)) {
        fold();
        folded = true;
      }
      t = t_next;
      buf ^= 1;
    }
    if (solo) continue;  // (then helps once; the lists are merged below)
    if (role == 1) {  // leave the top-k for the owner (own cand[] is done with)
#pragma unroll
      for (int j = 0; j < kK; ++j) sm.keys[j][lane] = tk[j];
      break;
    }
    // own task swept.  Nobody can take a tile any more; whoever did has set its bit before its first claim.
    if (kG > 1) {
      __threadfence_block();
      const unsigned helped = *reinterpret_cast<volatile unsigned*>(&sm.helped_by);
      if (helped != 0u) break;  // the helpers' lists are merged after the barrier
    }
    if (live) {
      const size_t o = ((size_t)b * Q + qi) * k;
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        if (j < k) {
          // fewer than k entries only if distances were NaN (outside the contract): index 0 keeps consumers in range
          idx_out[o + j] = tk[j] == ~0ull ? 0 : (int64_t)(unsigned)tk[j];
          if (dist_out) dist_out[o + j] = tk[j] == ~0ull ? kInf : __uint_as_float((unsigned)(tk[j] >> 32));
        }
      }
    }
  }
  if (kG == 1) return;
  if (solo) {  // every warp's sorted list to warp 0
    if (w != 0) {
#pragma unroll
      for (int j = 0; j < kK; ++j) sm.keys[j][lane] = tk[j];
    }
    __syncthreads();
    if (w != 0) return;
  } else {
    __syncthreads();
  }
  const unsigned helped = solo ? ((1u << kG) - 2u) : has_task ? sm.helped_by : 0u;
  if (helped == 0u) return;
  // (tk still holds the own task's top-k: a helped owner left the loop before anything overwrote it)
  for (int h = 0; h < kG; ++h) {
    if (!((helped >> h) & 1u)) continue;
#pragma unroll
    for (int j = 0; j < kK; ++j) {
      const u64 ok = tsm[h].keys[kK - 1 - j][lane];
      tk[j] = ok < tk[j] ? ok : tk[j];
    }
    bitonic_merge_u64<kK>(tk);
  }
  const int own_qi = __float_as_int(qv.w);
  if (own_qi < 0) return;
  const size_t o = ((size_t)b * Q + own_qi) * k;
#pragma unroll
  for (int j = 0; j < kK; ++j) {
    if (j < k) {
      idx_out[o + j] = tk[j] == ~0ull ? 0 : (int64_t)(unsigned)tk[j];
      if (dist_out) dist_out[o + j] = tk[j] == ~0ull ? kInf : __uint_as_float((unsigned)(tk[j] >> 32));
    }
  }
}

// Reference-range splits for the list kernel: only when the query count cannot fill the machine on its own.
static int knn_r_splits(int B, int R, int Q) {
  const long long ctas4 = (long long)B * ((Q + 31) / 32);
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

// Which kernel a call takes, shared by the workspace query and the launcher.
struct KnnPlan {
  bool select;  // thread-per-query selection (k <= 16) vs warp-per-query list
  int splits;   // select: warps per 32 queries; list: reference slices merged by knn_merge_kernel
  bool pruned;  // Hilbert-binned, box-culled selection (k <= 16, kPrMinR <= R <= kPrMaxN, Q <= kPrMaxN)
};

static long long env_ll(const char* name, long long fallback) {
  const char* e = std::getenv(name);
  return (e && std::atoll(e) > 0) ? std::atoll(e) : fallback;
}

static KnnPlan knn_plan(int B, int R, int Q, int k) {
  static const bool force_list = env_ll("IAPT_KNN_FORCE_LIST", 0) == 1;             // development knobs
  static const long long min_warps = env_ll("IAPT_KNN_SELECT_MIN_WARPS", 256);
  static const long long want_warps = env_ll("IAPT_KNN_SELECT_WANT_WARPS", 2048);
  static const long long force_split = env_ll("IAPT_KNN_SELECT_SPLIT", 0);
  // Culled path: measured against the select kernel at B = 48, k = 16 (profiles/r02y_knn_culled.txt): 2048 x 2048 0.10 -
  // 0.13 ms vs 0.18, 4096 x 4096 0.20 vs 0.52, 1024 x 1024 0.058 vs 0.072; with 512 queries per cloud (768 warps of
  // queries: a quarter of the machine) the select kernel's slices win, 0.064 vs 0.072.  IAPT_KNN_PRUNED = 1 forces it
  // wherever it applies, 2 disables it (read per call: the tests run both paths over the same cases).
  const long long pruned_mode = env_ll("IAPT_KNN_PRUNED", 0);
  const long long warps = (long long)B * ((Q + 31) / 32);
  const bool pruned_fits = k <= kSelMaxK && R >= kPrMinR && R <= kPrMaxN && Q >= 1 && Q <= kPrMaxN && B <= 65535;
  // (measured over B = 1 ... 48, profiles/r02y_knn_culled.txt: the select kernel's time grows with R, the culled path has
  // a fixed part of ~35 us - bin kernel + one task's latency - so at 1024 references it needs 1536 warps of queries to
  // win, at 2048 references 640, above that 128.  SPARSE queries - the grouper's 512 FPS centres against 2048 points -
  // make every warp's query box, hence its tile list, 2-3x larger: there the select kernel's slices stay ahead.)
  const bool dense = 2LL * Q >= R;
  const bool pruned_pays = dense && warps >= (R > 2048 ? 128 : R >= 2048 ? 640 : 1536);
  if (!force_list && pruned_fits && pruned_mode != 2 && (pruned_mode == 1 || pruned_pays)) return {false, 1, true};
  if (k <= kSelMaxK && R >= 64 && R <= 65536 && !force_list && warps >= min_warps) {
    int s = 1;
    // (slices shorter than 256 do not pay; 512 before every slice's warp folded its own list - r02y-9: 48 x 512 x 512
    // 0.0273 -> 0.0232 ms with two slices of 256)
    while (s < 4 && warps * s < want_warps && R / (2 * s) >= 256) s *= 2;
    if (force_split == 1 || force_split == 2 || force_split == 4) s = (int)force_split;
    return {true, s, false};
  }
  return {false, knn_r_splits(B, R, Q), false};
}

size_t knn_workspace_bytes(int B, int R, int Q, int k) {
  const KnnPlan plan = knn_plan(B, R, Q, k);
  if (plan.pruned) {  // sorted references + their tile boxes + sorted queries
    const size_t Rp = (size_t)(R + 31) / 32 * 32, Qp = (size_t)(Q + 31) / 32 * 32;
    return (size_t)B * (Rp * sizeof(float4) + Rp / 32 * sizeof(PrTileBox) + Qp * sizeof(float4)) + 256;
  }
  if (plan.select || plan.splits <= 1) return 0;
  return (sizeof(float) + sizeof(int)) * (size_t)B * Q * plan.splits * k;  // per-slice lists for knn_merge_kernel
}

template <int kSlots>
static void knn_list_launch(const float* ref, const float* qry, int B, int R, int Q, int k, int64_t* idx, float* dist,
                            void* split_ws, int rs, cudaStream_t stream) {
  const int q_blocks = (Q + kKnnWarps - 1) / kKnnWarps;
  if (rs > 1 && split_ws != nullptr) {
    dim3 grid((unsigned)((long long)B * q_blocks), rs);
    float* pd = static_cast<float*>(split_ws);
    int* pi = reinterpret_cast<int*>(pd + (size_t)B * Q * rs * k);
    knn_kernel<kSlots><<<grid, kKnnThreads, 0, stream>>>(ref, qry, R, Q, k, q_blocks, idx, dist, rs, pd, pi);
    const long long nq = (long long)B * Q;
    knn_merge_kernel<kSlots><<<(unsigned)((nq + kKnnWarps - 1) / kKnnWarps), kKnnThreads, 0, stream>>>(pd, pi, nq, k, rs,
                                                                                                     idx, dist);
    count_launch(2);
  } else {
    dim3 grid((unsigned)((long long)B * q_blocks), 1);
    knn_kernel<kSlots><<<grid, kKnnThreads, 0, stream>>>(ref, qry, R, Q, k, q_blocks, idx, dist, 1, nullptr, nullptr);
    count_launch();
  }
}

template <int kStride, int kK, int kSplit>
static void knn_select_launch(const float* ref, const float* qry, int B, int R, int Q, int k, int64_t* idx, float* dist,
                              cudaStream_t stream) {
  const int qb = (Q + 31) / 32;
  const int aligned = (reinterpret_cast<uintptr_t>(ref) % 16 == 0 && R % 4 == 0) ? 1 : 0;
  knn_select_kernel<kStride, kK, kSplit><<<(unsigned)((long long)B * qb), 32 * kSplit, 0, stream>>>(
      ref, qry, R, Q, k, qb, aligned, idx, dist);
  count_launch();
}

template <int kK>
static void knn_select_dispatch(const float* ref, const float* qry, int B, int R, int Q, int k, int splits, int64_t* idx,
                                float* dist, cudaStream_t stream) {
  static const bool sample_slices = env_ll("IAPT_KNN_SELECT_SAMPLE_SLICES", 1) == 1;  // (2 = off)
  if (splits == 4 && sample_slices && R >= 2048)
    knn_select_launch<2, kK, 4>(ref, qry, B, R, Q, k, idx, dist, stream);
  else if (splits == 4)
    knn_select_launch<1, kK, 4>(ref, qry, B, R, Q, k, idx, dist, stream);
  else if (splits == 2 && sample_slices && R >= 1024)
    knn_select_launch<2, kK, 2>(ref, qry, B, R, Q, k, idx, dist, stream);
  else if (splits == 2)
    knn_select_launch<1, kK, 2>(ref, qry, B, R, Q, k, idx, dist, stream);
  else if (R >= 1024)  // sampled sweep 1 only where the slice is long enough to pay for the extra candidates
    knn_select_launch<2, kK, 1>(ref, qry, B, R, Q, k, idx, dist, stream);
  else
    knn_select_launch<1, kK, 1>(ref, qry, B, R, Q, k, idx, dist, stream);
}

int knn_launch(const float* ref, const float* qry, int B, int R, int Q, int k, int64_t* idx, float* dist,
               void* workspace, cudaStream_t stream) {
  const KnnPlan plan = knn_plan(B, R, Q, k);
  if (plan.pruned && workspace != nullptr) {
    const int Rp = (R + 31) / 32 * 32, Qp = (Q + 31) / 32 * 32;
    unsigned char* w = static_cast<unsigned char*>(workspace);
    w += (256 - reinterpret_cast<uintptr_t>(w) % 256) % 256;
    float4* sref = reinterpret_cast<float4*>(w);
    PrTileBox* boxes = reinterpret_cast<PrTileBox*>(sref + (size_t)B * Rp);
    float4* sqry = reinterpret_cast<float4*>(boxes + (size_t)B * (Rp / 32));
    const bool self = ref == qry && R == Q;  // the self-query of DGCNN_Grouper: one sort serves both sides
    knn_bin_kernel<<<dim3(B, self ? 1 : 2), kBinThreads, 0, stream>>>(ref, R, Rp, sref, boxes, qry, Q, Qp, sqry);
    const int qb = Qp / 32;
    const int nt = Rp / 32;
    // development knobs (read per call): 12 seed tiles, every second one a probe; 4 tasks per CTA helping each other
    int n_seeds = (int)env_ll("IAPT_KNN_SEEDS", 12);
    int group = (int)env_ll("IAPT_KNN_CULL_GROUP", 4);
    if (group != 1 && group != 2) group = 4;
    const int probes = env_ll("IAPT_KNN_PROBES", 1) == 1 ? 1 : 0;  // (2 = off)
    n_seeds = std::max(1, std::min(n_seeds, std::min(nt, 32)));
    const size_t dyn = (size_t)2 * nt * sizeof(float4);
    const float4* sq = self ? sref : sqry;
    // few tasks: one per CTA and its 4 warps share the tile list from the start
    const long long solo_below = env_ll("IAPT_KNN_CULL_SOLO_BELOW", 900);  // (888 CTAs are resident at once)
    const int solo = (group == 4 && (long long)B * qb < solo_below) ? 1 : 0;
    const unsigned grid = (unsigned)((long long)B * (solo ? qb : (qb + group - 1) / group));
#define KNN_CULLED(KK, G)                                                                                               \
  knn_culled_kernel<KK, G><<<grid, 32 * G, dyn, stream>>>(sref, boxes, sq, R, Rp, Q, Qp, k, qb, n_seeds, probes, solo, idx, \
                                                          dist)
    if (k <= 8) {
      if (group == 1) KNN_CULLED(8, 1); else if (group == 2) KNN_CULLED(8, 2); else KNN_CULLED(8, 4);
    } else {
      if (group == 1) KNN_CULLED(16, 1); else if (group == 2) KNN_CULLED(16, 2); else KNN_CULLED(16, 4);
    }
#undef KNN_CULLED
    count_launch(2);
    return (int)cudaGetLastError();
  }
  if (plan.select || plan.pruned) {
    if (k <= 8)
      knn_select_dispatch<8>(ref, qry, B, R, Q, k, plan.splits, idx, dist, stream);
    else
      knn_select_dispatch<16>(ref, qry, B, R, Q, k, plan.splits, idx, dist, stream);
    return (int)cudaGetLastError();
  }
  const int rs = workspace == nullptr ? 1 : plan.splits;
  if (k <= 32)
    knn_list_launch<1>(ref, qry, B, R, Q, k, idx, dist, workspace, rs, stream);
  else
    knn_list_launch<2>(ref, qry, B, R, Q, k, idx, dist, workspace, rs, stream);
  return (int)cudaGetLastError();
}

// ---- square_distance (models/Transformer_utils.py:32-51) ---------------------------------------------------------------
// The reference's API: the full [B,N,M] matrix in the EXPANSION form it uses, in its op order
//     dist = -2 * (src @ dst^T);  dist += sum(src^2, -1)[:, :, None];  dist += sum(dst^2, -1)[:, None, :]
// (knn_point above never materialises it; this serves callers that want the matrix itself).  HBM-write bound:
// a CTA covers 8 rows x 1024 columns, a thread 4 consecutive columns of every row -> 128-bit coalesced stores.
constexpr int kSqRows = 8, kSqThreads = 256, kSqMaxC = 16;

__global__ void __launch_bounds__(kSqThreads) square_distance_kernel(const float* __restrict__ src,
                                                                     const float* __restrict__ dst, int N, int M, int C,
                                                                     float* __restrict__ out) {
  __shared__ float s_src[kSqRows][kSqMaxC + 1];
  const int b = blockIdx.z;
  const int n0 = blockIdx.y * kSqRows;
  const int m0 = (blockIdx.x * kSqThreads + threadIdx.x) * 4;
  for (int i = threadIdx.x; i < kSqRows * C; i += kSqThreads) {
    const int r = i / C, c = i - r * C;
    s_src[r][c] = (n0 + r < N) ? __ldg(src + ((size_t)b * N + n0 + r) * C + c) : 0.f;
  }
  __syncthreads();
  if (m0 >= M) return;
  float d[4][kSqMaxC];
  float dsum[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    dsum[j] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = (m0 + j < M) ? __ldg(dst + ((size_t)b * M + m0 + j) * C + c) : 0.f;
      d[j][c] = v;
      dsum[j] = c == 0 ? __fmul_rn(v, v) : __fadd_rn(dsum[j], __fmul_rn(v, v));
    }
  }
  const bool vec = (M & 3) == 0;
  for (int r = 0; r < kSqRows; ++r) {
    const int n = n0 + r;
    if (n >= N) break;
    float ssum = 0.f;
    for (int c = 0; c < C; ++c) ssum = c == 0 ? __fmul_rn(s_src[r][c], s_src[r][c]) : __fadd_rn(ssum, __fmul_rn(s_src[r][c], s_src[r][c]));
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = 0.f;
      for (int c = 0; c < C; ++c) t = c == 0 ? __fmul_rn(s_src[r][c], d[j][c]) : __fmaf_rn(s_src[r][c], d[j][c], t);
      o[j] = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, t), ssum), dsum[j]);
    }
    float* row = out + ((size_t)b * N + n) * M + m0;
    if (vec) {
      *reinterpret_cast<float4*>(row) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (m0 + j < M) row[j] = o[j];
    }
  }
}

int square_distance_launch(const float* src, const float* dst, int B, int N, int M, int C, float* out,
                           cudaStream_t stream) {
  if (C > kSqMaxC) return -3;  // IAPT_EUNSUPPORTED
  const dim3 grid((M + kSqThreads * 4 - 1) / (kSqThreads * 4), (N + kSqRows - 1) / kSqRows, B);
  square_distance_kernel<<<grid, kSqThreads, 0, stream>>>(src, dst, N, M, C, out);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
