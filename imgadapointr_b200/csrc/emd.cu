// emd.cu - earth mover's distance approximation by auction, for sm_100a.
//
// Replaces extensions/emd (emd_cuda.cu:23-308, driven by emd_module.py:31-79), the test-time EMDistance metric
// (utils/metrics.py:111-116: eps = 0.005, 100 iterations, 7 launches each).  Same algorithm, same arithmetic:
//   Bid    every unassigned prediction j scans all targets k with value
//              v = (float)(3.0 - (double)sqrtf(d2) - (double)price[k])          (`3.0` is a double literal upstream)
//          keeps the best and second best, bids best_i with increment best - better + eps and raises max_inc[best_i];
//   Pick   a bidder whose increment is within 1e-6 (double compare) of max_inc[bid] wins the target;
//   Assign the winner takes it (the previous owner becomes unassigned), price += increment, max_inc = -1e9; in the last
//          iteration every unassigned prediction simply takes its bid (the assignment need not be a bijection).
// The reference leaves two choices to thread scheduling (and races on them); they are fixed here and in the oracle:
// equal best values keep the LOWEST k, and of several bidders within 1e-6 of the maximum the HIGHEST j wins.
//
// B200 shape of the hot loop (Bid is the only O(n^2) step): the unassigned list is compacted per cloud; while it is long a
// CTA of 8 warps takes 32 list entries (4 per warp, so every target fetched from shared memory feeds 4 predictions and 4
// independent sqrt / double chains), targets stream through a shared-memory tile as float4 (x, y, z, price), lanes
// stride over the targets and the per-lane (best, second, index) triples are folded with shuffles.  As the list shrinks
// the same kernel switches to 8 entries per CTA and finally to one entry per CTA (all 256 threads on one prediction), so
// the ~90 late iterations with a handful of unassigned points are bound by launch latency, not by one warp's scan.
#include <cstdint>

#include "common.cuh"

namespace iapt {

constexpr int kEmdThreads = 256;
constexpr int kEmdWarps = kEmdThreads / 32;
constexpr int kEmdTile = 2048;                       // targets per shared-memory tile (32 KB)

struct EmdState {
  int* assignment;      // [B,n] output
  int* assignment_inv;  // [B,n]
  int* bid;             // [B,n]
  int* winner;          // [B,n]
  int* unass;           // [B,n] list of the unassigned predictions (any order: no result depends on it)
  int* unass_next;      // [B,n] the list being built for the next iteration
  int* cand;            // [B,2n] scratch of emd_resolve_kernel
  int* unass_cnt;       // [B]
  float* price;         // [B,n]
  float* inc;           // [B,n]
  float* max_inc;       // [B,n]
  // target-sliced bidding (short lists): per cloud and list entry, folded across the slices with atomics
  u64* top;             // [B, cap]  (orderable best value << 32) | (0xFFFFFFFF - target index)
  unsigned* second;     // [B, cap]  orderable second-best value
  int slices, cap;
};

// float <-> unsigned whose integer order is the float order (values here are finite and never -0)
__device__ __forceinline__ unsigned emd_ord(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float emd_unord(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ u64 emd_top_init() { return (u64)emd_ord(-1e9f) << 32; }  // index field 0 = target -1

__global__ void __launch_bounds__(256) emd_init_kernel(EmdState s, int n, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    s.assignment[t] = -1;
    s.assignment_inv[t] = -1;
    s.price[t] = 0.f;
    s.max_inc[t] = 0.f;  // torch.zeros upstream (emd_module.py:52); -1e9 only after a target has been taken
    s.winner[t] = -1;
    s.unass[t] = (int)(t % n);  // everybody is unassigned before the first iteration
    if (t % n == 0) s.unass_cnt[t / n] = n;
    if (t % n < s.cap) {
      s.top[(t / n) * s.cap + t % n] = emd_top_init();
      s.second[(t / n) * s.cap + t % n] = emd_ord(-1e9f);
    }
  }
}

// Upstream's `if (v > best) {better = best; best = v; best_i = k;} else if (v > better) better = v;` without the branches
// (v <= best: min(best, v) = v; v > best: min(best, v) = best >= better) - identical results for non-NaN values.
__device__ __forceinline__ void emd_offer(float v, int k, float& best, float& better, int& best_i) {
  better = fmaxf(better, fminf(best, v));
  best_i = v > best ? k : best_i;
  best = fmaxf(best, v);
}

__device__ __forceinline__ float atomic_max_float(float* addr, float val) {  // as emd_cuda.cu:10-21
  int ret = __float_as_int(*addr);
  while (val > __int_as_float(ret)) {
    const int old = ret;
    if ((ret = atomicCAS(reinterpret_cast<int*>(addr), old, __float_as_int(val))) == old) break;
  }
  return __int_as_float(ret);
}

// Fold (best, second best, index) over the lanes of a warp: the larger best wins (lower index on equal values), the
// second best is the best of the rest.  Every lane ends with the folded triple.
__device__ __forceinline__ void emd_warp_fold(float& best, float& better, int& best_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const float o2 = __shfl_xor_sync(0xffffffffu, better, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    const bool take = ob > best || (ob == best && (unsigned)oi < (unsigned)best_i);
    const float loser = take ? best : ob;
    better = fmaxf(fmaxf(better, o2), loser);
    if (take) {
      best = ob;
      best_i = oi;
    }
  }
}

__device__ __forceinline__ void emd_publish(const EmdState& s, size_t o, int j, float best, float better, int best_i,
                                            float eps) {
  const float increment = best - better + eps;
  s.bid[o + j] = best_i;
  s.inc[o + j] = increment;
  atomic_max_float(s.max_inc + o + best_i, increment);
}

// kQP unassigned points per warp, every warp scans all targets.
template <int kQP>
__device__ __forceinline__ void emd_bid_points_per_warp(const EmdState& s, const float* __restrict__ X1,
                                                        const float* __restrict__ X2, const float* __restrict__ price,
                                                        const int* __restrict__ list, int cnt, int first, int n, float eps,
                                                        size_t o, float4* tile) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int j[kQP];
  float ax[kQP], ay[kQP], az[kQP], best[kQP], better[kQP];
  int best_i[kQP];
#pragma unroll
  for (int u = 0; u < kQP; ++u) {
    const int e = first + warp * kQP + u;
    j[u] = e < cnt ? list[e] : -1;
    const int jj = j[u] >= 0 ? j[u] : 0;
    ax[u] = __ldg(X1 + 3 * (size_t)jj);
    ay[u] = __ldg(X1 + 3 * (size_t)jj + 1);
    az[u] = __ldg(X1 + 3 * (size_t)jj + 2);
    best[u] = -1e9;
    better[u] = -1e9;
    best_i[u] = -1;
  }
  for (int t0 = 0; t0 < n; t0 += kEmdTile) {
    const int tn = min(kEmdTile, n - t0);
    __syncthreads();
    for (int r = threadIdx.x; r < tn; r += kEmdThreads) {
      const float* p = X2 + 3 * (size_t)(t0 + r);
      tile[r] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(price + t0 + r));
    }
    __syncthreads();
    for (int r = lane; r < tn; r += 32) {  // ascending k inside a lane: strict '>' keeps the lane's lowest index
      const float4 q = tile[r];
#pragma unroll
      for (int u = 0; u < kQP; ++u) {
        const float d2 = sqdist(ax[u], ay[u], az[u], q.x, q.y, q.z);
        const float v = (float)(3.0 - (double)sqrtf(d2) - (double)q.w);
        emd_offer(v, t0 + r, best[u], better[u], best_i[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kQP; ++u) {
    emd_warp_fold(best[u], better[u], best_i[u]);
    if (lane == 0 && j[u] >= 0) emd_publish(s, o, j[u], best[u], better[u], best_i[u], eps);
  }
}

// Short lists (late iterations): the TARGETS are sliced over the CTAs instead of the predictions.  A CTA stages its
// slice (n / slices targets) in shared memory once and every thread runs its list entries against it (warp-uniform
// LDS).  The per-slice (best, second best, index) of an entry are folded across slices with two atomics:
//   old = atomicMax(top, (best << 32) | ~index)    - the larger value wins, the lower index on equal values;
//   atomicMax(second, max(second best of the slice, the loser of that exchange))
// Every best that is not the final top loses exactly one exchange (on arrival or when displaced later), so `second` ends
// as the largest of all other values: the same (best, second best, index) as one sequential scan.  All SMs work on a
// 100-entry list, nobody streams the whole cloud through one SM, and emd_resolve_kernel reads two words per entry.
__device__ __forceinline__ void emd_bid_sliced(const EmdState& s, const float* __restrict__ X1,
                                               const float* __restrict__ X2, const float* __restrict__ price,
                                               const int* __restrict__ list, int cnt, int n, int b, float4* tile) {
  const int slice = blockIdx.x;
  const int per = (n + s.slices - 1) / s.slices;  // <= kEmdTile by construction
  const int k_lo = min(n, slice * per), k_hi = min(n, k_lo + per);
  if (k_lo >= k_hi) return;
  for (int r = threadIdx.x; r < k_hi - k_lo; r += kEmdThreads) {
    const float* p = X2 + 3 * (size_t)(k_lo + r);
    tile[r] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(price + k_lo + r));
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cnt; e += kEmdThreads) {
    const int j = list[e];
    const float ax = __ldg(X1 + 3 * (size_t)j), ay = __ldg(X1 + 3 * (size_t)j + 1), az = __ldg(X1 + 3 * (size_t)j + 2);
    float best = -1e9, better = -1e9;
    int best_i = -1;
    for (int r = 0; r < k_hi - k_lo; ++r) {  // ascending k
      const float4 q = tile[r];
      const float v = (float)(3.0 - (double)sqrtf(sqdist(ax, ay, az, q.x, q.y, q.z)) - (double)q.w);
      emd_offer(v, k_lo + r, best, better, best_i);
    }
    const u64 key = ((u64)emd_ord(best) << 32) | (u64)(0xFFFFFFFFu - (unsigned)best_i);
    const u64 old = atomicMax(s.top + (size_t)b * s.cap + e, key);
    const float loser = key > old ? emd_unord((unsigned)(old >> 32)) : best;
    atomicMax(s.second + (size_t)b * s.cap + e, emd_ord(fmaxf(loser, better)));
  }
}

// Work per CTA adapts to the length of the cloud's unassigned list (known only on the device): 32 points per CTA (4 per
// warp) while that still fills the machine, then a warp per point, and for lists of at most `cap` entries the targets
// are sliced over the CTAs instead (emd_bid_sliced).  CTAs beyond the work exit at once.
__global__ void __launch_bounds__(kEmdThreads) emd_bid_kernel(EmdState s, const float* __restrict__ xyz1,
                                                              const float* __restrict__ xyz2, int n, float eps,
                                                              int fill_ctas) {
  __shared__ float4 tile[kEmdTile];
  const int b = blockIdx.y;
  const int cnt = s.unass_cnt[b];
  if (cnt == 0) return;
  const size_t o = (size_t)b * n;
  const float* X1 = xyz1 + o * 3;
  const float* X2 = xyz2 + o * 3;
  const float* price = s.price + o;
  const int* list = s.unass + o;
  if (cnt >= fill_ctas * kEmdWarps * 4) {
    const int first = blockIdx.x * kEmdWarps * 4;
    if (first < cnt) emd_bid_points_per_warp<4>(s, X1, X2, price, list, cnt, first, n, eps, o, tile);
  } else if (cnt > s.cap) {
    const int first = blockIdx.x * kEmdWarps;
    if (first < cnt) emd_bid_points_per_warp<1>(s, X1, X2, price, list, cnt, first, n, eps, o, tile);
  } else if ((int)blockIdx.x < s.slices) {
    emd_bid_sliced(s, X1, X2, price, list, cnt, n, b, tile);
  }
}

__global__ void __launch_bounds__(1024) emd_resolve_kernel(EmdState s, int n, int last, float eps, int fill_ctas) {
  __shared__ int warp_tot[32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t o = (size_t)b * n;
  const int cnt = s.unass_cnt[b];
  if (cnt == 0) return;  // nothing was unassigned: state, list and count stay as they are
  int* asg = s.assignment + o;
  int* list = s.unass + o;
  int* win = s.winner + o;
  if (cnt <= s.cap && cnt < fill_ctas * kEmdWarps * 4) {
    // the bids of this iteration sit in the slice-folded accumulators: publish them and re-arm the accumulators
    for (int e = threadIdx.x; e < cnt; e += 1024) {
      const u64 t = __ldcg(s.top + (size_t)b * s.cap + e);
      const float better = emd_unord(__ldcg(s.second + (size_t)b * s.cap + e));
      emd_publish(s, o, list[e], emd_unord((unsigned)(t >> 32)), better, (int)(0xFFFFFFFFu - (unsigned)(t & 0xFFFFFFFFu)),
                  eps);
      s.top[(size_t)b * s.cap + e] = emd_top_init();
      s.second[(size_t)b * s.cap + e] = emd_ord(-1e9f);
    }
    __syncthreads();
  }
  if (last) {
    for (int e = threadIdx.x; e < cnt; e += 1024) {
      const int j = list[e];
      asg[j] = s.bid[o + j];  // several predictions may share a target here; prices no longer matter
    }
    return;
  }
  for (int e = threadIdx.x; e < cnt; e += 1024) {
    const int j = list[e];
    const int k = s.bid[o + j];
    const double bi = (double)s.inc[o + j], mi = (double)__ldcg(s.max_inc + o + k);
    if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMax(win + k, j);
  }
  __syncthreads();
  // assign; the next list = the entries that did not win + the previous owners the winners displaced
  int* cand = s.cand + 2 * o;
  for (int e = threadIdx.x; e < cnt; e += 1024) {
    const int j = list[e];
    const int k = s.bid[o + j];
    const bool won = win[k] == j;
    int kicked = -1;
    if (won) {
      const int prev = s.assignment_inv[o + k];
      if (prev != -1) {  // prev was assigned when the iteration started: it is not in the list
        asg[prev] = -1;
        kicked = prev;
      }
      s.assignment_inv[o + k] = j;
      asg[j] = k;
      s.price[o + k] += s.inc[o + j];
      s.max_inc[o + k] = -1e9;
    }
    cand[e] = won ? -1 : j;
    cand[cnt + e] = kicked;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cnt; e += 1024) win[s.bid[o + list[e]]] = -1;  // re-arm the targets that were bid for
  int* next = s.unass_next + o;
  const int total_c = 2 * cnt;
  const int per = (total_c + 1023) / 1024;
  const int c_lo = min(total_c, (int)threadIdx.x * per), c_hi = min(total_c, c_lo + per);
  int mine = 0;
  for (int c0 = c_lo; c0 < c_hi; c0 += 8) {  // 8 independent loads at a time
    int v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = c0 + u < c_hi ? cand[c0 + u] : -1;
#pragma unroll
    for (int u = 0; u < 8; ++u) mine += v[u] >= 0 ? 1 : 0;
  }
  int incl = mine;
#pragma unroll
  for (int sh = 1; sh < 32; sh <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, sh);
    if (lane >= sh) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int before = incl - mine;
  int total = 0;
  for (int w = 0; w < 32; ++w) {
    const int t = warp_tot[w];
    if (w < warp) before += t;
    total += t;
  }
  for (int c = c_lo; c < c_hi; ++c) {
    const int v = cand[c];
    if (v >= 0) next[before++] = v;
  }
  if (threadIdx.x == 0) s.unass_cnt[b] = total;
}

__global__ void __launch_bounds__(256) emd_dist_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                       const int* __restrict__ assignment, int n, float* __restrict__ dist,
                                                       long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / n;
    const float* a = xyz1 + 3 * t;
    const float* c = xyz2 + 3 * ((size_t)b * n + assignment[t]);
    dist[t] = sqdist(__ldg(c), __ldg(c + 1), __ldg(c + 2), __ldg(a), __ldg(a + 1), __ldg(a + 2));
  }
}

// grad_xyz1[j] = 2 * grad_dist[j] * (xyz1[j] - xyz2[assignment[j]])   (emd_cuda.cu:284-300; xyz2 gets no gradient)
__global__ void __launch_bounds__(256) emd_backward_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                           const float* __restrict__ grad_dist,
                                                           const int* __restrict__ assignment, int n,
                                                           float* __restrict__ grad, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / n;
    const float* a = xyz1 + 3 * t;
    const float* c = xyz2 + 3 * ((size_t)b * n + assignment[t]);
    const float g = __fmul_rn(__ldg(grad_dist + t), 2.f);
    grad[3 * t + 0] = __fmul_rn(g, __fsub_rn(__ldg(a), __ldg(c)));
    grad[3 * t + 1] = __fmul_rn(g, __fsub_rn(__ldg(a + 1), __ldg(c + 1)));
    grad[3 * t + 2] = __fmul_rn(g, __fsub_rn(__ldg(a + 2), __ldg(c + 2)));
  }
}

// Target slices per cloud and the longest list the sliced mode takes: about two CTAs per SM over the batch, slices of at
// most one shared-memory tile.  Depends on the current device's SM count (148 when no device can be queried).
static void emd_plan(int B, int n, int* sms_out, int* slices, int* cap) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms <= 0) {
    (void)cudaGetLastError();
    sms = 148;
  }
  int sl = (2 * sms + B - 1) / B;
  const int min_sl = (n + kEmdTile - 1) / kEmdTile;
  if (sl < min_sl) sl = min_sl;
  if (sl > n) sl = n;
  *sms_out = sms;
  *slices = sl;
  *cap = n < 2048 ? n : 2048;
}

size_t emd_workspace_bytes(int B, int n) {
  const size_t bn = (size_t)B * n;
  const int cap = n < 2048 ? n : 2048;  // as emd_plan: independent of the device
  return sizeof(int) * (7 * bn + (size_t)B) + sizeof(float) * 3 * bn + 12 * (size_t)B * cap + 128;
}

int emd_forward_launch(const float* xyz1, const float* xyz2, int B, int n, float eps, int iters, float* dist,
                       int* assignment, void* workspace, cudaStream_t stream) {
  const size_t bn = (size_t)B * n;
  EmdState s;
  int* wi = static_cast<int*>(workspace);
  s.assignment = assignment;
  s.assignment_inv = wi;
  s.bid = wi + bn;
  s.winner = wi + 2 * bn;
  s.unass = wi + 3 * bn;
  s.unass_next = wi + 4 * bn;
  s.cand = wi + 5 * bn;
  s.unass_cnt = wi + 7 * bn;
  float* wf = reinterpret_cast<float*>(wi + 7 * bn + (((size_t)B + 3) & ~(size_t)3));
  s.price = wf;
  s.inc = wf + bn;
  s.max_inc = wf + 2 * bn;
  int sms = 148;
  emd_plan(B, n, &sms, &s.slices, &s.cap);
  s.top = reinterpret_cast<u64*>((reinterpret_cast<uintptr_t>(wf + 3 * bn) + 7) & ~(uintptr_t)7);
  s.second = reinterpret_cast<unsigned*>(s.top + (size_t)B * s.cap);
  const long long total = (long long)bn;
  const unsigned g1 = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  emd_init_kernel<<<g1, 256, 0, stream>>>(s, n, total);
  count_launch();
  const int fill_ctas = sms / B > 0 ? sms / B : 1;  // CTAs per cloud that keep every SM busy
  // enough CTAs for every mode of emd_bid_kernel: a warp per point needs n/8, the sliced mode one CTA per slice
  const int need_ctas = max((n + kEmdWarps - 1) / kEmdWarps, s.slices);
  const dim3 gbid(need_ctas, B);
  for (int it = 0; it < iters; ++it) {
    emd_bid_kernel<<<gbid, kEmdThreads, 0, stream>>>(s, xyz1, xyz2, n, eps, fill_ctas);
    emd_resolve_kernel<<<B, 1024, 0, stream>>>(s, n, it == iters - 1 ? 1 : 0, eps, fill_ctas);
    count_launch(2);
    int* t = s.unass;  // the list written by this iteration is the one the next iteration reads
    s.unass = s.unass_next;
    s.unass_next = t;
  }
  emd_dist_kernel<<<g1, 256, 0, stream>>>(xyz1, xyz2, assignment, n, dist, total);
  count_launch();
  return (int)cudaGetLastError();
}

int emd_backward_launch(const float* xyz1, const float* xyz2, const float* grad_dist, const int* assignment, int B, int n,
                        float* grad_xyz1, cudaStream_t stream) {
  const long long total = (long long)B * n;
  const unsigned g1 = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  emd_backward_kernel<<<g1, 256, 0, stream>>>(xyz1, xyz2, grad_dist, assignment, n, grad_xyz1, total);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
