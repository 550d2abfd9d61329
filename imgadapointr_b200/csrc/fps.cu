// fps.cu - furthest point sampling and index gathers for sm_100a.
//
// Replaces pointnet2_ops' furthest_point_sampling_kernel / gather_points_kernel / gather_points_grad_kernel
// (third-party pointnet2_ops_lib, NOT vendored in the reference; call sites utils/misc.py:37-38,
// models/base_blocks.py:550-558).  Upstream: one block per cloud, running-min array `temp` in GLOBAL memory,
// a 9-level shared-memory tree reduction with a __syncthreads per level per round.
//
// Here: one CTA per cloud; the cloud AND the running minimum live in registers (or shared memory for big
// clouds); each of the npoint-1 dependent rounds costs ONE __syncthreads: warp argmax via two REDUX
// instructions (max of the value bits, then min index among the lanes holding it), one shared-memory slot per
// warp (double-buffered by round parity), and every warp redundantly folds the <= 32 slots with two more REDUX.
// Tie rule: LOWEST index among equal maxima (documented in include/iapt_b200.h and DESIGN.md).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace iapt {

constexpr int kFpsMaxSmemPoints = 14000;  // 16 B per point (float4) -> 224 KB

// Tie order among equal maxima of the running minimum.  Default (lg < 0): the LOWEST index.  IAPT_FPS_TIE_UPSTREAM:
// upstream's order as published (pointnet2_ops furthest_point_sampling_kernel<block>: thread k mod block scans its
// points with strict `>`, the shared-memory tree keeps the lower thread) = the smallest (k mod block, k / block),
// block = 2^lg = the largest power of two <= N, at most 512.  The reductions below minimise an integer among the
// candidates holding the maximum; in upstream mode that integer is this key instead of the index.  The kernels get lg
// in bits 8..15 of `flags` (as lg + 1; 0 = lowest index).
__device__ __forceinline__ int fps_tie_key(int k, int lg) { return ((k & ((1 << lg) - 1)) << 16) | (k >> lg); }
__device__ __forceinline__ int fps_tie_unkey(int key, int lg) { return ((key & 0xffff) << lg) | (key >> 16); }
__device__ __forceinline__ int fps_tie_lg(int flags) { return ((flags >> 8) & 0xff) - 1; }

struct FpsSlots {
  int val[2][32];
  int idx[2][32];
  float cx[2][32], cy[2][32], cz[2][32];  // cluster kernel: coordinates of each warp's candidate
};

// kRegs: points cached in registers (N <= kT * kPPT); otherwise re-read from shared memory each round.
// Shared memory: the cloud as float4 (one LDS.128 fetches the selected point's coordinates) and one int2
// (value bits, index) slot per warp, double-buffered by round parity.
template <int kT, int kPPT, bool kRegs>
__global__ void __launch_bounds__(kT) fps_kernel(const float* __restrict__ xyz, int N, int npoint, int flags,
                                                 int32_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int2* wslot = reinterpret_cast<int2*>(smem_raw);          // [2][32]
  float4* sp = reinterpret_cast<float4*>(smem_raw + 512);   // [N]

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  constexpr int kNW = kT / 32;
  const float* __restrict__ P = xyz + (size_t)b * N * 3;
  int32_t* __restrict__ out = idx_out + (size_t)b * npoint;

  for (int k = tid; k < N; k += kT) {
    const float* p = P + 3 * (size_t)k;
    sp[k] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
  }
  __syncthreads();

  float px[kRegs ? kPPT : 1], py[kRegs ? kPPT : 1], pz[kRegs ? kPPT : 1];
  float tmp[kPPT];  // running min distance; -1 marks a point that can never be selected
#pragma unroll
  for (int i = 0; i < kPPT; ++i) {
    const int k = tid + i * kT;
    float x = 0.f, y = 0.f, z = 0.f;
    bool ok = k < N;
    if (ok) {
      const float4 p = sp[k];
      x = p.x;
      y = p.y;
      z = p.z;
      if (flags & 1) {
        // upstream: `if (mag <= 1e-3) continue;` - float magnitude against the DOUBLE literal
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        ok = !((double)mag <= 1e-3);
      }
    }
    if (kRegs) {
      px[i] = x;
      py[i] = y;
      pz[i] = z;
    }
    tmp[i] = ok ? 1e10f : -1.f;
  }

  int old = 0;
  if (tid == 0) out[0] = 0;
  const int lg = fps_tie_lg(flags);

  for (int j = 1; j < npoint; ++j) {
    const float4 o = sp[old];
    float best = -1.f;
    int besti = 0;
#pragma unroll
    for (int i = 0; i < kPPT; ++i) {
      const int k = tid + i * kT;
      float x, y, z;
      if (kRegs) {
        x = px[i];
        y = py[i];
        z = pz[i];
      } else {
        const float4 p = sp[k < N ? k : 0];
        x = p.x;
        y = p.y;
        z = p.z;
      }
      const float d = sqdist(o.x, o.y, o.z, x, y, z);
      const float t = fminf(tmp[i], d);  // stays -1 for ineligible points (d >= 0)
      tmp[i] = t;
      if (lg < 0) {
        if (t > best) {  // ascending k inside a thread + strict '>' keeps the thread's lowest index
          best = t;
          besti = k;
        }
      } else {
        const int key = fps_tie_key(k, lg);
        if (t > best || (t == best && key < besti)) {
          best = t;
          besti = key;
        }
      }
    }
    // warp argmax: value bits order like signed ints (-1.f is negative, distances are >= +0)
    const int vb = __float_as_int(best);
    const int vmax = redux_max_s32(vb);
    const int imin = redux_min_s32(vb == vmax ? besti : 0x7fffffff);
    const int par = (j & 1) * 32;
    if (lane == 0) wslot[par + warp] = make_int2(vmax, imin);
    __syncthreads();
    int2 s = make_int2((int)0x80000000, 0x7fffffff);
    if (lane < kNW) s = wslot[par + lane];
    const int gmax = redux_max_s32(s.x);
    old = redux_min_s32(s.x == gmax ? s.y : 0x7fffffff);
    if (lg >= 0) old = fps_tie_unkey(old, lg);
    if (tid == 0) out[j] = old;
  }
}

// ---- thread-block-cluster FPS for large clouds ------------------------------------------------------------------
// One CTA per cloud leaves 100 of 148 SMs idle at B = 48 and makes a round cost N/128 lane-cycles.  For N >= 8192 a
// cluster of kC CTAs shares one cloud: every CTA keeps N/kC points and their running minima in registers, finds its
// local argmax exactly as above, then warp 0 pushes one record {value bits, index, x, y, z} into the exchange slot of
// EVERY CTA of the cluster through distributed shared memory; after one cluster barrier each CTA folds the kC records
// (max value, min index) and already holds the winner's coordinates for the next round - no remote read on the
// critical path.  The exchange buffer is double-buffered by round parity (a CTA may run one round ahead).
struct FpsRecord {
  int val, idx;
  float x, y, z;
  int pad[3];
};

template <int kT, int kPPT, int kC>
__global__ void __launch_bounds__(kT) fps_cluster_kernel(const float* __restrict__ xyz, int N, int npoint, int flags,
                                                         int32_t* __restrict__ idx_out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FpsSlots* slots = reinterpret_cast<FpsSlots*>(smem_raw);
  FpsRecord* xchg = reinterpret_cast<FpsRecord*>(smem_raw + sizeof(FpsSlots));  // [2][kC]
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / kC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kNW = kT / 32;
  const int chunk = (N + kC - 1) / kC;  // points [rank*chunk, ...) belong to this CTA
  const int base = rank * chunk;
  const float* __restrict__ P = xyz + (size_t)b * N * 3;
  int32_t* __restrict__ out = idx_out + (size_t)b * npoint;

  float px[kPPT], py[kPPT], pz[kPPT], tmp[kPPT];
#pragma unroll
  for (int i = 0; i < kPPT; ++i) {
    const int local = tid + i * kT;
    const int k = base + local;
    bool ok = local < chunk && k < N;
    float x = 0.f, y = 0.f, z = 0.f;
    if (ok) {
      x = __ldg(P + 3 * (size_t)k);
      y = __ldg(P + 3 * (size_t)k + 1);
      z = __ldg(P + 3 * (size_t)k + 2);
      if (flags & 1) {
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        ok = !((double)mag <= 1e-3);
      }
    }
    px[i] = x;
    py[i] = y;
    pz[i] = z;
    tmp[i] = ok ? 1e10f : -1.f;
  }
  float ox = __ldg(P), oy = __ldg(P + 1), oz = __ldg(P + 2);  // round 0 selects index 0
  if (rank == 0 && tid == 0) out[0] = 0;
  const int lg = fps_tie_lg(flags);
  cluster.sync();  // every CTA of the cluster is resident before the first remote store

  for (int j = 1; j < npoint; ++j) {
    float best = -1.f;
    int besti = 0;
    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int i = 0; i < kPPT; ++i) {
      const float d = sqdist(ox, oy, oz, px[i], py[i], pz[i]);
      const float t = fminf(tmp[i], d);
      tmp[i] = t;
      const int kk = lg < 0 ? base + tid + i * kT : fps_tie_key(base + tid + i * kT, lg);
      if (t > best || (lg >= 0 && t == best && kk < besti)) {
        best = t;
        besti = kk;
        bx = px[i];
        by = py[i];
        bz = pz[i];
      }
    }
    const int vb = __float_as_int(best);
    const int vmax = redux_max_s32(vb);
    const int imin = redux_min_s32(vb == vmax ? besti : 0x7fffffff);
    const int par = j & 1;
    // the lane that owns the warp's winner publishes its coordinates with it
    if (vb == vmax && besti == imin) {
      slots->val[par][warp] = vmax;
      slots->idx[par][warp] = imin;
      reinterpret_cast<float*>(slots->cx[par])[warp] = bx;
      reinterpret_cast<float*>(slots->cy[par])[warp] = by;
      reinterpret_cast<float*>(slots->cz[par])[warp] = bz;
    }
    __syncthreads();
    if (warp == 0) {
      const int v = (lane < kNW) ? slots->val[par][lane] : (int)0x80000000;
      const int ii = (lane < kNW) ? slots->idx[par][lane] : 0x7fffffff;
      const int gmax = redux_max_s32(v);
      const int gidx = redux_min_s32(v == gmax ? ii : 0x7fffffff);
      // lane holding the CTA winner pushes the record to every CTA of the cluster (DSMEM stores)
      if (lane < kNW && v == gmax && ii == gidx) {
        FpsRecord r;
        r.val = gmax;
        r.idx = gidx;
        r.x = slots->cx[par][lane];
        r.y = slots->cy[par][lane];
        r.z = slots->cz[par][lane];
#pragma unroll
        for (int c = 0; c < kC; ++c) {
          FpsRecord* dst = cluster.map_shared_rank(xchg, c) + par * kC + rank;
          dst->val = r.val;
          dst->idx = r.idx;
          dst->x = r.x;
          dst->y = r.y;
          dst->z = r.z;
        }
      }
    }
    cluster.sync();  // records of all kC CTAs have landed (release/acquire at cluster scope)
    {
      const FpsRecord* rec = xchg + par * kC;
      int v = (int)0x80000000, ii = 0x7fffffff;
      if (lane < kC) {
        v = rec[lane].val;
        ii = rec[lane].idx;
      }
      const int gmax = redux_max_s32(v);
      const int old = redux_min_s32(v == gmax ? ii : 0x7fffffff);
      // coordinates of the winner: the record whose (val, idx) won
      int src = 0;
#pragma unroll
      for (int c = 0; c < kC; ++c)
        if (rec[c].val == gmax && rec[c].idx == old) src = c;
      if (gmax < 0) {  // nothing eligible anywhere: upstream selects index 0
        ox = __ldg(P);
        oy = __ldg(P + 1);
        oz = __ldg(P + 2);
      } else {
        ox = rec[src].x;
        oy = rec[src].y;
        oz = rec[src].z;
      }
      if (rank == 0 && tid == 0) out[j] = gmax < 0 ? 0 : (lg < 0 ? old : fps_tie_unkey(old, lg));
    }
  }
  cluster.sync();  // no CTA exits while a peer may still write into its shared memory
}

// Generic fallback for clouds too large for shared memory: points from global (L2), running min in a workspace.
__global__ void __launch_bounds__(1024) fps_generic_kernel(const float* __restrict__ xyz, int N, int npoint, int flags,
                                                           float* __restrict__ temp_all,
                                                           int32_t* __restrict__ idx_out) {
  __shared__ FpsSlots slots;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ P = xyz + (size_t)b * N * 3;
  float* __restrict__ temp = temp_all + (size_t)b * N;
  int32_t* __restrict__ out = idx_out + (size_t)b * npoint;
  for (int k = tid; k < N; k += 1024) {
    bool ok = true;
    if (flags & 1) {
      const float x = P[3 * (size_t)k], y = P[3 * (size_t)k + 1], z = P[3 * (size_t)k + 2];
      const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
      ok = !((double)mag <= 1e-3);
    }
    temp[k] = ok ? 1e10f : -1.f;
  }
  int old = 0;
  if (tid == 0) out[0] = 0;
  const int lg = fps_tie_lg(flags);
  __syncthreads();
  for (int j = 1; j < npoint; ++j) {
    const float ox = P[3 * (size_t)old], oy = P[3 * (size_t)old + 1], oz = P[3 * (size_t)old + 2];
    float best = -1.f;
    int besti = 0;
    for (int k = tid; k < N; k += 1024) {
      const float d = sqdist(ox, oy, oz, P[3 * (size_t)k], P[3 * (size_t)k + 1], P[3 * (size_t)k + 2]);
      const float t = fminf(temp[k], d);
      temp[k] = t;
      const int kk = lg < 0 ? k : fps_tie_key(k, lg);
      if (t > best || (lg >= 0 && t == best && kk < besti)) {
        best = t;
        besti = kk;
      }
    }
    const int vb = __float_as_int(best);
    const int vmax = redux_max_s32(vb);
    const int imin = redux_min_s32(vb == vmax ? besti : 0x7fffffff);
    const int par = j & 1;
    if (lane == 0) {
      slots.val[par][warp] = vmax;
      slots.idx[par][warp] = imin;
    }
    __syncthreads();
    const int v = slots.val[par][lane];
    const int ii = slots.idx[par][lane];
    const int gmax = redux_max_s32(v);
    old = redux_min_s32(v == gmax ? ii : 0x7fffffff);
    if (lg >= 0) old = fps_tie_unkey(old, lg);
    if (tid == 0) out[j] = old;
  }
}

size_t fps_workspace_bytes(int B, int N, int /*npoint*/) {
  return N > kFpsMaxSmemPoints ? sizeof(float) * (size_t)B * N : 0;
}

template <int kT, int kPPT, bool kRegs>
static int fps_launch_t(const float* xyz, int B, int N, int npoint, int flags, int32_t* idx, cudaStream_t stream) {
  const size_t smem = 512 + sizeof(float4) * (size_t)N;
  static PerDeviceOnce once;
  if (once.pending()) {
    cudaFuncSetAttribute(fps_kernel<kT, kPPT, kRegs>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    once.done();
  }
  fps_kernel<kT, kPPT, kRegs><<<B, kT, smem, stream>>>(xyz, N, npoint, flags, idx);
  count_launch();
  return (int)cudaGetLastError();
}

template <int kT, int kPPT, int kC>
static int fps_cluster_launch_t(const float* xyz, int B, int N, int npoint, int flags, int32_t* idx,
                                cudaStream_t stream) {
  const size_t smem = sizeof(FpsSlots) + sizeof(FpsRecord) * 2 * kC;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B * kC, 1, 1);
  cfg.blockDim = dim3(kT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, fps_cluster_kernel<kT, kPPT, kC>, xyz, N, npoint, flags, idx);
  count_launch();
  return (int)(err != cudaSuccess ? err : cudaGetLastError());
}

// Cluster width for a large cloud: the widest of {8,4,2} that keeps all B clusters resident at once (one CTA per SM)
// and leaves every CTA at least 2048 points; 1 = use the single-CTA kernels.
static int fps_cluster_width(int B, int N) {
  static int sms_of_device[256] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int sms = sms_of_device[dev & 255];
  if (sms == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    sms_of_device[dev & 255] = sms;
  }
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = std::getenv("IAPT_FPS_CLUSTER");  // development knob: 0 disables the cluster kernels
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  // Measured (profiles/r01d_fps_sizes.txt): a round costs ~0.12 + 0.075 us per 1024 points per CTA, plus ~0.55 us for
  // the cluster exchange (remote stores + barrier.cluster).  A single CTA holds up to 8192 points in registers
  // (0.73 us/round), so clusters pay off only beyond that - where one CTA would fall back to shared / global memory.
  // (A barrier-free variant - tagged records, st.release.cluster + polling - was measured SLOWER: 1.8 vs 1.44 us/round.)
  if (!enabled || N <= 8192) return 1;
  // Widest power-of-two cluster that keeps every cluster resident (one CTA per SM).  Widths 3 and 6 were tried for
  // B=48 x 16384 (144 of 148 SMs): clusters live inside a GPC (16/18/20 SMs), they did not all fit at once and the
  // kernel ran in two waves (6.0 ms vs 2.9 ms with width 2).
  const int widths[] = {8, 4, 2};
  for (int c : widths)
    if ((long long)B * c <= sms && N / c >= 2048) return c;
  return 1;
}

int fps_launch(const float* xyz, int B, int N, int npoint, int flags, int32_t* idx, void* workspace,
               cudaStream_t stream) {
  {  // bit 1 (IAPT_FPS_TIE_UPSTREAM) -> lg + 1 in bits 8..15 for the kernels
    int lg = -1;
    if (flags & 2) {
      lg = 0;
      while ((2 << lg) <= N && lg < 9) ++lg;  // block = largest power of two <= N, at most 512
    }
    flags = (flags & 1) | ((lg + 1) << 8);
  }
  const int cw = fps_cluster_width(B, N);
  if (cw > 1) {
    const int per = (N + cw - 1) / cw;  // points per CTA: registers hold up to 1024 x 8
    if (per <= 8192) {
#define IAPT_FPS_CLUSTER_CASE(C)                                                                          \
  if (cw == C)                                                                                            \
    return per <= 4096 ? fps_cluster_launch_t<512, 8, C>(xyz, B, N, npoint, flags, idx, stream)          \
                       : fps_cluster_launch_t<1024, 8, C>(xyz, B, N, npoint, flags, idx, stream);
      IAPT_FPS_CLUSTER_CASE(2)
      IAPT_FPS_CLUSTER_CASE(4)
      IAPT_FPS_CLUSTER_CASE(8)
#undef IAPT_FPS_CLUSTER_CASE
    }
  }
  if (N <= 128) return fps_launch_t<128, 1, true>(xyz, B, N, npoint, flags, idx, stream);
  if (N <= 256) return fps_launch_t<256, 1, true>(xyz, B, N, npoint, flags, idx, stream);
  if (N <= 512) return fps_launch_t<256, 2, true>(xyz, B, N, npoint, flags, idx, stream);
  if (N <= 1024) return fps_launch_t<512, 2, true>(xyz, B, N, npoint, flags, idx, stream);
  if (N <= 2048) return fps_launch_t<512, 4, true>(xyz, B, N, npoint, flags, idx, stream);  // 256x8: 0.184, 1024x2: 0.167, 512x4: 0.160 ms
  if (N <= 4096) return fps_launch_t<512, 8, true>(xyz, B, N, npoint, flags, idx, stream);
  if (N <= 8192) return fps_launch_t<1024, 8, true>(xyz, B, N, npoint, flags, idx, stream);
  if (N <= kFpsMaxSmemPoints) return fps_launch_t<1024, 14, false>(xyz, B, N, npoint, flags, idx, stream);
  fps_generic_kernel<<<B, 1024, 0, stream>>>(xyz, N, npoint, flags, static_cast<float*>(workspace), idx);
  count_launch();
  return (int)cudaGetLastError();
}

// ---- gathers -----------------------------------------------------------------------------------------------
// out[b,c,j] = features[b,c,idx[b,j]]
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                                                     int C, int N, int m, float* __restrict__ out, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % m);
    const long long bc = t / m;
    const int b = (int)(bc / C);
    out[t] = __ldg(feat + bc * N + __ldg(idx + (size_t)b * m + j));
  }
}

// out[b,j,:] = xyz[b,idx[b,j],:]   (utils.misc.fps without the transposes)
__global__ void __launch_bounds__(256) gather_xyz_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                                                         int N, int m, float* __restrict__ out, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % 3);
    const long long bj = t / 3;
    const int b = (int)(bj / m);
    out[t] = __ldg(xyz + ((size_t)b * N + __ldg(idx + bj)) * 3 + c);
  }
}

// grad_features[b,c,n] = sum_{j asc : idx[b,j]==n} grad_out[b,c,j].  Owner-computes (one thread per (b,n)),
// deterministic, every output written exactly once.  idx[b,:] is staged in shared memory.
__global__ void __launch_bounds__(256) gather_grad_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx,
                                                          int C, int N, int m, float* __restrict__ gfeat) {
  extern __shared__ int sidx[];
  const int b = blockIdx.y;
  for (int j = threadIdx.x; j < m; j += blockDim.x) sidx[j] = __ldg(idx + (size_t)b * m + j);
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int first[4];
  int cnt = 0;
  for (int j = 0; j < m; ++j) {
    if (sidx[j] == n) {
      if (cnt < 4) first[cnt] = j;
      ++cnt;
    }
  }
  for (int c = 0; c < C; ++c) {
    const float* g = gout + ((size_t)b * C + c) * m;
    float v = 0.f;
    if (cnt <= 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < cnt) v = __fadd_rn(v, __ldg(g + first[q]));
    } else {
      for (int j = 0; j < m; ++j)
        if (sidx[j] == n) v = __fadd_rn(v, __ldg(g + j));
    }
    gfeat[((size_t)b * C + c) * N + n] = v;
  }
}

static inline int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  if (g > 148LL * 32) g = 148LL * 32;
  if (g < 1) g = 1;
  return (int)g;
}

int gather_launch(const float* feat, const int32_t* idx, int B, int C, int N, int m, float* out, cudaStream_t stream) {
  const long long total = (long long)B * C * m;
  gather_kernel<<<grid_for(total, 256), 256, 0, stream>>>(feat, idx, C, N, m, out, total);
  count_launch();
  return (int)cudaGetLastError();
}

int gather_xyz_launch(const float* xyz, const int32_t* idx, int B, int N, int m, float* out, cudaStream_t stream) {
  const long long total = (long long)B * m * 3;
  gather_xyz_kernel<<<grid_for(total, 256), 256, 0, stream>>>(xyz, idx, N, m, out, total);
  count_launch();
  return (int)cudaGetLastError();
}

int gather_grad_launch(const float* gout, const int32_t* idx, int B, int C, int N, int m, float* gfeat,
                       cudaStream_t stream) {
  dim3 grid((N + 255) / 256, B);
  const size_t smem = sizeof(int) * (size_t)m;
  if (smem > 48 * 1024) {
    static PerDeviceOnce once;
    if (once.pending()) {
      cudaFuncSetAttribute(gather_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      once.done();
    }
  }
  gather_grad_kernel<<<grid, 256, smem, stream>>>(gout, idx, C, N, m, gfeat);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
