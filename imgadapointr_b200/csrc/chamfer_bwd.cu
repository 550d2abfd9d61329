// chamfer_bwd.cu - deterministic, floating-point-atomics-free Chamfer backward for sm_100a.
//
// Replaces chamfer_dist_grad_kernel x2 (reference extensions/chamfer_dist/chamfer.cu:173-201, launched :215-222),
// which scatters with 6 float atomicAdd per point from only 4096 threads (grid (1,16), block 256) and therefore
// sums in a run-dependent order.
//
// Owner-computes formulation.  For output point j of cloud side X (other side Y):
//     grad_X[j] = 2 gX[j] (x_j - y_{idxX[j]})                      "own" term   (chamfer.cu:193-195)
//               - sum_{k : idxY[k] == j} 2 gY[k] (y_k - x_j)       fan-in term  (chamfer.cu:196-198, roles swapped)
// A warp owns kBins consecutive output points and keeps their fan-in accumulators in shared memory.  It scans the
// other side's index array idxY[b, :] in ascending k (coalesced 128 B loads, 8 in flight per lane) and only
// COMPACTS the hits (k, bin) into a small shared-memory list; whenever the list fills it is drained 32 hits at a
// time: the lanes gather gY[k], y_k, x_j in parallel (one round trip to L2 for 32 hits instead of one per hit) and
// lanes that hit the same bin are then serialised in list order.  Every accumulator therefore receives its
// contributions in ascending k: the result is bit-reproducible and independent of the launch geometry, with no
// floating-point atomics.  Each key is read N / kBins times from L2 (~1 % of the forward search's work) and
// worst-case fan-in (all keys in one bin) degrades to a sequential sum inside one warp.
//
// Optional fused loss mode (iapt_chamfer_backward_fused): instead of reading grad_dist arrays the kernel forms
// g = scale / sqrt(dist) (L1) or g = scale (L2) on the fly from the saved distances - the analytic gradient of the
// reference's sqrt/mean epilogues (extensions/chamfer_dist/__init__.py:82-84) - times the upstream scalar gradient.
#include <cstdint>
#include <cstdlib>

#include "common.cuh"

namespace iapt {

constexpr int kBwdThreads = 256;
constexpr int kBwdWarps = kBwdThreads / 32;
constexpr int kBins = 512;     // output points per warp (maximum; the launcher picks 32..512 per side, see scan_bins)
constexpr int kHitCap = 256;   // compacted hits per warp before a drain (>= 128 + 32)

struct BwdParams {
  const float* xyz[2];     // [B, n[s], 3]
  const int32_t* idx[2];   // [B, n[s]]   idx[s][j] indexes into side 1-s
  const float* gdist[2];   // plain mode: grad_dist [B, n[s]] (null = zeros).  fused mode: dist [B, n[s]]
  float* grad[2];          // [B, n[s], 3]
  int n[2];
  int B;
  int bins[2];              // output points per warp of side s (power of two, 32..kBins)
  int blocks_per_cloud[2];  // ceil(n[s] / bins[s])
  long long warps_side0;    // B * blocks_per_cloud[0]
  long long total_warps;
  // fused loss mode: g[s] = upstream * scale[s] * (mode == 1 ? 1/sqrt(dist) : 1); upstream read from *grad_out
  int mode;                 // 0 plain, 1 L1 (sqrt), 2 L2
  float scale[2];
  const float* grad_out;    // device: upstream[side] = grad_out[side * grad_out_stride] (null = 1)
  int grad_out_stride;
  // fan-in lists (fast path): cnt[s][b*n[s] + j] = number of side-(1-s) points whose nearest neighbour is j,
  // list[s][(b*n[s] + j)*cap[s] + slot] = their indices in arrival order (sorted on the fly by the gather kernel).
  // cap[s] == 0: no lists for that side, every bin is handled by the scan kernel.  Null cnt = scan everything.
  int* cnt[2];
  int* list[2];
  int cap[2];
  int* overflow_flag;       // [2]: set when some bin of side s overflowed its list
};

struct BwdSmem {
  unsigned char over[kBwdWarps][kBins];
  float acc[kBwdWarps][kBins * 3];
  int hit_k[kBwdWarps][kHitCap];
  unsigned short hit_bin[kBwdWarps][kHitCap];
};

__device__ __forceinline__ float point_weight(int mode, float scaled_upstream, const float* __restrict__ g, int i) {
  // returns 2 * dLoss/d dist[i]
  if (mode == 0) return __fmul_rn(__ldg(g + i), 2.f);
  float w = scaled_upstream;
  if (mode == 1) w = __fdiv_rn(w, sqrtf(__ldg(g + i)));  // d/dd sqrt(d) = 0.5/sqrt(d); the 0.5 is in scale
  return __fmul_rn(w, 2.f);
}

__global__ void __launch_bounds__(kBwdThreads) chamfer_bwd_kernel(BwdParams p) {
  extern __shared__ __align__(16) unsigned char bwd_smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(bwd_smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * kBwdWarps + warp;
  if (gw >= p.total_warps) return;  // whole warp exits together; no block-level sync below

  int side;
  long long w = gw;
  if (w < p.warps_side0) {
    side = 0;
  } else {
    side = 1;
    w -= p.warps_side0;
  }
  const int bpc = side ? p.blocks_per_cloud[1] : p.blocks_per_cloud[0];
  const int b = (int)(w / bpc);
  const int bins = side ? p.bins[1] : p.bins[0];
  const int j0 = (int)(w - (long long)b * bpc) * bins;
  const int nX = side ? p.n[1] : p.n[0], nY = side ? p.n[0] : p.n[1];
  const float* __restrict__ X = (side ? p.xyz[1] : p.xyz[0]) + (size_t)b * nX * 3;
  const float* __restrict__ Y = (side ? p.xyz[0] : p.xyz[1]) + (size_t)b * nY * 3;
  const int32_t* __restrict__ idxX = (side ? p.idx[1] : p.idx[0]) + (size_t)b * nX;
  const int32_t* __restrict__ idxY = (side ? p.idx[0] : p.idx[1]) + (size_t)b * nY;
  const float* gXb = side ? p.gdist[1] : p.gdist[0];
  const float* gYb = side ? p.gdist[0] : p.gdist[1];
  const float* __restrict__ gX = gXb ? gXb + (size_t)b * nX : nullptr;
  const float* __restrict__ gY = gYb ? gYb + (size_t)b * nY : nullptr;
  float* __restrict__ out = (side ? p.grad[1] : p.grad[0]) + (size_t)b * nX * 3;
  const float scaleX = side ? p.scale[1] : p.scale[0], scaleY = side ? p.scale[0] : p.scale[1];
  float* a = sm.acc[warp];
  int* hk = sm.hit_k[warp];
  unsigned short* hb = sm.hit_bin[warp];
  float up0 = 1.f, up1 = 1.f;
  if (p.mode != 0 && p.grad_out != nullptr) {
    up0 = __ldg(p.grad_out);
    up1 = __ldg(p.grad_out + p.grad_out_stride);
  }
  const int mode = p.mode;
  const float wX = __fmul_rn(side ? up1 : up0, scaleX);  // upstream * scale of this side / of the other side
  const float wY = __fmul_rn(side ? up0 : up1, scaleY);
  const unsigned lt_mask = (1u << lane) - 1u;

  // With fan-in lists, this kernel only serves the bins whose list overflowed (cnt > cap): nothing to do otherwise.
  const int* cntX = side ? p.cnt[1] : p.cnt[0];
  const int capX = side ? p.cap[1] : p.cap[0];
  const bool lists = cntX != nullptr && capX > 0;
  if (lists) {
    if (p.overflow_flag[side] == 0) return;  // warp-uniform: no bin of this side overflowed anywhere
    bool mine = false;
    for (int i = lane; i < bins; i += 32) {
      const int j = j0 + i;
      const bool over = j < nX && cntX[(size_t)b * nX + j] > capX;
      sm.over[warp][i] = over ? 1 : 0;
      mine |= over;
    }
    if (!__any_sync(0xffffffffu, mine)) return;
    __syncwarp();  // over[] visible to every lane
  }
  for (int i = lane; i < bins * 3; i += 32) a[i] = 0.f;
  __syncwarp();

  int count = 0;  // warp-uniform
  auto drain = [&]() {
    for (int i0 = 0; i0 < count; i0 += 32) {
      const int i = i0 + lane;
      const bool on = i < count;
      float cx = 0.f, cy = 0.f, cz = 0.f;
      unsigned bin = 0, rank = 0;
      const unsigned active = __ballot_sync(0xffffffffu, on);
      if (on) {
        const int k = hk[i];
        bin = hb[i];
        const int j = j0 + (int)bin;
        const float g = point_weight(mode, wY, gY, k);
        cx = __fmul_rn(g, __fsub_rn(__ldg(Y + 3 * (size_t)k + 0), __ldg(X + 3 * (size_t)j + 0)));
        cy = __fmul_rn(g, __fsub_rn(__ldg(Y + 3 * (size_t)k + 1), __ldg(X + 3 * (size_t)j + 1)));
        cz = __fmul_rn(g, __fsub_rn(__ldg(Y + 3 * (size_t)k + 2), __ldg(X + 3 * (size_t)j + 2)));
        const unsigned peers = __match_any_sync(active, bin);
        rank = __popc(peers & lt_mask);
      }
      // hits of the same bin take turns in list (= ascending k) order
      for (unsigned r = 0;; ++r) {
        if (on && rank == r) {
          a[bin * 3 + 0] = __fadd_rn(a[bin * 3 + 0], cx);
          a[bin * 3 + 1] = __fadd_rn(a[bin * 3 + 1], cy);
          a[bin * 3 + 2] = __fadd_rn(a[bin * 3 + 2], cz);
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, on && rank > r)) break;
      }
    }
    count = 0;
  };

  if (gY != nullptr) {
    // ---- scan the other side's nearest-neighbour indices, 256 keys per step, ascending k ---------------------
    for (int k0 = 0; k0 < nY; k0 += 256) {
      int key[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e * 32 + lane;
        key[e] = (k < nY) ? __ldg(idxY + k) : -1;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned bin = (unsigned)(key[e] - j0);
        bool hit = (key[e] >= 0) && (bin < (unsigned)bins) && ((int)bin + j0 < nX);
        if (lists && hit) hit = sm.over[warp][bin] != 0;
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask != 0u) {  // warp-uniform
          if (hit) {
            const int pos = count + __popc(mask & lt_mask);
            hk[pos] = k0 + e * 32 + lane;
            hb[pos] = (unsigned short)bin;
          }
          count += __popc(mask);
          __syncwarp();
          if (count > kHitCap - 32) drain();
        }
      }
    }
    __syncwarp();
    drain();
  }
  __syncwarp();

  // ---- own term and write-out ------------------------------------------------------------------------------
  for (int i = lane; i < bins; i += 32) {
    const int j = j0 + i;
    if (j >= nX) break;
    if (lists && sm.over[warp][i] == 0) continue;  // written by the gather kernel
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (gX != nullptr) {
      const int j2 = __ldg(idxX + j);
      const float g = point_weight(mode, wX, gX, j);
      ox = __fmul_rn(g, __fsub_rn(__ldg(X + 3 * (size_t)j + 0), __ldg(Y + 3 * (size_t)j2 + 0)));
      oy = __fmul_rn(g, __fsub_rn(__ldg(X + 3 * (size_t)j + 1), __ldg(Y + 3 * (size_t)j2 + 1)));
      oz = __fmul_rn(g, __fsub_rn(__ldg(X + 3 * (size_t)j + 2), __ldg(Y + 3 * (size_t)j2 + 2)));
    }
    out[3 * (size_t)j + 0] = __fsub_rn(ox, a[i * 3 + 0]);
    out[3 * (size_t)j + 1] = __fsub_rn(oy, a[i * 3 + 1]);
    out[3 * (size_t)j + 2] = __fsub_rn(oz, a[i * 3 + 2]);
  }
}

// ---- fast path: fan-in lists ------------------------------------------------------------------------------------
// link: every point appends itself to the list of its nearest neighbour (integer atomics: the SET is deterministic, the
// order is not - the gather kernel below restores ascending order).
__global__ void __launch_bounds__(256) chamfer_link_kernel(BwdParams p, long long total0, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int src = t < total0 ? 0 : 1;  // the side the point belongs to; it links into side 1 - src
    const long long u = src ? t - total0 : t;
    const int dst = src ^ 1;
    const int cap = dst ? p.cap[1] : p.cap[0];
    if (cap == 0) continue;
    const float* g = src ? p.gdist[1] : p.gdist[0];
    if (g == nullptr) continue;  // this side contributes nothing
    const int nS = src ? p.n[1] : p.n[0], nD = src ? p.n[0] : p.n[1];
    const int b = (int)(u / nS);
    const int k = (int)(u - (long long)b * nS);
    const int j = __ldg((src ? p.idx[1] : p.idx[0]) + u);
    int* cnt = dst ? p.cnt[1] : p.cnt[0];
    const int slot = atomicAdd(cnt + (size_t)b * nD + j, 1);
    if (slot < cap)
      (dst ? p.list[1] : p.list[0])[((size_t)b * nD + j) * cap + slot] = k;
    else
      p.overflow_flag[dst] = 1;
  }
}

// gather: one thread per output point: own term, then its fan-in list in ASCENDING source index, so the sum order
// equals the scan kernel's.  Lists of <= 8 entries (the common case: mean fan-in ~1) are pulled into registers with two
// 128-bit loads and ordered by a 19-exchange sorting network; longer lists fall back to selection from memory.
__device__ __forceinline__ void cswap(int& a, int& b) {
  const int lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}

__global__ void __launch_bounds__(256) chamfer_gather_kernel(BwdParams p, long long total0, long long total) {
  float up0 = 1.f, up1 = 1.f;
  if (p.mode != 0 && p.grad_out != nullptr) {
    up0 = __ldg(p.grad_out);
    up1 = __ldg(p.grad_out + p.grad_out_stride);
  }
  const int mode = p.mode;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int side = t < total0 ? 0 : 1;
    const long long u = side ? t - total0 : t;
    const int cap = side ? p.cap[1] : p.cap[0];
    if (cap == 0) continue;  // scan kernel owns this side
    const int nX = side ? p.n[1] : p.n[0], nY = side ? p.n[0] : p.n[1];
    const int b = (int)(u / nX);
    const float* __restrict__ X = (side ? p.xyz[1] : p.xyz[0]);
    const float* __restrict__ Y = (side ? p.xyz[0] : p.xyz[1]) + (size_t)b * nY * 3;
    const float* gX = side ? p.gdist[1] : p.gdist[0];
    const float* gYb = side ? p.gdist[0] : p.gdist[1];
    const float* gY = gYb ? gYb + (size_t)b * nY : nullptr;
    const int c = gY ? (side ? p.cnt[1] : p.cnt[0])[u] : 0;
    if (c > cap) continue;  // overflowed: the scan kernel computes and writes this point
    const int* lst = (side ? p.list[1] : p.list[0]) + (size_t)u * cap;
    int e[8];
    if (c > 0 && c <= 8) {  // rows are 32-byte aligned (cap is a multiple of 8, base 256-byte aligned)
      const int4 lo = *reinterpret_cast<const int4*>(lst);
      const int4 hi = c > 4 ? *reinterpret_cast<const int4*>(lst + 4) : make_int4(0, 0, 0, 0);
      e[0] = lo.x; e[1] = lo.y; e[2] = lo.z; e[3] = lo.w;
      e[4] = hi.x; e[5] = hi.y; e[6] = hi.z; e[7] = hi.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i >= c) e[i] = 0x7fffffff;
      if (c > 1) {  // 8-input sorting network (19 compare-exchanges)
        cswap(e[0], e[1]); cswap(e[2], e[3]); cswap(e[4], e[5]); cswap(e[6], e[7]);
        cswap(e[0], e[2]); cswap(e[1], e[3]); cswap(e[4], e[6]); cswap(e[5], e[7]);
        cswap(e[1], e[2]); cswap(e[5], e[6]); cswap(e[0], e[4]); cswap(e[3], e[7]);
        cswap(e[1], e[5]); cswap(e[2], e[6]);
        cswap(e[1], e[4]); cswap(e[3], e[6]);
        cswap(e[2], e[4]); cswap(e[3], e[5]);
        cswap(e[3], e[4]);
      }
    }
    const float wX = __fmul_rn(side ? up1 : up0, side ? p.scale[1] : p.scale[0]);
    const float wY = __fmul_rn(side ? up0 : up1, side ? p.scale[0] : p.scale[1]);
    const float x0 = __ldg(X + 3 * u), x1 = __ldg(X + 3 * u + 1), x2 = __ldg(X + 3 * u + 2);
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (gX != nullptr) {
      const int j2 = __ldg((side ? p.idx[1] : p.idx[0]) + u);
      const float g = point_weight(mode, wX, gX + (size_t)b * nX, (int)(u - (long long)b * nX));
      ox = __fmul_rn(g, __fsub_rn(x0, __ldg(Y + 3 * (size_t)j2 + 0)));
      oy = __fmul_rn(g, __fsub_rn(x1, __ldg(Y + 3 * (size_t)j2 + 1)));
      oz = __fmul_rn(g, __fsub_rn(x2, __ldg(Y + 3 * (size_t)j2 + 2)));
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (c <= 8) {
      // all gathers first (independent), then the ordered sum
      float gw[8], yx[8], yy[8], yz[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < c) {
          const int k = e[i];
          gw[i] = point_weight(mode, wY, gY, k);
          yx[i] = __ldg(Y + 3 * (size_t)k + 0);
          yy[i] = __ldg(Y + 3 * (size_t)k + 1);
          yz[i] = __ldg(Y + 3 * (size_t)k + 2);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < c) {
          ax = __fadd_rn(ax, __fmul_rn(gw[i], __fsub_rn(yx[i], x0)));
          ay = __fadd_rn(ay, __fmul_rn(gw[i], __fsub_rn(yy[i], x1)));
          az = __fadd_rn(az, __fmul_rn(gw[i], __fsub_rn(yz[i], x2)));
        }
      }
    } else {
      int prev = -1;
      for (int r = 0; r < c; ++r) {
        int k = 0x7fffffff;  // next larger source index
        for (int q = 0; q < c; ++q) {
          const int v = lst[q];
          if (v > prev && v < k) k = v;
        }
        prev = k;
        const float g = point_weight(mode, wY, gY, k);
        ax = __fadd_rn(ax, __fmul_rn(g, __fsub_rn(__ldg(Y + 3 * (size_t)k + 0), x0)));
        ay = __fadd_rn(ay, __fmul_rn(g, __fsub_rn(__ldg(Y + 3 * (size_t)k + 1), x1)));
        az = __fadd_rn(az, __fmul_rn(g, __fsub_rn(__ldg(Y + 3 * (size_t)k + 2), x2)));
      }
    }
    float* out = (side ? p.grad[1] : p.grad[0]) + 3 * u;
    out[0] = __fsub_rn(ox, ax);
    out[1] = __fsub_rn(oy, ay);
    out[2] = __fsub_rn(oz, az);
  }
}

// List capacity for the side with nX bins fed by nY keys: 16 when the mean fan-in is <= 2 (a single overflowing bin
// makes the scan kernel run at its full per-warp latency, so the capacity must make overflow a real exception), up to
// 64 for mean <= 14, 0 (scan kernel only) beyond that.
static int list_cap(int nX, int nY) {
  const double mean = (double)nY / (double)(nX > 0 ? nX : 1);
  if (mean <= 2.0) return 16;
  if (mean <= 6.0) return 32;
  if (mean <= 14.0) return 64;
  return 0;
}

size_t chamfer_backward_workspace_bytes(int B, int N, int M) {
  const size_t c1 = list_cap(N, M), c2 = list_cap(M, N);
  return 256 + 32 + sizeof(int) * ((size_t)B * N * (c1 ? 1 + c1 : 0) + (size_t)B * M * (c2 ? 1 + c2 : 0));
}

static int bwd_launch_common(BwdParams& p, int B, int N, int M, void* workspace, cudaStream_t stream) {
  p.n[0] = N;
  p.n[1] = M;
  p.B = B;
  p.cnt[0] = p.cnt[1] = nullptr;
  p.list[0] = p.list[1] = nullptr;
  p.cap[0] = p.cap[1] = 0;
  p.overflow_flag = nullptr;
  static int use_lists = -1;
  if (use_lists < 0) {
    // Measured: at B=48 x 16384^2 lists 0.141 ms (link 37 us + gather 72 us + memset) vs scan-only 0.131-0.149 ms (a
    // tie); for small batches the scan kernel is bound by its per-warp latency (every warp walks all keys) and the
    // lists win clearly (1 x 2048 x 16384 fwd+bwd: 0.29 -> 0.18 ms).  Lists are the default whenever the caller
    // provides the workspace; IAPT_BWD_LISTS=0 forces the scan kernel.
    const char* e = std::getenv("IAPT_BWD_LISTS");
    use_lists = (e && e[0] == '0') ? 0 : 1;
  }
  if (workspace != nullptr && use_lists) {
    const int c1 = list_cap(N, M), c2 = list_cap(M, N);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    p.overflow_flag = reinterpret_cast<int*>(ws);
    int* base = reinterpret_cast<int*>(ws + 256);
    size_t zero_ints = 0;
    if (c1) {
      p.cnt[0] = base;
      base += (size_t)B * N;
      zero_ints += (size_t)B * N;
    }
    if (c2) {
      p.cnt[1] = base;
      base += (size_t)B * M;
      zero_ints += (size_t)B * M;
    }
    base = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(base) + 31) & ~(uintptr_t)31);  // 32-byte rows
    if (c1) {
      p.list[0] = base;
      base += (size_t)B * N * c1;
    }
    if (c2) p.list[1] = base;
    p.cap[0] = c1;
    p.cap[1] = c2;
    if (c1 || c2) {
      cudaError_t err = cudaMemsetAsync(ws, 0, 256 + sizeof(int) * zero_ints, stream);
      if (err != cudaSuccess) return (int)err;
      const long long total0 = (long long)B * N, total = total0 + (long long)B * M;
      long long g = (total + 255) / 256;  // one point per thread: both kernels are bound by dependent L2 round trips
      if (g > (1LL << 20)) g = 1LL << 20;
      chamfer_link_kernel<<<(unsigned)g, 256, 0, stream>>>(p, total0, total);
      chamfer_gather_kernel<<<(unsigned)g, 256, 0, stream>>>(p, total0, total);
      count_launch(2);
    } else {
      p.overflow_flag = nullptr;
    }
  }
  // Output points per warp: every warp walks ALL keys of the other side, so fewer, fatter warps do less total work
  // (512 is best at B=48 x 16384: 0.13 ms vs 0.144 at 256 and 0.27 at 1024) - but a small output side must still
  // spread over enough warps to hide that walk (B=48, 512 x 16384: 0.52 ms with one 512-bin warp per cloud).
  auto scan_bins = [&](int n) {
    int bins = kBins;
    while (bins > 32 && (long long)B * ((n + bins - 1) / bins) < 1184) bins >>= 1;
    return bins;
  };
  p.bins[0] = scan_bins(N);
  p.bins[1] = scan_bins(M);
  p.blocks_per_cloud[0] = (N + p.bins[0] - 1) / p.bins[0];
  p.blocks_per_cloud[1] = (M + p.bins[1] - 1) / p.bins[1];
  p.warps_side0 = (long long)B * p.blocks_per_cloud[0];
  p.total_warps = p.warps_side0 + (long long)B * p.blocks_per_cloud[1];
  const long long blocks = (p.total_warps + kBwdWarps - 1) / kBwdWarps;
  static PerDeviceOnce once;
  if (once.pending()) {
    cudaFuncSetAttribute(chamfer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
    once.done();
  }
  chamfer_bwd_kernel<<<(unsigned)blocks, kBwdThreads, sizeof(BwdSmem), stream>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

int chamfer_backward_launch(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                            const float* grad_dist1, const float* grad_dist2, int B, int N, int M, float* grad_xyz1,
                            float* grad_xyz2, void* workspace, cudaStream_t stream) {
  BwdParams p;
  p.xyz[0] = xyz1;
  p.xyz[1] = xyz2;
  p.idx[0] = idx1;
  p.idx[1] = idx2;
  p.gdist[0] = grad_dist1;
  p.gdist[1] = grad_dist2;
  p.grad[0] = grad_xyz1;
  p.grad[1] = grad_xyz2;
  p.mode = 0;
  p.scale[0] = p.scale[1] = 1.f;
  p.grad_out = nullptr;
  p.grad_out_stride = 0;
  return bwd_launch_common(p, B, N, M, workspace, stream);
}

// mode 1: loss = scale1 * sum sqrt(dist1) + scale2 * sum sqrt(dist2);  mode 2: loss = scale1 * sum dist1 + scale2 * sum dist2.
// (scale already contains the 1/(B*N) of the mean and, for mode 1, the 0.5 of d sqrt.)  A zero scale skips that direction.
int chamfer_backward_fused_launch(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                                  const float* dist1, const float* dist2, int mode, float scale1, float scale2,
                                  const float* grad_out, int grad_out_stride, int B, int N, int M, float* grad_xyz1,
                                  float* grad_xyz2, void* workspace, cudaStream_t stream) {
  BwdParams p;
  p.xyz[0] = xyz1;
  p.xyz[1] = xyz2;
  p.idx[0] = idx1;
  p.idx[1] = idx2;
  p.gdist[0] = scale1 != 0.f ? dist1 : nullptr;
  p.gdist[1] = scale2 != 0.f ? dist2 : nullptr;
  p.grad[0] = grad_xyz1;
  p.grad[1] = grad_xyz2;
  p.mode = mode;
  p.scale[0] = scale1;
  p.scale[1] = scale2;
  p.grad_out = grad_out;
  p.grad_out_stride = grad_out_stride;
  return bwd_launch_common(p, B, N, M, workspace, stream);
}

}  // namespace iapt
