// chamfer_fwd.cu - bidirectional brute-force nearest-neighbour search for sm_100a.
//
// Replaces chamfer_dist_kernel x2 (reference extensions/chamfer_dist/chamfer.cu:15-145, launched :159-164).
//
// Design:
//  * ONE pass evaluates every unordered (a, s) pair once and feeds BOTH directions: the row minimum (nearest
//    streamed point for each register point) and the column minimum (nearest register point for each streamed
//    point).  The reference evaluates each pair twice (two launches with the roles swapped).
//  * "A side" points live in registers, QT (8 or 16) per thread, packed two-per-f32x2 so the 3 SUB + 1 MUL + 2 FMA
//    of the direct form issue as FADD2/FMUL2/FFMA2 (half the FP32 issue slots).  "S side" points are streamed
//    through double-buffered shared memory in a duplicated (x,x,y,y | z,z) layout and read with broadcast LDS.128.
//  * No per-pair argmin bookkeeping.  Only minima are tracked in the hot loop (FMNMX3: one ALU slot per two
//    pairs per direction); the index is recovered exactly afterwards:
//      rows   : per 32-point chunk "did my running minimum strictly decrease?" -> remember the FIRST such chunk
//               (the snapshot and the chunk id live in shared memory, not registers);
//      columns: REDUX.MIN over the warp + ballot of the lanes that hold the minimum -> lowest QT-point group;
//    both are folded across tiles/CTAs with a 64-bit atomicMin on (distance bits << 32 | coarse id), whose
//    unsigned order is the lexicographic (distance, id) order = the reference's lowest-index tie rule
//    (chamfer.cu:47-50,137-140).  chamfer_resolve_kernel then re-evaluates the <= 32 candidates of the winning
//    chunk/group with the SAME rounding and returns the first exact match.
//  * Work items = (cloud, A tile of 256*QT points, S split), handed to a persistent grid (2 CTAs per SM) through an
//    atomic ticket so B = 48 does not quantise badly over 148 SMs (the reference strides batches over gridDim.x=32).
#include <cstdlib>
#include <utility>
#include <vector>

#include "common.cuh"

namespace iapt {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSStage = 512;  // streamed points per smem stage
constexpr int kChunk = 32;    // streamed points per row-index chunk

struct FwdParams {
  const float* A;  // [B, NA, 3] register side
  const float* S;  // [B, NS, 3] streamed side
  int B, NA, NS;
  u64* PA;  // [B, NA] packed (dist, chunk id of S)
  u64* PS;  // [B, NS] packed (dist, group id of A)
  unsigned* ticket;  // starts at 0xFFFFFFFF (same memset as the keys); item = atomicAdd(ticket, 1) + 1
  int a_tiles, s_splits, stages_per_split, total_items;
};

template <int QT>
struct __align__(16) FwdSmem {
  ulonglong2 sxy[2][kSStage];       // (x,x),(y,y)
  ulonglong2 szz[2][kSStage / 2];   // (z,z) of two consecutive points
  uint4 col[kWarps][kSStage / 2];   // per warp, per streamed point pair: (min bits, ballot) x2
  float4 prev[QT / 4][kThreads];    // running minima at the last chunk boundary
  unsigned bestc[QT][kThreads];     // first chunk in which the running minimum strictly decreased
  unsigned next_item;
};

__device__ __forceinline__ void stage_fill_load(const float* __restrict__ Sb, int NS, int base, int tid,
                                                float (&v)[2][3]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int s = base + 2 * tid + h;
    if (s < NS) {
      v[h][0] = __ldg(Sb + 3 * (size_t)s + 0);
      v[h][1] = __ldg(Sb + 3 * (size_t)s + 1);
      v[h][2] = __ldg(Sb + 3 * (size_t)s + 2);
    } else {
      v[h][0] = v[h][1] = v[h][2] = kPadCoord();
    }
  }
}

template <int QT>
__device__ __forceinline__ void stage_fill_store(FwdSmem<QT>& sm, int buf, int tid, const float (&v)[2][3]) {
  ulonglong2 zz;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    ulonglong2 xy;
    xy.x = pack2(v[h][0], v[h][0]);
    xy.y = pack2(v[h][1], v[h][1]);
    sm.sxy[buf][2 * tid + h] = xy;
  }
  zz.x = pack2(v[0][2], v[0][2]);
  zz.y = pack2(v[1][2], v[1][2]);
  sm.szz[buf][tid] = zz;
}

// QT: A-side points per thread (8 or 16).
template <int QT>
__global__ void __launch_bounds__(kThreads, 2) chamfer_pair_kernel(FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem<QT>& sm = *reinterpret_cast<FwdSmem<QT>*>(smem_raw);
  constexpr int kATile = kThreads * QT;
  constexpr int kPairs = QT / 2;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const float kInf = __int_as_float(0x7f800000);

  const int total_stages = (p.NS + kSStage - 1) / kSStage;

  for (;;) {
    __syncthreads();  // everyone is done with smem (incl. next_item) of the previous item
    if (tid == 0) sm.next_item = atomicAdd(p.ticket, 1u) + 1u;
    __syncthreads();
    const int item = (int)sm.next_item;
    if (item >= p.total_items) break;
    const int per_cloud = p.a_tiles * p.s_splits;
    const int b = item / per_cloud;
    const int rem = item - b * per_cloud;
    const int at = rem / p.s_splits;
    const int ss = rem - at * p.s_splits;
    const int stage0 = ss * p.stages_per_split;
    const int stage_end = min(total_stages, stage0 + p.stages_per_split);
    const float* __restrict__ Ab = p.A + (size_t)b * p.NA * 3;
    const float* __restrict__ Sb = p.S + (size_t)b * p.NS * 3;

    // ---- this thread's QT consecutive A points -> QT/2 f32x2 pairs per coordinate -----------------------------
    const int a0 = at * kATile + tid * QT;
    u64 qx[kPairs], qy[kPairs], qz[kPairs];
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      float c[2][3];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int a = a0 + 2 * i + h;
        if (a < p.NA) {
          c[h][0] = __ldg(Ab + 3 * (size_t)a + 0);
          c[h][1] = __ldg(Ab + 3 * (size_t)a + 1);
          c[h][2] = __ldg(Ab + 3 * (size_t)a + 2);
        } else {
          c[h][0] = c[h][1] = c[h][2] = -kPadCoord();
        }
      }
      qx[i] = pack2(c[0][0], c[1][0]);
      qy[i] = pack2(c[0][1], c[1][1]);
      qz[i] = pack2(c[0][2], c[1][2]);
    }
    float m[QT];
#pragma unroll
    for (int i = 0; i < QT; ++i) {
      m[i] = kInf;
      sm.bestc[i][tid] = 0u;
    }
#pragma unroll
    for (int i = 0; i < QT / 4; ++i) sm.prev[i][tid] = make_float4(kInf, kInf, kInf, kInf);

    {  // prologue: first stage -> buffer 0
      float v[2][3];
      stage_fill_load(Sb, p.NS, stage0 * kSStage, tid, v);
      stage_fill_store<QT>(sm, 0, tid, v);
    }
    __syncthreads();

    for (int st = stage0; st < stage_end; ++st) {
      const int buf = (st - stage0) & 1;
      const int s_base = st * kSStage;
      float nxt[2][3];
      const bool has_next = (st + 1 < stage_end);
      if (has_next) stage_fill_load(Sb, p.NS, s_base + kSStage, tid, nxt);

      // number of 32-point chunks that hold at least one real point
      const int n_valid = min(kSStage, p.NS - s_base);
      const int n_chunks = (n_valid + kChunk - 1) / kChunk;
      const ulonglong2* __restrict__ sxy = sm.sxy[buf];
      const ulonglong2* __restrict__ szz = sm.szz[buf];

      for (int c = 0; c < n_chunks; ++c) {
#pragma unroll 4
        for (int rr = 0; rr < kChunk; rr += 2) {
          const int r = c * kChunk + rr;
          const ulonglong2 cxy0 = sxy[r];
          const ulonglong2 cxy1 = sxy[r + 1];
          const ulonglong2 czz = szz[r >> 1];
          float c0 = kInf, c1 = kInf;
#pragma unroll
          for (int i = 0; i < kPairs; ++i) {
            const u64 d = sqdist2(qx[i], qy[i], qz[i], cxy0.x, cxy0.y, czz.x);
            const u64 e = sqdist2(qx[i], qy[i], qz[i], cxy1.x, cxy1.y, czz.y);
            float dl, dh, el, eh;
            unpack2(d, dl, dh);
            unpack2(e, el, eh);
            m[2 * i] = fmin3(m[2 * i], dl, el);
            m[2 * i + 1] = fmin3(m[2 * i + 1], dh, eh);
            c0 = fmin3(c0, dl, dh);
            c1 = fmin3(c1, el, eh);
          }
          const unsigned u0 = __float_as_uint(c0), u1 = __float_as_uint(c1);
          const unsigned w0 = redux_min_u32(u0);
          const unsigned w1 = redux_min_u32(u1);
          const unsigned b0 = __ballot_sync(0xffffffffu, u0 == w0);
          const unsigned b1 = __ballot_sync(0xffffffffu, u1 == w1);
          if (lane == 0) sm.col[warp][r >> 1] = make_uint4(w0, b0, w1, b1);
        }
        // chunk boundary: which running minima strictly decreased inside this chunk?
        const unsigned gchunk = (unsigned)((s_base >> 5) + c);
#pragma unroll
        for (int i = 0; i < QT / 4; ++i) {
          const float4 pv = sm.prev[i][tid];
          if (m[4 * i + 0] < pv.x) sm.bestc[4 * i + 0][tid] = gchunk;
          if (m[4 * i + 1] < pv.y) sm.bestc[4 * i + 1][tid] = gchunk;
          if (m[4 * i + 2] < pv.z) sm.bestc[4 * i + 2][tid] = gchunk;
          if (m[4 * i + 3] < pv.w) sm.bestc[4 * i + 3][tid] = gchunk;
          sm.prev[i][tid] = make_float4(m[4 * i + 0], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]);
        }
      }

      if (has_next) stage_fill_store<QT>(sm, buf ^ 1, tid, nxt);
      __syncthreads();  // col[][] of this stage complete; next stage's points visible

      {
        // fold the 8 warps' column minima (lowest warp wins ties -> lowest A index) and publish
        const uint2* colp = reinterpret_cast<const uint2*>(&sm.col[0][0]);
#pragma unroll
        for (int h = 0; h < kSStage / kThreads; ++h) {
          const int r = tid + h * kThreads;
          if (r < n_chunks * kChunk) {
            uint2 best = colp[r];
            int bw = 0;
#pragma unroll
            for (int w = 1; w < kWarps; ++w) {
              uint2 v = colp[w * kSStage + r];
              if (v.x < best.x) {
                best = v;
                bw = w;
              }
            }
            const int s = s_base + r;
            if (s < p.NS) {
              const unsigned gid = (unsigned)(at * kThreads + bw * 32 + (__ffs(best.y) - 1));
              atomicMin(p.PS + (size_t)b * p.NS + s, ((u64)best.x << 32) | gid);
            }
          }
        }
        __syncthreads();  // col[][] may be overwritten by the next stage
      }
    }

    // ---- publish row minima ------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < QT; ++i) {
      const int a = a0 + i;
      if (a < p.NA) atomicMin(p.PA + (size_t)b * p.NA + a, pack_key(m[i], sm.bestc[i][tid]));
    }
  }
}

// Index resolution + unpack, warp-cooperative: the lanes of a warp evaluate the candidates of one point in
// parallel (coalesced candidate reads), a ballot of the exact matches gives the first (lowest-index) one.
//   A side: candidates are the 32 streamed points of chunk `id` (one point per warp step);
//   S side: the `group` register-side points of group `id` (32/group points per warp step).
// The candidate distance is evaluated with the identical instruction sequence as the search, so exactly-equal
// compares are well defined.
__global__ void __launch_bounds__(256) chamfer_resolve_kernel(const float* __restrict__ A, const float* __restrict__ S,
                                                              int B, int NA, int NS, int group,
                                                              const u64* __restrict__ PA, const u64* __restrict__ PS,
                                                              float* __restrict__ distA, int32_t* __restrict__ idxA,
                                                              float* __restrict__ distS, int32_t* __restrict__ idxS) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long totalA = (long long)B * NA;
  const long long totalS = (long long)B * NS;

  // ---- A side: 32 points per warp iteration, one candidate chunk per point -------------------------------------
  // The candidate loads of kRU consecutive points are issued together so kRU gathers are in flight per lane.
  constexpr int kRU = 4;
  if (distA != nullptr) {
    for (long long t0 = warp_global * 32; t0 < totalA; t0 += n_warps * 32) {
      const long long t = t0 + lane;
      u64 key = 0ull;
      float ax = 0.f, ay = 0.f, az = 0.f;
      int bl = 0;
      if (t < totalA) {
        key = PA[t];
        ax = A[3 * t];
        ay = A[3 * t + 1];
        az = A[3 * t + 2];
        bl = (int)(t / NA);
      }
      int myidx = 0;
#pragma unroll 1
      for (int j0 = 0; j0 < 32; j0 += kRU) {
        if (t0 + j0 >= totalA) break;  // warp-uniform
        float cx[kRU], cy[kRU], cz[kRU];
        bool ok[kRU];
        unsigned kd[kRU];
        int base[kRU];
#pragma unroll
        for (int u = 0; u < kRU; ++u) {
          const int j = j0 + u;
          const u64 kj = __shfl_sync(0xffffffffu, key, j);
          const int b = __shfl_sync(0xffffffffu, bl, j);
          kd[u] = (unsigned)(kj >> 32);
          // a key that was never lowered (all-ones: every distance of this point was NaN - outside the contract)
          // must not be turned into an address
          const bool valid_key = kj != ~0ull;
          base[u] = valid_key ? (int)(unsigned)kj * kChunk : 0;
          const int cand = base[u] + lane;
          ok[u] = valid_key && (t0 + j < totalA) && (cand < NS);
          cx[u] = cy[u] = cz[u] = 0.f;
          if (ok[u]) {
            const float* c = S + ((size_t)b * NS + cand) * 3;
            cx[u] = __ldg(c);
            cy[u] = __ldg(c + 1);
            cz[u] = __ldg(c + 2);
          }
        }
#pragma unroll
        for (int u = 0; u < kRU; ++u) {
          const int j = j0 + u;
          const float x = __shfl_sync(0xffffffffu, ax, j);
          const float y = __shfl_sync(0xffffffffu, ay, j);
          const float z = __shfl_sync(0xffffffffu, az, j);
          const bool eq = ok[u] && (sqdist(x, y, z, cx[u], cy[u], cz[u]) == __uint_as_float(kd[u]));
          const unsigned hits = __ballot_sync(0xffffffffu, eq);
          if (lane == j) myidx = base[u] + (hits ? (__ffs(hits) - 1) : 0);
        }
      }
      if (t < totalA) {
        distA[t] = __uint_as_float((unsigned)(key >> 32));
        idxA[t] = myidx;
      }
    }
  }
  // ---- S side: `group` lanes per point, 32/group points per step ----------------------------------------------
  if (distS != nullptr) {
    const int per_step = 32 / group;  // group is 8 or 16
    const int sub = lane / group, off = lane % group;
    const unsigned gmask = (1u << group) - 1u;
    for (long long u0 = warp_global * 32; u0 < totalS; u0 += n_warps * 32) {
      const long long uu = u0 + lane;
      u64 key = 0ull;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      int bl = 0;
      if (uu < totalS) {
        key = PS[uu];
        sx = S[3 * uu];
        sy = S[3 * uu + 1];
        sz = S[3 * uu + 2];
        bl = (int)(uu / NS);
      }
      int myidx = 0;
#pragma unroll 1
      for (int j0 = 0; j0 < 32; j0 += kRU * per_step) {
        if (u0 + j0 >= totalS) break;  // warp-uniform
        float cx[kRU], cy[kRU], cz[kRU];
        bool ok[kRU];
        unsigned kd[kRU];
#pragma unroll
        for (int q = 0; q < kRU; ++q) {
          const int j = j0 + q * per_step + sub;  // the point this lane helps with
          const u64 kj = __shfl_sync(0xffffffffu, key, j & 31);
          const int b = __shfl_sync(0xffffffffu, bl, j & 31);
          kd[q] = (unsigned)(kj >> 32);
          const bool valid_key = kj != ~0ull;
          const int cand = valid_key ? (int)(unsigned)kj * group + off : 0;
          ok[q] = valid_key && (j < 32) && (u0 + j < totalS) && (cand < NA);
          cx[q] = cy[q] = cz[q] = 0.f;
          if (ok[q]) {
            const float* c = A + ((size_t)b * NA + cand) * 3;
            cx[q] = __ldg(c);
            cy[q] = __ldg(c + 1);
            cz[q] = __ldg(c + 2);
          }
        }
#pragma unroll
        for (int q = 0; q < kRU; ++q) {
          const int jb = j0 + q * per_step;  // first point of this step
          const int j = jb + sub;
          const float x = __shfl_sync(0xffffffffu, sx, j & 31);
          const float y = __shfl_sync(0xffffffffu, sy, j & 31);
          const float z = __shfl_sync(0xffffffffu, sz, j & 31);
          const bool eq = ok[q] && (sqdist(cx[q], cy[q], cz[q], x, y, z) == __uint_as_float(kd[q]));
          const unsigned hits = __ballot_sync(0xffffffffu, eq);
          const int owner_sub = lane - jb;  // lane jb + s owns sub-group s of this step
          if (owner_sub >= 0 && owner_sub < per_step) {
            const unsigned mine = (hits >> (owner_sub * group)) & gmask;
            myidx = (key != ~0ull) ? (int)(unsigned)key * group + (mine ? (__ffs(mine) - 1) : 0) : 0;
          }
        }
      }
      if (uu < totalS) {
        distS[uu] = __uint_as_float((unsigned)(key >> 32));
        idxS[uu] = myidx;
      }
    }
  }
}

// ---- optional per-launch timing of chamfer_pair_kernel (bench.py's roofline.achieved) ------------------------
// CUDA events are recorded on the launch stream immediately around the kernel; nothing is synchronised until collect.
static bool g_profile = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;

void profile_enable(bool on) {
  for (auto& e : g_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  g_events.clear();
  g_profile = on;
}

int profile_collect(double* total_ms, int* launches) {
  double tot = 0.0;
  int n = 0;
  for (auto& e : g_events) {
    cudaError_t err = cudaEventSynchronize(e.second);
    if (err != cudaSuccess) return (int)err;
    float ms = 0.f;
    if ((err = cudaEventElapsedTime(&ms, e.first, e.second)) != cudaSuccess) return (int)err;
    tot += ms;
    ++n;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return 0;
}

size_t chamfer_forward_workspace_bytes(int B, int N, int M) {
  return sizeof(u64) * ((size_t)B * N + (size_t)B * M) + 256;
}

// A-side points per thread.  QT=16 halves the per-pair LDS/REDUX overhead but makes 4096-point tiles: it is used only
// when even one-stage items of that size would still oversubscribe the persistent grid 4x (large batches); otherwise
// QT=8 (2048-point tiles).  IAPT_FWD_VARIANT=8|16 (development knob, read once) forces one.
// Measured at B=48 x 16384^2 (all parity-green): QT=8 3.22 ms, QT=16 3.09 ms; also tried and dropped for no gain:
// register prefetch of the next streamed pair, VIMNMX3 integer minima, a static contiguous schedule (3.17 ms) and
// guided decreasing-size tickets (3.09 ms) - see DESIGN.md section 3.1.
static int pick_qt(int B, int NA, int NS, int slots) {
  static int forced = -1;
  if (forced < 0) {
    forced = 0;
    if (const char* e = std::getenv("IAPT_FWD_VARIANT")) {
      const int q = std::atoi(e);
      if (q == 8 || q == 16) forced = q;
    }
  }
  if (forced > 0) return forced;
  const long long tiles16 = (NA + kThreads * 16 - 1) / (kThreads * 16);
  const long long stages = (NS + kSStage - 1) / kSStage;
  return (NA >= 4096 && (long long)B * tiles16 * stages >= 4LL * slots) ? 16 : 8;
}

template <int QT>
static cudaError_t launch_pair(const FwdParams& p, int grid, cudaStream_t stream) {
  static PerDeviceOnce once;
  if (once.pending()) {
    cudaFuncSetAttribute(chamfer_pair_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem<QT>));
    once.done();
  }
  chamfer_pair_kernel<QT><<<grid, kThreads, sizeof(FwdSmem<QT>), stream>>>(p);
  return cudaGetLastError();
}

// Host launcher.  Returns cudaError_t.
int chamfer_forward_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2,
                           int32_t* idx1, int32_t* idx2, void* workspace, cudaStream_t stream) {
  // Register side = the larger cloud (its tile is 256*QT points; the streamed side pads only to 32).
  const bool swap = N < M;
  const float* A = swap ? xyz2 : xyz1;
  const float* S = swap ? xyz1 : xyz2;
  const int NA = swap ? M : N, NS = swap ? N : M;
  float* distA = swap ? dist2 : dist1;
  float* distS = swap ? dist1 : dist2;
  int32_t* idxA = swap ? idx2 : idx1;
  int32_t* idxS = swap ? idx1 : idx2;

  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const size_t key_bytes = sizeof(u64) * ((size_t)B * NA + (size_t)B * NS);
  u64* PA = reinterpret_cast<u64*>(ws);
  u64* PS = PA + (size_t)B * NA;
  unsigned* ticket = reinterpret_cast<unsigned*>(ws + key_bytes);

  cudaError_t err;
  // one memset: keys = +inf-like all-ones, ticket = 0xFFFFFFFF (the kernel uses atomicAdd(ticket,1)+1)
  if ((err = cudaMemsetAsync(ws, 0xFF, key_bytes + 256, stream)) != cudaSuccess) return (int)err;

  static int sms_of_device[256] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int num_sms = sms_of_device[dev & 255];
  if (num_sms == 0) {
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
    sms_of_device[dev & 255] = num_sms;
  }
  const int slots = num_sms * 2;
  const int qt = pick_qt(B, NA, NS, slots);
  const int a_tile = kThreads * qt;

  FwdParams p;
  p.A = A;
  p.S = S;
  p.B = B;
  p.NA = NA;
  p.NS = NS;
  p.PA = PA;
  p.PS = PS;
  p.ticket = ticket;
  p.a_tiles = (NA + a_tile - 1) / a_tile;
  const int total_stages = (NS + kSStage - 1) / kSStage;
  const long long base_items = (long long)B * p.a_tiles;
  // Split the streamed side until there are >= 16 items per resident CTA slot (tail < ~3 %); an item never gets
  // less than one 512-point stage.
  const long long items_per_slot = 16;  // measured: 8 -> 3.20 ms, 16 -> 3.09 ms, 32/64 -> 3.23 ms at B=48 x 16384^2
  int splits = 1;
  if (base_items < items_per_slot * slots) {
    splits = (int)((items_per_slot * slots + base_items - 1) / base_items);
    if (splits > total_stages) splits = total_stages;
    if (splits < 1) splits = 1;
  }
  p.stages_per_split = (total_stages + splits - 1) / splits;
  p.s_splits = (total_stages + p.stages_per_split - 1) / p.stages_per_split;
  p.total_items = (int)(base_items * p.s_splits);

  const int grid = (int)((long long)p.total_items < (long long)slots ? p.total_items : slots);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_profile) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  err = qt == 16 ? launch_pair<16>(p, grid, stream) : launch_pair<8>(p, grid, stream);
  count_launch();
  if (g_profile) {
    cudaEventRecord(ev1, stream);
    g_events.emplace_back(ev0, ev1);
  }
  if (err != cudaSuccess) return (int)err;

  const long long pts = (long long)B * NA + (long long)B * NS;
  long long rblocks = (pts / 32 + 7) / 8;  // one warp per 32 points
  if (rblocks > num_sms * 16LL) rblocks = num_sms * 16LL;
  if (rblocks < 1) rblocks = 1;
  chamfer_resolve_kernel<<<(int)rblocks, 256, 0, stream>>>(A, S, B, NA, NS, qt, PA, PS, distA, idxA, distS, idxS);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
