// chamfer_fwd.cu - bidirectional brute-force nearest-neighbour search for sm_100a.
//
// Replaces chamfer_dist_kernel x2 (reference extensions/chamfer_dist/chamfer.cu:15-145, launched :159-164).
//
// Design:
//  * ONE pass evaluates every unordered (a, s) pair once and feeds BOTH directions: the row minimum (nearest
//    streamed point for each register point) and the column minimum (nearest register point for each streamed
//    point).  The reference evaluates each pair twice (two launches with the roles swapped).
//  * "A side" points live in registers, 8 per thread, packed two-per-f32x2 so the 3 SUB + 1 MUL + 2 FMA of the
//    direct form issue as FADD2/FMUL2/FFMA2 (half the FP32 issue slots).  "S side" points are streamed through
//    double-buffered shared memory in a duplicated (x,x,y,y | z,z) layout and read with broadcast LDS.128/LDS.64.
//  * No per-pair argmin bookkeeping.  Only minima are tracked in the hot loop (FMNMX3: one ALU slot per two
//    pairs per direction); the index is recovered exactly afterwards:
//      rows   : per 32-point chunk "did my running minimum strictly decrease?" -> remember the FIRST such chunk;
//      columns: REDUX.MIN over the warp + ballot of the lanes that hold the minimum -> lowest 8-point group;
//    both are folded across tiles/CTAs with a 64-bit atomicMin on (distance bits << 32 | coarse id), whose
//    unsigned order is the lexicographic (distance, id) order = the reference's lowest-index tie rule
//    (chamfer.cu:47-50,137-140).  chamfer_resolve_kernel then re-evaluates the <= 32 candidates of the winning
//    chunk/group with the SAME rounding and returns the first exact match.
//  * Work items = (cloud, 2048-point A tile, S split), handed to a persistent grid (2 CTAs per SM) through an
//    atomic ticket so B = 48 does not quantise badly over 148 SMs (the reference strides batches over gridDim.x=32).
#include <utility>
#include <vector>

#include "common.cuh"

namespace iapt {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kQT = 8;                  // A-side points per thread
constexpr int kATile = kThreads * kQT;  // 2048
constexpr int kSStage = 512;            // streamed points per smem stage
constexpr int kChunk = 32;              // streamed points per row-index chunk
constexpr int kGroup = kQT;             // A-side points per column-index group (= one thread's points)

struct FwdParams {
  const float* A;  // [B, NA, 3] register side
  const float* S;  // [B, NS, 3] streamed side
  int B, NA, NS;
  u64* PA;  // [B, NA] packed (dist, chunk id of S)
  u64* PS;  // [B, NS] packed (dist, group id of A)
  unsigned* ticket;
  int a_tiles, s_splits, stages_per_split, total_items;
};

struct __align__(16) FwdSmem {
  ulonglong2 sxy[2][kSStage];         // (x,x),(y,y)
  u64 sz[2][kSStage];                 // (z,z)
  uint4 col[kWarps][kSStage / 2];     // per warp, per streamed point pair: (min bits, ballot) x2
  unsigned next_item;
};

__device__ __forceinline__ void stage_fill_load(const float* __restrict__ Sb, int NS, int base, int tid, float (&v)[2][3]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int s = base + tid + h * kThreads;
    if (s < NS) {
      v[h][0] = __ldg(Sb + 3 * (size_t)s + 0);
      v[h][1] = __ldg(Sb + 3 * (size_t)s + 1);
      v[h][2] = __ldg(Sb + 3 * (size_t)s + 2);
    } else {
      v[h][0] = v[h][1] = v[h][2] = kPadCoord();
    }
  }
}

__device__ __forceinline__ void stage_fill_store(FwdSmem& sm, int buf, int tid, const float (&v)[2][3]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int r = tid + h * kThreads;
    ulonglong2 xy;
    xy.x = pack2(v[h][0], v[h][0]);
    xy.y = pack2(v[h][1], v[h][1]);
    sm.sxy[buf][r] = xy;
    sm.sz[buf][r] = pack2(v[h][2], v[h][2]);
  }
}

template <bool kCols>
__global__ void __launch_bounds__(kThreads, 2) chamfer_pair_kernel(FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const float kInf = __int_as_float(0x7f800000);

  for (;;) {
    __syncthreads();  // everyone is done with smem (incl. next_item) of the previous item
    if (tid == 0) sm.next_item = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const int item = (int)sm.next_item;
    if (item >= p.total_items) break;

    const int per_cloud = p.a_tiles * p.s_splits;
    const int b = item / per_cloud;
    const int rem = item - b * per_cloud;
    const int at = rem / p.s_splits;
    const int ss = rem - at * p.s_splits;
    const float* __restrict__ Ab = p.A + (size_t)b * p.NA * 3;
    const float* __restrict__ Sb = p.S + (size_t)b * p.NS * 3;

    // ---- this thread's 8 consecutive A points -> 4 f32x2 pairs per coordinate --------------------------------
    const int a0 = at * kATile + tid * kQT;
    u64 qx[kQT / 2], qy[kQT / 2], qz[kQT / 2];
#pragma unroll
    for (int i = 0; i < kQT / 2; ++i) {
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
    float m[kQT];
    unsigned bestc[kQT];
#pragma unroll
    for (int i = 0; i < kQT; ++i) {
      m[i] = kInf;
      bestc[i] = 0u;
    }

    const int stage0 = ss * p.stages_per_split;
    const int total_stages = (p.NS + kSStage - 1) / kSStage;
    const int stage_end = min(total_stages, stage0 + p.stages_per_split);

    {  // prologue: first stage -> buffer 0
      float v[2][3];
      stage_fill_load(Sb, p.NS, stage0 * kSStage, tid, v);
      stage_fill_store(sm, 0, tid, v);
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
      const u64* __restrict__ sz = sm.sz[buf];

      for (int c = 0; c < n_chunks; ++c) {
        float prev[kQT];
#pragma unroll
        for (int i = 0; i < kQT; ++i) prev[i] = m[i];
#pragma unroll 4
        for (int rr = 0; rr < kChunk; rr += 2) {
          const int r = c * kChunk + rr;
          const ulonglong2 xy0 = sxy[r];
          const ulonglong2 xy1 = sxy[r + 1];
          const u64 z0 = sz[r];
          const u64 z1 = sz[r + 1];
          float dl[kQT / 2], dh[kQT / 2], el[kQT / 2], eh[kQT / 2];
#pragma unroll
          for (int i = 0; i < kQT / 2; ++i) {
            u64 d = sqdist2(qx[i], qy[i], qz[i], xy0.x, xy0.y, z0);
            u64 e = sqdist2(qx[i], qy[i], qz[i], xy1.x, xy1.y, z1);
            unpack2(d, dl[i], dh[i]);
            unpack2(e, el[i], eh[i]);
            m[2 * i] = fmin3(m[2 * i], dl[i], el[i]);
            m[2 * i + 1] = fmin3(m[2 * i + 1], dh[i], eh[i]);
          }
          if (kCols) {
            float c0 = fmin3(dl[0], dh[0], dl[1]);
            c0 = fmin3(c0, dh[1], dl[2]);
            c0 = fmin3(c0, dh[2], dl[3]);
            c0 = fminf(c0, dh[3]);
            float c1 = fmin3(el[0], eh[0], el[1]);
            c1 = fmin3(c1, eh[1], el[2]);
            c1 = fmin3(c1, eh[2], el[3]);
            c1 = fminf(c1, eh[3]);
            const unsigned u0 = __float_as_uint(c0), u1 = __float_as_uint(c1);
            const unsigned w0 = redux_min_u32(u0);
            const unsigned w1 = redux_min_u32(u1);
            const unsigned b0 = __ballot_sync(0xffffffffu, u0 == w0);
            const unsigned b1 = __ballot_sync(0xffffffffu, u1 == w1);
            if (lane == 0) sm.col[warp][r >> 1] = make_uint4(w0, b0, w1, b1);
          }
        }
        const unsigned gchunk = (unsigned)((s_base >> 5) + c);
#pragma unroll
        for (int i = 0; i < kQT; ++i) bestc[i] = (m[i] < prev[i]) ? gchunk : bestc[i];
      }

      if (has_next) stage_fill_store(sm, buf ^ 1, tid, nxt);
      __syncthreads();  // col[][] of this stage complete; next stage's points visible

      if (kCols) {
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
              const unsigned gid = (unsigned)(at * (kATile / kGroup) + bw * 32 + (__ffs(best.y) - 1));
              atomicMin(p.PS + (size_t)b * p.NS + s, ((u64)best.x << 32) | gid);
            }
          }
        }
        __syncthreads();  // col[][] may be overwritten by the next stage
      }
    }

    // ---- publish row minima ------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < kQT; ++i) {
      const int a = a0 + i;
      if (a < p.NA) atomicMin(p.PA + (size_t)b * p.NA + a, pack_key(m[i], bestc[i]));
    }
  }
}

// Index resolution + unpack.  One thread per point of either side.
//   A side: candidates are the 32 streamed points of chunk `id`;  S side: the 8 register-side points of group `id`.
// The candidate distance is evaluated with the identical instruction sequence as the search, so exactly-equal
// compares are well defined; the first (lowest index) match is the reference's answer.
__global__ void __launch_bounds__(256) chamfer_resolve_kernel(const float* __restrict__ A, const float* __restrict__ S,
                                                              int B, int NA, int NS, const u64* __restrict__ PA,
                                                              const u64* __restrict__ PS, float* __restrict__ distA,
                                                              int32_t* __restrict__ idxA, float* __restrict__ distS,
                                                              int32_t* __restrict__ idxS) {
  const long long totalA = (long long)B * NA;
  const long long total = totalA + (long long)B * NS;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    if (t < totalA) {
      if (distA == nullptr) continue;
      const int b = (int)(t / NA);
      const u64 key = PA[t];
      const float d = __uint_as_float((unsigned)(key >> 32));
      const int c0 = (int)(unsigned)key * kChunk;
      const float ax = A[3 * t], ay = A[3 * t + 1], az = A[3 * t + 2];
      const float* Sb = S + (size_t)b * NS * 3;
      const int c1 = min(NS, c0 + kChunk);
      int found = c0;
      for (int j = c0; j < c1; ++j) {
        const float e = sqdist(ax, ay, az, Sb[3 * j], Sb[3 * j + 1], Sb[3 * j + 2]);
        if (e == d) {
          found = j;
          break;
        }
      }
      distA[t] = d;
      idxA[t] = found;
    } else {
      if (distS == nullptr) continue;
      const long long u = t - totalA;
      const int b = (int)(u / NS);
      const u64 key = PS[u];
      const float d = __uint_as_float((unsigned)(key >> 32));
      const int g0 = (int)(unsigned)key * kGroup;
      const float sx = S[3 * u], sy = S[3 * u + 1], sz = S[3 * u + 2];
      const float* Ab = A + (size_t)b * NA * 3;
      const int g1 = min(NA, g0 + kGroup);
      int found = g0;
      for (int j = g0; j < g1; ++j) {
        const float e = sqdist(Ab[3 * j], Ab[3 * j + 1], Ab[3 * j + 2], sx, sy, sz);
        if (e == d) {
          found = j;
          break;
        }
      }
      distS[u] = d;
      idxS[u] = found;
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

// Host launcher.  Returns cudaError_t.
int chamfer_forward_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2,
                           int32_t* idx1, int32_t* idx2, void* workspace, cudaStream_t stream) {
  // Register side = the larger cloud (its tile is 2048 points; the streamed side pads only to 32).
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
  if ((err = cudaMemsetAsync(ws, 0xFF, key_bytes, stream)) != cudaSuccess) return (int)err;
  if ((err = cudaMemsetAsync(ticket, 0, 256, stream)) != cudaSuccess) return (int)err;

  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
    cudaFuncSetAttribute(chamfer_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem));
    cudaFuncSetAttribute(chamfer_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem));
  }

  FwdParams p;
  p.A = A;
  p.S = S;
  p.B = B;
  p.NA = NA;
  p.NS = NS;
  p.PA = PA;
  p.PS = PS;
  p.ticket = ticket;
  p.a_tiles = (NA + kATile - 1) / kATile;
  const int total_stages = (NS + kSStage - 1) / kSStage;
  const int slots = num_sms * 2;
  const long long base_items = (long long)B * p.a_tiles;
  // Split the streamed side until there are >= 16 items per resident CTA slot (tail < ~3 %), but keep >= 2 stages
  // (1024 streamed points) per item so the per-item prologue stays negligible.
  int splits = 1;
  if (base_items < 16LL * slots) {
    splits = (int)((16LL * slots + base_items - 1) / base_items);
    const int max_splits = (total_stages + 1) / 2;
    if (splits > max_splits) splits = max_splits;
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
  chamfer_pair_kernel<true><<<grid, kThreads, sizeof(FwdSmem), stream>>>(p);
  count_launch();
  if (g_profile) {
    cudaEventRecord(ev1, stream);
    g_events.emplace_back(ev0, ev1);
  }
  if ((err = cudaGetLastError()) != cudaSuccess) return (int)err;

  const long long pts = (long long)B * NA + (long long)B * NS;
  long long rblocks = (pts + 255) / 256;
  if (rblocks > num_sms * 32LL) rblocks = num_sms * 32LL;
  chamfer_resolve_kernel<<<(int)rblocks, 256, 0, stream>>>(A, S, B, NA, NS, PA, PS, distA, idxA, distS, idxS);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
