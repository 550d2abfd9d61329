// chamfer_step.cu - the whole Chamfer step (search, index resolution, loss reduction, gradient) as ONE persistent,
// dependency-driven kernel for sm_100a.
//
// Replaces chamfer_dist_kernel x2 + chamfer_dist_grad_kernel x2 (reference extensions/chamfer_dist/chamfer.cu:15-145
// launched :159-164, and :173-201 launched :215-222) and the torch.sqrt / torch.mean epilogues of the loss modules
// (extensions/chamfer_dist/__init__.py:43-44,61-62,82-84,101-103).
//
// Work is cut into ITEMS handed out by one atomic ticket to a persistent grid (2 CTAs per SM).  Items are ordered by
// phase and, inside a phase, by cloud; an item only ever waits for items with SMALLER tickets (per-cloud completion
// counters), so whatever the residency of the grid the scheme cannot deadlock:
//
//   PAIR     (cloud, A tile of 256*QT points, range of 32-point S chunks)   the O(N*M) search, DESIGN.md section 3.1
//   RESOLVE  (cloud, side, rc points)   waits PAIR(cloud): exact index recovery, dist/idx write-out, per-32-point
//                                       partial sums (sum d, sum sqrt d, #sqrt d < th), fan-in COUNT of the neighbour
//   PREFIX   (cloud, side)              waits RESOLVE(cloud): exclusive scan of the fan-in counts -> CSR offsets
//   FILL     (cloud, side, rc points)   waits PREFIX(cloud): every point drops its index into its neighbour's segment
//   GATHER   (cloud, side, rc points)   waits FILL(cloud): own term + fan-in segment summed in ASCENDING source index
//
// The CTA that completes the last RESOLVE item of a cloud folds that cloud's partials in a fixed order (per-cloud sums,
// F-Score counts); the one that completes the last cloud forms the loss scalars.  So a fused loss step is one memset
// (0xFF over keys, counters and counts - every counter is read as value+1) and one kernel, instead of the reference's
// 4 memsets + 4 kernels + the autograd chain of sqrt/mean, and no phase leaves the machine idle while another drains.
//
// PAIR (unchanged arithmetic): every unordered (a, s) pair is evaluated once and feeds both
// directions; A points live in registers (QT per thread, packed f32x2: FADD2/FMUL2/FFMA2), S points stream through
// double-buffered shared memory read with broadcast LDS.128; only minima are tracked (FMNMX3) and the index is
// recovered exactly afterwards from (first 32-point chunk in which the row minimum strictly decreased | lowest lane
// group holding the column minimum), folded across items by a 64-bit atomicMin on (distance bits << 32 | coarse id)
// = the reference's lowest-index tie rule (chamfer.cu:47-50,137-140).
//
// GATHER is deterministic without floating-point atomics: a CSR segment of <= 8 entries (mean fan-in is ~1) is ordered
// by a sorting network in registers and summed in ascending source index; a longer one is summed EXACTLY in 64-bit
// fixed point (scale = a power of two from the segment's largest term), which needs no order at all - so results do
// not depend on scheduling or launch geometry, and a collapsed prediction (thousands of points sharing one neighbour)
// is a parallel reduction instead of the reference's atomicAdd contention (chamfer.cu:193-198).
#include <cmath>
#include <cstdlib>
#include <utility>
#include <vector>

#include "common.cuh"
#include "chamfer_step.h"

namespace iapt {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSStage = 512;                    // streamed points per smem stage
constexpr int kChunk = 32;                      // streamed points per row-index chunk
constexpr int kStageChunks = kSStage / kChunk;  // 16
constexpr int kWarpSeg = 1024;                  // fan-in entries a single warp sums; longer segments take the CTA
constexpr int kMaxRc = 2048;                    // points per RESOLVE / FILL / GATHER item (upper bound)

struct StepParams {
  // ---- geometry: side 0 = xyz1, side 1 = xyz2; the register ("A") side of the pair phase is xyz[reg] -------------
  const float* xyz[2];
  int n[2];
  int B;
  int reg;
  int mask_zero_sum;  // points with x+y+z == 0 neither query nor answer (ignore_zeros, chamfer_dist/__init__.py:37-41)
  // ---- search state ------------------------------------------------------------------------------------------
  u64* key[2];       // [B, n[s]] packed (distance bits, coarse id); all-ones = never lowered
  float* dist[2];    // [B, n[s]]   (a direction may be skipped in search-only launches: both null)
  int32_t* idx[2];
  // ---- counters: every one starts at 0xFFFFFFFF and is read as value + 1 -----------------------------------------
  unsigned* ticket;
  unsigned* done_pair;   // [B]
  unsigned* done_res;    // [B]
  unsigned* done_pre;    // [B]
  unsigned* done_fill;   // [B]
  unsigned* done_final;  // clouds whose partials were folded
  // ---- item layout -----------------------------------------------------------------------------------------------
  int a_tiles, s_splits, total_chunks;  // PAIR item ss of a tile streams chunks [ss*total/s_splits, (ss+1)*total/s_splits)
  int rc[2];            // points of side s per RESOLVE / FILL / GATHER item (multiple of 32)
  int rc_items[2];      // ceil(n[s] / rc[s])
  int pair_per_cloud, rc_per_cloud;
  // Ticket order: STEPS t = 0, 1, ...; step t holds, in phase order, the items of phase ph for cloud t - lag[ph]
  // (when that is a cloud).  lag[] is non-decreasing, so an item only waits for smaller tickets.
  int per_cloud[5];     // items per cloud of PAIR, RESOLVE, PREFIX, FILL, GATHER (0: phase absent)
  int lag[5];
  int n_steps;          // B + lag[4]
  int total_items;
  // ---- loss / metric epilogue ----------------------------------------------------------------------------------------
  float4* part[2];      // [B, ceil(n[s]/32)] {sum d, sum sqrt d, count(sqrt d < th) as int bits, #unmasked as int bits}
  float th;
  float* sums;          // [B,4] {sum d1, sum d2, sum sqrt d1, sum sqrt d2} or null
  int32_t* counts;      // [B,2] or null
  int32_t* valid;       // [B,2] points with a non-zero coordinate sum per cloud and side, or null
  int valid_is_input;   // gate: `valid` comes from an earlier launch; if it equals (n[0], n[1]) everywhere, do nothing
  float* loss;          // [2] or null
  int kind;             // IAPT_LOSS_* (0 = none)
  // ---- gradient ------------------------------------------------------------------------------------------------------
  int mode;             // 0: weights from gw = grad_dist;  1: scale/sqrt(gw = dist);  2: scale
  float scale[2];
  const float* grad_out;  // device upstream scalar(s) or null (= 1)
  int grad_out_stride;
  const float* gw[2];   // per-point weight source of side s (null: side s contributes nothing)
  float* grad[2];       // [B, n[s], 3] or null
  int need_list[2];     // fan-in list of side s is built (grad[s] wanted and side 1-s contributes)
  int* cnt[2];          // [B, n[s]]      fan-in count - 1 (starts at -1), consumed again by FILL
  int* off[2];          // [B, n[s] + 1]  CSR offsets
  int* csr[2];          // [B, n[1-s]]    source indices grouped by neighbour
  u64* stats;           // optional (globaltimer ns): [0] min start of PAIR items, [1] of the kernel, [2] max end of PAIR
                        // items, [3] of the kernel, [4 + ph] max end of the items of phase ph
};

template <int QT>
struct __align__(16) PairSmem {
  ulonglong2 sxy[2][kSStage];       // (x,x),(y,y)
  ulonglong2 szz[2][kSStage / 2];   // (z,z) of two consecutive points
  uint4 col[kWarps][kSStage / 2];   // per warp, per streamed point pair: (min bits, ballot) x2
  float4 prev[QT / 4][kThreads];    // running minima at the last chunk boundary
  unsigned bestc[QT][kThreads];     // first chunk in which the running minimum strictly decreased
  float sraw[kSStage * 3];          // the NEXT stage's [512,3] rows as they lie in global memory (cp.async landing zone)
};
template <>
struct __align__(16) PairSmem<0> {
  int unused;
};

struct __align__(16) TailSmem {
  long long redl[3 * kWarps];
  unsigned short deferred[kMaxRc];   // points of this item whose segment (> 8 entries) needs a cooperative pass
  unsigned short medium[kMaxRc];     // points of this item with 3..8 entries
  float redf[32];
  int redi[32];
  int wcnt[kWarps];
  int n_deferred, n_medium;
};

struct CtlSmem {
  int ph, b, q;  // next item: phase (-1: none left), cloud, index inside (phase, cloud)
  int flag;
};

template <int QT>
struct StepSmem {
  static constexpr size_t kCtlOff = sizeof(PairSmem<QT>) > sizeof(TailSmem) ? sizeof(PairSmem<QT>) : sizeof(TailSmem);
  static constexpr size_t kBytes = kCtlOff + sizeof(CtlSmem);
};

// ---- small device helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 globaltimer_ns() {
  u64 t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Thread 0 spins until `count` increments of *ctr (which started at 0xFFFFFFFF) are visible, then the CTA proceeds.
__device__ __forceinline__ void wait_done(const unsigned* ctr, int count) {
  if (threadIdx.x == 0 && count > 0) {
    unsigned ns = 20;
    u64 t0 = 0;
    while (ld_acquire_u32(ctr) != (unsigned)(count - 1)) {
      __nanosleep(ns);
      if (ns < 640) {
        ns <<= 1;
      } else {  // a producer always holds a smaller ticket and is running: waiting this long means a broken launch -
        const u64 t = globaltimer_ns();  // fail loudly (launch error) rather than hang the GPU
        if (t0 == 0) t0 = t;
        if (t - t0 > 4000000000ull) __trap();
      }
    }
  }
  __syncthreads();
}
// All of the CTA's global writes of this item -> one increment.  Returns (to every thread) whether this increment was
// the `count`-th, i.e. the item completed its group.
__device__ __forceinline__ bool signal_done(unsigned* ctr, int count, CtlSmem& ctl) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(ctr, 1u);
    ctl.flag = (old + 2u == (unsigned)count) ? 1 : 0;
    if (ctl.flag) __threadfence();  // acquire side of the last-arriver pattern
  }
  __syncthreads();
  return ctl.flag != 0;
}

__device__ __forceinline__ bool zero_sum(float x, float y, float z) { return __fadd_rn(__fadd_rn(x, y), z) == 0.f; }

__device__ __forceinline__ float block_sum_f(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < kWarps) ? sh[threadIdx.x] : 0.f;
  if (warp == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
  }
  return t;  // valid in thread 0
}
__device__ __forceinline__ int block_sum_i(int v, int* sh) {
  v = __reduce_add_sync(0xffffffffu, v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  int t = (threadIdx.x < kWarps) ? sh[threadIdx.x] : 0;
  if (warp == 0) t = __reduce_add_sync(0xffffffffu, t);
  return t;
}

// =====================================================================================================================
// PAIR
// =====================================================================================================================
// The streamed cloud reaches shared memory by cp.async (LDGSTS): the 1536 floats of a stage are copied as they lie, 6
// fully coalesced 4-byte copies per thread, while the previous stage is being evaluated - no register holds them during
// the hot loop (the register-staged version kept 6 live across it).  At the stage boundary the raw rows are expanded
// into the lane-duplicated layout the broadcast LDS.128 of the loop reads.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void stage_issue(const float* __restrict__ Sb, int base, int end, int tid, float* sraw) {
  const float* src = Sb + 3 * (size_t)base;
  const int nfl = 3 * (min(end - base, kSStage));  // floats of this stage that exist
#pragma unroll
  for (int i = 0; i < 3 * kSStage / kThreads; ++i) {
    const int f = tid + i * kThreads;
    if (f < nfl) cp_async4(sraw + f, src + f);
  }
  cp_async_commit();
}

// this thread's two points of the stage, from the landed rows (sentinels beyond the item's range / for masked points)
__device__ __forceinline__ void stage_fill_load(const float* sraw, int base, int end, int tid, bool mask,
                                                float (&v)[2][3]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int s = base + 2 * tid + h;
    if (s < end) {
      v[h][0] = sraw[6 * tid + 3 * h + 0];
      v[h][1] = sraw[6 * tid + 3 * h + 1];
      v[h][2] = sraw[6 * tid + 3 * h + 2];
      if (mask && zero_sum(v[h][0], v[h][1], v[h][2])) v[h][0] = v[h][1] = v[h][2] = kPadCoord();
    } else {
      v[h][0] = v[h][1] = v[h][2] = kPadCoord();
    }
  }
}

template <int QT>
__device__ __forceinline__ void stage_fill_store(PairSmem<QT>& sm, int buf, int tid, const float (&v)[2][3]) {
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

template <int QT>
__device__ __forceinline__ void pair_item(const StepParams& p, unsigned char* smem_raw, int b, int at, int ss) {
  PairSmem<QT>& sm = *reinterpret_cast<PairSmem<QT>*>(smem_raw);
  constexpr int kATile = kThreads * QT;
  constexpr int kPairs = QT / 2;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const float kInf = __int_as_float(0x7f800000);
  const int NA = p.n[p.reg], NS = p.n[p.reg ^ 1];
  const float* __restrict__ Ab = p.xyz[p.reg] + (size_t)b * NA * 3;
  const float* __restrict__ Sb = p.xyz[p.reg ^ 1] + (size_t)b * NS * 3;
  u64* PA = p.key[p.reg] + (size_t)b * NA;
  u64* PS = p.key[p.reg ^ 1] + (size_t)b * NS;
  const bool mask = p.mask_zero_sum != 0;

  // this item's streamed range, in points: [s_begin, s_end)
  const int s_begin = (int)((long long)ss * p.total_chunks / p.s_splits) * kChunk;
  const int s_end = min(NS, (int)((long long)(ss + 1) * p.total_chunks / p.s_splits) * kChunk);
  const int n_stages = (s_end - s_begin + kSStage - 1) / kSStage;

  // ---- this thread's QT consecutive A points -> QT/2 f32x2 pairs per coordinate ---------------------------------
  const int a0 = at * kATile + tid * QT;
  u64 qx[kPairs], qy[kPairs], qz[kPairs];
  {
    float f[3 * QT];
    const float* Ap = Ab + 3 * (size_t)a0;
    if (a0 + QT <= NA && (reinterpret_cast<uintptr_t>(Ap) & 15) == 0) {  // 128-bit loads: QT*12 contiguous bytes
#pragma unroll
      for (int i = 0; i < 3 * QT / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(Ap) + i);
        f[4 * i + 0] = v.x;
        f[4 * i + 1] = v.y;
        f[4 * i + 2] = v.z;
        f[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int h = 0; h < QT; ++h) {
        const bool in = a0 + h < NA;
        f[3 * h + 0] = in ? __ldg(Ap + 3 * h + 0) : -kPadCoord();
        f[3 * h + 1] = in ? __ldg(Ap + 3 * h + 1) : -kPadCoord();
        f[3 * h + 2] = in ? __ldg(Ap + 3 * h + 2) : -kPadCoord();
      }
    }
    if (mask) {
#pragma unroll
      for (int h = 0; h < QT; ++h)
        if (zero_sum(f[3 * h], f[3 * h + 1], f[3 * h + 2])) f[3 * h] = f[3 * h + 1] = f[3 * h + 2] = -kPadCoord();
    }
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      qx[i] = pack2(f[6 * i + 0], f[6 * i + 3]);
      qy[i] = pack2(f[6 * i + 1], f[6 * i + 4]);
      qz[i] = pack2(f[6 * i + 2], f[6 * i + 5]);
    }
    // Pass every pair through one packed add of +0 (x + 0 == x for the subtractions that follow; a multiply by 1 is
    // folded away by ptxas): the RESULT of an f32x2 instruction is an aligned register pair.  Without it ptxas leaves
    // the halves in the registers of the 128-bit loads and re-packs them with two moves in front of every FADD2 of the
    // hot loop (+75 % instructions, 3.09 -> 4.46 ms at config 2).
    const u64 zero2 = pack2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(qx[i]) : "l"(zero2));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(qy[i]) : "l"(zero2));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(qz[i]) : "l"(zero2));
    }
  }
  float m[QT];
#pragma unroll
  for (int i = 0; i < QT; ++i) {
    m[i] = kInf;
    sm.bestc[i][tid] = 0u;
  }
#pragma unroll
  for (int i = 0; i < QT / 4; ++i) sm.prev[i][tid] = make_float4(kInf, kInf, kInf, kInf);

  {  // prologue: first stage -> landing zone -> buffer 0
    stage_issue(Sb, s_begin, s_end, tid, sm.sraw);
    cp_async_wait_all();
    __syncthreads();
    float v[2][3];
    stage_fill_load(sm.sraw, s_begin, s_end, tid, mask, v);
    stage_fill_store<QT>(sm, 0, tid, v);
  }
  __syncthreads();

  for (int st = 0; st < n_stages; ++st) {
    const int buf = st & 1;
    const int s_base = s_begin + st * kSStage;
    const bool has_next = (st + 1 < n_stages);
    if (has_next) stage_issue(Sb, s_base + kSStage, s_end, tid, sm.sraw);  // lands while this stage is evaluated

    // number of 32-point chunks that hold at least one real point of this item
    const int n_valid = min(kSStage, s_end - s_base);
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

    cp_async_wait_all();
    __syncthreads();  // col[][] of this stage complete; the next stage's rows have landed
    if (has_next) {
      float nxt[2][3];
      stage_fill_load(sm.sraw, s_base + kSStage, s_end, tid, mask, nxt);
      stage_fill_store<QT>(sm, buf ^ 1, tid, nxt);
    }

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
          if (s < s_end) {
            const unsigned gid = (unsigned)(at * kThreads + bw * 32 + (__ffs(best.y) - 1));
            atomicMin(PS + s, ((u64)best.x << 32) | gid);
          }
        }
      }
      __syncthreads();  // col[][] may be overwritten by the next stage
    }
  }

  // ---- publish row minima: transposed through shared memory so that a warp's 32 atomics hit 32 consecutive keys -----
  {
    u64* keys = reinterpret_cast<u64*>(smem_raw);  // the stage buffers are free now: (kATile + kThreads) * 8 bytes
#pragma unroll
    for (int i = 0; i < QT; ++i) keys[tid * (QT + 1) + i] = pack_key(m[i], sm.bestc[i][tid]);
    // bestc[][] lies behind the staging area (sxy + szz + col >= 34.8 KB); all reads above precede this barrier
    __syncthreads();
#pragma unroll
    for (int r = 0; r < QT; ++r) {
      const int pt = r * kThreads + tid;
      const int a = at * kATile + pt;
      if (a < NA) atomicMin(PA + a, keys[pt + pt / QT]);
    }
  }
}

// =====================================================================================================================
// RESOLVE: exact index recovery + write-out + partial sums + fan-in count
// =====================================================================================================================
// Warp-cooperative: the lanes of a warp evaluate the candidates of one point in parallel (coalesced candidate reads), a
// ballot of the exact matches gives the first (lowest-index) one.
//   register side: candidates are the 32 streamed points of chunk `id` (one point per warp step);
//   streamed side: the `group` register-side points of group `id` (32/group points per warp step).
// The candidate distance is evaluated with the identical instruction sequence as the search, so exactly-equal compares
// are well defined.  kRU warp steps are issued together so kRU gathers are in flight per lane.
template <bool kRow>
__device__ __noinline__ void resolve_item(const StepParams& p, int b, int side, int p0, int p1, int group) {
#ifndef IAPT_RU
#define IAPT_RU 16
#endif
#ifndef IAPT_PREFETCH
#define IAPT_PREFETCH 1
#endif
  constexpr int kRU = IAPT_RU;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int other = side ^ 1;
  const int nX = p.n[side], nY = p.n[other];
  const float* __restrict__ Xb = p.xyz[side] + (size_t)b * nX * 3;
  const float* __restrict__ Yb = p.xyz[other] + (size_t)b * nY * 3;
  const u64* keys = p.key[side] + (size_t)b * nX;
  float* dist = p.dist[side] ? p.dist[side] + (size_t)b * nX : nullptr;
  int32_t* idx = p.idx[side] ? p.idx[side] + (size_t)b * nX : nullptr;
  float4* part = p.part[side] ? p.part[side] + (size_t)b * ((nX + 31) / 32) : nullptr;
  int* cnt_other = p.need_list[other] ? p.cnt[other] + (size_t)b * nY : nullptr;
  const bool mask = p.mask_zero_sum != 0;

  // this lane's point of a warp step: key + coordinates (fetched one warp step ahead)
  u64 nkey = 0ull;
  float nsx = 0.f, nsy = 0.f, nsz = 0.f;
  auto fetch = [&](int t) {
    nkey = 0ull;
    nsx = nsy = nsz = 0.f;
    if (t < p1) {
      nkey = __ldcg(keys + t);
      nsx = __ldg(Xb + 3 * (size_t)t);
      nsy = __ldg(Xb + 3 * (size_t)t + 1);
      nsz = __ldg(Xb + 3 * (size_t)t + 2);
    }
  };
#if IAPT_PREFETCH
  fetch(p0 + warp * 32 + lane);
#endif
  for (int t0 = p0 + warp * 32; t0 < p1; t0 += kWarps * 32) {
    const int t = t0 + lane;
    const bool valid = t < p1;
#if !IAPT_PREFETCH
    fetch(t);
#endif
    const u64 key = nkey;
    const float sx = nsx, sy = nsy, sz = nsz;
#if IAPT_PREFETCH
    fetch(t + kWarps * 32);
#endif
    const bool zs = zero_sum(sx, sy, sz);
    const bool masked = mask && zs;
    int myidx = 0;
    if (kRow) {
#pragma unroll 1
      for (int j0 = 0; j0 < 32; j0 += kRU) {
        if (t0 + j0 >= p1) break;  // warp-uniform
        float cx[kRU], cy[kRU], cz[kRU];
        unsigned okm = 0u;
#pragma unroll
        for (int u = 0; u < kRU; ++u) {
          const int j = j0 + u;
          const u64 kj = __shfl_sync(0xffffffffu, key, j);
          // a key that was never lowered (all-ones: every distance of this point was NaN - outside the contract)
          // must not be turned into an address
          const bool valid_key = kj != ~0ull;
          const int cand = (valid_key ? (int)(unsigned)kj * kChunk : 0) + lane;
          const bool ok = valid_key && (t0 + j < p1) && (cand < nY);
          cx[u] = cy[u] = cz[u] = 0.f;
          if (ok) {
            okm |= 1u << u;
            const float* c = Yb + (size_t)cand * 3;
            cx[u] = __ldg(c);
            cy[u] = __ldg(c + 1);
            cz[u] = __ldg(c + 2);
          }
        }
        if (mask) {  // after ALL loads of the batch were issued (a use inside the batch would serialise them)
#pragma unroll
          for (int u = 0; u < kRU; ++u)
            if (zero_sum(cx[u], cy[u], cz[u])) cx[u] = cy[u] = cz[u] = kPadCoord();
        }
#pragma unroll
        for (int u = 0; u < kRU; ++u) {
          const int j = j0 + u;
          float x = __shfl_sync(0xffffffffu, sx, j);
          float y = __shfl_sync(0xffffffffu, sy, j);
          float z = __shfl_sync(0xffffffffu, sz, j);
          if (mask && zero_sum(x, y, z)) x = y = z = -kPadCoord();
          const u64 kj = __shfl_sync(0xffffffffu, key, j);  // (again: cheaper than 2 x kRU live registers)
          const bool eq = ((okm >> u) & 1u) && (sqdist(x, y, z, cx[u], cy[u], cz[u]) == __uint_as_float((unsigned)(kj >> 32)));
          const unsigned hits = __ballot_sync(0xffffffffu, eq);
          if (lane == j) myidx = (kj != ~0ull ? (int)(unsigned)kj * kChunk : 0) + (hits ? (__ffs(hits) - 1) : 0);
        }
      }
    } else {
      const int per_step = 32 / group;  // group is 8 or 16
      const int sub = lane / group, offl = lane % group;
      const unsigned gmask = (1u << group) - 1u;
#pragma unroll 1
      for (int j0 = 0; j0 < 32; j0 += kRU * per_step) {
        if (t0 + j0 >= p1) break;  // warp-uniform
        float cx[kRU], cy[kRU], cz[kRU];
        unsigned okm = 0u;
#pragma unroll
        for (int q = 0; q < kRU; ++q) {
          const int j = j0 + q * per_step + sub;  // the point this lane helps with
          const u64 kj = __shfl_sync(0xffffffffu, key, j & 31);
          const bool valid_key = kj != ~0ull;
          const int cand = valid_key ? (int)(unsigned)kj * group + offl : 0;
          const bool ok = valid_key && (j < 32) && (t0 + j < p1) && (cand < nY);
          cx[q] = cy[q] = cz[q] = 0.f;
          if (ok) {
            okm |= 1u << q;
            const float* c = Yb + (size_t)cand * 3;
            cx[q] = __ldg(c);
            cy[q] = __ldg(c + 1);
            cz[q] = __ldg(c + 2);
          }
        }
        if (mask) {
#pragma unroll
          for (int q = 0; q < kRU; ++q)
            if (zero_sum(cx[q], cy[q], cz[q])) cx[q] = cy[q] = cz[q] = -kPadCoord();
        }
#pragma unroll
        for (int q = 0; q < kRU; ++q) {
          const int jb = j0 + q * per_step;  // first point of this step
          const int j = jb + sub;
          float x = __shfl_sync(0xffffffffu, sx, j & 31);
          float y = __shfl_sync(0xffffffffu, sy, j & 31);
          float z = __shfl_sync(0xffffffffu, sz, j & 31);
          if (mask && zero_sum(x, y, z)) x = y = z = kPadCoord();
          const unsigned kd = (unsigned)(__shfl_sync(0xffffffffu, key, j & 31) >> 32);
          const bool eq = ((okm >> q) & 1u) && (sqdist(cx[q], cy[q], cz[q], x, y, z) == __uint_as_float(kd));
          const unsigned hits = __ballot_sync(0xffffffffu, eq);
          const int owner_sub = lane - jb;  // lane jb + s owns sub-group s of this step
          if (owner_sub >= 0 && owner_sub < per_step) {
            const unsigned mine = (hits >> (owner_sub * group)) & gmask;
            myidx = (key != ~0ull) ? (int)(unsigned)key * group + (mine ? (__ffs(mine) - 1) : 0) : 0;
          }
        }
      }
    }
    const float d = __uint_as_float((unsigned)(key >> 32));
    if (valid) {
      if (dist) dist[t] = d;
      if (idx) idx[t] = myidx;
      if (cnt_other && !masked) atomicAdd(cnt_other + myidx, 1);
    }
    if (part) {  // fixed butterfly over the 32 points of this step: independent of the launch geometry
      const bool in = valid && !masked;
      const float r = sqrtf(d);  // IEEE sqrt.rn (no fast-math), same as torch.sqrt
      float s1 = in ? d : 0.f, s2 = in ? r : 0.f;
      int c1 = (in && r < p.th) ? 1 : 0, c2 = (valid && !zs) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 = __fadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, o));
        s2 = __fadd_rn(s2, __shfl_xor_sync(0xffffffffu, s2, o));
      }
      c1 = __reduce_add_sync(0xffffffffu, c1);
      c2 = __reduce_add_sync(0xffffffffu, c2);
      if (lane == 0) part[t0 >> 5] = make_float4(s1, s2, __int_as_float(c1), __int_as_float(c2));
    }
  }
}

// backward-only launches: the indices are given, RESOLVE items only count the fan-in
__device__ __noinline__ void count_item(const StepParams& p, int b, int side, int p0, int p1) {
  const int other = side ^ 1;
  if (!p.need_list[other]) return;
  const int32_t* __restrict__ idx = p.idx[side] + (size_t)b * p.n[side];
  int* cnt_other = p.cnt[other] + (size_t)b * p.n[other];
  for (int t = p0 + threadIdx.x; t < p1; t += kThreads) {
    const unsigned j = (unsigned)__ldg(idx + t);
    if (j < (unsigned)p.n[other]) atomicAdd(cnt_other + j, 1);  // out-of-range indices (caller garbage) are dropped
  }
}

// Fixed-order fold of one cloud's partials -> sums[b], counts[b], valid[b].
__device__ __noinline__ void finalize_cloud(const StepParams& p, TailSmem& ts, int b) {
  for (int side = 0; side < 2; ++side) {
    if (!p.part[side]) continue;
    const int P = (p.n[side] + 31) / 32;
    const float4* part = p.part[side] + (size_t)b * P;
    float s1 = 0.f, s2 = 0.f;
    int c1 = 0, c2 = 0;
    for (int i = threadIdx.x; i < P; i += kThreads) {
      const float4 v = __ldcg(part + i);
      s1 = __fadd_rn(s1, v.x);
      s2 = __fadd_rn(s2, v.y);
      c1 += __float_as_int(v.z);
      c2 += __float_as_int(v.w);
    }
    const float t1 = block_sum_f(s1, ts.redf);
    const float t2 = block_sum_f(s2, ts.redf);
    const int k1 = block_sum_i(c1, ts.redi);
    const int k2 = block_sum_i(c2, ts.redi);
    if (threadIdx.x == 0) {
      if (p.sums) {
        p.sums[b * 4 + side] = t1;
        p.sums[b * 4 + 2 + side] = t2;
      }
      if (p.counts) p.counts[b * 2 + side] = k1;
      if (p.valid && !p.valid_is_input) p.valid[b * 2 + side] = k2;
    }
  }
}

// Fixed-order fold of the per-cloud sums into the loss scalars of the four reference modules
// (extensions/chamfer_dist/__init__.py:43-44,61-62,82-84,101-103).  Means are over the UNMASKED points.
__device__ __noinline__ void finalize_loss(const StepParams& p, TailSmem& ts) {
  if (!p.loss || !p.sums) return;
  float t[4];
  for (int c = 0; c < 4; ++c) {
    float v = 0.f;
    for (int b = threadIdx.x; b < p.B; b += kThreads) v = __fadd_rn(v, __ldcg(p.sums + b * 4 + c));
    t[c] = block_sum_f(v, ts.redf);
  }
  int nv[2] = {0, 0};
  if (p.mask_zero_sum && p.valid) {
    for (int s = 0; s < 2; ++s) {
      int v = 0;
      for (int b = threadIdx.x; b < p.B; b += kThreads) v += __ldcg(p.valid + b * 2 + s);
      nv[s] = block_sum_i(v, ts.redi);
    }
  }
  if (threadIdx.x == 0) {
    const float n1 = p.mask_zero_sum ? (float)nv[0] : (float)p.B * (float)p.n[0];
    const float n2 = p.mask_zero_sum ? (float)nv[1] : (float)p.B * (float)p.n[1];
    const float m1 = t[0] / n1, m2 = t[1] / n2, q1 = t[2] / n1, q2 = t[3] / n2;
    float l0, l1 = 0.f;
    if (p.kind == 1) {
      l0 = (q1 + q2) / 2;
    } else if (p.kind == 2) {
      l0 = m1 + m2;
    } else if (p.kind == 3) {
      l0 = m1;
      l1 = m2;
    } else {
      l0 = q1;
    }
    p.loss[0] = l0;
    p.loss[1] = l1;
  }
}

// =====================================================================================================================
// PREFIX / FILL
// =====================================================================================================================
__device__ __noinline__ void prefix_item(const StepParams& p, TailSmem& ts, int b, int side) {
  if (!p.need_list[side]) return;
  const int n = p.n[side];
  const int* cnt = p.cnt[side] + (size_t)b * n;
  int* off = p.off[side] + (size_t)b * (n + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 16 counts per thread and iteration (4096 per CTA step): the loads of a step are independent, the steps are chained
  // through `carry`, so fewer, fatter steps shorten the serial chain (65536 bins: 64 -> 16 steps)
  constexpr int kPer = 16;
  int carry = 0;
  for (int base = 0; base < n; base += kThreads * kPer) {
    int v[kPer];
    const int i0 = base + tid * kPer;
    if (i0 + kPer <= n && ((reinterpret_cast<uintptr_t>(cnt + i0) & 15) == 0)) {
#pragma unroll
      for (int e = 0; e < kPer; e += 4) {
        const int4 q = __ldcg(reinterpret_cast<const int4*>(cnt + i0 + e));
        v[e] = q.x + 1;
        v[e + 1] = q.y + 1;
        v[e + 2] = q.z + 1;
        v[e + 3] = q.w + 1;
      }
    } else {
#pragma unroll
      for (int e = 0; e < kPer; ++e) v[e] = (i0 + e < n) ? __ldcg(cnt + i0 + e) + 1 : 0;
    }
    int mine = 0;
#pragma unroll
    for (int e = 0; e < kPer; ++e) mine += v[e];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    __syncthreads();  // wcnt free
    if (lane == 31) ts.wcnt[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int c = ts.wcnt[w];
      if (w < warp) wbase += c;
      total += c;
    }
    int run = carry + wbase + incl - mine;
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      if (i0 + e < n) off[i0 + e] = run;
      run += v[e];
    }
    carry += total;
  }
  if (tid == 0) off[n] = carry;
}

__device__ __noinline__ void fill_item(const StepParams& p, int b, int side, int p0, int p1) {
  const int other = side ^ 1;
  if (!p.need_list[other]) return;
  const int nX = p.n[side], nY = p.n[other];
  const float* __restrict__ Xb = p.xyz[side] + (size_t)b * nX * 3;
  const int32_t* idx = p.idx[side] + (size_t)b * nX;
  int* cnt = p.cnt[other] + (size_t)b * nY;
  const int* off = p.off[other] + (size_t)b * (nY + 1);
  int* csr = p.csr[other] + (size_t)b * nX;
  for (int t = p0 + threadIdx.x; t < p1; t += kThreads) {
    if (p.mask_zero_sum && zero_sum(__ldg(Xb + 3 * (size_t)t), __ldg(Xb + 3 * (size_t)t + 1), __ldg(Xb + 3 * (size_t)t + 2)))
      continue;
    const unsigned j = (unsigned)__ldcg(idx + t);
    if (j >= (unsigned)nY) continue;
    const int slot = atomicAdd(cnt + j, -1);  // remaining - 1: slots count down to 0
    csr[__ldcg(off + j) + slot] = t;
  }
}

// =====================================================================================================================
// GATHER
// =====================================================================================================================
__device__ __forceinline__ float point_weight(int mode, float scaled_upstream, const float* g, int i) {
  // returns 2 * dLoss/d dist[i]
  if (mode == 0) return __fmul_rn(__ldcg(g + i), 2.f);
  float w = scaled_upstream;
  if (mode == 1) w = __fdiv_rn(w, sqrtf(__ldcg(g + i)));  // d/dd sqrt(d) = 0.5/sqrt(d); the 0.5 is in scale
  return __fmul_rn(w, 2.f);
}

// the same weight from an already loaded source value (gw[i]; unused for mode 2)
__device__ __forceinline__ float weight_from(int mode, float scaled_upstream, float src) {
  if (mode == 0) return __fmul_rn(src, 2.f);
  float w = scaled_upstream;
  if (mode == 1) w = __fdiv_rn(w, sqrtf(src));
  return __fmul_rn(w, 2.f);
}

__device__ __forceinline__ void cswap(int& a, int& b) {
  const int lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}

__device__ __noinline__ void gather_item(const StepParams& p, TailSmem& ts, int b, int side, int p0, int p1) {
  if (!p.grad[side]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int other = side ^ 1;
  const int nX = p.n[side], nY = p.n[other];
  const float* __restrict__ Xb = p.xyz[side] + (size_t)b * nX * 3;
  const float* __restrict__ Yb = p.xyz[other] + (size_t)b * nY * 3;
  const float* gX = p.gw[side] ? p.gw[side] + (size_t)b * nX : nullptr;
  const float* gY = p.gw[other] ? p.gw[other] + (size_t)b * nY : nullptr;
  const bool fan = p.need_list[side] != 0;
  const int32_t* idxX = p.idx[side] + (size_t)b * nX;
  const int* off = fan ? p.off[side] + (size_t)b * (nX + 1) : nullptr;
  const int* csr = fan ? p.csr[side] + (size_t)b * nY : nullptr;
  float* out = p.grad[side] + (size_t)b * nX * 3;
  float up0 = 1.f, up1 = 1.f;
  if (p.mode != 0 && p.grad_out != nullptr) {
    up0 = __ldg(p.grad_out);
    up1 = __ldg(p.grad_out + p.grad_out_stride);
  }
  const int mode = p.mode;
  const float wX = __fmul_rn(side ? up1 : up0, p.scale[side]);
  const float wY = __fmul_rn(side ? up0 : up1, p.scale[other]);
  const bool mask = p.mask_zero_sum != 0;

  auto own_term = [&](int j, float x0, float x1, float x2, float& ox, float& oy, float& oz) {
    ox = oy = oz = 0.f;
    if (gX != nullptr && !(mask && zero_sum(x0, x1, x2))) {
      const unsigned j2 = (unsigned)__ldcg(idxX + j);
      if (j2 < (unsigned)nY) {
        const float g = point_weight(mode, wX, gX, j);
        ox = __fmul_rn(g, __fsub_rn(x0, __ldg(Yb + 3 * (size_t)j2 + 0)));
        oy = __fmul_rn(g, __fsub_rn(x1, __ldg(Yb + 3 * (size_t)j2 + 1)));
        oz = __fmul_rn(g, __fsub_rn(x2, __ldg(Yb + 3 * (size_t)j2 + 2)));
      }
    }
  };

  if (tid == 0) ts.n_deferred = ts.n_medium = 0;
  __syncthreads();
  // ---- pass 1: one thread per point, kG points of a thread in flight together (the chain offsets -> entries ->
  // coordinates is three dependent L2 round trips; at 16 warps per SM only instruction-level parallelism hides them).
  // Segments of <= 2 entries (92 % at a mean fan-in of 1) finish here; 3..8 go to pass 1b, longer ones to pass 2 / 3.
  constexpr int kG = 4;
  for (int jb = p0 + tid; jb < p1; jb += kG * kThreads) {
    int o0[kG], c[kG];
    unsigned nb[kG];
    float x[kG][3], wsrc[kG];
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const int j = jb + g * kThreads;
      const bool in = j < p1;
      o0[g] = (fan && in) ? __ldcg(off + j) : 0;
      c[g] = (fan && in) ? __ldcg(off + j + 1) : 0;
      nb[g] = (gX != nullptr && in) ? (unsigned)__ldcg(idxX + j) : 0xffffffffu;
      wsrc[g] = (gX != nullptr && in && mode != 2) ? __ldcg(gX + j) : 1.f;
#pragma unroll
      for (int e = 0; e < 3; ++e) x[g][e] = in ? __ldg(Xb + 3 * (size_t)j + e) : 0.f;
    }
    int e0[kG], e1[kG];
    float y[kG][3];
    bool fast[kG];
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const int j = jb + g * kThreads;
      c[g] -= o0[g];
      fast[g] = j < p1 && c[g] <= 2;
      if (j < p1 && c[g] > 2) {
        if (c[g] > 8)
          ts.deferred[atomicAdd(&ts.n_deferred, 1)] = (unsigned short)(j - p0);
        else
          ts.medium[atomicAdd(&ts.n_medium, 1)] = (unsigned short)(j - p0);
      }
      e0[g] = (fast[g] && c[g] >= 1) ? __ldcg(csr + o0[g]) : -1;
      e1[g] = (fast[g] && c[g] >= 2) ? __ldcg(csr + o0[g] + 1) : -1;
      if (nb[g] >= (unsigned)nY || (mask && zero_sum(x[g][0], x[g][1], x[g][2]))) nb[g] = 0xffffffffu;
#pragma unroll
      for (int e = 0; e < 3; ++e) y[g][e] = nb[g] != 0xffffffffu ? __ldg(Yb + 3 * (size_t)nb[g] + e) : 0.f;
    }
    float z0[kG][3], z1[kG][3], w0[kG], w1[kG];
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      if (e1[g] >= 0 && e1[g] < e0[g]) {  // ascending source index
        const int t = e0[g];
        e0[g] = e1[g];
        e1[g] = t;
      }
      w0[g] = (e0[g] >= 0 && mode != 2) ? __ldcg(gY + e0[g]) : 1.f;
      w1[g] = (e1[g] >= 0 && mode != 2) ? __ldcg(gY + e1[g]) : 1.f;
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        z0[g][e] = e0[g] >= 0 ? __ldg(Yb + 3 * (size_t)e0[g] + e) : 0.f;
        z1[g][e] = e1[g] >= 0 ? __ldg(Yb + 3 * (size_t)e1[g] + e) : 0.f;
      }
    }
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      if (!fast[g]) continue;
      const int j = jb + g * kThreads;
      float o[3] = {0.f, 0.f, 0.f}, a[3] = {0.f, 0.f, 0.f};
      if (nb[g] != 0xffffffffu) {
        const float gw = weight_from(mode, wX, wsrc[g]);
#pragma unroll
        for (int e = 0; e < 3; ++e) o[e] = __fmul_rn(gw, __fsub_rn(x[g][e], y[g][e]));
      }
      if (e0[g] >= 0) {
        const float gw = weight_from(mode, wY, w0[g]);
#pragma unroll
        for (int e = 0; e < 3; ++e) a[e] = __fadd_rn(a[e], __fmul_rn(gw, __fsub_rn(z0[g][e], x[g][e])));
      }
      if (e1[g] >= 0) {
        const float gw = weight_from(mode, wY, w1[g]);
#pragma unroll
        for (int e = 0; e < 3; ++e) a[e] = __fadd_rn(a[e], __fmul_rn(gw, __fsub_rn(z1[g][e], x[g][e])));
      }
#pragma unroll
      for (int e = 0; e < 3; ++e) out[3 * (size_t)j + e] = __fsub_rn(o[e], a[e]);
    }
  }
  __syncthreads();
  // ---- pass 1b: segments of 3..8 entries, one thread each: ordered by a sorting network in registers ------------------
  const int n_med = ts.n_medium;
  for (int d = tid; d < n_med; d += kThreads) {
    const int j = p0 + ts.medium[d];
    const int o0 = __ldcg(off + j);
    const int c = __ldcg(off + j + 1) - o0;
    int e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = (i < c) ? __ldcg(csr + o0 + i) : 0x7fffffff;
    const float x0 = __ldg(Xb + 3 * (size_t)j), x1 = __ldg(Xb + 3 * (size_t)j + 1), x2 = __ldg(Xb + 3 * (size_t)j + 2);
    float ox, oy, oz;
    own_term(j, x0, x1, x2, ox, oy, oz);
    // 8-input sorting network (19 compare-exchanges)
    cswap(e[0], e[1]); cswap(e[2], e[3]); cswap(e[4], e[5]); cswap(e[6], e[7]);
    cswap(e[0], e[2]); cswap(e[1], e[3]); cswap(e[4], e[6]); cswap(e[5], e[7]);
    cswap(e[1], e[2]); cswap(e[5], e[6]); cswap(e[0], e[4]); cswap(e[3], e[7]);
    cswap(e[1], e[5]); cswap(e[2], e[6]);
    cswap(e[1], e[4]); cswap(e[3], e[6]);
    cswap(e[2], e[4]); cswap(e[3], e[5]);
    cswap(e[3], e[4]);
    // all gathers first (independent), then the ordered sum
    float gw[8], yx[8], yy[8], yz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < c) {
        const int k = e[i];
        gw[i] = point_weight(mode, wY, gY, k);
        yx[i] = __ldg(Yb + 3 * (size_t)k + 0);
        yy[i] = __ldg(Yb + 3 * (size_t)k + 1);
        yz[i] = __ldg(Yb + 3 * (size_t)k + 2);
      }
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < c) {
        ax = __fadd_rn(ax, __fmul_rn(gw[i], __fsub_rn(yx[i], x0)));
        ay = __fadd_rn(ay, __fmul_rn(gw[i], __fsub_rn(yy[i], x1)));
        az = __fadd_rn(az, __fmul_rn(gw[i], __fsub_rn(yz[i], x2)));
      }
    }
    out[3 * (size_t)j + 0] = __fsub_rn(ox, ax);
    out[3 * (size_t)j + 1] = __fsub_rn(oy, ay);
    out[3 * (size_t)j + 2] = __fsub_rn(oz, az);
  }
  const int n_def = ts.n_deferred;
  if (n_def == 0) return;  // CTA-uniform

  // ---- pass 2 / 3: segments of more than 8 entries are summed EXACTLY, so no order has to be established: every term is
  // scaled by a power of two S chosen from the segment's largest |term| (|term * S| < 2^38), rounded once to a 64-bit
  // integer (absolute error <= 2^-39 of the largest term - far below the n * 2^-24 of any float32 summation order) and
  // the integers are added - associative, hence independent of the order of the CSR segment (which the FILL atomics
  // leave to scheduling), of the warp / CTA shape and of the launch geometry.  Non-finite terms (0 distance under L1:
  // inf * 0, like the reference's sqrt backward) bypass the integers: sums over {0, +-Inf, NaN} are order-free too.
  // A warp takes a segment of <= kWarpSeg entries, the whole CTA a longer one (a collapsed prediction: thousands of
  // points sharing one neighbour) - a parallel reduction instead of the reference's serialised atomicAdds
  // (chamfer.cu:193-198).
  // terms of the entries seg[i0 + u * stride], u < kSU (zero beyond c): all index loads, then all gathers, then the math -
  // two dependent L2 round trips per kSU entries instead of per entry
  constexpr int kSU = 4;
  auto terms = [&](const int* seg, int i0, int stride, int c, float x0, float x1, float x2, float (&tx)[kSU],
                   float (&ty)[kSU], float (&tz)[kSU]) {
    int k[kSU];
#pragma unroll
    for (int u = 0; u < kSU; ++u) k[u] = (i0 + u * stride < c) ? __ldcg(seg + i0 + u * stride) : -1;
    float w[kSU], yx[kSU], yy[kSU], yz[kSU];
#pragma unroll
    for (int u = 0; u < kSU; ++u) {
      const bool in = k[u] >= 0;
      w[u] = (in && mode != 2) ? __ldcg(gY + k[u]) : 1.f;
      yx[u] = in ? __ldg(Yb + 3 * (size_t)k[u] + 0) : 0.f;
      yy[u] = in ? __ldg(Yb + 3 * (size_t)k[u] + 1) : 0.f;
      yz[u] = in ? __ldg(Yb + 3 * (size_t)k[u] + 2) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kSU; ++u) {
      const float g = k[u] >= 0 ? weight_from(mode, wY, w[u]) : 0.f;
      tx[u] = k[u] >= 0 ? __fmul_rn(g, __fsub_rn(yx[u], x0)) : 0.f;
      ty[u] = k[u] >= 0 ? __fmul_rn(g, __fsub_rn(yy[u], x1)) : 0.f;
      tz[u] = k[u] >= 0 ? __fmul_rn(g, __fsub_rn(yz[u], x2)) : 0.f;
    }
  };
  auto finite_abs = [](float t) { return (__float_as_uint(t) & 0x7f800000u) != 0x7f800000u ? fabsf(t) : 0.f; };
  auto nonfinite = [](float t) { return (__float_as_uint(t) & 0x7f800000u) != 0x7f800000u ? 0.f : t; };
  // S = 2^(37 - e) for the largest finite |term| = 1.m * 2^e (so |term * S| < 2^38), clamped to float range
  auto scale_for = [](float mx, float& S, float& invS) {
    int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;  // denormal maximum: e = -127 (terms * S stay tiny)
    int se = 37 - e;
    se = max(-100, min(100, se));
    S = __uint_as_float((unsigned)(se + 127) << 23);
    invS = __uint_as_float((unsigned)(127 - se) << 23);
  };
  auto to_fixed = [](float t, float S) -> long long {
    return (__float_as_uint(t) & 0x7f800000u) != 0x7f800000u ? __float2ll_rn(__fmul_rn(t, S)) : 0ll;
  };

  bool any_huge = false;
  for (int d = warp; d < n_def; d += kWarps) {
    const int j = p0 + ts.deferred[d];
    const int o0 = __ldcg(off + j);
    const int c = __ldcg(off + j + 1) - o0;
    if (c > kWarpSeg) {
      any_huge = true;
      continue;
    }
    const float x0 = __ldg(Xb + 3 * (size_t)j), x1 = __ldg(Xb + 3 * (size_t)j + 1), x2 = __ldg(Xb + 3 * (size_t)j + 2);
    float mx = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    for (int i0 = lane; i0 < c; i0 += 32 * kSU) {
      float tx[kSU], ty[kSU], tz[kSU];
      terms(csr + o0, i0, 32, c, x0, x1, x2, tx, ty, tz);
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        mx = fmaxf(mx, fmaxf(finite_abs(tx[u]), fmaxf(finite_abs(ty[u]), finite_abs(tz[u]))));
        nx = __fadd_rn(nx, nonfinite(tx[u]));
        ny = __fadd_rn(ny, nonfinite(ty[u]));
        nz = __fadd_rn(nz, nonfinite(tz[u]));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      nx = __fadd_rn(nx, __shfl_xor_sync(0xffffffffu, nx, o));
      ny = __fadd_rn(ny, __shfl_xor_sync(0xffffffffu, ny, o));
      nz = __fadd_rn(nz, __shfl_xor_sync(0xffffffffu, nz, o));
    }
    float S, invS;
    scale_for(mx, S, invS);
    long long sx = 0, sy = 0, sz = 0;
    for (int i0 = lane; i0 < c; i0 += 32 * kSU) {
      float tx[kSU], ty[kSU], tz[kSU];
      terms(csr + o0, i0, 32, c, x0, x1, x2, tx, ty, tz);
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        sx += to_fixed(tx[u], S);
        sy += to_fixed(ty[u], S);
        sz += to_fixed(tz[u], S);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) {
      float ox, oy, oz;
      own_term(j, x0, x1, x2, ox, oy, oz);
      out[3 * (size_t)j + 0] = __fsub_rn(ox, __fadd_rn(__fmul_rn(__ll2float_rn(sx), invS), nx));
      out[3 * (size_t)j + 1] = __fsub_rn(oy, __fadd_rn(__fmul_rn(__ll2float_rn(sy), invS), ny));
      out[3 * (size_t)j + 2] = __fsub_rn(oz, __fadd_rn(__fmul_rn(__ll2float_rn(sz), invS), nz));
    }
  }
  if (!__syncthreads_or(any_huge ? 1 : 0)) return;

  for (int d = 0; d < n_def; ++d) {
    const int j = p0 + ts.deferred[d];
    const int o0 = __ldcg(off + j);
    const int c = __ldcg(off + j + 1) - o0;
    if (c <= kWarpSeg) continue;  // CTA-uniform
    const float x0 = __ldg(Xb + 3 * (size_t)j), x1 = __ldg(Xb + 3 * (size_t)j + 1), x2 = __ldg(Xb + 3 * (size_t)j + 2);
    float mx = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    for (int i0 = tid; i0 < c; i0 += kThreads * kSU) {
      float tx[kSU], ty[kSU], tz[kSU];
      terms(csr + o0, i0, kThreads, c, x0, x1, x2, tx, ty, tz);
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        mx = fmaxf(mx, fmaxf(finite_abs(tx[u]), fmaxf(finite_abs(ty[u]), finite_abs(tz[u]))));
        nx = __fadd_rn(nx, nonfinite(tx[u]));
        ny = __fadd_rn(ny, nonfinite(ty[u]));
        nz = __fadd_rn(nz, nonfinite(tz[u]));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      nx = __fadd_rn(nx, __shfl_xor_sync(0xffffffffu, nx, o));
      ny = __fadd_rn(ny, __shfl_xor_sync(0xffffffffu, ny, o));
      nz = __fadd_rn(nz, __shfl_xor_sync(0xffffffffu, nz, o));
    }
    __syncthreads();  // red arrays free
    if (lane == 0) {
      ts.redf[warp] = mx;
      ts.redf[8 + warp] = nx;
      ts.redf[16 + warp] = ny;
      ts.redf[24 + warp] = nz;
    }
    __syncthreads();
    mx = nx = ny = nz = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      mx = fmaxf(mx, ts.redf[w]);
      nx = __fadd_rn(nx, ts.redf[8 + w]);
      ny = __fadd_rn(ny, ts.redf[16 + w]);
      nz = __fadd_rn(nz, ts.redf[24 + w]);
    }
    float S, invS;
    scale_for(mx, S, invS);
    long long sx = 0, sy = 0, sz = 0;
    for (int i0 = tid; i0 < c; i0 += kThreads * kSU) {
      float tx[kSU], ty[kSU], tz[kSU];
      terms(csr + o0, i0, kThreads, c, x0, x1, x2, tx, ty, tz);
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        sx += to_fixed(tx[u], S);
        sy += to_fixed(ty[u], S);
        sz += to_fixed(tz[u], S);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) {
      ts.redl[warp] = sx;
      ts.redl[8 + warp] = sy;
      ts.redl[16 + warp] = sz;
    }
    __syncthreads();
    if (tid < 3) {
      long long t = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += ts.redl[8 * tid + w];
      float ox, oy, oz;
      own_term(j, x0, x1, x2, ox, oy, oz);
      const float own = tid == 0 ? ox : (tid == 1 ? oy : oz);
      const float nf = tid == 0 ? nx : (tid == 1 ? ny : nz);
      out[3 * (size_t)j + tid] = __fsub_rn(own, __fadd_rn(__fmul_rn(__ll2float_rn(t), invS), nf));
    }
  }
}

// =====================================================================================================================
// the persistent kernel
// =====================================================================================================================
// QT: A-side points per thread of the PAIR phase (8 or 16); QT = 0: no PAIR / RESOLVE code (backward-only launches).
template <int QT>
__global__ void __launch_bounds__(kThreads, 2) chamfer_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtlSmem& ctl = *reinterpret_cast<CtlSmem*>(smem_raw + StepSmem<QT>::kCtlOff);
  TailSmem& ts = *reinterpret_cast<TailSmem*>(smem_raw);
  const int tid = threadIdx.x;
  if (p.valid_is_input) {  // gated launch (ignore_zeros second pass): nothing to do when no cloud holds a masked point
    int differs = 0;
    for (int i = tid; i < 2 * p.B; i += kThreads) differs |= (__ldcg(p.valid + i) != p.n[i & 1]) ? 1 : 0;
    if (!__syncthreads_or(differs)) return;
  }
  if (p.stats && tid == 0) atomicMin(p.stats + 1, globaltimer_ns());

  for (;;) {
    __syncthreads();  // everyone is done with smem (incl. ctl) of the previous item
    if (tid == 0) {
      const long long item = (long long)(unsigned)(atomicAdd(p.ticket, 1u) + 1u);  // counters start at 0xFFFFFFFF
      int ph = -1, b = 0, q = 0;
      if (item < (long long)p.total_items) {
        // items before step t: C(t) = sum_ph per_cloud[ph] * clamp(t - lag[ph], 0, B); find the step that holds `item`
        auto before = [&](int t) {
          long long c = 0;
#pragma unroll
          for (int h = 0; h < 5; ++h) c += (long long)p.per_cloud[h] * min(max(t - p.lag[h], 0), p.B);
          return c;
        };
        int lo = 0, hi = p.n_steps;  // C(lo) <= item < C(hi)
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (before(mid) <= item) lo = mid; else hi = mid;
        }
        int rem = (int)(item - before(lo));
#pragma unroll
        for (int h = 0; h < 5; ++h) {
          const int c = lo - p.lag[h];
          if (ph < 0 && c >= 0 && c < p.B) {
            if (rem < p.per_cloud[h]) {
              ph = h;
              b = c;
              q = rem;
            } else {
              rem -= p.per_cloud[h];
            }
          }
        }
      }
      ctl.ph = ph;
      ctl.b = b;
      ctl.q = q;
    }
    __syncthreads();
    const int ph = ctl.ph, b = ctl.b, q = ctl.q;
    if (ph < 0) break;

    if (ph == 0) {
      if constexpr (QT != 0) {
        const int at = q / p.s_splits;
        const int ss = q - at * p.s_splits;
        if (p.stats && tid == 0) atomicMin(p.stats + 0, globaltimer_ns());
        pair_item<QT>(p, smem_raw, b, at, ss);
        signal_done(p.done_pair + b, p.pair_per_cloud, ctl);
        if (p.stats && tid == 0) {
          atomicMax(p.stats + 2, globaltimer_ns());
          atomicMax(p.stats + 4, globaltimer_ns());
        }
      }
      continue;
    }
    if (ph == 2) {
      wait_done(p.done_res + b, p.rc_per_cloud);
      prefix_item(p, ts, b, q);
      signal_done(p.done_pre + b, 2, ctl);
      if (p.stats && tid == 0) atomicMax(p.stats + 6, globaltimer_ns());
      continue;
    }
    // RESOLVE / FILL / GATHER items address `rc` points of one side
    const int side = q < p.rc_items[0] ? 0 : 1;
    const int p0 = (side ? q - p.rc_items[0] : q) * p.rc[side];
    const int p1 = min(p.n[side], p0 + p.rc[side]);
    if (ph == 1) {
      if constexpr (QT != 0) {
        wait_done(p.done_pair + b, p.pair_per_cloud);
        if (side == p.reg)
          resolve_item<true>(p, b, side, p0, p1, QT);
        else
          resolve_item<false>(p, b, side, p0, p1, QT);
      } else {
        count_item(p, b, side, p0, p1);
      }
      const bool last_of_cloud = signal_done(p.done_res + b, p.rc_per_cloud, ctl);
      if (last_of_cloud && (p.part[0] || p.part[1])) {
        finalize_cloud(p, ts, b);
        if (signal_done(p.done_final, p.B, ctl)) finalize_loss(p, ts);
      }
    } else if (ph == 3) {
      wait_done(p.done_pre + b, 2);
      fill_item(p, b, side, p0, p1);
      signal_done(p.done_fill + b, p.rc_per_cloud, ctl);
    } else {
      // without fan-in lists (e.g. L1_PM with only xyz1 wanted) there are no PREFIX / FILL items: dist / idx suffice
      wait_done(p.per_cloud[3] ? p.done_fill + b : p.done_res + b, p.rc_per_cloud);
      gather_item(p, ts, b, side, p0, p1);
    }
    if (p.stats) {
      __syncthreads();
      if (tid == 0) atomicMax(p.stats + 4 + ph, globaltimer_ns());
    }
  }
  if (p.stats && tid == 0) atomicMax(p.stats + 3, globaltimer_ns());
}

// =====================================================================================================================
// host side
// =====================================================================================================================
// ---- optional per-launch timing of the step kernel (bench.py's roofline.achieved) --------------------------------
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

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct StepLayout {
  size_t stats, memset_begin, key[2], cnt[2], counters, memset_end, off[2], csr[2], part[2], sums, valid, total;
};

// Workspace layout.  `search`: keys present; `lists`: cnt/off/csr present; `parts`: partial sums present.
static StepLayout step_layout(int B, int N, int M, bool search, bool lists, bool parts) {
  StepLayout L;
  const size_t n[2] = {(size_t)N, (size_t)M};
  L.stats = 0;  // fixed place: iapt_chamfer_step_stats reads it without knowing the launch's flags
  size_t o = 256;
  L.memset_begin = o;
  for (int s = 0; s < 2; ++s) {
    L.key[s] = o;
    if (search) o += align256(sizeof(u64) * B * n[s]);
  }
  for (int s = 0; s < 2; ++s) {
    L.cnt[s] = o;
    if (lists) o += align256(sizeof(int) * B * n[s]);
  }
  L.counters = o;
  o += align256(64 + sizeof(unsigned) * 4 * (size_t)B);
  L.memset_end = o;
  for (int s = 0; s < 2; ++s) {
    L.off[s] = o;
    if (lists) o += align256(sizeof(int) * B * (n[s] + 1));
  }
  for (int s = 0; s < 2; ++s) {
    L.csr[s] = o;
    if (lists) o += align256(sizeof(int) * B * n[1 - s]);
  }
  for (int s = 0; s < 2; ++s) {
    L.part[s] = o;
    if (parts) o += align256(sizeof(float4) * B * ((n[s] + 31) / 32));
  }
  L.sums = o;
  if (parts) o += align256(sizeof(float) * 4 * (size_t)B);
  L.valid = o;
  if (parts) o += align256(sizeof(int32_t) * 2 * (size_t)B);
  L.total = o;
  return L;
}

size_t chamfer_step_workspace_bytes(int B, int N, int M) { return step_layout(B, N, M, true, true, true).total; }
size_t chamfer_forward_workspace_bytes(int B, int N, int M) { return step_layout(B, N, M, true, false, false).total; }
size_t chamfer_backward_workspace_bytes(int B, int N, int M) { return step_layout(B, N, M, false, true, false).total; }

static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e && *e ? std::atoi(e) : dflt;
}

static int device_sms() {
  static int sms_of_device[256] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int num_sms = sms_of_device[dev & 255];
  if (num_sms == 0) {
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
    sms_of_device[dev & 255] = num_sms;
  }
  return num_sms;
}

// PAIR item shape: A-side points per thread (QT = 16 or 8: tiles of 4096 / 2048 points) and the number of ranges the
// streamed side is split into.  Measured (profiles/r02g_split_sweep.txt, B = 6..48 x 16384^2): the search time follows
//     ceil(items / SMs) / (items / SMs)  x  (1 + o * s_splits / chunks),     o ~ 0.35 chunks of per-item cost
// i.e. the SM (its two resident CTAs share the work fluidly) is the unit that quantises, and an item costs about a third
// of a 32-point chunk in tile loads and key publishing; QT = 16 needs ~4 % fewer issue slots per pair than QT = 8.
// The model is evaluated for every split count (a few hundred integer operations; the choice is cached per shape).
// IAPT_FWD_VARIANT = 8|16 and IAPT_S_SPLITS force either (read per call: the parity tests run both over every suite).
struct PairShape {
  int qt, s_splits;
};
static PairShape pick_pair_shape(int B, int NA, int NS, int num_sms) {
  struct Key {
    int B, NA, NS, sms, fq, fs;
    bool operator==(const Key& o) const { return B == o.B && NA == o.NA && NS == o.NS && sms == o.sms && fq == o.fq && fs == o.fs; }
  };
  thread_local Key last_key{-1, 0, 0, 0, 0, 0};
  thread_local PairShape last{8, 1};
  const int forced_qt = env_int("IAPT_FWD_VARIANT", 0), forced_s = env_int("IAPT_S_SPLITS", 0);
  const Key key{B, NA, NS, num_sms, forced_qt, forced_s};
  if (key == last_key) return last;
  const int chunks = (NS + kChunk - 1) / kChunk;
  const int min_chunks = 4;
  PairShape best{8, 1};
  double best_cost = 1e300;
  for (int qt = 8; qt <= 16; qt += 8) {
    if (forced_qt == 8 || forced_qt == 16) {
      if (qt != forced_qt) continue;
    } else if (qt == 16 && NA < 4096) {
      continue;
    }
    const int tile = kThreads * qt;
    const long long base = (long long)B * ((NA + tile - 1) / tile);
    const double per_chunk = (double)tile * (qt == 16 ? 1.0 : 1.04);  // issue time of one chunk against one tile
    const double overhead = qt == 16 ? 0.35 : 0.5;                    // per item, in chunks
    const int s_max = chunks / min_chunks > 1 ? chunks / min_chunks : 1;
    for (int s = 1; s <= s_max; ++s) {
      if (forced_s && s != (forced_s < s_max ? forced_s : s_max)) continue;
      const long long items = base * s;
      // an exact multiple of the SM count behaves like one item too many (any jitter makes some CTA start an extra
      // item alone): at least 3 % of an item of slack is required
      const long long rounds = (long long)std::ceil((double)items / num_sms + 0.03);
      const double csize = (double)chunks / s;
      // every SM runs `rounds` items of chunks / s chunks (+ overhead); fewer than 4 rounds leave nothing for the
      // dynamic ticket to even out, which is charged as half an item
      double cost = (double)rounds * (csize + overhead) * per_chunk;
      if (rounds < 4) cost += 0.5 * (csize + overhead) * per_chunk;
      if (cost < best_cost) {
        best_cost = cost;
        best = PairShape{qt, s};
      }
    }
  }
  last_key = key;
  last = best;
  return best;
}

template <int QT>
static cudaError_t launch_step(const StepParams& p, int grid, cudaStream_t stream) {
  static PerDeviceOnce once;
  if (once.pending()) {
    cudaFuncSetAttribute(chamfer_step_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)StepSmem<QT>::kBytes);
    once.done();
  }
  chamfer_step_kernel<QT><<<grid, kThreads, StepSmem<QT>::kBytes, stream>>>(p);
  return cudaGetLastError();
}

int chamfer_step_launch(const StepArgs& a, cudaStream_t stream) {
  const int B = a.B, N = a.N, M = a.M;
  const bool search = a.search;
  const bool want_grad[2] = {a.grad_xyz1 != nullptr, a.grad_xyz2 != nullptr};
  const bool parts = search && (a.sums || a.counts || a.valid || a.loss);

  StepParams p{};
  p.xyz[0] = a.xyz1;
  p.xyz[1] = a.xyz2;
  p.n[0] = N;
  p.n[1] = M;
  p.B = B;
  p.reg = N < M ? 1 : 0;  // register side = the larger cloud (its tile is 256*QT points; the streamed side pads to 32)
  p.mask_zero_sum = a.mask_zero_sum ? 1 : 0;
  p.dist[0] = a.dist1;
  p.dist[1] = a.dist2;
  p.idx[0] = a.idx1;
  p.idx[1] = a.idx2;
  p.th = a.th;
  p.kind = a.kind;
  p.mode = a.mode;
  p.scale[0] = a.scale1;
  p.scale[1] = a.scale2;
  p.grad_out = a.grad_out;
  p.grad_out_stride = a.grad_out_stride;
  p.gw[0] = a.gw1;
  p.gw[1] = a.gw2;
  p.grad[0] = a.grad_xyz1;
  p.grad[1] = a.grad_xyz2;
  bool lists = false;
  for (int s = 0; s < 2; ++s) {
    p.need_list[s] = (want_grad[s] && p.gw[1 - s] != nullptr) ? 1 : 0;
    lists = lists || p.need_list[s];
  }
  const bool grads = want_grad[0] || want_grad[1];

  const StepLayout L = step_layout(B, N, M, search, lists, parts);
  unsigned char* ws = static_cast<unsigned char*>(a.workspace);
  for (int s = 0; s < 2; ++s) {
    p.key[s] = search ? reinterpret_cast<u64*>(ws + L.key[s]) : nullptr;
    p.cnt[s] = lists ? reinterpret_cast<int*>(ws + L.cnt[s]) : nullptr;
    p.off[s] = lists ? reinterpret_cast<int*>(ws + L.off[s]) : nullptr;
    p.csr[s] = lists ? reinterpret_cast<int*>(ws + L.csr[s]) : nullptr;
    p.part[s] = parts ? reinterpret_cast<float4*>(ws + L.part[s]) : nullptr;
  }
  unsigned* ctr = reinterpret_cast<unsigned*>(ws + L.counters);
  p.ticket = ctr;
  p.done_final = ctr + 1;
  p.done_pair = ctr + 16;
  p.done_res = p.done_pair + B;
  p.done_pre = p.done_res + B;
  p.done_fill = p.done_pre + B;
  p.sums = parts ? (a.sums ? a.sums : reinterpret_cast<float*>(ws + L.sums)) : nullptr;
  p.counts = parts ? a.counts : nullptr;
  p.valid = parts ? (a.valid ? a.valid : reinterpret_cast<int32_t*>(ws + L.valid)) : nullptr;
  p.valid_is_input = (parts && a.valid && a.valid_is_input) ? 1 : 0;
  p.loss = parts ? a.loss : nullptr;
  p.stats = a.stats ? reinterpret_cast<u64*>(ws + L.stats) : nullptr;

  cudaError_t err;
  // one memset: keys = all-ones, every counter and fan-in count = 0xFFFFFFFF (read as value + 1)
  if ((err = cudaMemsetAsync(ws + L.memset_begin, 0xFF, L.memset_end - L.memset_begin, stream)) != cudaSuccess)
    return (int)err;
  if (p.stats) {  // measurement only: two more memsets (min slots all-ones, max slots zero)
    if ((err = cudaMemsetAsync(ws + L.stats, 0xFF, 16, stream)) != cudaSuccess) return (int)err;
    if ((err = cudaMemsetAsync(ws + L.stats + 16, 0, 56, stream)) != cudaSuccess) return (int)err;
  }

  const int num_sms = device_sms();
  int qt = 0;
  const int slots = num_sms * 2;
  if (search) {
    const int NA = p.n[p.reg], NS = p.n[p.reg ^ 1];
    const PairShape shape = pick_pair_shape(B, NA, NS, num_sms);
    qt = shape.qt;
    const int a_tile = kThreads * qt;
    p.a_tiles = (NA + a_tile - 1) / a_tile;
    const int total_chunks = (NS + kChunk - 1) / kChunk;
    p.s_splits = shape.s_splits;
    p.total_chunks = total_chunks;
    p.pair_per_cloud = p.a_tiles * p.s_splits;
  } else {
    p.a_tiles = p.s_splits = p.total_chunks = 0;
    p.pair_per_cloud = 0;
  }
  // RESOLVE / FILL / GATHER items: rc[s] points of side s (a power of two), small enough that the side alone gives
  // every resident slot 2 items.  A side whose points collect many sources each (n[s] << n[1-s]: the coarse 512-point
  // prediction against 16384 ground-truth points) gets items down to 32 points - its GATHER work per point is a whole
  // warp's; otherwise an item is at least 256 points (one per thread).
  for (int sd = 0; sd < 2; ++sd) {
    const long long fan = p.n[sd] > 0 ? (long long)p.n[1 - sd] / p.n[sd] : 1;
    int lo = 256;
    while (fan > 2 && lo > 32 && (long long)(256 / lo) < fan) lo >>= 1;
    int rc = kMaxRc / 2;
    while (rc > lo && (long long)B * ((p.n[sd] + rc - 1) / rc) < 2LL * slots) rc >>= 1;
    if (const int forced_rc = env_int("IAPT_RC", 0)) rc = forced_rc;  // development knob: 32 .. 2048, power of two
    p.rc[sd] = rc;
    p.rc_items[sd] = (p.n[sd] + rc - 1) / rc;
  }
  p.rc_per_cloud = p.rc_items[0] + p.rc_items[1];
  p.per_cloud[0] = p.pair_per_cloud;
  p.per_cloud[1] = p.rc_per_cloud;
  p.per_cloud[2] = lists ? 2 : 0;
  p.per_cloud[3] = lists ? p.rc_per_cloud : 0;
  p.per_cloud[4] = grads ? p.rc_per_cloud : 0;
  long long total = 0;
  for (int h = 0; h < 5; ++h) total += (long long)B * p.per_cloud[h];
  if (total >= (1LL << 31) - 2) return -3;  // IAPT_EUNSUPPORTED
  p.total_items = (int)total;
  // Ticket order: phase-major (all PAIR items, then all RESOLVE items, ...), clouds in order inside a phase.  The
  // general form - phase ph of cloud b at step b + lag[ph] - exists to MEASURE the alternative of running the tail
  // phases of finished clouds beside the search of later ones (IAPT_LAG = L: RESOLVE lags the search by L clouds,
  // PREFIX / FILL / GATHER one more step each).  It loses at every lag (config 2: L = 2 / 4 / 8 -> 4.4 / 3.85 / 3.36 ms
  // against 3.26 phase-major): the search needs both resident CTAs of an SM to fill the FP32 issue port, so a CTA slot
  // spent on a latency-bound tail item costs more search time than the tail it hides.
  {
    int L = B;
    if (const int forced = env_int("IAPT_LAG", -1); forced >= 0) L = forced;
    const bool phase_major = L >= B;
    for (int h = 0; h < 5; ++h) p.lag[h] = phase_major ? h * B : (h == 0 ? 0 : L + h - 1);
    p.n_steps = B + p.lag[4];
  }

  const int grid = (int)(total < (long long)slots ? total : (long long)slots);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_profile && search) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  err = qt == 16 ? launch_step<16>(p, grid, stream) : (qt == 8 ? launch_step<8>(p, grid, stream) : launch_step<0>(p, grid, stream));
  count_launch();
  if (g_profile && search) {
    cudaEventRecord(ev1, stream);
    g_events.emplace_back(ev0, ev1);
  }
  return (int)err;
}

// dst[i] = src[i] * factor[0]  (the upstream gradient of a loss whose unit gradient was formed in the forward launch)
__global__ void __launch_bounds__(256) scale_kernel(const float* __restrict__ src, const float* __restrict__ factor,
                                                    long long n, float* __restrict__ dst) {
  const float f = __ldg(factor);
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (; i < n4; i += stride) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      v.x = __fmul_rn(v.x, f);
      v.y = __fmul_rn(v.y, f);
      v.z = __fmul_rn(v.z, f);
      v.w = __fmul_rn(v.w, f);
      reinterpret_cast<float4*>(dst)[i] = v;
    }
    for (long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
      dst[t] = __fmul_rn(__ldg(src + t), f);
  } else {
    for (; i < n; i += stride) dst[i] = __fmul_rn(__ldg(src + i), f);
  }
}

int scale_launch(const float* src, const float* factor, long long n, float* dst, cudaStream_t stream) {
  long long blocks = ((n >> 2) + 255) / 256;
  const long long cap = (long long)device_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  scale_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, factor, n, dst);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace iapt
