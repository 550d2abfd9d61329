// pipe_mix.cu - sm_100a issue/pipe microbenchmark behind DESIGN.md section 3.1: how many FMA-pipe cycles does a
// packed f32x2 op hold, and how much of the pipe survives when FMNMX3 / REDUX / LDS share the issue port?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu ; run: ./pipe_mix
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned redux(unsigned v) { unsigned r; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }

constexpr int ITER = 4096;
// MODE 0: 8 independent FFMA2 chains            (packed only)
// MODE 1: 16 independent scalar FFMA chains     (scalar only, same FLOPs as MODE 0)
// MODE 2: 8 FFMA2 (x,x,acc form) + 8 sub2       (our FP mix: 2 distinct 64-bit sources)
// MODE 3: MODE 2 + 4 FMNMX3 per 16 packed       (ratio of the pair kernel: 1 ALU per 3 packed ... here 1:4)
// MODE 4: MODE 2 + 6 FMNMX3 per 16 packed       (1 : 2.67)
// MODE 5: MODE 4 + 1 REDUX + 1 LDS.128 per 16 packed
// MODE 6: MODE 2 + 8 FMNMX (2-input)      MODE 7: MODE 2 + 8 IMNMX.U32 (min.u32)   MODE 8: MODE 2 + 4 VIMNMX3.U32
// MODE 9: MODE 2 + 8 IADD3                MODE 10: 16 FMNMX3 only                  MODE 11: 16 VIMNMX3.U32 only
// MODE 12: 16 FMNMX (2-input) only
template <int MODE>
__global__ void k(float* out, const float4* smem_src, int iters) {
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = smem_src[threadIdx.x];
  __syncthreads();
  u64 a[8], b[8];
  float s[16], m[8];
  for (int i = 0; i < 8; ++i) { a[i] = (u64)(threadIdx.x + i) * 0x3f8000013f800001ull; b[i] = a[i] ^ 0x0000100000001000ull; m[i] = 1e30f + i; }
  for (int i = 0; i < 16; ++i) s[i] = 1.0f + threadIdx.x * 1e-6f + i;
  unsigned rx = threadIdx.x;
  float4 ld = sm[threadIdx.x & 63];
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], b[i], a[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], b[i], a[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] = fma1(s[i], s[(i + 1) & 15], s[i]);
    } else if (MODE == 13 || MODE == 14) {
      // packed ops + independent SCALAR FMAs: do the scalar ones ride on the other FMA pipe for free?
#pragma unroll
      for (int i = 0; i < 8; ++i) { u64 d = sub2(b[i], a[(i + 1) & 7]); a[i] = fma2(d, d, a[i]); }
#pragma unroll
      for (int i = 0; i < (MODE == 13 ? 8 : 16); ++i) s[i] = fma1(s[i], s[(i + 1) & 15], s[i]);
    } else if (MODE == 10) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = min3(s[i], s[(i + 1) & 15], s[(i + 5) & 15]);
    } else if (MODE == 11) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        s[i] = __uint_as_float(__vimin3_u32(__float_as_uint(s[i]), __float_as_uint(s[(i + 1) & 15]), __float_as_uint(s[(i + 5) & 15])));
    } else if (MODE == 12) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = fminf(s[i], s[(i + 1) & 15]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { u64 d = sub2(b[i], a[(i + 1) & 7]); a[i] = fma2(d, d, a[i]); }
      if (MODE >= 6 && MODE <= 9) {
        float lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo[i]), "=f"(hi[i]) : "l"(a[i]));
        if (MODE == 6) {
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = fminf(m[i], lo[i]);
        } else if (MODE == 7) {
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = __uint_as_float(min(__float_as_uint(m[i]), __float_as_uint(lo[i])));
        } else if (MODE == 8) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            m[i] = __uint_as_float(__vimin3_u32(__float_as_uint(m[i]), __float_as_uint(lo[i]), __float_as_uint(hi[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = __uint_as_float(__float_as_uint(m[i]) + __float_as_uint(lo[i]) + 3u);
        }
      }
      if (MODE >= 3 && MODE <= 5) {
        float lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo[i]), "=f"(hi[i]) : "l"(a[i]));
        m[0] = min3(m[0], lo[0], hi[0]); m[1] = min3(m[1], lo[1], hi[1]);
        m[2] = min3(m[2], lo[2], hi[2]); m[3] = min3(m[3], lo[3], hi[3]);
        if (MODE >= 4) { m[4] = min3(m[4], lo[4], hi[4]); m[5] = min3(m[5], lo[5], hi[5]); }
        if (MODE >= 5) {
          rx = redux(__float_as_uint(m[5]) ^ rx);
          ld = sm[(it + threadIdx.x) & 63];
          m[6] = min3(m[6], ld.x, ld.y);
        }
      }
    }
  }
  float acc = 0;
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32)) + m[i];
  for (int i = 0; i < 16; ++i) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + rx + ld.z;
}

template <int MODE>
void run(const char* name, int packed_per_iter, int scalar_per_iter, int other_per_iter, float* out, float4* src) {
  for (int threads : {512, 1024}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(out, src, 64);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, src, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int dev; cudaGetDevice(&dev); int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    double cycles = ms * 1e-3 * khz * 1e3;           // at the max SM clock (upper bound on cycles)
    double warps_per_smsp = threads / 32.0 / 4.0;
    double per_warp_iter = cycles / ITER / warps_per_smsp;  // SMSP cycles per warp-iteration
    printf("%-34s warps/SMSP=%4.1f  cyc/warp-iter=%7.2f  packed/cyc=%5.3f scalar/cyc=%5.3f all-inst/cyc=%5.3f\n", name, warps_per_smsp,
           per_warp_iter, packed_per_iter / per_warp_iter, scalar_per_iter / per_warp_iter,
           (packed_per_iter + scalar_per_iter + other_per_iter) / per_warp_iter);
  }
}

int main() {
  float* out; float4* src;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&src, 64 * 16); cudaMemset(src, 0, 64 * 16);
  run<0>("0: 16 FFMA2 (3 distinct regs)", 16, 0, 0, out, src);
  run<1>("1: 32 FFMA scalar", 0, 32, 0, out, src);
  run<2>("2: 8 FADD2 + 8 FFMA2(d,d,acc)", 16, 0, 0, out, src);
  run<3>("3: mode2 + 4 FMNMX3", 16, 0, 4, out, src);
  run<4>("4: mode2 + 6 FMNMX3", 16, 0, 6, out, src);
  run<5>("5: mode4 + REDUX + LDS + FMNMX3", 16, 0, 9, out, src);
  run<6>("6: mode2 + 8 FMNMX", 16, 0, 8, out, src);
  run<7>("7: mode2 + 8 IMNMX.U32", 16, 0, 8, out, src);
  run<8>("8: mode2 + 4 VIMNMX3.U32", 16, 0, 4, out, src);
  run<9>("9: mode2 + 8 IADD3", 16, 0, 8, out, src);
  run<13>("13: mode2 + 8 scalar FFMA", 16, 8, 0, out, src);
  run<14>("14: mode2 + 16 scalar FFMA", 16, 16, 0, out, src);
  run<10>("10: 16 FMNMX3 only", 0, 0, 16, out, src);
  run<11>("11: 16 VIMNMX3.U32 only", 0, 0, 16, out, src);
  run<12>("12: 16 FMNMX only", 0, 0, 16, out, src);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
