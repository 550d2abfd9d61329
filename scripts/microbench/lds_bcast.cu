// lds_bcast.cu - sm_100a microbenchmark: what does a warp-uniform (broadcast) LDS.128 cost next to FP32 math?
// The kNN select kernel issues one broadcast LDS.128 per ~8.5 other instructions; ncu shows 40 % short-scoreboard stalls.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu ; run: ./lds_bcast
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 2048;
// MODE 0: math only (NF FFMA per ref, operands in registers)
// MODE 1: one broadcast LDS.128 per ref + NF FFMA
// MODE 2: three broadcast LDS.32 per ref + NF FFMA
// MODE 3: one broadcast LDS.64 + one LDS.32 per ref + NF FFMA
// MODE 4: one LDS.128 per ref with lane-distinct, conflict-free addresses + NF FFMA
template <int MODE, int NF>
__global__ void __launch_bounds__(256) k(float* out, const float4* src, int iters) {
  __shared__ float4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = src[i];
  __syncthreads();
  float acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  const float* sf = reinterpret_cast<const float*>(sm);
  for (int it = 0; it < iters; ++it) {
    const int base = (it & 63) * 32;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      float x, y, z;
      if (MODE == 0) { x = acc[(r + 1) & 7]; y = acc[(r + 2) & 7]; z = acc[(r + 3) & 7]; }
      if (MODE == 1) { float4 p = sm[base + r]; x = p.x; y = p.y; z = p.z; }
      if (MODE == 2) { x = sf[4 * (base + r)]; y = sf[4 * (base + r) + 1]; z = sf[4 * (base + r) + 2]; }
      if (MODE == 3) { float2 p = *reinterpret_cast<const float2*>(sf + 4 * (base + r)); x = p.x; y = p.y; z = sf[4 * (base + r) + 2]; }
      if (MODE == 4) { float4 p = sm[(base + r * 32 + (threadIdx.x & 31)) & 2047]; x = p.x; y = p.y; z = p.z; }
      float a = acc[r & 7];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const float s = (f % 3 == 0) ? x : (f % 3 == 1) ? y : z;
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a) : "f"(s), "f"(s));
      }
      acc[r & 7] = a;
    }
  }
  float t = 0.f;
  for (int i = 0; i < 8; ++i) t += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int NF>
void run(const char* name, int ctas_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; float4* src;
  cudaMalloc(&out, sizeof(float) * 256 * sms * ctas_per_sm);
  cudaMalloc(&src, sizeof(float4) * 2048);
  cudaMemset(src, 0, sizeof(float4) * 2048);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, NF><<<sms * ctas_per_sm, 256>>>(out, src, 64);
  cudaEventRecord(e0);
  k<MODE, NF><<<sms * ctas_per_sm, 256>>>(out, src, ITER);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  // cycles per (ref x warp) per scheduler: each SM scheduler runs ctas_per_sm*8/4 warps
  const double refs_per_sched = (double)ITER * 32 * ctas_per_sm * 2;
  printf("%-44s ctas/SM=%d NF=%d  %.3f ms  %.2f cycles per ref-warp per scheduler (at %d MHz nominal)\n", name, ctas_per_sm, NF, ms,
         ms * 1e-3 * khz * 1e3 / refs_per_sched, khz / 1000);
  cudaFree(out); cudaFree(src);
}

int main() {
  for (int c = 1; c <= 3; ++c) {
    run<0, 8>("math only", c);
    run<1, 8>("broadcast LDS.128 + math", c);
    run<2, 8>("3x broadcast LDS.32 + math", c);
    run<3, 8>("broadcast LDS.64 + LDS.32 + math", c);
    run<4, 8>("lane-distinct LDS.128 + math", c);
  }
  run<1, 4>("broadcast LDS.128 + math", 3);
  run<1, 2>("broadcast LDS.128 + math", 3);
  run<1, 1>("broadcast LDS.128 + math", 3);
  run<4, 1>("lane-distinct LDS.128 + math", 3);
  return 0;
}
