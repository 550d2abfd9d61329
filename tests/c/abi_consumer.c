/* A plain-C consumer of libiapt_b200.so: no Python, no torch, no C++ - just the CUDA runtime API and the header.
 * Computes the Chamfer search + gradient and FPS + kNN of a tiny hand-checkable cloud and prints the results.
 * Built and run by tests/test_c_consumer.py. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "iapt_b200.h"

#define CK(x)                                                        \
  do {                                                               \
    int rc_ = (int)(x);                                              \
    if (rc_ != 0) {                                                  \
      fprintf(stderr, "FAIL %s -> %d (%s)\n", #x, rc_, iapt_error_string(rc_)); \
      return 1;                                                      \
    }                                                                \
  } while (0)

int main(void) {
  /* xyz1: 4 points on the x axis at 0,1,2,3;  xyz2: 3 points at 0.1, 2.2, 10 */
  const float h1[12] = {0, 0, 0, 1, 0, 0, 2, 0, 0, 3, 0, 0};
  const float h2[9] = {0.1f, 0, 0, 2.2f, 0, 0, 10, 0, 0};
  const int B = 1, N = 4, M = 3;
  float *x1, *x2, *d1, *d2, *g1, *g2, *gx1, *gx2;
  int32_t *i1, *i2, *fi;
  int64_t* ki;
  void* ws;
  printf("abi %d\n", iapt_abi_version());
  CK(cudaMalloc((void**)&x1, sizeof h1));
  CK(cudaMalloc((void**)&x2, sizeof h2));
  CK(cudaMalloc((void**)&d1, 4 * N));
  CK(cudaMalloc((void**)&d2, 4 * M));
  CK(cudaMalloc((void**)&i1, 4 * N));
  CK(cudaMalloc((void**)&i2, 4 * M));
  CK(cudaMalloc((void**)&g1, 4 * N));
  CK(cudaMalloc((void**)&g2, 4 * M));
  CK(cudaMalloc((void**)&gx1, 12 * N));
  CK(cudaMalloc((void**)&gx2, 12 * M));
  CK(cudaMalloc((void**)&fi, 4 * 3));
  CK(cudaMalloc((void**)&ki, 8 * N * 2));
  CK(cudaMemcpy(x1, h1, sizeof h1, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(x2, h2, sizeof h2, cudaMemcpyHostToDevice));
  size_t wsb = iapt_chamfer_forward_workspace_bytes(B, N, M);
  CK(cudaMalloc(&ws, wsb));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CK(iapt_chamfer_forward(x1, x2, B, N, M, d1, d2, i1, i2, ws, wsb, st));
  const float ones[4] = {1, 1, 1, 1};
  CK(cudaMemcpyAsync(g1, ones, 4 * N, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g2, ones, 4 * M, cudaMemcpyHostToDevice, st));
  void* bws;
  size_t bwsb = iapt_chamfer_backward_workspace_bytes(B, N, M);
  CK(cudaMalloc(&bws, bwsb));
  CK(iapt_chamfer_backward(x1, x2, i1, i2, g1, g2, B, N, M, gx1, gx2, bws, bwsb, st));
  CK(iapt_fps(x1, B, N, 3, 0, fi, NULL, 0, st));
  size_t kwb = iapt_knn_workspace_bytes(B, N, N, 2);
  void* kws = NULL;
  if (kwb) CK(cudaMalloc(&kws, kwb));
  CK(iapt_knn(x1, x1, B, N, N, 2, ki, NULL, kws, kwb, st));
  CK(cudaStreamSynchronize(st));
  float hd1[4], hd2[3], hg1[12];
  int32_t hi1[4], hi2[3], hf[3];
  int64_t hk[8];
  CK(cudaMemcpy(hd1, d1, sizeof hd1, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hd2, d2, sizeof hd2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hi1, i1, sizeof hi1, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hi2, i2, sizeof hi2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hg1, gx1, sizeof hg1, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hf, fi, sizeof hf, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hk, ki, sizeof hk, cudaMemcpyDeviceToHost));
  printf("idx1 %d %d %d %d\n", hi1[0], hi1[1], hi1[2], hi1[3]);
  printf("idx2 %d %d %d\n", hi2[0], hi2[1], hi2[2]);
  printf("dist1 %.6f %.6f %.6f %.6f\n", hd1[0], hd1[1], hd1[2], hd1[3]);
  printf("dist2 %.6f %.6f %.6f\n", hd2[0], hd2[1], hd2[2]);
  printf("gx1x %.6f %.6f %.6f %.6f\n", hg1[0], hg1[3], hg1[6], hg1[9]);
  printf("fps %d %d %d\n", hf[0], hf[1], hf[2]);
  printf("knn %lld %lld | %lld %lld | %lld %lld | %lld %lld\n", (long long)hk[0], (long long)hk[1], (long long)hk[2],
         (long long)hk[3], (long long)hk[4], (long long)hk[5], (long long)hk[6], (long long)hk[7]);
  printf("launches %llu\n", iapt_launch_count());
  return 0;
}
