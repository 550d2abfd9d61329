// chamfer_step.h - host-side argument block of the persistent Chamfer step kernel (chamfer_step.cu), shared with c_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace iapt {

struct StepArgs {
  const float* xyz1 = nullptr;  // [B,N,3]
  const float* xyz2 = nullptr;  // [B,M,3]
  int B = 0, N = 0, M = 0;
  bool search = true;           // run PAIR + RESOLVE; false: idx1/idx2 are inputs (backward-only launch)
  bool mask_zero_sum = false;   // ignore_zeros inside the kernel (search launches only)
  float* dist1 = nullptr;       // outputs of the search (a direction may be skipped in search-only launches)
  float* dist2 = nullptr;
  int32_t* idx1 = nullptr;
  int32_t* idx2 = nullptr;
  // loss / metric epilogue (search launches): any of them non-null enables the per-cloud partial sums
  float th = 0.f;
  int kind = 0;                 // IAPT_LOSS_* for `loss`
  float* sums = nullptr;        // [B,4]
  int32_t* counts = nullptr;    // [B,2]
  int32_t* valid = nullptr;     // [B,2]
  bool valid_is_input = false;  // gate on an earlier launch's `valid` (see IAPT_STEP_GATE_ON_VALID)
  float* loss = nullptr;        // [2]
  // gradient (GATHER): grad pointers non-null select the sides; gw = per-point weight source of each side
  int mode = 0;                 // 0: 2*gw[i];  1: 2*up*scale/sqrt(gw[i]);  2: 2*up*scale
  float scale1 = 1.f, scale2 = 1.f;
  const float* grad_out = nullptr;
  int grad_out_stride = 0;
  const float* gw1 = nullptr;
  const float* gw2 = nullptr;
  float* grad_xyz1 = nullptr;
  float* grad_xyz2 = nullptr;
  void* workspace = nullptr;
  bool stats = false;           // record device timestamps into the first 32 bytes of the workspace
};

size_t chamfer_step_workspace_bytes(int B, int N, int M);
size_t chamfer_forward_workspace_bytes(int B, int N, int M);
size_t chamfer_backward_workspace_bytes(int B, int N, int M);
int chamfer_step_launch(const StepArgs& a, cudaStream_t stream);
int scale_launch(const float* src, const float* factor, long long n, float* dst, cudaStream_t stream);
void profile_enable(bool on);
int profile_collect(double* total_ms, int* launches);

}  // namespace iapt
