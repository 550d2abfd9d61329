// torch_binding.cpp - the compiled PyTorch extension over the C ABI of libiapt_b200.so (include/iapt_b200.h).
//
// This is the binding the reference builds as the pybind module `chamfer` (extensions/chamfer_dist/chamfer_cuda.cpp:
// 22-39: forward -> [dist1, dist2, idx1, idx2], backward -> [grad_xyz1, grad_xyz2]) and what `pointnet2_ops._ext`
// (furthest_point_sampling / gather_points / gather_points_grad; call sites utils/misc.py:37-38,
// models/base_blocks.py:550-558) binds upstream, plus knn_point (models/Transformer_utils.py:19-30) and the fused loss
// step.  It owns NO kernels: every entry validates its tensors (TORCH_CHECK: CUDA, float32 / int32 / int64, contiguous,
// rank, sizes - the checks the reference leaves to data_ptr<float>()), guards the device (c10::cuda::CUDAGuard),
// allocates outputs and workspaces with at::empty on that device, takes at::cuda::getCurrentCUDAStream() and calls
// the C ABI with raw pointers.  A non-zero return code raises (the reference printf()s and continues,
// chamfer.cu:166-169).  The ctypes table in imgadapointr_b200/_lib.py stays as the torch-free way to the same ABI.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "iapt_b200.h"

namespace {

using at::Tensor;

void check_code(int code, const char* what) {
  TORCH_CHECK(code == 0, what, " failed (", code, "): ", iapt_error_string(code));
}

void check_cuda(const Tensor& t, const char* name, at::ScalarType dtype, int64_t dim) {
  TORCH_CHECK(t.defined(), name, " must be a tensor");
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (imgadapointr_b200 has no CPU path)");
  TORCH_CHECK(t.scalar_type() == dtype, name, " must be ", dtype, ", got ", t.scalar_type());
  TORCH_CHECK(t.dim() == dim, name, " must have ", dim, " dims, got ", t.dim());
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

void check_clouds(const Tensor& xyz1, const Tensor& xyz2) {
  check_cuda(xyz1, "xyz1", at::kFloat, 3);
  check_cuda(xyz2, "xyz2", at::kFloat, 3);
  TORCH_CHECK(xyz1.size(2) == 3 && xyz2.size(2) == 3, "xyz1 / xyz2 must be [B, n, 3]");
  TORCH_CHECK(xyz1.size(0) == xyz2.size(0), "xyz1 and xyz2 must have the same batch size");
  TORCH_CHECK(xyz1.device() == xyz2.device(), "xyz1 and xyz2 must be on the same device");
  TORCH_CHECK(xyz1.size(0) < (1LL << 31) && xyz1.size(1) < (1LL << 31) && xyz2.size(1) < (1LL << 31), "sizes exceed int");
}

Tensor workspace(size_t bytes, const Tensor& like) {
  return at::empty({(int64_t)(bytes > 8 ? bytes : 8)}, like.options().dtype(at::kByte));
}

template <typename T>
T* ptr(const Tensor& t) {
  return t.defined() ? t.data_ptr<T>() : nullptr;
}
template <typename T>
T* ptr(const c10::optional<Tensor>& t) {
  return (t.has_value() && t->defined()) ? t->data_ptr<T>() : nullptr;
}

iapt_stream_t stream_of(const Tensor& t) { return (iapt_stream_t)at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

// chamfer.forward (chamfer_cuda.cpp:22-25)
std::vector<Tensor> chamfer_forward(const Tensor& xyz1, const Tensor& xyz2) {
  check_clouds(xyz1, xyz2);
  const c10::cuda::CUDAGuard guard(xyz1.device());
  const int B = (int)xyz1.size(0), N = (int)xyz1.size(1), M = (int)xyz2.size(1);
  Tensor dist1 = at::empty({B, N}, xyz1.options()), dist2 = at::empty({B, M}, xyz1.options());
  Tensor idx1 = at::empty({B, N}, xyz1.options().dtype(at::kInt)), idx2 = at::empty({B, M}, xyz1.options().dtype(at::kInt));
  const size_t wb = iapt_chamfer_forward_workspace_bytes(B, N, M);
  Tensor ws = workspace(wb, xyz1);
  check_code(iapt_chamfer_forward(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), B, N, M, dist1.data_ptr<float>(),
                                  dist2.data_ptr<float>(), idx1.data_ptr<int32_t>(), idx2.data_ptr<int32_t>(),
                                  ws.data_ptr(), wb, stream_of(xyz1)),
             "iapt_chamfer_forward");
  return {dist1, dist2, idx1, idx2};
}

// chamfer.backward (chamfer_cuda.cpp:27-34); need1 / need2 = false skips that gradient (undefined tensor -> None)
std::vector<Tensor> chamfer_backward(const Tensor& xyz1, const Tensor& xyz2, const Tensor& idx1, const Tensor& idx2,
                                     const Tensor& grad_dist1, const Tensor& grad_dist2, bool need1, bool need2) {
  check_clouds(xyz1, xyz2);
  const int B = (int)xyz1.size(0), N = (int)xyz1.size(1), M = (int)xyz2.size(1);
  check_cuda(idx1, "idx1", at::kInt, 2);
  check_cuda(idx2, "idx2", at::kInt, 2);
  check_cuda(grad_dist1, "grad_dist1", at::kFloat, 2);
  check_cuda(grad_dist2, "grad_dist2", at::kFloat, 2);
  TORCH_CHECK(idx1.size(0) == B && idx1.size(1) == N && idx2.size(0) == B && idx2.size(1) == M &&
                  grad_dist1.size(0) == B && grad_dist1.size(1) == N && grad_dist2.size(0) == B && grad_dist2.size(1) == M,
              "backward: shape mismatch between xyz, idx and grad_dist");
  const c10::cuda::CUDAGuard guard(xyz1.device());
  Tensor gx1, gx2;
  if (!need1 && !need2) return {gx1, gx2};
  if (need1) gx1 = at::empty_like(xyz1);
  if (need2) gx2 = at::empty_like(xyz2);
  const size_t wb = iapt_chamfer_backward_workspace_bytes(B, N, M);
  Tensor ws = workspace(wb, xyz1);
  check_code(iapt_chamfer_backward(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), idx1.data_ptr<int32_t>(),
                                   idx2.data_ptr<int32_t>(), grad_dist1.data_ptr<float>(), grad_dist2.data_ptr<float>(), B,
                                   N, M, ptr<float>(gx1), ptr<float>(gx2), ws.data_ptr(), wb, stream_of(xyz1)),
             "iapt_chamfer_backward");
  return {gx1, gx2};
}

// The whole loss / metric step in one launch (iapt_chamfer_step).  Returns
// [dist1, dist2, idx1, idx2, sums, counts, valid, loss, grad1, grad2, workspace]; entries not requested are undefined.
std::vector<Tensor> chamfer_step(const Tensor& xyz1, const Tensor& xyz2, int64_t kind, int64_t flags, double th,
                                 bool want_sums, bool want_counts, bool want_valid,
                                 const c10::optional<Tensor>& gate_valid, bool grad1, bool grad2,
                                 const c10::optional<Tensor>& grad_loss) {
  check_clouds(xyz1, xyz2);
  const c10::cuda::CUDAGuard guard(xyz1.device());
  const int B = (int)xyz1.size(0), N = (int)xyz1.size(1), M = (int)xyz2.size(1);
  const auto f32 = xyz1.options(), i32 = xyz1.options().dtype(at::kInt);
  Tensor dist1 = at::empty({B, N}, f32), dist2 = at::empty({B, M}, f32);
  Tensor idx1 = at::empty({B, N}, i32), idx2 = at::empty({B, M}, i32);
  Tensor sums, counts, valid, loss, g1, g2;
  if (want_sums) sums = at::empty({B, 4}, f32);
  if (want_counts) counts = at::empty({B, 2}, i32);
  if (gate_valid.has_value() && gate_valid->defined()) {
    check_cuda(*gate_valid, "gate_valid", at::kInt, 2);
    TORCH_CHECK(gate_valid->size(0) == B && gate_valid->size(1) == 2, "gate_valid must be the [B,2] `valid` of an earlier step");
    valid = *gate_valid;
    flags |= IAPT_STEP_GATE_ON_VALID;
  } else if (want_valid) {
    valid = at::empty({B, 2}, i32);
  }
  if (kind != 0) loss = at::empty({2}, f32);
  if (grad1) g1 = at::empty_like(xyz1);
  if (grad2) g2 = at::empty_like(xyz2);
  if (grad_loss.has_value() && grad_loss->defined()) {
    TORCH_CHECK(grad_loss->is_cuda() && grad_loss->scalar_type() == at::kFloat && grad_loss->is_contiguous() &&
                    grad_loss->numel() == (kind == IAPT_LOSS_L2_SPLIT ? 2 : 1),
                "grad_loss must be a contiguous CUDA float32 tensor with 1 element (2 for L2_split)");
  }
  const size_t wb = iapt_chamfer_step_workspace_bytes(B, N, M);
  Tensor ws = workspace(wb, xyz1);
  check_code(iapt_chamfer_step(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), B, N, M, (int)kind, (int)flags, (float)th,
                               dist1.data_ptr<float>(), dist2.data_ptr<float>(), idx1.data_ptr<int32_t>(),
                               idx2.data_ptr<int32_t>(), ptr<float>(sums), ptr<int32_t>(counts), ptr<int32_t>(valid),
                               ptr<float>(loss), ptr<float>(grad_loss), ptr<float>(g1), ptr<float>(g2), ws.data_ptr(), wb,
                               stream_of(xyz1)),
             "iapt_chamfer_step");
  return {dist1, dist2, idx1, idx2, sums, counts, valid, loss, g1, g2, ws};
}

// Gradient of a fused loss from saved indices and distances (iapt_chamfer_backward_fused)
std::vector<Tensor> chamfer_loss_backward(const Tensor& xyz1, const Tensor& xyz2, const Tensor& idx1, const Tensor& idx2,
                                          const Tensor& dist1, const Tensor& dist2, int64_t kind,
                                          const c10::optional<Tensor>& grad_loss, bool need1, bool need2) {
  check_clouds(xyz1, xyz2);
  const int B = (int)xyz1.size(0), N = (int)xyz1.size(1), M = (int)xyz2.size(1);
  check_cuda(idx1, "idx1", at::kInt, 2);
  check_cuda(idx2, "idx2", at::kInt, 2);
  check_cuda(dist1, "dist1", at::kFloat, 2);
  check_cuda(dist2, "dist2", at::kFloat, 2);
  TORCH_CHECK(idx1.size(0) == B && idx1.size(1) == N && idx2.size(0) == B && idx2.size(1) == M && dist1.size(0) == B &&
                  dist1.size(1) == N && dist2.size(0) == B && dist2.size(1) == M,
              "loss_backward: shape mismatch between xyz, idx and dist");
  if (grad_loss.has_value() && grad_loss->defined()) {
    TORCH_CHECK(grad_loss->is_cuda() && grad_loss->scalar_type() == at::kFloat && grad_loss->is_contiguous() &&
                    grad_loss->numel() == (kind == IAPT_LOSS_L2_SPLIT ? 2 : 1),
                "grad_loss must be a contiguous CUDA float32 tensor with 1 element (2 for L2_split)");
  }
  const c10::cuda::CUDAGuard guard(xyz1.device());
  Tensor gx1, gx2;
  if (!need1 && !need2) return {gx1, gx2};
  if (need1) gx1 = at::empty_like(xyz1);
  if (need2) gx2 = at::empty_like(xyz2);
  const size_t wb = iapt_chamfer_backward_workspace_bytes(B, N, M);
  Tensor ws = workspace(wb, xyz1);
  check_code(iapt_chamfer_backward_fused(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), idx1.data_ptr<int32_t>(),
                                         idx2.data_ptr<int32_t>(), dist1.data_ptr<float>(), dist2.data_ptr<float>(),
                                         (int)kind, ptr<float>(grad_loss), B, N, M, ptr<float>(gx1), ptr<float>(gx2),
                                         ws.data_ptr(), wb, stream_of(xyz1)),
             "iapt_chamfer_backward_fused");
  return {gx1, gx2};
}

// dst = src * factor[0], factor read on the device (iapt_scale)
Tensor scale(const Tensor& src, const Tensor& factor) {
  TORCH_CHECK(src.is_cuda() && src.scalar_type() == at::kFloat && src.is_contiguous(), "scale: src must be contiguous CUDA float32");
  TORCH_CHECK(factor.is_cuda() && factor.scalar_type() == at::kFloat && factor.numel() >= 1 && factor.is_contiguous(),
              "scale: factor must be a contiguous CUDA float32 tensor");
  const c10::cuda::CUDAGuard guard(src.device());
  Tensor dst = at::empty_like(src);
  check_code(iapt_scale(src.data_ptr<float>(), factor.data_ptr<float>(), (long long)src.numel(), dst.data_ptr<float>(),
                        stream_of(src)),
             "iapt_scale");
  return dst;
}

// pointnet2_ops._ext.furthest_point_sampling(xyz, npoint) -> int32 [B, npoint]
Tensor furthest_point_sampling(const Tensor& xyz, int64_t npoint, int64_t flags) {
  check_cuda(xyz, "xyz", at::kFloat, 3);
  TORCH_CHECK(xyz.size(2) == 3, "xyz must be [B, N, 3]");
  TORCH_CHECK(npoint >= 0, "npoint must be non-negative");
  const c10::cuda::CUDAGuard guard(xyz.device());
  const int B = (int)xyz.size(0), N = (int)xyz.size(1);
  Tensor idx = at::empty({B, npoint}, xyz.options().dtype(at::kInt));
  const size_t wb = iapt_fps_workspace_bytes(B, N, (int)npoint);
  Tensor ws = workspace(wb, xyz);
  check_code(iapt_fps(xyz.data_ptr<float>(), B, N, (int)npoint, (int)flags, idx.data_ptr<int32_t>(), ws.data_ptr(), wb,
                      stream_of(xyz)),
             "iapt_fps");
  return idx;
}

// pointnet2_ops._ext.gather_points(features [B,C,N], idx [B,npoint]) -> [B,C,npoint]
Tensor gather_points(const Tensor& features, const Tensor& idx) {
  check_cuda(features, "features", at::kFloat, 3);
  check_cuda(idx, "idx", at::kInt, 2);
  TORCH_CHECK(idx.size(0) == features.size(0), "features / idx batch mismatch");
  const c10::cuda::CUDAGuard guard(features.device());
  const int B = (int)features.size(0), C = (int)features.size(1), N = (int)features.size(2), m = (int)idx.size(1);
  Tensor out = at::empty({B, C, m}, features.options());
  check_code(iapt_gather(features.data_ptr<float>(), idx.data_ptr<int32_t>(), B, C, N, m, out.data_ptr<float>(),
                         stream_of(features)),
             "iapt_gather");
  return out;
}

// pointnet2_ops._ext.gather_points_grad(grad_out [B,C,npoint], idx, N) -> [B,C,N]
Tensor gather_points_grad(const Tensor& grad_out, const Tensor& idx, int64_t N) {
  check_cuda(grad_out, "grad_out", at::kFloat, 3);
  check_cuda(idx, "idx", at::kInt, 2);
  TORCH_CHECK(idx.size(0) == grad_out.size(0) && idx.size(1) == grad_out.size(2), "grad_out / idx shape mismatch");
  const c10::cuda::CUDAGuard guard(grad_out.device());
  const int B = (int)grad_out.size(0), C = (int)grad_out.size(1), m = (int)grad_out.size(2);
  Tensor grad = at::empty({B, C, N}, grad_out.options());
  check_code(iapt_gather_grad(grad_out.data_ptr<float>(), idx.data_ptr<int32_t>(), B, C, (int)N, m, grad.data_ptr<float>(),
                              stream_of(grad_out)),
             "iapt_gather_grad");
  return grad;
}

// utils.misc.fps's gather without the transposes: xyz [B,N,3], idx [B,m] -> [B,m,3]
Tensor gather_xyz(const Tensor& xyz, const Tensor& idx) {
  check_cuda(xyz, "xyz", at::kFloat, 3);
  check_cuda(idx, "idx", at::kInt, 2);
  TORCH_CHECK(xyz.size(2) == 3 && idx.size(0) == xyz.size(0), "xyz must be [B,N,3] and idx [B,m]");
  const c10::cuda::CUDAGuard guard(xyz.device());
  const int B = (int)xyz.size(0), N = (int)xyz.size(1), m = (int)idx.size(1);
  Tensor out = at::empty({B, m, 3}, xyz.options());
  check_code(iapt_gather_xyz(xyz.data_ptr<float>(), idx.data_ptr<int32_t>(), B, N, m, out.data_ptr<float>(), stream_of(xyz)),
             "iapt_gather_xyz");
  return out;
}

// knn_point(nsample, xyz = ref, new_xyz = query) -> int64 [B,Q,k] (+ float32 distances when asked)
std::vector<Tensor> knn(const Tensor& ref, const Tensor& query, int64_t k, bool want_dist) {
  check_cuda(ref, "xyz", at::kFloat, 3);
  check_cuda(query, "new_xyz", at::kFloat, 3);
  TORCH_CHECK(ref.size(2) == 3 && query.size(2) == 3 && ref.size(0) == query.size(0), "xyz [B,R,3], new_xyz [B,Q,3]");
  TORCH_CHECK(ref.device() == query.device(), "xyz and new_xyz must be on the same device");
  const c10::cuda::CUDAGuard guard(ref.device());
  const int B = (int)ref.size(0), R = (int)ref.size(1), Q = (int)query.size(1);
  Tensor idx = at::empty({B, Q, k}, ref.options().dtype(at::kLong));
  Tensor dist;
  if (want_dist) dist = at::empty({B, Q, k}, ref.options());
  const size_t wb = iapt_knn_workspace_bytes(B, R, Q, (int)k);
  Tensor ws = workspace(wb, ref);
  check_code(iapt_knn(ref.data_ptr<float>(), query.data_ptr<float>(), B, R, Q, (int)k, idx.data_ptr<int64_t>(),
                      ptr<float>(dist), ws.data_ptr(), wb, stream_of(ref)),
             "iapt_knn");
  return {idx, dist};
}

// square_distance(src [B,N,C], dst [B,M,C]) -> [B,N,M] (models/Transformer_utils.py:32-51)
Tensor square_distance(const Tensor& src, const Tensor& dst) {
  check_cuda(src, "src", at::kFloat, 3);
  check_cuda(dst, "dst", at::kFloat, 3);
  TORCH_CHECK(src.size(0) == dst.size(0) && src.size(2) == dst.size(2), "src [B,N,C], dst [B,M,C]");
  TORCH_CHECK(src.device() == dst.device(), "src and dst must be on the same device");
  const c10::cuda::CUDAGuard guard(src.device());
  const int B = (int)src.size(0), N = (int)src.size(1), M = (int)dst.size(1), C = (int)src.size(2);
  Tensor out = at::empty({B, N, M}, src.options());
  check_code(iapt_square_distance(src.data_ptr<float>(), dst.data_ptr<float>(), B, N, M, C, out.data_ptr<float>(),
                                  stream_of(src)),
             "iapt_square_distance");
  return out;
}

// ---- the fused loss as a C++ autograd node ---------------------------------------------------------------------------------
// ChamferDistanceL1 / L2 / L2_split / L1_PM (extensions/chamfer_dist/__init__.py:28-103) as ONE differentiable op: forward =
// one iapt_chamfer_step launch that also forms the gradient for upstream 1 (when an input needs one), backward = iapt_scale
// with the upstream scalar.  Same arithmetic as the Python autograd.Function in extensions/chamfer_dist/__init__.py
// (FusedChamferLoss), without its per-call Python cost (~40 us of a 105 us module forward + backward).
struct FusedChamferLossFn : public torch::autograd::Function<FusedChamferLossFn> {
  static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, const Tensor& xyz1_in,
                                                const Tensor& xyz2_in, int64_t kind, bool need1, bool need2) {
    const bool one_launch = (need1 || need2) && kind != IAPT_LOSS_L2_SPLIT;
    const Tensor xyz1 = xyz1_in.detach().contiguous(), xyz2 = xyz2_in.detach().contiguous();
    auto r = chamfer_step(xyz1, xyz2, kind, 0, 0.0, false, false, false, c10::nullopt, one_launch && need1,
                          one_launch && need2, c10::nullopt);
    ctx->saved_data["kind"] = kind;
    ctx->saved_data["one_launch"] = one_launch;
    ctx->saved_data["need1"] = need1;
    ctx->saved_data["need2"] = need2;
    if (one_launch) {
      torch::autograd::variable_list keep;
      if (need1) keep.push_back(r[8]);
      if (need2) keep.push_back(r[9]);
      ctx->save_for_backward(keep);
    } else if (need1 || need2) {
      ctx->save_for_backward({xyz1, xyz2, r[2], r[3], r[0], r[1]});
    }
    const Tensor& loss = r[7];
    if (kind == IAPT_LOSS_L2_SPLIT) return {loss.select(0, 0), loss.select(0, 1)};
    return {loss.select(0, 0)};
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::variable_list grads) {
    const int64_t kind = ctx->saved_data["kind"].toInt();
    const bool one_launch = ctx->saved_data["one_launch"].toBool();
    const bool need1 = ctx->saved_data["need1"].toBool(), need2 = ctx->saved_data["need2"].toBool();
    const auto saved = ctx->get_saved_variables();
    Tensor gx1, gx2;
    if (one_launch) {
      const Tensor g = grads[0].to(at::kFloat).reshape({1}).contiguous();
      size_t i = 0;
      if (need1) gx1 = scale(saved[i++], g);
      if (need2) gx2 = scale(saved[i++], g);
    } else if (need1 || need2) {
      Tensor g;
      if (kind == IAPT_LOSS_L2_SPLIT) {
        const Tensor z = at::zeros({}, saved[4].options());
        g = at::stack({grads[0].defined() ? grads[0].to(at::kFloat) : z, grads[1].defined() ? grads[1].to(at::kFloat) : z})
                .contiguous();
      } else {
        g = grads[0].to(at::kFloat).reshape({1}).contiguous();
      }
      auto out = chamfer_loss_backward(saved[0], saved[1], saved[2], saved[3], saved[4], saved[5], kind, g, need1, need2);
      gx1 = out[0];
      gx2 = out[1];
    }
    return {gx1, gx2, Tensor(), Tensor(), Tensor()};
  }
};

std::vector<Tensor> fused_chamfer_loss(const Tensor& xyz1, const Tensor& xyz2, int64_t kind) {
  TORCH_CHECK(kind >= IAPT_LOSS_L1 && kind <= IAPT_LOSS_L1_PM, "kind must be one of IAPT_LOSS_*");
  // which gradients the step should form is decided here, where grad mode is visible (it is off inside forward)
  const bool on = at::GradMode::is_enabled();
  return FusedChamferLossFn::apply(xyz1, xyz2, kind, on && xyz1.requires_grad(), on && xyz2.requires_grad());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "Compiled PyTorch binding of libiapt_b200.so (C ABI, include/iapt_b200.h)";
  m.def("abi_version", []() { return iapt_abi_version(); });
  // the reference's `chamfer` module surface (chamfer_cuda.cpp:36-39)
  m.def("forward", &chamfer_forward, "Chamfer forward (CUDA)", py::arg("xyz1"), py::arg("xyz2"));
  m.def("backward", &chamfer_backward, "Chamfer backward (CUDA)", py::arg("xyz1"), py::arg("xyz2"), py::arg("idx1"),
        py::arg("idx2"), py::arg("grad_dist1"), py::arg("grad_dist2"), py::arg("need1") = true, py::arg("need2") = true);
  m.def("step", &chamfer_step);
  m.def("loss_backward", &chamfer_loss_backward);
  m.def("fused_chamfer_loss", &fused_chamfer_loss, "loss module as one differentiable op", py::arg("xyz1"), py::arg("xyz2"),
        py::arg("kind"));
  m.def("scale", &scale);
  // pointnet2_ops._ext surface
  m.def("furthest_point_sampling", &furthest_point_sampling, py::arg("xyz"), py::arg("npoint"), py::arg("flags") = 1);
  m.def("gather_points", &gather_points);
  m.def("gather_points_grad", &gather_points_grad);
  m.def("gather_xyz", &gather_xyz);
  m.def("square_distance", &square_distance);
  m.def("knn", &knn, py::arg("xyz"), py::arg("new_xyz"), py::arg("k"), py::arg("want_dist") = false);
}
