"""CUDA-graph replay of the fused Chamfer loss step for small / launch-bound problems.

The fused loss path is 1 memset + 1 kernel (+ 1 to apply the upstream gradient); for tiny problems even that is
shorter than the host cost of issuing it from Python.  ``GraphedChamferLoss`` captures search + loss reduction + analytic
gradient once per (kind, B, N, M, device) into a ``torch.cuda.CUDAGraph`` (every launch of libiapt_b200.so goes to
torch's current stream, so stream capture sees them) and replays it with one ``cudaGraphLaunch``.  Values and gradients
are bit-identical to ``ChamferDistanceL1/L2/L2_split/L1_PM`` (same kernels); it is an optimisation of HOW the kernels
are issued, with the same ``forward(xyz1, xyz2)`` signature, differentiable w.r.t. both clouds.
"""
from __future__ import annotations

import torch

from . import chamfer

_KINDS = {"L1": chamfer.LOSS_L1, "L2": chamfer.LOSS_L2, "L2_split": chamfer.LOSS_L2_SPLIT, "L1_PM": chamfer.LOSS_L1_PM}


class _Captured:
    def __init__(self, kind, B, N, M, dev):
        self.x1 = torch.zeros((B, N, 3), dtype=torch.float32, device=dev)
        self.x2 = torch.zeros((B, M, 3), dtype=torch.float32, device=dev)
        n_out = 2 if kind == chamfer.LOSS_L2_SPLIT else 1
        self.unit = torch.ones((n_out,), dtype=torch.float32, device=dev)  # upstream gradient 1 (scaled in backward)
        # unit upstream gradients of the two L2_split outputs (created BEFORE capture: no host copies inside a graph)
        self.e0 = torch.tensor([1.0, 0.0], dtype=torch.float32, device=dev)
        self.e1 = torch.tensor([0.0, 1.0], dtype=torch.float32, device=dev)
        self.kind = kind
        self.x1.uniform_(-0.5, 0.5)
        self.x2.uniform_(-0.5, 0.5)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture: one-time kernel attribute configuration happens here
            for _ in range(2):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    def _run(self):
        if self.kind == chamfer.LOSS_L2_SPLIT:
            # the two outputs have independent upstream gradients: keep the per-output gradients separately
            r = chamfer.step(self.x1, self.x2, self.kind, grad1=True, grad2=True, grad_loss=self.e0)
            self.g1a, self.g2a = r.grad1, r.grad2
            self.g1b, self.g2b = chamfer.loss_backward(self.x1, self.x2, r.idx1, r.idx2, r.dist1, r.dist2, self.kind,
                                                       self.e1)
        else:
            r = chamfer.step(self.x1, self.x2, self.kind, grad1=True, grad2=True)
            self.g1, self.g2 = r.grad1, r.grad2
        self.loss = r.loss


class _GraphedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, cap):
        cap.x1.copy_(xyz1)
        cap.x2.copy_(xyz2)
        cap.graph.replay()
        if cap.kind == chamfer.LOSS_L2_SPLIT:
            ctx.save_for_backward(cap.g1a.clone(), cap.g2a.clone(), cap.g1b.clone(), cap.g2b.clone())
            ctx.split = True
            return cap.loss[0].clone(), cap.loss[1].clone()
        ctx.save_for_backward(cap.g1.clone(), cap.g2.clone())
        ctx.split = False
        return cap.loss[0].clone()

    @staticmethod
    def backward(ctx, *grads):
        if ctx.split:
            g1a, g2a, g1b, g2b = ctx.saved_tensors
            ga = grads[0] if grads[0] is not None else torch.zeros((), device=g1a.device)
            gb = grads[1] if grads[1] is not None else torch.zeros((), device=g1a.device)
            return ga * g1a + gb * g1b, ga * g2a + gb * g2b, None
        g1, g2 = ctx.saved_tensors
        return grads[0] * g1, grads[0] * g2, None


class GraphedChamferLoss(torch.nn.Module):
    """``GraphedChamferLoss("L1")(pred, gt)`` == ``ChamferDistanceL1()(pred, gt)`` issued as one graph launch.
    Shapes are cached per (B, N, M, device); ``ignore_zeros`` filtering changes shapes per call and is not supported."""

    def __init__(self, kind: str = "L1", max_cached: int = 8):
        super().__init__()
        if kind not in _KINDS:
            raise ValueError(f"kind must be one of {sorted(_KINDS)}")
        self.kind = _KINDS[kind]
        self.max_cached = max_cached
        self._cache = {}

    def forward(self, xyz1, xyz2):
        key = (xyz1.size(0), xyz1.size(1), xyz2.size(1), xyz1.device)
        cap = self._cache.get(key)
        if cap is None:
            if len(self._cache) >= self.max_cached:
                self._cache.pop(next(iter(self._cache)))
            cap = self._cache[key] = _Captured(self.kind, key[0], key[1], key[2], key[3])
        return _GraphedFn.apply(xyz1, xyz2, cap)
