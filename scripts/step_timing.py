"""Dev: time the one-launch Chamfer step (iapt_chamfer_step) at the BASELINE shapes, with the device-side split between
the search items and the rest, and check it against the separate entry points.
    python scripts/step_timing.py [B N M] ...   (default: the config 2 / 4 / 5 / 1 shapes)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imgadapointr_b200 import _lib, chamfer  # noqa: E402
from imgadapointr_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 148 * 128 * 2 * 1.965e9


def ev(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


shapes = [(48, 16384, 16384), (32, 16384, 16384), (24, 16384, 16384), (12, 16384, 16384), (6, 16384, 16384),
          (2, 65536, 65536), (1, 16384, 16384), (1, 2048, 16384), (48, 512, 16384), (48, 2048, 2048)]
if len(sys.argv) > 3:
    a = list(map(int, sys.argv[1:]))
    shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
for (B, N, M) in shapes:
    p, g = synthetic.pcn_pair(B, max(N, M), 5)
    p, g = p[:, :N].contiguous().to(dev), g[:, :M].contiguous().to(dev)
    t_fwd = ev(lambda: chamfer.forward(p, g))
    t_step = ev(lambda: chamfer.step(p, g, chamfer.LOSS_L1, grad1=True))
    t_step2 = ev(lambda: chamfer.step(p, g, chamfer.LOSS_L1, grad1=True, grad2=True))
    r = chamfer.step(p, g, chamfer.LOSS_L1, grad1=True, stats=True)
    search_ms, kern_ms = chamfer.step_stats(r.workspace)
    ends = " ".join("-" if e is None else f"{e:.3f}" for e in chamfer.step_phase_ends(r.workspace))
    d1, d2, i1, i2 = chamfer.forward(p, g)
    t_bwd = ev(lambda: chamfer.loss_backward(p, g, i1, i2, d1, d2, chamfer.LOSS_L1, None))
    ok = torch.equal(r.dist1, d1) and torch.equal(r.idx2, i2)
    fl = 16.0 * B * N * M
    print(f"B={B:3d} {N:5d}x{M:5d}  fwd {t_fwd:7.4f}  step(g1) {t_step:7.4f} ({fl / t_step / 1e-3 / PEAK:.3f} of fp32 peak)  "
          f"step(g1,g2) {t_step2:7.4f}  bwd-only {t_bwd:7.4f}  [device: search {search_ms:7.4f} kernel {kern_ms:7.4f}]  "
          f"same={ok}  phase ends [{ends}]", flush=True)

# ---- collapsed predictions (what an untrained network emits): every gt point picks one of 4 prediction points --------
gt = synthetic.pcn_gt(48, 16384, 303).to(dev)
for n in (512, 16384):
    p = torch.rand(48, 4, 3, device=dev).repeat(1, n // 4, 1).contiguous()
    t_step = ev(lambda: chamfer.step(p, gt, chamfer.LOSS_L1, grad1=True))
    r = chamfer.step(p, gt, chamfer.LOSS_L1, grad1=True, stats=True)
    ends = " ".join("-" if e is None else f"{e:.3f}" for e in chamfer.step_phase_ends(r.workspace))
    d1, d2, i1, i2 = chamfer.forward(p, gt)
    t_bwd = ev(lambda: chamfer.loss_backward(p, gt, i1, i2, d1, d2, chamfer.LOSS_L1, None))
    print(f"collapsed B=48 {n}x16384 (4 distinct): step(g1) {t_step:.4f}  bwd-only {t_bwd:.4f}  phase ends [{ends}]", flush=True)

# ---- knob sweep at config 2 (env is read per launch) -------------------------------------------------------------------
if os.environ.get("SWEEP", "1") == "1":
    B, N, M = 48, 16384, 16384
    p, g = synthetic.pcn_pair(B, N, 5)
    p, g = p.to(dev), g.to(dev)
    for (B, N, M, qts, splits) in [(48, 16384, 16384, (16, 8), (3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 20, 23, 26, 29, 32, 40, 43)),
                                   (6, 16384, 16384, (16, 8), (12, 24, 37, 49, 61, 74, 86, 98)),
                                   (12, 16384, 16384, (16, 8), (6, 12, 18, 24, 30, 37, 43, 49)),
                                   (32, 16384, 16384, (16,), (9, 16, 18, 23, 25, 27, 30, 32))]:
        p, g = synthetic.pcn_pair(B, N, 5)
        p, g = p.to(dev), g.to(dev)
        for qt in qts:
            for sp in splits:
                if qt == 8 and sp % 2 == 0:
                    sp //= 2
                elif qt == 8:
                    continue
                os.environ["IAPT_FWD_VARIANT"] = str(qt)
                os.environ["IAPT_S_SPLITS"] = str(sp)
                t = ev(lambda: chamfer.step(p, g, chamfer.LOSS_L1, grad1=True))
                r = chamfer.step(p, g, chamfer.LOSS_L1, grad1=True, stats=True)
                search_ms, kern_ms = chamfer.step_stats(r.workspace)
                items = B * (N // (256 * qt)) * sp
                print(f"B={B} qt={qt} s_splits={sp} items/slot={items / 296:.2f}: step(g1) {t:.4f} ms  [device: search {search_ms:.4f} kernel {kern_ms:.4f}]", flush=True)
    os.environ.pop("IAPT_FWD_VARIANT")
    os.environ.pop("IAPT_S_SPLITS")
    os.environ["IAPT_RC"] = os.environ["IAPT_LAG"] = "0"
    os.environ.pop("IAPT_RC")
    os.environ.pop("IAPT_LAG")
    # the module path (python + autograd overhead included), config 2 and config 1
    from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1
    mod = ChamferDistanceL1()
    for (B, N, M) in [(48, 16384, 16384), (1, 2048, 16384), (1, 16384, 16384)]:
        p, g = synthetic.pcn_pair(B, max(N, M), 5)
        p, g = p[:, :N].contiguous().to(dev), g[:, :M].contiguous().to(dev)

        def fb():
            x = p.detach().requires_grad_(True)
            mod(x, g).backward()

        print(f"module fwd+bwd B={B} {N}x{M}: {ev(fb, 5, 20):.4f} ms", flush=True)
