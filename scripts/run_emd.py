"""Timing + agreement: auction EMD, ours vs the reference's unmodified extension (oracle/_ref/emd*.so) on the same GPU,
at the test loop's shape (utils/metrics.py:111-116: eps 0.005, 100 iterations, one 16384-point cloud pair)."""
import glob, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imgadapointr_b200.extensions.emd import emd_module as emd
from imgadapointr_b200.utils import synthetic
dev = torch.device("cuda:0")
hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "emd*.so"))
ref = None
if hits:
    spec = importlib.util.spec_from_file_location("emd", hits[0]); ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
def run_ref(x, y, eps, iters):
    B, n, _ = x.shape
    o = dict(device=dev)
    dist = torch.zeros(B, n, **o); asg = torch.zeros(B, n, dtype=torch.int32, **o) - 1
    inv = torch.zeros(B, n, dtype=torch.int32, **o) - 1; price = torch.zeros(B, n, **o)
    bid = torch.zeros(B, n, dtype=torch.int32, **o); inc = torch.zeros(B, n, **o); mx = torch.zeros(B, n, **o)
    ui = torch.zeros(B * n, dtype=torch.int32, **o); mi = torch.zeros(B * n, dtype=torch.int32, **o)
    c1, c2, c3 = (torch.zeros(512, dtype=torch.int32, **o) for _ in range(3))
    ref.forward(x, y, dist, asg, price, inv, bid, inc, mx, ui, c1, c2, c3, mi, eps, iters)
    return dist, asg
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for (B, n, kind) in [(1, 16384, "pcn"), (1, 16384, "uniform"), (8, 8192, "uniform"), (32, 2048, "uniform")]:
    if kind == "pcn":   # a completion-like pair: gt and a noisy resampling of it, normalised to [0, 1]
        gt = synthetic.pcn_gt(B, n, 5)
        pred = synthetic.pcn_pred(gt, n, 6) if hasattr(synthetic, "pcn_pred") else gt[:, torch.randperm(n)] + 0.01 * torch.randn(B, n, 3)
        lo = gt.amin((1, 2), keepdim=True); hi = gt.amax((1, 2), keepdim=True)
        x, y = ((pred - lo) / (hi - lo)).clamp(0, 1).to(dev), ((gt - lo) / (hi - lo)).to(dev)
    else:
        g = torch.Generator().manual_seed(n); x, y = torch.rand((B, n, 3), generator=g).to(dev), torch.rand((B, n, 3), generator=g).to(dev)
    d, a = emd.emdModule()(x, y, 0.005, 100)
    ms = t(lambda: emd.emdModule()(x, y, 0.005, 100))
    line = "B=%d n=%d %s: ours %.2f ms EMD %.5f unique %.3f" % (B, n, kind, ms, float(torch.sqrt(d).mean()), len(torch.unique(a[0])) / n)
    if ref is not None and n % 1024 == 0:
        rd, ra = run_ref(x, y, 0.005, 100)
        rms = t(lambda: run_ref(x, y, 0.005, 100))
        line += " | reference ext %.2f ms EMD %.5f | assignment agreement %.4f" % (rms, float(torch.sqrt(rd).mean()), float((ra == a).float().mean()))
    print(line, flush=True)
