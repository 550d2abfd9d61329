"""Dev: time chamfer.forward for the variant selected by IAPT_FWD_VARIANT and check parity vs the reference .so."""
import os, sys, glob, importlib.util
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200 import chamfer, _lib
from imgadapointr_b200.utils import synthetic

dev = torch.device("cuda:0")
so = glob.glob(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "chamfer*.so"))
spec = importlib.util.spec_from_file_location("chamfer", so[0]); ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
ok = True
for (B, N, M) in [(2, 700, 1300), (1, 2048, 16384), (2, 8192, 8192), (1, 16384, 16384), (3, 4097, 5000)]:
    p, g = synthetic.pcn_pair(B, max(N, M), 5)
    a, b = p[:, :N].contiguous().to(dev), g[:, :M].contiguous().to(dev)
    r = ref.forward(a, b); torch.cuda.synchronize()
    o = chamfer.forward(a, b)
    good = all(torch.equal(x, y) for x, y in zip(o, r))
    ok &= good
    if not good:
        print("MISMATCH", B, N, M, [int((x != y).sum()) for x, y in zip(o, r)])
B = 48
pred, gt = synthetic.pcn_pair(B, 16384, 202)
p, g = pred.to(dev), gt.to(dev)
for _ in range(3): chamfer.forward(p, g)
torch.cuda.synchronize()
_lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): chamfer.forward(p, g)
e1.record(); torch.cuda.synchronize()
ms, n = _lib.profile_collect()
print(f"variant={os.environ.get('IAPT_FWD_VARIANT','default')} parity={'OK' if ok else 'FAIL'} fwd_total_ms={e0.elapsed_time(e1)/10:.4f} pair_kernel_ms={ms/n:.4f} frac={16*B*16384*16384/(ms/n*1e-3)/74.45e12:.4f}")
