"""Dev: search time of the Chamfer step against the number of streamed-side splits (IAPT_S_SPLITS) and the tile variant.
    python scripts/split_sweep.py B N M qt s0 s1 [step]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imgadapointr_b200 import chamfer  # noqa: E402
from imgadapointr_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda:0")
B, N, M, qt, s0, s1 = map(int, sys.argv[1:7])
st = int(sys.argv[7]) if len(sys.argv) > 7 else 1
p, g = synthetic.pcn_pair(B, max(N, M), 5)
p, g = p[:, :N].contiguous().to(dev), g[:, :M].contiguous().to(dev)
os.environ["IAPT_FWD_VARIANT"] = str(qt)
tiles = -(-max(N, M) // (256 * qt))
for s in range(s0, s1 + 1, st):
    os.environ["IAPT_S_SPLITS"] = str(s)
    ts = []
    for _ in range(9):
        r = chamfer.step(p, g, chamfer.LOSS_L1, grad1=True, stats=True)
        ts.append(chamfer.step_stats(r.workspace)[0])
    ts.sort()
    print(f"B={B} qt={qt} s={s:3d} items/SM={B * tiles * s / 148:7.2f}  search min {ts[0]:.4f} med {ts[4]:.4f}", flush=True)
