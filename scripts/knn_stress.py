"""Randomised stress of the culled kNN path against the select kernel (itself checked against the oracle by the tests):
random shapes inside the culled range, clustered / duplicated / padded clouds, every tasks-per-CTA variant, solo and
grouped mode.  The deques and the helpers' merge are the concurrent part of that kernel: any race shows up as an index or
distance mismatch.  Usage: python scripts/knn_stress.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.utils import synthetic


def make_cloud(rng, B, R, kind):
    g = torch.Generator().manual_seed(int(rng.integers(0, 2**31)))
    if kind == 0:
        x = torch.rand((B, R, 3), generator=g) - 0.5
    elif kind == 1:
        x = synthetic.pcn_gt(B, R, int(rng.integers(0, 10000)))
    elif kind == 2:  # zero padding
        x = synthetic.pcn_gt(B, R, int(rng.integers(0, 10000)))
        x[:, int(R * rng.uniform(0.2, 0.95)):] = 0.0
    elif kind == 3:  # a few tight clusters
        c = torch.rand((B, 6, 3), generator=g) - 0.5
        x = c[:, torch.randint(0, 6, (R,), generator=g)] + 1e-3 * torch.randn((B, R, 3), generator=g)
    elif kind == 4:  # heavy duplication
        base = torch.rand((B, 40, 3), generator=g) - 0.5
        x = base[:, torch.randint(0, 40, (R,), generator=g)]
    else:  # quantised grid: exact ties between different positions
        x = torch.round((torch.rand((B, R, 3), generator=g) - 0.5) * 16) / 16
    return x.contiguous()


def run(seconds=60.0, seed=1, verbose=True):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    t0, n = time.time(), 0
    while time.time() - t0 < seconds:
        B = int(rng.choice([1, 2, 3, 5, 8, 17]))
        R = int(rng.integers(1024, 4097))
        kind = int(rng.integers(0, 6))
        ref = make_cloud(rng, B, R, kind).to(dev)
        mode = int(rng.integers(0, 3))
        if mode == 0:
            qry = ref  # self-query (one binning serves both sides)
        elif mode == 1:
            Q = int(rng.integers(1, 4097))
            qry = (torch.rand((B, Q, 3)) - 0.5).to(dev)
        else:
            Q = int(rng.integers(1, R + 1))
            qry = ref[:, torch.randperm(R)[:Q]].contiguous()
        k = int(rng.integers(1, 17))
        os.environ["IAPT_KNN_PRUNED"] = "2"
        want_i, want_d = knn_point(k, ref, qry, return_dist=True)
        os.environ["IAPT_KNN_PRUNED"] = "1"
        os.environ["IAPT_KNN_CULL_GROUP"] = str(int(rng.choice([1, 2, 4])))
        os.environ["IAPT_KNN_CULL_SOLO_BELOW"] = str(int(rng.choice([1, 900, 1 << 30])))
        os.environ["IAPT_KNN_SEEDS"] = str(int(rng.choice([1, 3, 12, 32])))
        os.environ["IAPT_KNN_PROBES"] = str(int(rng.choice([1, 2])))
        for rep in range(2):
            got_i, got_d = knn_point(k, ref, qry, return_dist=True)
            ok = torch.equal(got_i, want_i) and torch.equal(got_d, want_d)
            assert ok, "kNN mismatch: B=%d R=%d Q=%d k=%d kind=%d mode=%d env=%s" % (
                B, R, qry.shape[1], k, kind, mode, {e: os.environ[e] for e in os.environ if e.startswith("IAPT_KNN")})
        n += 1
    for e in ("IAPT_KNN_PRUNED", "IAPT_KNN_CULL_GROUP", "IAPT_KNN_CULL_SOLO_BELOW", "IAPT_KNN_SEEDS", "IAPT_KNN_PROBES"):
        os.environ.pop(e, None)
    if verbose:
        print("knn stress: %d cases in %.0f s, no mismatch" % (n, time.time() - t0))
    return n


if __name__ == "__main__":
    run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
