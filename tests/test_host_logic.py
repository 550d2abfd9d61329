"""CPU: host-side logic - batch sharding, the single all-reduce (gloo, world_size 2), metric accumulation, drop-in
module registration, synthetic generators."""
import os
import subprocess
import sys

import pytest
import torch

from imgadapointr_b200.utils import dist_utils, synthetic
from imgadapointr_b200.utils.metrics import MetricAccumulator, Metrics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 48, 1200, 1201):
        for world in (1, 2, 3, 4, 8):
            spans = [dist_utils.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_metric_accumulator_macro_average():
    acc = MetricAccumulator(3, torch.device("cpu"))
    acc.update(0, torch.tensor(0.5), torch.tensor(10.0), torch.tensor(1.0))
    acc.update(0, torch.tensor(1.0), torch.tensor(20.0), torch.tensor(3.0))
    acc.update_batch(torch.tensor([2, 2]), torch.tensor([0.2, 0.4]), torch.tensor([1.0, 3.0]), torch.tensor([5.0, 7.0]))
    per = acc.per_category()
    assert torch.allclose(per[0], torch.tensor([0.75, 15.0, 2.0], dtype=torch.float64))
    assert torch.allclose(per[2], torch.tensor([0.3, 2.0, 6.0], dtype=torch.float64))
    # category 1 is empty and must not enter the macro average (runner.py:676-678 averages existing categories)
    assert torch.allclose(acc.overall(), torch.tensor([0.525, 8.5, 4.0], dtype=torch.float64))


def test_metrics_class_api():
    assert Metrics.names() == ['F-Score', 'CDL1', 'CDL2', 'EMDistance']
    a = Metrics('CDL1', {'F-Score': 0.8, 'CDL1': 6.5, 'CDL2': 0.2, 'EMDistance': 0})
    b = Metrics('CDL1', [0.7, 7.0, 0.3, 0])
    assert a.better_than(b) and not b.better_than(a) and a.better_than(None)
    f = Metrics('F-Score', [0.9, 9.0, 0.3, 0])
    assert f.better_than(Metrics('F-Score', [0.7, 7.0, 0.3, 0]))
    assert a.state_dict()['CDL1'] == 6.5
    with pytest.raises(Exception):
        Metrics('nope', [0, 0, 0, 0]).better_than(b)


_WORKER = r"""
import os, sys, torch
sys.path.insert(0, %r)
import torch.distributed as dist
from imgadapointr_b200.utils import dist_utils
from imgadapointr_b200.utils.metrics import MetricAccumulator
rank, world = dist_utils.init_dist('pytorch', backend='gloo')
assert world == 2
# 11 clouds in 3 categories, sharded contiguously; every rank must end with the same global table
cats = torch.tensor([0, 0, 1, 2, 2, 2, 1, 0, 2, 1, 0])
vals = torch.arange(11, dtype=torch.float32)
b, e = dist_utils.shard_range(11, rank, world)
acc = MetricAccumulator(3, torch.device('cpu'))
acc.update_batch(cats[b:e], vals[b:e] / 10, vals[b:e], vals[b:e] * 2)
acc.reduce()
full = MetricAccumulator(3, torch.device('cpu'))
full.update_batch(cats, vals / 10, vals, vals * 2)
assert torch.allclose(acc.table, full.table), (rank, acc.table, full.table)
t = dist_utils.reduce_tensor(torch.tensor(float(rank + 1)))
assert float(t) == 1.5
g = dist_utils.gather_tensor(torch.tensor([float(rank)]))
assert g.tolist() == [0.0, 1.0]
s = dist_utils.allreduce_sums(torch.tensor([1.0, float(e - b)]))
assert s.tolist() == [2.0, 11.0]
# the deferred form bench.py uses: issue now, consume one step later - same value as reduce_tensor
pending = None
seen = []
for k in range(4):
    if pending is not None:
        seen.append(float(dist_utils.finish_reduce(*pending)))
    pending = dist_utils.reduce_tensor_async(torch.tensor(float(rank + 1 + k)))
seen.append(float(dist_utils.finish_reduce(*pending)))
assert seen == [1.5, 2.5, 3.5, 4.5], seen
dist.destroy_process_group()
print('rank', rank, 'ok')
"""


def test_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_reduce_tensor_async_is_identity_without_a_process_group():
    from imgadapointr_b200.utils import dist_utils

    rt, work = dist_utils.reduce_tensor_async(torch.tensor(3.0))
    assert float(dist_utils.finish_reduce(rt, work)) == 3.0


def test_dropin_registers_reference_module_names():
    code = ("import sys; sys.path.insert(0, %r); from imgadapointr_b200 import dropin; dropin.install(); "
            "import chamfer; from pointnet2_ops import pointnet2_utils; import extensions.chamfer_dist as cd; "
            "assert callable(chamfer.forward) and callable(chamfer.backward); "
            "assert hasattr(pointnet2_utils, 'furthest_point_sample') and hasattr(pointnet2_utils, 'gather_operation'); "
            "assert all(hasattr(cd, n) for n in ['ChamferFunction','ChamferDistanceL1','ChamferDistanceL2',"
            "'ChamferDistanceL2_split','ChamferDistanceL1_PM'])" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_pointnet2_cold_ops_have_no_cpu_path():
    from imgadapointr_b200._lib import IaptError
    from imgadapointr_b200.pointnet2_ops import pointnet2_utils as pu

    x = torch.zeros((1, 8, 3))
    f = torch.zeros((1, 2, 8))
    i = torch.zeros((1, 4, 3), dtype=torch.int32)
    for call in (lambda: pu.three_nn(x, x), lambda: pu.three_interpolate(f, i, torch.zeros(1, 4, 3)),
                 lambda: pu.grouping_operation(f, i), lambda: pu.ball_query(0.1, 4, x, x)):
        with pytest.raises(IaptError):
            call()


def test_synthetic_generators_are_deterministic_and_shaped():
    a, b = synthetic.pcn_pair(2, 512, 5)
    a2, b2 = synthetic.pcn_pair(2, 512, 5)
    assert torch.equal(a, a2) and torch.equal(b, b2) and a.shape == (2, 512, 3)
    assert float(b.abs().max()) <= 0.5
    part = synthetic.pcn_partial(b, 256, 6)
    assert part.shape == (2, 256, 3)
    g = synthetic.grid(1, 100, 1, bits=4, dup_frac=0.2, zero_frac=0.1)
    assert torch.equal(g * 16, torch.round(g * 16)) and not g[0, 90:].any()
