"""GPU parity, randomised: scripts/fuzz_parity.py for a bounded time with a fixed seed (random batch and cloud sizes,
tie grids, zero padding, duplicates) through Chamfer forward/backward, FPS, kNN and the graph-feature backward, each
against the CPU oracle - bit-exact indices and distances, 1e-5 gradients."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_parity_sweep(dev):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_parity.py"), "20", "77"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "failures 0" in r.stdout and int(r.stdout.split("cases")[1].split()[0]) > 50


def test_randomised_culled_knn_against_the_select_kernel(dev):
    """scripts/knn_stress.py for 15 s: random shapes inside the culled path's range, clustered / duplicated / padded
    clouds, 1 / 2 / 4 tasks per CTA, solo and grouped mode, odd seed counts - indices and distances must equal the select
    kernel's bit for bit, twice in a row (the helpers' deques and merges are the concurrent part)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "knn_stress.py"), "15", "5"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "no mismatch" in r.stdout and int(r.stdout.split("stress:")[1].split()[0]) > 20
