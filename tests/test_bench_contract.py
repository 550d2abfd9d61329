"""CPU: the bench.py contract that can be checked without a GPU - the reference arm prints the required JSON line
(metric/unit/config of our arm, impl=reference, cpu_baseline, e2e with zero copy bytes), non-zero ranks of a torchrun
launch stay silent, and our arm refuses to run (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "chamfer_l1_fwd_bwd_point_pairs_per_s"
    assert line["unit"] == "point-pairs/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 1e6 and line["steps"] == 1 and line["dtype"] == "f32" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]


def test_reference_arm_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_our_arm_has_no_cpu_fallback():
    out = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "3"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)


@pytest.mark.gpu
def test_our_arm_json_line_on_gpu(dev):
    """The ours-arm line carries every key of the contract: device-timed value, e2e with its copy bytes, the roofline of
    the one kernel, the strong-scaling block, launches counted from our library, clocks sampled during the run."""
    out = subprocess.run([sys.executable, BENCH, "--steps", "3", "--warmup", "3", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["metric"] == "chamfer_l1_fwd_bwd_point_pairs_per_s" and line["n_gpus"] == 1 and line["steps"] == 3
    assert line["scaling"] == "weak" and line["dtype"] == "f32" and line["vs_baseline"] is None
    assert line["value"] > 1e12 and 0.5 < line["frac_fp32_peak"] < 1.0
    assert line["gpu_launches"] >= 3  # at least one kernel of ours per step
    r = line["roofline"]
    assert r["bound"] == "fp32_fma" and 0.5 < r["frac"] < 1.0 and r["achieved"] < r["peak"] and r["peak_measured"] < r["peak"]
    assert r["kernel_launches_timed"] == 3 and 0.9 < r["kernel_share_of_step"] <= 1.0
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * 48 * 16384 * 12 and e["d2h_bytes_per_step"] == 4 and 0 < e["value"] < line["value"]
    s = line["strong"]
    for key in ("config2_B48_16384_L1", "config5_B256_16384_L2", "config5_B16_65536_L2", "config4_sweep_1200_clouds"):
        assert 0.3 < s[key]["frac_fp32_peak_per_gpu"] < 1.0, key
    assert line["clocks"]["sm_mhz"] and "reasons" in line["clocks"]
    assert line["secondary"]["fps_2048_to_512"] > 1e4 and line["secondary"]["knn_k16_2048x2048"] > 1e4
