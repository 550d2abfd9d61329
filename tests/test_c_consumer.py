"""The C ABI is usable from plain C: tests/c/abi_consumer.c includes only include/iapt_b200.h and the CUDA runtime API.
CPU: it compiles with gcc as C11 (-Wall -Werror) and links against libiapt_b200.so.  GPU: it runs and prints the
hand-checkable answers."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_consumer.c")
EXE = os.path.join(ROOT, "tests", "c", "abi_consumer")
LIBDIR = os.path.join(ROOT, "imgadapointr_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build():
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           SRC, "-o", EXE, "-L", LIBDIR, "-liapt_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
           f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_c_consumer_compiles_and_links_as_plain_c():
    _build()
    assert os.path.exists(EXE)
    out = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libiapt_b200.so" in out and "libtorch" not in out and "libpython" not in out


@pytest.mark.gpu
def test_c_consumer_runs_on_gpu(dev):
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    assert lines["abi"] == "2"
    assert lines["idx1"] == "0 0 1 1"          # 0->0.1, 1->0.1 (0.81 < 1.44), 2->2.2, 3->2.2
    assert lines["idx2"] == "0 2 3"            # 0.1->0, 2.2->2, 10->3
    d1 = [float(x) for x in lines["dist1"].split()]
    assert abs(d1[0] - 0.01) < 1e-6 and abs(d1[1] - 0.81) < 1e-6 and abs(d1[2] - 0.04) < 1e-6 and abs(d1[3] - 0.64) < 1e-6
    g = [float(x) for x in lines["gx1x"].split()]
    # grad_x1[0] = 2(0-0.1) - 2(0.1-0) = -0.4 ; [1] = 2(1-0.1) = 1.8 ; [2] = 2(2-2.2) - 2(2.2-2) = -0.8 ; [3] = 2(3-2.2) - 2(10-3) = -12.4
    for got, want in zip(g, (-0.4, 1.8, -0.8, -12.4)):
        assert abs(got - want) < 1e-5
    assert lines["fps"] == "0 3 1"             # start 0, farthest 3, then 1 (ties 1 vs 2 -> lowest index)
    assert lines["knn"] == "0 1 | 1 0 | 2 1 | 3 2"
    assert int(lines["launches"]) >= 4
