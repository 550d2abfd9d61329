"""CPU: the product package never touches the oracle (or the reference), and ships no CPU fallback."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "imgadapointr_b200")


def _sources():
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                yield os.path.join(d, f)


def test_product_does_not_import_oracle_or_reference():
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|/root/reference|oracle/_ref", re.M)
    for path in _sources():
        assert not bad.search(open(path).read()), path


def test_importing_product_does_not_load_oracle():
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r); import imgadapointr_b200, imgadapointr_b200.chamfer, "
            "imgadapointr_b200.utils.metrics, imgadapointr_b200.utils.misc, imgadapointr_b200.models.Transformer_utils; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_no_forbidden_toolchains():
    pat = re.compile(r"^\s*(import|from)\s+(triton|tilelang)\b|torch\.compile\(", re.M)
    for path in _sources():
        assert not pat.search(open(path).read()), path
