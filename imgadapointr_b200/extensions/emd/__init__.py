"""Drop-in for the reference's ``extensions/emd`` package (auction EMD) on the B200 kernels."""
from . import emd_module  # noqa: F401
from .emd_module import emdFunction, emdModule  # noqa: F401
