"""Builds the compiled PyTorch binding (torch_binding.cpp) IN-TREE as imgadapointr_b200/iapt_torch_ext*.so.

One g++ command against the installed torch's headers and libraries (torch.utils.cpp_extension supplies the paths);
the module links libiapt_b200.so next to it ($ORIGIN rpath) and owns no device code - the sm_100a kernels are all in
that library (csrc/Makefile).  Skips the build when the .so is newer than its inputs."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
NAME = "iapt_torch_ext"


def target():
    return os.path.join(PKG, NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False):
    import torch
    from torch.utils import cpp_extension as ce

    out = target()
    src = os.path.join(HERE, "torch_binding.cpp")
    deps = [src, os.path.join(ROOT, "include", "iapt_b200.h"), os.path.join(PKG, "libiapt_b200.so")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps[:2]):
        return out
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = ce.include_paths() + [os.path.join(cuda_home, "include"), sysconfig.get_paths()["include"],
                                os.path.join(ROOT, "include")]
    libdirs = ce.library_paths() + [os.path.join(cuda_home, "lib64"), PKG]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in inc] + [src, "-o", out] + [f"-L{p}" for p in libdirs]
    cmd += ["-liapt_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
            "-Wl,-rpath,$ORIGIN"] + [f"-Wl,-rpath,{p}" for p in ce.library_paths()]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
