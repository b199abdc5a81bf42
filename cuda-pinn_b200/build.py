"""Build libpinn_b200.so (the C-ABI shared library of include/pinn_abi.h) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libpinn_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "pinn_capi.cu"), os.path.join(HERE, "csrc", "pinn_tensor_ops.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + \
       [os.path.join(ROOT, "include", "pinn_abi.h"), os.path.join(ROOT, "include", "pinn_tensor_abi.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if p and os.path.exists(p):
            return p
    raise RuntimeError("nvcc not found")


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + SOURCES + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)   # the image exports a wrapper gcc; let nvcc pick /usr/bin/gcc
    subprocess.run(cmd, check=True, env=env)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
