"""Builds libbdeflate.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "bdeflate.cu")
OUT = os.path.join(HERE, "libbdeflate.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(os.path.dirname(HERE), "include", "bdeflate.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


OUT_CHECK = os.path.join(HERE, "libbdeflate_check.so")


def build(force=False, verbose=False, check=False, defines=(), out=None):
    """check=True: the debug build with device-side assertions (-DBDF_CHECK), libbdeflate_check.so.
    defines / out: an experiment build (-DNAME=VALUE ...) under another file name, loaded with BDF_LIBRARY."""
    out = os.path.join(HERE, out) if out else (OUT_CHECK if check else OUT)
    stale = not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS)
    if not force and not stale:
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-DBDF_CHECK"] if check else []) + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, check="--check" in sys.argv, defines=defs,
                out=outs[0] if outs else None))
