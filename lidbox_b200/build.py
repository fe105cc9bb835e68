"""Builds lidbox_b200/liblidbox_b200.so (the C-ABI library, include/lidbox_b200.h) with nvcc for sm_100a, in-tree."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "liblidbox_b200.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
    "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return sources() + hdrs + [os.path.join(REPO_DIR, "include", "lidbox_b200.h")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every CUDA translation unit into one shared library. Returns the library path."""
    if out is not None:
        return _compile(out, list(extra_flags), verbose)
    if not force and not is_stale():
        return LIB_PATH
    return _compile(LIB_PATH, [], verbose)


def _compile(LIB_PATH, extra, verbose):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: lidbox_b200 has no CPU fallback and cannot run without its CUDA library")
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", os.path.join(REPO_DIR, "include"), "-I", CSRC, "-o", LIB_PATH + ".tmp"] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d)" % proc.returncode)
    with open(os.path.join(PKG_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
