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
    if force:
        shutil.rmtree(os.path.join(CSRC, "gen", "obj", "default"), ignore_errors=True)
    return _compile(LIB_PATH, [], verbose)


def _compile(LIB_PATH, extra, verbose):
    """One nvcc -c per translation unit, in parallel, objects cached under csrc/gen/obj (rebuilt when the source or any
    header is newer), then one link step."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: lidbox_b200 has no CPU fallback and cannot run without its CUDA library")
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)] + extra
    tag = "default" if not extra else "x%08x" % (hash(tuple(extra)) & 0xffffffff)
    obj_dir = os.path.join(CSRC, "gen", "obj", tag)
    os.makedirs(obj_dir, exist_ok=True)
    hdr_time = max(os.path.getmtime(p) for p in _deps() if not p.endswith((".cu", ".cpp")))
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append([nvcc] + compile_flags + ["-I", os.path.join(REPO_DIR, "include"), "-I", CSRC, "-c", "-o", obj, src])
    log = []
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        procs = list(ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs))
    failed = False
    for cmd, proc in zip(jobs, procs):
        log.append(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
        failed = failed or proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
            "-o", LIB_PATH + ".tmp"] + objs
    proc = subprocess.run(link, capture_output=True, text=True)
    log.append(" ".join(link) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("link failed (exit %d)" % proc.returncode)
    with open(os.path.join(PKG_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
