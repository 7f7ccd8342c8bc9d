"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  liblvreg.so          CUDA kernels + C ABI (csrc/lvreg.cu), sm_100a only
  liblvreg_host.so     C++ host mirror of mapOptimization + synthetic generator (host/*.cpp)
  lvreg_replay         standalone replay harness binary (host/replay_main.cpp)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "liblvreg.so")
HOSTLIB = os.path.join(HERE, "liblvreg_host.so")
REPLAY = os.path.join(HERE, "lvreg_replay")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-fmad=false",              # one rounding per op, like the reference's x86-64 build
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_cuda(force=False, verbose=False):
    srcs = _sources(CSRC, (".cu", ".cuh")) + [os.path.join(HERE, "..", "include", "lvreg.h")]
    if force or _stale(LIB, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB, os.path.join(CSRC, "lvreg.cu")]
        subprocess.check_call(cmd)
    return LIB


def build_host(force=False):
    if not os.path.isdir(HOST):
        return None
    cpps = [s for s in _sources(HOST, (".cpp",)) if not s.endswith("replay_main.cpp")]
    if not cpps:
        return None
    deps = _sources(HOST, (".cpp", ".hpp", ".h")) + [LIB]
    if force or _stale(HOSTLIB, deps):
        cmd = [_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", os.path.join(HERE, "..", "include"),
               "-o", HOSTLIB] + cpps + ["-L", HERE, "-llvreg", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    main = os.path.join(HOST, "replay_main.cpp")
    if os.path.exists(main) and (force or _stale(REPLAY, deps + [main, HOSTLIB])):
        cmd = [_cxx(), "-O2", "-std=c++17", "-pthread", "-I", os.path.join(HERE, "..", "include"), "-o", REPLAY, main,
               "-L", HERE, "-llvreg_host", "-llvreg", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return HOSTLIB


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
