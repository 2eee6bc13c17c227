"""Build libbayhunter_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree.

nvcc cross-compiles without a GPU; the resulting shared object travels with
the repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbayhunter_b200.so")
SOURCES = ["engine.cu", "prep_kernel.cu", "swd_kernel.cu", "swd_lockstep.cu", "swd_pool.cu", "swd_general.cu", "rf_kernel.cu", "loglik_kernel.cu", "sampler.cu", "noise_kernel.cu"]
HEADERS = ["bh_common.cuh", "bh_math.cuh", "swd_core.cuh", "swd_eval.cuh", "swd_general_core.cuh", "sampler_core.cuh", "rf_core.cuh", "kernels.h",
           os.path.join("..", "..", "include", "bayhunter_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
    "-cudart", "static",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libbayhunter_b200.so")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra=()):
    """Compile every CUDA source for sm_100a into one shared library."""
    if not force and not needs_build():
        return LIB
    # several ranks of one torchrun may get here at once: one builds (into a temporary file, renamed into
    # place when complete), the others wait for the lock and find the library up to date
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB
        tmp = "%s.%d.tmp" % (LIB, os.getpid())
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            print(" ".join(cmd).replace(tmp, LIB))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed building libbayhunter_b200.so")
        os.replace(tmp, LIB)
        if verbose and (res.stdout or res.stderr):
            print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra=("-Xptxas", "-v") if "--ptxas" in sys.argv else ())
    print(LIB)
