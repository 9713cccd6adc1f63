"""Build recipe for librtr_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  The shared library lands IN-TREE next to this file so that it
travels to the GPU box with the gpurun snapshot (it is git-ignored, not gpurun-ignored).
Flags: -fmad=false because every FP operation on the parity path is spelled with an explicit
_rn / fma intrinsic (csrc/rtr_common.cuh); -lineinfo so ncu's source page maps to our code.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# RTR_LIB_SUFFIX / RTR_NVCC_FLAGS / RTR_EXPERIMENTS=1: experiment builds next to the shipped library (tools/experiments)
LIB = os.path.join(HERE, "librtr_b200" + os.environ.get("RTR_LIB_SUFFIX", "") + ".so")
SOURCES = ["rtr_point_kernels.cu", "rtr_point_ring.cu", "rtr_image_kernels.cu", "rtr_cull.cu", "rtr_peer.cu", "rtr_synth.cu", "rtr_microbench.cu", "rtr_io.cu", "rtr_renderer.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=hidden", "--threads", "4",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rtr_b200.h"), os.path.join(HERE, "..", "include", "rtr_b200_io.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into librtr_b200.so; returns the library path."""
    if not force and not _stale():
        return LIB
    extra = os.environ.get("RTR_NVCC_FLAGS", "").split() + (["-DRTR_EXPERIMENTS"] if os.environ.get("RTR_EXPERIMENTS") == "1" else [])
    cmd = [NVCC, *FLAGS, *extra, "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
