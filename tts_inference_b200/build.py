"""Build ``libsnacb.so`` (sm_100a only) in-tree with nvcc.

    python -m tts_inference_b200.build [--force] [--verbose]

The library has no torch / Python dependency: it is a plain C-ABI shared object
(include/snacb.h) that links the CUDA runtime statically and fetches
``cuTensorMapEncodeTiled`` from the driver at run time.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsnacb.so")
SOURCES = ["snacb.cu", "kernels_simt.cu", "kernels_tc.cu", "kernels_res2.cu", "kernels_chain.cu", "kernels_chain_ws.cu", "kernels_convt.cu", "kernels_io.cu", "encoder.cu", "batcher.cpp", "streamer.cpp"]
HEADERS = ["common.cuh", "chain_span.cuh", "kernels.h", "ptx.cuh", os.path.join("..", "..", "include", "snacb.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """SNACB_EXPERIMENTS=1 in the environment also compiles the measured-and-dropped kernel variants (k_chain_ws); the
    default product build leaves them out."""
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(PKG, "build", os.path.splitext(s)[0] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if os.environ.get("SNACB_EXPERIMENTS") == "1":
            cmd.insert(1, "-DSNACB_EXPERIMENTS")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libsnacb.so")
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB, *objs]
    subprocess.run(cmd, check=True)
    return LIB


BENCH_SRC = os.path.join(os.path.dirname(PKG), "tests", "native", "batcher_bench.cpp")
BENCH_LIB = os.path.join(os.path.dirname(PKG), "tests", "native", "libbatcherbench.so")


def build_test_helpers(force: bool = False) -> str:
    """tests/native/libbatcherbench.so: the multi-threaded producer harness of tests/gpu_batcher_bench.py (g++, links
    libsnacb.so by path; test infrastructure, not shipped)."""
    if not os.path.exists(BENCH_SRC):
        return ""
    if not force and os.path.exists(BENCH_LIB) and os.path.getmtime(BENCH_LIB) >= max(os.path.getmtime(BENCH_SRC), os.path.getmtime(LIB)):
        return BENCH_LIB
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", BENCH_SRC, "-o", BENCH_LIB,
           "-L" + PKG, "-l:libsnacb.so", "-Wl,-rpath," + PKG, "-Wl,-rpath,$ORIGIN/../../tts_inference_b200"]
    subprocess.run(cmd, check=True)
    return BENCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_test_helpers(force="--force" in sys.argv))
