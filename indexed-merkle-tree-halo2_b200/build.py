"""Builds libimt_b200.so (the C-ABI library) in-tree with nvcc for sm_100a. No JIT cache: the .so travels with the repo."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libimt_b200.so")
SOURCES = ["imt_capi.cu", "imt_indexed.cu", "imt_spec.cu", "imt_latency.cu", "imt_comm.cu", "imt_io.cu", "poseidon_params.cpp"]
# per-unit flags: which multiply-accumulate chains start without reading the carry flag (csrc/fr.cuh IMT_FREE_MASK). All 32 masks
# were swept on a B200 for both kernel families (tools/lab/latency_lab.cu, profiles/r02_latency_lab.md):
#   22 = odd product chains + even reduction chains + squarings free: the thread-per-hash kernels (-2.4 % time at full occupancy,
#        -12 % on a 32768-node level) — more freedom (31) makes ptxas spill carry predicates, +8 %
#   29 = the 3-lanes-per-hash latency kernels (one warp per scheduler): 283 -> 233 us per small level
UNIT_FLAGS = {"imt_capi.cu": ["-DIMT_FREE_MASK=" + os.environ.get("IMT_THROUGHPUT_FREE_MASK", "22")],
              "imt_latency.cu": ["-DIMT_FREE_MASK=" + os.environ.get("IMT_LATENCY_FREE_MASK", "29")]}
DEPS = SOURCES + ["kernels.cuh", "kernels_common.cuh", "poseidon_coop.cuh", "poseidon_lh.cuh", "poseidon_lh_math.cuh", "imt_internal.h", "poseidon.cuh", "poseidon_spec.cuh", "fr.cuh",
                  "poseidon_params.h"]
OBJ_DIR = os.path.join(HERE, "build")


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(ROOT, "include", "imt_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """one nvcc -c per translation unit, in parallel, then a link step: sm_100a only, -lineinfo for ncu's source page"""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    # the image's CC/CXX wrappers lack some specs; let nvcc find the system g++
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
             "-diag-suppress", "550", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        flags.append("-Xptxas=-v")
    flags += os.environ.get("IMT_NVCC_FLAGS", "").split()  # experiments only, e.g. -DIMT_ILP

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        subprocess.run([_nvcc(), *flags, *UNIT_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj], check=True, env=env)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", *objs, "-ldl", "-o", LIB], check=True, env=env)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
