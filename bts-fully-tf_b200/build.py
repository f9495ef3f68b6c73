"""Build libbtslpg.so (hand-written sm_100a CUDA + the C ABI of include/btslpg.h) in-tree with nvcc.

    python bts-fully-tf_b200/build.py [--force] [--verbose]

The shared library lands in bts-fully-tf_b200/lib/ (git-ignored, but it travels to the GPU box
with the repo snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libbtslpg.so")

SOURCES = ["btslpg_api.cu"]
DEPS = ["btslpg_api.cu", "head_api.inl", "head_kernels.cuh", "lpg_kernels.cuh", "common.cuh", "tma_pipe.cuh", "lpg_dir_tables.h", "tail_kernels.cuh", "tail_api.inl", "concat_kernels.cuh", "concat_api.inl", "upsample_kernels.cuh", "upsample_api.inl", "slice_kernels.cuh", "slice_api.inl", "depthconv_kernels.cuh", "depthconv_api.inl",
        os.path.join("..", "..", "include", "btslpg.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-shared",
    # IEEE sqrt/div and no flush-to-zero: the kernels pick their approximations explicitly
    "--fmad=true", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: experiment builds (e.g. -DBTSLPG_HEAD_STAGES=2 into lib/libbtslpg_x.so, loaded with BTSLPG_LIB=...)."""
    target = out or LIB
    if not force and out is None and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ([nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] +
           [os.path.join(CSRC, s) for s in SOURCES])
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="--verbose" in sys.argv, defines=defs,
                out=os.path.join(LIB_DIR, outs[0]) if outs else None))
