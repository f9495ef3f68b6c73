"""Build libbtslpg.so (hand-written sm_100a CUDA + the C ABI of include/btslpg.h) in-tree with nvcc.

    python bts-fully-tf_b200/build.py [--force] [--verbose] [-DNAME[=V] ...] [--out=libname.so]

One translation unit per kernel family (csrc/*_api.cu), compiled in parallel into lib/obj/*.o and linked into
lib/libbtslpg.so (git-ignored, but it travels to the GPU box with the repo snapshot).  Only the units whose
sources changed are recompiled.  nvcc cross-compiles without a GPU.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libbtslpg.so")
HEADER = os.path.join(HERE, "..", "include", "btslpg.h")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    # IEEE sqrt/div and no flush-to-zero: the kernels pick their approximations explicitly
    "--fmad=true", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


_INC = re.compile(r'^\s*#include\s+"([^"]+)"', re.M)


def deps_of(path, seen=None):
    """Transitive closure of the quoted includes of a source file (paths relative to its directory)."""
    seen = seen if seen is not None else set()
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in _INC.findall(f.read()):
            deps_of(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _stale(target, inputs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in inputs)


def needs_build():
    return any(_stale(LIB, deps_of(os.path.join(CSRC, s))) for s in sources())


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: experiment builds (e.g. -DBTSLPG_HEAD_STAGES=2 into lib/libbtslpg_x.so, loaded with BTSLPG_LIB=...)."""
    target = out or LIB
    experiment = bool(defines) or out is not None
    obj_dir = OBJ_DIR + ("_x" if experiment else "")
    os.makedirs(obj_dir, exist_ok=True)
    jobs = []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s[:-3] + ".o")
        if force or experiment or _stale(obj, deps_of(src)):
            cmd = ([nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
            jobs.append(cmd)
    if not jobs and os.path.exists(target) and not force:
        return target

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, p.returncode, p.stdout

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        results = list(ex.map(run, jobs))
    for cmd, rc, text in results:
        if text and (verbose or rc):
            print(text, flush=True)
        if rc:
            raise subprocess.CalledProcessError(rc, cmd, text)
    objs = [os.path.join(obj_dir, s[:-3] + ".o") for s in sources()]
    link = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", target] + objs
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.check_call(link)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs,
                out=os.path.join(LIB_DIR, outs[0]) if outs else None))
