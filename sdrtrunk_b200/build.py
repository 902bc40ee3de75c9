"""Builds libsdrgpu.so in-tree with nvcc for sm_100a (no torch involved).  Every source is compiled to its own object
(in parallel, rebuilt only when it or a header changed), then linked."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsdrgpu.so")

COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # Java never contracts a*b+c: keep every float op separately rounded; FMA only where written (fmaf)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [__file__]


def obj_of(src):
    return os.path.join(OBJ, os.path.basename(src) + ".o")


def needs_build():
    if not os.path.exists(LIB):
        return True
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in sources() + headers())


def _compile(args):
    src, verbose = args
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + COMPILE_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj_of(src), src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in headers())
    todo = [s for s in sources()
            if force or verbose or not os.path.exists(obj_of(s))
            or os.path.getmtime(obj_of(s)) < max(os.path.getmtime(s), newest_header)]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 4))) as pool:
        for src, res in pool.map(_compile, [(s, verbose) for s in todo]):
            if verbose:
                sys.stderr.write(res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
    nvcc = os.environ.get("NVCC", "nvcc")
    res = subprocess.run([nvcc] + LINK_FLAGS + ["-o", LIB] + [obj_of(s) for s in sources()], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
