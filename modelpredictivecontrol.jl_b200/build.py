"""Builds libbmpc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.
The translation units (API + general kernel, and one unit per warp-kernel specialisation) are
compiled in parallel and linked into one shared object."""
import concurrent.futures
import glob
import zlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(os.environ.get("TMPDIR", "/tmp"), "bmpc_build_objs")  # outside the tree: only libbmpc.so travels
UNITS = [os.path.join(CSRC, "bmpc_api.cu"), os.path.join(CSRC, "bmpc_mhe_api.cu")] + sorted(
    glob.glob(os.path.join(CSRC, "warp_inst_*.cu")))
HEADERS = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))) + [
    os.path.join(os.path.dirname(HERE), "include", "bmpc.h")]
OUT = os.path.join(HERE, "libbmpc.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--expt-extended-lambda", "-Xcompiler", "-fPIC"]


EXTRA = os.environ.get("BMPC_NVCC_EXTRA", "").split()  # study builds: e.g. -DBMPC_PHASE_CLK (with BMPC_OUT / TMPDIR set)
if EXTRA:
    FLAGS = FLAGS + EXTRA
    OBJDIR = OBJDIR + "_%08x" % zlib.crc32(" ".join(EXTRA).encode())
OUT = os.environ.get("BMPC_OUT", OUT)


def _obj(src):
    return os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src, verbose):
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(src), src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [u for u in UNITS if force or _stale(_obj(u), [u] + HEADERS)]
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            for src, r in ex.map(lambda u: _compile(u, verbose), todo):
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"nvcc failed on {src}")
                if verbose:
                    sys.stderr.write(r.stderr)
    objs = [_obj(u) for u in UNITS]
    if todo or _stale(OUT, objs):
        r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc link failed for libbmpc.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
