"""Build libhsp2g.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.  `python -m hspsquared_b200.build`.

The shared object is git-ignored but travels to the GPU box with the working tree.  nvcc cross-compiles without a
GPU.  -fmad=false is REQUIRED for parity: the reference's Numba code has no FMA contraction.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libhsp2g.so")
# tuning experiments only (profiles/): extra -D flags and an alternative output name, e.g.
#   HSP2G_NVCC_EXTRA="-DHSP_MINB_HOT=3" HSP2G_LIB_OUT=/root/repo/gpurun_out/../exp/libhsp2g_m3.so python -m hspsquared_b200.build --force
EXTRA = os.environ.get("HSP2G_NVCC_EXTRA", "").split()
LIB_OUT = os.environ.get("HSP2G_LIB_OUT") or LIB
UNITS = ("hsp2g_api.cu", "perlnd_kernel.cu", "implnd_kernel.cu", "qual_kernel.cu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("HSP2G_NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise RuntimeError("nvcc not found")


def _sources():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "hsp2g.h")]


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources())


def build(force=False, verbose=False):
    if not force and LIB_OUT == LIB and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(unit):
        obj = os.path.join(OBJ, unit.replace(".cu", ".o" if LIB_OUT == LIB else "." + os.path.basename(LIB_OUT) + ".o"))
        cmd = [nvcc] + NVCC_FLAGS + EXTRA + ["-c", os.path.join(CSRC, unit), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(len(UNITS)) as ex:
        results = list(ex.map(compile_one, UNITS))
    log = "\n".join(err for _o, err in results)
    with open(os.path.join(OBJ, "ptxas.log" if LIB_OUT == LIB else os.path.basename(LIB_OUT) + ".ptxas.log"), "w") as f:
        f.write(log)
    if verbose:
        print(log)
    objs = [o for o, _e in results]
    r = subprocess.run([nvcc, "-shared", "-o", LIB_OUT] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
