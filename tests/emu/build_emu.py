"""Build the host-compiled test double of libhsp2g (see README.md in this directory).  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "hspsquared_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libhsp2g_emu.so")
UNITS = ("hsp2g_api.cu", "perlnd_kernel.cu", "implnd_kernel.cu", "qual_kernel.cu")
# strict IEEE like the device build (-fmad=false): no contraction, no fast-math
FLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c++17", "-DHSP2G_HOST_EMU", "-I", HERE, "-I", CSRC]


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(HERE, f) for f in ("emu_cuda_stubs.h", "emu_globals.cpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    procs = []
    for u in UNITS:
        o = os.path.join(HERE, "_build", u.replace(".cu", ".o"))
        objs.append(o)
        procs.append(subprocess.Popen(["g++"] + FLAGS + ["-x", "c++", "-c", os.path.join(CSRC, u), "-o", o],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    o = os.path.join(HERE, "_build", "emu_globals.o")
    objs.append(o)
    procs.append(subprocess.Popen(["g++"] + FLAGS + ["-c", os.path.join(HERE, "emu_globals.cpp"), "-o", o],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate()
        if p.returncode:
            raise RuntimeError("emulation build failed:\n" + out)
    subprocess.run(["g++", "-shared", "-o", OUT] + objs, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
