"""Kernels built for ONE batch's exact configuration, at run time.

The land kernels are templates (csrc/land_kernels.cuh, csrc/qual.cuh) over everything that is constant for a batch: the
activity mask, whether the step is hourly, which forcing rows exist and which outputs are saved (StaticIO), the element
types and the series layout (StaticLay) and -- when every segment carries the same 64-bit option word, i.e. a replicated or
calibration ensemble -- that word (StaticFlags), so that every option test folds and the sections the word switches off
(degree-day snow, ARM routing, UZINF2, monthly-table paths ...) vanish from the code.  libhsp2g.so itself ships only the
general instantiations (run-time row tables, option words and layout).  This module writes the one-line instantiation for a
batch, compiles it with `nvcc -cubin` for sm_100a (1-3 s; cached in hspsquared_b200/_kcache, keyed by the generated text and
the kernel sources) and hands the image to the library (hsp2g_set_kernel_image), which checks the image's own description
against the batch before every launch.  Same arithmetic, same bits as the general kernels (tests/test_jit.py, -m gpu
test_option_word_kernels_equal_the_general_ones).

If nvcc is missing the batch simply runs the library's general kernels (LandBatch.kernel_name() tells which ran).
HSP2G_JIT=0 switches this module off.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import tempfile
import threading
import time

from . import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CACHE = os.environ.get("HSP2G_KCACHE") or os.path.join(HERE, "_kcache")
ENTRY = "hsp2g_jit_kernel"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-Xptxas", "-v"]
ACT_NAMES = ("ATEMP", "SNOW", "PWATER", "SEDMNT", "PSTEMP", "PWTGAS", "IWATER", "SOLIDS", "IWTGAS", "PQUAL", "IQUAL")
CHAIN_ACTS = 1 | 8 | 16 | 32

_src_hash = None
_failed = {}


# a batch below this size is not worth a 1-3 s build (the drop-in activities run one segment per call): LandBatch(jit=None)
MIN_SEGMENTS = int(os.environ.get("HSP2G_JIT_MIN_SEGMENTS", "2048"))
SPILL_OK = 32  # bytes of register spill tolerated before the register cap is raised


def enabled(n_segments=None):
    d = "1" if _lib.build_kind() == _lib.CUDA_BUILD else "0"  # the CPU test double uses it only where a test asks
    if os.environ.get("HSP2G_JIT", d) == "0":
        return False
    return n_segments is None or n_segments >= MIN_SEGMENTS


def _nvcc():
    for c in (os.environ.get("HSP2G_NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    import shutil

    return shutil.which("nvcc")


def _sources_hash():
    global _src_hash
    if _src_hash is None:
        h = hashlib.sha1()
        for f in sorted(os.listdir(CSRC)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(CSRC, f), "rb") as fh:
                    h.update(f.encode() + b"\0" + fh.read())
        _src_hash = h.hexdigest()
    return _src_hash


def spec_for(store, delt, forcing_names, save, layout="plain", forcing_dtype="float64", out_dtype="float64", shared=False,
             word=True, opts=None, met_table=False, linked=False):
    """The compile-time description of a batch that does not exist yet (no GPU needed): store = SegmentStore, forcing_names /
    save = the names LandBatch would be given (save may be 'all')."""
    import numpy as np

    F, O = _lib.index("forcings"), _lib.index("outputs")
    fmask = 0
    for nme in forcing_names:
        fmask |= 1 << F[nme]
    act = int(store.act)
    if isinstance(save, str) and save == "all":
        from .engine import output_names

        save = output_names(store.operation, act)
    s = [0, 0, 0]
    for nme in save:
        s[O[nme] // 64] |= 1 << (O[nme] % 64)
    qual = bool(act & (_lib.ACT_BITS["PQUAL"] | _lib.ACT_BITS["IQUAL"]))
    # the option bits every segment agrees on are fixed at compile time (all 64 for a replicated ensemble)
    f_and = int(np.bitwise_and.reduce(store.flags)) if len(store.flags) else 0
    f_or = int(np.bitwise_or.reduce(store.flags)) if len(store.flags) else 0
    mask = ~(f_and ^ f_or) & 0xFFFFFFFFFFFFFFFF
    f32in = np.dtype(forcing_dtype).itemsize == 4 and not shared
    if opts is None:  # the batch-wide option bits LandBatch would set (HSP2G_OPT_*)
        from .engine import forcing_fill

        opts = forcing_fill(store.operation, act, tuple(forcing_names))[1] | (_lib.OPT_METRIC if getattr(store, "units", 1) == 2 else 0)
    return {
        "op": 0 if store.operation == "PERLND" else 1, "qual": qual, "act": act, "hourly": bool(delt >= 60),
        "fmask": fmask, "s0": s[0], "s1": s[1], "s2": s[2], "f32out": bool(np.dtype(out_dtype).itemsize == 4),
        "plain": bool(layout == "plain"), "f32in": bool(f32in), "shared": bool(shared),
        "word": (f_and & mask) if (word and not qual and mask) else None, "mask": mask if (word and not qual and mask) else 0,
        "opts": None if qual else int(opts), "mettbl": bool(shared and met_table), "linked": bool(linked and not shared),
    }


def spec_of(batch, word=True):
    """Everything of `batch` (engine.LandBatch / QualBatch) that the kernel can fix at compile time."""
    return spec_for(batch.store, batch.cal.delt, batch.forcing_rows, batch.save_rows, batch.layout, batch.forcing_dtype,
                    batch.out_dtype, batch.shared, word, getattr(batch, "opts", None), getattr(batch, "met_table", False),
                    getattr(batch, "linked", False))


def _popc(x):
    return bin(x).count("1")


def default_maxreg(spec):
    """Register cap = resident warps per SM (65536 / (32 * cap)): 128 -> 16 warps for the water-budget kernels, 168 -> 12 for
    the full chain and for SAVE sets so large that 128 would spill (ensure() moves up on spills by itself)."""
    env = os.environ.get("HSP2G_JIT_MAXREG")
    if env:
        return int(env)
    if spec["qual"]:
        return 128
    if spec["op"] == 1:
        # IMPLND SNOW + IWATER needs 106 registers at the 128 cap and fits 96 without spilling: 10 instead of 9 blocks per SM,
        # 0.758 -> 0.772 of measured HBM over a year (80: spills, 0.592; profiles/r3i_implnd_regcap.jsonl)
        return 96
    ns = _popc(spec["s0"]) + _popc(spec["s1"]) + _popc(spec["s2"])
    return 168 if (spec["act"] & CHAIN_ACTS) or ns > 32 else 128


def geometry(spec, kind=None):
    """(threads per block, lock-step level) of the kernel built for `spec`.  Lock-step blocks (csrc/land_common.cuh
    HSP_LOCKSTEP): the warps of a block take adjacent tiles and meet at a block barrier before every interval."""
    if spec["qual"] or (kind or _lib.build_kind()) != _lib.CUDA_BUILD:
        return 64, 0
    block = int(os.environ.get("HSP2G_JIT_BLOCK", spec.get("block") or 64))
    lockstep = int(os.environ.get("HSP2G_JIT_LOCKSTEP", spec.get("lockstep") or 0))
    return block, lockstep


def write_back_stores(spec):
    """Ordinary write-back stores instead of streaming ones for the saved series (csrc/land_common.cuh HSP_STORE_POLICY)?"""
    ns = _popc(spec["s0"]) + _popc(spec["s1"]) + _popc(spec["s2"])
    return bool(spec["plain"]) and ns > 48 and "HSP_STORE_POLICY" not in os.environ.get("HSP2G_JIT_FLAGS", "")


def display_name(spec, maxreg, kind=None):
    acts = "|".join(n for i, n in enumerate(ACT_NAMES) if spec["act"] >> i & 1)
    kern = "qual_kernel" if spec["qual"] else ("perlnd_kernel" if spec["op"] == 0 else "implnd_kernel")
    parts = [acts, "hourly" if spec["hourly"] else "subhourly",
             "forc=%d:%#x%s" % (_popc(spec["fmask"]), spec["fmask"], ",f32" if spec["f32in"] else ""),
             "save=%d:%#x.%#x.%#x%s" % (_popc(spec["s0"]) + _popc(spec["s1"]) + _popc(spec["s2"]), spec["s0"], spec["s1"],
                                         spec["s2"], ",f32" if spec["f32out"] else ""),
             "plain" if spec["plain"] else "tiled"]
    if spec.get("linked"):
        parts.append("in=linked-rows")
    if spec["shared"]:
        parts.append("met=shared/table" if spec.get("mettbl") else "met=shared")
    if spec["word"] is None:
        parts.append("options=per-segment")
    elif spec.get("mask", 0xFFFFFFFFFFFFFFFF) == 0xFFFFFFFFFFFFFFFF:
        parts.append("options=%#x" % spec["word"])
    else:
        parts.append("options=%#x/fixed:%#x(%d bits per segment)" % (spec["word"], spec["mask"],
                                                                      _popc(~spec["mask"] & ((1 << len(_lib.names("flags"))) - 1))))
    if spec.get("opts"):
        parts.append("batch-opts=%#x" % spec["opts"])
    if write_back_stores(spec):
        parts.append("stores=write-back")
    block, lockstep = geometry(spec, kind)
    if block != 64 or lockstep:
        parts.append("block=%d%s" % (block, ",lockstep=%d" % lockstep if lockstep else ""))
    parts.append("regs<=%d" % maxreg)
    return "%s<%s,jit>" % (kern, ",".join(parts))


def source(spec, maxreg, kind=None):
    """The translation unit: one instantiation of the body template + the info words hsp2g_set_kernel_image checks."""
    u = lambda v: "%#xull" % v
    b = lambda v: "true" if v else "false"
    act = spec["act"]
    lines = ["// generated by hspsquared_b200/jit.py for one batch configuration -- not a source file of the repository",
             "// %s" % display_name(spec, maxreg, kind)]
    inline_div = spec["word"] is not None and spec.get("mask", 0xFFFFFFFFFFFFFFFF) == 0xFFFFFFFFFFFFFFFF and not (act & CHAIN_ACTS)
    if os.environ.get("HSP2G_JIT_INLINE_DIV"):  # measurement switch
        inline_div = os.environ["HSP2G_JIT_INLINE_DIV"] == "1"
    if inline_div:
        # small kernels (a known option word removes whole sections): the IEEE division sequence inline at every site
        lines.append("#define HSP_INLINE_DIV")
    if write_back_stores(spec):
        # many saved rows in the plain layout: a warp-interval scatters ns pieces of 256 bytes, each in another DRAM page; ordinary
        # write-back stores let the L2 gather neighbouring tiles' pieces before they leave (all 84 outputs of the full chain: +5.5 %,
        # profiles/r3d_fullchain_layout_policy.jsonl; no effect on the 24-row headline, profiles/r2z_store_policy.jsonl)
        lines.append("#define HSP_STORE_POLICY 1")
    block, lockstep = geometry(spec, kind)
    if block != 64:
        lines.append("#define LAND_BLOCK %d" % block)
    if lockstep:
        lines.append("#define HSP_LOCKSTEP %d" % lockstep)
    lines += ['#include "jit_info.h"', '#include "land_kernels.cuh"', '#include "qual.cuh"',
              "typedef StaticIO<%s, %s, %s, %s, %s> IO;" % (u(spec["fmask"]), u(spec["s0"]), u(spec["s1"]), b(spec["f32out"]), u(spec["s2"])),
              "typedef %s FW;" % ("StaticFlags<%s, %s, %d>" % (u(spec["word"] or 0), u(spec.get("mask", 0) if spec["word"] is not None else 0),
                                                                 spec["opts"])
                                  if not spec["qual"] else "DynFlags"),
              "typedef StaticLay<%s, %s, %d, %s> LAY;" % (b(spec["plain"]), b(spec["f32in"]),
                                                          int(bool(spec.get("mettbl"))) if spec["shared"] else -1,
                                                          b(spec.get("linked"))),
              "constexpr int ACT = %d;" % act, "constexpr bool HOURLY = %s, SHARED = %s;" % (b(spec["hourly"]), b(spec["shared"])),
              "constexpr int FESZ = %d;" % (4 if spec["f32in"] else 8)]
    if spec["qual"]:
        body = "qual_body<%s, IO, LAY>(L);" % b(spec["op"] == 1)
        smem = "land_smem_bytes(IO::NF, 0ull, 0, false, FESZ)"
    elif spec["op"] == 0:
        body = "perlnd_body<ACT, HOURLY, IO, SHARED, FW, LAY>(L);"
        smem = "land_smem_bytes(IO::NF, perlnd_hot_io<IO>(ACT), perlnd_ncold<HOURLY, IO, ACT>(), SHARED, FESZ)"
    else:
        body = "implnd_body<ACT, HOURLY, IO, SHARED, FW, LAY>(L);"
        smem = "land_smem_bytes(IO::NF, IMPLND_HOT, implnd_ncold<HOURLY, IO>(), SHARED, FESZ)"
    bits = " | ".join(n for n, on in (("JB_F32OUT", spec["f32out"]), ("JB_PLAIN", spec["plain"]), ("JB_F32IN", spec["f32in"]),
                                      ("JB_SHARED", spec["shared"]), ("JB_WORD", spec["word"] is not None), ("JB_IOSTATIC", True),
                                      ("JB_LAYSTATIC", True), ("JB_QUAL", spec["qual"]),
                                      ("JB_OPTS", spec.get("opts") is not None), ("JB_METTBL", spec.get("mettbl")),
                                      ("JB_LINKED", spec.get("linked")), ("JB_LOCKSTEP", lockstep)) if on)
    info = ("{HSP2G_JIT_MAGIC, sizeof(LandLaunch), %s, %d, %s, %d, %s, %s, %s, %s, %s, %s, LAND_BLOCK, SB_ROWS, %s, %d}"
            % (smem, spec["op"], u(act), int(spec["hourly"]), u(spec["fmask"]), u(spec["s0"]), u(spec["s1"]), u(spec["s2"]), bits,
               u(spec["word"] or 0), u(spec.get("mask", 0xFFFFFFFFFFFFFFFF) if spec["word"] is not None else 0), spec.get("opts") or 0))
    lines += ["#ifdef HSP2G_HOST_EMU  /* the CPU test suite's double: a host function with the same body */",
              'extern "C" void %s(const LandLaunch &L) { %s }' % (ENTRY, body),
              'extern "C" const unsigned long long %s_info[JI_N] = %s;' % (ENTRY, info),
              "#else",
              "// register cap %d through the resident-blocks bound (__maxnreg__ cannot be combined with __launch_bounds__ on a"
              " non-template kernel)" % maxreg,
              ('extern "C" __global__ void __launch_bounds__(LAND_BLOCK, %d) %s(const __grid_constant__ LandLaunch L) { %s }'
               % (max(1, 65536 // (block * maxreg)), ENTRY, body)) if not os.environ.get("HSP2G_JIT_MAXNREG") else
              ('extern "C" __global__ void %s(const __grid_constant__ LandLaunch L) { %s } // cap %d'
               % (ENTRY, body, maxreg)),  # measurement switch: an exact cap through -maxrregcount (_compile) -- the resident-blocks
                                          # bound only reaches 128, 96, 80 ...
              'extern "C" __device__ const unsigned long long %s_info[JI_N] = %s;' % (ENTRY, info),
              "#endif", ""]
    return "\n".join(lines)


def _compile(text, out, kind, maxreg=128):
    d = tempfile.mkdtemp(prefix="hsp2g_jit_")
    src = os.path.join(d, "k.cu")
    with open(src, "w") as f:
        f.write(text)
    tmp = out + ".tmp%d_%d" % (os.getpid(), threading.get_ident())  # unique per process AND thread
    if kind == _lib.CUDA_BUILD:
        nvcc = _nvcc()
        if not nvcc:
            raise RuntimeError("nvcc not found")
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("HSP2G_JIT_FLAGS", "").split() + ["-I", CSRC, "-cubin", src, "-o", tmp]
        if os.environ.get("HSP2G_JIT_MAXNREG"):
            cmd += ["-maxrregcount", str(maxreg)]
    else:  # host double (tests/emu): same text, g++, linked against the double for its thread-index globals
        from tests.emu import build_emu

        cmd = ["g++"] + build_emu.FLAGS + ["-shared", "-x", "c++", src, "-o", tmp, "-x", "none", build_emu.OUT,
                                           "-Wl,-rpath," + os.path.dirname(build_emu.OUT)]
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        os.remove(src)
        os.rmdir(d)
    except OSError:
        pass
    if r.returncode != 0:
        raise RuntimeError("kernel build failed (%s):\n%s\n%s" % (" ".join(cmd), r.stdout[-3000:], r.stderr[-3000:]))
    os.replace(tmp, out)
    return r.stderr, time.time() - t


def _ptxas(log):
    """registers and spill bytes of the entry point from `ptxas -v`"""
    regs = re.search(r"Used (\d+) registers", log)
    spill = [int(x) for x in re.findall(r"(\d+) bytes spill stores", log)]
    return (int(regs.group(1)) if regs else None), (max(spill) if spill else 0)


def ensure(spec, kind=None, maxreg=None):
    """-> (image path, entry, display name, meta): built now or taken from the cache."""
    kind = kind or _lib.build_kind()
    os.makedirs(CACHE, exist_ok=True)
    tried = []
    mr = maxreg or default_maxreg(spec)
    while True:
        text = source(spec, mr, kind)
        key = hashlib.sha1((text + "\0" + _sources_hash() + "\0" + kind + "\0" + " ".join(NVCC_FLAGS) +
                            os.environ.get("HSP2G_JIT_FLAGS", "")).encode()).hexdigest()[:20]
        out = os.path.join(CACHE, key + (".cubin" if kind == _lib.CUDA_BUILD else ".so"))
        meta_p = os.path.join(CACHE, key + ".json")
        if os.path.exists(out) and os.path.exists(meta_p):
            with open(meta_p) as f:
                meta = json.load(f)
        else:
            log, secs = _compile(text, out, kind, mr)
            regs, spill = _ptxas(log) if kind == _lib.CUDA_BUILD else (None, 0)
            meta = {"spec": spec, "name": display_name(spec, mr, kind), "maxreg": mr, "registers": regs, "spill_bytes": spill,
                    "compile_s": round(secs, 2), "kind": kind, "sources": _sources_hash()}
            mtmp = meta_p + ".tmp%d_%d" % (os.getpid(), threading.get_ident())
            with open(mtmp, "w") as f:
                json.dump(meta, f)
            os.replace(mtmp, meta_p)
        tried.append((mr, meta.get("spill_bytes", 0)))
        # a kernel that spills in earnest loses more than the occupancy it keeps (profiles/README.md: 280 B of spill cost 14 points):
        # give it registers; a handful of bytes is cheaper than a quarter of the resident warps
        if maxreg is None and not os.environ.get("HSP2G_JIT_MAXREG") and meta.get("spill_bytes", 0) > SPILL_OK and mr < 232:
            mr = 128 if mr < 128 else 168 if mr < 168 else 232
            continue
        meta["tried"] = tried
        return out, ENTRY, meta["name"], meta


def specialize(batch, word=True):
    """Build (or fetch) the kernel for `batch` and hand it to the library.  False (and batch.jit_error) if that was not possible;
    the batch then runs the library's general kernels."""
    kind = _lib.build_kind()
    spec = spec_of(batch, word=word)
    key = json.dumps(spec, sort_keys=True) + kind
    if key in _failed:
        batch.jit_error = _failed[key]
        return False
    try:
        path, entry, name, meta = ensure(spec, kind)
    except Exception as ex:  # no nvcc, compile error: recorded, not fatal (the general kernels compute the same bits)
        _failed[key] = batch.jit_error = repr(ex)
        return False
    if kind == _lib.CUDA_BUILD:
        with open(path, "rb") as f:
            image = f.read()
    else:
        image = path.encode()
        if os.environ.get("HSP2G_JIT_PREBUILD"):  # CPU runs of the test suite can fill the cubin cache for the GPU box
            try:
                ensure(spec, _lib.CUDA_BUILD)
            except Exception:
                pass
    _lib.check(batch.lib.hsp2g_set_kernel_image(batch.h, image, len(image), entry.encode(), name.encode()))
    batch.jit_meta = meta
    return True


def drop(batch):
    _lib.check(batch.lib.hsp2g_set_kernel_image(batch.h, None, 0, None, None))


def prebuild(specs, kind=None):
    """Build the kernels of `specs` now (hspsquared_b200.build / __graft_entry__.build: the cubins travel with the tree, so a
    fresh GPU box does not spend its first seconds in nvcc).  Returns the display names."""
    from concurrent.futures import ThreadPoolExecutor

    kind = kind or _lib.CUDA_BUILD
    with ThreadPoolExecutor(min(8, max(1, len(specs)))) as ex:
        names = [r[2] for r in ex.map(lambda sp: ensure(sp, kind), specs)]
    prune()
    return names


def prune():
    """Drop cached images built from other versions of the kernel headers (their keys can never be asked for again)."""
    import glob
    import json

    n = 0
    for meta_p in glob.glob(os.path.join(CACHE, "*.json")):
        try:
            with open(meta_p) as f:
                stale = json.load(f).get("sources") != _sources_hash()
        except (OSError, ValueError):
            stale = True
        if stale:
            for ext in (".json", ".cubin", ".so"):
                try:
                    os.remove(meta_p[:-5] + ext)
                    n += ext != ".json"
                except OSError:
                    pass
    return n
