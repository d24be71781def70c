"""Host emulation of the PRODUCT library, for the CPU-only test run (test infrastructure; see tests/support/hostemu/).

``build()`` compiles landhydrology.jl_b200/csrc/*.cu — the same sources nvcc compiles for sm_100a — with g++ against
tests/support/hostemu/cuda_runtime.h and links them with the fiber scheduler (hostemu.cpp) into
tests/support/hostemu/_build/liblh_soil_hostemu.so (git-ignored, cached on source mtimes).  ``library()`` opens it as a
``SoilLibrary`` with the product's own symbol names, so every test written against the CUDA library can also run against the
emulated build.  Never imported by the package, by bench.py or by ``__graft_entry__``."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "landhydrology.jl_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "support", "hostemu")
BUILD = os.path.join(EMU, "_build")
LIB = os.path.join(BUILD, "liblh_soil_hostemu.so")

# -O1: the 230-odd kernel variants compile in ~1 min on 8 cores and run ~20x faster than at -O0.  -ffp-contract=off: an FMA
# is fused exactly where the source says lh_fma / fma, as nvcc's --fmad=true would only ADD contractions of a*b+c written
# out — the closures avoid those on purpose (every such expression is spelled with fma), so both builds round alike.
CXXFLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-DLH_HOSTEMU=1", "-DLH_MATH_HOST=1",
            "-I", EMU, "-I", CSRC, "-I", os.path.join(ROOT, "include")]


def _sources():
    cu = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc")))
    deps += [os.path.join(EMU, f) for f in ("cuda_runtime.h", "hostemu.cpp", "exports.map")] + [os.path.join(ROOT, "include", "lh_soil.h"), __file__]
    return cu, deps


def build(force: bool = False) -> str:
    cu, deps = _sources()
    newest = max(os.path.getmtime(p) for p in cu + deps)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}

    def compile_one(src):
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), max(os.path.getmtime(d) for d in deps)):
            return obj
        cmd = ["g++", *CXXFLAGS, "-x", "c++", "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("host-emulation build failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj

    with ThreadPoolExecutor(max(1, len(os.sched_getaffinity(0)))) as pool:
        objs = list(pool.map(compile_one, cu + [os.path.join(EMU, "hostemu.cpp")]))
    # -Bsymbolic + the export map: calls inside the library bind inside the library whatever else the test process has loaded
    cmd = ["g++", "-shared", "-Wl,-Bsymbolic", "-Wl,--version-script=" + os.path.join(EMU, "exports.map"), "-o", LIB + ".tmp", *objs, "-ldl", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("host-emulation link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


def build_mutant(name: str, filename: str, old: str, new: str) -> str:
    """A copy of the emulated library in which csrc/<filename> has `old` replaced by `new` — a deliberately broken product, for
    checking that the tests which claim to detect a class of bug really do (tests/test_hostemu.py).  A mutation of
    lh_soil_api.cu recompiles only that translation unit (the kernels' objects are those of the regular build); a mutation of
    a header recompiles everything from a patched copy of csrc/."""
    build()
    src = open(os.path.join(CSRC, filename)).read()
    assert src.count(old) == 1, "mutation site not found exactly once"
    text = src.replace(old, new)
    mdir = os.path.join(BUILD, "mutants", name)
    mlib, stamp = os.path.join(mdir, "libmutant.so"), os.path.join(mdir, "patched_source")
    if os.path.exists(mlib) and os.path.exists(stamp) and open(stamp).read() == text and os.path.getmtime(mlib) >= os.path.getmtime(LIB):
        return mlib
    os.makedirs(mdir, exist_ok=True)
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    link = ["g++", "-shared", "-Wl,-Bsymbolic", "-Wl,--version-script=" + os.path.join(EMU, "exports.map"), "-o", mlib]
    if filename == "lh_soil_api.cu":
        msrc, mobj = os.path.join(mdir, "lh_soil_api.cu"), os.path.join(mdir, "lh_soil_api.cu.o")
        open(msrc, "w").write(text)
        subprocess.run(["g++", *CXXFLAGS, "-x", "c++", "-c", msrc, "-o", mobj], check=True, env=env)
        objs = [os.path.join(BUILD, f) for f in os.listdir(BUILD) if f.endswith(".o") and f != "lh_soil_api.cu.o"] + [mobj]
    else:
        import shutil

        mcsrc = os.path.join(mdir, "csrc")
        shutil.rmtree(mcsrc, ignore_errors=True)
        shutil.copytree(CSRC, mcsrc, ignore=shutil.ignore_patterns("*.so", "*.o"))
        open(os.path.join(mcsrc, filename), "w").write(text)
        flags = [f if f != CSRC else mcsrc for f in CXXFLAGS]
        units = sorted(os.path.join(mcsrc, f) for f in os.listdir(mcsrc) if f.endswith(".cu")) + [os.path.join(EMU, "hostemu.cpp")]

        def compile_one(src_path):
            obj = os.path.join(mdir, os.path.basename(src_path) + ".o")
            subprocess.run(["g++", *flags, "-x", "c++", "-c", src_path, "-o", obj], check=True, env=env)
            return obj

        with ThreadPoolExecutor(max(1, len(os.sched_getaffinity(0)))) as pool:
            objs = list(pool.map(compile_one, units))
    subprocess.run([*link, *objs, "-ldl", "-lpthread"], check=True, env=env)
    open(stamp, "w").write(text)
    return mlib


def build_fake_nccl() -> str:
    """tests/support/hostemu/fake_nccl.cpp as .../_build/fake_nccl/libnccl.so.2 (soname libnccl.so.2): the multi-rank path of the
    emulated build, with host threads as ranks.  Load it (RTLD_GLOBAL) before the product library first asks for NCCL."""
    src = os.path.join(EMU, "fake_nccl.cpp")
    out = os.path.join(BUILD, "fake_nccl", "libnccl.so.2")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wl,-soname,libnccl.so.2", src, "-o", out, "-lpthread"], check=True, env=env)
    return out


_lib = None


def library(lh):
    """The emulated build as a ``SoilLibrary`` (product symbol names, prefix ``lh_``)."""
    global _lib
    if _lib is None:
        _lib = lh.SoilLibrary(build(), "lh_")
    return _lib


def controls(path=None):
    """ctypes handle with the emulator's knobs: lh_emu_set_schedule, lh_emu_set_cp_async_lazy, lh_emu_set_sm_count, ..."""
    h = C.CDLL(path or build())
    h.lh_emu_set_schedule.argtypes = [C.c_int, C.c_uint64]
    h.lh_emu_set_schedule.restype = C.c_int
    h.lh_emu_set_cp_async_lazy.argtypes = [C.c_int]
    h.lh_emu_set_cp_async_lazy.restype = C.c_int
    h.lh_emu_set_device_count.argtypes = [C.c_int]
    h.lh_emu_set_device_count.restype = C.c_int
    h.lh_emu_set_sm_count.argtypes = [C.c_int]
    h.lh_emu_set_sm_count.restype = C.c_int
    h.lh_emu_set_async.argtypes = [C.c_int, C.c_uint64]
    h.lh_emu_set_async.restype = C.c_int
    h.lh_emu_live_handles.argtypes = []
    h.lh_emu_live_handles.restype = C.c_int64
    h.lh_emu_fail_allocation_in.argtypes = [C.c_int64]
    h.lh_emu_fail_allocation_in.restype = None
    h.lh_emu_live_allocations.argtypes = []
    h.lh_emu_live_allocations.restype = C.c_uint64
    h.lh_emu_launch_count.argtypes = []
    h.lh_emu_launch_count.restype = C.c_uint64
    return h
