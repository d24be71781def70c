"""The C-ABI shared library loads and exports every symbol include/lh_soil.h declares (no compute
calls: this box has no GPU), the config struct layouts agree, and the product fails loudly —
never falls back — without a device."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import workloads as w

lh = w.lh
abi = w.abi
ROOT = w.ROOT
HEADER = os.path.join(ROOT, "include", "lh_soil.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lh_soil_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    assert "lh_soil_rhs" in names and "lh_soil_step_ssprk33" in names and "lh_soil_budgets_allreduce" in names
    assert set(names) == set(abi.ABI_SYMBOLS), "ctypes binding and header disagree"


def test_cuda_library_exports_every_declared_symbol():
    import __graft_entry__ as graft

    path = graft.build_cuda()
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, f"missing exports: {missing}"
    lib = lh.cuda_library()           # dlopen + resolve through ctypes
    assert lib.soil_abi_version() == abi.ABI_VERSION


def test_cuda_library_has_sm100a_code():
    import __graft_entry__ as graft

    out = subprocess.run(["cuobjdump", "-lelf", graft.build_cuda()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_oracle_exports_same_abi(oracle):
    for name in abi.ABI_SYMBOLS:
        assert hasattr(oracle, name[3:])


def test_struct_layout_matches_c(tmp_path):
    """sizeof/offsetof of the config structs as the C compiler sees them vs ctypes."""
    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "lh_soil.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(lh_soil_params), sizeof(lh_soil_face_bc),"
        " sizeof(lh_soil_config), offsetof(lh_soil_config, params), offsetof(lh_soil_config, top),"
        " offsetof(lh_soil_config, flags), offsetof(lh_soil_params, visc_gamma));return 0;}\n"
    )
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True,
                   env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    vals = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals == [
        C.sizeof(abi.lh_soil_params), C.sizeof(abi.lh_soil_face_bc), C.sizeof(abi.lh_soil_config),
        abi.lh_soil_config.params.offset, abi.lh_soil_config.top.offset, abi.lh_soil_config.flags.offset,
        abi.lh_soil_params.visc_gamma.offset,
    ]


def _no_gpu():
    try:
        import torch

        return not torch.cuda.is_available()
    except Exception:
        return True


def test_create_validation_mirrors_reference_errors(oracle):
    """Config validation happens before any device work and is identical in both libraries:
    @assert zlim[1] < zlim[2] (domain.jl:30); MethodError for BC/component pairs that have no
    vertical_flux method (boundary_conditions.jl:295-444)."""
    D, F, FD, N = abi.LH_BC_DIRICHLET, abi.LH_BC_FLUX, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE
    libs = [oracle, lh.cuda_library()]
    for lib in libs:
        wl = w.coupled_workload(ncol=2, nlayer=4)
        wl.zmin, wl.zmax = 0.0, 0.0
        with pytest.raises(lh.DomainAssertionError):
            lh.SoilContext(lib, wl.config())
        for top in ((FD, 0.0, F, 0.0), (N, 0.0, F, 0.0), (F, 0.0, N, 0.0)):   # FreeDrainage/NoBC on energy, NoBC on water
            wl = w.coupled_workload(ncol=2, nlayer=4, top=top)
            with pytest.raises(lh.UnsupportedBCError):
                lh.SoilContext(lib, wl.config())
        wl = w.richards_workload(ncol=2, nlayer=4, top=(D, 288.0, D, 0.2))    # Dirichlet T on a prescribed-T model
        with pytest.raises(lh.UnsupportedBCError):
            lh.SoilContext(lib, wl.config())
        wl = w.heat_workload(ncol=2, nlayer=4, bottom=(D, 280.0, FD, 0.0))    # FreeDrainage on prescribed hydrology
        with pytest.raises(lh.UnsupportedBCError):
            lh.SoilContext(lib, wl.config())
        cfg = w.coupled_workload(ncol=2, nlayer=4).config()
        h = C.c_void_p()
        cfg.struct_size = 8
        assert lib.soil_create(C.byref(cfg), C.byref(h)) == abi.LH_ERR_INVALID_ARG


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    """Without a CUDA device the product raises NoDeviceError; nothing routes to the CPU."""
    wl = w.coupled_workload(ncol=2, nlayer=4)
    with pytest.raises(lh.NoDeviceError):
        lh.SoilContext(lh.cuda_library(), wl.config())


def test_config_validation_needs_no_device():
    """lh_soil_create validates the config before it looks for a device: the reference constructor stores
    m = 1 - 1/n (SoilWaterParameterizations.jl:162-169) and the closures rely on it."""
    wl = w.richards_workload(ncol=2, nlayer=4)
    wl.params.vg_m = 0.6
    with pytest.raises(lh._abi.SoilError) as e:
        lh.SoilContext(lh.cuda_library(), wl.config())
    assert e.value.status == lh._abi.LH_ERR_INVALID_ARG and "vg_m" in str(e.value)


def test_product_never_references_oracle():
    """Nothing in the package directory names the oracle directory or library."""
    pkg = os.path.join(ROOT, "landhydrology.jl_b200")
    hits = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"liblho_soil|lho_soil|oracle/|np_soil", text):
                    hits.append(os.path.join(dirpath, f))
    assert not hits, hits
    # nor can any product or measurement path load the host-emulated test build (tests/support/hostemu): no Python file of the
    # package, bench.py or __graft_entry__.py mentions it (the csrc headers name it in comments only: lh_ptx.cuh's LH_HOSTEMU guard)
    emu_hits = []
    for path in [os.path.join(d, f) for d, _, fs in os.walk(pkg) for f in fs if f.endswith(".py")] + \
            [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + \
            [os.path.join(ROOT, "tools", f) for f in os.listdir(os.path.join(ROOT, "tools")) if f.endswith(".py")]:
        if re.search(r"hostemu|LH_HOSTEMU", open(path, errors="ignore").read()):
            emu_hits.append(path)
    assert not emu_hits, emu_hits


def test_device_code_is_byte_identical_to_the_last_gpu_verified_build():
    """Everything after commit 24b093f was done without a GPU.  What was allowed to change is HOST code (launch spelling, the
    transfer pipelines of the C ABI, bench.py): every kernel's SASS must still hash to the fingerprint of the build whose `-m gpu`
    suite last ran green on a B200 (profiles/r02_sass_fingerprints_gpu_verified.json, tools/sass_identity.py)."""
    import shutil

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    w.graft.build_cuda()
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_identity.py"), "--check",
                          os.path.join(ROOT, "profiles", "r02_sass_fingerprints_gpu_verified.json")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "268 identical, 0 changed, 0 new, 0 gone" in res.stdout, res.stdout
