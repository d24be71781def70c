"""ctypes binding of the C ABI declared in ``include/lh_soil.h``.

The product library is ``csrc/liblh_soil.so`` (hand-written sm_100a CUDA behind ``extern "C"``).
There is NO CPU fallback: :func:`cuda_library` raises if the shared object is missing, and
``lh_soil_create`` returns ``LH_ERR_NO_DEVICE`` when no B200 is visible.

:class:`SoilLibrary` is generic over (path, symbol prefix) only so that the parity tests can
drive their CPU checker library (same ABI under another prefix) through the very same harness;
nothing in this package ever names or loads it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

ABI_VERSION = 1

# status codes (include/lh_soil.h)
LH_OK = 0
LH_ERR_INVALID_ARG = -1
LH_ERR_DOMAIN = -2
LH_ERR_UNSUPPORTED_BC = -3
LH_ERR_CUDA = -4
LH_ERR_NO_DEVICE = -5
LH_ERR_NCCL = -6
LH_ERR_NONFINITE = -7
LH_ERR_STATE = -8

LH_MODEL_RICHARDS, LH_MODEL_HEAT, LH_MODEL_COUPLED = 0, 1, 2
LH_BC_NONE, LH_BC_FLUX, LH_BC_DIRICHLET, LH_BC_FREE_DRAINAGE = 0, 1, 2, 3
LH_FIELD_THETA_L, LH_FIELD_THETA_I, LH_FIELD_RHO_E_INT, LH_FIELD_T = 0, 1, 2, 3
LH_NUM_FIELDS = 4
LH_DIAG_K, LH_DIAG_PSI, LH_DIAG_KAPPA, LH_DIAG_T = 0, 1, 2, 3
LH_BCV_TOP_ENERGY, LH_BCV_TOP_HYDROLOGY, LH_BCV_BOTTOM_ENERGY, LH_BCV_BOTTOM_HYDROLOGY = 0, 1, 2, 3
LH_FLAG_CHECK_FINITE = 1
LH_FLAG_GENERAL_VG = 2
LH_FLAG_STAGE_LAUNCHES = 4
LH_FLAG_PERSISTENT = 8
LH_FLAG_NO_CHAIN = 16


class SoilError(RuntimeError):
    """Base class; ``status`` holds the LH_ERR_* code."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[lh_soil status {status}] {message}")
        self.status = status


class UnsupportedBCError(SoilError, TypeError):
    """The reference raises ``MethodError`` (no ``vertical_flux`` method) for this pair."""


class DomainAssertionError(SoilError, AssertionError):
    """``@assert zlim[1] < zlim[2]`` (reference src/Domains/domain.jl:30)."""


class NonFiniteStateError(SoilError, ArithmeticError):
    """The reference raises ``DomainError`` from ``^`` once ϑ_l leaves the valid range."""


class NoDeviceError(SoilError):
    """No CUDA device: the product has no CPU path."""


_ERRORS = {
    LH_ERR_UNSUPPORTED_BC: UnsupportedBCError,
    LH_ERR_DOMAIN: DomainAssertionError,
    LH_ERR_NONFINITE: NonFiniteStateError,
    LH_ERR_NO_DEVICE: NoDeviceError,
}


class lh_soil_params(C.Structure):
    _fields_ = [
        ("nu", C.c_double), ("S_s", C.c_double), ("nu_ss_gravel", C.c_double),
        ("nu_ss_om", C.c_double), ("nu_ss_quartz", C.c_double), ("rho_c_ds", C.c_double),
        ("kappa_solid", C.c_double), ("rho_p", C.c_double), ("kappa_sat_unfrozen", C.c_double),
        ("kappa_sat_frozen", C.c_double), ("a", C.c_double), ("b", C.c_double),
        ("kappa_dry_parameter", C.c_double), ("z_0m", C.c_double), ("z_0s", C.c_double),
        ("vg_n", C.c_double), ("vg_alpha", C.c_double), ("vg_m", C.c_double),
        ("theta_r", C.c_double), ("Ksat", C.c_double),
        ("viscosity_factor", C.c_int32), ("impedance_factor", C.c_int32),
        ("visc_gamma", C.c_double), ("visc_T_ref", C.c_double), ("imp_Omega", C.c_double),
        ("rho_cloud_liq", C.c_double), ("rho_cloud_ice", C.c_double), ("cp_l", C.c_double),
        ("cp_i", C.c_double), ("T_0", C.c_double), ("LH_f0", C.c_double), ("K_therm", C.c_double),
    ]


class lh_soil_face_bc(C.Structure):
    _fields_ = [
        ("energy_kind", C.c_int32), ("hydrology_kind", C.c_int32),
        ("energy_value", C.c_double), ("hydrology_value", C.c_double),
    ]


class lh_soil_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("ncol", C.c_int64),
        ("nlayer", C.c_int32), ("model", C.c_int32), ("zmin", C.c_double), ("zmax", C.c_double),
        ("params", lh_soil_params), ("top", lh_soil_face_bc), ("bottom", lh_soil_face_bc),
        ("flags", C.c_int32), ("reserved", C.c_int32),
    ]


LH_MAX_STAGES = 16
LH_STEPPER_SHU_OSHER, LH_STEPPER_2N = 0, 1
LH_METHOD_EULER, LH_METHOD_SSPRK22, LH_METHOD_SSPRK33, LH_METHOD_SSPRK43, LH_METHOD_CK2N54 = 0, 1, 2, 3, 4


class lh_soil_stepper(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("nstages", C.c_int32),
        ("a", C.c_double * LH_MAX_STAGES), ("b", C.c_double * LH_MAX_STAGES),
        ("g", C.c_double * LH_MAX_STAGES), ("c", C.c_double * LH_MAX_STAGES),
    ]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


class lh_soil_atmos(C.Structure):
    """PrescribedAtmosForcing + the constants its fluxes need (include/lh_soil.h)."""
    _fields_ = [("struct_size", C.c_int32), ("reserved", C.c_int32)] + [(n, C.c_double) for n in (
        "u_atm", "theta_atm", "z_atm", "theta_scale", "rho_a_sfc", "q_atm",
        "R_v", "R_d", "grav", "cp_d", "cp_v", "LH_v0", "press_triple", "T_triple", "von_karman", "Pr_0", "a_m", "a_h")]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.struct_size = C.sizeof(lh_soil_atmos)


class lh_soil_run_opts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("save_first", C.c_int32), ("bc_table", _dp),
        ("budget_every", C.c_int64), ("budgets_out", _dp), ("save_every", C.c_int64),
        ("nsave_fields", C.c_int32), ("save_fields", C.c_int32 * LH_NUM_FIELDS), ("reserved", C.c_int32),
        ("save_out", _dp), ("snapshot_stride", C.c_int64), ("field_stride", C.c_int64),
        ("col_stride", C.c_int64), ("layer_stride", C.c_int64),
    ]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.struct_size = C.sizeof(lh_soil_run_opts)


# name -> (argtypes, restype); every entry is declared in include/lh_soil.h
_SIGNATURES = {
    "soil_abi_version": ([], C.c_int32),
    "soil_create": ([C.POINTER(lh_soil_config), C.POINTER(_vp)], C.c_int32),
    "soil_destroy": ([_vp], C.c_int32),
    "soil_last_error": ([_vp], C.c_char_p),
    "soil_get_zc": ([_vp, _dp], C.c_int32),
    "soil_set_state": ([_vp, C.c_int32, _dp, C.c_int64, C.c_int64], C.c_int32),
    "soil_get_state": ([_vp, C.c_int32, _dp, C.c_int64, C.c_int64], C.c_int32),
    "soil_set_aux": ([_vp, C.c_int32, _dp, C.c_int64, C.c_int64], C.c_int32),
    "soil_set_column_params": ([_vp, _dp, _dp, _dp, _dp, _dp], C.c_int32),
    "soil_set_column_heat_params": ([_vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp], C.c_int32),
    "soil_set_cell_params": ([_vp, _dp, _dp, _dp, _dp, _dp, C.c_int64, C.c_int64], C.c_int32),
    "soil_set_column_fluxes": ([_vp, C.POINTER(_dp)], C.c_int32),
    "soil_set_atmos_forcing": ([_vp, C.POINTER(lh_soil_atmos)], C.c_int32),
    "soil_atmos_fluxes": ([_vp, _dp, _dp, _dp, C.c_int64, _dp, _dp], C.c_int32),
    "soil_set_bc_values": ([_vp, _dp], C.c_int32),
    "soil_rhs": ([_vp, C.c_double], C.c_int32),
    "soil_get_tendency": ([_vp, C.c_int32, _dp, C.c_int64, C.c_int64], C.c_int32),
    "soil_stage_ssprk33": ([_vp, C.c_int32, C.c_double], C.c_int32),
    "soil_step_ssprk33": ([_vp, C.c_double, C.c_double, C.c_int64, _dp], C.c_int32),
    "soil_stepper_named": ([C.c_int32, C.POINTER(lh_soil_stepper)], C.c_int32),
    "soil_step": ([_vp, C.POINTER(lh_soil_stepper), C.c_double, C.c_double, C.c_int64, _dp], C.c_int32),
    "soil_budgets": ([_vp, _dp], C.c_int32),
    "soil_set_aux_table": ([_vp, C.c_int32, _dp, C.c_int64], C.c_int32),
    "soil_run": ([_vp, C.c_double, C.c_double, C.c_int64, C.POINTER(lh_soil_run_opts)], C.c_int32),
    "soil_checkpoint_bytes": ([_vp], C.c_int64),
    "soil_checkpoint_save": ([_vp, _vp, C.c_int64], C.c_int32),
    "soil_checkpoint_load": ([_vp, _vp, C.c_int64], C.c_int32),
    "soil_alloc_host": ([C.c_int64, C.POINTER(_vp)], C.c_int32),
    "soil_free_host": ([_vp], C.c_int32),
    "soil_budgets_async": ([_vp, C.POINTER(C.c_int64)], C.c_int32),
    "soil_budgets_wait": ([_vp, C.c_int64, _dp], C.c_int32),
    "soil_diagnostic": ([_vp, C.c_int32, _dp, C.c_int64, C.c_int64], C.c_int32),
    "soil_sync": ([_vp], C.c_int32),
    "soil_eval_math": ([_vp, C.c_int32, _dp, _dp, C.c_int64], C.c_int32),
    "soil_last_step_timing": ([_vp, _dp, C.POINTER(C.c_int64)], C.c_int32),
    "soil_kernel_info": ([_vp, C.c_char_p, C.c_int64], C.c_int32),
    "soil_device_ptr": ([_vp, C.c_int32, C.POINTER(_vp), C.POINTER(C.c_int64)], C.c_int32),
    "soil_comm_unique_id": ([C.POINTER(C.c_uint8)], C.c_int32),
    "soil_comm_init": ([_vp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)], C.c_int32),
    "soil_budgets_allreduce": ([_vp, _dp], C.c_int32),
}

ABI_SYMBOLS = tuple("lh_" + k for k in _SIGNATURES)


def _as_double_ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class SoilLibrary:
    """A loaded shared object exporting the lh_soil C ABI under ``prefix``."""

    def __init__(self, path: str, prefix: str = "lh_"):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'); "
                "this package has no CPU fallback"
            )
        self.path = os.path.abspath(path)
        self.prefix = prefix
        self._dll = C.CDLL(self.path, mode=C.RTLD_GLOBAL if prefix == "lh_" else C.RTLD_LOCAL)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(self._dll, prefix + name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = restype
            setattr(self, name, fn)
        v = self.soil_abi_version()
        if v != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI version {v}, expected {ABI_VERSION}")

    def raw(self, symbol: str):
        """Any other exported symbol (used by the tests for the oracle's scalar closures)."""
        return getattr(self._dll, symbol)

    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        st = self.soil_comm_unique_id(buf)
        if st != LH_OK:
            raise SoilError(st, "lh_soil_comm_unique_id failed (is libnccl.so.2 loadable?)")
        return bytes(buf)

    def __repr__(self):
        return f"SoilLibrary({self.path!r}, prefix={self.prefix!r})"


_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIBRARY_PATH = os.path.join(_HERE, "csrc", "liblh_soil.so")
_cuda_lib: Optional[SoilLibrary] = None


def cuda_library() -> SoilLibrary:
    """The product library.  Raises (never falls back) when it has not been built."""
    global _cuda_lib
    if _cuda_lib is None:
        # LH_SOIL_LIBRARY: an alternative BUILD of the same CUDA library (kernel tuning experiments)
        _cuda_lib = SoilLibrary(os.environ.get("LH_SOIL_LIBRARY", CUDA_LIBRARY_PATH), "lh_")
    return _cuda_lib


class SoilContext:
    """Owner of one ``lh_soil_ctx``: a shard of ``ncol`` columns on one device."""

    def __init__(self, lib: SoilLibrary, cfg: lh_soil_config):
        self.lib = lib
        self.cfg = cfg
        self.ncol = int(cfg.ncol)
        self.nlayer = int(cfg.nlayer)
        self.model = int(cfg.model)
        self._h = _vp()
        cfg.struct_size = C.sizeof(lh_soil_config)
        st = lib.soil_create(C.byref(cfg), C.byref(self._h))
        if st != LH_OK:
            msg = lib.soil_last_error(None)
            self._h = _vp()
            self._raise(st, msg.decode() if msg else "lh_soil_create failed")

    # -- plumbing ---------------------------------------------------------------------------
    def _raise(self, st: int, msg: Optional[str] = None):
        if msg is None:
            raw = self.lib.soil_last_error(self._h)
            msg = raw.decode() if raw else ""
        raise _ERRORS.get(st, SoilError)(st, msg)

    def _check(self, st: int):
        if st != LH_OK:
            self._raise(st)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.soil_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- array helpers: shape (ncol, nlayer) C-order == reference layout (layer fastest) ----
    def _strides(self, a: np.ndarray):
        if a.dtype != np.float64:
            raise TypeError("fp64 arrays only")
        if a.ndim == 1:
            if self.ncol != 1 or a.shape[0] != self.nlayer:
                raise ValueError(f"expected shape ({self.nlayer},) for a single column, got {a.shape}")
            return 0, a.strides[0] // 8
        if a.shape != (self.ncol, self.nlayer):
            raise ValueError(f"expected shape ({self.ncol}, {self.nlayer}), got {a.shape}")
        if a.strides[0] % 8 or a.strides[1] % 8:
            raise ValueError("unaligned strides")
        return a.strides[0] // 8, a.strides[1] // 8

    def zc(self) -> np.ndarray:
        out = np.empty(self.nlayer, dtype=np.float64)
        self._check(self.lib.soil_get_zc(self._h, _as_double_ptr(out)))
        return out

    def set_state(self, field: int, a: np.ndarray):
        cs, ls = self._strides(a)
        self._check(self.lib.soil_set_state(self._h, field, _as_double_ptr(a), cs, ls))

    def set_aux(self, field: int, a: np.ndarray, per_layer: bool = False):
        if per_layer:
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (self.nlayer,):
                raise ValueError(f"per-layer profile must have shape ({self.nlayer},)")
            self._check(self.lib.soil_set_aux(self._h, field, _as_double_ptr(a), 0, 1))
        else:
            cs, ls = self._strides(a)
            self._check(self.lib.soil_set_aux(self._h, field, _as_double_ptr(a), cs, ls))

    def _empty(self, like: Optional[np.ndarray]) -> np.ndarray:
        if like is not None:
            return like
        return np.empty((self.ncol, self.nlayer), dtype=np.float64)

    def get_state(self, field: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = self._empty(out)
        cs, ls = self._strides(out)
        self._check(self.lib.soil_get_state(self._h, field, _as_double_ptr(out), cs, ls))
        return out

    def get_tendency(self, field: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = self._empty(out)
        cs, ls = self._strides(out)
        self._check(self.lib.soil_get_tendency(self._h, field, _as_double_ptr(out), cs, ls))
        return out

    def diagnostic(self, which: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = self._empty(out)
        cs, ls = self._strides(out)
        self._check(self.lib.soil_diagnostic(self._h, which, _as_double_ptr(out), cs, ls))
        return out

    def set_column_params(self, nu=None, theta_r=None, vg_n=None, vg_alpha=None, Ksat=None):
        """Per-column hydraulic parameters (``lh_soil_set_column_params``); ``None`` keeps the model's scalar."""
        ptrs, keep = [], []
        for a in (nu, theta_r, vg_n, vg_alpha, Ksat):
            if a is None:
                ptrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (self.ncol,):
                raise ValueError(f"per-column parameter must have shape ({self.ncol},)")
            keep.append(a)
            ptrs.append(_as_double_ptr(a))
        self._check(self.lib.soil_set_column_params(self._h, *ptrs))

    def set_column_heat_params(self, rho_c_ds=None, kappa_sat_unfrozen=None, kappa_sat_frozen=None, kappa_solid=None,
                               nu_ss_om=None, nu_ss_quartz=None, nu_ss_gravel=None):
        """``lh_soil_set_column_heat_params``: per-column heat parameters (arrays of ncol doubles; None keeps the scalar)."""
        arrs = []
        for a in (rho_c_ds, kappa_sat_unfrozen, kappa_sat_frozen, kappa_solid, nu_ss_om, nu_ss_quartz, nu_ss_gravel):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (self.ncol,):
                raise ValueError(f"per-column parameter must have shape ({self.ncol},)")
            arrs.append(a)
        self._check(self.lib.soil_set_column_heat_params(self._h, *[None if a is None else _as_double_ptr(a) for a in arrs]))

    def set_cell_params(self, nu=None, theta_r=None, vg_n=None, vg_alpha=None, Ksat=None):
        """``lh_soil_set_cell_params``: per-cell (layered) hydraulic parameters, arrays of shape (ncol, nlayer); None keeps the
        per-column value or the model's scalar."""
        arrs = []
        for a in (nu, theta_r, vg_n, vg_alpha, Ksat):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (self.ncol, self.nlayer):
                raise ValueError(f"per-cell parameter must have shape ({self.ncol}, {self.nlayer})")
            arrs.append(a)
        self._check(self.lib.soil_set_cell_params(self._h, *[None if a is None else _as_double_ptr(a) for a in arrs], self.nlayer, 1))

    def set_column_fluxes(self, top_energy=None, top_hydrology=None, bottom_energy=None, bottom_hydrology=None):
        """``lh_soil_set_column_fluxes``: per-column VerticalFlux values (arrays of ncol doubles) for faces of kind LH_BC_FLUX."""
        arrs = []
        for a in (top_energy, top_hydrology, bottom_energy, bottom_hydrology):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (self.ncol,):
                raise ValueError(f"per-column flux must have shape ({self.ncol},)")
            arrs.append(a)
        ptrs = (_dp * 4)(*[None if a is None else _as_double_ptr(a) for a in arrs])
        self._check(self.lib.soil_set_column_fluxes(self._h, ptrs))

    def set_atmos_forcing(self, atmos: Optional["lh_soil_atmos"]):
        """``lh_soil_set_atmos_forcing``: PrescribedAtmosForcing at the top face (None: back to the configured top BC)."""
        self._check(self.lib.soil_set_atmos_forcing(self._h, None if atmos is None else C.byref(atmos)))

    def atmos_fluxes(self, theta_l, theta_i, T):
        """``compute_turbulent_surface_fluxes`` for given surface states: (heat flux, water volume flux) arrays."""
        th, ti, T = (np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64) for x in (theta_l, theta_i, T))
        heat, water = np.empty_like(th), np.empty_like(th)
        self._check(self.lib.soil_atmos_fluxes(self._h, _as_double_ptr(th), _as_double_ptr(ti), _as_double_ptr(T), th.size,
                                                _as_double_ptr(heat), _as_double_ptr(water)))
        return heat, water

    def set_bc_values(self, values: Sequence[float]):
        v = np.asarray(values, dtype=np.float64)
        if v.shape != (4,):
            raise ValueError("bc values: 4 doubles [top energy, top hydrology, bottom energy, bottom hydrology]")
        self._check(self.lib.soil_set_bc_values(self._h, _as_double_ptr(v)))

    # -- hot path ---------------------------------------------------------------------------
    def rhs(self, t: float = 0.0):
        self._check(self.lib.soil_rhs(self._h, float(t)))

    def stage(self, stage: int, dt: float):
        self._check(self.lib.soil_stage_ssprk33(self._h, int(stage), float(dt)))

    def step(self, t: float, dt: float, nsteps: int = 1, bc_table: Optional[np.ndarray] = None):
        if bc_table is not None:
            bc_table = np.ascontiguousarray(bc_table, dtype=np.float64)
            if bc_table.size != nsteps * 12:
                raise ValueError("bc_table must hold nsteps*3*4 doubles")
            ptr = _as_double_ptr(bc_table)
        else:
            ptr = None
        self._check(self.lib.soil_step_ssprk33(self._h, float(t), float(dt), int(nsteps), ptr))

    def step_with(self, stepper: "lh_soil_stepper", t: float, dt: float, nsteps: int = 1,
                  bc_table: Optional[np.ndarray] = None):
        """``lh_soil_step``: nsteps of a Shu-Osher / 2N low-storage stepper (one launch per stage)."""
        if bc_table is not None:
            bc_table = np.ascontiguousarray(bc_table, dtype=np.float64)
            if bc_table.size != nsteps * stepper.nstages * 4:
                raise ValueError("bc_table must hold nsteps*nstages*4 doubles")
            ptr = _as_double_ptr(bc_table)
        else:
            ptr = None
        self._check(self.lib.soil_step(self._h, C.byref(stepper), float(t), float(dt), int(nsteps), ptr))

    def budgets(self) -> np.ndarray:
        out = np.empty(2, dtype=np.float64)
        self._check(self.lib.soil_budgets(self._h, _as_double_ptr(out)))
        return out

    def set_aux_table(self, field: int, table: Optional[np.ndarray]):
        """``lh_soil_set_aux_table``: rows of a prescribed profile for the coming stage launches ([nrows, nlayer])."""
        if table is None:
            self._check(self.lib.soil_set_aux_table(self._h, int(field), None, 0))
            return
        table = np.ascontiguousarray(table, dtype=np.float64)
        if table.ndim != 2 or table.shape[1] != self.nlayer:
            raise ValueError(f"aux table must have shape (nrows, {self.nlayer})")
        self._check(self.lib.soil_set_aux_table(self._h, int(field), _as_double_ptr(table), table.shape[0]))

    def run(self, t0: float, dt: float, nsteps: int, *, bc_table: Optional[np.ndarray] = None, budget_every: int = 0,
            save_every: int = 0, save_first: bool = False, save_fields: Sequence[int] = (), save_out: Optional[np.ndarray] = None):
        """``lh_soil_run``: nsteps SSPRK33 steps with budgets every ``budget_every`` and snapshots every ``save_every`` steps
        in one call.  Returns ``(budgets [nb, 2] or None, snapshots [ns, nfields, ncol, nlayer] or None)``; ``save_out``
        may be a preallocated (ideally pinned) array of that shape."""
        o = lh_soil_run_opts()
        keep = []
        if bc_table is not None:
            bc_table = np.ascontiguousarray(bc_table, dtype=np.float64)
            if bc_table.size != nsteps * 12:
                raise ValueError("bc_table must hold nsteps*3*4 doubles")
            o.bc_table = _as_double_ptr(bc_table)
            keep.append(bc_table)
        budgets = None
        if budget_every > 0:
            budgets = np.zeros((nsteps // budget_every, 2), dtype=np.float64)
            o.budget_every = int(budget_every)
            o.budgets_out = _as_double_ptr(budgets)
        snaps = None
        fields = [int(f) for f in save_fields]
        if (save_every > 0 or save_first) and fields:
            ns = (nsteps // save_every if save_every > 0 else 0) + (1 if save_first else 0)
            shape = (ns, len(fields), self.ncol, self.nlayer)
            if save_out is not None:
                if save_out.shape != shape or save_out.dtype != np.float64 or not save_out.flags.c_contiguous:
                    raise ValueError(f"save_out must be a C-contiguous float64 array of shape {shape}")
                snaps = save_out
            else:
                snaps = np.empty(shape, dtype=np.float64)
            o.save_every = int(save_every)
            o.save_first = 1 if save_first else 0
            o.nsave_fields = len(fields)
            for k, f in enumerate(fields):
                o.save_fields[k] = f
            o.save_out = _as_double_ptr(snaps)
            o.snapshot_stride = len(fields) * self.ncol * self.nlayer
            o.field_stride = self.ncol * self.nlayer
            o.col_stride, o.layer_stride = self.nlayer, 1
        self._check(self.lib.soil_run(self._h, float(t0), float(dt), int(nsteps), C.byref(o)))
        return budgets, snaps

    def checkpoint(self) -> np.ndarray:
        """``lh_soil_checkpoint_save`` into a fresh byte array."""
        n = int(self.lib.soil_checkpoint_bytes(self._h))
        buf = np.empty(n, dtype=np.uint8)
        self._check(self.lib.soil_checkpoint_save(self._h, buf.ctypes.data_as(_vp), n))
        return buf

    def restore(self, buf: np.ndarray):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self._check(self.lib.soil_checkpoint_load(self._h, buf.ctypes.data_as(_vp), buf.size))

    def budgets_async(self) -> int:
        """``lh_soil_budgets_async``: enqueue the budget read behind the work already on the stream; returns a ticket."""
        t = C.c_int64()
        self._check(self.lib.soil_budgets_async(self._h, C.byref(t)))
        return t.value

    def budgets_wait(self, ticket: int) -> np.ndarray:
        out = np.empty(2, dtype=np.float64)
        self._check(self.lib.soil_budgets_wait(self._h, int(ticket), _as_double_ptr(out)))
        return out

    def budgets_allreduce(self) -> np.ndarray:
        out = np.empty(2, dtype=np.float64)
        self._check(self.lib.soil_budgets_allreduce(self._h, _as_double_ptr(out)))
        return out

    def sync(self):
        self._check(self.lib.soil_sync(self._h))

    def eval_math(self, fn: int, x: np.ndarray) -> np.ndarray:
        """y = f(x) with the library's own elementary functions (LH_MATH_*); DIV: x = [num..., den...]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.size // 2 if fn == 6 else x.size
        y = np.empty(n, dtype=np.float64)
        self._check(self.lib.soil_eval_math(self._h, int(fn), _as_double_ptr(x), _as_double_ptr(y), n))
        return y

    def last_step_timing(self):
        ms = C.c_double()
        n = C.c_int64()
        self._check(self.lib.soil_last_step_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_info(self) -> str:
        """``lh_soil_kernel_info``: the kernel variant, launch shape and strategy the next step call will use."""
        buf = C.create_string_buffer(512)
        self._check(self.lib.soil_kernel_info(self._h, buf, 512))
        return buf.value.decode()

    def device_ptr(self, field: int):
        p = _vp()
        n = C.c_int64()
        self._check(self.lib.soil_device_ptr(self._h, field, C.byref(p), C.byref(n)))
        return p.value, n.value

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        if len(unique_id) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.soil_comm_init(self._h, int(nranks), int(rank), buf))
