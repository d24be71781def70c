"""State containers and ``initialize_states``.

Mirror of reference src/SoilModel/initial_conditions.jl:14-107.  ``FieldVector`` stands in for
``ClimaCore.Fields.FieldVector``: ``Y.soil.ϑ_l`` is a NumPy array of shape ``(nlayer,)`` for a
``Column`` (index 0 = bottom, like ``parent(Y.soil.ϑ_l)[:]``) and ``(ncolumns, nlayer)`` for a
``HybridBox`` — C-order, i.e. the layer index is fastest, which is the reference's per-column
layout and what ``lh_soil_set_state`` takes with ``layer_stride = 1``.
"""
from __future__ import annotations

import unicodedata
from typing import Callable, Dict, Iterable, Mapping

import numpy as np

from .domains import make_function_space
from .models import PrescribedHydrologyModel, PrescribedTemperatureModel, SoilModel


def nf(name: str) -> str:
    """NFKC-normalise a field name exactly like Python normalises identifiers (ϑ -> θ)."""
    return unicodedata.normalize("NFKC", name)


class NamedFields:
    """A NamedTuple-of-fields: ``x.ϑ_l``, ``x["ϑ_l"]``, iteration in insertion order."""

    def __init__(self, **fields):
        object.__setattr__(self, "_names", [])
        for k, v in fields.items():
            self._set(k, v)

    def _set(self, name, value):
        name = nf(name)
        if name not in self._names:
            self._names.append(name)
        object.__setattr__(self, name, value)

    def __setattr__(self, name, value):
        self._set(name, value)

    def __getitem__(self, name):
        return getattr(self, nf(name))

    def __contains__(self, name):
        return nf(name) in self._names

    def keys(self):
        return list(self._names)

    def items(self):
        return [(k, getattr(self, k)) for k in self._names]

    def __len__(self):
        return len(self._names)

    def __eq__(self, other):
        return (
            isinstance(other, NamedFields)
            and self._names == other._names
            and all(np.array_equal(getattr(self, k), getattr(other, k)) for k in self._names)
        )

    def __repr__(self):
        return "NamedFields(" + ", ".join(f"{k}={getattr(self, k).shape}" for k in self._names) + ")"


class FieldVector:
    """``Fields.FieldVector(; zc = ..., soil = ...)`` stand-in."""

    def __init__(self, **components):
        object.__setattr__(self, "_names", [])
        for k, v in components.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        name = nf(name)
        if name not in self._names:
            self._names.append(name)
        object.__setattr__(self, name, value)

    def keys(self):
        return list(self._names)

    def __eq__(self, other):
        if not isinstance(other, FieldVector) or self._names != other._names:
            return False
        for k in self._names:
            a, b = getattr(self, k), getattr(other, k)
            if isinstance(a, np.ndarray):
                if not np.array_equal(a, b):
                    return False
            elif a != b:
                return False
        return True

    def __repr__(self):
        return "FieldVector(" + ", ".join(self._names) + ")"


def parent(x):
    """``parent(field)``: the raw array."""
    return x


def similar(Y: FieldVector) -> FieldVector:
    """``similar(Y)``: same structure, uninitialised (here zero-filled) storage."""
    out = FieldVector()
    for k in Y.keys():
        v = getattr(Y, k)
        if isinstance(v, NamedFields):
            setattr(out, k, NamedFields(**{n: np.zeros_like(a) for n, a in v.items()}))
        else:
            setattr(out, k, np.zeros_like(v))
    return out


def copy(Y: FieldVector) -> FieldVector:
    out = FieldVector()
    for k in Y.keys():
        v = getattr(Y, k)
        if isinstance(v, NamedFields):
            setattr(out, k, NamedFields(**{n: a.copy() for n, a in v.items()}))
        else:
            setattr(out, k, v.copy())
    return out


def coordinates(cs) -> np.ndarray:
    """right_hand_side.jl:7-8: the z coordinates of the centre space."""
    return cs.z


def _field_shape(model: SoilModel):
    n = model.domain.nelements
    ncol = model.domain.ncolumns
    return (n,) if ncol == 1 and model.domain.column_shape == () else (ncol, n)


def _broadcast_over_cells(fn: Callable, zc: np.ndarray, shape, names_hint=None) -> Dict[str, np.ndarray]:
    """Evaluate ``fn(z)`` (returning a mapping of scalars) at every centre and stack by name.

    A function may declare ``fn.vectorized = True`` to be called ONCE with ``zc`` broadcast to
    the full field shape and return arrays (needed for million-column initial conditions).
    """
    if getattr(fn, "vectorized", False):
        z = np.broadcast_to(zc, shape)
        res = fn(z)
        return {nf(k): np.array(np.broadcast_to(np.asarray(v, dtype=np.float64), shape), dtype=np.float64, order="C") for k, v in dict(res).items()}
    per_layer = [dict(fn(float(z))) for z in zc]
    names = list(per_layer[0].keys()) if per_layer else list(names_hint or [])
    out = {}
    for k in names:
        col = np.array([d[k] for d in per_layer], dtype=np.float64)
        out[nf(k)] = np.array(np.broadcast_to(col, shape), dtype=np.float64, order="C")
    return out


def aux_vars(model_or_component):
    """initial_conditions.jl:27-77: a function (t, z) -> mapping of auxiliary values."""
    m = model_or_component
    if isinstance(m, SoilModel):
        fe, fh = aux_vars(m.energy_model), aux_vars(m.hydrology_model)
        return lambda t, z: {**fe(t, z), **fh(t, z)}
    if isinstance(m, PrescribedTemperatureModel):
        return lambda t, z: {"T": m.T_profile(z, t)}
    if isinstance(m, PrescribedHydrologyModel):
        return lambda t, z: {"ϑ_l": m.ϑ_l_profile(z, t), "θ_i": m.θ_i_profile(z, t)}
    return lambda t, z: {}


def initialize_auxiliary(model: SoilModel, t0: float, zc: np.ndarray) -> FieldVector:
    """initial_conditions.jl:14-17"""
    init_aux_soil = aux_vars(model)
    shape = _field_shape(model)
    fields = _broadcast_over_cells(lambda z: init_aux_soil(t0, z), zc, shape)
    return FieldVector(**{"zc": np.array(zc, dtype=np.float64), model.name: NamedFields(**fields)})


def initialize_prognostic(model: SoilModel, f: Callable, zc: np.ndarray) -> FieldVector:
    """initial_conditions.jl:85-89: ``f.(zc, Ref(model))``."""
    shape = _field_shape(model)
    g = lambda z: f(z, model)
    if getattr(f, "vectorized", False):
        g.vectorized = True
    fields = _broadcast_over_cells(g, zc, shape)
    return FieldVector(**{model.name: NamedFields(**fields)})


def initialize_states(model: SoilModel, f: Callable, t0: float):
    """initial_conditions.jl:101-107"""
    space_c, _ = make_function_space(model.domain)
    zc = coordinates(space_c)
    Y0 = initialize_prognostic(model, f, zc)
    Ya0 = initialize_auxiliary(model, t0, zc)
    return Y0, Ya0
