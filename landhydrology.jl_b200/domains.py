"""Domains: ``Column`` (reference src/Domains/domain.jl:12-69) and ``HybridBox``.

``HybridBox`` does not exist in the reference snapshot; BASELINE.json names it.  It is defined
here as nx*ny laterally independent columns sharing one vertical mesh: every reference RHS uses
vertical (C2F/F2C) operators only (right_hand_side.jl:170-181, 249-259, 337-365), so each of its
columns must evolve exactly like a ``Column`` with the same ``zlim``/``nelements``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np


class AbstractDomain:
    pass


class AbstractVerticalDomain(AbstractDomain):
    pass


def _check_zlim(zlim):
    # domain.jl:30  @assert zlim[1] < zlim[2]
    assert zlim[0] < zlim[1], "zlim[1] < zlim[2]"


@dataclass(frozen=True)
class Column(AbstractVerticalDomain):
    """``Column(FT; zlim, nelements)``; boundary tags are (:bottom, :top) (domain.jl:29-33)."""

    zlim: Tuple[float, float]
    nelements: int
    boundary_tags: Tuple[str, str] = ("bottom", "top")
    FT: type = np.float64

    def __init__(self, FT=np.float64, *, zlim, nelements):
        _check_zlim(zlim)
        object.__setattr__(self, "FT", FT)
        object.__setattr__(self, "zlim", (FT(zlim[0]), FT(zlim[1])))
        object.__setattr__(self, "nelements", int(np.int32(nelements)))
        object.__setattr__(self, "boundary_tags", ("bottom", "top"))

    @property
    def ncolumns(self) -> int:
        return 1

    @property
    def column_shape(self) -> Tuple[int, ...]:
        return ()

    # Base.ndims / Base.length / Base.size (domain.jl:35-39)
    def ndims(self) -> int:
        return 1

    def length(self):
        return self.zlim[1] - self.zlim[0]

    def size(self):
        return self.length()

    def __len__(self):  # pragma: no cover - convenience only
        raise TypeError("use length(domain): the reference's length is a float")

    def __str__(self):  # Base.show, domain.jl:41-49
        return "[%0.1f, %0.1f]" % (self.zlim[0], self.zlim[1])


@dataclass(frozen=True)
class HybridBox(AbstractVerticalDomain):
    """``HybridBox(FT; xlim, ylim, zlim, nelements = (nx, ny, nz))``: nx*ny independent columns."""

    xlim: Tuple[float, float]
    ylim: Tuple[float, float]
    zlim: Tuple[float, float]
    nelements3: Tuple[int, int, int]
    boundary_tags: Tuple[str, str] = ("bottom", "top")
    FT: type = np.float64

    def __init__(self, FT=np.float64, *, xlim=(0.0, 1.0), ylim=(0.0, 1.0), zlim, nelements):
        _check_zlim(zlim)
        assert xlim[0] < xlim[1] and ylim[0] < ylim[1]
        nx, ny, nz = (int(v) for v in nelements)
        assert nx >= 1 and ny >= 1 and nz >= 1
        object.__setattr__(self, "FT", FT)
        object.__setattr__(self, "xlim", (FT(xlim[0]), FT(xlim[1])))
        object.__setattr__(self, "ylim", (FT(ylim[0]), FT(ylim[1])))
        object.__setattr__(self, "zlim", (FT(zlim[0]), FT(zlim[1])))
        object.__setattr__(self, "nelements3", (nx, ny, nz))
        object.__setattr__(self, "boundary_tags", ("bottom", "top"))

    @property
    def nelements(self) -> int:
        """Vertical element count (what the soil RHS sees)."""
        return self.nelements3[2]

    @property
    def ncolumns(self) -> int:
        return self.nelements3[0] * self.nelements3[1]

    @property
    def column_shape(self) -> Tuple[int, ...]:
        return (self.nelements3[0], self.nelements3[1])

    def ndims(self) -> int:
        return 3

    def length(self):
        return self.zlim[1] - self.zlim[0]

    def size(self):
        return (self.xlim[1] - self.xlim[0], self.ylim[1] - self.ylim[0], self.length())


def ndims(domain) -> int:
    return domain.ndims()


def length(domain):
    return domain.length()


def size(domain):
    return domain.size()


@dataclass(frozen=True)
class CenterFiniteDifferenceSpace:
    """What the soil path needs from ClimaCore's centre space: z_c and Δz of a uniform mesh."""

    zmin: float
    zmax: float
    nelements: int

    @property
    def Δz(self) -> float:
        return (self.zmax - self.zmin) / self.nelements

    @property
    def z(self) -> np.ndarray:
        n = self.nelements
        j = np.arange(n + 1, dtype=np.float64)
        zf = self.zmin + (self.zmax - self.zmin) * j / n
        return (zf[:-1] + zf[1:]) / 2.0


@dataclass(frozen=True)
class FaceFiniteDifferenceSpace:
    zmin: float
    zmax: float
    nelements: int

    @property
    def z(self) -> np.ndarray:
        n = self.nelements
        j = np.arange(n + 1, dtype=np.float64)
        return self.zmin + (self.zmax - self.zmin) * j / n


def make_function_space(domain: AbstractVerticalDomain):
    """domain.jl:58-69: uniform IntervalMesh -> (center_space, face_space)."""
    zmin, zmax = float(domain.zlim[0]), float(domain.zlim[1])
    n = int(domain.nelements)
    return CenterFiniteDifferenceSpace(zmin, zmax, n), FaceFiniteDifferenceSpace(zmin, zmax, n)
