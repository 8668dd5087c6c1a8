"""Problem descriptors: the numerics of the reference's ``example_problems`` re-expressed as data.

The reference defines its problems as ExaSlang files that only the (Java) ExaStencils generator can
read; nothing of them is copied here.  Each descriptor cites the file:line it restates and yields

* the rediscretised system operator of every level as a coefficient table ``[nf, nf, 27]``
  (index :func:`evostencils_b200.oplist.stencil_index`),
* the initial guess (with Dirichlet boundary values) and the right-hand side of the finest level,
* the settings of the generated solver (``generate solver`` block: tolerance, iteration cap,
  default smoother, coarse-grid solver).

Host arrays are dense, C-ordered ``[z, y, x]`` (x = ExaStencils ``i0`` fastest), ``2^l + 1``
nodes per dimension with the boundary layer stored (reference: exastencils.py:97-103).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import oplist as ol


@dataclass
class SolverSettings:
    """`generate solver` block of the .exa3 files."""
    tol: float = 1e-12           # solver_targetResReduction
    max_iters: int = 100         # solver_maxNumIts
    num_pre: int = 2             # solver_smoother_numPre
    num_post: int = 1            # solver_smoother_numPost
    damping: float = 1.0         # solver_smoother_damping
    red_black: bool = True       # solver_smoother_coloring = "red-black", jacobiType = false
    cgs_max_iters: int = 1000    # solver_cgs_maxNumIts
    cgs_tol: float = 1e-12       # solver_cgs_targetResReduction


@dataclass
class Problem:
    name: str
    dim: int
    fields: Tuple[str, ...]
    rhs_names: Tuple[str, ...]
    equation_names: Tuple[str, ...]
    min_level: int
    max_level: int
    settings: SolverSettings = field(default_factory=SolverSettings)
    kind: int = ol.PROBLEM_LINEAR
    complex_valued: bool = False
    gamma: float = 0.0           # FAS
    wave_number: complex = 0.0   # Helmholtz k
    parameters: Dict[str, float] = field(default_factory=dict)

    # ---- to be provided by subclasses ---------------------------------------------------------
    def operator(self, level: int) -> np.ndarray:            # [nf, nf, 27]
        raise NotImplementedError

    def boundary_value(self, field_index: int, level: int, *coords: np.ndarray) -> Optional[np.ndarray]:
        """Dirichlet value on the boundary nodes; None = homogeneous."""
        return None

    def rhs_value(self, field_index: int, level: int, *coords: np.ndarray) -> Optional[np.ndarray]:
        """Right-hand side at the nodes; None = zero."""
        return None

    # ---- derived ------------------------------------------------------------------------------
    @property
    def n_fields(self) -> int:
        return len(self.fields)

    @property
    def dtype(self):
        return np.complex128 if self.complex_valued else np.float64

    def with_levels(self, min_level: int, max_level: int) -> "Problem":
        import copy
        p = copy.copy(self)
        p.min_level, p.max_level = int(min_level), int(max_level)
        return p

    def nodes(self, level: int) -> int:
        return (1 << level) + 1

    def spacing(self, level: int) -> float:
        return 1.0 / float(1 << level)

    def restrict_weights(self) -> np.ndarray:
        return ol.full_weighting(self.dim)

    def prolong_weights(self) -> np.ndarray:
        return ol.linear_interpolation(self.dim)

    def initial_solution(self, field_index: int, level: Optional[int] = None) -> np.ndarray:
        """u = 0 inside, Dirichlet values on the boundary layer (InitFields + `apply bc`)."""
        level = self.max_level if level is None else level
        n = self.nodes(level)
        h = self.spacing(level)
        u = np.zeros((n,) * self.dim, dtype=self.dtype)
        ax = np.arange(n, dtype=np.float64) * h
        # evaluate the boundary function face by face (cheap also for 513^3)
        for d in range(self.dim):                      # d counts from the slowest axis
            for side in (0, n - 1):
                sl = [slice(None)] * self.dim
                sl[d] = side
                grids = []
                for dd in range(self.dim):
                    grids.append(np.array([ax[side]]) if dd == d else ax)
                mesh = np.meshgrid(*grids, indexing="ij")
                coords = tuple(reversed(mesh))         # (x, y[, z]) : last axis is x
                val = self.boundary_value(field_index, level, *coords)
                if val is not None:
                    u[tuple(sl)] = np.squeeze(np.asarray(val, dtype=self.dtype), axis=d)
        return u

    def rhs(self, field_index: int, level: Optional[int] = None) -> np.ndarray:
        level = self.max_level if level is None else level
        n = self.nodes(level)
        h = self.spacing(level)
        ax = np.arange(n, dtype=np.float64) * h
        probe = self.rhs_value(field_index, level, *(np.zeros(1),) * self.dim)
        if probe is None:
            return np.zeros((n,) * self.dim, dtype=self.dtype)
        mesh = np.meshgrid(*([ax] * self.dim), indexing="ij", sparse=True)
        coords = tuple(reversed(mesh))
        val = np.asarray(self.rhs_value(field_index, level, *coords), dtype=self.dtype)
        return np.ascontiguousarray(np.broadcast_to(val, (n,) * self.dim))


def _table(nf: int, complex_valued: bool = False) -> np.ndarray:
    return np.zeros((nf, nf, ol.STENCIL_POINTS), dtype=np.complex128 if complex_valued else np.float64)


# ------------------------------------------------------------------------------------------------
class Poisson2D(Problem):
    """example_problems/Poisson/2D_FD_Poisson_fromL2.exa2:2-19, .exa3:2-15, .knowledge:1-4."""

    def __init__(self, min_level: int = 5, max_level: int = 9):
        super().__init__(name="2D_FD_Poisson_fromL2", dim=2, fields=("u",), rhs_names=("RHS_u",),
                         equation_names=("solEq",), min_level=min_level, max_level=max_level,
                         settings=SolverSettings(damping=1.15))

    def operator(self, level):
        # Laplace: centre 2/hx^2 + 2/hy^2, neighbours -1/h^2          (exa2:9-15)
        h = self.spacing(level)
        t = _table(1)
        t[0, 0, ol.stencil_index((0, 0))] = 2.0 / (h ** 2) + 2.0 / (h ** 2)
        for o in ((-1, 0), (1, 0), (0, -1), (0, 1)):
            t[0, 0, ol.stencil_index(o)] = -1.0 / (h ** 2)
        return t

    def boundary_value(self, fi, level, x, y):
        # u on boundary = cos(PI x) - sin(2 PI y)                     (exa2:5)
        return np.cos(math.pi * x) - np.sin(2.0 * math.pi * y)

    def rhs_value(self, fi, level, x, y):
        # RHS_u = PI^2 cos(PI x) - 4 PI^2 sin(2 PI y)                 (exa2:7)
        return math.pi ** 2 * np.cos(math.pi * x) - 4.0 * math.pi ** 2 * np.sin(2.0 * math.pi * y)


class Poisson3D(Problem):
    """example_problems/Poisson/3D_FD_Poisson_fromL2.exa2:2-23, .exa3:1-14, .knowledge:1-4."""

    def __init__(self, min_level: int = 2, max_level: int = 6):
        super().__init__(name="3D_FD_Poisson_fromL2", dim=3, fields=("u",), rhs_names=("RHS_u",),
                         equation_names=("solEq",), min_level=min_level, max_level=max_level,
                         settings=SolverSettings(damping=1.25))

    def operator(self, level):
        h = self.spacing(level)
        t = _table(1)
        t[0, 0, ol.stencil_index((0, 0, 0))] = 2.0 / (h ** 2) + 2.0 / (h ** 2) + 2.0 / (h ** 2)   # exa2:12
        for o in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
            t[0, 0, ol.stencil_index(o)] = -1.0 / (h ** 2)                                        # exa2:13-18
        return t

    def boundary_value(self, fi, level, x, y, z):
        # u@finest on boundary = x^2 - 0.5 y^2 - 0.5 z^2 ; all coarser levels 0   (exa2:6-7)
        if level != self.max_level:
            return None
        return x * x - 0.5 * y * y - 0.5 * z * z

    def rhs_value(self, fi, level, x, y, z):
        return None                                                    # RHS_u = 0 (exa2:9)


class LinearElasticity2D(Problem):
    """example_problems/LinearElasticity/2D_FD_LinearElasticity_fromL2.exa2:2-55, .exa3:2-17.

    uEq: (lambda+mu)(dxx u + dxy v) + lambda Laplace u = RHS_u
    vEq: (lambda+mu)(dxy u + dyy v) + lambda Laplace v = RHS_v      (exa2:45-50)
    with the negative-definite Laplace of exa2:23-29."""

    def __init__(self, min_level: int = 4, max_level: int = 8, lam: float = 195.0, mu: float = 130.0):
        super().__init__(name="2D_FD_LinearElasticity_fromL2", dim=2, fields=("u", "v"),
                         rhs_names=("RHS_u", "RHS_v"), equation_names=("uEq", "vEq"),
                         min_level=min_level, max_level=max_level, settings=SolverSettings(damping=1.25),
                         parameters={"lambda": lam, "mu": mu})

    def operator(self, level):
        h = self.spacing(level)
        lam, mu = self.parameters["lambda"], self.parameters["mu"]
        t = _table(2)

        def add(i, j, off, val):
            t[i, j, ol.stencil_index(off)] += val

        dxx = {(0, 0): -2.0 / (h ** 2), (-1, 0): 1.0 / (h ** 2), (1, 0): 1.0 / (h ** 2)}            # exa2:11-15
        dyy = {(0, 0): -2.0 / (h ** 2), (0, -1): 1.0 / (h ** 2), (0, 1): 1.0 / (h ** 2)}            # exa2:17-21
        lap = {(0, 0): -2.0 / (h ** 2) - 2.0 / (h ** 2), (-1, 0): 1.0 / (h ** 2), (1, 0): 1.0 / (h ** 2),
               (0, -1): 1.0 / (h ** 2), (0, 1): 1.0 / (h ** 2)}                                     # exa2:23-29
        dxy = {(-1, 1): -1.0 / (4 * h * h), (1, 1): 1.0 / (4 * h * h),
               (-1, -1): 1.0 / (4 * h * h), (1, -1): -1.0 / (4 * h * h)}                            # exa2:31-36
        for o, v in dxx.items():
            add(0, 0, o, (lam + mu) * v)
        for o, v in lap.items():
            add(0, 0, o, lam * v)
            add(1, 1, o, lam * v)
        for o, v in dyy.items():
            add(1, 1, o, (lam + mu) * v)
        for o, v in dxy.items():
            add(0, 1, o, (lam + mu) * v)
            add(1, 0, o, (lam + mu) * v)
        return t

    def boundary_value(self, fi, level, x, y):
        if fi == 0:
            return None                                                # u on boundary = 0 (exa2:5)
        # v on boundary = 0.4 sin(PI x)(1 - x) x y                     (exa2:7)
        return 4e-1 * np.sin(math.pi * x) * (1.0 - x) * x * y + 0.0 * y

    def rhs_value(self, fi, level, x, y):
        return None                                                    # exa2:8-9


class FAS2D(Problem):
    """example_problems/FAS_2D_Basic/FAS_2D_Basic_template.exa4: -Lap u + gamma u e^u = f,
    gamma = 20 (:34), u* = (x^2 - x^3) sin(3 PI y) (:54-56), f :48-53, u0 = 0, homogeneous Dirichlet;
    Solve loop tol 1e-10 / 300 iterations (:146); CGS = 200 damped Newton-Jacobi sweeps (:58-73)."""

    def __init__(self, min_level: int = 6, max_level: int = 10, gamma: float = 20.0):
        super().__init__(name="FAS_2D_Basic", dim=2, fields=("u",), rhs_names=("RHS_u",),
                         equation_names=("solEq",), min_level=min_level, max_level=max_level,
                         settings=SolverSettings(tol=1e-10, max_iters=300, num_pre=2, num_post=2, damping=0.8,
                                                 red_black=False, cgs_max_iters=200, cgs_tol=0.0),
                         kind=ol.PROBLEM_FAS, gamma=gamma)

    def operator(self, level):
        h = self.spacing(level)
        t = _table(1)
        t[0, 0, ol.stencil_index((0, 0))] = (2.0 / (h * h) + 2.0 / (h * h))          # template.exa4:20
        for o in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            t[0, 0, ol.stencil_index(o)] = (-1.0 / (h * h))                          # :21-24
        return t

    def exact_solution(self, x, y):
        return (x ** 2 - x ** 3) * np.sin(3.0 * math.pi * y)

    def rhs_value(self, fi, level, x, y):
        sol = self.exact_solution(x, y)
        return ((9.0 * math.pi ** 2 + self.gamma * np.exp(sol)) * (x ** 2 - x ** 3) + 6.0 * x - 2.0) \
            * np.sin(3.0 * math.pi * y)


class Helmholtz2D(Problem):
    """example_problems/Helmholtz/2D_FD_Helmholtz_fromL3.exa3: A = -Lap_h - k^2 (:55-61),
    preconditioner operator M = -Lap_h - k^2 * shift, shift = 1 + 0.5i (:63-69, :80-84), k = 80,
    RHS = product of hat functions at (0.5, 0.5) (:24); Robin x-boundaries (exa4:25-145);
    outer preconditioned BiCGStab, stop 1e-7 or 10000 iterations (:144-200); coarsest-level
    BiCGStab <= 1000 iterations to 1e-6 (:396-433)."""

    def __init__(self, min_level: int = 3, max_level: int = 7, k: float = 80.0, shift: complex = 1.0 + 0.5j,
                 omega_relax: float = 0.6):
        super().__init__(name="2D_FD_Helmholtz_fromL3", dim=2, fields=("u",), rhs_names=("f",),
                         equation_names=("PrecEq",), min_level=min_level, max_level=max_level,
                         settings=SolverSettings(tol=1e-7, max_iters=10000, num_pre=2, num_post=1,
                                                 damping=omega_relax, red_black=True, cgs_max_iters=1000,
                                                 cgs_tol=1e-6),
                         kind=ol.PROBLEM_HELMHOLTZ, complex_valued=True, wave_number=complex(k),
                         parameters={"k": k, "shift_re": shift.real, "shift_im": shift.imag})

    @property
    def shift(self) -> complex:
        return complex(self.parameters["shift_re"], self.parameters["shift_im"])

    def _lap(self, level, diag_extra):
        h = self.spacing(level)
        t = _table(1, True)
        t[0, 0, ol.stencil_index((0, 0))] = 2.0 / (h ** 2) + 2.0 / (h ** 2) - diag_extra
        for o in ((-1, 0), (1, 0), (0, -1), (0, 1)):
            t[0, 0, ol.stencil_index(o)] = -1.0 / (h ** 2)
        return t

    def operator(self, level):
        """Preconditioner operator M (the evolved cycle works on PrecEq: M u = f, exa3:71-73)."""
        k = self.parameters["k"]
        return self._lap(level, k ** 2 * self.shift)

    def outer_operator(self, level):
        """A of the outer Krylov iteration (exa3:55-61)."""
        k = self.parameters["k"]
        return self._lap(level, k ** 2 + 0j)

    def rhs_value(self, fi, level, x, y):
        h = self.spacing(self.max_level)
        fx = np.maximum(0.0, -(np.abs(x - 0.5) - h) / h ** 2)
        fy = np.maximum(0.0, -(np.abs(y - 0.5) - h) / h ** 2)
        return (fx * fy).astype(np.complex128)


PROBLEMS = {
    "poisson2d": Poisson2D,
    "poisson3d": Poisson3D,
    "elasticity2d": LinearElasticity2D,
    "fas2d": FAS2D,
    "helmholtz2d": Helmholtz2D,
}


def make_problem(name: str, **kw) -> Problem:
    return PROBLEMS[name](**kw)
