"""Op list ("program") of one lowered multigrid cycle and the ctypes mirror of
``include/evostencils_b200.h``.

One :class:`Op` is one statement the reference's emitter would print for the cycle function
(reference: evostencils/code_generation/exastencils.py:684-925, ``generate_multigrid``).  The
op list replaces the ExaSlang text + Java code generation + ``make`` of the reference
(exastencils.py:485-504): it is handed to the CUDA library through ``evo_cycle_build``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

ABI_VERSION = 1
MAX_FIELDS = 2
MAX_UNKNOWNS = 8
MAX_DIM = 3
MAX_LEVELS = 16
STENCIL_POINTS = 27

# evo_buffer
BUF_SOL, BUF_RHS, BUF_RES, BUF_COR, BUF_APX = 0, 1, 2, 3, 4
BUF_NEXT = 100   # evo_cycle_buffer only: the [next] slot of SOL
BUF_NAMES = {BUF_SOL: "SOL", BUF_RHS: "RHS", BUF_RES: "RES", BUF_COR: "COR", BUF_APX: "APX"}

# evo_opcode
OP_ZERO, OP_COPY, OP_RESIDUAL, OP_RICHARDSON, OP_SMOOTH = 1, 2, 3, 4, 5
OP_RESTRICT, OP_PROLONG_ADD, OP_PROLONG_SET, OP_COARSE_SOLVE = 6, 7, 8, 9
OP_FAS_RESTRICT_SOL, OP_FAS_COARSE_RHS, OP_FAS_SUB_APX = 10, 11, 12
OP_RESIDUAL_RESTRICT, OP_SMOOTH_FUSED = 32, 33
OP_NAMES = {
    OP_ZERO: "ZERO", OP_COPY: "COPY", OP_RESIDUAL: "RESIDUAL", OP_RICHARDSON: "RICHARDSON", OP_SMOOTH: "SMOOTH",
    OP_RESTRICT: "RESTRICT", OP_PROLONG_ADD: "PROLONG_ADD", OP_PROLONG_SET: "PROLONG_SET",
    OP_COARSE_SOLVE: "COARSE_SOLVE", OP_FAS_RESTRICT_SOL: "FAS_RESTRICT_SOL", OP_FAS_COARSE_RHS: "FAS_COARSE_RHS",
    OP_FAS_SUB_APX: "FAS_SUB_APX", OP_RESIDUAL_RESTRICT: "RESIDUAL_RESTRICT", OP_SMOOTH_FUSED: "SMOOTH_FUSED",
}

# evo_smooth_mode / evo_smooth_kind
MODE_JACOBI, MODE_REDBLACK, MODE_LEX = 0, 1, 2
KIND_LINEAR, KIND_FAS_PICARD, KIND_FAS_NEWTON = 0, 1, 2

# evo_problem_kind
PROBLEM_LINEAR, PROBLEM_FAS, PROBLEM_HELMHOLTZ = 0, 1, 2

SOLVE_NO_GRAPH = 1
SOLVE_KEEP_STATE = 2
SOLVE_SOLO_TIMING = 4


class CEvoOp(C.Structure):
    _fields_ = [
        ("code", C.c_int32), ("level", C.c_int32), ("dst", C.c_int32), ("src", C.c_int32),
        ("mode", C.c_int32), ("kind", C.c_int32), ("n_unknowns", C.c_int32), ("count", C.c_int32),
        ("unk_field", C.c_int32 * MAX_UNKNOWNS),
        ("unk_off", (C.c_int32 * MAX_DIM) * MAX_UNKNOWNS),
        ("omega", C.c_double), ("tol", C.c_double),
    ]


class CEvoLevelOperator(C.Structure):
    _fields_ = [
        ("level", C.c_int32), ("pad_", C.c_int32),
        ("coef", (((C.c_double * 2) * STENCIL_POINTS) * MAX_FIELDS) * MAX_FIELDS),
    ]


class CEvoProblemDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("dim", C.c_int32), ("n_fields", C.c_int32), ("scalar_words", C.c_int32),
        ("min_level", C.c_int32), ("max_level", C.c_int32), ("kind", C.c_int32), ("device", C.c_int32),
        ("gamma", C.c_double), ("k_re", C.c_double), ("k_im", C.c_double),
        ("restrict_w", C.c_double * STENCIL_POINTS), ("prolong_w", C.c_double * STENCIL_POINTS),
    ]


class CEvoSolveParams(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iters", C.c_int32), ("samples", C.c_int32), ("flags", C.c_int32),
                ("timeout_ms", C.c_int32)]


class CEvoSolveResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("time_ms", C.c_double),
                ("time_ms_min", C.c_double), ("initial_residual", C.c_double), ("final_residual", C.c_double),
                ("kernel_launches", C.c_int64)]


def stencil_index(offset: Sequence[int]) -> int:
    """Table index p = (dz+1)*9 + (dy+1)*3 + (dx+1) of a stencil offset (2-D: dz = 0)."""
    try:
        return _STENCIL_INDEX[tuple(offset)]
    except (KeyError, TypeError):
        pass
    o = tuple(int(v) for v in offset) + (0,) * (3 - len(offset))
    if any(abs(v) > 1 for v in o):
        raise ValueError(f"stencil offset {offset} outside the 3^d neighbourhood")
    return (o[2] + 1) * 9 + (o[1] + 1) * 3 + (o[0] + 1)


# every offset of the 3^d neighbourhoods (the lowering calls stencil_index ~400 times per individual)
_STENCIL_INDEX = {}
for _dz in (-1, 0, 1):
    for _dy in (-1, 0, 1):
        for _dx in (-1, 0, 1):
            _STENCIL_INDEX[(_dx, _dy, _dz)] = (_dz + 1) * 9 + (_dy + 1) * 3 + (_dx + 1)
            if _dz == 0:
                _STENCIL_INDEX[(_dx, _dy)] = 9 + (_dy + 1) * 3 + (_dx + 1)


def stencil_offset(p: int, dim: int) -> Tuple[int, ...]:
    o = (p % 3 - 1, (p // 3) % 3 - 1, p // 9 - 1)
    return o[:dim]


@dataclass
class Op:
    """One statement of the cycle function."""
    code: int
    level: int
    dst: int = BUF_SOL
    src: int = BUF_SOL
    mode: int = MODE_JACOBI
    kind: int = KIND_LINEAR
    count: int = 1
    omega: float = 1.0
    tol: float = 0.0
    unknowns: Tuple[Tuple[int, Tuple[int, ...]], ...] = ()  # ((field, offset), ...)

    def to_c(self) -> CEvoOp:
        c = CEvoOp()
        c.code, c.level, c.dst, c.src = self.code, self.level, self.dst, self.src
        c.mode, c.kind, c.count = self.mode, self.kind, self.count
        c.n_unknowns = len(self.unknowns)
        if len(self.unknowns) > MAX_UNKNOWNS:
            raise ValueError("local system larger than EVO_MAX_UNKNOWNS")
        for a, (fld, off) in enumerate(self.unknowns):
            c.unk_field[a] = fld
            for d in range(MAX_DIM):
                c.unk_off[a][d] = int(off[d]) if d < len(off) else 0
        c.omega, c.tol = float(self.omega), float(self.tol)
        return c

    def key(self):
        """Structural key used by the tests (weights rounded to the printed precision)."""
        return (OP_NAMES[self.code], self.level, BUF_NAMES[self.dst], BUF_NAMES[self.src], self.mode, self.kind,
                self.count, repr(float(self.omega)), self.unknowns)

    def to_json(self) -> dict:
        return {"code": OP_NAMES[self.code], "level": self.level, "dst": BUF_NAMES[self.dst],
                "src": BUF_NAMES[self.src], "mode": self.mode, "kind": self.kind, "count": self.count,
                "omega": float(self.omega), "tol": float(self.tol),
                "unknowns": [[f, list(o)] for f, o in self.unknowns]}

    @staticmethod
    def from_json(d: dict) -> "Op":
        codes = {v: k for k, v in OP_NAMES.items()}
        bufs = {v: k for k, v in BUF_NAMES.items()}
        return Op(code=codes[d["code"]], level=d["level"], dst=bufs[d["dst"]], src=bufs[d["src"]],
                  mode=d.get("mode", 0), kind=d.get("kind", 0), count=d.get("count", 1),
                  omega=d.get("omega", 1.0), tol=d.get("tol", 0.0),
                  unknowns=tuple((int(f), tuple(int(v) for v in o)) for f, o in d.get("unknowns", [])))


@dataclass
class Program:
    """A lowered individual: statements + the rediscretised operators they refer to.

    ``operators[level]`` is an array ``[n_fields, n_fields, 27]`` (complex128 when the problem
    is complex) of stencil coefficients, index :func:`stencil_index`."""
    dim: int
    n_fields: int
    min_level: int
    max_level: int
    ops: List[Op] = field(default_factory=list)
    operators: Dict[int, np.ndarray] = field(default_factory=dict)
    restrict_w: np.ndarray | None = None
    prolong_w: np.ndarray | None = None

    def c_ops(self):
        arr = (CEvoOp * max(1, len(self.ops)))()
        for i, op in enumerate(self.ops):
            arr[i] = op.to_c()
        return arr

    def c_operators(self):
        levels = sorted(self.operators)
        arr = (CEvoLevelOperator * max(1, len(levels)))()
        for t, lvl in enumerate(levels):
            arr[t].level = lvl
            coef = np.asarray(self.operators[lvl], dtype=np.complex128)
            for i in range(self.n_fields):
                for j in range(self.n_fields):
                    for p in range(STENCIL_POINTS):
                        arr[t].coef[i][j][p][0] = coef[i, j, p].real
                        arr[t].coef[i][j][p][1] = coef[i, j, p].imag
        return arr, len(levels)

    def structure(self):
        return [op.key() for op in self.ops]

    def to_json(self) -> dict:
        def enc(a):
            a = np.asarray(a)
            if np.iscomplexobj(a):
                return {"re": a.real.tolist(), "im": a.imag.tolist()}
            return {"re": a.tolist()}
        return {"dim": self.dim, "n_fields": self.n_fields, "min_level": self.min_level,
                "max_level": self.max_level, "ops": [o.to_json() for o in self.ops],
                "operators": {str(k): enc(v) for k, v in self.operators.items()},
                "restrict_w": None if self.restrict_w is None else np.asarray(self.restrict_w).tolist(),
                "prolong_w": None if self.prolong_w is None else np.asarray(self.prolong_w).tolist()}

    @staticmethod
    def from_json(d: dict) -> "Program":
        def dec(e):
            a = np.asarray(e["re"], dtype=np.float64)
            if "im" in e:
                a = a + 1j * np.asarray(e["im"], dtype=np.float64)
            return a
        p = Program(dim=d["dim"], n_fields=d["n_fields"], min_level=d["min_level"], max_level=d["max_level"])
        p.ops = [Op.from_json(o) for o in d["ops"]]
        p.operators = {int(k): dec(v) for k, v in d["operators"].items()}
        p.restrict_w = None if d.get("restrict_w") is None else np.asarray(d["restrict_w"], dtype=np.float64)
        p.prolong_w = None if d.get("prolong_w") is None else np.asarray(d["prolong_w"], dtype=np.float64)
        return p


def full_weighting(dim: int) -> np.ndarray:
    """'default restriction on Node with linear' = full weighting, 1/4^d * prod(2-|o|)
    (reference: example_problems/Helmholtz/2D_FD_Helmholtz_fromL3.exa3:75)."""
    w = np.zeros(STENCIL_POINTS)
    for p in range(STENCIL_POINTS):
        o = (p % 3 - 1, (p // 3) % 3 - 1, p // 9 - 1)
        if dim == 2 and o[2] != 0:
            continue
        v = 1.0
        for d in range(dim):
            v *= (2 - abs(o[d]))
        w[p] = v / (4.0 ** dim)
    return w


def linear_interpolation(dim: int) -> np.ndarray:
    """'default prolongation on Node with linear' = bi/tri-linear, prod(2-|o|)/2^d
    (reference: Helmholtz...exa3:76; weights (2-|i|)(2-|j|)/4 in SURVEY.md Appendix F)."""
    w = np.zeros(STENCIL_POINTS)
    for p in range(STENCIL_POINTS):
        o = (p % 3 - 1, (p // 3) % 3 - 1, p // 9 - 1)
        if dim == 2 and o[2] != 0:
            continue
        v = 1.0
        for d in range(dim):
            v *= (2 - abs(o[d])) / 2.0
        w[p] = v
    return w
