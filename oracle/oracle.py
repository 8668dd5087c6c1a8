"""TEST INFRASTRUCTURE: ctypes wrapper of ``oracle/liboracle.so`` (the CPU restatement).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs import this module; the product package never does.  The interface mirrors
``evostencils_b200.backend`` (DeviceProblem / DeviceCycle) so that the parity tests run the same op
list through both and compare.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from evostencils_b200 import oplist as ol
from evostencils_b200.backend import SolveOutcome, make_desc, _as_doubles

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O2 -fopenmp -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "mg_ops.inc", "mg_krylov.inc", "mg_fas.inc", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "evostencils_b200.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.POINTER(ol.CEvoProblemDesc)]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_set_operators.argtypes = [C.c_void_p, C.POINTER(ol.CEvoLevelOperator), C.c_int]
        lib.orc_set_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        lib.orc_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        lib.orc_reset.argtypes = [C.c_void_p]
        lib.orc_run_ops.argtypes = [C.c_void_p, C.POINTER(ol.CEvoOp), C.c_int, C.c_int]
        lib.orc_residual_norm.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.orc_solve.argtypes = [C.c_void_p, C.POINTER(ol.CEvoOp), C.c_int, C.POINTER(ol.CEvoSolveParams),
                                  C.POINTER(ol.CEvoSolveResult), C.POINTER(C.c_double)]
        lib.orc_helmholtz_solve.argtypes = [C.c_void_p, C.POINTER(ol.CEvoOp), C.c_int, C.POINTER(ol.CEvoLevelOperator),
                                            C.POINTER(ol.CEvoSolveParams), C.POINTER(ol.CEvoSolveResult),
                                            C.POINTER(C.c_double)]
        lib.orc_cg_iterations.restype = C.c_long
        lib.orc_cg_iterations.argtypes = [C.c_void_p]
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(load().orc_num_threads())


def set_num_threads(n: int):
    load().orc_set_num_threads(int(n))


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle: {what} failed with status {rc}")


class OracleCycle:
    def __init__(self, problem: "OracleProblem", program: ol.Program):
        self.problem = problem
        self.program = program
        self._lib = problem._lib
        self._ops = program.c_ops()
        self._n_ops = len(program.ops)
        operators, n = program.c_operators()
        _check(self._lib.orc_set_operators(problem._h, operators, n), "orc_set_operators")

    def _bind(self):
        # one hierarchy per OracleProblem: (re)install this cycle's operators before use
        if self.problem._bound is not self:
            operators, n = self.program.c_operators()
            _check(self._lib.orc_set_operators(self.problem._h, operators, n), "orc_set_operators")
            self.problem._bound = self

    def close(self):
        pass

    def reset(self):
        _check(self._lib.orc_reset(self.problem._h), "orc_reset")

    def apply(self, repeat: int = 1):
        self._bind()
        _check(self._lib.orc_run_ops(self.problem._h, self._ops, self._n_ops, repeat), "orc_run_ops")

    def get_field(self, level: int, buf: int, field: int = 0) -> np.ndarray:
        p = self.problem.problem
        n = p.nodes(level)
        out = np.empty((n,) * p.dim, dtype=p.dtype)
        flat = out.view(np.float64).reshape(-1)
        _check(self._lib.orc_get_field(self.problem._h, level, buf, field, flat.ctypes.data, flat.size), "orc_get_field")
        return out

    def set_field(self, level: int, buf: int, field: int, data: np.ndarray):
        flat = _as_doubles(data, self.problem.problem.complex_valued)
        _check(self._lib.orc_set_field(self.problem._h, level, buf, field, flat.ctypes.data, flat.size), "orc_set_field")

    def residual_norm(self) -> float:
        self._bind()
        v = C.c_double()
        _check(self._lib.orc_residual_norm(self.problem._h, C.byref(v)), "orc_residual_norm")
        return float(v.value)

    def solve(self, tol: float, max_iters: int, samples: int = 1, flags: int = 0, timeout_ms: int = 0) -> SolveOutcome:
        self._bind()
        prm = ol.CEvoSolveParams(tol, max_iters, samples, flags, 0)
        res = ol.CEvoSolveResult()
        hist = np.zeros(max_iters + 1, dtype=np.float64)
        _check(self._lib.orc_solve(self.problem._h, self._ops, self._n_ops, C.byref(prm), C.byref(res),
                                   hist.ctypes.data_as(C.POINTER(C.c_double))), "orc_solve")
        return SolveOutcome(res, hist)

    def helmholtz_solve(self, tol: float, max_iters: int, samples: int = 1, timeout_ms: int = 0) -> SolveOutcome:
        """Outer preconditioned BiCGStab with this cycle as the preconditioner (Helmholtz problems)."""
        self._bind()
        prob = self.problem.problem
        outer = ol.Program(dim=2, n_fields=1, min_level=prob.max_level, max_level=prob.max_level,
                           operators={prob.max_level: prob.outer_operator(prob.max_level)})
        arr, _ = outer.c_operators()
        prm = ol.CEvoSolveParams(tol, max_iters, samples, 0, 0)
        res = ol.CEvoSolveResult()
        hist = np.zeros(max_iters + 1, dtype=np.float64)
        _check(self._lib.orc_helmholtz_solve(self.problem._h, self._ops, self._n_ops, arr, C.byref(prm), C.byref(res),
                                             hist.ctypes.data_as(C.POINTER(C.c_double))), "orc_helmholtz_solve")
        return SolveOutcome(res, hist)

    def cg_iterations(self) -> int:
        return int(self._lib.orc_cg_iterations(self.problem._h))


class OracleProblem:
    backend_name = "oracle"

    def __init__(self, problem, device: int = 0):
        self._lib = load()
        self.problem = problem
        self._bound = None
        desc = make_desc(problem, 0)
        self._h = self._lib.orc_create(C.byref(desc))
        if not self._h:
            raise RuntimeError("oracle: invalid problem descriptor")
        for fi in range(problem.n_fields):
            for buf, arr in ((ol.BUF_SOL, problem.initial_solution(fi)), (ol.BUF_RHS, problem.rhs(fi))):
                flat = _as_doubles(arr, problem.complex_valued)
                _check(self._lib.orc_set_field(self._h, problem.max_level, buf, fi, flat.ctypes.data, flat.size),
                       "orc_set_field")

    def close(self):
        if self._h:
            self._lib.orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, program: ol.Program) -> OracleCycle:
        return OracleCycle(self, program)
