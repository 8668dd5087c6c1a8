"""ctypes binding of the CUDA library (``evostencils_b200/csrc`` -> ``libevostencils_b200.so``).

This is the only route from the Python host to the numerics: there is no CPU fallback.  Loading the
library without a CUDA device works (the CPU test-suite checks the exported symbols), creating a
:class:`DeviceProblem` without one raises ``RuntimeError`` -- the analogue of the reference raising
``RuntimeError("Compiler not found. Aborting.")`` when its tool chain is missing
(reference: evostencils/code_generation/exastencils.py:104-108).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import oplist as ol
from .problems import Problem

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libevostencils_b200.so")

# every symbol include/evostencils_b200.h declares
EXPORTED_SYMBOLS = (
    "evo_abi_version", "evo_last_error", "evo_device_count", "evo_device_name",
    "evo_problem_create", "evo_problem_destroy", "evo_problem_set_field",
    "evo_cycle_build", "evo_cycle_destroy", "evo_cycle_reset", "evo_cycle_apply",
    "evo_cycle_get_field", "evo_cycle_set_field", "evo_cycle_residual_norm", "evo_cycle_profile_op",
    "evo_cycle_solve", "evo_helmholtz_solve", "evo_batch_solve",
    "evo_problem_set_slab", "evo_problem_set_slab_ex", "evo_problem_slab_info", "evo_cycle_set_stream", "evo_cycle_exec_ops", "evo_cycle_buffer",
    "evo_cycle_residual_plane_sums", "evo_cycle_vecsum", "evo_cycle_vecsum_async", "evo_cycle_read_sum",
    "evo_cycle_swap_slots", "evo_cycle_exec_part", "evo_set_option", "evo_get_option",
    "evo_cycle_solve_begin", "evo_cycle_solve_end", "evo_cycle_solve_retime",
)

_lib = None


ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_OOM = -1, -2, -3, -4, -5


class BackendError(RuntimeError):
    """Failure of a C-ABI call; ``status`` is the library's status code (include/evostencils_b200.h)."""

    def __init__(self, message: str, status: int = ERR_INVALID):
        super().__init__(message)
        self.status = status

    @property
    def infrastructure(self) -> bool:
        """True for failures that say nothing about the individual (no device, CUDA error, out of memory, a
        statement the library does not implement): callers must not turn these into a fitness value."""
        return self.status in (ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_OOM)


def load_library(path: Optional[str] = None):
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise BackendError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(nvcc, sm_100a). There is no CPU fallback for the evaluate path.")
    lib = C.CDLL(path)
    lib.evo_last_error.restype = C.c_char_p
    lib.evo_abi_version.restype = C.c_int
    lib.evo_device_count.restype = C.c_int
    lib.evo_device_name.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
    lib.evo_problem_create.argtypes = [C.POINTER(ol.CEvoProblemDesc), C.POINTER(C.c_void_p)]
    lib.evo_problem_destroy.argtypes = [C.c_void_p]
    lib.evo_problem_set_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.evo_cycle_build.argtypes = [C.c_void_p, C.POINTER(ol.CEvoOp), C.c_int, C.POINTER(ol.CEvoLevelOperator),
                                    C.c_int, C.POINTER(C.c_void_p)]
    lib.evo_cycle_destroy.argtypes = [C.c_void_p]
    lib.evo_cycle_reset.argtypes = [C.c_void_p]
    lib.evo_cycle_apply.argtypes = [C.c_void_p, C.c_int]
    lib.evo_cycle_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.evo_cycle_set_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.evo_cycle_residual_norm.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.evo_cycle_profile_op.argtypes = [C.c_void_p, C.POINTER(ol.CEvoOp), C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_int64)]
    lib.evo_cycle_solve.argtypes = [C.c_void_p, C.POINTER(ol.CEvoSolveParams), C.POINTER(ol.CEvoSolveResult),
                                    C.POINTER(C.c_double)]
    lib.evo_helmholtz_solve.argtypes = [C.c_void_p, C.POINTER(ol.CEvoLevelOperator), C.POINTER(ol.CEvoSolveParams),
                                        C.POINTER(ol.CEvoSolveResult), C.POINTER(C.c_double)]
    lib.evo_batch_solve.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(ol.CEvoSolveParams),
                                    C.POINTER(ol.CEvoSolveResult), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.evo_problem_set_slab.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.evo_problem_set_slab_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.evo_problem_slab_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
    lib.evo_cycle_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.evo_cycle_exec_ops.argtypes = [C.c_void_p, C.POINTER(ol.CEvoOp), C.c_int, C.c_int, C.c_int]
    lib.evo_cycle_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.evo_cycle_residual_plane_sums.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.evo_cycle_vecsum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    lib.evo_cycle_vecsum_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.evo_cycle_read_sum.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.evo_cycle_swap_slots.argtypes = [C.c_void_p, C.c_int]
    lib.evo_cycle_exec_part.argtypes = [C.c_void_p, C.POINTER(ol.CEvoOp), C.c_int, C.c_int, C.c_int]
    lib.evo_cycle_solve_begin.argtypes = [C.c_void_p, C.POINTER(ol.CEvoSolveParams)]
    lib.evo_cycle_solve_end.argtypes = [C.c_void_p, C.POINTER(ol.CEvoSolveParams), C.POINTER(ol.CEvoSolveResult), C.POINTER(C.c_double)]
    lib.evo_cycle_solve_retime.argtypes = [C.c_void_p, C.POINTER(ol.CEvoSolveParams), C.POINTER(ol.CEvoSolveResult)]
    lib.evo_set_option.argtypes = [C.c_char_p, C.c_int]
    lib.evo_get_option.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    if lib.evo_abi_version() != ol.ABI_VERSION:
        raise BackendError("ABI version mismatch between the Python host and libevostencils_b200.so")
    if path == LIB_PATH:
        _lib = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.evo_last_error()
        raise BackendError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}", status=rc)


def set_option(name: str, value: int) -> None:
    """Tuning switch of the CUDA library (``EVO_RB_VARIANT`` ...); applies to solver graphs captured afterwards."""
    lib = load_library()
    _check(lib, lib.evo_set_option(name.encode(), int(value)), f"evo_set_option({name})")


def get_option(name: str) -> int:
    lib = load_library()
    v = C.c_int(0)
    _check(lib, lib.evo_get_option(name.encode(), C.byref(v)), f"evo_get_option({name})")
    return int(v.value)


def device_count() -> int:
    return int(load_library().evo_device_count())


def make_desc(problem: Problem, device: int = 0) -> ol.CEvoProblemDesc:
    d = ol.CEvoProblemDesc()
    d.abi_version = ol.ABI_VERSION
    d.dim, d.n_fields = problem.dim, problem.n_fields
    d.scalar_words = 2 if problem.complex_valued else 1
    d.min_level, d.max_level = problem.min_level, problem.max_level
    d.kind, d.device = problem.kind, device
    d.gamma = problem.gamma
    d.k_re, d.k_im = complex(problem.wave_number).real, complex(problem.wave_number).imag
    rw, pw = problem.restrict_weights(), problem.prolong_weights()
    for p in range(ol.STENCIL_POINTS):
        d.restrict_w[p] = rw[p]
        d.prolong_w[p] = pw[p]
    return d


def _as_doubles(a: np.ndarray, complex_valued: bool) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.complex128 if complex_valued else np.float64)
    return a.view(np.float64).reshape(-1)


class SolveOutcome:
    """Raw result of the outer solver loop (what the generated binary prints, unparsed)."""

    def __init__(self, res: ol.CEvoSolveResult, hist: np.ndarray):
        self.status = int(res.status)
        self.iterations = int(res.iterations)
        self.time_ms = float(res.time_ms)
        self.time_ms_min = float(res.time_ms_min)
        self.initial_residual = float(res.initial_residual)
        self.final_residual = float(res.final_residual)
        self.kernel_launches = int(res.kernel_launches)
        self.residuals = np.array(hist[: self.iterations + 1], dtype=np.float64)


class DeviceCycle:
    """One lowered individual resident on the GPU (ops, operators, working hierarchy, CUDA graph)."""

    def __init__(self, problem: "DeviceProblem", program: ol.Program):
        self.problem = problem
        self.program = program
        self._lib = problem._lib
        self._h = C.c_void_p()
        ops = program.c_ops()
        operators, n_operators = program.c_operators()
        _check(self._lib, self._lib.evo_cycle_build(problem._h, ops, len(program.ops), operators, n_operators,
                                                    C.byref(self._h)), "evo_cycle_build")

    def close(self):
        if self._h:
            self._lib.evo_cycle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _check(self._lib, self._lib.evo_cycle_reset(self._h), "evo_cycle_reset")

    def apply(self, repeat: int = 1):
        _check(self._lib, self._lib.evo_cycle_apply(self._h, repeat), "evo_cycle_apply")

    def get_field(self, level: int, buf: int, field: int = 0) -> np.ndarray:
        p = self.problem.problem
        n = p.nodes(level)
        out = np.empty((n,) * p.dim, dtype=p.dtype)
        flat = out.view(np.float64).reshape(-1)
        _check(self._lib, self._lib.evo_cycle_get_field(self._h, level, buf, field, flat.ctypes.data, flat.size),
               "evo_cycle_get_field")
        return out

    def set_field(self, level: int, buf: int, field: int, data: np.ndarray):
        flat = _as_doubles(data, self.problem.problem.complex_valued)
        _check(self._lib, self._lib.evo_cycle_set_field(self._h, level, buf, field, flat.ctypes.data, flat.size),
               "evo_cycle_set_field")

    def residual_norm(self) -> float:
        v = C.c_double()
        _check(self._lib, self._lib.evo_cycle_residual_norm(self._h, C.byref(v)), "evo_cycle_residual_norm")
        return float(v.value)

    def profile_op(self, op: ol.Op, repeat: int = 10):
        """(average ms per execution, kernel launches per execution) of one statement."""
        ms, n = C.c_double(), C.c_int64()
        cop = op.to_c()
        _check(self._lib, self._lib.evo_cycle_profile_op(self._h, C.byref(cop), repeat, C.byref(ms), C.byref(n)),
               "evo_cycle_profile_op")
        return float(ms.value), int(n.value)

    # -- host-orchestrated execution (evostencils_b200.domain) ------------------------------------
    def set_stream(self, cuda_stream: int):
        _check(self._lib, self._lib.evo_cycle_set_stream(self._h, C.c_void_p(cuda_stream)), "evo_cycle_set_stream")

    def exec_ops(self, c_ops, n: int, zc_lo: int = -1, zc_hi: int = -1):
        _check(self._lib, self._lib.evo_cycle_exec_ops(self._h, c_ops, n, zc_lo, zc_hi), "evo_cycle_exec_ops")

    def exec_part(self, c_op, z_lo: int, z_hi: int, no_swap: bool):
        _check(self._lib, self._lib.evo_cycle_exec_part(self._h, c_op, z_lo, z_hi, 1 if no_swap else 0), "evo_cycle_exec_part")

    def buffer_ptr(self, level: int, buf: int, field: int = 0) -> int:
        p = C.c_void_p()
        _check(self._lib, self._lib.evo_cycle_buffer(self._h, level, buf, field, C.byref(p)), "evo_cycle_buffer")
        return int(p.value)

    def residual_plane_sums(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_int()
        _check(self._lib, self._lib.evo_cycle_residual_plane_sums(self._h, C.byref(p), C.byref(n)),
               "evo_cycle_residual_plane_sums")
        return int(p.value), int(n.value)

    def vecsum(self, device_ptr: int, m: int) -> float:
        v = C.c_double()
        _check(self._lib, self._lib.evo_cycle_vecsum(self._h, C.c_void_p(device_ptr), m, C.byref(v)), "evo_cycle_vecsum")
        return float(v.value)

    def vecsum_async(self, device_ptr: int, m: int):
        _check(self._lib, self._lib.evo_cycle_vecsum_async(self._h, C.c_void_p(device_ptr), m), "evo_cycle_vecsum_async")

    def read_sum(self) -> float:
        v = C.c_double()
        _check(self._lib, self._lib.evo_cycle_read_sum(self._h, C.byref(v)), "evo_cycle_read_sum")
        return float(v.value)

    def swap_slots(self, level: int):
        _check(self._lib, self._lib.evo_cycle_swap_slots(self._h, level), "evo_cycle_swap_slots")

    def solve(self, tol: float, max_iters: int, samples: int = 1, flags: int = 0, timeout_ms: int = 0) -> SolveOutcome:
        prm = ol.CEvoSolveParams(tol, max_iters, samples, flags, int(timeout_ms))
        res = ol.CEvoSolveResult()
        hist = np.zeros(max_iters + 1, dtype=np.float64)
        _check(self._lib, self._lib.evo_cycle_solve(self._h, C.byref(prm), C.byref(res),
                                                    hist.ctypes.data_as(C.POINTER(C.c_double))), "evo_cycle_solve")
        return SolveOutcome(res, hist)


def _solve_begin(cycle: "DeviceCycle", tol: float, max_iters: int, timeout_ms: int = 0):
    """Enqueue one complete solve on the cycle's stream and return at once (pipeline of many individuals)."""
    cycle._prm = ol.CEvoSolveParams(tol, max_iters, 1, 0, int(timeout_ms))
    _check(cycle._lib, cycle._lib.evo_cycle_solve_begin(cycle._h, C.byref(cycle._prm)), "evo_cycle_solve_begin")


def _solve_end(cycle: "DeviceCycle") -> SolveOutcome:
    prm = cycle._prm
    cycle._res = ol.CEvoSolveResult()
    hist = np.zeros(prm.max_iters + 1, dtype=np.float64)
    _check(cycle._lib, cycle._lib.evo_cycle_solve_end(cycle._h, C.byref(prm), C.byref(cycle._res),
                                                      hist.ctypes.data_as(C.POINTER(C.c_double))), "evo_cycle_solve_end")
    cycle._hist = hist
    return SolveOutcome(cycle._res, hist)


def _solve_retime(cycle: "DeviceCycle") -> SolveOutcome:
    """Time the finished solve again with the device to itself (call when nothing else is in flight)."""
    _check(cycle._lib, cycle._lib.evo_cycle_solve_retime(cycle._h, C.byref(cycle._prm), C.byref(cycle._res)), "evo_cycle_solve_retime")
    return SolveOutcome(cycle._res, cycle._hist)


DeviceCycle.solve_begin = _solve_begin
DeviceCycle.solve_end = _solve_end
DeviceCycle.solve_retime = _solve_retime


def _helmholtz_solve(cycle: "DeviceCycle", tol: float, max_iters: int, samples: int = 1, timeout_ms: int = 0) -> SolveOutcome:
    prob = cycle.problem.problem
    outer = ol.Program(dim=2, n_fields=1, min_level=prob.max_level, max_level=prob.max_level,
                       operators={prob.max_level: prob.outer_operator(prob.max_level)})
    arr, _ = outer.c_operators()
    prm = ol.CEvoSolveParams(tol, max_iters, samples, 0, int(timeout_ms))
    res = ol.CEvoSolveResult()
    hist = np.zeros(max_iters + 1, dtype=np.float64)
    _check(cycle._lib, cycle._lib.evo_helmholtz_solve(cycle._h, arr, C.byref(prm), C.byref(res),
                                                      hist.ctypes.data_as(C.POINTER(C.c_double))), "evo_helmholtz_solve")
    return SolveOutcome(res, hist)


DeviceCycle.helmholtz_solve = _helmholtz_solve


class DeviceProblem:
    """Discretisation hierarchy + initial guess / rhs on one GPU (``evo_problem``)."""

    backend_name = "cuda"

    def __init__(self, problem: Problem, device: int = 0, lib=None, slab: Optional[Tuple[int, int, int]] = None):
        """slab = (rank, world, coarsest distributed level[, ghost planes per side = 2]): hold only this rank's z-slab of every level
        >= that level (domain decomposition of one grid; see evostencils_b200.domain)."""
        self._lib = lib or load_library()
        n = self._lib.evo_device_count()
        if n <= 0:
            raise BackendError("no CUDA device visible: the B200 backend has no CPU fallback "
                               f"({(self._lib.evo_last_error() or b'').decode()})")
        self.problem = problem
        self.device = device
        self._h = C.c_void_p()
        desc = make_desc(problem, device)
        _check(self._lib, self._lib.evo_problem_create(C.byref(desc), C.byref(self._h)), "evo_problem_create")
        self.slab = slab
        if slab is not None:
            vals = [int(v) for v in slab]
            if len(vals) == 3:
                vals.append(2)
            _check(self._lib, self._lib.evo_problem_set_slab_ex(self._h, *vals), "evo_problem_set_slab_ex")
        for fi in range(problem.n_fields):
            for buf, arr in ((ol.BUF_SOL, problem.initial_solution(fi)), (ol.BUF_RHS, problem.rhs(fi))):
                flat = _as_doubles(arr, problem.complex_valued)
                _check(self._lib, self._lib.evo_problem_set_field(self._h, problem.max_level, buf, fi,
                                                                  flat.ctypes.data, flat.size),
                       "evo_problem_set_field")

    def close(self):
        if self._h:
            self._lib.evo_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, program: ol.Program) -> DeviceCycle:
        return DeviceCycle(self, program)

    def slab_info(self, level: int) -> dict:
        info = (C.c_longlong * 8)()
        _check(self._lib, self._lib.evo_problem_slab_info(self._h, level, info), "evo_problem_slab_info")
        keys = ("zoff", "nz", "zlo", "zhi", "g0", "g1", "pitch", "plane")
        return {k: int(v) for k, v in zip(keys, info)}

    def batch_solve(self, cycles: Sequence[DeviceCycle], tol: float, max_iters: int, samples: int = 1,
                    flags: int = 0, solo_timing: bool = False, timeout_ms: int = 0) -> Tuple[List[SolveOutcome], float]:
        """All cycles solve concurrently (one stream each).  ``solo_timing``: ``time_ms`` of every converged member is
        re-measured with the GPU to itself (a few iterations, extrapolated) -- comparable with a single ``solve``."""
        if solo_timing:
            flags |= ol.SOLVE_SOLO_TIMING
        n = len(cycles)
        handles = (C.c_void_p * n)(*[c._h for c in cycles])
        prm = ol.CEvoSolveParams(tol, max_iters, samples, flags, int(timeout_ms))
        results = (ol.CEvoSolveResult * n)()
        hist = np.zeros((n, max_iters + 1), dtype=np.float64)
        ms = C.c_double()
        _check(self._lib, self._lib.evo_batch_solve(handles, n, C.byref(prm), results,
                                                    hist.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ms)),
               "evo_batch_solve")
        return [SolveOutcome(results[i], hist[i]) for i in range(n)], float(ms.value)
