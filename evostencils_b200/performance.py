"""Roofline-aware runtime estimate of a cycle on a B200.

Counterpart of the reference's ``PerformanceEvaluator``
(evostencils/model_based_prediction/performance.py:6-148): same constructor arguments, ``estimate_runtime`` returning
seconds (the optimiser multiplies by 1e3, optimization/program.py:347, :381) and ``set_runtime_of_coarse_grid_solver``
(:826).  The reference walks the expression tree with a generic operations/words roofline for a CPU; here the tree is
lowered to the op list the GPU executes and every statement is charged

    max( algorithmic bytes / (efficiency x measured HBM bandwidth),  launch floor )

with the algorithmic bytes of SURVEY.md 8(d) and the efficiencies measured for the hand-written kernels
(profiles/README.md).  It is a surrogate for pre-screening, not a measurement: `generate_and_evaluate` stays the
fitness.

Two modes:

* analytic (default, needs no device): the roofline formula above with the efficiency table below;
* measured (``measured=True`` with a ``generator``): every distinct statement shape (kind, level, smoother mode, local
  system, repetitions) is timed ONCE on the device with CUDA events (``evo_cycle_profile_op``, on the realistic data one
  application of the cycle leaves behind) and cached; a cycle's estimate is the sum of its statements' measured costs.
  After the first few individuals of a run every shape is known and an estimate costs no device work at all -- the
  B200 counterpart of the reference's operation / word counting.
"""
from __future__ import annotations

from typing import Dict, Optional

from . import oplist as ol

# fraction of the measured HBM peak the kernel family reaches at large sizes (profiles/, round 2)
EFFICIENCY: Dict[str, float] = {
    "rbgs3d": 0.95, "jacobi3d": 0.87, "residual3d": 0.87, "residual_restrict3d": 0.95, "restrict": 0.40,
    "prolong3d": 0.98, "generic": 0.35,
    # 2-D streaming kernels (>= 513^2): one pass per (up to) two sweeps
    "sweep2d": 0.75, "residual_restrict2d": 0.70, "prolong2d": 0.90, "residual2d": 0.55,
}
LAUNCH_FLOOR_S = 3.0e-6          # a kernel node of the solver graph on a latency-bound level
ROWSEQ_STEP_S = 1.3e-6           # one row step of the order-dependent coloured sweep (one CTA)


def _is_star5(program: ol.Program, level: int) -> bool:
    """Scalar 2-D 5-point operator on this level (the register-streamed kernels of evo_kernels_warp2d.cuh apply)."""
    table = (program.operators or {}).get(level)
    if table is None:
        return False
    nz = {int(p) for p in range(ol.STENCIL_POINTS) if table[0][0][p] != 0}
    return nz == {10, 12, 13, 14, 16}


class B200PerformanceEvaluator:
    def __init__(self, peak_performance: float = 40e12, peak_bandwidth: float = 6554.6e9, bytes_per_word: int = 8,
                 runtime_coarse_grid_solver: float = 0.0, generator=None, problem=None, measured: bool = False):
        self._peak_performance = peak_performance
        self._peak_bandwidth = peak_bandwidth
        self._bytes_per_word = bytes_per_word
        self._runtime_coarse_grid_solver = runtime_coarse_grid_solver
        self.generator = generator       # a B200ProgramGenerator: lowers expression trees (not needed for op lists)
        self.problem = problem           # ... or a problem description: lowered with lowering.lower_cycle (no device)
        if measured and generator is None:
            raise RuntimeError("measured statement costs need a program generator (device)")
        self.measured = measured
        self._table: Dict[tuple, float] = {}     # statement shape -> measured seconds
        self.device_measurements = 0             # statements timed on the device so far

    peak_performance = property(lambda self: self._peak_performance)
    peak_bandwidth = property(lambda self: self._peak_bandwidth)
    bytes_per_word = property(lambda self: self._bytes_per_word)
    runtime_coarse_grid_solver = property(lambda self: self._runtime_coarse_grid_solver)

    def set_runtime_of_coarse_grid_solver(self, runtime_coarse_grid_solver: float):
        self._runtime_coarse_grid_solver = runtime_coarse_grid_solver

    # ------------------------------------------------------------------------------------------------------
    def _bytes(self, words_per_dof: float, dofs: float) -> float:
        return words_per_dof * self.bytes_per_word * dofs

    def op_cost(self, op: ol.Op, program: ol.Program) -> float:
        """Seconds for one statement."""
        dim, nf = program.dim, program.n_fields
        dofs = float(((1 << op.level) - 1) ** dim) * nf
        star3 = dim == 3 and nf == 1 and op.level >= 5
        star2 = dim == 2 and nf == 1 and op.level >= 9 and _is_star5(program, op.level)
        bw = self.peak_bandwidth
        c = op.code
        if c == ol.OP_SMOOTH:
            nu = max(1, len(op.unknowns or ()))
            sweeps = max(1, op.count)
            if op.kind != ol.KIND_LINEAR:                      # FAS: exp + division per node, fp64 pipe bound
                t = max(self._bytes(3, dofs) / (0.37 * bw), 100.0 * dofs * max(1, op.count) / self.peak_performance)
                return max(t, LAUNCH_FLOOR_S)
            if nf > 1 and op.mode == ol.MODE_REDBLACK and nu == nf:
                return ((1 << op.level) + 4 * sweeps) * ROWSEQ_STEP_S   # row-sequential pipeline
            if star2 and nu == 1 and op.mode in (ol.MODE_REDBLACK, ol.MODE_JACOBI):
                passes2 = (sweeps + 1) // 2                      # two consecutive sweeps per pass over HBM
                return max(passes2 * self._bytes(3, dofs) / (EFFICIENCY["sweep2d"] * bw), passes2 * LAUNCH_FLOOR_S)
            key = ("rbgs3d" if op.mode == ol.MODE_REDBLACK else "jacobi3d") if (star3 and nu == 1) else "generic"
            passes = 2 if (op.mode == ol.MODE_REDBLACK and key == "generic") else 1
            return sweeps * max(passes * self._bytes(3, dofs) / (EFFICIENCY[key] * bw), passes * LAUNCH_FLOOR_S)
        if c == ol.OP_RESIDUAL:
            return max(self._bytes(3, dofs) / (EFFICIENCY["residual3d" if star3 else ("residual2d" if star2 else "generic")] * bw), LAUNCH_FLOOR_S)
        if c == ol.OP_RESIDUAL_RESTRICT:
            w = 2 + 1.0 / 2 ** dim
            return max(self._bytes(w, dofs) / (EFFICIENCY["residual_restrict3d" if star3 else ("residual_restrict2d" if star2 else "generic")] * bw), LAUNCH_FLOOR_S)
        if c in (ol.OP_RESTRICT, ol.OP_FAS_RESTRICT_SOL, ol.OP_FAS_COARSE_RHS):
            return max(self._bytes(1 + 1.0 / 2 ** dim, dofs) / (EFFICIENCY["restrict"] * bw), LAUNCH_FLOOR_S)
        if c in (ol.OP_PROLONG_ADD, ol.OP_PROLONG_SET):
            w = 2 + 1.0 / 2 ** dim
            return max(self._bytes(w, dofs) / (EFFICIENCY["prolong3d" if star3 else ("prolong2d" if star2 else "generic")] * bw), LAUNCH_FLOOR_S)
        if c in (ol.OP_ZERO, ol.OP_COPY, ol.OP_FAS_SUB_APX, ol.OP_RICHARDSON):
            return max(self._bytes(2, dofs) / (0.8 * bw), LAUNCH_FLOOR_S)
        if c == ol.OP_COARSE_SOLVE:
            if self._runtime_coarse_grid_solver:
                return self._runtime_coarse_grid_solver
            n = (1 << op.level) - 1
            if op.kind != ol.KIND_LINEAR or program.operators is None:
                return max(1, op.count) * 1.3e-6
            return min(max(1, op.count), 3 * n) * 2.5e-6        # CG: ~3 n iterations of ~2.5 us in one CTA
        return LAUNCH_FLOOR_S

    # ---- measured mode -------------------------------------------------------------------------------------
    @staticmethod
    def shape_of(op: ol.Op) -> tuple:
        """What determines the cost of a statement (the relaxation factor and the buffers' roles do not)."""
        return (op.code, op.level, op.mode, op.kind, tuple(op.unknowns or ()), max(1, op.count) if op.code == ol.OP_SMOOTH else 0)

    def _measure_missing(self, program: ol.Program) -> None:
        missing = {}
        for op in program.ops:
            k = self.shape_of(op)
            if k not in self._table and k not in missing:
                missing[k] = op
        if not missing:
            return
        g = self.generator
        dev = g._device_problem(program.min_level, program.max_level)
        cyc = dev.build(program)
        try:
            cyc.apply(1)                      # realistic data on every level (the coarse solver's work depends on it)
            for k, op in missing.items():
                ms, _ = cyc.profile_op(op, repeat=5)
                self._table[k] = ms * 1e-3
                self.device_measurements += 1
        finally:
            cyc.close()

    def estimate_program(self, program: ol.Program) -> float:
        if self.measured:
            self._measure_missing(program)
            return sum(self._table[self.shape_of(op)] for op in program.ops)
        return sum(self.op_cost(op, program) for op in program.ops)

    def estimate_runtime(self, expression) -> float:
        """Seconds per application of the cycle (the reference's contract); accepts an expression tree (lowered through
        the generator) or an already lowered `oplist.Program`."""
        if isinstance(expression, ol.Program):
            return self.estimate_program(expression)
        cached = getattr(expression, "runtime", None)
        if cached is not None:
            return cached
        if self.generator is not None:
            g = self.generator
            program = g._finalise(g.lower(expression, g.min_level))
        elif self.problem is not None:
            from . import lowering
            pr = self.problem
            program = lowering.optimise(lowering.lower_cycle(
                expression, pr.min_level, pr.max_level, pr.n_fields, pr.dim, cgs_max_iters=pr.settings.cgs_max_iters,
                cgs_tol=pr.settings.cgs_tol, default_restrict=pr.restrict_weights(), default_prolong=pr.prolong_weights()))
        else:
            raise RuntimeError("B200PerformanceEvaluator needs a program generator or a problem to lower expression trees")
        runtime = self.estimate_program(program)
        try:
            expression.runtime = runtime
        except AttributeError:
            pass
        return runtime
