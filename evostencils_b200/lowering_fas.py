"""Lowering of FAS cycle trees (nonlinear full-approximation scheme).

Counterpart of ``ProgramGeneratorFAS.traverse_graph`` (reference:
evostencils/code_generation/exastencils_FAS.py:99-319), which walks the tree built by the FAS variant
of the grammar (grammar/multigrid.py:275-343 with ``FAS=True``) and appends layer-4 statements; here
the same walk appends ops.  The reference keeps per-level ``update_rhs`` flags (:38, :138-147, :308) so
that the coarse right-hand side is computed once when a coarse level is first used and again after a
prolongation from it; the same flags are kept here.

Deliberate deviation: the reference prints the restriction of the solution as
``Approximation@(l-1) = RestrictionNode@l * Solution@(l-1)`` (coarse field name, golden text
FAS_2D_Basic.exa4:222); the intended FAS restricts the *fine* solution, which is what
``OP_FAS_RESTRICT_SOL`` does.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import oplist as ol
from .lowering import LoweringError, operator_table, transfer_table, _is


def _tname(obj) -> str:
    return type(obj).__name__


class FASLowering:
    def __init__(self, min_level: int, max_level: int, dim: int, cgs_sweeps: int, cgs_omega: float):
        self.min_level, self.max_level, self.dim = min_level, max_level, dim
        self.cgs_sweeps, self.cgs_omega = cgs_sweeps, cgs_omega
        self.ops: List[ol.Op] = []
        self.operators: Dict[int, np.ndarray] = {}
        self.update_rhs = {l: True for l in range(min_level, max_level + 1)}
        self.relaxation_factors: List[float] = []
        self.partitioning: List[object] = []
        self.restrict_w = None
        self.prolong_w = None

    @staticmethod
    def _level(obj) -> int:
        g = obj.grid
        return int((g[0] if isinstance(g, (list, tuple)) else g).level)

    def _register(self, operator, level):
        if operator is None or not hasattr(operator, "entries"):
            return
        try:
            t = operator_table(operator, 1)
        except Exception:
            return
        if np.any(t != 0):
            self.operators.setdefault(level, t)

    # -- helpers named after the reference's local functions ---------------------------------------------
    def _update_rhs(self, rhs_obj, cur):
        if _tname(rhs_obj) != "RightHandSide":                       # :139
            if cur >= self.max_level:
                raise LoweringError("coarse right-hand side on the finest level")
            if self.update_rhs[cur]:                                   # :143
                self._rhs_expression(rhs_obj, cur)
                self.update_rhs[cur] = False

    def _rhs_expression(self, expr, cur):
        """RHS@cur = R * Residual@(cur+1) + (A + N)(R * approximation)   (:138-147 with the tree of
        grammar/multigrid.py:287-293).  Emission order = traversal order: residual statement of the fine
        level, FAS restriction of the solution, then the right-hand side loop."""
        if _tname(expr) != "Addition":
            raise LoweringError("FAS: unexpected coarse right-hand side expression")
        a, b = expr.operand1, expr.operand2
        # operand1 = R * Residual  -> generic multiplication branch: traverse(R), traverse(Residual)
        if _tname(a) != "Multiplication" or "Residual" not in _tname(a.operand2):
            raise LoweringError("FAS: expected R * residual")
        self._traverse(a.operand1)
        self._traverse(a.operand2)
        # operand2 = A_c * (R * approximation) -> Operator branch: traverse(operand) = updateFASApproximation
        if _tname(b) != "Multiplication":
            raise LoweringError("FAS: expected A * (R * u)")
        self._traverse(b.operand2)
        self._register(b.operand1, cur)
        self.ops.append(ol.Op(ol.OP_FAS_COARSE_RHS, cur + 1, dst=ol.BUF_RHS, src=ol.BUF_RES))

    def _smoothing(self, expression, cur):
        residual = expression.operand2
        self._update_rhs(residual.rhs, cur)
        omega = self.relaxation_factors.pop()
        part = self.partitioning.pop()
        pname = part.__name__ if isinstance(part, type) else _tname(part)
        red_black = pname == "RedBlack"
        operator = expression.operand1.operand
        newton_steps = 0
        if _tname(operator) == "Addition":                          # ElementwiseDiagonal + Jacobian (ir/smoother.py:45-46)
            newton_steps = int(operator.operand2.n_newton_steps)
        self._register(residual.operator, cur)
        zero = (0,) * self.dim
        self.ops.append(ol.Op(ol.OP_SMOOTH, cur, mode=ol.MODE_REDBLACK if red_black else ol.MODE_JACOBI,
                              kind=ol.KIND_FAS_NEWTON if newton_steps > 0 else ol.KIND_FAS_PICARD,
                              count=max(1, newton_steps), omega=float(omega), unknowns=((0, zero),)))

    # -- the walk ---------------------------------------------------------------------------------------
    def _traverse(self, expression):
        """Returns a tag describing what the reference's traverse_graph would return (a field / an
        operator / None)."""
        cur = self._level(expression)
        t = _tname(expression)
        if t == "Cycle":
            self.relaxation_factors.append(expression.relaxation_factor)
            self.partitioning.append(expression.partitioning)
            self._traverse(expression.approximation)
            correction = self._traverse(expression.correction)
            if correction is not None:                                # :106-117
                omega = self.relaxation_factors.pop()
                self.partitioning.pop()
                if correction != ("P*SOL", cur):
                    raise LoweringError("FAS: unsupported correction expression")
                self.ops.append(ol.Op(ol.OP_PROLONG_ADD, cur, dst=ol.BUF_SOL, src=ol.BUF_SOL, omega=float(omega)))
            return ("SOL", cur)
        if t == "Multiplication":
            op_type = _tname(expression.operand1)
            operand = expression.operand2
            if op_type == "CoarseGridSolver":                         # solve(), :185-194
                self._update_rhs(operand, cur)
                self._register(expression.operand1.operator, cur)
                self.ops.append(ol.Op(ol.OP_COARSE_SOLVE, cur, count=self.cgs_sweeps, omega=self.cgs_omega))
                return ("SOL", cur)
            if op_type == "Inverse":                                  # smoothing(), :196-252
                self._smoothing(expression, cur)
                return None
            if op_type == "Operator":
                self._traverse(operand)
                self._register(expression.operand1, cur)
                return ("A*x", cur)
            if op_type == "Restriction" and ("Approximation" in _tname(operand) or _tname(operand) == "Cycle"):
                # updateFASApproximation(), :121-136 (operand1 is traversed, operand2 is NOT)
                w = transfer_table(expression.operand1)
                self.restrict_w = w if self.restrict_w is None else self.restrict_w
                self.ops.append(ol.Op(ol.OP_FAS_RESTRICT_SOL, cur + 1, dst=ol.BUF_APX, src=ol.BUF_SOL))
                return ("APX", cur)
            a = self._traverse(expression.operand1)
            b = self._traverse(expression.operand2)
            if a == ("P", cur - 1) and b == ("SOL", cur - 1):
                return ("P*SOL", cur)
            return ("mul", cur)
        if t == "Addition":
            self._traverse(expression.operand1)
            self._traverse(expression.operand2)
            return ("add", cur)
        if t == "Subtraction":                                        # updateFASerror(), :173-183
            self._traverse(expression.operand1)
            self.ops.append(ol.Op(ol.OP_FAS_SUB_APX, cur, dst=ol.BUF_SOL, src=ol.BUF_APX))
            return ("SOL", cur)
        if "Residual" in t:                                           # :297-300
            self._update_rhs(expression.rhs, cur)
            self._register(expression.operator, cur)
            self.ops.append(ol.Op(ol.OP_RESIDUAL, cur, dst=ol.BUF_RES))
            return ("RES", cur)
        if "Approximation" in t:
            return ("SOL", cur)
        if t == "RightHandSide":
            return ("RHS", cur)
        if t == "Prolongation":
            self.update_rhs[cur - 1] = True                           # :308
            self.prolong_w = transfer_table(expression) if self.prolong_w is None else self.prolong_w
            return ("P", cur - 1)
        if t == "Restriction":
            self.restrict_w = transfer_table(expression) if self.restrict_w is None else self.restrict_w
            return ("R", cur + 1)
        if t == "Operator":
            return ("A", cur)
        raise LoweringError(f"FAS: unsupported node {t}")


def lower_fas_cycle(expression, min_level: int, max_level: int, dim: int, cgs_sweeps: int = 200, cgs_omega: float = 0.8,
                    default_restrict=None, default_prolong=None, operators=None) -> ol.Program:
    lo = FASLowering(min_level, max_level, dim, cgs_sweeps, cgs_omega)
    lo._traverse(expression)
    prog = ol.Program(dim=dim, n_fields=1, min_level=min_level, max_level=max_level, ops=lo.ops, operators=lo.operators)
    if operators:
        for l, t in operators.items():
            prog.operators.setdefault(l, t)
    prog.restrict_w = lo.restrict_w if lo.restrict_w is not None and np.any(lo.restrict_w) else default_restrict
    prog.prolong_w = lo.prolong_w if lo.prolong_w is not None and np.any(lo.prolong_w) else default_prolong
    return prog
