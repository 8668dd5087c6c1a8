"""Hand-written op lists of the classical cycles the reference uses as baselines.

* :func:`default_solver_cycle` -- what ExaStencils' ``generate solver`` block produces and
  ``scripts/evaluate_reference_solver.py`` times (reference: scripts/evaluate_reference_solver.py:5-48,
  example_problems/Poisson/2D_FD_Poisson_fromL2.exa3:2-15): V(numPre, numPost) with damped
  red-black Gauss-Seidel and a CG coarse-grid solve.
* :func:`v_cycle` / :func:`w_cycle` -- general (nu1, nu2) cycles with pointwise Jacobi or RB-GS.

The statement order is the one the reference's emitter produces for the equivalent grammar tree
(SURVEY.md Appendix F.4): smooth, residual, restrict, zero the coarse error, recurse, prolongate and
correct, smooth.
"""
from __future__ import annotations

from typing import List, Optional

from . import oplist as ol
from .problems import Problem


def pointwise_unknowns(problem: Problem, collective: bool = True):
    zero = (0,) * problem.dim
    if collective:
        return [tuple((f, zero) for f in range(problem.n_fields))]
    return [((f, zero),) for f in range(problem.n_fields)]


def smoother_ops(problem: Problem, level: int, omega: float, red_black: bool, sweeps: int,
                 collective: bool = True) -> List[ol.Op]:
    ops = []
    mode = ol.MODE_REDBLACK if red_black else ol.MODE_JACOBI
    for _ in range(sweeps):
        for unk in pointwise_unknowns(problem, collective):
            ops.append(ol.Op(ol.OP_SMOOTH, level, mode=mode, omega=omega, unknowns=tuple(unk)))
    return ops


def _cycle(problem: Problem, level: int, pre: int, post: int, omega: float, red_black: bool, gamma: int,
           cgc_weight: float, collective: bool) -> List[ol.Op]:
    s = problem.settings
    if level == problem.min_level:
        return [ol.Op(ol.OP_COARSE_SOLVE, level, count=s.cgs_max_iters, tol=s.cgs_tol)]
    ops = smoother_ops(problem, level, omega, red_black, pre, collective)
    ops.append(ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES))
    ops.append(ol.Op(ol.OP_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES))
    if level - 1 > problem.min_level:
        ops.append(ol.Op(ol.OP_ZERO, level - 1, dst=ol.BUF_SOL))
    for g in range(gamma if level - 1 > problem.min_level else 1):
        ops.extend(_cycle(problem, level - 1, pre, post, omega, red_black, gamma, cgc_weight, collective))
    ops.append(ol.Op(ol.OP_PROLONG_ADD, level, src=ol.BUF_SOL, omega=cgc_weight))
    ops.extend(smoother_ops(problem, level, omega, red_black, post, collective))
    return ops


def build_program(problem: Problem, ops: List[ol.Op]) -> ol.Program:
    prog = ol.Program(dim=problem.dim, n_fields=problem.n_fields, min_level=problem.min_level,
                      max_level=problem.max_level, ops=list(ops))
    for lvl in range(problem.min_level, problem.max_level + 1):
        prog.operators[lvl] = problem.operator(lvl)
    prog.restrict_w = problem.restrict_weights()
    prog.prolong_w = problem.prolong_weights()
    return prog


def v_cycle(problem: Problem, pre: int = 2, post: int = 2, omega: float = 1.0, red_black: bool = True,
            cgc_weight: float = 1.0, collective: bool = True) -> ol.Program:
    return build_program(problem, _cycle(problem, problem.max_level, pre, post, omega, red_black, 1, cgc_weight,
                                         collective))


def w_cycle(problem: Problem, pre: int = 2, post: int = 2, omega: float = 1.0, red_black: bool = True,
            cgc_weight: float = 1.0, collective: bool = True) -> ol.Program:
    return build_program(problem, _cycle(problem, problem.max_level, pre, post, omega, red_black, 2, cgc_weight,
                                         collective))


def default_solver_cycle(problem: Problem) -> ol.Program:
    """The cycle of the problem's own `generate solver` block (config C1a of SURVEY.md 8d)."""
    s = problem.settings
    return v_cycle(problem, s.num_pre, s.num_post, s.damping, s.red_black)


# ------------------------------------------------------------------------------------------------
def _fas_cycle(problem: Problem, level: int, pre: int, post: int, omega: float, newton_steps: int, red_black: bool,
               cgc_weight: float) -> List[ol.Op]:
    """FAS V-cycle in the statement order of the reference's FAS emitter (exastencils_FAS.py:99-319;
    golden text example_problems/FAS_2D_Basic/FAS_2D_Basic.exa4:213-269)."""
    s = problem.settings
    zero = (0,) * problem.dim
    kind = ol.KIND_FAS_NEWTON if newton_steps > 0 else ol.KIND_FAS_PICARD
    mode = ol.MODE_REDBLACK if red_black else ol.MODE_JACOBI

    def smooth(n):
        return [ol.Op(ol.OP_SMOOTH, level, mode=mode, kind=kind, count=max(1, newton_steps), omega=omega,
                      unknowns=((0, zero),)) for _ in range(n)]

    if level == problem.min_level:
        return [ol.Op(ol.OP_COARSE_SOLVE, level, count=s.cgs_max_iters, omega=s.damping)]
    ops = smooth(pre)
    ops.append(ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES))
    ops.append(ol.Op(ol.OP_FAS_RESTRICT_SOL, level, dst=ol.BUF_APX, src=ol.BUF_SOL))
    ops.append(ol.Op(ol.OP_FAS_COARSE_RHS, level, dst=ol.BUF_RHS, src=ol.BUF_RES))
    ops.extend(_fas_cycle(problem, level - 1, pre, post, omega, newton_steps, red_black, cgc_weight))
    ops.append(ol.Op(ol.OP_FAS_SUB_APX, level - 1, dst=ol.BUF_SOL, src=ol.BUF_APX))
    ops.append(ol.Op(ol.OP_PROLONG_ADD, level, src=ol.BUF_SOL, omega=cgc_weight))
    ops.extend(smooth(post))
    return ops


def fas_v_cycle(problem: Problem, pre: int = 2, post: int = 2, omega: float = 0.8, newton_steps: int = 1,
                red_black: bool = False, cgc_weight: float = 1.0) -> ol.Program:
    """The template's own FAS cycle: V(2,2), damped Newton-Jacobi omega = 0.8 (template.exa4:65-73)."""
    return build_program(problem, _fas_cycle(problem, problem.max_level, pre, post, omega, newton_steps, red_black,
                                             cgc_weight))
