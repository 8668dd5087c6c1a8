"""From the residual history of a solve to the reference's fitness tuple.

The reference never sees residuals: it regex-parses the stdout of the generated binary
(reference: evostencils/code_generation/exastencils.py:540-584 ``parse_output``; FAS:
exastencils_FAS.py:370-394).  The generated program prints with the C++ ``cout`` default precision
of 6 significant digits, so what the reference averages are *rounded* per-iteration factors --
reproducing its numbers bit for bit needs the same rounding (the tutorial
known-answer test only matches with it).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple


def _cout(x: float, digits: int = 6) -> float:
    """Value of ``x`` after a round trip through ``std::cout << x`` (``%.{digits}g``)."""
    if math.isnan(x) or math.isinf(x):
        return x
    return float(f"%.{digits}g" % x)


def convergence_factor(residuals: Sequence[float], infinity: float = 1e100,
                       print_digits: Optional[int] = 6) -> float:
    """Geometric mean of the printed per-iteration factors rho_k = res_k / res_{k-1}
    (exastencils.py:547-554, :569-576).  A non-finite rho counts as sqrt(infinity) (:543, :550-551)
    but does not make the count non-zero (:554); no finite rho at all -> ``infinity`` (:574-576)."""
    rho_inf = math.sqrt(infinity)
    factors = []
    count = 0
    for k in range(1, len(residuals)):
        prev, cur = residuals[k - 1], residuals[k]
        try:
            rho = cur / prev
        except ZeroDivisionError:
            rho = math.nan if cur == 0 or math.isnan(cur) else math.copysign(math.inf, cur)
        if print_digits is not None:
            rho = _cout(rho, print_digits)
        if math.isinf(rho) or math.isnan(rho):
            factors.append(rho_inf)
        else:
            factors.append(rho)
            count += 1
    if count == 0:
        return infinity
    exponent = 1.0 / len(factors)
    cf = 1.0
    for rho in factors:
        cf *= math.pow(rho, exponent)
    return cf


def fitness_from_history(residuals: Sequence[float], time_ms: float, max_iters: int, infinity: float = 1e100,
                         solver_iteration_limit: Optional[int] = None,
                         print_digits: Optional[int] = 6) -> Tuple[float, float, float]:
    """(time_to_solution_ms, convergence_factor, number_of_iterations) as ``parse_output`` returns it.

    ``residuals`` = res_0 .. res_n of the executed iterations.  If the residual became non-finite
    the generated binary would still run to ``max_iters`` printing non-finite factors; the history
    is padded accordingly so that the geometric mean has the reference's exponent."""
    res = list(residuals)
    n = len(res) - 1
    if n >= 1 and not math.isfinite(res[-1]) and n < max_iters:
        res.extend([math.nan] * (max_iters - n))
        n = max_iters
    cf = convergence_factor(res, infinity, print_digits)
    iters: float = n
    if iters == 0:                                       # exastencils.py:580-581
        iters = infinity
    if solver_iteration_limit is not None and iters >= solver_iteration_limit:   # :582-583
        iters = infinity
    return float(time_ms), cf, iters


def fas_fitness(residuals: Sequence[float], time_ms: float, infinity: float = 1e100) -> Tuple[float, float, float]:
    """FAS variant (exastencils_FAS.py:370-394): c = (res_final / res_initial)^(1/n) from the values
    the template prints with 4 significant digits (FAS_2D_Basic_template.exa4:117-181; fewer digits
    below 1e-9, 'EFFECTIVELY ZERO' -- i.e. no print, previous value kept -- below 1e-12)."""
    def printed(x, previous):
        if x <= 1e-12:
            return previous
        if x <= 1e-11:
            return _cout(x, 1)
        if x <= 1e-10:
            return _cout(x, 2)
        if x <= 1e-9:
            return _cout(x, 3)
        return _cout(x, 4)

    n = len(residuals) - 1
    if n <= 0:
        return infinity, infinity, infinity
    res_initial = printed(residuals[0], 0.0)
    res_final = 0.0
    for r in residuals[1:]:
        if math.isnan(r):
            return infinity, infinity, infinity
        res_final = printed(r, res_final)
    try:
        c = (res_final / res_initial) ** (1.0 / n)
    except (ZeroDivisionError, OverflowError):
        return infinity, infinity, infinity
    if isinstance(c, complex) or math.isinf(c) or math.isnan(c):
        return infinity, infinity, infinity
    return float(time_ms), c, n


def helmholtz_fitness(residuals: Sequence[float], time_ms: float, max_iters: int, infinity: float = 1e100,
                      solver_iteration_limit: Optional[int] = None, tol: float = 1e-7) -> Tuple[float, float, float]:
    """What ``parse_output`` (exastencils.py:540-584) makes of the Helmholtz solver's prints
    (example_problems/Helmholtz/2D_FD_Helmholtz_fromL3.exa3:192-199): ONE line
    "Residual after <curStep> iterations is ... --- convergence factor is <|res|/|res0|>" when the loop ends,
    curStep being the 0-based index of the last iteration, so the "convergence factor" is the total
    reduction and an immediate convergence (curStep = 0) counts as infinity (:580-581); hitting the cap
    prints "Maximum number of solver iterations" first -> iterations = infinity (:555-557)."""
    n = len(residuals) - 1
    if n <= 0:
        return float(time_ms), infinity, infinity
    res0, res = residuals[0], residuals[-1]
    rho = _cout(res / res0) if res0 != 0 else math.nan
    cf = math.sqrt(infinity) if (math.isinf(rho) or math.isnan(rho)) else rho
    if math.isinf(rho) or math.isnan(rho):
        cf = infinity            # count == 0 (:554, :574-576)
    converged = res < tol * res0
    iters: float = n - 1 if converged else infinity
    if iters == 0:
        iters = infinity
    if solver_iteration_limit is not None and iters >= solver_iteration_limit:
        iters = infinity
    return float(time_ms), cf, iters
