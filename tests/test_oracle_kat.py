"""Pins the CPU oracle against the reference's known-answer fixture (CPU only)."""
import math

import numpy as np

from evostencils_b200 import cycles, fitness, problems
from tests import kat


def _solve(oracle_mod, prob, ops):
    prog = cycles.build_program(prob, ops)
    cyc = oracle_mod.OracleProblem(prob).build(prog)
    return cyc.solve(prob.settings.tol, prob.settings.max_iters, 1)


def test_tutorial_known_answer(oracle_mod):
    prob = problems.Poisson2D(5, 9)
    out = _solve(oracle_mod, prob, kat.drop_jacobi(kat.tutorial_ops()))
    t, cf, iters = fitness.fitness_from_history(out.residuals, out.time_ms, prob.settings.max_iters)
    assert iters == kat.EXPECTED_ITERS
    # the notebook value and the restatement differ in the last bit (order of the pow products)
    assert abs(cf - kat.EXPECTED_CF) <= 2e-16 * kat.EXPECTED_CF * 2
    # first per-iteration factors as the generated binary would print them
    rhos = [float("%.6g" % (out.residuals[k + 1] / out.residuals[k])) for k in range(6)]
    assert rhos == [1.14101, 0.60437, 0.650816, 0.719709, 0.790328, 0.841953]
    assert abs(out.initial_residual - 1.1875651441228032e7) < 1e-6
    assert abs(out.final_residual / 3.263274e3 - 1) < 1e-6


def test_print_rounding_matters(oracle_mod):
    prob = problems.Poisson2D(5, 9)
    out = _solve(oracle_mod, prob, kat.drop_jacobi(kat.tutorial_ops()))
    cf_raw = fitness.convergence_factor(out.residuals, print_digits=None)
    assert abs(cf_raw - kat.EXPECTED_CF) > 1e-9      # without the 6-digit print rounding the KAT is missed


def test_intended_jacobi_semantics(oracle_mod):
    """With working `with jacobi` statements the same cycle converges faster (~0.8651)."""
    prob = problems.Poisson2D(5, 9)
    out = _solve(oracle_mod, prob, kat.tutorial_ops())
    cf = fitness.fitness_from_history(out.residuals, out.time_ms, 100)[1]
    assert abs(cf - 0.8651) < 5e-4


def test_classical_cycles(oracle_mod):
    """Restatement-derived sanity numbers of SURVEY.md Appendix C."""
    prob = problems.Poisson2D(5, 9)
    for prog, its, cf_ref in ((cycles.default_solver_cycle(prob), 7, 0.01781),
                              (cycles.v_cycle(prob, 2, 2, 1.0, True), 9, 0.0370),
                              (cycles.v_cycle(prob, 2, 2, 0.8, False), 15, 0.1558)):
        out = oracle_mod.OracleProblem(prob).build(prog).solve(1e-12, 100, 1)
        cf = fitness.fitness_from_history(out.residuals, out.time_ms, 100)[1]
        assert out.iterations == its
        assert abs(cf - cf_ref) < 2e-4


def test_poisson_solution_error(oracle_mod):
    """The converged solve reproduces the analytic solution cos(pi x) - sin(2 pi y) to O(h^2)."""
    from evostencils_b200 import oplist as ol
    prob = problems.Poisson2D(3, 6)
    cyc = oracle_mod.OracleProblem(prob).build(cycles.default_solver_cycle(prob))
    out = cyc.solve(1e-12, 100, 1)
    assert out.iterations < 15
    u = cyc.get_field(6, ol.BUF_SOL)
    n = prob.nodes(6)
    ax = np.arange(n) / (n - 1)
    exact = np.cos(math.pi * ax)[None, :] - np.sin(2 * math.pi * ax)[:, None]
    assert np.abs(u - exact).max() < 2e-3
