"""The tutorial known-answer test through the host stack: grammar string -> tree -> lowering (CPU)."""
from evostencils_b200 import fitness, lowering, oplist as ol, problems, tree
from tests import kat


def _lowered():
    prob = problems.Poisson2D(5, 9)
    expression = tree.build_tree(prob, kat.TUTORIAL_INDIVIDUAL)
    return prob, lowering.lower_cycle(expression, 5, 9, 1, 2, cgs_max_iters=1000, cgs_tol=1e-12)


def test_tutorial_individual_lowers_to_the_printed_cycle():
    prob, prog = _lowered()
    want = kat.tutorial_ops()
    assert [o.key() for o in prog.ops] == [o.key() for o in want]


def test_compat_mode_reproduces_the_notebook_number(oracle_mod):
    prob, prog = _lowered()
    prog = lowering.apply_jacobi_compat(prog, "exastencils_v1_1_noop")
    out = oracle_mod.OracleProblem(prob).build(prog).solve(1e-12, 100, 1)
    t, cf, its = fitness.fitness_from_history(out.residuals, out.time_ms, 100)
    assert its == kat.EXPECTED_ITERS and abs(cf - kat.EXPECTED_CF) < 1e-15


def test_fitness_sentinels_follow_parse_output():
    inf = 1e100
    # no iteration printed -> iterations = infinity (exastencils.py:580-581), no factor -> cf = infinity (:574-576)
    assert fitness.fitness_from_history([1.0], 0.5, 100)[1:] == (inf, inf)
    # iteration limit of the optimiser (:582-583)
    assert fitness.fitness_from_history([1.0, 0.5, 0.25], 1.0, 100, solver_iteration_limit=2)[2] == inf
    # non-finite residual: remaining iterations count with rho = sqrt(infinity) (:543, :550-551)
    t, cf, its = fitness.fitness_from_history([1.0, 0.5, float("inf")], 1.0, 4)
    assert its == 4 and cf > 1e20
    # all factors non-finite -> count == 0 -> cf = infinity
    assert fitness.fitness_from_history([1.0, float("nan")], 1.0, 3)[1] == inf
    # FAS variant: c = (res_final/res_initial)^(1/n) from 4-digit prints (exastencils_FAS.py:370-394)
    t, c, n = fitness.fas_fitness([123.456789, 1.23456789, 0.0123456789], 2.0)
    assert n == 2 and abs(c - (0.01235 / 123.5) ** 0.5) < 1e-15
