"""GPU: Helmholtz 2-D (complex, Robin x-boundaries): evolved/shifted-Laplacian cycle as preconditioner of the
outer BiCGStab, against the oracle.  Explicit Smith division / plain complex products / canonical reductions on
both sides make the histories bit-identical."""
import numpy as np
import pytest

from evostencils_b200 import cycles, oplist as ol, problems

pytestmark = pytest.mark.gpu


def test_helmholtz_cycle_statements_bit_exact(cuda_backend, oracle_mod):
    prob = problems.Helmholtz2D(3, 6, k=40.0)
    prog = cycles.default_solver_cycle(prob)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    # a non-trivial right-hand side for the preconditioner equation M u = f
    rng = np.random.default_rng(3)
    n = prob.nodes(6)
    f = np.zeros((n, n), dtype=np.complex128)
    f[1:-1, 1:-1] = rng.standard_normal((n - 2, n - 2)) + 1j * rng.standard_normal((n - 2, n - 2))
    for c in (gc, oc):
        c.set_field(6, ol.BUF_SOL, 0, np.zeros((n, n), dtype=np.complex128))
    # the finest rhs of a Helmholtz cycle is private (rewritten per application): set it on both
    gc.set_field(6, ol.BUF_RHS, 0, f)
    oc.set_field(6, ol.BUF_RHS, 0, f)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        for l in range(3, 7):
            for b in (ol.BUF_SOL, ol.BUF_RHS, ol.BUF_RES):
                assert np.array_equal(gc.get_field(l, b), oc.get_field(l, b)), (l, b)


# (80, 3..7) is the shipped configuration; (160, 4..8) and (320, 5..9) are the generalisation steps of BASELINE configs[3]
# (one level finer, k doubled: optimization/program.py:110-146 -> exastencils.py:196-215)
@pytest.mark.parametrize("k,levels", [(40.0, (3, 6)), (80.0, (3, 7)), (160.0, (4, 8)), (320.0, (5, 9))])
def test_helmholtz_outer_solver_parity(cuda_backend, oracle_mod, k, levels):
    prob = problems.Helmholtz2D(levels[0], levels[1], k=k)
    prog = cycles.default_solver_cycle(prob)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    # the shipped sizes are solved to the end; the generalisation steps need thousands of outer iterations with the
    # template's V(2,1) preconditioner: their first 400 iterations (800 cycle applications) are compared
    max_iters = prob.settings.max_iters if k <= 80.0 else 400
    a = gc.helmholtz_solve(prob.settings.tol, max_iters, 1)
    b = oc.helmholtz_solve(prob.settings.tol, max_iters, 1)
    assert a.iterations == b.iterations and a.iterations > 10
    assert np.array_equal(a.residuals, b.residuals)
    if k <= 80.0:
        assert a.iterations < 2000 and a.final_residual < 1e-7 * a.initial_residual


def test_helmholtz_through_the_drop_in(cuda_backend, oracle_mod):
    """generate_and_evaluate on the Helmholtz problem: outer BiCGStab iterations / total reduction as the
    reference's parse_output would report them, and the k, 2k, 4k triple run (exastencils.py:518-532)."""
    from evostencils_b200 import fitness, tree
    from evostencils_b200.program_generator import B200ProgramGenerator
    prob = problems.Helmholtz2D(3, 6, k=20.0)
    pg = B200ProgramGenerator(problem=prob, solver_iteration_limit=10000)
    storages = pg.generate_storage(3, 6, pg.finest_grid)
    s = tree.v_cycle_individual(3, 2, 1, 10)            # RB-GS omega = 0.6 V(2,1): the template's preconditioner
    expression = tree.build_tree(prob, s)
    t, cf, its = pg.generate_and_evaluate(expression, storages, 3, 6, "", evaluation_samples=1)
    prog = pg._finalise(pg.lower(expression, 3))
    ref = oracle_mod.OracleProblem(prob).build(prog).helmholtz_solve(prob.settings.tol, prob.settings.max_iters, 1)
    want = fitness.helmholtz_fitness(ref.residuals, ref.time_ms, prob.settings.max_iters, 1e100, 10000, prob.settings.tol)
    assert (cf, its) == want[1:] and its == ref.iterations - 1 and cf < 1e-7
    # triple run: averages over k, 2k, 4k on the same grid
    t3, cf3, its3 = pg.generate_and_evaluate(expression, storages, 3, 6, "", evaluation_samples=1,
                                             global_variable_values={"k": 10.0})
    singles = []
    for k in (10.0, 20.0, 40.0):
        pk = problems.Helmholtz2D(3, 6, k=k)
        pgk = B200ProgramGenerator(problem=pk, solver_iteration_limit=10000)
        singles.append(pgk.generate_and_evaluate(tree.build_tree(pk, s), storages, 3, 6, "", evaluation_samples=1))
        pgk.close()
    assert abs(its3 - sum(x[2] for x in singles) / 3) < 1e-12
    assert abs(cf3 - sum(x[1] for x in singles) / 3) < 1e-18
    pg.close()
