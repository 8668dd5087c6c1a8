"""GPU: FAS_2D_Basic (nonlinear, -Lap u + gamma u e^u = f) against the oracle; both sides use the shared
evo_exp, so the histories are bit-identical."""
import numpy as np
import pytest

from evostencils_b200 import cycles, fitness, oplist as ol, problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("newton_steps,red_black", [(1, False), (2, False), (0, False), (1, True), (3, True)])
def test_fas_cycle_parity(cuda_backend, oracle_mod, newton_steps, red_black):
    prob = problems.FAS2D(3, 6)
    prog = cycles.fas_v_cycle(prob, 2, 2, 0.8, newton_steps, red_black)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    gc.apply(1)
    oc.apply(1)
    for l in range(3, 7):
        for b in (ol.BUF_SOL, ol.BUF_RHS, ol.BUF_RES, ol.BUF_APX):
            assert np.array_equal(gc.get_field(l, b), oc.get_field(l, b)), (l, b)
    a = gc.solve(prob.settings.tol, prob.settings.max_iters, 1)
    b = oc.solve(prob.settings.tol, prob.settings.max_iters, 1)
    assert a.iterations == b.iterations and a.iterations < 100
    assert np.array_equal(a.residuals, b.residuals)
    assert fitness.fas_fitness(a.residuals, a.time_ms)[1:] == fitness.fas_fitness(b.residuals, b.time_ms)[1:]


@pytest.mark.parametrize("lo,hi", [(7, 9), (8, 10), (5, 7)])
def test_fas_large_coarsest_grid_uses_multi_cta_sweeps(cuda_backend, oracle_mod, lo, hi):
    """Coarsest grids beyond one SM: rows split over a thread-block cluster with distributed-shared-memory halos
    (<= 129^2), or one launch per Newton-Jacobi sweep (BASELINE config shape: 257^2 below 4097^2) -- same
    arithmetic, same bits."""
    prob = problems.FAS2D(lo, hi)
    prog = cycles.fas_v_cycle(prob)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    a = gc.solve(prob.settings.tol, 6, 1)
    b = oc.solve(prob.settings.tol, 6, 1)
    assert a.iterations == b.iterations == 6
    assert np.array_equal(a.residuals, b.residuals)
    for l in range(lo, hi + 1):
        assert np.array_equal(gc.get_field(l, ol.BUF_SOL), oc.get_field(l, ol.BUF_SOL)), l


def test_fas_solution_converges_to_manufactured_solution(cuda_backend):
    prob = problems.FAS2D(3, 7)
    cyc = cuda_backend.DeviceProblem(prob).build(cycles.fas_v_cycle(prob))
    out = cyc.solve(1e-10, 300, 1)
    assert out.iterations < 30
    n = prob.nodes(7)
    ax = np.arange(n) / (n - 1)
    exact = prob.exact_solution(ax[None, :], ax[:, None])
    assert np.abs(cyc.get_field(7, ol.BUF_SOL) - exact).max() < 5e-3


def test_fas_golden_on_gpu_and_drop_in(cuda_backend):
    from evostencils_b200 import tree
    from evostencils_b200.program_generator import B200ProgramGeneratorFAS
    from tests.test_fas_golden import load
    prob, recs = load()
    pg = B200ProgramGeneratorFAS(problem=prob, cumulative_timer=False)
    assert pg.uses_FAS and pg.generate_storage(1, 2, 3) == []
    for rec in recs:
        if "oracle" not in rec:
            continue
        expression = tree.build_tree(prob, rec["individual"])
        t, c, n = pg.generate_and_evaluate(expression, [], prob.min_level, prob.max_level, "", evaluation_samples=1)
        want = np.array([float.fromhex(h) for h in rec["oracle"]["residuals"]])
        assert np.array_equal(pg.last_outcome.residuals, want, equal_nan=True), rec["individual"]
        assert c == rec["oracle"]["convergence_factor"] and n == rec["oracle"]["fitness_iterations"]
    pg.close()
