"""A problem that is NOT one of the reference's five example problems, written in ExaSlang layer 2/3 for these tests
(tests/fixtures/user_problem: anisotropic diffusion-reaction with a mixed derivative -> 9-point stencil, its own
globals, boundary / right-hand-side expressions and solver block).  The drop-in constructor must take it from the
configuration triple exactly like the reference's ProgramGenerator (exastencils.py:39-110; parser.py:25-143)."""
import os

import numpy as np
import pytest

from evostencils_b200 import cycles, fitness, frontend, oplist as ol
from evostencils_b200.program_generator import _problem_from_paths

BASE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "user_problem")
SETTINGS, KNOWLEDGE = "AnisoDiffusion/2D_FD_AnisoDiffusion.settings", "AnisoDiffusion/2D_FD_AnisoDiffusion.knowledge"


def test_front_end_reads_the_user_problem(oracle_mod):
    p = _problem_from_paths(SETTINGS, KNOWLEDGE, BASE)
    assert isinstance(p, frontend.ExaProblem) and p.name == "2D_FD_AnisoDiffusion"
    assert (p.dim, p.min_level, p.max_level) == (2, 3, 7)
    assert p.fields == ("w",) and p.rhs_names == ("RHS_w",) and p.equation_names == ("anisoEq",)
    assert p.parameters == {"eps": 0.25, "beta": 0.125, "sigma": 3.0}
    s = p.settings
    assert (s.tol, s.max_iters, s.num_pre, s.num_post, s.damping, s.red_black, s.cgs_max_iters, s.cgs_tol) == \
        (1e-10, 60, 2, 2, 1.0, True, 500, 1e-10)
    # rediscretised 9-point operator at level 5 (h = 1/32): literal evaluation of the stencil expressions
    h = 1.0 / 32
    expect = {(0, 0): 2.0 * 0.25 / h ** 2 + 2.0 / h ** 2 + 3.0, (-1, 0): -0.25 / h ** 2, (1, 0): -0.25 / h ** 2,
              (0, -1): -1.0 / h ** 2, (0, 1): -1.0 / h ** 2, (-1, -1): 0.25 / (4 * h * h), (1, 1): 0.25 / (4 * h * h),
              (-1, 1): -0.25 / (4 * h * h), (1, -1): -0.25 / (4 * h * h)}
    table = p.operator(5)
    for off, v in expect.items():
        assert table[0, 0, ol.stencil_index(off)] == v
    assert np.count_nonzero(table) == 9
    # the problem is solvable with its own solver block and second-order accurate against the manufactured solution
    out = oracle_mod.OracleProblem(p).build(cycles.default_solver_cycle(p))
    res = out.solve(s.tol, s.max_iters, 1)
    assert res.iterations < 30 and res.residuals[-1] < 1e-10 * res.residuals[0]
    n = p.nodes(7)
    x = np.linspace(0.0, 1.0, n)
    X, Y = np.meshgrid(x, x)
    err = np.abs(out.get_field(7, ol.BUF_SOL, 0) - (np.sin(np.pi * X) * np.exp(Y) + X * Y)).max()
    assert err < 5e-5


def test_constructor_errors_like_the_reference():
    with pytest.raises(RuntimeError):
        _problem_from_paths("Nope/missing.settings", "Nope/missing.knowledge", BASE)


@pytest.mark.gpu
def test_user_problem_through_the_drop_in(cuda_backend, oracle_mod):
    """Sixth problem through B200ProgramGenerator(base_path, settings_path, knowledge_path): the evolved-cycle path
    (random individuals incl. coloured sweeps on the 9-point operator, which are order dependent) agrees with the
    oracle bit for bit."""
    import random
    from evostencils_b200 import tree
    from evostencils_b200.program_generator import B200ProgramGenerator
    pg = B200ProgramGenerator(None, BASE, SETTINGS, KNOWLEDGE, None, mpi_rank=0)
    prob = pg.problem
    assert prob.name == "2D_FD_AnisoDiffusion" and pg.dimension == 2 and (pg.min_level, pg.max_level) == (3, 7)
    storages = pg.generate_storage(3, 7, pg.finest_grid)
    rng = random.Random(11)
    strings = [tree.v_cycle_individual(4, 2, 2, 18)] + [tree.random_individual(prob, rng) for _ in range(6)]
    ref = oracle_mod.OracleProblem(prob)
    for s in strings:
        expr = tree.build_tree(prob, s)
        t, cf, its = pg.generate_and_evaluate(expr, storages, 3, 7, "", evaluation_samples=1)
        prog = pg._finalise(pg.lower(expr, 3))
        o = ref.build(prog).solve(prob.settings.tol, prob.settings.max_iters, 1)
        assert np.array_equal(pg.last_outcome.residuals, o.residuals), s
        rt, rcf, rits = fitness.fitness_from_history(o.residuals, o.time_ms, prob.settings.max_iters)
        assert its == rits and (cf == rcf or (np.isnan(cf) and np.isnan(rcf)))
    pg.close()
