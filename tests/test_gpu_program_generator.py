"""GPU: the drop-in ProgramGenerator surface (exastencils.py:485 generate_and_evaluate etc.)."""
import json
import os
import random

import numpy as np
import pytest

from evostencils_b200 import fitness, problems, tree
from evostencils_b200.program_generator import B200ProgramGenerator
from tests import kat
from tests.test_lowering_golden import PROBLEMS, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_generate_and_evaluate_matches_golden(cuda_backend, name):
    prob, recs = load(name)
    pg = B200ProgramGenerator(problem=prob)
    assert pg.dimension == prob.dim and pg.min_level == prob.min_level and pg.max_level == prob.max_level
    assert len(pg.fields) == prob.n_fields and pg.uses_FAS is False and pg.mpi_rank == 0
    storages = pg.generate_storage(pg.min_level, pg.max_level, pg.finest_grid)
    pg.initialize_code_generation(pg.min_level, pg.max_level)
    for rec in recs:
        expression = tree.build_tree(prob, rec["individual"])
        t, cf, its = pg.generate_and_evaluate(expression, storages, pg.min_level, pg.max_level, "", evaluation_samples=2)
        assert cf == rec["oracle"]["convergence_factor"], rec["individual"]
        assert its == rec["oracle"]["fitness_iterations"]
        assert 0 < t < 1e5
        text = pg.generate_cycle_function(expression, storages, pg.min_level, pg.max_level, pg.max_level)
        assert text.startswith(f"Function gen_mgCycle@{pg.max_level} {{")
    assert pg.total_kernel_launches > 0
    pg.close()


def test_tutorial_known_answer_through_the_drop_in(cuda_backend):
    """Reference fixture notebooks/tutorial.ipynb:3373 through string -> tree -> generate_and_evaluate."""
    prob = problems.Poisson2D(5, 9)
    pg = B200ProgramGenerator(problem=prob, jacobi_compat="exastencils_v1_1_noop")
    storages = pg.generate_storage(5, 9, pg.finest_grid)
    pg.initialize_code_generation(5, 9)
    expression = tree.build_tree(prob, kat.TUTORIAL_INDIVIDUAL)
    t, cf, its = pg.generate_and_evaluate(expression, storages, 5, 9, "", evaluation_samples=3)
    assert its == kat.EXPECTED_ITERS
    assert abs(cf - kat.EXPECTED_CF) < 1e-15
    # intended Jacobi semantics: the same individual converges faster
    pg2 = B200ProgramGenerator(problem=prob)
    t2, cf2, its2 = pg2.generate_and_evaluate(expression, storages, 5, 9, "", evaluation_samples=1)
    assert abs(cf2 - 0.8651) < 5e-4
    pg.close(); pg2.close()


def test_bad_individuals_return_sentinels(cuda_backend):
    prob = problems.Poisson2D(3, 5)
    pg = B200ProgramGenerator(problem=prob)
    storages = pg.generate_storage(3, 5, pg.finest_grid)
    # not a tree at all -> "code generation failed" -> (infinity,)*3, never raises (exastencils.py:499-510)
    assert pg.generate_and_evaluate(object(), storages, 3, 5, "") == (1e100, 1e100, 1e100)
    # diverging cycle: cf > 1 is returned as is (exastencils.py:436-437)
    s = tree.v_cycle_individual(2, 2, 2, 36, partitioning="single")
    t, cf, its = pg.generate_and_evaluate(tree.build_tree(prob, s), storages, 3, 5, "")
    assert cf > 1 and its == 100
    pg.close()


def test_population_evaluation_equals_one_by_one(cuda_backend):
    prob = problems.Poisson2D(3, 6)
    pg = B200ProgramGenerator(problem=prob)
    storages = pg.generate_storage(3, 6, pg.finest_grid)
    rng = random.Random(7)
    strings = [tree.random_individual(prob, rng) for _ in range(24)]
    one_by_one = [pg.generate_and_evaluate(tree.build_tree(prob, s), storages, 3, 6, "", evaluation_samples=1) for s in strings]
    batch, ms = pg.evaluate_population([tree.build_tree(prob, s) for s in strings], max_in_flight=16)
    assert ms > 0
    for a, b in zip(one_by_one, batch):
        assert a[1] == b[1] or (np.isnan(a[1]) and np.isnan(b[1]))
        assert a[2] == b[2]
    pg.close()


def test_timeout_gives_the_sentinel(cuda_backend):
    """evaluation_timeout (exastencils.py:42, :430-433, :476-483): an evaluation that runs longer returns
    (infinity,)*3; here the watchdog lives in the device-side solver loop."""
    prob = problems.Poisson2D(5, 9)
    slow = tree.build_tree(prob, tree.v_cycle_individual(4, 1, 1, 2, partitioning="single"))   # weak damping: many iterations
    pg = B200ProgramGenerator(problem=prob, evaluation_timeout=0.002)
    storages = pg.generate_storage(5, 9, pg.finest_grid)
    assert pg.generate_and_evaluate(slow, storages, 5, 9, "", evaluation_samples=1) == (1e100, 1e100, 1e100)
    assert pg.last_outcome.status == 2 and 0 < pg.last_outcome.iterations < 100
    pg.close()
    pg = B200ProgramGenerator(problem=prob, evaluation_timeout=300)
    t, cf, its = pg.generate_and_evaluate(slow, storages, 5, 9, "", evaluation_samples=1)
    assert pg.last_outcome.status == 0 and cf < 1e100
    pg.close()


def test_infrastructure_errors_are_not_fitness_values(cuda_backend):
    """A statement the library does not implement must raise, not masquerade as a diverged individual."""
    from evostencils_b200 import backend, cycles, oplist as ol
    prob = problems.Poisson2D(3, 5)
    pg = B200ProgramGenerator(problem=prob)
    unk = ((0, (0, 0)), (0, (1, 0)))
    prog = cycles.build_program(prob, [ol.Op(ol.OP_SMOOTH, 5, mode=ol.MODE_REDBLACK, omega=1.0, unknowns=unk)])
    with pytest.raises(backend.BackendError) as info:
        pg._evaluate_program(prog, 3, None, 1e100, 1)
    assert info.value.status == backend.ERR_UNSUPPORTED and info.value.infrastructure
    pg.close()


def test_population_time_objective_is_contention_free(cuda_backend):
    """solo_timing (default): the time entry of a batch member is measured with the GPU to itself and is
    comparable with generate_and_evaluate; cf / iterations are unchanged by it."""
    prob = problems.Poisson2D(5, 9)
    pg = B200ProgramGenerator(problem=prob)
    storages = pg.generate_storage(5, 9, pg.finest_grid)
    rng = random.Random(3)
    strings = [tree.random_individual(prob, rng) for _ in range(32)]
    trees = [tree.build_tree(prob, s) for s in strings]
    pg.evaluate_population(trees[:4])                                      # warm-up
    solo, _ = pg.evaluate_population(trees, max_in_flight=32, solo_timing=True)
    crowd, _ = pg.evaluate_population(trees, max_in_flight=32, solo_timing=False)
    single = [pg.generate_and_evaluate(t, storages, 5, 9, "", evaluation_samples=3) for t in trees]
    checked = 0
    for a, b, c in zip(solo, crowd, single):
        assert a[1] == b[1] == c[1] or (np.isnan(a[1]) and np.isnan(c[1]))
        assert a[2] == b[2] == c[2]
        if c[2] < 100 and c[1] < 1:
            checked += 1
            assert abs(a[0] - c[0]) <= 0.25 * c[0] + 0.05, (a[0], c[0])     # ms; launch jitter of a short solve
    assert checked >= 4
    pg.close()
