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
