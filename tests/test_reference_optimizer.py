"""The call site the drop-in boundary exists for: the REFERENCE's own ``Optimizer``
(evostencils/optimization/program.py:67-108) constructed around ``B200ProgramGenerator`` and driven through
``generate_and_evaluate_program_from_grammar_representation`` (:904-932) and ``evaluate_single_objective`` /
``evaluate_multiple_objectives`` (:386-453).

The reference checkout only exists in the build container (not on the GPU box) and the GPU only exists on the box, so
this CPU test substitutes the two things the host logic does not own: DEAP (tests/deap_stub.py) and the device -- the
generator's ``backend`` is patched with an adapter over the CPU oracle, which reproduces the reference's known answer
(tests/test_oracle_kat.py).  Everything in between is the product's host code: constructor from the reference's
configuration files (front-end), attributes the Optimizer reads, lowering of the reference's IR tree, fitness
extraction, sentinels."""
import os
import sys

import pytest

from tests import kat

REF = os.environ.get("EVOSTENCILS_REFERENCE", "/root/reference")
BASE = os.path.join(REF, "example_problems")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "evostencils")), reason="reference checkout not available")


class _OracleDevice:
    """DeviceProblem look-alike over the CPU oracle (test infrastructure)."""

    def __init__(self, problem, device=0):
        from oracle import oracle as orc
        self._impl = orc.OracleProblem(problem)
        self.problem = problem

    def build(self, program):
        cyc = self._impl.build(program)
        solve = cyc.solve
        # timing repeats (evaluation_samples = 20 at program.py:926) do not change cf / iterations: once is enough here
        cyc.solve = lambda tol, max_iters, samples=1, flags=0, timeout_ms=0: solve(tol, max_iters, 1, flags)
        return cyc

    def close(self):
        self._impl.close()


@pytest.fixture
def reference_optimizer(monkeypatch):
    from tests import deap_stub
    deap_stub.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from evostencils_b200 import backend

    class _Lib:
        @staticmethod
        def evo_device_count():
            return 1

    monkeypatch.setattr(backend, "load_library", lambda path=None: _Lib)
    monkeypatch.setattr(backend, "DeviceProblem", _OracleDevice)
    from evostencils.optimization.program import Optimizer
    return Optimizer


def _optimizer(Optimizer, pg):
    return Optimizer(pg.dimension, pg.finest_grid, pg.coarsening_factor, pg.min_level, pg.max_level, pg.equations,
                     pg.operators, pg.fields, program_generator=pg, mpi_rank=0, number_of_mpi_processes=1)


def test_tutorial_individual_through_the_reference_optimizer(reference_optimizer):
    """notebooks/tutorial.ipynb:3373: cf = 0.9212764170210773 after 100 iterations, through the reference's
    own Optimizer.generate_and_evaluate_program_from_grammar_representation."""
    from evostencils_b200.program_generator import B200ProgramGenerator
    pg = B200ProgramGenerator(None, BASE, "Poisson/2D_FD_Poisson_fromL2.settings", "Poisson/2D_FD_Poisson_fromL2.knowledge",
                              "lib/linux.platform", mpi_rank=0, jacobi_compat="exastencils_v1_1_noop")
    assert type(pg.problem).__name__ == "ExaProblem"           # built by the front-end from the user's files
    opt = _optimizer(reference_optimizer, pg)
    t, cf, its = opt.generate_and_evaluate_program_from_grammar_representation(kat.TUTORIAL_INDIVIDUAL, 8)
    assert its == kat.EXPECTED_ITERS == 100
    assert abs(cf - kat.EXPECTED_CF) < 1e-15      # last bit but one, like tests/test_oracle_kat.py
    assert t > 0
    assert pg._counter >= 1                                     # the attribute the Optimizer's logging reads


def test_objective_functions_of_the_reference(reference_optimizer):
    """evaluate_single_objective / evaluate_multiple_objectives on DEAP-style individuals (program.py:386-453): a
    converging V-cycle gives (time,) and (cf, time/iterations); a diverging one gives the sqrt(cf * iters) penalty."""
    from deap import gp
    from evostencils.grammar import multigrid as mg
    from evostencils_b200 import problems, tree
    from evostencils_b200.program_generator import B200ProgramGenerator
    prob = problems.Poisson2D(3, 6)
    pg = B200ProgramGenerator(problem=prob)
    opt = _optimizer(reference_optimizer, pg)
    pset, _ = mg.generate_primitive_set(opt.approximation, opt.rhs, opt.dimension, opt.coarsening_factors, opt.max_level,
                                        opt.equations, opt.operators, opt.fields, maximum_local_system_size=4,
                                        depth=opt.max_level - opt.min_level)
    storages = pg.generate_storage(3, 6, pg.finest_grid)

    class _Str(gp.PrimitiveTree):           # an individual whose str() is the grammar string (what gp.compile consumes)
        def __init__(self, s):
            super().__init__([None] * 10)
            self._s = s

        def __str__(self):
            return self._s

    good = _Str(tree.v_cycle_individual(3, 2, 1, 18))
    bad = _Str(tree.v_cycle_individual(3, 2, 2, 36, partitioning="single"))
    (time_,) = opt.evaluate_single_objective(good, pset, storages, 3, 6, "", evaluation_samples=1)
    assert 0 < time_ < 1e6
    opt.clear_individual_cache()
    cf, t_per_it = opt.evaluate_multiple_objectives(good, pset, storages, 3, 6, "", evaluation_samples=1)
    assert 0 < cf < 0.3 and 0 < t_per_it < 1e6
    (penalty,) = opt.evaluate_single_objective(bad, pset, storages, 3, 6, "", evaluation_samples=1)
    assert penalty > 10            # sqrt(cf) * sqrt(1e100): the iteration sentinel of a solver that hit the limit


def test_generator_attributes_the_optimizer_reads(reference_optimizer):
    from evostencils_b200.program_generator import B200ProgramGenerator
    pg = B200ProgramGenerator(None, BASE, "LinearElasticity/2D_FD_LinearElasticity_fromL2.settings",
                              "LinearElasticity/2D_FD_LinearElasticity_fromL2.knowledge", "lib/linux.platform")
    opt = _optimizer(reference_optimizer, pg)
    assert opt.dimension == 2 and (opt.min_level, opt.max_level) == (4, 8)
    assert [f.name for f in opt.fields] == ["u", "v"]
    assert len(opt.equations) == 2 * 5 and {e.rhs_name for e in opt.equations} == {"RHS_u", "RHS_v"}
