"""the roofline-aware surrogate mirrors the reference's PerformanceEvaluator interface
(model_based_prediction/performance.py:6-48) and reproduces the measured cycle times of profiles/ to ~30 %."""
import json
import os

import pytest

from evostencils_b200 import cycles, lowering, oplist as ol, problems
from evostencils_b200.performance import B200PerformanceEvaluator

HERE = os.path.dirname(os.path.abspath(__file__))


def test_interface_of_the_reference_evaluator():
    ev = B200PerformanceEvaluator(40e12, 6554.6e9, 8)
    assert ev.peak_performance == 40e12 and ev.peak_bandwidth == 6554.6e9 and ev.bytes_per_word == 8
    assert ev.runtime_coarse_grid_solver == 0
    ev.set_runtime_of_coarse_grid_solver(1e-3)
    assert ev.runtime_coarse_grid_solver == 1e-3
    with pytest.raises(RuntimeError):
        ev.estimate_runtime(object())          # a tree needs a generator to be lowered


def test_estimate_matches_the_measured_513_cycle():
    prob = problems.Poisson3D(2, 9)
    prog = lowering.optimise(cycles.default_solver_cycle(prob))
    est_ms = B200PerformanceEvaluator().estimate_runtime(prog) * 1e3
    line = json.load(open(os.path.join(HERE, "..", "profiles", "r2_e_bench_grid513.json")))
    cycle_ms = line["ms_per_cycle"] - 0.40      # the bench figure includes the norm residual of the solver loop
    assert abs(est_ms - cycle_ms) / cycle_ms < 0.3, (est_ms, cycle_ms)


def test_estimates_order_cycles_sensibly():
    prob = problems.Poisson3D(2, 8)
    ev = B200PerformanceEvaluator()
    v21 = ev.estimate_runtime(cycles.v_cycle(prob, 2, 1, 1.25, True))
    v11 = ev.estimate_runtime(cycles.v_cycle(prob, 1, 1, 1.25, True))
    w21 = ev.estimate_runtime(cycles.w_cycle(prob, 2, 1, 1.25, True))
    assert v11 < v21 < w21
    fused = ev.estimate_runtime(lowering.optimise(cycles.v_cycle(prob, 2, 1, 1.25, True)))
    assert fused < v21
    small = ev.estimate_runtime(cycles.v_cycle(problems.Poisson3D(2, 5), 2, 1, 1.25, True))
    assert small < 0.2 * v21 and small > 20 * 3e-6      # latency floor of ~40 statements


def test_tree_input_is_lowered_without_a_device():
    from evostencils_b200 import tree
    from tests import kat
    prob = problems.Poisson2D(5, 9)
    expr = tree.build_tree(prob, kat.TUTORIAL_INDIVIDUAL)
    ev = B200PerformanceEvaluator(problem=prob)
    r = ev.estimate_runtime(expr)
    assert 20e-6 < r < 2e-3
    assert ev.estimate_runtime(expr) == r          # cached on the node like the reference does (performance.py:51-52)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["poisson2d", "poisson3d", "elasticity"])
def test_measured_statement_costs_predict_the_cycle_time(cuda_backend, name):
    """measured=True: statement shapes are timed once on the device and cached; the sum over a cycle's statements must
    predict the measured time per cycle of random individuals (median error < 35 %) and rank them (Spearman > 0.5: the
    individuals of one problem differ by ~20 % in cost, so the ranks are noisy)."""
    import random
    import numpy as np
    from evostencils_b200 import tree
    from evostencils_b200.program_generator import B200ProgramGenerator
    prob = {"poisson2d": problems.Poisson2D(5, 9), "poisson3d": problems.Poisson3D(2, 6), "elasticity": problems.LinearElasticity2D(4, 8)}[name]
    pg = B200ProgramGenerator(problem=prob)
    ev = B200PerformanceEvaluator(generator=pg, measured=True)
    rng = random.Random(2)
    pred, meas = [], []
    s = prob.settings
    for i in range(14):
        expr = tree.build_tree(prob, tree.random_individual(prob, rng, maximum_local_system_size=4))
        prog = pg._finalise(pg.lower(expr, prob.min_level))
        before = ev.device_measurements
        p = ev.estimate_runtime(prog)
        if i >= 10:
            assert ev.device_measurements - before <= 6      # late individuals mostly reuse cached shapes
        cyc = pg._device_problem(prob.min_level, prob.max_level).build(prog)
        out = cyc.solve(s.tol, 12, 3)                          # 12 iterations are enough to time a cycle
        cyc.close()
        if out.iterations < 3:
            continue
        # one iteration of the solver loop = the cycle + the norm residual of the finest level
        norm = ev.estimate_runtime(ol.Program(dim=prob.dim, n_fields=prob.n_fields, min_level=prob.min_level, max_level=prob.max_level,
                                              ops=[ol.Op(ol.OP_RESIDUAL, prob.max_level, dst=ol.BUF_RES)], operators=prog.operators,
                                              restrict_w=prog.restrict_w, prolong_w=prog.prolong_w))
        pred.append((p + norm) * 1e3)
        meas.append(out.time_ms / out.iterations)
    pred, meas = np.array(pred), np.array(meas)
    assert len(pred) >= 8
    rel = np.abs(pred - meas) / meas
    rank = lambda v: np.argsort(np.argsort(v)).astype(float)
    rho = np.corrcoef(rank(pred), rank(meas))[0, 1]
    assert np.median(rel) < 0.35, (pred, meas)
    assert rho > 0.5, (rho, pred, meas)
    assert ev.estimate_runtime(prog) == p and ev.device_measurements > 0
    pg.close()
