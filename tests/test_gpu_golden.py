"""GPU: the golden individuals (lowered from the reference's own trees) through the CUDA path; residual
histories must equal the recorded oracle histories bit for bit, fitness tuples exactly."""
import numpy as np
import pytest

from evostencils_b200 import fitness, lowering, oplist as ol
from tests.test_lowering_golden import PROBLEMS, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(PROBLEMS))
@pytest.mark.parametrize("fuse", [False, True])
def test_golden_individuals_on_gpu(cuda_backend, name, fuse):
    prob, recs = load(name)
    dev = cuda_backend.DeviceProblem(prob)
    for rec in recs:
        prog = ol.Program.from_json(rec["program"])
        if fuse:
            prog = lowering.optimise(prog)
        cyc = dev.build(prog)
        out = cyc.solve(prob.settings.tol, prob.settings.max_iters, 1)
        want = np.array([float.fromhex(h) for h in rec["oracle"]["residuals"]])
        assert out.iterations == rec["oracle"]["iterations"], rec["individual"]
        assert np.array_equal(out.residuals, want, equal_nan=True), rec["individual"]
        t, cf, its = fitness.fitness_from_history(out.residuals, out.time_ms, prob.settings.max_iters)
        assert cf == rec["oracle"]["convergence_factor"] and its == rec["oracle"]["fitness_iterations"]
        cyc.close()
