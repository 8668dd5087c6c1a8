"""Domain decomposition: the slab-wise solve of one grid must be BIT-IDENTICAL to the
undecomposed solve (which the other GPU tests pin against the oracle) for any number of slabs.  All slabs
live on cuda:0 here (LocalComm); the NCCL exchange path is the same schedule with send/recv instead of copies
and is covered on CPU by tests/test_domain_layout.py (gloo) and on GPUs by bench.py --gpus N."""
import numpy as np
import pytest

from evostencils_b200 import cycles, domain, oplist as ol, problems

pytestmark = pytest.mark.gpu


def _reference(cuda_backend, prob, prog):
    dev = cuda_backend.DeviceProblem(prob)
    cyc = dev.build(prog)
    s = prob.settings
    out = cyc.solve(s.tol, s.max_iters, 1, 2)   # keep state
    sol = cyc.get_field(prob.max_level, ol.BUF_SOL)
    cyc.close(); dev.close()
    return out, sol


@pytest.mark.parametrize("ghost", [2, 6])
@pytest.mark.parametrize("max_level,world,lc,red_black", [
    (6, 2, 5, True), (6, 3, 5, True), (6, 4, 6, True), (7, 2, 5, True), (7, 5, 6, True), (6, 2, 5, False), (7, 3, 7, True),
])
def test_slab_solve_is_bit_identical(cuda_backend, max_level, world, lc, red_black, ghost):
    """ghost = 2: an exchange after (almost) every statement; ghost = 6: consecutive sweeps recompute the halo on the
    extended planes and exchange once per level and cycle."""
    prob = problems.Poisson3D(2, max_level)
    s = prob.settings
    prog = cycles.v_cycle(prob, 2, 1, s.damping if red_black else 0.8, red_black)
    ref, ref_sol = _reference(cuda_backend, prob, prog)
    dd = domain.DomainSolver.emulate(prob, prog, world, lc, ghost=ghost)
    try:
        out = dd.solve(s.tol, s.max_iters)
        assert out.iterations == ref.iterations
        assert np.array_equal(out.residuals, ref.residuals)
        sol = dd.gather_solution()
        assert np.array_equal(sol, ref_sol)
        assert out.exchanges > 0
    finally:
        dd.close()


def test_w_cycle_and_default_lc(cuda_backend):
    prob = problems.Poisson3D(2, 6)
    prog = cycles.w_cycle(prob, 1, 1, 1.1, True)
    ref, ref_sol = _reference(cuda_backend, prob, prog)
    dd = domain.DomainSolver.emulate(prob, prog, 2)
    try:
        out = dd.solve(prob.settings.tol, prob.settings.max_iters)
        assert out.iterations == ref.iterations and np.array_equal(out.residuals, ref.residuals)
        assert np.array_equal(dd.gather_solution(), ref_sol)
    finally:
        dd.close()


@pytest.mark.parametrize("world,lc", [(2, 5), (3, 6), (2, 7)])
def test_fused_residual_restrict_in_slabs(cuda_backend, world, lc):
    from evostencils_b200 import lowering
    prob = problems.Poisson3D(2, 7)
    prog = lowering.optimise(cycles.default_solver_cycle(prob))
    assert any(o.code == ol.OP_RESIDUAL_RESTRICT for o in prog.ops)
    ref, ref_sol = _reference(cuda_backend, prob, prog)
    dd = domain.DomainSolver.emulate(prob, prog, world, lc)
    try:
        out = dd.solve(prob.settings.tol, prob.settings.max_iters)
        assert out.iterations == ref.iterations and np.array_equal(out.residuals, ref.residuals)
        assert np.array_equal(dd.gather_solution(), ref_sol)
    finally:
        dd.close()


@pytest.mark.parametrize("ghost", [2, 4, 6])
@pytest.mark.parametrize("world,lc,red_black", [(2, 6, True), (3, 5, True), (2, 6, False)])
def test_boundary_first_execution_is_identical(cuda_backend, world, lc, red_black, ghost):
    """overlap mode: every smoothing statement runs on the boundary planes first, then (while the halo travels) on the
    interior; the partial launches and the deferred exchange of SOL and its [next] slot must not change a bit."""
    prob = problems.Poisson3D(2, 7)
    prog = cycles.v_cycle(prob, 2, 1, 1.25 if red_black else 0.8, red_black)
    ref, ref_sol = _reference(cuda_backend, prob, prog)
    dd = domain.DomainSolver.emulate(prob, prog, world, lc, ghost=ghost)
    dd.overlap = True
    try:
        out = dd.solve(prob.settings.tol, prob.settings.max_iters)
        assert out.iterations == ref.iterations and np.array_equal(out.residuals, ref.residuals)
        assert np.array_equal(dd.gather_solution(), ref_sol)
    finally:
        dd.close()


def test_unsupported_statements_are_refused(cuda_backend):
    prob = problems.Poisson3D(2, 6)
    prog = cycles.v_cycle(prob, 1, 1, 1.0, True)
    prog.ops.insert(0, ol.Op(ol.OP_RICHARDSON, prob.max_level, omega=0.1))
    with pytest.raises(ValueError):
        domain.DomainSolver.emulate(prob, prog, 2)
