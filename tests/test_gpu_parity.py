"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same
inputs.  Bar (BASELINE.json north_star): identical iteration counts, per-iteration residual norms
within 1e-10 relative, convergence factor within 1e-6; statements without reductions are bit-exact."""
import numpy as np
import pytest

from evostencils_b200 import cycles, fitness, oplist as ol, problems
from tests import kat

pytestmark = pytest.mark.gpu

RES_RTOL = 1e-10
CF_ATOL = 1e-6


def _pair(cuda_backend, oracle_mod, prob, prog):
    dev = cuda_backend.DeviceProblem(prob)
    ref = oracle_mod.OracleProblem(prob)
    return dev.build(prog), ref.build(prog), dev, ref


def _assert_solve_parity(gc, oc, prob, flags=0):
    s = prob.settings
    a = gc.solve(s.tol, s.max_iters, 1, flags)
    b = oc.solve(s.tol, s.max_iters, 1)
    assert a.iterations == b.iterations
    assert a.status == b.status
    # every reduction uses the shared canonical order -> the histories are bit-identical, far inside
    # the 1e-10 bar of the north star
    np.testing.assert_allclose(a.residuals, b.residuals, rtol=RES_RTOL, atol=0)
    assert np.array_equal(a.residuals, b.residuals)
    cfa = fitness.fitness_from_history(a.residuals, a.time_ms, s.max_iters)[1]
    cfb = fitness.fitness_from_history(b.residuals, b.time_ms, s.max_iters)[1]
    assert abs(cfa - cfb) < CF_ATOL
    assert a.kernel_launches > 0
    return a, b


def _fields_equal(gc, oc, prob, levels, bufs=(ol.BUF_SOL, ol.BUF_RHS, ol.BUF_RES), exact=True):
    for l in levels:
        for b in bufs:
            for f in range(prob.n_fields):
                x, y = gc.get_field(l, b, f), oc.get_field(l, b, f)
                if exact:
                    assert np.array_equal(x, y), f"level {l} buf {b} field {f}: max diff {np.abs(x - y).max()}"
                else:
                    scale = max(np.abs(y).max(), 1e-300)
                    assert np.abs(x - y).max() <= 1e-12 * scale, f"level {l} buf {b} field {f}"


def _no_cg(prog):
    """Same cycle with the coarse solve replaced by 3 RB-GS sweeps -> no reduction feeds back, bit-exact."""
    ops = []
    for o in prog.ops:
        if o.code == ol.OP_COARSE_SOLVE:
            zero = (0,) * prog.dim
            unk = tuple((f, zero) for f in range(prog.n_fields))
            ops.append(ol.Op(ol.OP_ZERO, o.level))
            ops += [ol.Op(ol.OP_SMOOTH, o.level, mode=ol.MODE_JACOBI, omega=0.8, unknowns=unk) for _ in range(3)]
        else:
            ops.append(o)
    import copy
    p = copy.copy(prog)
    p.ops = ops
    return p


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("red_black", [True, False])
def test_poisson2d_statements_bit_exact(cuda_backend, oracle_mod, red_black):
    prob = problems.Poisson2D(3, 6)
    prog = _no_cg(cycles.v_cycle(prob, 2, 1, 1.15 if red_black else 0.8, red_black))
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, prog)
    for _ in range(3):
        gc.apply(1)
        oc.apply(1)
        _fields_equal(gc, oc, prob, range(3, 7))


def test_poisson2d_default_solver(cuda_backend, oracle_mod):
    prob = problems.Poisson2D(3, 7)
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.default_solver_cycle(prob))
    a, _ = _assert_solve_parity(gc, oc, prob)
    assert a.iterations < 12


def test_graph_and_direct_launch_agree(cuda_backend, oracle_mod):
    prob = problems.Poisson2D(3, 6)
    prog = cycles.default_solver_cycle(prob)
    dev = cuda_backend.DeviceProblem(prob)
    c1, c2 = dev.build(prog), dev.build(prog)
    a = c1.solve(1e-12, 100, 1)
    b = c2.solve(1e-12, 100, 1, ol.SOLVE_NO_GRAPH)
    assert a.iterations == b.iterations
    assert np.array_equal(a.residuals, b.residuals)


def test_tutorial_known_answer_on_gpu(cuda_backend, oracle_mod):
    """Full-size KAT (Poisson 2D 513^2, levels 9..5) on the GPU: notebooks/tutorial.ipynb:3373."""
    prob = problems.Poisson2D(5, 9)
    prog = cycles.build_program(prob, kat.drop_jacobi(kat.tutorial_ops()))
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    out = gc.solve(prob.settings.tol, prob.settings.max_iters, 1)
    _, cf, iters = fitness.fitness_from_history(out.residuals, out.time_ms, prob.settings.max_iters)
    assert iters == kat.EXPECTED_ITERS
    assert abs(cf - kat.EXPECTED_CF) < 1e-9


def test_tutorial_cycle_intended_jacobi(cuda_backend, oracle_mod):
    """Block Jacobi (3x2 and 1x6 local systems) + pointwise Jacobi + RB-GS, full size."""
    prob = problems.Poisson2D(5, 9)
    prog = cycles.build_program(prob, kat.tutorial_ops())
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, prog)
    _assert_solve_parity(gc, oc, prob)


@pytest.mark.parametrize("shape", [(2, 1), (1, 2), (2, 2), (3, 1), (1, 4), (4, 2), (3, 2)])
def test_block_jacobi_bit_exact(cuda_backend, oracle_mod, shape):
    prob = problems.Poisson2D(3, 5)
    unk = tuple((0, (i, j)) for i in range(shape[0]) for j in range(shape[1]))
    ops = [ol.Op(ol.OP_SMOOTH, 5, mode=ol.MODE_JACOBI, omega=0.7, unknowns=unk),
           ol.Op(ol.OP_SMOOTH, 5, mode=ol.MODE_JACOBI, omega=0.9, unknowns=unk),
           ol.Op(ol.OP_RESIDUAL, 5, dst=ol.BUF_RES)]
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.build_program(prob, ops))
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _fields_equal(gc, oc, prob, [5])   # dense solves share pivot rule and operation order with the oracle: bitwise


def test_elasticity_statements(cuda_backend, oracle_mod):
    """2-field system: collective RB-GS (order dependent -> row-sequential kernel), collective and
    decoupled Jacobi, decoupled RB-GS."""
    prob = problems.LinearElasticity2D(3, 6)
    z = (0, 0)
    coll = ((0, z), (1, z))
    ops = [ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=coll),
           ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_JACOBI, omega=0.8, unknowns=coll),
           ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_JACOBI, omega=0.7, unknowns=((0, z),)),
           ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_JACOBI, omega=0.7, unknowns=((1, z),)),
           ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_REDBLACK, omega=1.1, unknowns=((0, z),)),
           ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_REDBLACK, omega=1.1, unknowns=((1, z),)),
           ol.Op(ol.OP_RESIDUAL, 6, dst=ol.BUF_RES),
           ol.Op(ol.OP_RESTRICT, 6, dst=ol.BUF_RHS, src=ol.BUF_RES),
           ol.Op(ol.OP_ZERO, 5),
           ol.Op(ol.OP_SMOOTH, 5, mode=ol.MODE_REDBLACK, omega=1.0, unknowns=coll),
           ol.Op(ol.OP_PROLONG_ADD, 6, omega=0.9)]
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.build_program(prob, ops))
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _fields_equal(gc, oc, prob, [5, 6], exact=False)


def test_elasticity_default_solver(cuda_backend, oracle_mod):
    prob = problems.LinearElasticity2D(3, 6)
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.default_solver_cycle(prob))
    _assert_solve_parity(gc, oc, prob)


def test_elasticity_block_jacobi(cuda_backend, oracle_mod):
    prob = problems.LinearElasticity2D(3, 5)
    unk = ((0, (0, 0)), (0, (1, 0)), (1, (0, 0)), (1, (0, 1)))      # block shapes ((2,1),(1,2))
    ops = [ol.Op(ol.OP_SMOOTH, 5, mode=ol.MODE_JACOBI, omega=0.6, unknowns=unk),
           ol.Op(ol.OP_RESIDUAL, 5, dst=ol.BUF_RES)]
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.build_program(prob, ops))
    gc.apply(2)
    oc.apply(2)
    _fields_equal(gc, oc, prob, [5], exact=False)


@pytest.mark.parametrize("red_black", [True, False])
def test_poisson3d_statements_bit_exact(cuda_backend, oracle_mod, red_black):
    prob = problems.Poisson3D(2, 5)
    prog = _no_cg(cycles.v_cycle(prob, 2, 1, 1.25 if red_black else 0.8, red_black))
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _fields_equal(gc, oc, prob, range(2, 6))


def test_poisson3d_default_solver_and_w_cycle(cuda_backend, oracle_mod):
    prob = problems.Poisson3D(2, 5)
    for prog in (cycles.default_solver_cycle(prob), cycles.w_cycle(prob, 1, 1, 1.0, True)):
        gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, prog)
        _assert_solve_parity(gc, oc, prob)


def test_fused_residual_restrict(cuda_backend, oracle_mod):
    for prob in (problems.Poisson2D(3, 6), problems.Poisson3D(2, 4), problems.LinearElasticity2D(3, 5)):
        l = prob.max_level
        z = (0,) * prob.dim
        unk = tuple((f, z) for f in range(prob.n_fields))
        ops = [ol.Op(ol.OP_SMOOTH, l, mode=ol.MODE_JACOBI, omega=0.8, unknowns=unk),
               ol.Op(ol.OP_RESIDUAL_RESTRICT, l, dst=ol.BUF_RHS, src=ol.BUF_RES)]
        gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.build_program(prob, ops))
        gc.apply(1)
        oc.apply(1)
        _fields_equal(gc, oc, prob, [l - 1], bufs=(ol.BUF_RHS,))


@pytest.mark.parametrize("max_level", [5, 6])
def test_fused_residual_restrict_streaming_kernel(cuda_backend, oracle_mod, max_level):
    """3-D 7-point fast path (n >= 33): fine residual planes staged in shared memory, never written."""
    prob = problems.Poisson3D(2, max_level)
    z = (0, 0, 0)
    ops = []
    for l in range(max_level, 4, -1):
        ops += [ol.Op(ol.OP_SMOOTH, l, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=((0, z),)),
                ol.Op(ol.OP_RESIDUAL_RESTRICT, l, dst=ol.BUF_RHS, src=ol.BUF_RES)]
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.build_program(prob, ops))
    gc.apply(1)
    oc.apply(1)
    _fields_equal(gc, oc, prob, range(4, max_level), bufs=(ol.BUF_RHS,))
    assert np.abs(gc.get_field(4, ol.BUF_RHS)).max() > 0


def test_fused_cycle_gives_the_same_fitness(cuda_backend, oracle_mod):
    from evostencils_b200 import lowering
    prob = problems.Poisson3D(2, 6)
    plain = cycles.default_solver_cycle(prob)
    fused = lowering.optimise(plain)
    assert any(o.code == ol.OP_RESIDUAL_RESTRICT for o in fused.ops)
    dev = cuda_backend.DeviceProblem(prob)
    s = prob.settings
    a = dev.build(plain).solve(s.tol, s.max_iters, 1)
    b = dev.build(fused).solve(s.tol, s.max_iters, 1)
    assert a.iterations == b.iterations and np.array_equal(a.residuals, b.residuals)
    ref = oracle_mod.OracleProblem(prob).build(fused).solve(s.tol, s.max_iters, 1)
    assert np.array_equal(b.residuals, ref.residuals)


def test_divergent_cycle_reports_like_reference(cuda_backend, oracle_mod):
    """Over-relaxed Jacobi diverges: same iteration count / status handling on both paths."""
    prob = problems.Poisson2D(3, 5)
    prog = cycles.v_cycle(prob, 2, 2, 1.9, False)
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, prog)
    a = gc.solve(1e-12, 100, 1)
    b = oc.solve(1e-12, 100, 1)
    assert a.iterations == b.iterations and a.status == b.status
    fa = fitness.fitness_from_history(a.residuals, a.time_ms, 100)
    fb = fitness.fitness_from_history(b.residuals, b.time_ms, 100)
    assert fa[2] == fb[2]
    assert fa[1] > 1 and fb[1] > 1
    assert abs(fa[1] - fb[1]) <= 1e-6 * fb[1]


def test_batch_solve_matches_single(cuda_backend, oracle_mod):
    prob = problems.Poisson2D(3, 6)
    dev = cuda_backend.DeviceProblem(prob)
    progs = [cycles.v_cycle(prob, p, q, w, rb) for (p, q, w, rb) in
             ((2, 1, 1.15, True), (1, 1, 1.0, True), (2, 2, 0.8, False), (3, 3, 0.7, False), (1, 0, 1.0, True))]
    cyc = [dev.build(p) for p in progs]
    singles = [c.solve(1e-12, 100, 1) for c in cyc]
    outs, ms = dev.batch_solve(cyc, 1e-12, 100, samples=2)
    assert ms > 0
    for s, o in zip(singles, outs):
        assert s.iterations == o.iterations
        assert np.array_equal(s.residuals, o.residuals)


@pytest.mark.parametrize("level,sweeps", [(5, 1), (5, 2), (6, 3), (7, 2), (7, 1)])
def test_poisson3d_streaming_rbgs_bit_exact(cuda_backend, oracle_mod, level, sweeps):
    """The default 3-D RB-GS path (register-carried pair-column kernel k3_rbgs_col: both colours of ONE sweep per launch,
    z-slabs, XY tiles with halo recomputation; consecutive sweeps are separate launches by default -- the two-sweeps-per-
    launch kernel is covered by tests/test_gpu_variants.py::test_rbgs_two_sweeps_per_launch) against the plain
    colour-by-colour loops of the oracle: bit-identical."""
    prob = problems.Poisson3D(level - 1, level)
    z = (0, 0, 0)
    # start from a non-trivial state: one Jacobi sweep first, then `sweeps` RB-GS sweeps, then the residual
    ops = [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.9, unknowns=((0, z),))]
    ops += [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=((0, z),)) for _ in range(sweeps)]
    ops += [ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES)]
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.build_program(prob, ops))
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _fields_equal(gc, oc, prob, [level], bufs=(ol.BUF_SOL, ol.BUF_RES))


def test_poisson3d_129_default_solver(cuda_backend, oracle_mod):
    prob = problems.Poisson3D(2, 7)
    gc, oc, *_ = _pair(cuda_backend, oracle_mod, prob, cycles.default_solver_cycle(prob))
    _assert_solve_parity(gc, oc, prob)
