"""Kernel variants and modes that the default configuration does not select: every 3-D RB-GS kernel (lean tile
kernel, register-carried pair-column kernel, generic multi-stage kernel, two sweeps fused per launch) and the
lexicographic in-place sweeps of the reference's model-based mode (exastencils.py:64-70, :781-822) -- all
bit-identical to the plain loops of the oracle."""
import numpy as np
import pytest

from evostencils_b200 import cycles, lowering, oplist as ol, problems

pytestmark = pytest.mark.gpu


@pytest.fixture
def option(cuda_backend):
    """Set tuning switches of the library for one test and restore them afterwards."""
    saved = {}

    def setter(name, value):
        if name not in saved:
            saved[name] = cuda_backend.get_option(name)
        cuda_backend.set_option(name, value)

    yield setter
    for name, value in saved.items():
        cuda_backend.set_option(name, value)


def _equal(gc, oc, prob, levels, bufs):
    for l in levels:
        for b in bufs:
            for f in range(prob.n_fields):
                x, y = gc.get_field(l, b, f), oc.get_field(l, b, f)
                assert np.array_equal(x, y), f"level {l} buf {b} field {f}: max diff {np.abs(x - y).max()}"


def test_option_api(cuda_backend, option):
    assert cuda_backend.get_option("EVO_RB_FUSE2") in (0, 1)
    option("EVO_RB_VARIANT", 30)
    assert cuda_backend.get_option("EVO_RB_VARIANT") == 30
    with pytest.raises(cuda_backend.BackendError):
        cuda_backend.set_option("EVO_NO_SUCH_SWITCH", 1)


# variants: 0 default, 10 lean, 20 generic multi-stage, 30-35 pair-column kernels
@pytest.mark.parametrize("variant", [0, 10, 20, 30, 31, 32, 33, 34, 35, 36, 37])
@pytest.mark.parametrize("level,sweeps", [(5, 1), (6, 3), (7, 2)])
def test_rbgs_kernel_variants_bit_exact(cuda_backend, oracle_mod, option, variant, level, sweeps):
    option("EVO_RB_VARIANT", variant)
    prob = problems.Poisson3D(level - 1, level)
    z = (0, 0, 0)
    ops = [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.9, unknowns=((0, z),))]
    ops += [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=((0, z),)) for _ in range(sweeps)]
    ops += [ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES)]
    prog = cycles.build_program(prob, ops)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, [level], (ol.BUF_SOL, ol.BUF_RES))


@pytest.mark.parametrize("variant", [0, 30])
@pytest.mark.parametrize("fuse2", [0, 1])
def test_rbgs_two_sweeps_per_launch(cuda_backend, oracle_mod, option, variant, fuse2):
    """EVO_RB_FUSE2: consecutive identical sweeps are executed two per launch (temporal blocking); the solve
    history must not change."""
    option("EVO_RB_VARIANT", variant)
    option("EVO_RB_FUSE2", fuse2)
    prob = problems.Poisson3D(2, 6)
    prog = cycles.default_solver_cycle(prob)       # V(2,1): the two pre-smoothing sweeps are merged
    s = prob.settings
    a = cuda_backend.DeviceProblem(prob).build(prog).solve(s.tol, s.max_iters, 1)
    b = oracle_mod.OracleProblem(prob).build(prog).solve(s.tol, s.max_iters, 1)
    assert a.iterations == b.iterations
    assert np.array_equal(a.residuals, b.residuals)


# ---- lexicographic in-place sweeps ---------------------------------------------------------------------------
def _lex_ops(level, unknowns, n=2, omega=0.9):
    return [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_LEX, omega=omega, unknowns=unknowns) for _ in range(n)] + \
           [ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES)]


@pytest.mark.parametrize("lex_variant", [0, 1])
@pytest.mark.parametrize("name,level", [("p2", 5), ("p2", 7), ("p3", 4), ("p3", 5), ("el", 5)])
def test_lexicographic_pointwise_bit_exact(cuda_backend, oracle_mod, option, name, level, lex_variant):
    option("EVO_LEX_VARIANT", lex_variant)
    prob = {"p2": problems.Poisson2D, "p3": problems.Poisson3D, "el": problems.LinearElasticity2D}[name](level - 1, level)
    z = (0,) * prob.dim
    if prob.n_fields == 1:
        ops = _lex_ops(level, ((0, z),))
    else:   # decoupled (one statement per field) and collective
        ops = [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_LEX, omega=0.8, unknowns=((0, z),)),
               ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_LEX, omega=0.8, unknowns=((1, z),))] + _lex_ops(level, ((0, z), (1, z)))
    prog = cycles.build_program(prob, ops)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, [level], (ol.BUF_SOL, ol.BUF_RES))


@pytest.mark.parametrize("shape", [(2, 1), (1, 2), (2, 2), (3, 1), (1, 3)])
def test_lexicographic_block_bit_exact(cuda_backend, oracle_mod, shape):
    """Overlapping blocks written in place: the hyperplane skew has to respect the wider footprint."""
    prob = problems.Poisson2D(4, 5)
    unk = tuple((0, (i, j)) for i in range(shape[0]) for j in range(shape[1]))
    prog = cycles.build_program(prob, _lex_ops(5, unk, n=2, omega=0.7))
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    gc.apply(1)
    oc.apply(1)
    _equal(gc, oc, prob, [5], (ol.BUF_SOL, ol.BUF_RES))


def test_lexicographic_block_3d(cuda_backend, oracle_mod):
    prob = problems.Poisson3D(3, 4)
    unk = ((0, (0, 0, 0)), (0, (1, 0, 0)), (0, (0, 0, 1)))
    prog = cycles.build_program(prob, _lex_ops(4, unk, n=1, omega=0.8))
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    gc.apply(1)
    oc.apply(1)
    _equal(gc, oc, prob, [4], (ol.BUF_SOL, ol.BUF_RES))


def test_model_based_mode_through_the_drop_in(cuda_backend, oracle_mod):
    """model_based_estimation=True (the default of the reference's scripts/optimize.py) lowers every
    Single-partitioned smoother to an in-place lexicographic sweep: the drop-in must evaluate it (not refuse it)
    and agree with the oracle."""
    from evostencils_b200 import fitness, tree
    from evostencils_b200.program_generator import B200ProgramGenerator
    prob = problems.Poisson2D(3, 6)
    gen = B200ProgramGenerator(problem=prob, model_based_estimation=True)
    expr = tree.build_tree(prob, tree.v_cycle_individual(prob.max_level - prob.min_level, pre=2, post=1, partitioning="single"))
    prog = gen._finalise(gen.lower(expr, prob.min_level))
    assert any(o.code == ol.OP_SMOOTH and o.mode == ol.MODE_LEX for o in prog.ops)
    t, cf, its = gen.generate_and_evaluate(expr, gen.generate_storage(3, 6), 3, 6, "", evaluation_samples=1)
    ref = oracle_mod.OracleProblem(prob).build(prog).solve(prob.settings.tol, prob.settings.max_iters, 1)
    rt, rcf, rits = fitness.fitness_from_history(ref.residuals, ref.time_ms, prob.settings.max_iters)
    assert its == rits and abs(cf - rcf) < 1e-12 and cf < 1
    assert np.array_equal(gen.last_outcome.residuals, ref.residuals)
    gen.close()


# ---- fused runs of statements on the small levels (evo_kernels_run.cuh) -------------------------------------------
@pytest.mark.parametrize("mode", [1, 2, 3, 5])
@pytest.mark.parametrize("name", ["p2", "el", "p3"])
def test_fused_runs_equal_single_launches(cuda_backend, oracle_mod, option, name, mode):
    """EVO_COARSE_FUSE: maximal runs of statements on small levels -- coarse-grid CG included -- are interpreted by one
    kernel (one CTA or a cluster); the residual history must be bit-identical to one launch per statement (and to the
    oracle), with fewer launches.  Modes: 1 one CTA, 2 clusters on the larger levels, 3 smallest levels, 5 field arrays resident in shared memory."""
    import random
    from evostencils_b200 import lowering, tree
    prob = {"p2": problems.Poisson2D(3, 8), "el": problems.LinearElasticity2D(3, 7), "p3": problems.Poisson3D(2, 5)}[name]
    rng = random.Random(5)
    progs = [cycles.default_solver_cycle(prob), lowering.optimise(cycles.default_solver_cycle(prob)),
             lowering.optimise(cycles.w_cycle(prob, 2, 1, 1.1, True))]
    for _ in range(5):
        s = tree.random_individual(prob, rng, maximum_local_system_size=4)
        progs.append(lowering.optimise(lowering.lower_cycle(tree.build_tree(prob, s), prob.min_level, prob.max_level, prob.n_fields,
                                                            prob.dim, cgs_max_iters=prob.settings.cgs_max_iters,
                                                            cgs_tol=prob.settings.cgs_tol,
                                                            default_restrict=prob.restrict_weights(),
                                                            default_prolong=prob.prolong_weights())))
    dev = cuda_backend.DeviceProblem(prob)
    ref = oracle_mod.OracleProblem(prob)
    st = prob.settings
    fewer = 0
    for prog in progs:
        option("EVO_COARSE_FUSE", 0)
        a = dev.build(prog).solve(st.tol, st.max_iters, 1)
        option("EVO_COARSE_FUSE", mode)
        b = dev.build(prog).solve(st.tol, st.max_iters, 1)
        c = dev.build(prog).solve(st.tol, st.max_iters, 1, ol.SOLVE_NO_GRAPH)
        o = ref.build(prog).solve(st.tol, st.max_iters, 1)
        assert a.iterations == b.iterations == c.iterations == o.iterations
        assert np.array_equal(a.residuals, b.residuals, equal_nan=True)
        assert np.array_equal(b.residuals, c.residuals, equal_nan=True)
        assert np.array_equal(b.residuals, o.residuals, equal_nan=True)
        fewer += b.kernel_launches < a.kernel_launches
    assert fewer >= len(progs) - 1


# ---- register-streamed 2-D sweeps (evo_kernels_warp2d.cuh) ---------------------------------------------------------
@pytest.mark.parametrize("star2d", [65, 1, 0])
@pytest.mark.parametrize("level,mode,sweeps", [(7, "rb", 1), (8, "rb", 2), (9, "rb", 3), (10, "rb", 2), (7, "jac", 1), (8, "jac", 2),
                                               (9, "jac", 3), (10, "jac", 1)])
def test_streamed_2d_sweeps_bit_exact(cuda_backend, oracle_mod, option, star2d, level, mode, sweeps):
    """Pointwise Jacobi / RB-GS on large 2-D grids: up to two consecutive sweeps per pass (temporal blocking), strips and
    row chunks with redundant halo work -- bit-identical to the plain loops of the oracle; EVO_STAR2D = 0: generic kernels,
    1: streaming from 513^2 (default), n: streaming from n^2."""
    option("EVO_STAR2D", star2d)
    prob = problems.Poisson2D(level - 1, level)
    z = (0, 0)
    m = ol.MODE_REDBLACK if mode == "rb" else ol.MODE_JACOBI
    ops = [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.9, unknowns=((0, z),))]
    ops += [ol.Op(ol.OP_SMOOTH, level, mode=m, omega=1.15 if mode == "rb" else 0.8, unknowns=((0, z),)) for _ in range(sweeps)]
    ops += [ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES)]
    prog = cycles.build_program(prob, ops)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, [level], (ol.BUF_SOL, ol.BUF_RES))


def test_streamed_2d_default_solver_1025(cuda_backend, oracle_mod):
    prob = problems.Poisson2D(5, 10)
    prog = cycles.default_solver_cycle(prob)
    s = prob.settings
    a = cuda_backend.DeviceProblem(prob).build(prog).solve(s.tol, s.max_iters, 1)
    b = oracle_mod.OracleProblem(prob).build(prog).solve(s.tol, s.max_iters, 1)
    assert a.iterations == b.iterations and np.array_equal(a.residuals, b.residuals)


@pytest.mark.parametrize("star2d", [65, 1, 0])
def test_streamed_fas_sweeps(cuda_backend, oracle_mod, option, star2d):
    """FAS Newton-Jacobi smoother and the 200-sweep coarse solver on a 129^2 coarsest grid (4 / 2 / 1 sweeps per launch)."""
    option("EVO_STAR2D", star2d)
    prob = problems.FAS2D(7, 9)
    prog = cycles.fas_v_cycle(prob)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, [7, 8, 9], (ol.BUF_SOL,))


@pytest.mark.parametrize("star2d", [65, 1])
@pytest.mark.parametrize("level", [7, 8, 9, 10])
def test_streamed_2d_transfers_bit_exact(cuda_backend, oracle_mod, option, star2d, level):
    """Fused residual + restriction and prolongation + correction of the 2-D streaming path against the oracle."""
    option("EVO_STAR2D", star2d)
    prob = problems.Poisson2D(level - 2, level)
    z = (0, 0)
    ops = [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.9, unknowns=((0, z),)),
           ol.Op(ol.OP_RESIDUAL_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES),
           ol.Op(ol.OP_ZERO, level - 1, dst=ol.BUF_SOL),
           ol.Op(ol.OP_SMOOTH, level - 1, mode=ol.MODE_REDBLACK, omega=1.1, unknowns=((0, z),)),
           ol.Op(ol.OP_PROLONG_ADD, level, src=ol.BUF_SOL, omega=0.95),
           ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES)]
    prog = cycles.build_program(prob, ops)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, [level], (ol.BUF_SOL, ol.BUF_RES))
        _equal(gc, oc, prob, [level - 1], (ol.BUF_SOL, ol.BUF_RHS))


# fused residual + restriction, 3-D: 0-4 shared-memory ring kernel (tile shapes), 10-17 register-carried column kernel
@pytest.mark.parametrize("variant", [0, 9, 1, 2, 3, 4, 10, 11, 12, 13, 14, 15, 16, 17])
@pytest.mark.parametrize("level", [6, 7])
def test_residual_restrict_3d_variants_bit_exact(cuda_backend, oracle_mod, option, variant, level):
    option("EVO_RR_VARIANT", variant)
    prob = problems.Poisson3D(level - 2, level)
    z = (0, 0, 0)
    ops = [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.9, unknowns=((0, z),)),
           ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.2, unknowns=((0, z),)),
           ol.Op(ol.OP_RESIDUAL_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES),
           ol.Op(ol.OP_ZERO, level - 1, dst=ol.BUF_SOL),
           ol.Op(ol.OP_SMOOTH, level - 1, mode=ol.MODE_REDBLACK, omega=1.1, unknowns=((0, z),)),
           ol.Op(ol.OP_PROLONG_ADD, level, src=ol.BUF_SOL, omega=0.95)]
    prog = cycles.build_program(prob, ops)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, [level], (ol.BUF_SOL,))
        _equal(gc, oc, prob, [level - 1], (ol.BUF_SOL, ol.BUF_RHS))


@pytest.mark.parametrize("name", ["p2", "el", "p3"])
@pytest.mark.parametrize("no_fuse", [0, 1])
def test_zero_folded_into_the_restriction(cuda_backend, oracle_mod, option, name, no_fuse):
    """`RHS@(l-1) = R (f - A u)` followed by `SOL@(l-1) = 0` (every coarse-grid correction, exastencils.py:698-716): the
    generic and the 2-D streaming restriction kernels store the zeros themselves (EVO_NO_ZERO_FUSE = 1: separate memset
    node).  Both ways must equal the oracle on every level after complete cycles."""
    option("EVO_NO_ZERO_FUSE", no_fuse)
    option("EVO_STAR2D", 65)
    prob = {"p2": problems.Poisson2D(3, 8), "el": problems.LinearElasticity2D(3, 6), "p3": problems.Poisson3D(2, 5)}[name]
    prog = lowering.optimise(cycles.v_cycle(prob, 2, 1, 1.1, True))
    assert any(o.code == ol.OP_RESIDUAL_RESTRICT for o in prog.ops) and any(o.code == ol.OP_ZERO for o in prog.ops)
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(3):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, range(prob.min_level, prob.max_level + 1), (ol.BUF_SOL, ol.BUF_RHS))


@pytest.mark.parametrize("name,lo,hi", [("p3", 2, 4), ("p3", 2, 3), ("p2", 2, 6), ("p2", 3, 5), ("el", 2, 5)])
def test_tiny_grid_kernels_bit_exact(cuda_backend, oracle_mod, name, lo, hi):
    """Grids up to 4096 inner nodes: all repetitions and both colours of a red-black sweep run in one single-CTA launch
    (k_smooth_rb_small), the fused residual+restriction uses one warp per coarse node (k_residual_restrict_warp).  W-cycle
    with three merged pre-smoothing sweeps against the oracle."""
    prob = {"p2": problems.Poisson2D, "el": problems.LinearElasticity2D, "p3": problems.Poisson3D}[name](lo, hi)
    prog = lowering.optimise(cycles.w_cycle(prob, 3, 2, 1.15, True))
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, range(prob.min_level, prob.max_level + 1), (ol.BUF_SOL, ol.BUF_RHS))


def test_zero_fold_is_dropped_after_a_user_write_to_a_coarse_solution(cuda_backend, oracle_mod):
    """The folded `SOL@(l-1) = 0` clears the inner nodes only (the boundary layer of a correction level is never written
    by a kernel).  After evo_cycle_set_field put values on that boundary layer the statement must clear the whole array
    again, like the oracle's memset -- also in a solver graph that was captured before the write."""
    prob = problems.Poisson2D(3, 6)
    prog = lowering.optimise(cycles.v_cycle(prob, 2, 1, 1.1, True))
    gc = cuda_backend.DeviceProblem(prob).build(prog)
    oc = oracle_mod.OracleProblem(prob).build(prog)
    a = gc.solve(1e-30, 2, 1)
    b = oc.solve(1e-30, 2, 1)
    assert np.array_equal(a.residuals, b.residuals)
    rng = np.random.default_rng(3)
    junk = rng.standard_normal(gc.get_field(4, ol.BUF_SOL, 0).shape)      # non-zero boundary layer included
    for c in (gc, oc):
        c.set_field(4, ol.BUF_SOL, 0, junk)
    for _ in range(2):
        gc.apply(1)
        oc.apply(1)
        _equal(gc, oc, prob, range(prob.min_level, prob.max_level + 1), (ol.BUF_SOL, ol.BUF_RHS))
