"""An independent check of the oracle (test infrastructure checking test infrastructure): the discrete equations are
assembled with scipy.sparse straight from the problem descriptions (stencil tables, boundary / right-hand-side
functions) -- no code shared with oracle/*.c -- and the oracle's converged solutions must satisfy them; for the
linear problems the oracle's solution must also agree with a sparse direct solve.  This pins operator application,
Dirichlet handling, right-hand sides and the coupled 2x2 system for the problems the reference has no fixture for."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from evostencils_b200 import cycles, oplist as ol, problems


def assemble(prob, level):
    """(A, b, u_boundary): A over the inner unknowns of all fields (field-major), b = f - A_boundary * g."""
    n, dim, nf = prob.nodes(level), prob.dim, prob.n_fields
    ni = n - 2
    m = ni ** dim
    table = prob.operator(level)
    idx = np.arange(m).reshape((ni,) * dim)
    init = [prob.initial_solution(f, level) if level == prob.max_level else np.zeros((n,) * dim, dtype=prob.dtype)
            for f in range(nf)]
    rhs = [prob.rhs(f, level) for f in range(nf)]
    rows, cols, vals = [], [], []
    b = np.concatenate([r[(slice(1, -1),) * dim].reshape(-1) for r in rhs]).astype(prob.dtype)
    inner = (slice(1, -1),) * dim
    for a in range(nf):
        for j in range(nf):
            for p in range(ol.STENCIL_POINTS):
                cf = table[a, j, p]
                if cf == 0:
                    continue
                off = ol.stencil_offset(p, dim)              # (dx, dy[, dz])
                shift = tuple(reversed(off))                  # array axes are ([z,] y, x)
                # neighbour coordinates of every inner node
                src = tuple(slice(1 + s, n - 1 + s) for s in shift)
                coords = np.meshgrid(*[np.arange(1, n - 1) + s for s in shift], indexing="ij")
                inside = np.ones((ni,) * dim, dtype=bool)
                for cgrid in coords:
                    inside &= (cgrid >= 1) & (cgrid <= n - 2)
                r_idx = idx[inside] + a * m
                c_idx = idx[tuple(cg[inside] - 1 for cg in coords)] + j * m
                rows.append(r_idx); cols.append(c_idx); vals.append(np.full(r_idx.shape, cf, dtype=prob.dtype))
                # boundary neighbours move to the right-hand side
                bvals = init[j][src]
                contrib = np.where(inside, 0.0, cf * bvals)
                b[a * m:(a + 1) * m] -= contrib.reshape(-1)
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nf * m, nf * m))
    return A, b


def oracle_solution(oracle_mod, prob, prog, tol=1e-12, iters=100):
    cyc = oracle_mod.OracleProblem(prob).build(prog)
    out = cyc.solve(tol, iters, 1, ol.SOLVE_KEEP_STATE) if False else cyc.solve(tol, iters, 1)
    inner = (slice(1, -1),) * prob.dim
    u = np.concatenate([cyc.get_field(prob.max_level, ol.BUF_SOL, f)[inner].reshape(-1) for f in range(prob.n_fields)])
    return out, u


@pytest.mark.parametrize("prob", [problems.Poisson2D(2, 5), problems.Poisson3D(2, 4), problems.LinearElasticity2D(2, 5)],
                         ids=["poisson2d", "poisson3d", "elasticity2d"])
def test_linear_problems_against_a_sparse_direct_solve(oracle_mod, prob):
    prog = cycles.default_solver_cycle(prob)
    out, u = oracle_solution(oracle_mod, prob, prog)
    assert out.residuals[-1] < 1e-11 * out.residuals[0]
    A, b = assemble(prob, prob.max_level)
    direct = spla.spsolve(A.tocsc(), b)
    scale = np.abs(direct).max()
    assert np.abs(u - direct).max() < 1e-9 * scale
    # the oracle's residual norm is the norm of the independently assembled residual
    assert abs(np.linalg.norm(b - A @ u) - out.residuals[-1]) <= 1e-6 * out.residuals[0]
    # ... and so is the initial one (u = 0 inside): pins right-hand side + boundary handling
    assert abs(np.linalg.norm(b) - out.residuals[0]) <= 1e-12 * out.residuals[0]


def test_fas_solution_satisfies_the_nonlinear_equations(oracle_mod):
    prob = problems.FAS2D(2, 5)
    out, u = oracle_solution(oracle_mod, prob, cycles.fas_v_cycle(prob), prob.settings.tol, prob.settings.max_iters)
    A, b = assemble(prob, prob.max_level)
    F = A @ u + prob.gamma * u * np.exp(u) - b
    assert np.linalg.norm(F) < 1e-9 * np.linalg.norm(b)
    assert abs(np.linalg.norm(b) - out.residuals[0]) <= 1e-12 * out.residuals[0]      # u0 = 0: F(0) = -f
    # second-order accurate against the manufactured solution
    n = prob.nodes(prob.max_level)
    x = np.linspace(0, 1, n)
    exact = prob.exact_solution(x[None, 1:-1], x[1:-1, None]).reshape(-1)
    assert np.abs(u - exact).max() < 5e-3


def test_helmholtz_solution_satisfies_independently_assembled_equations(oracle_mod):
    """Outer operator -Lap_h - k^2 with the Robin closure u_b = u_inner / (1 - i k h) on the x boundaries and u = 0 on the
    y boundaries (Helmholtz exa4:43-108), assembled here from the stencil table; the oracle's BiCGStab solution (stop
    1e-7) must satisfy it and agree with a sparse direct solve."""
    prob = problems.Helmholtz2D(3, 6, k=40.0)
    level = prob.max_level
    n, ni = prob.nodes(level), prob.nodes(level) - 2
    h = prob.spacing(level)
    table = prob.outer_operator(level)
    rden = 1.0 / (1.0 - 1j * prob.wave_number * h)
    idx = np.arange(ni * ni).reshape(ni, ni)                 # [y-1, x-1]
    A = sp.lil_matrix((ni * ni, ni * ni), dtype=np.complex128)
    for p in range(ol.STENCIL_POINTS):
        cf = table[0, 0, p]
        if cf == 0:
            continue
        dx, dy = ol.stencil_offset(p, 2)
        for y in range(1, n - 1):
            for x in range(1, n - 1):
                xx, yy = x + dx, y + dy
                if yy < 1 or yy > n - 2:
                    continue                                  # Dirichlet 0
                if xx < 1 or xx > n - 2:
                    A[idx[y - 1, x - 1], idx[yy - 1, x - 1]] += cf * rden     # Robin: boundary value = rden * inner neighbour
                else:
                    A[idx[y - 1, x - 1], idx[yy - 1, xx - 1]] += cf
    b = prob.rhs(0, level)[1:-1, 1:-1].reshape(-1).astype(np.complex128)
    cyc = oracle_mod.OracleProblem(prob).build(cycles.default_solver_cycle(prob))
    out = cyc.helmholtz_solve(prob.settings.tol, prob.settings.max_iters, 1)
    u = cyc.get_field(level, ol.BUF_COR)[1:-1, 1:-1].reshape(-1)      # the outer solution is left in COR@finest
    A = A.tocsc()
    assert out.residuals[-1] < prob.settings.tol * out.residuals[0]
    assert abs(np.linalg.norm(b) - out.residuals[0]) <= 1e-12 * out.residuals[0]
    assert np.linalg.norm(b - A @ u) < 5e-7 * np.linalg.norm(b)
    direct = spla.spsolve(A, b)
    assert np.abs(u - direct).max() < 1e-4 * np.abs(direct).max()
