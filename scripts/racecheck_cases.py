#!/usr/bin/env python
"""Small invocations of every shared-memory kernel of the hot path, to be run under
`compute-sanitizer --tool racecheck`.  Each case is also checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evostencils_b200 import backend, cycles, lowering, oplist as ol, problems  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def check(name, prob, prog, applies=1, solve=False):
    gc = backend.DeviceProblem(prob).build(prog)
    oc = orc.OracleProblem(prob).build(prog)
    if solve:
        a = gc.solve(prob.settings.tol, 3, 1, ol.SOLVE_NO_GRAPH)
        b = oc.solve(prob.settings.tol, 3, 1)
        ok = a.iterations == b.iterations and np.array_equal(a.residuals, b.residuals)
    else:
        gc.apply(applies)
        oc.apply(applies)
        ok = all(np.array_equal(gc.get_field(prob.max_level, ol.BUF_SOL, f), oc.get_field(prob.max_level, ol.BUF_SOL, f))
                 for f in range(prob.n_fields))
    print(f"{name}: {'identical to the oracle' if ok else 'MISMATCH'}", flush=True)
    gc.close()
    return ok


def main():
    ok = True
    z3 = (0, 0, 0)
    # k3_rbgs_col (TMA ring, first colour through shared memory) + k3_residual_restrict_tma + k_coarse_cg_smem
    p3 = problems.Poisson3D(2, 5)
    ok &= check("3-D V(2,1) RB-GS cycle, 33^3 (k3_rbgs_col, k3_residual_restrict_tma, k_coarse_cg_smem)", p3,
                lowering.optimise(cycles.default_solver_cycle(p3)), solve=True)
    for variant in (10, 20):
        backend.set_option("EVO_RB_VARIANT", variant)
        ok &= check(f"3-D RB-GS sweep, variant {variant} (k3_rbgs_lean / k3_rbgs_stream)", p3, cycles.build_program(
            p3, [ol.Op(ol.OP_SMOOTH, 5, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=((0, z3),))]), applies=2)
    backend.set_option("EVO_RB_VARIANT", 0)
    # order-dependent coloured sweep of the elasticity system (k2_smooth_rowseq_pipe, cp.async window) + register CG
    pe = problems.LinearElasticity2D(3, 5)
    ok &= check("elasticity V(2,1) collective RB-GS, 33^2 (k2_smooth_rowseq_pipe, k2_coarse_cg_reg)", pe, cycles.default_solver_cycle(pe),
                solve=True)
    # FAS coarse solver in a thread-block cluster (distributed shared memory halos)
    pf = problems.FAS2D(5, 7)
    ok &= check("FAS V(2,2), coarsest 33^2 (k2_fas_coarse_cluster)", pf, cycles.fas_v_cycle(pf), applies=1)
    # Helmholtz coarse BiCGStab
    ph = problems.Helmholtz2D(3, 5, k=20.0)
    gc = backend.DeviceProblem(ph).build(cycles.default_solver_cycle(ph))
    a = gc.helmholtz_solve(ph.settings.tol, 5, 1)
    b = orc.OracleProblem(ph).build(cycles.default_solver_cycle(ph)).helmholtz_solve(ph.settings.tol, 5, 1)
    same = np.array_equal(a.residuals, b.residuals)
    print(f"Helmholtz outer BiCGStab, 5 iterations (k2_coarse_bicgstab): {'identical to the oracle' if same else 'MISMATCH'}", flush=True)
    ok &= same
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
