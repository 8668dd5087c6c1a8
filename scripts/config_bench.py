#!/usr/bin/env python
"""BASELINE.json configs[2] (FAS_2D_Basic at 4097^2) and configs[3] (Helmholtz 2D, MG-preconditioned BiCGStab)
on one GPU, CUDA-event timed through the C-ABI; optional CPU-port time beside it (--cpu).  One JSON line each."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evostencils_b200 import backend, cycles, fitness, oplist as ol, problems  # noqa: E402


def fas(level, cpu):
    prob = problems.FAS2D(level - 4, level)
    prog = cycles.fas_v_cycle(prob)
    s = prob.settings
    dev = backend.DeviceProblem(prob)
    cyc = dev.build(prog)
    cyc.solve(s.tol, s.max_iters, 1)
    out = cyc.solve(s.tol, s.max_iters, 3)
    ndof = float((prob.nodes(level) - 2) ** 2)
    line = {"config": f"FAS_2D_Basic {prob.nodes(level)}^2, levels {level}..{level - 4}, FAS V(2,2) Newton-Jacobi w=0.8",
            "iterations": out.iterations, "ms_per_eval": out.time_ms, "ms_per_cycle": out.time_ms / max(out.iterations, 1),
            "evals_per_s": 1e3 / out.time_ms, "final_rel_residual": out.final_residual / out.initial_residual,
            "cycle_gdof_s": ndof * out.iterations / (out.time_ms * 1e-3) / 1e9, "kernel_launches": out.kernel_launches}
    if cpu:
        from oracle import oracle as orc
        small = problems.FAS2D(level - 4 - 2, level - 2) if level > 10 else prob
        oc = orc.OracleProblem(small).build(cycles.fas_v_cycle(small))
        oc.apply(1)
        t0 = time.perf_counter(); oc.apply(1); oc.residual_norm(); t = time.perf_counter() - t0
        scale = ndof / float((small.nodes(small.max_level) - 2) ** 2)
        line["cpu_port"] = {"threads": orc.num_threads(), "ms_per_cycle_scaled": t * 1e3 * scale,
                            "sample": f"1 cycle at {small.nodes(small.max_level)}^2 x {scale:g}",
                            "evals_per_s": 1.0 / (t * scale * out.iterations)}
    print(json.dumps(line), flush=True)


def helmholtz(level, k, cpu):
    prob = problems.Helmholtz2D(level - 4, level, k=k)
    prog = cycles.default_solver_cycle(prob)
    s = prob.settings
    dev = backend.DeviceProblem(prob)
    cyc = dev.build(prog)
    cyc.helmholtz_solve(s.tol, s.max_iters, 1)
    out = cyc.helmholtz_solve(s.tol, s.max_iters, 3)
    line = {"config": f"Helmholtz 2D {prob.nodes(level)}^2 k={k}, V(2,1) RB-GS w=0.6 on the shifted operator inside BiCGStab",
            "outer_iterations": out.iterations, "ms_per_eval": out.time_ms, "evals_per_s": 1e3 / out.time_ms,
            "ms_per_outer_iteration": out.time_ms / max(out.iterations, 1), "kernel_launches": out.kernel_launches,
            "final_rel_residual": out.final_residual / out.initial_residual}
    if cpu:
        from oracle import oracle as orc
        oc = orc.OracleProblem(prob).build(prog)
        t0 = time.perf_counter(); ref = oc.helmholtz_solve(s.tol, s.max_iters, 1); t = time.perf_counter() - t0
        line["cpu_port"] = {"threads": orc.num_threads(), "ms_per_eval": t * 1e3, "outer_iterations": ref.iterations,
                            "history_identical": bool((ref.residuals == out.residuals).all())}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--fas-level", type=int, default=12)
    ap.add_argument("--fas-only", action="store_true")
    args = ap.parse_args()
    fas(args.fas_level, args.cpu)
    fas(10, args.cpu)
    if not args.fas_only:
        for lvl, k in ((7, 80.0), (8, 160.0), (9, 320.0)):
            helmholtz(lvl, k, args.cpu)
