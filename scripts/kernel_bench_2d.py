#!/usr/bin/env python
"""Per-statement timing of the finest-level 2-D kernels at 4097^2 (Poisson 2D and FAS_2D_Basic), CUDA events via
evo_cycle_profile_op.  Usage: kernel_bench_2d.py [level]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evostencils_b200 import backend, cycles, oplist as ol, problems  # noqa: E402


def run(prob, prog, rows, peak):
    cyc = backend.DeviceProblem(prob).build(prog)
    level = prob.max_level
    ndof = float((prob.nodes(level) - 2) ** 2)
    print(f"{prob.name} level {level}: {ndof / 1e6:.1f} MDOF, peak {peak} GB/s")
    for name, op, bpd in rows:
        try:
            ms, n = cyc.profile_op(op, repeat=10)
        except backend.BackendError as e:
            print(f"  {name:22s} error: {e}")
            continue
        gbs = bpd * ndof / (ms * 1e-3) / 1e9
        print(f"  {name:22s} {ms:8.3f} ms  {n} launches  {gbs:8.1f} GB/s alg  {100 * gbs / peak:5.1f}% of peak  {ndof / (ms * 1e-3) / 1e9:7.1f} GDOF/s",
              flush=True)
    cyc.close()


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    peak = 6554.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    unk = ((0, (0, 0)),)
    prob = problems.Poisson2D(level - 4, level)
    prog = cycles.build_program(prob, [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.8, unknowns=unk),
                                       ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.0, unknowns=unk)])
    run(prob, prog, [
        ("rbgs x1", ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.15, unknowns=unk, count=1), 24.0),
        ("rbgs x2", ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.15, unknowns=unk, count=2), 48.0),
        ("jacobi x1", ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.8, unknowns=unk, count=1), 24.0),
        ("jacobi x2", ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.8, unknowns=unk, count=2), 48.0),
        ("residual", ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES), 24.0),
        ("restrict", ol.Op(ol.OP_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES), 10.0),
        ("residual+restrict", ol.Op(ol.OP_RESIDUAL_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES), 18.0),
        ("prolong_add", ol.Op(ol.OP_PROLONG_ADD, level, src=ol.BUF_SOL, omega=1.0), 18.0),
    ], peak)
    fas = problems.FAS2D(level - 4, level)
    fprog = cycles.fas_v_cycle(fas)
    nj = [o for o in fprog.ops if o.code == ol.OP_SMOOTH and o.level == level][0]
    run(fas, fprog, [
        ("newton-jacobi x1", nj, 24.0),
        ("fas residual", ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES), 24.0),
        ("fas restrict sol", ol.Op(ol.OP_FAS_RESTRICT_SOL, level, dst=ol.BUF_APX, src=ol.BUF_SOL), 10.0 + 4.0),
        ("fas coarse rhs", ol.Op(ol.OP_FAS_COARSE_RHS, level, dst=ol.BUF_RHS, src=ol.BUF_RES), 10.0 + 6.0),
        ("prolong_add", ol.Op(ol.OP_PROLONG_ADD, level, src=ol.BUF_SOL, omega=1.0), 18.0),
    ], peak)


if __name__ == "__main__":
    main()
