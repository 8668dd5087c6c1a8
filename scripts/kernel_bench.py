#!/usr/bin/env python
"""Per-statement timing of the finest-level kernels (CUDA events via evo_cycle_profile_op)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evostencils_b200 import backend, cycles, oplist as ol, problems  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    prob = problems.Poisson3D(2, level) if dim == 3 else problems.Poisson2D(4, level)
    peak = 6554.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    z = (0,) * dim
    unk = ((0, z),)
    # a cycle that owns every buffer the statements below touch (jacobi slot included)
    prog = cycles.build_program(prob, [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.8, unknowns=unk),
                                       ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.0, unknowns=unk)])
    cyc = backend.DeviceProblem(prob).build(prog)
    ndof = float((prob.nodes(level) - 2) ** dim)
    rows = [
        ("rbgs x1", ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=unk, count=1), 24.0, 1),
        ("rbgs x2", ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=unk, count=2), 24.0, 2),
        ("jacobi x1", ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_JACOBI, omega=0.8, unknowns=unk, count=1), 24.0, 1),
        ("residual", ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES), 24.0, 1),
        ("restrict", ol.Op(ol.OP_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES), 8 + 8 / 2 ** dim, 1),
        ("residual+restrict", ol.Op(ol.OP_RESIDUAL_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES), 16 + 8 / 2 ** dim, 1),
        ("prolong_add", ol.Op(ol.OP_PROLONG_ADD, level, src=ol.BUF_SOL, omega=1.0), 16 + 8 / 2 ** dim, 1),
    ]
    print(f"level {level} dim {dim}: {ndof / 1e6:.1f} MDOF, peak {peak} GB/s")
    for name, op, bpd, sweeps in rows:
        try:
            ms, n = cyc.profile_op(op, repeat=10)
        except backend.BackendError as e:
            print(f"{name:20s} error: {e}")
            continue
        gbs = bpd * ndof / (ms * 1e-3) / 1e9
        print(f"{name:20s} {ms:8.3f} ms  {n} launches  {gbs:8.1f} GB/s alg  {100 * gbs / peak:5.1f}% of peak  "
              f"{sweeps * ndof / (ms * 1e-3) / 1e9:7.1f} GDOF/s")


if __name__ == "__main__":
    main()
