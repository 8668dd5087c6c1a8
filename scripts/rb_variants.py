#!/usr/bin/env python
"""Time every 3-D RB-GS kernel variant (EVO_RB_VARIANT) on one finest-level sweep; CUDA events through
evo_cycle_profile_op.  Usage: rb_variants.py [level] [variants...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evostencils_b200 import backend, cycles, oplist as ol, problems  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    variants = [int(v) for v in sys.argv[2:]] or [10, 30, 31, 32, 33, 34, 35, 36, 37]
    prob = problems.Poisson3D(2, level)
    peak = 6554.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    unk = ((0, (0, 0, 0)),)
    prog = cycles.build_program(prob, [ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.0, unknowns=unk)])
    cyc = backend.DeviceProblem(prob).build(prog)
    ndof = float((prob.nodes(level) - 2) ** 3)
    for v in variants:
        backend.set_option("EVO_RB_VARIANT", v)
        for sweeps in (1, 2):
            backend.set_option("EVO_RB_FUSE2", 1 if sweeps == 2 else 0)
            op = ol.Op(ol.OP_SMOOTH, level, mode=ol.MODE_REDBLACK, omega=1.25, unknowns=unk, count=sweeps)
            try:
                ms, n = cyc.profile_op(op, repeat=10)
            except backend.BackendError as e:
                print(f"variant {v} x{sweeps}: error {e}")
                continue
            gbs = 24.0 * ndof / (ms * 1e-3) / 1e9
            print(f"variant {v:3d} sweeps {sweeps} (fuse2={sweeps == 2}): {ms:7.3f} ms {n} launches  {gbs:7.1f} GB/s per launch-set "
                  f"= {100 * gbs / peak:5.1f}% of peak (per sweep: {100 * gbs * sweeps / peak:5.1f}%)  {sweeps * ndof / (ms * 1e-3) / 1e9:6.1f} GDOF/s",
                  flush=True)
    backend.set_option("EVO_RB_VARIANT", 0)
    backend.set_option("EVO_RB_FUSE2", 0)


if __name__ == "__main__":
    main()
