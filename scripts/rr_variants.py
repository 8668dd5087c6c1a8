#!/usr/bin/env python
"""Time every fused residual+restriction variant (EVO_RR_VARIANT) on one finest-level statement; CUDA events through
evo_cycle_profile_op.  Usage: rr_variants.py [level] [variants...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evostencils_b200 import backend, cycles, oplist as ol, problems  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    variants = [int(v) for v in sys.argv[2:]] or [0, 9, 1, 2, 3, 4, 10, 11, 12, 13, 14, 15, 16, 17]
    prob = problems.Poisson3D(level - 1, level)
    peak = 6554.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    op = ol.Op(ol.OP_RESIDUAL_RESTRICT, level, dst=ol.BUF_RHS, src=ol.BUF_RES)
    prog = cycles.build_program(prob, [op])
    cyc = backend.DeviceProblem(prob).build(prog)
    ndof = float((prob.nodes(level) - 2) ** 3)
    for v in variants:
        backend.set_option("EVO_RR_VARIANT", v)
        try:
            ms, n = cyc.profile_op(op, repeat=10)
        except backend.BackendError as e:
            print(f"variant {v}: error {e}")
            continue
        gbs = 17.0 * ndof / (ms * 1e-3) / 1e9
        print(f"variant {v:3d}: {ms:7.3f} ms {n} launches  {gbs:7.1f} GB/s = {100 * gbs / peak:5.1f}% of peak", flush=True)
    backend.set_option("EVO_RR_VARIANT", 0)


if __name__ == "__main__":
    main()
