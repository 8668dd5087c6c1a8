#!/usr/bin/env python
"""Per-statement cost (inside a graph, evo_cycle_profile_op) of a few individuals of bench.py's generation.
Usage: op_costs.py [problem index 0|1] [individual index ...]"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from evostencils_b200 import oplist as ol, tree  # noqa: E402
from evostencils_b200.program_generator import B200ProgramGenerator  # noqa: E402

NAMES = {getattr(ol, n): n for n in dir(ol) if n.startswith("OP_") and isinstance(getattr(ol, n), int)}


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    which = [int(v) for v in sys.argv[2:]] or [0, 1, 2]
    probs, individuals = bench.population_individuals(256)
    mine = [s for kk, s in individuals if kk == k]
    g = B200ProgramGenerator(problem=probs[k])
    for i in which:
        prog = g._finalise(g.lower(tree.build_tree(probs[k], mine[i]), g.min_level))
        cyc = g._device_problem(g.min_level, g.max_level).build(prog)
        cyc.apply(1)
        out = cyc.solve(probs[k].settings.tol, probs[k].settings.max_iters, samples=1)
        total, by = 0.0, defaultdict(float)
        for op in prog.ops:
            ms, n = cyc.profile_op(op, repeat=5)
            total += ms
            key = (NAMES[op.code], op.level, op.mode if op.code == ol.OP_SMOOTH else "", len(op.unknowns or ()) if op.code == ol.OP_SMOOTH else "",
                   op.count if op.code == ol.OP_SMOOTH else "")
            by[key] += ms
        print(f"individual {i}: {len(prog.ops)} statements, sum of statements {total * 1e3:.1f} us per cycle; solve: {out.iterations} iterations, "
              f"{out.time_ms / max(1, out.iterations) * 1e3:.1f} us per iteration")
        for key, ms in sorted(by.items(), key=lambda kv: -kv[1])[:12]:
            print(f"    {ms * 1e3:8.1f} us  {key}")
        cyc.close()


if __name__ == "__main__":
    main()
