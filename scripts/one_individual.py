#!/usr/bin/env python
"""Solve one individual of the population workload with direct launches (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from evostencils_b200 import tree, oplist as ol
from evostencils_b200.program_generator import B200ProgramGenerator
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
probs, inds = bench.population_individuals(idx + 1)
k, s = inds[idx]
g = B200ProgramGenerator(problem=probs[k])
prog = g._finalise(g.lower(tree.build_tree(probs[k], s), g.min_level))
cyc = g._device_problem(g.min_level, g.max_level).build(prog)
out = cyc.solve(1e-12, iters, 1, ol.SOLVE_NO_GRAPH)
print("problem", k, "ops", len(prog.ops), "iters", out.iterations, "ms", out.time_ms, "launches", out.kernel_launches)
for o in prog.ops:
    print("   ", ol.OP_NAMES[o.code], o.level, o.mode, len(o.unknowns), o.count)
