#!/usr/bin/env python
"""Where does a generation's time go?  Per-individual solve times (one at a time, no concurrency)."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from evostencils_b200 import tree, oplist as ol
from evostencils_b200.program_generator import B200ProgramGenerator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
probs, inds = bench.population_individuals(n)
gens = [B200ProgramGenerator(problem=p) for p in probs]
rows = []
for k, s in inds:
    g = gens[k]
    prog = g._finalise(g.lower(tree.build_tree(probs[k], s), g.min_level))
    dev = g._device_problem(g.min_level, g.max_level)
    cyc = dev.build(prog)
    out = cyc.solve(1e-12, 100, 1)
    out = cyc.solve(1e-12, 100, 1)
    kinds = collections.Counter()
    for o in prog.ops:
        if o.code == ol.OP_SMOOTH:
            kinds[("rb" if o.mode == ol.MODE_REDBLACK else "jac") + str(len(o.unknowns))] += 1
    rows.append((k, out.time_ms, out.iterations, out.kernel_launches, len(prog.ops), dict(kinds)))
    cyc.close()
for k in (0, 1):
    sel = [r for r in rows if r[0] == k]
    tot = sum(r[1] for r in sel)
    print(f"problem {k}: {len(sel)} individuals, total {tot:.1f} ms, mean {tot/len(sel):.2f} ms, max {max(r[1] for r in sel):.1f} ms, "
          f"mean iterations {sum(r[2] for r in sel)/len(sel):.1f}, mean us/launch {1e3*tot/sum(r[3] for r in sel):.2f}")
for r in sorted(rows, key=lambda r: -r[1])[:12]:
    print(f"  p{r[0]} {r[1]:9.2f} ms  its={r[2]:3d} launches={r[3]:6d} ops={r[4]:3d} us/launch={1e3*r[1]/max(r[3],1):6.2f} {r[5]}")
