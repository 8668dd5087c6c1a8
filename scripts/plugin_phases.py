#!/usr/bin/env python
"""Where does the time of ONE generate_and_evaluate call go?  Phases of the sequential plugin path for the first N
individuals of bench.py's generation: tree, lowering, evo_cycle_build, evo_cycle_solve (device time inside), close."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from evostencils_b200 import tree  # noqa: E402
from evostencils_b200.program_generator import B200ProgramGenerator  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    probs, individuals = bench.population_individuals(256)
    gens = [B200ProgramGenerator(problem=p) for p in probs]
    acc = {"tree": 0.0, "lower": 0.0, "build": 0.0, "solve_call": 0.0, "device_ms": 0.0, "close": 0.0, "iters": 0}
    picked = individuals[:n // 2] + individuals[128:128 + n // 2]
    for rep in range(2):
        for key in acc:
            acc[key] = 0
        for k, s in picked:
            g, pr = gens[k], probs[k]
            t0 = time.perf_counter()
            e = tree.build_tree(pr, s)
            t1 = time.perf_counter()
            prog = g._finalise(g.lower(e, g.min_level))
            t2 = time.perf_counter()
            dev = g._device_problem(g.min_level, g.max_level)
            cyc = dev.build(prog)
            t3 = time.perf_counter()
            out = cyc.solve(pr.settings.tol, pr.settings.max_iters, samples=1)
            t4 = time.perf_counter()
            cyc.close()
            t5 = time.perf_counter()
            acc["tree"] += t1 - t0; acc["lower"] += t2 - t1; acc["build"] += t3 - t2; acc["solve_call"] += t4 - t3
            acc["close"] += t5 - t4; acc["device_ms"] += out.time_ms; acc["iters"] += out.iterations
        m = len(picked)
        print(f"pass {rep}: per individual: tree {acc['tree'] / m * 1e3:.2f} ms, lower {acc['lower'] / m * 1e3:.2f}, build {acc['build'] / m * 1e3:.2f}, "
              f"solve call {acc['solve_call'] / m * 1e3:.2f} (device {acc['device_ms'] / m:.2f}), close {acc['close'] / m * 1e3:.2f}, "
              f"iterations {acc['iters'] / m:.1f}", flush=True)


if __name__ == "__main__":
    main()
