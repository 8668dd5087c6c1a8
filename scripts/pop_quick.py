#!/usr/bin/env python
"""One generation (256 individuals, bench.py's population) on one GPU: evals/s with the current tuning switches.
Usage: pop_quick.py [steps]   (environment: EVO_COARSE_FUSE etc.)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import bench  # noqa: E402
from evostencils_b200 import tree  # noqa: E402
from evostencils_b200.program_generator import B200ProgramGenerator  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    probs, individuals = bench.population_individuals(256)
    gens = [B200ProgramGenerator(problem=p) for p in probs]
    t0 = time.perf_counter()
    progs = [[], []]
    for k, s in individuals:
        progs[k].append(gens[k].lower(tree.build_tree(probs[k], s), gens[k].min_level))
    t_lower = time.perf_counter() - t0

    def evaluate(solo):
        ms = 0.0
        for k in (0, 1):
            _, t = gens[k].evaluate_population([], programs=progs[k], max_in_flight=int(os.environ.get('IN_FLIGHT', '48')), solo_timing=solo)
            ms += t
        return ms

    evaluate(False)
    for solo in (False, True):
        l0 = sum(g.total_kernel_launches for g in gens)
        t0 = time.perf_counter()
        busy = sum(evaluate(solo) for _ in range(steps))
        t = time.perf_counter() - t0
        ln = sum(g.total_kernel_launches for g in gens) - l0
        print(f"solo_timing={solo}: {256 * steps / t:7.1f} evals/s  device busy {busy / steps:7.1f} ms/generation  "
              f"{ln / steps / 256:7.0f} launches/individual  (lowering {t_lower * 1e3 / 256:.2f} ms/individual)", flush=True)


if __name__ == "__main__":
    main()
