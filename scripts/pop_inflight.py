#!/usr/bin/env python
"""Throughput of the population evaluation as a function of the number of individuals in flight (one problem type)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import bench  # noqa: E402
from evostencils_b200 import tree  # noqa: E402
from evostencils_b200.program_generator import B200ProgramGenerator  # noqa: E402


def main():
    probs, individuals = bench.population_individuals(256)
    for k in (0, 1):
        gen = B200ProgramGenerator(problem=probs[k])
        progs = [gen.lower(tree.build_tree(probs[k], s), gen.min_level) for kk, s in individuals if kk == k]
        gen.evaluate_population([], programs=progs[:8], solo_timing=False)
        for inflight in (1, 2, 4, 8, 16, 32, 64, 128):
            t0 = time.perf_counter()
            res, busy = gen.evaluate_population([], programs=progs, max_in_flight=inflight, solo_timing=False)
            t = time.perf_counter() - t0
            print(f"problem {probs[k].name} in flight {inflight:4d}: {len(progs) / t:7.1f} evals/s wall, device busy {busy:7.1f} ms "
                  f"-> {len(progs) / busy * 1e3:7.1f} evals/s device", flush=True)
        gen.close()


if __name__ == "__main__":
    main()
