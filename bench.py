#!/usr/bin/env python
"""bench.py -- headline measurement of the evaluate hot path (contract: see the task statement).

BASELINE.json's metric is "evolved-cycle fitness evals/s on 8xB200; smoother GDOF/s vs HBM roofline".  The default
line therefore carries both halves:

* ``value`` / ``e2e``: ONE G3P GENERATION (configs[4]) -- 256 evolved cycles, half Poisson 2D (513^2, levels 9..5), half
  LinearElasticity (257^2, 2 fields, levels 8..4), sharded round-robin over the N GPUs (``i % N == rank``, reference:
  optimization/program.py:534-538), fitness tuples gathered on the host.  A *step* is one generation = 256 fitness
  evaluations (each one: lower the tree, run the generated solver's outer loop to 1e-12, return (ms, cf, iterations)).
  The work is fixed as N grows: ``scaling = "strong"``.
* ``grid513``: ONE evaluation of the Poisson 3D 513^3 V(2,1) red-black GS cycle (configs[1]) -- on one GPU at N = 1, the
  same grid domain-decomposed into z-slabs over all N GPUs at N > 1 (halo exchange over NCCL; strong scaling), with the
  residual history compared bit for bit against the single-GPU run.
* ``roofline``: the dominant kernel of the 513^3 evaluation (finest-level RB-GS sweep) against the measured HBM peak.

  python bench.py --gpus 1 --steps 5 --warmup 3            our arm (CUDA library through the C-ABI / drop-in generator)
  python bench.py --impl reference ...                     the CPU arm: oracle port on the host cores, same workload
  torchrun ... bench.py --gpus N ...                       one rank per GPU
  python bench.py --workload poisson3d_513 [--no-graph]    single-workload modes (profiling, other BASELINE configs)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one stream per in-flight individual

import numpy as np  # noqa: E402

from evostencils_b200 import cycles, fitness, oplist as ol, problems  # noqa: E402

HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
DEFAULT_WORKLOAD = "generation256+grid513"
METRIC = ("evolved-cycle fitness evals/s (one G3P generation of 256 evolved cycles = 256 evals per step; "
          "one eval = lower the tree + solve to 1e-12 with the cycle); smoother GDOF/s in roofline")
GRID_WORKLOAD = "poisson3d_513"
# individuals of the generation the CPU arm solves per step (2 Poisson 2D + 2 elasticity, evenly spread)
CPU_SAMPLE = (0, 99, 132, 231)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def make_workload(name: str, fuse: bool = True):
    """The problem and its `generate solver` cycle, lowered the way the drop-in ProgramGenerator lowers it
    (program_generator.py: fuse=True -> lowering.optimise merges residual + restriction where the residual is dead)."""
    wcycle = name.endswith("_w")           # same problem, W-cycle (gamma = 2) instead of the V-cycle of the solver block
    if wcycle:
        name = name[:-2]
    if name == "poisson3d_513":
        prob = problems.Poisson3D(2, 9)
    elif name == "poisson3d_257":
        prob = problems.Poisson3D(2, 8)
    elif name == "poisson3d_129":
        prob = problems.Poisson3D(2, 7)
    elif name == "poisson2d_513":
        prob = problems.Poisson2D(5, 9)
    elif name == "poisson2d_4097":
        prob = problems.Poisson2D(5, 12)
    else:
        raise SystemExit(f"unknown workload {name}")
    s_ = prob.settings
    prog = cycles.w_cycle(prob, s_.num_pre, s_.num_post, s_.damping, s_.red_black) if wcycle else cycles.default_solver_cycle(prob)
    from evostencils_b200 import lowering
    expr = None if wcycle else workload_tree(name, prob)
    if expr is not None:
        # the cycle as the grammar individual Optimizer hands over, lowered like the drop-in lowers it: the op list of the
        # device-resident run, the CPU arm and the plugin call are then the same one (the grammar's weight grid
        # linspace(0.1, 1.9, 37) holds 0.9999999999999999 for the coarse-grid correction, not 1.0)
        prog = lowering.lower_cycle(expr, prob.min_level, prob.max_level, prob.n_fields, prob.dim, cgs_max_iters=s_.cgs_max_iters,
                                    cgs_tol=s_.cgs_tol, default_restrict=prob.restrict_weights(), default_prolong=prob.prolong_weights())
    if fuse:
        prog = lowering.optimise(prog)
    return prob, prog


def make_workload_cycle(name: str, prob):
    """The workload's cycle for another level range of the same problem (CPU samples on coarser hierarchies)."""
    from evostencils_b200 import lowering
    s_ = prob.settings
    if name.endswith("_w"):
        return lowering.optimise(cycles.w_cycle(prob, s_.num_pre, s_.num_post, s_.damping, s_.red_black))
    expr = workload_tree(name, prob)
    if expr is None:
        return cycles.default_solver_cycle(prob)
    return lowering.optimise(lowering.lower_cycle(expr, prob.min_level, prob.max_level, prob.n_fields, prob.dim,
                                                  cgs_max_iters=s_.cgs_max_iters, cgs_tol=s_.cgs_tol,
                                                  default_restrict=prob.restrict_weights(), default_prolong=prob.prolong_weights()))


def workload_tree(name: str, prob):
    """The same cycle as a grammar individual (what Optimizer hands to generate_and_evaluate)."""
    from evostencils_b200 import tree
    s_ = prob.settings
    if name.endswith("_w"):
        return None
    # weight index of the solver block's damping on the grammar's grid linspace(0.1, 1.9, 37)
    grid = np.linspace(0.1, 1.9, 37)
    idx = int(np.argmin(np.abs(grid - s_.damping)))
    if abs(grid[idx] - s_.damping) > 1e-12:
        return None
    string = tree.v_cycle_individual(prob.max_level - prob.min_level, s_.num_pre, s_.num_post, idx,
                                     partitioning="red_black" if s_.red_black else "single", cgc_weight_index=18)
    return tree.build_tree(prob, string)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1])); mx.append(float(t[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sorted(sm)[len(sm) // 2:])   # median of the upper half = under load
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


class Dist:
    """torch.distributed plumbing (NCCL, one rank per GPU); a no-op at world size 1."""

    def __init__(self, rank, world, local_rank):
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dist = None
        self.torch = None
        if world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max(self, *vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_int(self, v):
        if self.dist is None:
            return int(v)
        t = self.torch.tensor([int(v)], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C/OpenMP restatement of the generated solver) on the host cores
def _cpu_threads(threads=None):
    """All host threads unless told otherwise.  torchrun exports OMP_NUM_THREADS=1 to its workers: that is a launcher
    default, not a property of the reference, so the thread count is set explicitly."""
    from oracle import oracle as orc
    n = threads or os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    orc.set_num_threads(max(1, n))
    return orc.num_threads()


def cpu_grid_sample(workload: str, threads=None, budget_s: float = 25.0):
    """Oracle on the grid workload: ONE COMPLETE SOLVE (initial residual, V-cycles to 1e-12) at the finest level whose
    solve fits the time budget, scaled to the workload's DOF count.  Multigrid cost is linear in the DOFs and the
    iteration count is h-independent (the parity tests assert equal counts), so the only extrapolation is the DOF
    ratio; when level == the workload's finest level there is none."""
    from oracle import oracle as orc
    prob, _ = make_workload(workload)
    nthreads = _cpu_threads(threads)
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 16 << 30
    # probe: one full solve two levels below, predicts the cost of the finer ones (x 2^dim per level)
    probe_level = max(prob.min_level + 1, prob.max_level - 2)
    probe = prob.with_levels(prob.min_level, probe_level)
    oc = orc.OracleProblem(probe).build(make_workload_cycle(workload, probe))
    oc.solve(probe.settings.tol, probe.settings.max_iters, 1)
    t0 = time.perf_counter()
    ref = oc.solve(probe.settings.tol, probe.settings.max_iters, 1)
    t_probe = time.perf_counter() - t0
    level, t_solve, its = probe_level, t_probe, ref.iterations
    for cand in range(prob.max_level, probe_level, -1):
        need = 5.5 * 8 * (prob.nodes(cand) ** prob.dim) * prob.n_fields
        predicted = t_probe * (2 ** prob.dim) ** (cand - probe_level)
        if need < 0.4 * avail and need < (24 << 30) and predicted <= budget_s:
            sp = prob.with_levels(prob.min_level, cand)
            oc = orc.OracleProblem(sp).build(make_workload_cycle(workload, sp))
            t0 = time.perf_counter()
            ref = oc.solve(sp.settings.tol, sp.settings.max_iters, 1)
            t_solve, level, its = time.perf_counter() - t0, cand, ref.iterations
            break
    scale = float((prob.nodes(prob.max_level) - 2) ** prob.dim) / float((prob.nodes(level) - 2) ** prob.dim)
    return {"value": 1.0 / (t_solve * scale), "unit": "evals/s", "cores": nthreads, "kind": "port",
            "sample": f"one complete solve ({its} iterations to 1e-12) of the workload's cycle at level {level} "
                      f"({prob.nodes(level)}^{prob.dim} nodes, {t_solve:.2f} s) x {scale:.3g} (DOF ratio to level {prob.max_level}); "
                      f"oracle = C/OpenMP restatement, gcc -O3 -fopenmp, {nthreads} threads",
            "seconds_per_solve_at_sample_level": t_solve, "iterations": its}


def population_individuals(n_total: int, seed: int = 0):
    """BASELINE.json configs[4]: grammar-valid random individuals, alternating Poisson 2D (levels 5..9) and
    LinearElasticity (levels 4..8)."""
    import random
    from evostencils_b200 import tree
    probs = [problems.Poisson2D(5, 9), problems.LinearElasticity2D(4, 8)]
    rng = random.Random(seed)
    out = []
    for i in range(n_total):
        prob = probs[i % 2]
        out.append((i % 2, tree.random_individual(prob, rng, maximum_local_system_size=4)))
    return probs, out


def cpu_population_sample(n_total: int, threads=None, sample=CPU_SAMPLE):
    """Oracle on the generation workload: complete solves of a fixed, evenly spread sample of the SAME individuals."""
    from evostencils_b200 import lowering, tree
    from oracle import oracle as orc
    nthreads = _cpu_threads(threads)
    probs, individuals = population_individuals(n_total)
    idx = [i for i in sample if i < n_total] or list(range(min(4, n_total)))
    hosts = [orc.OracleProblem(p) for p in probs]
    t0 = time.perf_counter()
    for i in idx:
        k, s = individuals[i]
        p = probs[k]
        prog = lowering.lower_cycle(tree.build_tree(p, s), p.min_level, p.max_level, p.n_fields, p.dim,
                                    cgs_max_iters=p.settings.cgs_max_iters, cgs_tol=p.settings.cgs_tol,
                                    default_restrict=p.restrict_weights(), default_prolong=p.prolong_weights())
        prog = lowering.optimise(prog)
        out = hosts[k].build(prog).solve(p.settings.tol, p.settings.max_iters, 1)
        fitness.fitness_from_history(out.residuals, out.time_ms, p.settings.max_iters)
    t = time.perf_counter() - t0
    return {"value": len(idx) / t, "unit": "evals/s", "cores": nthreads, "kind": "port",
            "sample": f"complete evaluations (lowering + solve to 1e-12 or 100 iterations) of individuals {list(idx)} of the "
                      f"{n_total}-individual generation ({sum(1 for i in idx if i % 2 == 0)} Poisson 2D 513^2, "
                      f"{sum(1 for i in idx if i % 2 == 1)} LinearElasticity 257^2), one after the other, oracle = C/OpenMP "
                      f"restatement with {nthreads} threads, {t:.1f} s; solve time only -- the reference additionally runs "
                      f"the Java generator twice and make per individual",
            "seconds": t, "individuals": len(idx)}


def workload_config(name, prob, world):
    return {"workload": name, "problem": prob.name, "finest_nodes": f"{prob.nodes(prob.max_level)}^{prob.dim}",
            "levels": f"{prob.max_level}..{prob.min_level}",
            "cycle": f"{'W' if name.endswith('_w') else 'V'}({prob.settings.num_pre},{prob.settings.num_post}) red-black GS omega={prob.settings.damping} + CG",
            "tol": prob.settings.tol, "max_iters": prob.settings.max_iters,
            "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas (one evaluation per GPU per step)",
            "l2_policy": "inputs larger than L2 (finest fields 1.1 GB each)" if prob.dim == 3 and prob.max_level >= 8
            else "working set fits L2 (latency-bound regime; no flush)"}


def generation_config(n_total, world, in_flight):
    return {"workload": DEFAULT_WORKLOAD, "individuals": n_total,
            "problems": "Poisson 2D levels 9..5 (513^2) + LinearElasticity 2D levels 8..4 (257^2, 2 fields), alternating",
            "generator": "evostencils_b200.tree.random_individual, seed 0, local systems <= 4",
            "tol": 1e-12, "max_iters": 100, "in_flight_per_gpu": in_flight,
            "parallelism": f"population sharded round-robin over {world} GPU(s), fitness gathered on the host; "
                           f"grid513: one 513^3 evaluation " + ("on one GPU" if world == 1 else f"domain-decomposed into {world} z-slabs"),
            "grid513": "Poisson 3D 7-point, 513^3, levels 9..2, V(2,1) red-black GS omega=1.25 + CG, tol 1e-12",
            "l2_policy": "generation: working sets fit L2 (latency / launch bound regime, no flush); grid513 and roofline: "
                         "inputs larger than L2 (finest fields 1.1 GB each)"}


def run_reference(args, rank, world):
    """The reference's CPU path (oracle port: ExaStencils cannot be generated here) on the arm's workload, all host
    threads, rank 0 only; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    vals, info = [], None
    if args.workload == DEFAULT_WORKLOAD:
        cfg = generation_config(args.population, world, args.in_flight)
        for i in range(args.warmup + args.steps):
            info = cpu_population_sample(args.population)
            if i >= args.warmup:
                vals.append(info["value"])
    else:
        prob, _ = make_workload(args.workload)
        cfg = workload_config(args.workload, prob, world)
        for i in range(args.warmup + args.steps):
            info = cpu_grid_sample(args.workload)
            if i >= args.warmup:
                vals.append(info["value"])
    value = statistics.mean(vals)
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
            "scaling": "strong" if args.workload == DEFAULT_WORKLOAD else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "cpu_baseline": info,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def roofline_of(cyc, prob, workload):
    """Dominant kernel of the grid evaluation: finest-level RB-GS sweep, CUDA-event timed through evo_cycle_profile_op."""
    peak, peak_src = measured_peak()
    s = prob.settings
    ndof = float((prob.nodes(prob.max_level) - 2) ** prob.dim) * prob.n_fields
    zero = (0,) * prob.dim
    sm_op = ol.Op(ol.OP_SMOOTH, prob.max_level, mode=ol.MODE_REDBLACK, omega=s.damping,
                  unknowns=tuple((f, zero) for f in range(prob.n_fields)))
    ms, n_launch = cyc.profile_op(sm_op, repeat=10)
    alg_bytes = 24.0 * ndof                      # read u, f, write u per full sweep
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes of one launch from the committed ncu --set full capture of this kernel at this size
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as fh:
            traffic = json.load(fh).get(workload, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roof = {"bound": "hbm", "kernel": "RB-GS sweep, finest level (both colours): k3_rbgs_col", "achieved": achieved,
            "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "algorithmic_bytes_per_sweep": alg_bytes, "ms_per_sweep": ms, "launches_per_sweep": n_launch,
            "smoother_gdof_s": ndof / (ms * 1e-3) / 1e9}
    other = {}
    for nm, op, b in (("residual", ol.Op(ol.OP_RESIDUAL, prob.max_level, dst=ol.BUF_RES), 24.0),
                      ("residual+restrict", ol.Op(ol.OP_RESIDUAL_RESTRICT, prob.max_level, dst=ol.BUF_RHS, src=ol.BUF_RES),
                       16.0 + 8.0 / 2 ** prob.dim),
                      ("prolong_add", ol.Op(ol.OP_PROLONG_ADD, prob.max_level, src=ol.BUF_SOL, omega=1.0),
                       16.0 + 8.0 / 2 ** prob.dim)):
        try:
            m2, _ = cyc.profile_op(op, repeat=10)
            other[nm] = {"ms": m2, "GB/s": b * ndof / (m2 * 1e-3) / 1e9, "frac": b * ndof / (m2 * 1e-3) / 1e9 / peak}
        except Exception as e:   # pragma: no cover
            other[nm] = {"error": str(e)}
    roof["other_kernels"] = other
    return roof


def measure_grid(args, D: Dist, workload: str, want_roofline: bool):
    """One evaluation of the grid workload: device-resident solves (value), the reference-facing plugin call with the
    tree as input (e2e), and at N > 1 the same evaluation domain-decomposed over all GPUs."""
    from evostencils_b200 import backend
    from evostencils_b200.program_generator import B200ProgramGenerator
    prob, prog = make_workload(workload)
    s = prob.settings
    flags = ol.SOLVE_NO_GRAPH if args.no_graph else 0
    ndof = float((prob.nodes(prob.max_level) - 2) ** prob.dim) * prob.n_fields
    res = {"workload": workload}
    single = None
    roof = None
    if D.world == 1 or D.rank == 0 or args.replicas:
        dev = backend.DeviceProblem(prob, device=D.local_rank)
        cyc = dev.build(prog)
        for _ in range(args.warmup):
            out = cyc.solve(s.tol, s.max_iters, 1, flags)
        t_dev, launches = 0.0, 0
        for _ in range(args.steps):
            out = cyc.solve(s.tol, s.max_iters, 1, flags)
            t_dev += out.time_ms
            launches += out.kernel_launches
        single = out
        res.update({"single_gpu_ms_per_eval": t_dev / args.steps, "single_gpu_evals_per_s": args.steps / (t_dev * 1e-3),
                    "iterations": out.iterations, "gpu_launches": launches,
                    "convergence_factor": fitness.fitness_from_history(out.residuals, out.time_ms, s.max_iters)[1],
                    "ms_per_cycle": t_dev / args.steps / max(out.iterations, 1),
                    "cycle_gdof_s": ndof * out.iterations * args.steps / (t_dev * 1e-3) / 1e9})
        if want_roofline and D.rank == 0:
            roof = roofline_of(cyc, prob, workload)
        cyc.close()
        # e2e: the call Optimizer makes -- generate_and_evaluate(tree, storages, ...): lowering, op list / operator tables
        # host -> device, solve, residual history device -> host, fitness tuple
        expr = workload_tree(workload, prob)
        if expr is not None and D.rank == 0 and not args.no_graph:
            pg = B200ProgramGenerator(problem=prob, device=D.local_rank)
            storages = pg.generate_storage(prob.min_level, prob.max_level, None)
            for _ in range(min(args.warmup, 2)):
                fit = pg.generate_and_evaluate(expr, storages, prob.min_level, prob.max_level, "", evaluation_samples=1)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fit = pg.generate_and_evaluate(expr, storages, prob.min_level, prob.max_level, "", evaluation_samples=1)
            t_e2e = time.perf_counter() - t0
            lowered = pg._finalise(pg.lower(expr, prob.min_level))
            res["e2e"] = {"value": args.steps / t_e2e, "unit": "evals/s", "ms_per_eval": 1e3 * t_e2e / args.steps,
                          "h2d_bytes_per_step": len(lowered.ops) * 160 + len(lowered.operators) * (8 + 2 * 2 * 27 * 2 * 8),
                          "d2h_bytes_per_step": (s.max_iters + 1) * 8 + 48,
                          "fitness": [float(v) for v in fit], "identical_history": bool(np.array_equal(pg.last_outcome.residuals, out.residuals)),
                          "note": "B200ProgramGenerator.generate_and_evaluate(tree, storages, ...): lowering + evo_cycle_build + "
                                  "evo_cycle_solve + history read-back + fitness; the call carries no field data (the problem is analytic)"}
            pg.close()
        dev.close()
    if D.world > 1 and prob.dim == 3 and not args.no_domain:
        import signal
        from evostencils_b200 import domain

        def give_up(signum, frame):   # the decomposed measurement must never cost the line
            raise TimeoutError("domain-decomposed measurement did not finish within the time limit")

        signal.signal(signal.SIGALRM, give_up)
        signal.alarm(int(os.environ.get("EVO_DOMAIN_TIME_LIMIT", "300")))
        try:
            D.barrier()
            solver = domain.DomainSolver.distributed(prob, prog, D.rank, D.world, D.local_rank, lc=args.lc or None)
            solver.overlap = not args.domain_no_overlap
            dd_solve = solver.solve if args.domain_eager else solver.solve_captured
            o3 = solver.solve(s.tol, s.max_iters)          # creates the NCCL communicators (not capturable)
            for _ in range(2):
                o3 = dd_solve(s.tol, s.max_iters)
            D.barrier()
            t_dd = 0.0
            for _ in range(args.steps):
                o3 = dd_solve(s.tol, s.max_iters)
                t_dd += o3.time_ms
            D.barrier()
            (t_dd,) = D.max(t_dd)
            same = None
            if single is not None:
                same = bool(np.array_equal(o3.residuals, single.residuals))
            res["domain_decomposition"] = {
                "n_gpus": D.world, "ms_per_eval": t_dd / args.steps, "evals_per_s": args.steps / (t_dd * 1e-3), "scaling": "strong",
                "iterations": o3.iterations, "identical_history": same, "halo_exchanges_per_eval": o3.exchanges,
                "overlap": solver.overlap, "levels_distributed": f"{prob.max_level}..{solver.layout.lc}",
                "speedup_vs_one_gpu": (res["single_gpu_ms_per_eval"] / (t_dd / args.steps)) if "single_gpu_ms_per_eval" in res else None,
                "note": "ONE evaluation split into z-slabs over all GPUs, halos over NCCL send/recv; "
                        + ("host-orchestrated statements" if args.domain_eager else "each iteration (kernels + exchanges) replayed as one CUDA graph")}
            solver.close()
        except Exception as e:   # pragma: no cover - reported, never fatal
            res["domain_decomposition"] = {"error": f"{type(e).__name__}: {e}"}
        signal.alarm(0)
    # headline numbers of this block: N = 1 -> the single GPU, N > 1 -> the decomposed evaluation
    dd = res.get("domain_decomposition")
    if dd and "ms_per_eval" in dd:
        res.update({"n_gpus": D.world, "ms_per_eval": dd["ms_per_eval"], "evals_per_s": dd["evals_per_s"], "scaling": "strong"})
    elif "single_gpu_ms_per_eval" in res:
        res.update({"n_gpus": 1, "ms_per_eval": res["single_gpu_ms_per_eval"], "evals_per_s": res["single_gpu_evals_per_s"],
                    "scaling": "strong"})
    return res, roof


class _LazyTrees:
    """Grammar strings of saved individuals -> expression trees, built when iterated (by the lowering thread)."""

    def __init__(self, problem, strings):
        self.problem, self.strings = problem, strings

    def __len__(self):
        return len(self.strings)

    def __iter__(self):
        from evostencils_b200 import tree
        return (tree.build_tree(self.problem, s) for s in self.strings)


def measure_generation(args, D: Dist):
    """One G3P generation sharded over the ranks.  Returns the timings (max over ranks) and the gathered fitness list."""
    from evostencils_b200 import population as popmod, tree
    from evostencils_b200.program_generator import B200ProgramGenerator
    n_total = args.population
    probs, individuals = population_individuals(n_total)
    mine = [individuals[i] for i in popmod.shard_indices(n_total, D.rank, D.world)]
    gens = [B200ProgramGenerator(problem=p, device=D.local_rank) for p in probs]
    for g in gens:
        g.initialize_code_generation(g.min_level, g.max_level)

    def lower_all():   # host side of the hot path: grammar string -> tree -> lowered program
        progs = [[], []]
        for k, s in mine:
            progs[k].append(gens[k].lower(tree.build_tree(probs[k], s), gens[k].min_level))
        return progs

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=2)

    def both(fn):
        """The two problems of the generation are evaluated by two host threads at the same time (the C calls release
        the interpreter lock; captures are thread local): with few individuals per GPU neither pipeline alone keeps
        enough solves in flight."""
        futures = [pool.submit(fn, k) for k in (0, 1)]
        return [f.result() for f in futures]

    batch_ms = [0.0]

    def evaluate(progs, solo=True):
        launches0 = sum(g.total_kernel_launches for g in gens)
        # the contention-free re-timing needs an idle device: concurrent pipelines only without it, then one after the other
        out = both(lambda k: gens[k].evaluate_population([], programs=progs[k], max_in_flight=args.in_flight, solo_timing=False,
                                                         keep_for_retime=solo))
        ms, res = max(out[0][1], out[1][1]), []
        batch_ms[0] += ms            # the concurrent phase alone (both pipelines at once)
        for k in (0, 1):
            r, t = gens[k].finish_retime() if solo else (out[k][0], 0.0)
            ms += t
            res += r
        return ms, res, sum(g.total_kernel_launches for g in gens) - launches0

    def ordered(progs, res):
        order, out, n0 = {0: 0, 1: 0}, [], len(progs[0])
        split = {0: res[:n0], 1: res[n0:]}
        for k, _ in mine:
            out.append(split[k][order[k]])
            order[k] += 1
        return out

    progs = lower_all()
    for _ in range(args.warmup):
        evaluate(progs)
    # ---- device-resident: programs already lowered ------------------------------------------------------------
    D.barrier()
    t_dev, launches = 0.0, 0
    batch_ms[0] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms, res, ln = evaluate(progs)
        t_dev += ms
        launches += ln
        all_fitness = popmod.evaluate_sharded(individuals, lambda _m: ordered(progs, res), D.rank, D.world, D.dist)
    D.barrier()
    t_wall = time.perf_counter() - t0
    # ---- throughput of the concurrent batch alone (the timed steps without their solo re-timing phase) ----------------
    t_batch_only = batch_ms[0] * 1e-3 / args.steps
    # ---- end to end: strings -> trees -> lowering -> build -> solve -> fitness tuples gathered on the host --------
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        # strings in, fitness tuples out: both problems at once (two host threads), trees lowered by background threads
        # while the device works, then the contention-free re-timing of the time objective on the idle device
        def run(k):
            trees_k = _LazyTrees(probs[k], [s for kk, s in mine if kk == k])     # built by the lowering thread, one ahead of the device
            return gens[k].evaluate_population(trees_k, max_in_flight=args.in_flight, solo_timing=False, keep_for_retime=True)
        both(run)
        res2 = []
        for k in (0, 1):
            res2 += gens[k].finish_retime()[0]
        all_fitness = popmod.evaluate_sharded(individuals, lambda _m: ordered(progs, res2), D.rank, D.world, D.dist)
    D.barrier()
    t_e2e = time.perf_counter() - t0
    # ---- the literal plugin call, one individual after the other (what Optimizer's toolbox.map does) ---------------
    seq = None
    if D.rank == 0:
        k_seq = min(16, len(mine))
        trees = [(k, tree.build_tree(probs[k], s)) for k, s in mine[:k_seq]]
        storages = [g.generate_storage(g.min_level, g.max_level, None) for g in gens]
        t0 = time.perf_counter()
        for k, e in trees:
            gens[k].generate_and_evaluate(e, storages[k], gens[k].min_level, gens[k].max_level, "", evaluation_samples=1)
        seq = k_seq / (time.perf_counter() - t0)
    t_dev, t_wall, t_e2e, t_batch_only = D.max(t_dev, t_wall, t_e2e, t_batch_only)
    launches = D.sum_int(launches)
    h2d = D.sum_int(sum(len(p.ops) * 160 + len(p.operators) * 1736 for ps in progs for p in ps))
    for g in gens:
        g.close()
    return {"t_dev_ms": t_dev, "t_wall": t_wall, "t_e2e": t_e2e, "t_batch_only": t_batch_only, "launches": launches, "h2d": h2d,
            "d2h": n_total * (101 * 8 + 48), "fitness": all_fitness, "sequential_plugin_evals_per_s": seq}


def run_default(args, rank, world, local_rank):
    D = Dist(rank, world, local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.2 s to start: begin before the warm-up, sample through the timed regions
    gen = measure_generation(args, D)
    grid, roof = measure_grid(args, D, GRID_WORKLOAD, want_roofline=True)
    clocks = sampler.stop() if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_population_sample(args.population)
        grid["cpu_port"] = cpu_grid_sample(GRID_WORKLOAD)
    if rank == 0:
        n_total = args.population
        evals = n_total * args.steps
        converged = sum(1 for r in gen["fitness"] if r[1] < 1)
        line = {"metric": METRIC, "value": evals / gen["t_wall"], "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": gen["t_wall"] * 1e3 / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": generation_config(n_total, world, args.in_flight),
                "device_busy_ms_per_step": gen["t_dev_ms"] / args.steps,
                "value_without_solo_retiming": n_total / gen["t_batch_only"],
                "e2e": {"value": evals / gen["t_e2e"], "unit": "evals/s", "h2d_bytes_per_step": gen["h2d"],
                        "d2h_bytes_per_step": gen["d2h"],
                        "sequential_plugin_evals_per_s": gen["sequential_plugin_evals_per_s"],
                        "note": "grammar strings -> trees -> lowering -> B200ProgramGenerator.evaluate_population (evo_cycle_build, "
                                "evo_batch_solve incl. the contention-free re-timing of the time objective) -> fitness tuples gathered "
                                "on the host; sequential_plugin = generate_and_evaluate one individual after the other (1 GPU)"},
                "gpu_launches": gen["launches"] + int(grid.get("gpu_launches") or 0), "clocks": clocks,
                "converging_individuals": converged, "grid513": grid, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    D.close()


def run_single(args, rank, world, local_rank):
    """One grid workload (profiling / other BASELINE configurations); at N > 1 every rank solves a replica."""
    D = Dist(rank, world, local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    args.replicas = True
    prob, _ = make_workload(args.workload)
    D.barrier()
    grid, roof = measure_grid(args, D, args.workload, want_roofline=not args.no_graph)
    (t_ms,) = D.max(grid["single_gpu_ms_per_eval"])
    launches = D.sum_int(grid["gpu_launches"])
    clocks = sampler.stop() if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_grid_sample(args.workload)
    if rank == 0:
        e2e = grid.get("e2e") or {"value": None, "unit": "evals/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}
        line = {"metric": METRIC, "value": world / (t_ms * 1e-3), "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, prob, world),
                "iterations_per_eval": grid["iterations"], "convergence_factor": grid["convergence_factor"],
                "cycle_gdof_s": grid["cycle_gdof_s"], "ms_per_cycle": grid["ms_per_cycle"], "e2e": e2e,
                "domain_decomposition": grid.get("domain_decomposition"), "gpu_launches": launches, "clocks": clocks,
                "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    D.close()


def main():
    # NCCL kernels of the halo exchanges should not queue behind the thousands of CTAs of an interior sweep
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--population", type=int, default=256)
    ap.add_argument("--in-flight", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch kernels directly (host-side solver loop) so that ncu can see them; not a bench value")
    ap.add_argument("--domain-no-overlap", action="store_true",
                    help="domain decomposition: exchange after the whole sweep instead of boundary planes first")
    ap.add_argument("--domain-eager", action="store_true", help="domain decomposition without CUDA-graph capture")
    ap.add_argument("--no-domain", action="store_true", help="N > 1: skip the domain-decomposed measurement")
    ap.add_argument("--lc", type=int, default=0, help="domain decomposition: coarsest distributed level (default: automatic)")
    args = ap.parse_args()
    args.replicas = False
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "population":
        args.workload = DEFAULT_WORKLOAD
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == DEFAULT_WORKLOAD:
        run_default(args, rank, world, local_rank)
    else:
        run_single(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
