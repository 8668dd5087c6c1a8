#!/usr/bin/env python
"""bench.py -- headline measurement of the evaluate hot path (contract: see the task statement).

A *step* is one fitness evaluation = one pass of the hot path: the generated solver's outer loop
(residual, V-cycles until ``res < 1e-12 res0``) for one individual on synthetic (analytic) input that
is already resident in HBM.  Default workload (BASELINE.json configs[1]): Poisson 3D 7-point, 513^3
finest grid, levels 9..2, the problem's own V(2,1) red-black Gauss-Seidel cycle (omega 1.25) + CG.

  python bench.py --gpus 1 --steps 5 --warmup 3            our arm (CUDA library through the C-ABI)
  python bench.py --impl reference ...                     the CPU arm: oracle port on the host cores
  torchrun ... bench.py --gpus N ...                       one rank per GPU, population-sharded

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
    if fuse:
        from evostencils_b200 import lowering
        prog = lowering.optimise(prog)
    return prob, prog


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


# ------------------------------------------------------------------------------------------------
def cpu_arm_sample(workload: str, threads: int | None = None):
    """Oracle (CPU port of the generated solver) on a bounded sample of the workload.

    Sample = ONE V-cycle + residual norm on the largest level that fits comfortably in host memory;
    the iteration count of the full solve comes from running the same cycle to convergence on a
    65^3 / 129^2-class grid (multigrid convergence is h-independent; the GPU parity tests assert the
    counts agree).  Returns (evals_per_s, dict)."""
    from oracle import oracle as orc
    prob, _ = make_workload(workload)
    if threads:
        orc.set_num_threads(threads)
    nthreads = orc.num_threads()
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 16 << 30
    level = prob.max_level
    while level > prob.min_level + 1:
        need = 5.5 * 8 * (prob.nodes(level) ** prob.dim) * prob.n_fields
        if need < 0.4 * avail and need < (24 << 30):
            break
        level -= 1
    sample_prob = prob.with_levels(prob.min_level, level)
    prog = cycles.default_solver_cycle(sample_prob)
    oc = orc.OracleProblem(sample_prob).build(prog)
    oc.apply(1)                              # warm the pages
    t0 = time.perf_counter()
    oc.apply(1)
    oc.residual_norm()
    t_cycle = time.perf_counter() - t0
    # iterations to convergence on a small grid of the same problem
    small_level = min(level, 6 if prob.dim == 3 else 8)
    small = prob.with_levels(prob.min_level, small_level)
    its = orc.OracleProblem(small).build(cycles.default_solver_cycle(small)).solve(
        small.settings.tol, small.settings.max_iters, 1).iterations
    scale = float((prob.nodes(prob.max_level) - 2) ** prob.dim) / float((prob.nodes(level) - 2) ** prob.dim)
    t_eval = t_cycle * scale * its
    # the reference's own OpenMP setting is 4 threads (example_problems/lib/parallelization_pureOmp.knowledge:3)
    value_4 = None
    if threads is None and nthreads > 4:
        orc.set_num_threads(4)
        t0 = time.perf_counter()
        oc.apply(1)
        oc.residual_norm()
        value_4 = 1.0 / ((time.perf_counter() - t0) * scale * its)
        orc.set_num_threads(nthreads)
    info = {"value": 1.0 / t_eval, "unit": "evals/s", "cores": nthreads, "kind": "port",
            "sample": f"1 V-cycle + residual norm of the workload's cycle at level {level} "
                      f"({prob.nodes(level)}^{prob.dim} nodes, {t_cycle * 1e3:.1f} ms) x {scale:.3g} (DOF ratio to "
                      f"level {prob.max_level}) x {its} iterations (full solve at level {small_level}); "
                      f"oracle = C/OpenMP restatement, gcc -O3 -fopenmp, {nthreads} threads",
            "ms_per_cycle_at_sample_level": t_cycle * 1e3, "iterations": its, "value_4_threads": value_4}
    return info


def run_reference(args, rank, world):
    if rank != 0:
        return
    prob, _ = make_workload(args.workload)
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_arm_sample(args.workload)
        if i >= args.warmup:
            vals.append(info["value"])
        if i == 0 and args.warmup + args.steps > 1 and 1.0 / info["value"] > 0:
            pass
    value = statistics.mean(vals)
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, prob, world),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


METRIC = "evolved-cycle fitness evals/s (one eval = solve to 1e-12 with the cycle); smoother GDOF/s in roofline"


def workload_config(name, prob, world):
    return {"workload": name, "problem": prob.name, "finest_nodes": f"{prob.nodes(prob.max_level)}^{prob.dim}",
            "levels": f"{prob.max_level}..{prob.min_level}",
            "cycle": f"{'W' if name.endswith('_w') else 'V'}({prob.settings.num_pre},{prob.settings.num_post}) red-black GS omega={prob.settings.damping} + CG",
            "tol": prob.settings.tol, "max_iters": prob.settings.max_iters,
            "parallelism": "1 GPU" if world == 1 else f"population-sharded x{world} (one evaluation per GPU per step)",
            "l2_policy": "inputs larger than L2 (finest fields 1.1 GB each)" if prob.dim == 3 and prob.max_level >= 8
            else "working set fits L2 (latency-bound regime; no flush)"}


def run_ours(args, rank, world, local_rank):
    from evostencils_b200 import backend
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prob, prog = make_workload(args.workload)
    dev = backend.DeviceProblem(prob, device=local_rank)
    cyc = dev.build(prog)
    s = prob.settings

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------------------
    flags = ol.SOLVE_NO_GRAPH if args.no_graph else 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.2 s to start: begin before the warm-up, sample through the timed region
    for _ in range(args.warmup):
        out = cyc.solve(s.tol, s.max_iters, 1, flags)
    barrier()
    t_dev = 0.0
    launches = 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        out = cyc.solve(s.tol, s.max_iters, 1, flags)
        t_dev += out.time_ms
        launches += out.kernel_launches
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        import torch
        t = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln.item())
    evals = args.steps * world
    value = evals / (t_dev * 1e-3)
    ndof = float((prob.nodes(prob.max_level) - 2) ** prob.dim) * prob.n_fields
    cf = fitness.fitness_from_history(out.residuals, out.time_ms, s.max_iters)[1]

    # ---- end to end through the host API: build (lowered op list -> device), solve, read back --------
    h2d = len(prog.ops) * 160 + len(prog.operators) * (8 + 2 * 2 * 27 * 2 * 8)
    d2h = (s.max_iters + 1) * 8 + 48
    for _ in range(min(args.warmup, 2)):
        c2 = dev.build(prog); c2.solve(s.tol, s.max_iters, 1); c2.close()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c2 = dev.build(prog)
        o2 = c2.solve(s.tol, s.max_iters, 1)
        fitness.fitness_from_history(o2.residuals, o2.time_ms, s.max_iters)
        c2.close()
    barrier()
    t_e2e = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = evals / t_e2e

    # ---- roofline of the dominant kernel: finest-level RB-GS sweep -----------------------------------
    roof = None
    cpu = None
    if rank == 0:
        peak, peak_src = measured_peak()
        zero = (0,) * prob.dim
        sm_op = ol.Op(ol.OP_SMOOTH, prob.max_level, mode=ol.MODE_REDBLACK, omega=s.damping,
                      unknowns=tuple((f, zero) for f in range(prob.n_fields)))
        ms, n_launch = cyc.profile_op(sm_op, repeat=10)
        alg_bytes = 24.0 * ndof                      # read u, f, write u per full sweep
        achieved = alg_bytes / (ms * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes of one launch from the committed ncu --set full capture of this kernel at this size
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload, {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        roof = {"bound": "hbm", "kernel": "RB-GS sweep, finest level (both colours)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_sweep": alg_bytes, "ms_per_sweep": ms, "launches_per_sweep": n_launch,
                "smoother_gdof_s": ndof / (ms * 1e-3) / 1e9}
        other = {}
        for nm, op, b in (("residual", ol.Op(ol.OP_RESIDUAL, prob.max_level, dst=ol.BUF_RES), 24.0),
                          ("restrict", ol.Op(ol.OP_RESTRICT, prob.max_level, dst=ol.BUF_RHS, src=ol.BUF_RES),
                           8.0 + 8.0 / 2 ** prob.dim),
                          ("prolong_add", ol.Op(ol.OP_PROLONG_ADD, prob.max_level, src=ol.BUF_SOL, omega=1.0),
                           16.0 + 8.0 / 2 ** prob.dim)):
            try:
                m2, _ = cyc.profile_op(op, repeat=10)
                other[nm] = {"ms": m2, "GB/s": b * ndof / (m2 * 1e-3) / 1e9, "frac": b * ndof / (m2 * 1e-3) / 1e9 / peak}
            except Exception as e:   # pragma: no cover
                other[nm] = {"error": str(e)}
        roof["other_kernels"] = other
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_arm_sample(args.workload)
    # ---- N > 1: the same evaluation decomposed over the GPUs, reported beside the weak number ----
    dd_info = None
    line_state = {}

    def emit_line():
        if rank == 0 and not line_state.get("printed"):
            line_state["printed"] = True
            line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": t_dev / args.steps, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "domain_decomposition": dd_info,
                    "config": workload_config(args.workload, prob, world),
                    "iterations_per_eval": out.iterations, "convergence_factor": cf,
                    "cycle_gdof_s": ndof * out.iterations * evals / (t_dev * 1e-3) / 1e9,
                    "ms_per_cycle": t_dev / args.steps / max(out.iterations, 1),
                    "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                            "note": "build (op list + operator tables H2D) + solve + history D2H through the C-ABI; "
                                    "the reference-facing call carries no field data (the problem is analytic)"},
                    "gpu_launches": launches, "wall_s_timed_region": t_wall, "clocks": clocks,
                    "roofline": roof, "cpu_baseline": cpu}
            print(json.dumps(line), flush=True)

    if dist is not None and prob.dim == 3 and not args.no_domain:
        import signal
        import torch
        from evostencils_b200 import domain

        def give_up(signum, frame):   # the secondary measurement must never cost the primary line
            nonlocal dd_info
            dd_info = {"error": "domain-decomposed measurement did not finish within the time limit"}
            emit_line()
            os._exit(0)

        signal.signal(signal.SIGALRM, give_up)
        signal.alarm(int(os.environ.get("EVO_DOMAIN_TIME_LIMIT", "240")))
        cyc.close()
        solver = domain.DomainSolver.distributed(prob, prog, rank, world, local_rank)
        solver.overlap = not args.domain_no_overlap
        dd_solve = solver.solve if args.domain_eager else solver.solve_captured
        o3 = solver.solve(s.tol, s.max_iters)          # creates the NCCL communicators (not capturable)
        for _ in range(2):
            o3 = dd_solve(s.tol, s.max_iters)
        barrier()
        t_dd = 0.0
        for _ in range(args.steps):
            o3 = dd_solve(s.tol, s.max_iters)
            t_dd += o3.time_ms
        barrier()
        t = torch.tensor([t_dd], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dd = float(t.item())
        dd_info = {"value": args.steps / (t_dd * 1e-3), "unit": "evals/s", "scaling": "strong", "ms_per_eval": t_dd / args.steps,
                   "iterations": o3.iterations, "identical_history": bool(np.array_equal(o3.residuals, out.residuals)),
                   "halo_exchanges_per_eval": o3.exchanges, "overlap": solver.overlap, "levels_distributed": f"{prob.max_level}..{solver.layout.lc}",
                   "note": "ONE evaluation split into z-slabs over all GPUs, NCCL send/recv halos; "
                           + ("host-orchestrated statements" if args.domain_eager else
                              "each iteration (kernels + exchanges) replayed as one CUDA graph")}
        solver.close()
        signal.alarm(0)
    emit_line()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# ONE evaluation spread over the GPUs (z-slab domain decomposition, halo exchange over NCCL).
# Strong scaling; results are bit-identical to the single-GPU evaluation (tests/test_gpu_domain.py).
def run_domain(args, rank, world, local_rank):
    import torch
    from evostencils_b200 import domain
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prob, prog = make_workload(args.workload)
    s = prob.settings
    slabs = world if world > 1 else max(1, args.slabs)
    if world > 1:
        solver = domain.DomainSolver.distributed(prob, prog, rank, world, local_rank, lc=args.lc or None)
    else:
        solver = domain.DomainSolver.emulate(prob, prog, slabs, lc=args.lc or None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    solver.overlap = not args.domain_no_overlap
    dd_solve = solver.solve if (args.domain_eager or world == 1) else solver.solve_captured
    out = solver.solve(s.tol, s.max_iters)
    for _ in range(args.warmup):
        out = dd_solve(s.tol, s.max_iters)
    barrier()
    t_dev = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = dd_solve(s.tol, s.max_iters)
        t_dev += out.time_ms
    barrier()
    t_wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())
    value = args.steps / (t_dev * 1e-3)
    ndof = float((prob.nodes(prob.max_level) - 2) ** prob.dim)
    cf = fitness.fitness_from_history(out.residuals, out.time_ms, s.max_iters)[1]
    if rank == 0:
        cfg = workload_config(args.workload, prob, world)
        cfg["parallelism"] = (f"z-slab domain decomposition x{slabs} "
                              f"({'NCCL send/recv' if world > 1 else 'slabs emulated on one GPU'}), "
                              f"levels < {solver.layout.lc} replicated")
        line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_dev / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "iterations_per_eval": out.iterations, "convergence_factor": cf,
                "residuals_last": float(out.residuals[-1]),
                "ms_per_cycle": t_dev / args.steps / max(out.iterations, 1),
                "cycle_gdof_s": ndof * out.iterations * args.steps / (t_dev * 1e-3) / 1e9,
                "halo_exchanges_per_eval": out.exchanges, "wall_s_timed_region": t_wall, "clocks": clocks,
                "e2e": None, "gpu_launches": None, "roofline": None, "cpu_baseline": None,
                "note": "secondary mode (--domain): the default bench line is the population-sharded one"}
        print(json.dumps(line), flush=True)
    solver.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: one G3P generation = 256 evolved cycles, half on Poisson 2D (levels 5..9),
# half on LinearElasticity (levels 4..8), individuals sharded round-robin over the GPUs
# (reference: optimization/program.py:534-535 distributes `i % nprocs == rank`).
def population_individuals(n_total: int, seed: int = 0):
    import random
    from evostencils_b200 import tree
    probs = [problems.Poisson2D(5, 9), problems.LinearElasticity2D(4, 8)]
    rng = random.Random(seed)
    out = []
    for i in range(n_total):
        prob = probs[i % 2]
        out.append((i % 2, tree.random_individual(prob, rng, maximum_local_system_size=4)))
    return probs, out


def run_population(args, rank, world, local_rank):
    from evostencils_b200 import backend, tree
    from evostencils_b200.program_generator import B200ProgramGenerator
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    from evostencils_b200 import population as popmod
    n_total = args.population
    probs, individuals = population_individuals(n_total)
    mine = [individuals[i] for i in popmod.shard_indices(n_total, rank, world)]
    gens = [B200ProgramGenerator(problem=p, device=local_rank) for p in probs]
    for g in gens:
        g.initialize_code_generation(g.min_level, g.max_level)
    # host side of the hot path: string -> tree -> lowered program (done per step, inside the e2e region)
    def lower_all():
        progs = [[], []]
        for k, s in mine:
            progs[k].append(gens[k].lower(tree.build_tree(probs[k], s), gens[k].min_level))
        return progs
    progs = lower_all()

    def evaluate(progs):
        ms, res, launches0 = 0.0, [], sum(g.total_kernel_launches for g in gens)
        for k in (0, 1):
            r, t = gens[k].evaluate_population([], programs=progs[k], max_in_flight=args.in_flight)
            ms += t
            res += r
        return ms, res, sum(g.total_kernel_launches for g in gens) - launches0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        evaluate(progs)
    barrier()
    t_dev, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms, res, ln = evaluate(progs)
        t_dev += ms
        launches += ln
    barrier()
    t_wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: strings -> trees -> lowering -> build -> solve -> fitness tuples on the host
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        evaluate(lower_all())
    barrier()
    t_e2e = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([t_dev, t_wall, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_wall, t_e2e = [float(v) for v in t.tolist()]
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln.item())
    # the complete, ordered fitness list on every rank (the reference's allgather, program.py:285-291)
    order = {0: 0, 1: 0}
    local_fitness = []
    n0 = len(progs[0])
    split = {0: res[:n0], 1: res[n0:]}
    for k, _ in mine:
        local_fitness.append(split[k][order[k]])
        order[k] += 1
    all_fitness = popmod.evaluate_sharded(individuals, lambda _m: local_fitness, rank, world, dist)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU port on a bounded sample: the first 4 individuals (2 per problem), complete solves
        from oracle import oracle as orc
        t0c = time.perf_counter()
        n_cpu = 0
        for k, s_ in individuals[:4]:
            g = gens[k]
            prog = g._finalise(g.lower(tree.build_tree(probs[k], s_), g.min_level))
            orc.OracleProblem(probs[k]).build(prog).solve(probs[k].settings.tol, probs[k].settings.max_iters, 1)
            n_cpu += 1
        t_cpu = time.perf_counter() - t0c
        cpu = {"value": n_cpu / t_cpu, "unit": "evals/s", "cores": orc.num_threads(), "kind": "port",
               "sample": f"complete solves of the first {n_cpu} individuals of the generation (2 Poisson 2D, 2 elasticity) "
                         f"with the C/OpenMP oracle, one after the other, {orc.num_threads()} threads; solve time only "
                         f"(the reference additionally runs the Java generator twice and make per individual)"}
    if rank == 0:
        evals = n_total * args.steps
        converged = sum(1 for r in all_fitness if r[1] < 1)
        line = {"metric": METRIC, "value": evals / t_wall, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_wall * 1e3 / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "population", "individuals": n_total,
                           "problems": "Poisson 2D levels 9..5 (513^2) + LinearElasticity 2D levels 8..4 (257^2, 2 fields), alternating",
                           "generator": "evostencils_b200.tree.random_individual, seed 0, local systems <= 4",
                           "tol": 1e-12, "max_iters": 100, "in_flight_per_gpu": args.in_flight,
                           "parallelism": f"population sharded round-robin over {world} GPU(s)",
                           "l2_policy": "working sets fit L2 (latency / launch bound regime; no flush)"},
                "device_busy_ms_per_step": t_dev / args.steps,
                "e2e": {"value": evals / t_e2e, "unit": "evals/s",
                        "h2d_bytes_per_step": sum(len(p.ops) * 160 + len(p.operators) * 1736 for ps in progs for p in ps),
                        "d2h_bytes_per_step": len(mine) * (101 * 8 + 48),
                        "note": "grammar strings -> trees -> lowering -> evo_cycle_build -> evo_batch_solve -> fitness tuples"},
                "gpu_launches": launches, "clocks": clocks, "converging_individuals": converged,
                "roofline": None, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    # NCCL kernels of the halo exchanges should not queue behind the thousands of CTAs of an interior sweep
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poisson3d_513")
    ap.add_argument("--population", type=int, default=256)
    ap.add_argument("--in-flight", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch kernels directly (host-side solver loop) so that ncu can see them; not a bench value")
    ap.add_argument("--domain", action="store_true",
                    help="strong scaling: ONE evaluation decomposed into z-slabs over the GPUs (SURVEY.md 8e.2)")
    ap.add_argument("--domain-no-overlap", action="store_true",
                    help="domain decomposition: exchange after the whole sweep instead of boundary planes first")
    ap.add_argument("--domain-eager", action="store_true", help="domain decomposition without CUDA-graph capture")
    ap.add_argument("--no-domain", action="store_true", help="N > 1: skip the additional domain-decomposed measurement")
    ap.add_argument("--slabs", type=int, default=2, help="--domain on one GPU: number of emulated slabs")
    ap.add_argument("--lc", type=int, default=0, help="--domain: coarsest distributed level (default: automatic)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "population" and args.impl != "reference":
        run_population(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    elif args.domain:
        run_domain(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
