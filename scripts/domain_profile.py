"""Per-step timeline of the domain-decomposed cycle (run under torchrun or alone with --slabs)."""
import argparse, os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evostencils_b200 import domain, cycles, problems, oplist as ol

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=9)
ap.add_argument("--slabs", type=int, default=2)
ap.add_argument("--lc", type=int, default=0)
ap.add_argument("--cycles", type=int, default=4)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
prob = problems.Poisson3D(2, args.level)
prog = cycles.default_solver_cycle(prob)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    sv = domain.DomainSolver.distributed(prob, prog, rank, world, lr, lc=args.lc or None)
else:
    sv = domain.DomainSolver.emulate(prob, prog, args.slabs, lc=args.lc or None)
sv.solve(1e-12, 3)
names = {ol.OP_SMOOTH: "smooth", ol.OP_RESIDUAL: "residual", ol.OP_RESTRICT: "restrict", ol.OP_PROLONG_ADD: "prolong",
         ol.OP_ZERO: "zero", ol.OP_COARSE_SOLVE: "cgs", ol.OP_COPY: "copy"}
acc = collections.defaultdict(float); cnt = collections.Counter()
host_total = 0.0; gpu_total = 0.0
orig_run = sv._run
with sv._streams():
    for cyc in range(args.cycles):
        evs = []; labels = []
        def run(st):
            orig_run(st)
            e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
            if st.kind == "halo": labels.append(("halo", st.a))
            elif st.kind == "gather": labels.append(("gather", sv.program.ops[st.a].level))
            else: labels.append((names.get(sv.program.ops[st.a].code, "?"), sv.program.ops[st.a].level))
        sv._run = run
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        t0 = time.perf_counter()
        sv.cycle()
        host = time.perf_counter() - t0
        torch.cuda.synchronize()
        host_total += host; gpu_total += e0.elapsed_time(evs[-1])
        prev = e0
        for e, lab in zip(evs, labels):
            acc[lab] += prev.elapsed_time(e); cnt[lab] += 1; prev = e
sv._run = orig_run
if rank == 0:
    print(f"world {world} slabs {sv.layout.world} lc {sv.layout.lc}: host {host_total/args.cycles*1e3:.3f} ms/cycle (enqueue only), device timeline {gpu_total/args.cycles:.3f} ms/cycle")
    for lab in sorted(acc, key=lambda k: -acc[k]):
        print(f"  {lab[0]:9s} L{lab[1]}  {acc[lab]/args.cycles:8.3f} ms/cycle  ({cnt[lab]//args.cycles} per cycle)")
sv.close()
if world > 1:
    dist.destroy_process_group()
