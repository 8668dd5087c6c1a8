"""Probe: does the captured (CUDA graph) iteration of the domain solver work with this torch/NCCL?  Prints progress."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from evostencils_b200 import domain, cycles, problems, lowering
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
prob = problems.Poisson3D(2, level)
prog = lowering.optimise(cycles.default_solver_cycle(prob))
sv = domain.DomainSolver.distributed(prob, prog, rank, world, lr)
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
o = sv.solve(1e-12, 100); say("eager", o.iterations, o.time_ms)
with sv._streams():
    say("capturing"); sv._capture(); say("captured", len(sv._graphs), "flipping", sv._flipping)
    torch.cuda.synchronize()
o2 = sv.solve_captured(1e-12, 100); say("captured solve", o2.iterations, o2.time_ms, (o2.residuals == o.residuals).all())
o3 = sv.solve_captured(1e-12, 100); say("captured solve 2", o3.iterations, o3.time_ms, (o3.residuals == o.residuals).all())
o4 = sv.solve(1e-12, 100); say("eager again", o4.iterations, (o4.residuals == o.residuals).all())
sv.close(); say('closed'); dist.barrier(); say('barrier'); dist.destroy_process_group(); say('destroyed')
