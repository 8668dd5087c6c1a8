#!/usr/bin/env python
"""Per-statement cost of one cycle: every distinct statement of the op list timed alone (evo_cycle_profile_op,
CUDA events, 20 repeats) x its multiplicity.  Usage: op_profile.py {poisson3d|poisson2d|fas2d|elasticity2d} MAXLEVEL [MINLEVEL]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evostencils_b200 import backend, cycles, lowering, oplist as ol, problems  # noqa: E402

kind = sys.argv[1]
hi = int(sys.argv[2])
lo = int(sys.argv[3]) if len(sys.argv) > 3 else None
if kind == "poisson3d":
    prob = problems.Poisson3D(lo or 2, hi); prog = lowering.optimise(cycles.default_solver_cycle(prob))
elif kind == "poisson2d":
    prob = problems.Poisson2D(lo or hi - 4, hi); prog = lowering.optimise(cycles.default_solver_cycle(prob))
elif kind == "elasticity2d":
    prob = problems.LinearElasticity2D(lo or hi - 4, hi); prog = lowering.optimise(cycles.default_solver_cycle(prob))
elif kind == "fas2d":
    prob = problems.FAS2D(lo or hi - 4, hi); prog = cycles.fas_v_cycle(prob)
else:
    raise SystemExit("unknown problem")
names = {v: k for k, v in vars(ol).items() if k.startswith("OP_") and isinstance(v, int)}
groups = collections.OrderedDict()
for op in prog.ops:
    key = (op.code, op.level, op.mode, op.kind, op.count, op.omega, op.dst, op.src, tuple(op.unknowns or ()))
    groups.setdefault(key, [op, 0])[1] += 1
cyc = backend.DeviceProblem(prob).build(prog)
total = 0.0
rows = []
for key, (op, n) in groups.items():
    try:
        ms, launches = cyc.profile_op(op, repeat=20)
    except backend.BackendError as e:
        ms, launches = float("nan"), 0
    rows.append((ms * n, names.get(op.code, op.code), op.level, n, ms, launches))
    total += ms * n
s = prob.settings
out = cyc.solve(s.tol, s.max_iters, 3)
print(f"{prob.name} levels {prob.max_level}..{prob.min_level}: sum of statements {total:.3f} ms/cycle; solve {out.time_ms:.2f} ms, "
      f"{out.iterations} iterations -> {out.time_ms / max(out.iterations, 1):.3f} ms per iteration (cycle + residual norm)")
for tot, name, lvl, n, ms, launches in sorted(rows, reverse=True)[:18]:
    print(f"  {name:22s} L{lvl:<2d} x{n:<3d} {ms:8.4f} ms each ({launches} launches)  {tot:8.4f} ms  {100 * tot / total:5.1f}%")
