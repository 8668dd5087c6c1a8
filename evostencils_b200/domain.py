"""Domain decomposition of ONE grid over several GPUs: z-slabs, halo exchange between statements.

Reference: the generated ExaStencils code can split a grid into blocks/fragments and exchanges ghost
layers with MPI (`communicate Solution` statements, example_problems/lib/domain_onePatch.knowledge sets one
block); every shipped configuration runs one block on one node.  Here the same lowered op
list is executed slab-wise on N GPUs, one process per GPU, with `torch.distributed` (NCCL over NVLink) moving the
ghost planes between statements.  The numerics are BIT-IDENTICAL to the undecomposed solve for any N:

* every statement is evaluated per node with the same arithmetic, red-black half sweeps are order
  independent within a colour, and ghost planes are refreshed before a statement reads them;
* residual norms use the canonical reduction (rows -> planes -> total); the plane sums are all-gathered and
  reduced in the canonical order on every rank.

Layout (mirrors `evo_problem_set_slab`, csrc/evo_runtime.cu): level `lc` splits its inner planes evenly; a
rank owning planes [a, b] of level l owns [2a-1, 2b] of level l+1 (the last rank also 2b+1).  Coarser levels
are replicated: the restriction from level lc is computed plane-wise by the owners and all-gathered, and every
rank runs the (tiny) coarse part of the cycle redundantly.  Each slab carries `ghost` planes per side (layout
parameter, default GHOST = 6).  One red-black sweep consumes two of them (two dependent half sweeps): with wide ghost
zones consecutive sweeps and the residual after them recompute the halo redundantly (a statement runs on the owned
planes extended by e ghost planes, bit-identical to what the neighbour computes) instead of exchanging after every
statement -- one exchange of 6 planes per level and cycle replaces four exchanges of 2 (communication avoiding:
the exchanges are latency, not bandwidth, bound).

`SlabLayout` and the exchange schedule are pure Python (tested on CPU with gloo); the statements themselves
run only through the CUDA library (no CPU fallback).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import oplist as ol

GHOST = int(os.environ.get("EVO_DOMAIN_GHOST", "6"))     # default ghost planes per side of a slab (even, 2..16)


class SlabLayout:
    """Ownership of global z planes per level and rank."""

    def __init__(self, max_level: int, coarsest_distributed_level: int, world: int, ghost: Optional[int] = None):
        lc = coarsest_distributed_level
        ghost = GHOST if ghost is None else int(ghost)
        if ghost < 2 or ghost > 16 or ghost % 2:
            raise ValueError("ghost planes per side: an even number in [2, 16]")
        if lc < 5 or lc > max_level:
            raise ValueError("coarsest distributed level must be in [5, max_level]")
        inner = (1 << lc) - 1
        if inner // world < ghost:
            raise ValueError("too many ranks for the coarsest distributed level")
        self.max_level, self.lc, self.world, self.ghost = max_level, lc, world, ghost
        self.owned: Dict[int, List[Tuple[int, int]]] = {}
        base, rem = divmod(inner, world)
        rows = []
        for r in range(world):
            a = 1 + r * base + min(r, rem)
            rows.append((a, a + base + (1 if r < rem else 0) - 1))
        self.owned[lc] = rows
        for l in range(lc + 1, max_level + 1):
            rows = [(2 * a - 1, 2 * b + (1 if r == world - 1 else 0)) for r, (a, b) in enumerate(self.owned[l - 1])]
            self.owned[l] = rows
        # "virtual" ownership of the first replicated level: coarse planes Z whose fine planes 2Z-1..2Z+1 the rank
        # holds (owned + one ghost plane)
        self.owned[lc - 1] = [((a + 1) // 2, b // 2) for (a, b) in self.owned[lc]]

    def distributed(self, level: int) -> bool:
        return level >= self.lc

    def local(self, level: int, rank: int) -> dict:
        """Local array geometry of a distributed level (same numbers as evo_problem_slab_info)."""
        a, b = self.owned[level][rank]
        G = self.ghost
        return {"zoff": a - G, "nz": b - a + 1 + 2 * G, "zlo": G, "zhi": G + b - a, "g0": a, "g1": b}


# --------------------------------------------------------------------------------------------------------------
# Ghost-plane validity tracking.  valid[(level, buf)] = number of ghost planes per side that hold the neighbour's
# current values.  A statement run on the owned planes extended by e ghost planes recomputes what the neighbour
# computes (same inputs, same arithmetic -> same bits) and saves an exchange.  The schedule depends only on the op
# list, so every rank takes the same decisions (the exchanges are collective).
INF = 99


class Step:
    """One scheduled action: ('halo', level, buf) or ('op', index, extension) or ('gather', index)."""
    __slots__ = ("kind", "a", "b")

    def __init__(self, kind, a, b=0):
        self.kind, self.a, self.b = kind, a, b

    def __repr__(self):
        return f"Step({self.kind}, {self.a}, {self.b})"


def schedule(program: ol.Program, layout: SlabLayout, valid: Dict[Tuple[int, int], int], overlap: bool = False) -> List[Step]:
    """Plan one pass over the op list; `valid` is updated in place (carry it from cycle to cycle)."""
    steps: List[Step] = []
    dist_ = layout.distributed
    G = layout.ghost

    def v(l, b):
        return INF if not dist_(l) else valid.get((l, b), 0)

    def need(l, b, depth):
        if v(l, b) < depth:
            steps.append(Step("halo", l, b))
            valid[(l, b)] = G

    for idx, op in enumerate(program.ops):
        c, l = op.code, op.level
        if c == ol.OP_RESTRICT and l == layout.lc:
            need(l, op.src, 1)
            steps.append(Step("gather", idx))
            continue
        if c == ol.OP_RESIDUAL_RESTRICT and dist_(l):
            # coarse plane Z reads fine residual planes 2Z-1..2Z+1, i.e. SOL planes 2Z-2..2Z+2
            need(l, ol.BUF_SOL, 2)
            need(l, ol.BUF_RHS, 1)
            if l == layout.lc:
                steps.append(Step("gather", idx))
            else:
                steps.append(Step("op", idx, 0))
                valid[(l - 1, ol.BUF_RHS)] = 0
            continue
        if not dist_(l):
            steps.append(Step("op", idx, -1))
            continue
        if c == ol.OP_SMOOTH:
            # one sweep consumes `depth` valid ghost planes of SOL (RB-GS: two dependent half sweeps) and depth - 1 of
            # RHS; whatever validity is left lets the sweep run on e extra ghost planes, which stay valid afterwards
            depth = 2 if op.mode == ol.MODE_REDBLACK else 1
            need(l, ol.BUF_SOL, depth)
            if depth == 2:
                need(l, ol.BUF_RHS, 1)
            e = max(0, min(v(l, ol.BUF_SOL) - depth, v(l, ol.BUF_RHS) - (depth - 1), G - depth))
            if overlap and e == 0:
                # boundary planes, then their exchange travels while the interior planes are swept
                steps.append(Step("smooth_overlapped", idx))
                valid[(l, ol.BUF_SOL)] = G
            else:
                steps.append(Step("op", idx, e))
                valid[(l, ol.BUF_SOL)] = e
        elif c == ol.OP_RESIDUAL:
            need(l, ol.BUF_SOL, 1)
            e = max(0, min(v(l, ol.BUF_SOL) - 1, v(l, ol.BUF_RHS), 1))
            steps.append(Step("op", idx, e))
            valid[(l, op.dst)] = e
        elif c == ol.OP_RESTRICT:
            need(l, op.src, 1)
            steps.append(Step("op", idx, 0))
            valid[(l - 1, op.dst)] = 0
        elif c == ol.OP_PROLONG_ADD:
            need(l - 1, op.src, 1)
            e = 2 if (v(l, ol.BUF_SOL) >= 2 and v(l - 1, op.src) >= 2) else (1 if v(l, ol.BUF_SOL) >= 1 else 0)
            steps.append(Step("op", idx, e))
            valid[(l, ol.BUF_SOL)] = e
        elif c == ol.OP_ZERO:
            steps.append(Step("op", idx, 0))
            valid[(l, op.dst)] = INF
        elif c == ol.OP_COPY:
            steps.append(Step("op", idx, 0))
            valid[(l, op.dst)] = v(l, op.src)
        else:
            raise ValueError(f"statement {c} on a distributed level is not supported by the slab decomposition")
    return steps


def check_supported(program: ol.Program, layout: SlabLayout):
    ok = {ol.OP_ZERO, ol.OP_COPY, ol.OP_RESIDUAL, ol.OP_SMOOTH, ol.OP_RESTRICT, ol.OP_PROLONG_ADD, ol.OP_COARSE_SOLVE,
          ol.OP_RESIDUAL_RESTRICT}
    for op in program.ops:
        if layout.distributed(op.level) and op.code not in ok:
            raise ValueError(f"statement {op.code} on a distributed level is not supported by the slab decomposition")
        if op.code == ol.OP_COARSE_SOLVE and layout.distributed(op.level):
            raise ValueError("the coarse-grid solver must run on a replicated level")


# --------------------------------------------------------------------------------------------------------------
class _DevArray:
    """`__cuda_array_interface__` view of a library-owned device array (zero copy into torch)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (ptr, False), "version": 2}


class SlabRank:
    """One rank's slab problem + cycle on its GPU."""

    def __init__(self, problem, program: ol.Program, rank: int, layout: SlabLayout, device: int, stream):
        import torch
        from .backend import DeviceProblem
        self.torch = torch
        self.rank, self.layout, self.device, self.stream = rank, layout, device, stream
        self.problem = problem
        with torch.cuda.device(device):
            self.dp = DeviceProblem(problem, device, slab=(rank, layout.world, layout.lc, layout.ghost))
            self.cycle = self.dp.build(program)
            self.cycle.set_stream(stream.cuda_stream)   # statements and exchanges share one stream per device
            self.cycle.reset()
        self.info = {l: self.dp.slab_info(l) for l in range(problem.min_level, problem.max_level + 1)}
        self._c_ops = [(op, (ol.CEvoOp * 1)(op.to_c())) for op in program.ops]

    def extended(self, level: int, e: int) -> Tuple[int, int]:
        """Owned local plane range extended by e ghost planes, clipped to the inner planes of the global grid."""
        i = self.info[level]
        n = self.problem.nodes(level)
        zin0, zin1 = max(0, 1 - i["zoff"]), min(i["nz"] - 1, n - 2 - i["zoff"])
        return max(i["zlo"] - e, zin0), min(i["zhi"] + e, zin1)

    def view(self, level: int, buf: int):
        """torch view [planes, n, pitch] of the CURRENT device array of (level, buf)."""
        i = self.info[level]
        n = self.problem.nodes(level)
        ptr = self.cycle.buffer_ptr(level, buf)
        return self.torch.as_tensor(_DevArray(ptr, (i["nz"], n, i["pitch"])), device=f"cuda:{self.device}")

    def close(self):
        self.cycle.close()
        self.dp.close()


class LocalComm:
    """All slabs live in this process (single-GPU emulation or one process driving several GPUs): copies."""

    def __init__(self, ranks: Sequence[SlabRank]):
        self.ranks = list(ranks)
        self.world = len(ranks)

    def halo(self, level: int, buf: int):
        views = [r.view(level, buf) for r in self.ranks]
        G = self.ranks[0].layout.ghost
        for r in range(self.world - 1):
            lo, hi = self.ranks[r], self.ranks[r + 1]
            il, ih = lo.info[level], hi.info[level]
            # top owned planes of r -> bottom ghosts of r+1; bottom owned planes of r+1 -> top ghosts of r
            views[r + 1][ih["zlo"] - G:ih["zlo"]].copy_(views[r][il["zhi"] - G + 1:il["zhi"] + 1])
            views[r][il["zhi"] + 1:il["zhi"] + 1 + G].copy_(views[r + 1][ih["zlo"]:ih["zlo"] + G])

    def halo_start(self, level: int, buf: int):
        self.halo(level, buf)      # copies on the one stream: nothing to overlap in the emulation
        return []

    def halo_finish(self, reqs):
        pass

    def gather_planes(self, level: int, buf: int, ranges: Sequence[Tuple[int, int]]):
        """Replicated level: rank r computed planes ranges[r]; make every copy complete."""
        views = [r.view(level, buf) for r in self.ranks]
        for src, (a, b) in enumerate(ranges):
            if b < a:
                continue
            for dst in range(self.world):
                if dst != src:
                    views[dst][a:b + 1].copy_(views[src][a:b + 1])

    def gather_sums(self, parts: Sequence, sizes: Sequence[int]):
        torch = self.ranks[0].torch
        dev = parts[0].device
        return torch.cat([p.to(dev) for p in parts])


class DistComm:
    """One slab per process: torch.distributed point-to-point (NCCL on GPUs; gloo in the CPU tests)."""

    def __init__(self, rank_state, rank: int, world: int, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.ranks = [rank_state]
        self.rank, self.world, self.group = rank, world, group

    def halo(self, level: int, buf: int):
        dist = self.dist
        me = self.ranks[0]
        G = me.layout.ghost
        v, i = me.view(level, buf), me.info[level]
        ops = []
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, v[i["zhi"] - G + 1:i["zhi"] + 1], self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, v[i["zhi"] + 1:i["zhi"] + 1 + G], self.rank + 1, self.group))
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, v[i["zlo"]:i["zlo"] + G], self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, v[i["zlo"] - G:i["zlo"]], self.rank - 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def halo_start(self, level: int, buf: int):
        """Post the exchange of (level, buf) -- buf may be BUF_NEXT, the target of an out-of-place smoother -- and return
        the requests; the current stream's work enqueued so far is what the sends wait for."""
        dist = self.dist
        me = self.ranks[0]
        G = me.layout.ghost
        v, i = me.view(level, buf), me.info[level]
        ops = []
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, v[i["zhi"] - G + 1:i["zhi"] + 1], self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, v[i["zhi"] + 1:i["zhi"] + 1 + G], self.rank + 1, self.group))
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, v[i["zlo"]:i["zlo"] + G], self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, v[i["zlo"] - G:i["zlo"]], self.rank - 1, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def halo_finish(self, reqs):
        for req in reqs:
            req.wait()

    def gather_planes(self, level: int, buf: int, ranges: Sequence[Tuple[int, int]]):
        """Replicated level: rank r computed planes ranges[r]; ONE all-gather (padded to the largest share) + one
        gather kernel make every copy complete (was: one broadcast per rank)."""
        torch = self.ranks[0].torch
        v = self.ranks[0].view(level, buf)
        counts = [max(0, b - a + 1) for (a, b) in ranges]
        m = max(counts)
        if m == 0:
            return
        key = ("planes", level, buf)
        cache = self.__dict__.setdefault("_bufs", {})
        if key not in cache:
            send = torch.zeros((m,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            recv = torch.empty((self.world * m,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            lo = min(a for (a, b) in ranges if b >= a)
            hi = max(b for (a, b) in ranges if b >= a)
            src = {}
            for r, (a, b) in enumerate(ranges):
                for z in range(a, b + 1):
                    src[z] = r * m + (z - a)
            idx = torch.tensor([src[z] for z in range(lo, hi + 1)], dtype=torch.int64, device=v.device)
            cache[key] = (send, recv, idx, lo, hi)
        send, recv, idx, lo, hi = cache[key]
        a, b = ranges[self.rank]
        if b >= a:
            send[:b - a + 1].copy_(v[a:b + 1])
        self.dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
        torch.index_select(recv, 0, idx, out=v[lo:hi + 1])

    def gather_sums(self, parts: Sequence, sizes: Sequence[int]):
        """Concatenation of every rank's plane sums in rank order: one padded all-gather + one gather kernel."""
        torch = self.ranks[0].torch
        mine = parts[0]
        m = max(sizes)
        key = ("sums", tuple(sizes))
        cache = self.__dict__.setdefault("_bufs", {})
        if key not in cache:
            send = torch.zeros(m, dtype=mine.dtype, device=mine.device)
            recv = torch.empty(self.world * m, dtype=mine.dtype, device=mine.device)
            idx = torch.tensor([r * m + k for r, sz in enumerate(sizes) for k in range(sz)], dtype=torch.int64, device=mine.device)
            cache[key] = (send, recv, idx)
        send, recv, idx = cache[key]
        send[:sizes[self.rank]].copy_(mine)
        self.dist.all_gather_into_tensor(recv, send, group=self.group)
        return recv.index_select(0, idx)


class DomainOutcome:
    def __init__(self, iterations, residuals, time_ms, exchanges):
        self.iterations = iterations
        self.residuals = np.asarray(residuals, dtype=np.float64)
        self.initial_residual = float(residuals[0])
        self.final_residual = float(residuals[-1])
        self.time_ms = time_ms
        self.exchanges = exchanges


class DomainSolver:
    """Runs the generated solver's outer loop (same semantics as evo_cycle_solve) slab-wise.

    `ranks` are the slabs this process drives: all of them with :class:`LocalComm`, exactly one with
    :class:`DistComm`."""

    def __init__(self, problem, program: ol.Program, layout: SlabLayout, ranks: Sequence[SlabRank], comm):
        check_supported(program, layout)
        self.problem, self.program, self.layout = problem, program, layout
        self.ranks, self.comm = list(ranks), comm
        self.exchanges = 0
        self.torch = self.ranks[0].torch
        self._plans = {}
        self.overlap = False        # sweep the boundary planes first and exchange them while the interior is swept
        self._reset_validity()

    def _streams(self):
        """Context: make every local slab's stream the current stream of its device."""
        import contextlib
        stack = contextlib.ExitStack()
        seen = set()
        for r in self.ranks:
            if r.device not in seen:
                seen.add(r.device)
                stack.enter_context(self.torch.cuda.stream(r.stream))
        return stack

    # -- factory helpers -------------------------------------------------------------------------------
    @classmethod
    def emulate(cls, problem, program, world: int, lc: Optional[int] = None, devices: Optional[Sequence[int]] = None,
                ghost: Optional[int] = None):
        """All `world` slabs driven by this process (device list: one entry per slab, default all on cuda:0)."""
        layout = SlabLayout(problem.max_level, lc or default_lc(problem, world, ghost), world, ghost)
        import torch
        devices = list(devices) if devices is not None else [0] * world
        streams = {dv: torch.cuda.Stream(dv) for dv in set(devices)}
        ranks = [SlabRank(problem, program, r, layout, devices[r], streams[devices[r]]) for r in range(world)]
        return cls(problem, program, layout, ranks, LocalComm(ranks))

    @classmethod
    def distributed(cls, problem, program, rank: int, world: int, device: int, lc: Optional[int] = None, group=None,
                    ghost: Optional[int] = None):
        layout = SlabLayout(problem.max_level, lc or default_lc(problem, world, ghost), world, ghost)
        import torch
        me = SlabRank(problem, program, rank, layout, device, torch.cuda.Stream(device))
        return cls(problem, program, layout, [me], DistComm(me, rank, world, group))

    def close(self):
        # captured graphs hold NCCL work: release them (and wait for the device) before the communicator goes away
        if getattr(self, "_graphs", None):
            for dv in {r.device for r in self.ranks}:
                self.torch.cuda.synchronize(dv)
            self._graphs = None
            self._keep = None
            import gc
            gc.collect()
        for r in self.ranks:
            r.close()

    # -- execution -------------------------------------------------------------------------------------
    def _run(self, st: Step):
        lay = self.layout
        if st.kind == "halo":
            self.comm.halo(st.a, st.b)
            self.exchanges += 1
        elif st.kind == "gather":
            # fine level distributed, coarse level replicated: owners compute their planes, then all-gather
            op = self.program.ops[st.a]
            ranges = lay.owned[lay.lc - 1]
            for r in self.ranks:
                a, b = ranges[r.rank]
                if b >= a:
                    r.cycle.exec_ops(r._c_ops[st.a][1], 1, a, b)
            self.comm.gather_planes(op.level - 1, ol.BUF_RHS if op.code == ol.OP_RESIDUAL_RESTRICT else op.dst, ranges)
            self.exchanges += 1
        elif st.kind == "smooth_overlapped":
            op = self.program.ops[st.a]
            lvl = op.level
            target = ol.BUF_NEXT if self._out_of_place(op) else ol.BUF_SOL
            G = lay.ghost
            for r in self.ranks:
                i = r.info[lvl]
                if i["zhi"] - i["zlo"] + 1 < 2 * G + 1:
                    r.cycle.exec_part(r._c_ops[st.a][1], i["zlo"], i["zhi"], True)
                else:
                    r.cycle.exec_part(r._c_ops[st.a][1], i["zlo"], i["zlo"] + G - 1, True)
                    r.cycle.exec_part(r._c_ops[st.a][1], i["zhi"] - G + 1, i["zhi"], True)
            reqs = self.comm.halo_start(lvl, target)
            self.exchanges += 1
            for r in self.ranks:
                i = r.info[lvl]
                if i["zhi"] - i["zlo"] + 1 < 2 * G + 1:
                    r.cycle.exec_part(r._c_ops[st.a][1], 1, 0, False)          # only the exchange of the slots
                else:
                    r.cycle.exec_part(r._c_ops[st.a][1], i["zlo"] + G, i["zhi"] - G, False)
            self.comm.halo_finish(reqs)
        else:
            op = self.program.ops[st.a]
            for r in self.ranks:
                if st.b > 0:
                    lo, hi = r.extended(op.level, st.b)
                    r.cycle.exec_ops(r._c_ops[st.a][1], 1, lo, hi)
                else:
                    r.cycle.exec_ops(r._c_ops[st.a][1], 1)

    def _run_cycle_eager(self):
        plan = self._plans.get(self._valid_key())
        if plan is None:
            key = self._valid_key()
            valid = dict(self.valid)
            plan = (schedule(self.program, self.layout, valid, self.overlap), valid)
            self._plans[key] = plan
        steps, after = plan
        for st in steps:
            self._run(st)
        self.valid = dict(after)

    def cycle(self):
        plan = self._plans.get(self._valid_key())
        if plan is None:
            key = self._valid_key()
            valid = dict(self.valid)
            plan = (schedule(self.program, self.layout, valid, self.overlap), valid)
            self._plans[key] = plan
        steps, after = plan
        with self._streams():
            for st in steps:
                self._run(st)
        self.valid = dict(after)

    @staticmethod
    def _out_of_place(op: ol.Op) -> bool:
        """Pointwise smoothers on distributed levels write the [next] slot (Jacobi and the streaming RB-GS kernel)."""
        return op.code == ol.OP_SMOOTH

    def _valid_key(self):
        return (self.overlap,) + tuple(sorted(self.valid.items()))

    def _reset_validity(self):
        top = self.problem.max_level
        G = self.layout.ghost
        self.valid = {(top, ol.BUF_SOL): G, (top, ol.BUF_RHS): G}   # the upload filled every local plane

    def residual_norm(self) -> float:
        torch = self.torch
        with self._streams():
            top_ = self.problem.max_level
            if self.valid.get((top_, ol.BUF_SOL), 0) < 1:
                self.comm.halo(top_, ol.BUF_SOL)
                self.exchanges += 1
                self.valid[(top_, ol.BUF_SOL)] = self.layout.ghost
            self.valid[(top_, ol.BUF_RES)] = 0
            parts = []
            for r in self.ranks:
                ptr, n = r.cycle.residual_plane_sums()
                parts.append(torch.as_tensor(_DevArray(ptr, (n,)), device=f"cuda:{r.device}"))
            top = self.problem.max_level
            sizes = [b - a + 1 for (a, b) in self.layout.owned[top]]
            full = self.comm.gather_sums(parts, sizes).contiguous()
            assert full.numel() == self.problem.nodes(top) - 2
            s = self.ranks[0].cycle.vecsum(full.data_ptr(), full.numel())
        return math.sqrt(s)

    # -- captured execution: one CUDA graph per iteration (kernels + NCCL exchanges) -------------------------------
    def _norm_partial(self):
        """Residual on the owned planes + plane sums + all-gather + canonical total, all stream ordered; the sum is
        read with read_sum() afterwards."""
        torch = self.torch
        top = self.problem.max_level
        if self.valid.get((top, ol.BUF_SOL), 0) < 1:
            self.comm.halo(top, ol.BUF_SOL)
            self.exchanges += 1
            self.valid[(top, ol.BUF_SOL)] = self.layout.ghost
        self.valid[(top, ol.BUF_RES)] = 0
        parts = []
        for r in self.ranks:
            ptr, n = r.cycle.residual_plane_sums()
            parts.append(torch.as_tensor(_DevArray(ptr, (n,)), device=f"cuda:{r.device}"))
        sizes = [b - a + 1 for (a, b) in self.layout.owned[top]]
        full = self.comm.gather_sums(parts, sizes).contiguous()
        self._keep = full                      # the captured graph reads this buffer on every replay
        self.ranks[0].cycle.vecsum_async(full.data_ptr(), full.numel())

    def _canonical_validity(self):
        top = self.problem.max_level
        G = self.layout.ghost
        self.valid = {(top, ol.BUF_SOL): G, (top, ol.BUF_RHS): G}

    def _capture(self):
        """Two graphs: an iteration starting from the canonical buffer assignment (A) and one starting from the
        assignment A leaves behind (B); statements that work out of place exchange SOL and its [next] slot."""
        torch = self.torch
        if len(self.ranks) != 1:
            raise RuntimeError("captured execution is for one slab per process")
        r = self.ranks[0]
        levels = range(self.layout.lc, self.problem.max_level + 1)
        self._graphs = []
        self._graph_exchanges = []             # halo exchanges / gathers one replay of each graph performs
        before = {l: r.cycle.buffer_ptr(l, ol.BUF_SOL) for l in levels}
        for k in range(2):
            g = torch.cuda.CUDAGraph()
            self._canonical_validity()
            e0 = self.exchanges
            with torch.cuda.graph(g, stream=r.stream):
                self._run_cycle_eager()
                self._norm_partial()
            self._graphs.append(g)
            self._graph_exchanges.append(self.exchanges - e0)
            if k == 0:
                self._flipping = [l for l in levels if r.cycle.buffer_ptr(l, ol.BUF_SOL) != before[l]]
                if not self._flipping:
                    self._graphs.append(g)     # nothing swaps: one graph serves every iteration
                    self._graph_exchanges.append(self._graph_exchanges[0])
                    break
        after = {l: r.cycle.buffer_ptr(l, ol.BUF_SOL) for l in levels}
        if after != before:
            raise RuntimeError("two cycles did not return to the canonical buffer assignment")

    def solve_captured(self, tol: float, max_iters: int) -> DomainOutcome:
        """Same loop as solve(), each iteration one graph launch (host cost per iteration: a launch and a read-back
        instead of ~60 statement / exchange calls).  NCCL communicators must exist (run solve() once before)."""
        torch = self.torch
        r = self.ranks[0]
        with self._streams():
            if getattr(self, "_odd", False):   # undo the host-side exchange of the previous (odd) solve
                for l in self._flipping:
                    r.cycle.swap_slots(l)
                self._odd = False
            if not getattr(self, "_graphs", None):
                self._capture()
            r.cycle.reset()
            self._reset_validity()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.exchanges = 0
            torch.cuda.synchronize(r.device)
            ev0.record()
            hist = [self.residual_norm()]
            it = 0
            while it < max_iters and math.isfinite(hist[-1]):
                self._graphs[it & 1].replay()
                self.exchanges += self._graph_exchanges[it & 1]
                it += 1
                hist.append(math.sqrt(r.cycle.read_sum()))
                if hist[-1] < tol * hist[0]:
                    break
            ev1.record()
            torch.cuda.synchronize(r.device)
            if it & 1:
                for l in self._flipping:   # the data of an odd solve lives in the other buffers
                    r.cycle.swap_slots(l)
                self._odd = True
            self._canonical_validity()
        return DomainOutcome(it, hist, ev0.elapsed_time(ev1), self.exchanges)

    def solve(self, tol: float, max_iters: int) -> DomainOutcome:
        """`repeat until res < tol * res0 or it >= maxIts` (2D_FD_Poisson_fromL2.exa3:3-4), timed on the device."""
        torch = self.torch
        with self._streams():
            for r in self.ranks:
                r.cycle.reset()
            self._reset_validity()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.exchanges = 0
            for dv in {r.device for r in self.ranks}:
                torch.cuda.synchronize(dv)
            ev0.record()
            hist = [self.residual_norm()]
            it = 0
            res0 = hist[0]
            while it < max_iters and math.isfinite(hist[-1]):
                self.cycle()
                it += 1
                hist.append(self.residual_norm())
                if hist[-1] < tol * res0:
                    break
            ev1.record()
            for dv in {r.device for r in self.ranks}:
                torch.cuda.synchronize(dv)
        return DomainOutcome(it, hist, ev0.elapsed_time(ev1), self.exchanges)

    def gather_solution(self) -> np.ndarray:
        """Dense finest-level solution assembled from the owned planes of the local slabs (tests)."""
        p = self.problem
        n = p.nodes(p.max_level)
        out = np.full((n, n, n), np.nan)
        for dv in {r.device for r in self.ranks}:
            self.torch.cuda.synchronize(dv)
        for r in self.ranks:
            i = r.info[p.max_level]
            v = r.view(p.max_level, ol.BUF_SOL)
            lo, hi = i["zlo"], i["zhi"]
            if r.rank == 0:
                lo -= 1
            if r.rank == self.layout.world - 1:
                hi += 1
            out[lo + i["zoff"]:hi + 1 + i["zoff"]] = v[lo:hi + 1, :, :n].cpu().numpy()
        return out


def default_lc(problem, world: int, ghost: Optional[int] = None) -> int:
    """Coarsest distributed level: only the two finest levels are split (measured at 513^3 on 2 GPUs: lc = 8 31.3 ms,
    7 31.9 ms, 5 36.2 ms per evaluation -- exchanges on small levels cost more than computing them redundantly),
    at least 8 planes per rank, never below level 5 (33^3)."""
    per_rank = max(8, GHOST if ghost is None else ghost)
    lc = max(5, problem.max_level - 1)
    while lc < problem.max_level and ((1 << lc) - 1) < per_rank * world:
        lc += 1
    return min(lc, problem.max_level)
