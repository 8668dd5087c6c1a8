"""Host logic of the slab decomposition (no GPU): ownership arithmetic and the torch.distributed exchange
schedule (gloo, world_size 2 and 3) on CPU tensors standing in for the device arrays."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from evostencils_b200 import cycles, domain, oplist as ol, problems


@pytest.mark.parametrize("max_level,lc,world,ghost", [(9, 5, 2, 2), (9, 5, 4, 6), (9, 6, 8, 6), (7, 5, 3, 6), (6, 6, 5, 6),
                                                      (9, 5, 15, 2), (9, 8, 8, 6), (9, 7, 8, 4)])
def test_layout_partitions_every_level(max_level, lc, world, ghost):
    lay = domain.SlabLayout(max_level, lc, world, ghost)
    assert lay.ghost == ghost
    for l in range(lc, max_level + 1):
        n = (1 << l) + 1
        covered = []
        for r in range(world):
            a, b = lay.owned[l][r]
            assert b - a + 1 >= lay.ghost               # a slab can fill its neighbour's ghost planes
            covered += list(range(a, b + 1))
        assert covered == list(range(1, n - 1))         # inner planes, no gap, no overlap, ascending
        if l > lc:
            for r in range(world):                       # nested: fine planes over the rank's coarse planes
                a, b = lay.owned[l - 1][r]
                fa, fb = lay.owned[l][r]
                assert fa == 2 * a - 1 and fb in (2 * b, 2 * b + 1)
    # virtual ownership of the first replicated level covers its inner planes once and is computable locally
    nc = (1 << (lc - 1)) + 1
    covered = []
    for r in range(world):
        a, b = lay.owned[lc - 1][r]
        fa, fb = lay.owned[lc][r]
        covered += list(range(a, b + 1))
        for z in range(a, b + 1):
            assert fa - 1 <= 2 * z - 1 and 2 * z + 1 <= fb + 1   # needs one ghost plane at most
    assert covered == list(range(1, nc - 1))


def test_layout_rejects_bad_arguments():
    with pytest.raises(ValueError):
        domain.SlabLayout(6, 4, 2)
    with pytest.raises(ValueError):
        domain.SlabLayout(6, 5, 16)
    with pytest.raises(ValueError):
        domain.SlabLayout(6, 7, 2)


@pytest.mark.parametrize("ghost", [2, 4, 6])
def test_exchange_schedule_of_a_v_cycle(ghost):
    prob = problems.Poisson3D(2, 7)
    prog = cycles.v_cycle(prob, 2, 1, 1.25, True)
    lay = domain.SlabLayout(7, 6, 2, ghost)
    domain.check_supported(prog, lay)
    valid = {(7, ol.BUF_SOL): ghost, (7, ol.BUF_RHS): ghost}
    steps = domain.schedule(prog, lay, valid)
    ops = [s for s in steps if s.kind == "op"]
    halos = [(s.a, s.b) for s in steps if s.kind == "halo"]
    assert len(ops) + sum(s.kind == "gather" for s in steps) == len(prog.ops)
    assert all(l >= 6 for l, _ in halos)
    # the statements' dependences are honoured: replay the plan and check the requirement of every statement
    v = {(7, ol.BUF_SOL): ghost, (7, ol.BUF_RHS): ghost}
    get = lambda l, b: 99 if l < 6 else v.get((l, b), 0)
    for s in steps:
        if s.kind == "halo":
            v[(s.a, s.b)] = ghost
            continue
        op = prog.ops[s.a]
        l = op.level
        if s.kind == "gather":
            assert get(l, op.src) >= 1
            continue
        if l < 6:
            continue
        if op.code == ol.OP_SMOOTH:
            # a red-black sweep on the owned planes extended by e = s.b ghost planes reads e + 2 valid ghost planes of
            # SOL and e + 1 of RHS, and leaves e valid ones behind
            assert 0 <= s.b <= ghost - 2
            assert get(l, ol.BUF_SOL) >= s.b + 2 and get(l, ol.BUF_RHS) >= s.b + 1
            v[(l, ol.BUF_SOL)] = s.b
        elif op.code == ol.OP_RESIDUAL:
            assert get(l, ol.BUF_SOL) >= s.b + 1 and get(l, ol.BUF_RHS) >= s.b
            v[(l, op.dst)] = s.b
        elif op.code == ol.OP_RESTRICT:
            assert get(l, op.src) >= 1
            v[(l - 1, op.dst)] = 0
        elif op.code == ol.OP_PROLONG_ADD:
            assert get(l, ol.BUF_SOL) >= s.b and get(l - 1, op.src) >= max(1, s.b)
            v[(l, ol.BUF_SOL)] = s.b
        elif op.code == ol.OP_ZERO:
            v[(l, op.dst)] = 99
    # fewer exchanges than "after every write" (6 per level and cycle); wide ghost zones need fewer still
    assert len(halos) <= (9 if ghost == 2 else 6)
    # the second cycle starts from the validity the first one left behind and is planned deterministically
    again = domain.schedule(prog, lay, dict(valid))
    assert [repr(s) for s in again] == [repr(s) for s in domain.schedule(prog, lay, dict(valid))]


class _FakeRank:
    """CPU stand-in of SlabRank: global field g[z, y, x] = z*10000 + y*100 + x, ghosts poisoned."""

    def __init__(self, rank, layout, level):
        self.torch = torch
        self.rank, self.device, self.layout = rank, "cpu", layout
        self.info = {level: layout.local(level, rank)}
        i = self.info[level]
        n = (1 << level) + 1
        z = torch.arange(i["zoff"], i["zoff"] + i["nz"], dtype=torch.float64).view(-1, 1, 1)
        y = torch.arange(n, dtype=torch.float64).view(1, -1, 1)
        x = torch.arange(n, dtype=torch.float64).view(1, 1, -1)
        self.expected = (z * 10000 + y * 100 + x).contiguous()
        self.data = self.expected.clone()
        self.data[: i["zlo"]] = -1.0
        self.data[i["zhi"] + 1:] = -1.0

    def view(self, level, buf):
        return self.data


def _worker(rank, world, port, level, lc):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = domain.SlabLayout(level, lc, world)
        me = _FakeRank(rank, lay, level)
        comm = domain.DistComm(me, rank, world)
        comm.halo(level, ol.BUF_SOL)
        i = me.info[level]
        lo = i["zlo"] - (lay.ghost if rank > 0 else 0)
        hi = i["zhi"] + (lay.ghost if rank < world - 1 else 0)
        assert torch.equal(me.data[lo:hi + 1], me.expected[lo:hi + 1])
        # outer ghosts of the first / last rank have no neighbour: untouched
        if rank == 0:
            assert (me.data[: i["zlo"]] == -1.0).all()
        if rank == world - 1:
            assert (me.data[i["zhi"] + 1:] == -1.0).all()
        # replicated-level gather: every rank computed its own plane range
        nc = (1 << (lc - 1)) + 1
        full = torch.zeros(nc, 4, 4, dtype=torch.float64)
        a, b = lay.owned[lc - 1][rank]
        full[a:b + 1] = torch.arange(a, b + 1, dtype=torch.float64).view(-1, 1, 1)
        me.data = full
        comm.gather_planes(lc - 1, ol.BUF_RHS, lay.owned[lc - 1])
        want = torch.arange(nc, dtype=torch.float64).view(-1, 1, 1).expand(nc, 4, 4).clone()
        want[0] = 0.0
        want[-1] = 0.0
        assert torch.equal(full, want)
        # plane sums all-gather in rank order
        sizes = [q - p + 1 for (p, q) in lay.owned[level]]
        p, q = lay.owned[level][rank]
        mine = torch.arange(p, q + 1, dtype=torch.float64)
        got = comm.gather_sums([mine], sizes)
        assert torch.equal(got, torch.arange(1, (1 << level), dtype=torch.float64))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_halo_exchange(world):
    mp.spawn(_worker, args=(world, _free_port(), 6, 5), nprocs=world, join=True)
