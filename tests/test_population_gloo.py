"""N>1 host path on CPU: world_size-2 gloo process group, population sharded round-robin, fitness tuples
gathered on every rank (no GPU needed: the evaluator is a deterministic stand-in)."""
import os
import sys

import pytest

from evostencils_b200 import population

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_and_merge_roundtrip():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            shards = [[f"item{i}" for i in population.shard_indices(n, r, world)] for r in range(world)]
            assert sum(len(s) for s in shards) == n
            assert population.merge_shards(shards, n) == [f"item{i}" for i in range(n)]


def _worker(rank, world, port, n_items, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from evostencils_b200 import population as pop
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    items = [f"individual_{i}" for i in range(n_items)]
    seen = []

    def evaluate_local(mine):
        seen.extend(mine)
        return [(float(len(s)), 0.5 + int(s.split("_")[1]) / 1000.0, float(rank)) for s in mine]
    res = pop.evaluate_sharded(items, evaluate_local, rank, world, dist)
    assert seen == [items[i] for i in pop.shard_indices(n_items, rank, world)]
    assert len(res) == n_items
    for i, (t, cf, r) in enumerate(res):
        assert cf == 0.5 + i / 1000.0 and r == float(i % world)
    with open(os.path.join(out_dir, f"ok_{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_gather(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 11, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok_0", "ok_1"]
