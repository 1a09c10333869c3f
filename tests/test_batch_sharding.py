"""Batch sharding of independent OCP instances over ranks (lpopc_b200/batch.py): host logic on the
gloo backend, world_size 2 (SURVEY.md 8e -- no collective on the evaluation path; a setup
broadcast of the mesh and a final gather of per-instance results only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lpopc_b200 import batch


def test_shard_range_partitions_exactly():
    for nbatch in (0, 1, 7, 8, 9, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [batch.shard_range(nbatch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nbatch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) <= (nbatch + world - 1) // world
            assert sum(sizes) == nbatch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nbatch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # setup broadcast: rank 0's mesh wins
        mesh = np.linspace(-1.0, 1.0, 9) if rank == 0 else np.array([-1.0, 0.3, 1.0])
        nodes = np.full(8, 8) if rank == 0 else np.array([3, 4])
        mp_, nd_ = batch.broadcast_mesh(mesh, nodes, dist)
        assert np.array_equal(mp_, np.linspace(-1.0, 1.0, 9)) and np.array_equal(nd_, np.full(8, 8))
        # every rank "evaluates" its contiguous shard of instances; the gather restores batch order
        lo, hi = batch.shard_range(nbatch, rank, world)
        local = np.arange(lo, hi, dtype=np.float64) ** 2 + 0.5
        full = batch.gather_results(local, nbatch, dist)
        assert np.array_equal(full, np.arange(nbatch, dtype=np.float64) ** 2 + 0.5)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nbatch", [9, 64])
def test_broadcast_and_gather_world2_gloo(tmp_path, nbatch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nbatch, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b) and a.size == nbatch


def test_single_process_paths_need_no_process_group():
    mp_, nd_ = batch.broadcast_mesh([-1.0, 1.0], [20])
    assert np.array_equal(mp_, [-1.0, 1.0]) and np.array_equal(nd_, [20])
    assert np.array_equal(batch.gather_results([1.0, 2.0], 2), [1.0, 2.0])
