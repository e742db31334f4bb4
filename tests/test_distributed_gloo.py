"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + the single statistics all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from momentum_b200.distributed import aggregate_solve_stats, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, rank, world)
    # each rank "solves" its shard: iterations = 10 per instance, error = instance index (checkable sums)
    its = 10.0 * (hi - lo)
    err = float(np.arange(lo, hi).sum())
    tot_it, tot_err, max_ms = aggregate_solve_stats(its, err, 5.0 + rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([lo, hi, tot_it, tot_err, max_ms]))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [65536, 1001])
def test_shard_and_aggregate_world2(tmp_path, total):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{k}.npy") for k in range(world)]
    assert r[0][0] == 0 and r[0][1] == r[1][0] and r[1][1] == total  # contiguous cover
    for k in range(world):
        assert r[k][2] == 10.0 * total and r[k][3] == total * (total - 1) / 2 and r[k][4] == 6.0


def test_shard_bounds_cover_and_balance():
    for total in (0, 1, 7, 8192, 65536):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
