"""N > 1 host path on CPU: world_size 2 over gloo -- contiguous sharding of the instance batch and the gather of
per-instance costs / statuses into global instance order (the only collective of the path)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pino_locoman_b200.sharding import gather_instance_results, shard_bounds


def _worker(rank, world, total, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, rank, world)
    # per-instance "cost" and "status" as functions of the global instance index
    idx = torch.arange(lo, hi, dtype=torch.float64)
    cost = idx * idx + 0.5
    status = torch.stack([idx, -idx], 1)
    g_cost = gather_instance_results(cost, total)
    g_status = gather_instance_results(status, total)
    np.save(os.path.join(out_dir, f"cost_{rank}.npy"), g_cost.numpy())
    np.save(os.path.join(out_dir, f"status_{rank}.npy"), g_status.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_shard_and_gather_world2(tmp_path, total):
    world = 2
    port = 29500 + (os.getpid() % 1000) + total
    mp.spawn(_worker, args=(world, total, port, str(tmp_path)), nprocs=world, join=True)
    ref = np.arange(total, dtype=np.float64)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"cost_{r}.npy"), ref * ref + 0.5)
        assert np.array_equal(np.load(tmp_path / f"status_{r}.npy"), np.stack([ref, -ref], 1))


def test_shard_bounds_cover_everything():
    for total in (1, 5, 8, 65536):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
