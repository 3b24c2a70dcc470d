"""N>1 host logic on CPU: contiguous batch shards and the all-gather of per-shard cost / flags with the
gloo backend, world_size 2 and 3 (uneven shards)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_partition():
    from class_files.sharded import shard_bounds
    for B in (1, 7, 8, 4096, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_bounds(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [h - l for l, h in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "iterative-linear-quadratic-regulator_b200"))
    from class_files.sharded import shard_bounds, gather_shards
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(B, rank, world)
    cost = torch.arange(lo, hi, dtype=torch.float64) * 0.5 + 1.0      # stands in for the shard's solve result
    status = (torch.arange(lo, hi, dtype=torch.int32) % 3)
    g_cost = gather_shards(cost, B)
    g_status = gather_shards(status, B)
    ok = torch.equal(g_cost, torch.arange(B, dtype=torch.float64) * 0.5 + 1.0) and \
        torch.equal(g_status, torch.arange(B, dtype=torch.int32) % 3)
    # the global stop decision every rank derives from the gathered flags must agree
    n_active = torch.tensor([int((g_status == 2).sum())])
    ref = n_active.clone()
    dist.broadcast(ref, 0)
    ok = ok and bool(ref.item() == n_active.item())
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 10), (2, 4096), (3, 10)])
def test_gather_shards_gloo(world, B, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(np.load(tmp_path / f"ok{r}.npy")[0])
