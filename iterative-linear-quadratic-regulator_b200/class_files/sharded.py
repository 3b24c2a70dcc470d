"""Batch data parallelism over the GPUs of one node: one process per GPU (torchrun), each rank solves a
contiguous shard of the trajectories with its own iLQR object.  Trajectories are independent
optimisations (the reference solves exactly one per iLQR object, iLQR_class.py:18-76), so no data
moves between ranks during a solve; the only exchange is an all-gather of the per-shard cost,
exit status and iteration count afterwards (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(B, rank, world):
    """Contiguous shard [lo, hi) of B items for `rank`; the first B % world ranks hold one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_shards(local, B, group=None):
    """All-gather 1-D per-trajectory values (cost, status, iterations) of every rank's shard into the
    global order.  `local` holds this rank's shard (length hi-lo of shard_bounds); returns length B."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(B, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} values, its shard has {hi - lo}")
    width = -(-B // world)                       # shards differ by at most one item: pad to the widest
    buf = local.new_zeros((width,) + tuple(local.shape[1:]))
    buf[: hi - lo] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = []
    for r, part in enumerate(parts):
        l, h = shard_bounds(B, r, world)
        out.append(part[: h - l])
    return torch.cat(out)


class ShardedILQR:
    """iLQR over a global batch sharded across the ranks of `group`.

    x_0 (B, n_x) and, if batched, U_init (B, n_u, N) are the GLOBAL arrays (every rank passes the same);
    each rank keeps rows [lo, hi) only.  optimize_trajectory() returns this rank's X, U and the global
    cost vector; `.status` / `.iterations` are global too.
    """

    def __init__(self, system, T, x_0, U_init, group=None, **kw):
        from .iLQR_class import iLQR
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.B_global = int(x_0.shape[0])
        self.lo, self.hi = shard_bounds(self.B_global, self.rank, self.world)
        U_loc = U_init[self.lo:self.hi] if len(U_init.shape) == 3 else U_init
        self.local = iLQR(system, T, x_0[self.lo:self.hi], U_loc, **kw)
        self.status = self.iterations = self.cost = None

    def optimize_trajectory(self):
        X, U, _ = self.local.optimize_trajectory()
        s = self.local
        self.cost = gather_shards(s._cost, self.B_global, self.group)
        self.status = gather_shards(s._status, self.B_global, self.group)
        self.iterations = gather_shards(s._iters, self.B_global, self.group)
        return X, U, self.cost

    @property
    def all_converged(self):
        """global stop decision from the gathered flags"""
        return bool((self.status != 3).all().item())
