"""Batches larger than one solver's buffers: BASELINE.json config 4 (n=12, m=4, N=1000, B=262144) needs
100 GB for the gains alone, so the batch is solved in sub-batches of `chunk` trajectories through ONE
iLQR object whose device buffers (state, gains, line-search candidates) are re-used for every sub-batch.
Trajectories are independent optimisations (the reference solves one per iLQR object,
iLQR_class.py:18-76), so chunking changes nothing in any result.
"""
import torch

from . import _device as D
from .iLQR_class import iLQR


def solve_chunked(system, T, x_0, U_init, chunk, phi=None, keep=("X", "U", "cost"), **ilqr_kw):
    """optimize_trajectory() for x_0 (B, n_x) in sub-batches of `chunk`.

    U_init: (n_u, N) shared or (B, n_u, N).  phi: (B,) per-trajectory phase (LTV system) or None.
    keep: which results to retain for the whole batch on the device, from {"X", "U", "K", "U_ff", "cost"}.
    Returns a dict of CUDA tensors in the reference's layout with a leading batch axis (views, no copy):
    X (B,n_x,N+1), U (B,n_u,N), K (B,N,n_u,n_x), U_ff (B,n_u,N), cost (B,), plus iterations (B,) and
    status (B,) int32 and "total_iterations".
    """
    D.require_cuda()
    tdt = D.torch_dtype(system.dtype)
    x0 = D.to_device(x_0, tdt).reshape(-1, system.n_x)
    B = x0.shape[0]
    chunk = min(int(chunk), B)
    ilqr_kw.setdefault("verbose", False)
    ph = D.to_device(phi, tdt).reshape(-1) if phi is not None else None
    U0 = D.to_device(U_init, tdt)
    sol = iLQR(system, T, x0[:chunk], U0[:chunk] if U0.ndim == 3 else U0,
               phi=ph[:chunk] if ph is not None else None, **ilqr_kw)
    n, m, N = sol.n_x, sol.n_u, sol.N
    dev = dict(dtype=tdt, device="cuda")
    out = {}
    shapes = {"X": (N + 1, n, B), "U": (N, m, B), "K": (N, m, n, B), "U_ff": (N, m, B), "cost": (B,)}
    src = {"X": "_X", "U": "_U", "K": "_K", "U_ff": "_k", "cost": "_cost"}
    for key in keep:
        out[key] = torch.empty(shapes[key], **dev)
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    status = torch.empty((B,), dtype=torch.int32, device="cuda")
    total = 0
    for lo in range(0, B, chunk):
        hi = min(lo + chunk, B)
        cnt = hi - lo
        # a short last chunk is padded with copies of its first trajectory; the padding is discarded
        idx = torch.arange(lo, lo + chunk, device="cuda").clamp_(max=hi - 1) if cnt < chunk else slice(lo, hi)
        sol._x0.copy_(x0[idx].t())
        if U0.ndim == 3:
            sol._U.copy_(U0[idx].permute(2, 1, 0))
        else:
            sol._U.copy_(U0.t().unsqueeze(-1).expand(N, m, chunk))
        if ph is not None:
            sol._phi.copy_(ph[idx])
        sol.reset_state()
        sol.solve_device(sync=True)
        total += int(sol._iters[:cnt].sum().item())
        for key in keep:
            out[key][..., lo:hi].copy_(getattr(sol, src[key])[..., :cnt])
        iters[lo:hi].copy_(sol._iters[:cnt])
        status[lo:hi].copy_(sol._status[:cnt])
    res = {"iterations": iters, "status": status, "total_iterations": total, "launches": sol.launches()}
    for key in keep:
        t = out[key]
        if key == "K":
            res[key] = t.permute(3, 0, 1, 2)
        elif key == "cost":
            res[key] = t
        else:
            res[key] = t.permute(2, 1, 0)
    return res
