"""Device plumbing shared by System and iLQR: torch tensors as raw device buffers, layout
transposes between the reference's (dim,time) arrays and the kernels' batch-innermost arrays,
and a thin handle wrapper over the C ABI.  No numerics happen here."""
import ctypes as C
import sys

import numpy as np
import torch

from . import _cabi


class HostArray(np.ndarray):
    """numpy array returned to callers; scripts call .block_until_ready() on results
    (reference run_iLQR_open_loop.py:83,93)."""

    def block_until_ready(self):
        return self


def host(a):
    return np.asarray(a).view(HostArray)


class _PinnedPool:
    """Page-locked staging buffers for device->host results.  A buffer is handed out again only
    when the caller has dropped every array that views it (refcount check), so results a caller
    keeps (e.g. X_bar of an earlier MPC tick) are never overwritten."""

    def __init__(self):
        self._bufs = {}

    def take(self, shape, dtype):
        key = (tuple(shape), dtype)
        lst = self._bufs.setdefault(key, [])
        for buf, arr in lst:
            if sys.getrefcount(arr) == 3:      # lst tuple + loop variable + getrefcount argument
                return buf, arr
        buf = torch.empty(shape, dtype=dtype, pin_memory=True)
        arr = buf.numpy()
        lst.append((buf, arr))
        if len(lst) > 8:                       # bound the pool; dropped buffers are freed once unreferenced
            lst.pop(0)
        return buf, arr


_pool = _PinnedPool()


def to_host(t):
    """CUDA tensor (any strides) -> C-contiguous HostArray through pinned memory."""
    t = t.contiguous()
    if t.numel() * t.element_size() < (1 << 16):
        return host(t.cpu().numpy())
    buf, arr = _pool.take(t.shape, t.dtype)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return arr.view(HostArray)


class PendingSolve:
    """Handle on the results of iLQR.optimize_trajectory_async()."""

    def __init__(self, owner, slot):
        self._owner, self._slot = owner, slot

    def done(self):
        return self._slot["done"].query()

    def result(self):
        """(X, U, cost) in the reference layout; blocks until the device->host copies of this solve have landed"""
        sl, sol = self._slot, self._owner.sol
        sl["done"].synchronize()
        self.total_iterations = int(sl["h_tot"][0])
        sol.total_iterations = self.total_iterations
        X, U, cost = sl["h_X"].numpy(), sl["h_U"].numpy(), sl["h_cost"].numpy()
        if not sol.batched:
            return host(X[0]), host(U[0]), host(cost[0])
        return host(X), host(U), host(cost)


class AsyncResults:
    """Two result slots (device staging in the reference layout + pinned host buffers) and a copy stream: the
    device->host copies of one solve overlap the kernels of the next (iLQR.optimize_trajectory_async)."""

    def __init__(self, sol):
        self.sol = sol
        B, n, m, N = sol.B, sol.n_x, sol.n_u, sol.N
        dev = dict(dtype=sol._tdt, device="cuda")
        self.copy_stream = torch.cuda.Stream()
        self.slots = []
        for _ in range(2):
            self.slots.append(dict(
                d_X=torch.empty((B, n, N + 1), **dev), d_U=torch.empty((B, m, N), **dev), d_cost=torch.empty((B,), **dev),
                d_tot=torch.zeros((1,), dtype=torch.int64, device="cuda"),
                h_X=torch.empty((B, n, N + 1), dtype=sol._tdt, pin_memory=True),
                h_U=torch.empty((B, m, N), dtype=sol._tdt, pin_memory=True),
                h_cost=torch.empty((B,), dtype=sol._tdt, pin_memory=True),
                h_tot=torch.zeros((1,), dtype=torch.int64, pin_memory=True),
                ready=torch.cuda.Event(), done=torch.cuda.Event(), used=False))
        self.next = 0

    def enqueue(self):
        sol, sl = self.sol, self.slots[self.next]
        self.next ^= 1
        cur = torch.cuda.current_stream()
        if sl["used"]:
            cur.wait_event(sl["done"])               # the slot's previous copies have left its device staging
        # device [time][dim][B] -> reference layout (B, dim, time), on the solve's stream
        sl["d_X"].copy_(sol._X.permute(2, 1, 0))
        sl["d_U"].copy_(sol._U.permute(2, 1, 0))
        sl["d_cost"].copy_(sol._cost)
        torch.sum(sol._iters, dim=0, keepdim=True, dtype=torch.int64, out=sl["d_tot"])
        sl["ready"].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(sl["ready"])
            for k in ("X", "U", "cost", "tot"):
                sl["h_" + k].copy_(sl["d_" + k], non_blocking=True)
            sl["done"].record(self.copy_stream)
        sl["used"] = True
        return PendingSolve(self, sl)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("iLQR (B200 build) needs a CUDA device: the solver has no CPU fallback.")


def torch_dtype(name):
    return torch.float64 if name == "float64" else torch.float32


def to_device(a, dtype):
    """numpy / list / torch (any device) -> contiguous CUDA tensor of `dtype`.  A pinned CPU tensor is copied
    asynchronously on the current stream."""
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype, non_blocking=a.device.type == "cpu" and a.is_pinned())
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=dtype).cuda()


def is_torch(a):
    return isinstance(a, torch.Tensor)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Handle:
    """Owns one ilqr_handle_t (problem constants + launch bookkeeping)."""

    def __init__(self, problem, lib=None):
        require_cuda()
        self.lib = lib if lib is not None else _cabi.load()     # lib: the generated library of a user-defined system
        self.problem = problem
        self._h = C.c_void_p()
        rc = self.lib.ilqr_create(C.byref(problem), C.byref(self._h))
        if rc == -1:
            raise ValueError("libilqr_b200 rejected the problem definition (model/integrator/dimensions)")
        _cabi.check(rc)
        self._ws = None

    def __del__(self):
        try:
            if self._h:
                self.lib.ilqr_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def h(self):
        return self._h

    def workspace(self):
        if self._ws is None:
            nbytes = self.lib.ilqr_workspace_bytes(self._h)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        return self._ws

    def launches(self):
        return int(self.lib.ilqr_launch_count(self._h))

    def check(self, rc):
        _cabi.check(rc, self._h)
