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


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("iLQR (B200 build) needs a CUDA device: the solver has no CPU fallback.")


def torch_dtype(name):
    return torch.float64 if name == "float64" else torch.float32


def to_device(a, dtype):
    """numpy / list / torch (any device) -> contiguous CUDA tensor of `dtype`."""
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype)
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
