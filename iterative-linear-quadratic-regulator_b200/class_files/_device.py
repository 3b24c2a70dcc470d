"""Device plumbing shared by System and iLQR: torch tensors as raw device buffers, layout
transposes between the reference's (dim,time) arrays and the kernels' batch-innermost arrays,
and a thin handle wrapper over the C ABI.  No numerics happen here."""
import ctypes as C

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

    def __init__(self, problem):
        require_cuda()
        self.lib = _cabi.load()
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
