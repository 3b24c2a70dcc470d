"""`jax.numpy` subset used by the reference, on float64 torch tensors."""
import math
import numpy as _np
import torch

pi = math.pi
newaxis = None
ndarray = torch.Tensor
float64 = torch.float64


def _to_tensor(v):
    if isinstance(v, torch.Tensor):
        return v.to(torch.float64) if not v.dtype.is_floating_point else v
    if isinstance(v, (list, tuple)):
        if len(v) == 0:
            return torch.zeros(0, dtype=torch.float64)
        return torch.stack([_to_tensor(e) for e in v])
    return torch.as_tensor(_np.asarray(v, dtype=_np.float64))


def array(v, dtype=None):
    return _to_tensor(v)


asarray = array


def zeros(shape, dtype=None):
    return torch.zeros(shape, dtype=torch.float64)


def ones(shape, dtype=None):
    return torch.ones(shape, dtype=torch.float64)


def zeros_like(a):
    return torch.zeros_like(_to_tensor(a))


def eye(n):
    return torch.eye(n, dtype=torch.float64)


def diag(v):
    return torch.diag(_to_tensor(v))


def arange(start, stop=None, step=1):
    # numpy length rule ceil((stop-start)/step), same as jnp.arange
    return torch.as_tensor(_np.arange(start, stop, step, dtype=_np.float64))


def sin(x):
    return torch.sin(_to_tensor(x))


def cos(x):
    return torch.cos(_to_tensor(x))


def exp(x):
    return torch.exp(_to_tensor(x))


def log(x):
    return torch.log(_to_tensor(x))


def sqrt(x):
    return torch.sqrt(_to_tensor(x))


def tanh(x):
    return torch.tanh(_to_tensor(x))


def tan(x):
    return torch.tan(_to_tensor(x))


def concatenate(seq, axis=0):
    return torch.cat([_to_tensor(s) for s in seq], dim=axis)


def vstack(seq):
    return torch.vstack([_to_tensor(s) for s in seq])


def repeat(a, n):
    return _to_tensor(a).reshape(-1).repeat_interleave(n)


def abs(x):  # noqa: A001
    return torch.abs(_to_tensor(x))


class linalg:
    @staticmethod
    def solve(a, b):
        return torch.linalg.solve(a, b)

    @staticmethod
    def norm(x):
        return torch.linalg.norm(x)
