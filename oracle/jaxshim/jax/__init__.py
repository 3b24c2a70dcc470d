"""Minimal `jax` API stand-in over torch.func (float64, CPU). TEST INFRASTRUCTURE ONLY.
See ../README.md.  Only the symbols the reference imports are provided:
system_base.py:1-4, iLQR_class.py:1-3, pendulum_sys.py:1-6."""
import torch
import torch.func as _tf

torch.set_default_dtype(torch.float64)

from . import numpy  # noqa: E402,F401
from . import lax    # noqa: E402,F401
from . import scipy  # noqa: E402,F401


def jit(fn=None, **_kw):
    # tracing compiler is irrelevant to results: identity
    if fn is None:
        return lambda f: f
    return fn


def grad(fn, argnums=0):
    return _tf.grad(fn, argnums=argnums)


def jacfwd(fn, argnums=0):
    return _tf.jacfwd(fn, argnums=argnums)


def jacrev(fn, argnums=0):
    return _tf.jacrev(fn, argnums=argnums)


def hessian(fn, argnums=0):
    # jax.hessian = jacfwd(jacrev(f))
    return _tf.jacfwd(_tf.jacrev(fn, argnums=argnums), argnums=argnums)


def _block_until_ready(self):
    return self


# run scripts call .block_until_ready() on returned arrays
torch.Tensor.block_until_ready = _block_until_ready
