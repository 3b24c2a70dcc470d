"""OPT-IN alias, not JAX: put `iterative-linear-quadratic-regulator_b200/jax_alias` on sys.path (only where the real
jax is absent or unwanted) and a System subclass file written for the reference -- `import jax`, `import jax.numpy as
jnp`, `from .system_base import System` -- loads unchanged: `jax.numpy` is class_files.symbolic, the namespace the three
user methods are traced with (class_files/codegen.py).  Everything else the reference's system files touch at import or
construction time is a harmless stand-in; numerics never run through this module.
"""
import sys

from class_files import symbolic as numpy

sys.modules[__name__ + ".numpy"] = numpy


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(fn=None, **kw):
    """the kernels are compiled by NVRTC from the traced expressions; jit is the identity here"""
    return fn if fn is not None else (lambda f: f)


def _no_autodiff(*a, **k):
    raise NotImplementedError("derivatives come from class_files/codegen.py (analytic, from the traced expressions)")


grad = jacfwd = jacrev = hessian = _no_autodiff


# jax.lax: cond / select / switch / fori_loop / scan are traced through; while_loop is staged and becomes a real loop in
# the generated device code (class_files/symbolic.py)
lax = numpy.lax
sys.modules[__name__ + ".lax"] = lax
