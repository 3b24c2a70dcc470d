"""A small `jax.numpy`-shaped namespace over sympy, for user-defined System subclasses.

The reference lets a user subclass `System` and write `_f_cont_fcn`, `_l_fcn`, `_l_f_fcn` with jax.numpy
(system_base.py:255-275); JAX then traces them and autodiff supplies every derivative.  A CUDA kernel cannot
call Python, so this package traces the same three methods ONCE, symbolically: written against this module

    from class_files import symbolic as jnp        # instead of: import jax.numpy as jnp

they receive arrays of sympy symbols and return sympy expressions, from which class_files/codegen.py derives
the analytic Jacobians/Hessians and generates the device code.  Only what dynamics/cost definitions use is
provided: elementwise math, array construction, @ / dot, small linear solves, data-dependent selects (where / maximum /
minimum / clip).
"""
import numpy as np
import sympy as sp

pi = float(np.pi)
float32 = np.float32
float64 = np.float64
ndarray = np.ndarray           # annotations in reference-style subclass files (`x: jnp.ndarray`)


def _obj(a):
    return a if isinstance(a, np.ndarray) else np.asarray(a, dtype=object)


def _elementwise(fn):
    def wrapped(a):
        if isinstance(a, np.ndarray):
            return np.frompyfunc(lambda v: fn(sp.sympify(v)), 1, 1)(a)
        return fn(sp.sympify(a))
    return wrapped


sin, cos, tan = _elementwise(sp.sin), _elementwise(sp.cos), _elementwise(sp.tan)
exp, log, sqrt, tanh = _elementwise(sp.exp), _elementwise(sp.log), _elementwise(sp.sqrt), _elementwise(sp.tanh)
abs = absolute = _elementwise(sp.Abs)            # noqa: A001  (mirrors jnp.abs)
square = _elementwise(lambda v: v * v)


def arctan2(y, x):
    return sp.atan2(sp.sympify(y), sp.sympify(x))


def power(a, p):
    return _obj(a) ** p if isinstance(a, np.ndarray) else sp.sympify(a) ** p


def array(v, dtype=None):
    return np.array(v, dtype=object)


asarray = array


def zeros(shape, dtype=None):
    out = np.empty(shape, dtype=object)
    out[...] = sp.Integer(0)
    return out


def ones(shape, dtype=None):
    out = np.empty(shape, dtype=object)
    out[...] = sp.Integer(1)
    return out


def eye(n, dtype=None):
    out = zeros((n, n))
    for i in range(n):
        out[i, i] = sp.Integer(1)
    return out


def diag(v):
    v = _obj(v)
    if v.ndim == 1:
        out = zeros((len(v), len(v)))
        for i, e in enumerate(v):
            out[i, i] = e
        return out
    return np.array([v[i, i] for i in range(min(v.shape))], dtype=object)


def stack(arrs, axis=0):
    return np.stack([_obj(a) for a in arrs], axis=axis)


def concatenate(arrs, axis=0):
    return np.concatenate([np.atleast_1d(_obj(a)) for a in arrs], axis=axis)


def hstack(arrs):
    return np.hstack([np.atleast_1d(_obj(a)) for a in arrs])


def vstack(arrs):
    return np.vstack([_obj(a) for a in arrs])


def dot(a, b):
    return np.dot(_obj(a), _obj(b))


def matmul(a, b):
    return np.matmul(_obj(a), _obj(b))


def sum(a, axis=None):                               # noqa: A001  (mirrors jnp.sum)
    return np.sum(_obj(a), axis=axis)


def transpose(a):
    return _obj(a).T


def where(cond, a, b):
    """jnp.where on traced values: a data-dependent SELECT becomes a Piecewise expression, which differentiates branch by
    branch and is generated as a conditional expression in the device code (the select form of lax.cond; loops with a
    data-dependent trip count, lax.while_loop, cannot be traced)."""
    def one(c, x, y):
        return sp.Piecewise((sp.sympify(x), c), (sp.sympify(y), True))
    if isinstance(cond, np.ndarray) or isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        c, x, y = np.broadcast_arrays(_obj(cond), _obj(a), _obj(b))
        return np.frompyfunc(one, 3, 1)(c, x, y)
    return one(cond, a, b)


def maximum(a, b):
    return where(_gt(a, b), a, b)


def minimum(a, b):
    return where(_gt(a, b), b, a)


def clip(a, lo, hi):
    return minimum(maximum(a, lo), hi)


def _gt(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        x, y = np.broadcast_arrays(_obj(a), _obj(b))
        return np.frompyfunc(lambda u, v: sp.StrictGreaterThan(sp.sympify(u), sp.sympify(v)), 2, 1)(x, y)
    return sp.StrictGreaterThan(sp.sympify(a), sp.sympify(b))


class linalg:
    @staticmethod
    def solve(A, b):
        """Symbolic solve of a small dense system (e.g. the 2x2 mass matrix of a double pendulum)."""
        A, b = _obj(A), _obj(b)
        M = sp.Matrix(A.tolist())
        rhs = sp.Matrix(b.tolist()) if b.ndim == 2 else sp.Matrix(b.reshape(-1, 1).tolist())
        sol = M.LUsolve(rhs)
        out = np.array(sol.tolist(), dtype=object)
        return out if b.ndim == 2 else out.reshape(-1)

    @staticmethod
    def inv(A):
        return np.array(sp.Matrix(_obj(A).tolist()).inv().tolist(), dtype=object)

    @staticmethod
    def norm(a):
        a = _obj(a).reshape(-1)
        return sp.sqrt(np.sum(a * a))
