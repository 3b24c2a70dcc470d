"""A small `jax.numpy`-shaped namespace over sympy, for user-defined System subclasses.

The reference lets a user subclass `System` and write `_f_cont_fcn`, `_l_fcn`, `_l_f_fcn` with jax.numpy
(system_base.py:255-275); JAX then traces them and autodiff supplies every derivative.  A CUDA kernel cannot
call Python, so this package traces the same three methods ONCE, symbolically: written against this module

    from class_files import symbolic as jnp        # instead of: import jax.numpy as jnp

they receive arrays of sympy symbols and return sympy expressions, from which class_files/codegen.py derives
the analytic Jacobians/Hessians and generates the device code.  Only what dynamics/cost definitions use is
provided: elementwise math, array construction, @ / dot, small linear solves, data-dependent selects (where / maximum /
minimum / clip) and the structured control flow of `jax.lax` (`lax.cond`, `select`, `switch`, `fori_loop`, `scan`, and
`while_loop` with a data-dependent trip count, which becomes a real loop in the device code -- see `lax` below).
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
arctan, arcsin, arccos = _elementwise(sp.atan), _elementwise(sp.asin), _elementwise(sp.acos)
sinh, cosh = _elementwise(sp.sinh), _elementwise(sp.cosh)
# sign as a select (its derivative is 0 on both sides, as jax has it; sympy's own sign differentiates to a DiracDelta)
sign = _elementwise(lambda v: sp.Piecewise((sp.Integer(1), v > 0), (sp.Integer(-1), v < 0), (sp.Integer(0), True)))


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


def reshape(a, shape):
    return _obj(a).reshape(shape)


def zeros_like(a, dtype=None):
    return zeros(np.shape(a))


def ones_like(a, dtype=None):
    return ones(np.shape(a))


def outer(a, b):
    return np.outer(_obj(a), _obj(b))


def cross(a, b):
    a, b = _obj(a), _obj(b)
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=object)


def trace(a):
    a = _obj(a)
    return np.sum(np.array([a[i, i] for i in range(min(a.shape))], dtype=object))


def mean(a, axis=None):
    return np.mean(_obj(a), axis=axis)


def prod(a, axis=None):
    return np.prod(_obj(a), axis=axis)


def where(cond, a, b):
    """jnp.where on traced values: a data-dependent SELECT becomes a Piecewise expression, which differentiates branch by
    branch and is generated as a conditional expression in the device code (the select form of lax.cond)."""
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
    def det(A):
        return sp.Matrix(_obj(A).tolist()).det(method="berkowitz")

    @staticmethod
    def norm(a):
        a = _obj(a).reshape(-1)
        return sp.sqrt(np.sum(a * a))


def logical_and(a, b):
    return sp.And(a, b)


def logical_or(a, b):
    return sp.Or(a, b)


def logical_not(a):
    return sp.Not(a)


# ------------------------------------------------------------------------------------------ jax.lax
# The reference's own integrator uses lax.while_loop (system_base.py:139) and a user may write one inside a dynamics
# function (a Newton or fixed-point iteration with a data-dependent trip count).  Unrolling such a loop into nested
# selects is exponential for a symbolic tracer, so a while_loop is STAGED instead: cond_fun and body_fun are traced once
# with fresh symbols for the carry, the loop's results enter the surrounding expressions as opaque leaves (LoopLeaf),
# and class_files/codegen.py emits a genuine `while` loop in the device code that carries, beside the values, their
# tangents with respect to x and u (forward mode: T <- (d body / d carry) T + d body / d(x,u)) -- what jax.jacfwd does
# with a while_loop.  Derivatives of a leaf are the loop's final tangents (LoopTangent).
class LoopLeaf(sp.Function):
    """leaf `i` of the final carry of loop `id`, as a function of the trace's base symbols (x..., u...)"""

    @classmethod
    def eval(cls, *args):
        return None

    def _eval_is_real(self):
        return True

    def fdiff(self, argindex=1):
        lp = _loop_by_id(int(self.args[0]))
        i, v = int(self.args[1]), argindex - 3
        if v < 0:
            raise sp.function.ArgumentIndexError(self, argindex)
        if not lp.nz[i][v]:
            return sp.Integer(0)              # structurally zero: a trip counter, or a loop that never sees this symbol
        return LoopTangent(*self.args[:2], sp.Integer(v), *self.args[2:])


class LoopTangent(sp.Function):
    """d LoopLeaf(id, i) / d base symbol v"""

    @classmethod
    def eval(cls, *args):
        return None

    def _eval_is_real(self):
        return True

    def fdiff(self, argindex=1):
        raise NotImplementedError(
            "second derivatives through lax.while_loop are not generated: a data-dependent loop may appear in "
            "_f_cont_fcn (first derivatives) but not inside _l_fcn / _l_f_fcn, whose Hessians the backward pass needs "
            "(the reference cannot either: its grad / hessian are reverse mode, which JAX does not define for while_loop)")


class _Loop:
    def __init__(self, ident, base, carry, init, cond, body, rebuild):
        self.id, self.base, self.carry, self.init, self.cond, self.body = ident, base, carry, init, cond, body
        # nz[i][v]: can leaf i depend on base symbol v -- directly, through an earlier loop's leaves (their fdiff knows) or
        # through another leaf of this loop?  A fixed point over the structural zeros of the partial derivatives; only
        # these (leaf, symbol) pairs get tangent variables in the generated loop (a trip counter gets none at all).
        L, V = len(carry), len(base)
        nz = [[sp.diff(init[i], base[v]) != 0 or sp.diff(body[i], base[v]) != 0 for v in range(V)] for i in range(L)]
        dep = [[sp.diff(body[i], carry[j]) != 0 for j in range(L)] for i in range(L)]
        changed = True
        while changed:
            changed = False
            for i in range(L):
                for v in range(V):
                    if not nz[i][v] and any(dep[i][j] and nz[j][v] for j in range(L)):
                        nz[i][v] = changed = True
        self.nz = nz
        self.active = [any(row) for row in nz]
        self.leaves = [LoopLeaf(sp.Integer(ident), sp.Integer(i), *base) for i in range(len(carry))]
        self.result = rebuild(self.leaves)


class _Trace:
    def __init__(self, base):
        self.base, self.loops, self.depth = list(base), [], 0


_TRACE = None
_ALL_LOOPS = {}          # by a process-wide id (sympy caches derivatives by expression, so ids are never reused);
                         # the generated code numbers the loops of one system from 0 (codegen._Printer)


def _loop_by_id(ident):
    return _ALL_LOOPS[ident]


def begin_trace(base_symbols):
    """codegen: the symbols every traced expression is a function of; loops met until end_trace() are collected"""
    global _TRACE
    _TRACE = _Trace(base_symbols)


def end_trace():
    global _TRACE
    t, _TRACE = _TRACE, None
    return t.loops if t is not None else []


def _flatten(tree):
    """-> (leaves, rebuild): tuples / lists / dicts / object arrays of scalars"""
    if isinstance(tree, (tuple, list)):
        parts = [_flatten(t) for t in tree]
        sizes = [len(p[0]) for p in parts]

        def rebuild(leaves, kind=type(tree)):
            out, o = [], 0
            for (_, rb), k in zip(parts, sizes):
                out.append(rb(leaves[o:o + k]))
                o += k
            return kind(out)
        return [l for p in parts for l in p[0]], rebuild
    if isinstance(tree, dict):
        keys = list(tree)
        leaves, rb = _flatten([tree[k] for k in keys])
        return leaves, lambda lv: dict(zip(keys, rb(lv)))
    if isinstance(tree, np.ndarray):
        shape = tree.shape

        def rebuild_arr(leaves):
            out = np.empty(len(leaves), dtype=object)
            for i, l in enumerate(leaves):
                out[i] = l
            return out.reshape(shape)
        return [sp.sympify(v) for v in tree.reshape(-1)], rebuild_arr
    return [sp.sympify(tree)], lambda lv: lv[0]


def _select_tree(c, a, b):
    la, rb = _flatten(a)
    lb, _ = _flatten(b)
    if len(la) != len(lb):
        raise TypeError("lax.cond / select: both branches must return the same structure")
    return rb([where(c, x, y) for x, y in zip(la, lb)])


class lax:
    """`jax.lax` for traced user methods"""

    @staticmethod
    def select(pred, on_true, on_false):
        return where(pred, on_true, on_false)

    @staticmethod
    def cond(pred, true_fun, false_fun, *operands):
        """both branches are traced and the results selected leaf by leaf (what lax.cond becomes under vmap as well)"""
        if isinstance(pred, (bool, np.bool_)):
            return true_fun(*operands) if pred else false_fun(*operands)
        return _select_tree(pred, true_fun(*operands), false_fun(*operands))

    @staticmethod
    def switch(index, branches, *operands):
        if isinstance(index, (int, np.integer)):
            return branches[min(max(int(index), 0), len(branches) - 1)](*operands)
        out = branches[-1](*operands)
        for i in range(len(branches) - 2, -1, -1):
            c = sp.Le(sp.sympify(index), i) if i == 0 else sp.Eq(sp.sympify(index), i)
            out = _select_tree(c, branches[i](*operands), out)
        return out

    @staticmethod
    def fori_loop(lower, upper, body_fun, init_val):
        """static bounds: unrolled at trace time; traced bounds: a while_loop over (i, value)"""
        if isinstance(lower, (int, np.integer)) and isinstance(upper, (int, np.integer)):
            val = init_val
            for i in range(int(lower), int(upper)):
                val = body_fun(i, val)
            return val
        return lax.while_loop(lambda c: c[0] < upper, lambda c: (c[0] + 1, body_fun(c[0], c[1])), (lower, init_val))[1]

    @staticmethod
    def scan(f, init, xs=None, length=None, reverse=False):
        """unrolled over the leading axis (trace-time length); xs and the per-step outputs may be tuples / lists of arrays;
        -> (carry, stacked ys)"""
        multi = isinstance(xs, (tuple, list))
        n = int(length) if xs is None else (len(xs[0]) if multi else len(xs))
        order = range(n - 1, -1, -1) if reverse else range(n)
        carry, ys = init, [None] * n
        for i in order:
            x_i = None if xs is None else (type(xs)(a[i] for a in xs) if multi else xs[i])
            carry, ys[i] = f(carry, x_i)
        if n == 0 or ys[0] is None:
            return carry, None
        if isinstance(ys[0], (tuple, list)):
            return carry, type(ys[0])(np.stack([_obj(y[k]) for y in ys]) for k in range(len(ys[0])))
        return carry, np.stack([_obj(y) for y in ys])

    @staticmethod
    def while_loop(cond_fun, body_fun, init_val):
        """a loop whose trip count depends on the data: staged (see above), a real `while` in the generated code"""
        t = _TRACE
        if t is None:
            raise RuntimeError("lax.while_loop outside a traced System method")
        if t.depth:
            raise NotImplementedError("nested lax.while_loop")
        init, rebuild = _flatten(init_val)
        ident = len(_ALL_LOOPS)
        carry = [sp.Symbol(f"wl{ident}_c{i}", real=True) for i in range(len(init))]
        t.depth += 1
        try:
            cond = sp.sympify(cond_fun(rebuild(carry)))
            body, _ = _flatten(body_fun(rebuild(carry)))
        finally:
            t.depth -= 1
        if len(body) != len(init):
            raise TypeError("lax.while_loop: body_fun must return the structure of init_val")
        lp = _Loop(ident, t.base, carry, init, cond, body, rebuild)
        _ALL_LOOPS[ident] = lp
        t.loops.append(lp)
        return lp.result
