"""`jax.lax.scan` / `while_loop` as plain Python loops (same semantics, eager)."""
import torch


def _leaves(xs):
    return list(xs) if isinstance(xs, (tuple, list)) else [xs]


def scan(f, init, xs, length=None, reverse=False):
    leaves = _leaves(xs)
    n = leaves[0].shape[0] if length is None else length
    idx = range(n - 1, -1, -1) if reverse else range(n)
    carry = init
    ys = [None] * n
    for i in idx:
        x_i = tuple(l[i] for l in leaves) if isinstance(xs, (tuple, list)) else leaves[0][i]
        carry, y = f(carry, x_i)
        ys[i] = y
    if isinstance(ys[0], (tuple, list)):
        stacked = tuple(torch.stack([torch.as_tensor(y[j]) for y in ys]) for j in range(len(ys[0])))
    else:
        stacked = torch.stack([torch.as_tensor(y) for y in ys])
    return carry, stacked


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def select(pred, on_true, on_false):
    return torch.where(torch.as_tensor(pred), torch.as_tensor(on_true), torch.as_tensor(on_false))
