"""`jax.scipy.linalg.lu_factor/lu_solve` (LAPACK getrf/getrs, partial pivoting)."""
import torch


def lu_factor(a):
    return torch.linalg.lu_factor(a)


def lu_solve(lu_and_piv, b):
    lu, piv = lu_and_piv
    if b.ndim == 1:
        return torch.linalg.lu_solve(lu, piv, b.unsqueeze(-1)).squeeze(-1)
    return torch.linalg.lu_solve(lu, piv, b)
