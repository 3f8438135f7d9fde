"""linalg restated (cost.py:13,17; invariant.py:930)."""
import torch


def batch_quadratic_product(X, A):
    """(N,n),(n,n) -> (N,) x_i^T A x_i."""
    return torch.einsum('ij,ij->i', X @ A, X)


def batch_batch_product(X, A):
    """X (B,n), A (B,m,n) -> (B,m) = A_b x_b."""
    return (A @ X.unsqueeze(2)).squeeze(2)


def batch_outer_product(u, v):
    return torch.einsum('bi,bj->bij', u, v)
