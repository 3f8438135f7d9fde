"""tensor_utils: decorators that flatten batch dims by VIEW (aliasing matters)."""
import functools
import numpy as np
import torch


def is_tensor_like(x):
    return torch.is_tensor(x) or type(x) is np.ndarray


def ensure_2d_input(func):
    """1-D tensor args of a METHOD become (1,-1) views (online_controller.py:60, invariant.py)."""

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        args = [v.view(1, -1) if (torch.is_tensor(v) and v.dim() == 1) else
                (v.reshape(1, -1) if (type(v) is np.ndarray and v.ndim == 1) else v) for v in args]
        return func(self, *args, **kwargs)

    return wrapper


def handle_batch_input(func):
    """Flatten all leading dims of >2-D tensor args into one (by view), call, restore leading dims."""

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        batch_dims = None
        for arg in args:
            if is_tensor_like(arg) and len(arg.shape) > 2:
                batch_dims = tuple(arg.shape[:-1])
                break
        if batch_dims is None:
            return func(*args, **kwargs)
        args = [(v.view(-1, v.shape[-1]) if torch.is_tensor(v) else v.reshape(-1, v.shape[-1]))
                if (is_tensor_like(v) and len(v.shape) > 2) else v for v in args]
        ret = func(*args, **kwargs)

        def restore(v):
            if not is_tensor_like(v) or len(v.shape) == 0:
                return v
            if len(v.shape) == 2:
                return v.reshape(*batch_dims, v.shape[-1])
            return v.reshape(*batch_dims)

        if type(ret) is tuple:
            return tuple(restore(v) for v in ret)
        return restore(ret)

    return wrapper


def ensure_tensor(device, dtype, *args):
    tensors = tuple(x.to(device=device, dtype=dtype) if torch.is_tensor(x) else
                    torch.tensor(np.asarray(x), device=device, dtype=dtype) for x in args)
    return tensors if len(tensors) > 1 else tensors[0]


def ensure_diagonal(v, n):
    """scalar -> v*I_n ; vector -> diag ; matrix returned as tensor."""
    if torch.is_tensor(v):
        t = v
    else:
        t = torch.tensor(np.asarray(v, dtype=np.float64))
    if t.dim() == 0:
        return torch.eye(n, dtype=t.dtype) * t
    if t.dim() == 1:
        return torch.diag(t)
    return t
