"""math_utils restated."""
import math
import numpy as np
import torch


def clip(a, min_val, max_val):
    """max(min(a, hi), lo) elementwise (controller.py:303)."""
    if torch.is_tensor(a):
        return torch.max(torch.min(a, max_val), min_val)
    return np.clip(a, min_val, max_val)


def get_bounds(u_min=None, u_max=None):
    """Mirror a missing bound (controller.py:199)."""
    if u_max is not None and u_min is None:
        u_min = -u_max
    if u_min is not None and u_max is None:
        u_max = -u_min
    return u_min, u_max


def angular_diff_batch(a, b):
    """a - b with a single wrap into [-pi, pi] (block_push.py:957)."""
    d = a - b
    if torch.is_tensor(d):
        d = d.clone()
    else:
        d = np.array(d, copy=True)
    d[d > math.pi] -= 2 * math.pi
    d[d < -math.pi] += 2 * math.pi
    return d


def angular_diff(a, b):
    d = a - b
    if d > math.pi:
        d -= 2 * math.pi
    elif d < -math.pi:
        d += 2 * math.pi
    return d


def angle_normalize(a):
    """((a + pi) mod 2pi) - pi (push_main.py:192, pendulum.py:246)."""
    return ((a + math.pi) % (2 * math.pi)) - math.pi


def rotate_wrt_origin(xy, theta):
    return (xy[0] * math.cos(theta) - xy[1] * math.sin(theta),
            xy[0] * math.sin(theta) + xy[1] * math.cos(theta))


def batch_rotate_wrt_origin(xy, theta):
    c, s = torch.cos(theta), torch.sin(theta)
    x = xy[:, 0] * c - xy[:, 1] * s
    y = xy[:, 0] * s + xy[:, 1] * c
    return torch.stack((x, y), dim=1)
