import numpy as np


def discrete_array_to_value_ranges(arr):
    """Yield (value, start, end_inclusive) for each run of equal values (env/env.py:42)."""
    a = np.asarray(arr).reshape(-1)
    out = []
    if a.size == 0:
        return out
    start = 0
    for i in range(1, a.size + 1):
        if i == a.size or a[i] != a[start]:
            out.append((a[start], start, i - 1))
            start = i
    return out
