import random
import numpy as np
import torch


def seed(s=None):
    if s is None:
        s = random.randint(0, 2 ** 31 - 1)
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)
    return s
