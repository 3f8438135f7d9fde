import torch


def get_device():
    """The oracle always runs the reference on the host."""
    return torch.device('cpu')
