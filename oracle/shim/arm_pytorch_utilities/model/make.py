"""make_sequential_network restated: Sequential(Linear, act, Linear, act, ..., Sequential(Linear)), float64.
Layer structure verified from checkpoint keys '0','2','4.0'; activation LeakyReLU pinned empirically
(SURVEY.md §7 #3)."""
import torch


def make_sequential_network(config, h_units=(32, 32), activation_factory=torch.nn.LeakyReLU, end_block_factory=None,
                            **kwargs):
    in_dim = config.input_dim()
    out_dim = config.ny
    layers = []
    for h in h_units:
        layers.append(torch.nn.Linear(in_dim, h, bias=True))
        layers.append(activation_factory())
        in_dim = h
    layers.append(torch.nn.Sequential(torch.nn.Linear(in_dim, out_dim, bias=True)))
    return torch.nn.Sequential(*layers).double()
