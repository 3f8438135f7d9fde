from pytorch_mppi.mppi import MPPI  # noqa: F401
