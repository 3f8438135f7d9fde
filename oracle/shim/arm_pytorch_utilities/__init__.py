"""Minimal restatement of the arm_pytorch_utilities names tampc's MPPI path uses (see ../README.md)."""
