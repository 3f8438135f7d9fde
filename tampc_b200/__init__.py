"""tampc_b200 — B200-native MPPI rollout hot path, a drop-in for the `pytorch_mppi` `MPPI.command()`
call made by tampc's OnlineMPPI / MPPI_MPC controllers (reference: tampc/controller/controller.py:316-435,
tampc/controller/online_controller.py:60-94,497).  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
