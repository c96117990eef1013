"""dcbf-mpc-b200: batched D-CBF ALIP/LIP MPC solver for NVIDIA B200 (sm_100a).

Drop-in for the reference planner's hot path (MPC_LIP_sig_step.py, MPC_LIP_modi.py, MPC_DD_sig_step.py,
ALIP_plan/planner.py of shaygong322/Mujoco-LIP-MPC-Simulation): same call surface, CUDA behind a thin C ABI
(include/dcbf_mpc.h).  No CPU fallback.
"""
from . import _lib, scenarios  # noqa: F401

__all__ = ["_lib", "scenarios", "DcbfSolver", "HostPipeline", "default_params"]


def __getattr__(name):
    if name in ("DcbfSolver", "HostPipeline", "default_params", "SolveResult"):
        from . import batch
        return getattr(batch, name)
    raise AttributeError(name)
