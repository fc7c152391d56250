"""quadsim-b200: the quadrotor racing environments of tudelft/optimal_quad_control_RL as sm_100a CUDA kernels
behind the reference's own ``Quadcopter3DGates(VecEnv)`` interface."""
from .tracks import rectangle_track, training_disturbance_ranges, zigzag_track  # noqa: F401
from .sharding import ObsAllGather, ObsPeerGather, shard_range  # noqa: F401
from ._lib import QuadsimError  # noqa: F401


def __getattr__(name):  # envs needs torch + the CUDA library: import lazily so CPU-only tooling can import the package
    if name in ("Quadcopter3DGates", "Quadcopter3DGatesINDI", "load_residual_weights"):
        from . import envs
        return getattr(envs, name)
    if name in ("PPO", "VecMonitor", "ActorCriticPolicy", "a8_bootstrap_"):
        from . import ppo
        return getattr(ppo, name)
    if name == "MlpPolicy":
        from . import policy
        return policy.MlpPolicy
    if name in ("TrajectoryLog", "log_policy_run"):
        from . import trajectory
        return getattr(trajectory, name)
    if name in ("export_controller", "build_controller", "CController", "track_spec"):
        from . import codegen
        return getattr(codegen, name)
    raise AttributeError(name)
