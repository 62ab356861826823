"""alphatank_b200 - B200-native batched simulator for alphaTank's environment step.

    from alphatank_b200 import MultiAgentEnv, MultiAgentTeamEnv
    from alphatank_b200.configs.config_teams import team_vs_bot_configs
    env = MultiAgentTeamEnv(team_vs_bot_configs, num_envs=65536, device="cuda:0", seed=0)

The classes keep the reference's gym_env / gym_env_multi reset()/step() API; the step itself is one fused
sm_100a CUDA kernel behind the C ABI in include/alphatank_b200.h."""
from .gym_env import MultiAgentEnv  # noqa: F401
from .gym_env_multi import MultiAgentTeamEnv  # noqa: F401

__all__ = ["MultiAgentEnv", "MultiAgentTeamEnv"]
__version__ = "0.1.0"
