"""Learner side of the batched env: PyTorch PPO / SAC pieces that mirror the reference's
models/ppo_utils.py, models/sac_utils.py and train_multi_ppo.py on top of the CUDA env step.  The env step itself
(alphatank_b200/csrc) is the product's hot path; this package is plain PyTorch by design (BASELINE.json north_star:
"PyTorch for the PPO/SAC side", NCCL only for the gradient all-reduce)."""
from .ppo_utils import PPOAgentBot, PPOAgentPPO, RunningMeanStd, tick_normalize  # noqa: F401
from .ppo import PPOConfig, PPOLearner, RolloutBuffer, RolloutRunner, collect_rollout, load_team_checkpoint, save_team_checkpoint  # noqa: F401
from .sac_utils import ContinuousToDiscrete, DeviceReplayBuffer, PolicyNetwork, QNetwork, SACAgent  # noqa: F401
