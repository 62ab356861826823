"""Batched drop-in for the reference's env/gym_env.py (MultiAgentEnv over GamingENV).

Same constructor keywords, attributes and reset()/step() contract as env/gym_env.py:10-103, with a
leading batch dimension: N independent 1v1 battles stepped by one CUDA kernel launch.  Differences that
follow from batching are listed in INTEGRATION.md (tensors instead of numpy arrays, per-env auto-reset)."""
import torch

from . import _capi
from .batched_env import BatchedTankEnv, LazyInfo, TankView
from .configs import config_basic
from .spaces import Box, MultiDiscrete

_1V1_MODES = {"agent": "agent", "bot_agent": "bot_agent", "bot_vs_bot": "arena"}


class _GameView:
    """``env.game_env`` stand-in (gym_env.py:14): exposes what the reference's callers touch."""

    def __init__(self, owner):
        self._o = owner
        self.tanks = [TankView(owner, i) for i in range(owner.num_tanks)]

    bot_type = property(lambda s: s._o.bot_type)
    mode = property(lambda s: s._o.mode)
    weakness = property(lambda s: s._o.weakness)


class MultiAgentEnv:
    """N batched copies of ``MultiAgentEnv(mode, type, bot_type, weakness)`` (env/gym_env.py:11).

    mode: "agent" (both tanks policy-driven, the "AI only" branch gaming_env.py:309-374), "bot_agent"
    (tank 0 scripted bot, tank 1 agent, gaming_env.py:139-199) or "bot_vs_bot" (both scripted, decisions on
    the pre-tick state as bot_arena.py:51-58; ``bot_types=(b0, b1)``).  The keyboard modes "bot" /
    "human_play" are rejected.  reset() -> (obs[N, n, D], info); step(actions[N, 2, 3]) ->
    (obs[N, n*D], rewards[N, n] cumulative, done[N], False, info) - shapes as gym_env.py:69-71, 100-103 (F2).
    """

    def __init__(self, mode="agent", type="train", bot_type="smart", weakness=1.0, *, num_envs=1, device="cuda:0",
                 seed=0, env_id_base=0, auto_reset=True, map=None, terminate_time="config", buff_on=None,
                 debuff_on=None, bot_types=None, reward_mode="cumulative"):
        if reward_mode not in ("cumulative", "delta"):
            raise ValueError("reward_mode must be 'cumulative' (the reference's per-episode running sum) or 'delta'")
        self.reward_mode = reward_mode
        if mode not in _1V1_MODES:
            raise ValueError("mode %r is not supported by the batched env (keyboard modes 'bot', 'human_play' and "
                             "'human' need a display); valid: %s" % (mode, sorted(_1V1_MODES)))
        self.mode, self.type, self.bot_type, self.weakness = mode, type, bot_type, float(weakness)
        self._kw = dict(num_envs=num_envs, device=device, seed=seed, env_id_base=env_id_base, auto_reset=auto_reset,
                        map=map, terminate_time=terminate_time, buff_on=buff_on, debuff_on=debuff_on,
                        bot_types=bot_types)
        self.training_step = 0
        self.max_bullets_per_tank = 6
        self._build()

    def _build(self):
        kw = self._kw
        m = kw["map"] or ("octagon" if config_basic.USE_OCTAGON else "battlefield")
        tt = config_basic.TERMINATE_TIME if kw["terminate_time"] == "config" else kw["terminate_time"]
        buff = config_basic.BUFF_ON if kw["buff_on"] is None else kw["buff_on"]
        debuff = config_basic.DEBUFF_ON if kw["debuff_on"] is None else kw["debuff_on"]
        kind = _1V1_MODES[self.mode]
        if kind == "agent":
            ctrl, bots = ["agent", "agent"], [None, None]
        elif kind == "bot_agent":
            ctrl, bots = ["bot", "agent"], [self.bot_type, None]
        else:
            bt = kw["bot_types"] or (self.bot_type, self.bot_type)
            ctrl, bots = ["bot", "bot"], list(bt)
        cfg = _capi.make_config(kind, [0, 1], ctrl, bots, map=m, terminate_time=tt, buff_on=buff, debuff_on=debuff,
                                auto_reset=kw["auto_reset"], weakness=self.weakness)
        old = getattr(self, "core", None)
        if old is not None:
            old.close()
        self.core = BatchedTankEnv(cfg, kw["num_envs"], kw["device"], kw["seed"], kw["env_id_base"])
        if self.reward_mode == "delta":
            self.core.set_reward_mode("delta")          # per-step increments, computed by the step kernel
        # gaming_env.py:176-177: outside training ('inference') the bot_agent branch reads the agent's action from row 0
        self._agent_row0 = kind == "bot_agent" and self.type != "train"
        self.num_envs = self.core.num_envs
        self.num_tanks = 2
        self.num_agents = 2 if kind == "agent" else (1 if kind == "bot_agent" else 0)
        self.num_walls = self.core.num_walls
        self.obs_dim_per_tank = self.core.D
        obs_dim = self.core.D * self.num_tanks                       # gym_env.py:53-61
        self.observation_space = Box(low=-1, high=max(config_basic.WIDTH, config_basic.HEIGHT), shape=(obs_dim,))
        self.action_space = MultiDiscrete([3, 3, 2] * self.num_tanks)  # gym_env.py:23
        self.game_env = _GameView(self)

    def set_bot_type(self, bot_type):
        """gym_env.py:25-29: change the scripted opponent and reset."""
        self.bot_type = bot_type
        self._build()
        return self.reset()

    def set_weakness(self, weakness):
        """Curriculum hook (train_ppo_cycle.py:82-92, 296): probability that the scripted opponent acts on a tick,
        changed in place (no reset, no rebuild)."""
        from . import _capi
        _capi.check(self.core.L.at_set_weakness(self.core.h, float(weakness)), "at_set_weakness")
        self.weakness = float(weakness)

    def reset(self, mask=None):
        obs = self.core.reset(mask)
        info = {"Tank:%d-%s" % (i, t): 0 for i, t in enumerate(("TeamA", "TeamB"))}   # gym_env.py:70
        return obs.view(self.num_envs, self.num_tanks, self.core.D), info

    def _agent_rows(self, actions):
        """bot_agent with type != 'train' (inference.py): callers pass the agent's action as row 0 - [N, 1, 3] or
        [N, 2, 3] with row 1 unused.  The engine's agent tank is tank 1, so the row is handed to it as row 1; row 0
        keeps the caller's values, whose repeats change_time[0] counts (gym_env.py:77-85)."""
        import torch
        a = torch.as_tensor(actions).reshape(self.num_envs, -1, 3)
        return torch.stack((a[:, 0], a[:, 0]), dim=1).contiguous()

    def step(self, actions=None):
        self.training_step += 1
        if self._agent_row0:
            actions = self._agent_rows(actions)
        obs, rewards, done = self.core.step(actions)
        return obs, rewards, done, False, LazyInfo(self.core, ("change_time", "winner"))

    def step_host(self, h_actions=None):
        self.training_step += 1
        if self._agent_row0:
            h_actions = self._agent_rows(h_actions).to("cpu", dtype=__import__("torch").int32)
        return self.core.step_host(h_actions)

    def render(self, mode="human"):
        return None   # no pixels in the batched simulator; kept because every trainer calls it once

    def close(self):
        self.core.close()
