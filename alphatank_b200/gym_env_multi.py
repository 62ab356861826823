"""Batched drop-in for the reference's env/gym_env_multi.py (MultiAgentTeamEnv over GamingTeamENV)."""
from . import _capi
from .batched_env import BatchedTankEnv, LazyInfo, TankView
from .configs import config_basic
from .spaces import Box, MultiDiscrete


class _TeamGameView:
    def __init__(self, owner):
        self._o = owner
        self.game_configs = owner.game_configs                     # inference_multi.py:101
        self.tanks = [TankView(owner, i) for i in range(owner.num_tanks)]

    def get_observation_order(self):
        return self._o.get_observation_order()


class MultiAgentTeamEnv:
    """N batched copies of ``MultiAgentTeamEnv(game_configs)`` (env/gym_env_multi.py:13).

    ``game_configs`` is a reference-style dict (configs/config_teams.py): name -> {"team", "color", "mode",
    ["bot_type"]}.  Tanks act in the reference's order: agent tanks in config order, then bot tanks in config
    order (gaming_env_multi.py:82-85).  reset() -> (obs[N, A*D], info); step(actions[N, A, 3]) ->
    (obs[N, A*D], rewards[N, A] cumulative, done[N], False, info{"hits", "be_hits"}).  ``mode: "human"`` tanks
    are rejected (keyboard)."""

    def __init__(self, game_configs, *, num_envs=1, device="cuda:0", seed=0, env_id_base=0, auto_reset=True, map=None,
                 terminate_time="config", buff_on=None, debuff_on=None, reward_mode="cumulative"):
        if reward_mode not in ("cumulative", "delta"):
            raise ValueError("reward_mode must be 'cumulative' (the reference's per-episode running sum) or 'delta'")
        self.reward_mode = reward_mode
        self.game_configs = game_configs
        teams, ctrl, bots, ids = [], [], [], {}
        self._agent_names = []
        for name, tc in game_configs.items():
            if tc["mode"] == "human":
                raise ValueError("tank %s: mode 'human' needs a keyboard and is not supported by the batched env" % name)
            if tc["mode"] not in ("agent", "bot"):
                raise ValueError("tank %s: unknown mode %r" % (name, tc["mode"]))
            teams.append(ids.setdefault(tc["team"], len(ids)))
            ctrl.append(tc["mode"])
            bots.append(tc.get("bot_type"))
            if tc["mode"] == "agent":
                self._agent_names.append(name)
        m = map or ("octagon" if config_basic.USE_OCTAGON else "battlefield")
        tt = config_basic.TERMINATE_TIME if terminate_time == "config" else terminate_time
        cfg = _capi.make_config("team", teams, ctrl, bots, map=m, terminate_time=tt,
                                buff_on=config_basic.BUFF_ON if buff_on is None else buff_on,
                                debuff_on=config_basic.DEBUFF_ON if debuff_on is None else debuff_on,
                                auto_reset=auto_reset)
        self.core = BatchedTankEnv(cfg, num_envs, device, seed, env_id_base)
        if self.reward_mode == "delta":
            self.core.set_reward_mode("delta")          # per-step increments, computed by the step kernel
        self.training_step = 0
        self.num_envs = self.core.num_envs
        self.num_tanks = len(teams)
        self.num_agents = len(self._agent_names)
        self.num_walls = self.core.num_walls
        self.max_bullets_per_tank = 6
        self.obs_dim_per_agent = self.core.D
        self.observation_space = Box(low=-1, high=max(config_basic.WIDTH, config_basic.HEIGHT),
                                     shape=(self.core.D * self.num_agents,))
        self.action_space = MultiDiscrete([3, 3, 2] * self.num_agents)
        self.game_env = _TeamGameView(self)

    def get_observation_order(self):
        return list(self._agent_names)                              # gaming_env_multi.py get_observation_order

    def reset(self, mask=None):
        obs = self.core.reset(mask)
        info = {"Tank:%d-%s" % (i, tc["team"]): 0 for i, tc in enumerate(self.game_configs.values())}
        return obs, info

    def step(self, actions=None):
        self.training_step += 1
        obs, rewards, done = self.core.step(actions)
        return obs, rewards, done, False, LazyInfo(self.core, ("hits", "be_hits"))

    def step_host(self, h_actions=None):
        self.training_step += 1
        return self.core.step_host(h_actions)

    def render(self, mode="human"):
        return None

    def close(self):
        self.core.close()
