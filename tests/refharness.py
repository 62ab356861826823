"""TEST-ONLY harness that runs the UNMODIFIED reference (/root/reference) headless
and deterministically, to pin the oracle and to generate tests/golden/*.npz.

What it does (SURVEY.md 8c "determinism recipe", App. D):
  * puts tests/refstubs (pygame, gym) and the reference root on sys.path;
  * replaces pygame.time.get_ticks by a virtual clock: during the k-th call of
    env.step() (k = 1, 2, ... per env instance, never reset) it reads k*1000//60;
  * serves every ``random.*`` / ``np.random.*`` draw of the env modules from the
    ATRNG counter stream (oracle/atrng.py) so that reference, oracle and CUDA
    kernels see the same spawn cells, angles, zones and bot randomness;
  * skips the render-only GIF decoding in Tank.__init__ / Explosion.__init__
    (env/sprite.py:248, env/util.py:112-119; quirk F21) - no simulation effect;
  * silences the reference's prints.

Nothing here is imported by the product (alphatank_b200/).
"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_ROOT = os.environ.get("ALPHATANK_REFERENCE", "/root/reference")
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle.atrng import Stream  # noqa: E402


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "env", "gym_env.py"))


class _StreamRandom:
    """Stands in for the ``random`` module inside the reference's env modules."""

    def __init__(self):
        self.stream = None

    def randint(self, a, b):
        return self.stream.randint(a, b)

    def randrange(self, start, stop=None, step=1):
        if stop is None:
            start, stop = 0, start
        n = (stop - start + step - 1) // step
        return start + step * self.stream.below(n)

    def random(self):
        return self.stream.unit53()

    def uniform(self, a, b):
        return self.stream.uniform(a, b)

    def sample(self, population, k):
        pool = list(population)
        out = []
        for _ in range(k):
            out.append(pool.pop(self.stream.below(len(pool))))
        return out

    def shuffle(self, x):
        for i in reversed(range(1, len(x))):
            j = self.stream.below(i + 1)
            x[i], x[j] = x[j], x[i]

    def choice(self, seq):
        return seq[self.stream.below(len(seq))]


class _NpRandom:
    def __init__(self, sr):
        self._sr = sr

    def choice(self, a, *args, **kw):
        assert not args and not kw
        a = list(a) if not isinstance(a, int) else list(range(a))
        return a[self._sr.stream.below(len(a))]

    def random(self):
        return self._sr.stream.unit53()


class _NpProxy:
    """``np`` inside env.gaming_env(_multi): real numpy except ``np.random``."""

    def __init__(self, sr):
        self.random = _NpRandom(sr)

    def __getattr__(self, name):
        return getattr(np, name)


_STATE = {"loaded": False}
RANDOM = _StreamRandom()


def load():
    """Import the reference's env package once, with stubs and patches."""
    if _STATE["loaded"]:
        return _STATE
    assert available(), "reference not present at %s" % REF_ROOT
    stubs = os.path.join(HERE, "refstubs")
    for p in (REF_ROOT, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):
        mods = {}
        for name in ("pygame", "configs.config_basic", "configs.config_teams", "env.util", "env.sprite",
                     "env.bfs", "env.maze", "env.gaming_env", "env.gaming_env_multi", "env.gym_env",
                     "env.gym_env_multi", "env.controller", "env.bots.simple_bots", "env.bots.strategy_bot",
                     "env.bots.bot_factory"):
            mods[name] = importlib.import_module(name)
    sprite, util = mods["env.sprite"], mods["env.util"]
    # render-only asset loading (F21)
    sprite.Tank.load_and_colorize_gif = lambda self, path, color, size: []

    def _explosion_init(self, x, y, env):
        self.x, self.y, self.env, self.alive = x, y, env, True
        self.frame_index = self.tick = 0
        self.frames, self.total_frames, self.frame_rate, self.played_once = [], 0, 1, False
    util.Explosion.__init__ = _explosion_init
    # every draw of the env modules comes from the ATRNG stream
    sprite.random = RANDOM
    mods["env.maze"].random = RANDOM
    mods["env.bots.simple_bots"].random = RANDOM
    for m in ("env.gaming_env", "env.gaming_env_multi"):
        mods[m].random = RANDOM
        mods[m].np = _NpProxy(RANDOM)
    _STATE.update(loaded=True, mods=mods, orig_aim=sprite.Tank._aiming_reward)
    return _STATE


def set_config(terminate_time=None, use_octagon=True, buff_on=True, debuff_on=True, elide_aim=False):
    """Patch the star-imported constants in every module that copied them."""
    m = load()["mods"]
    for name in ("env.sprite", "env.gym_env", "env.gym_env_multi"):
        m[name].TERMINATE_TIME = terminate_time
    for name in ("env.gaming_env", "env.gaming_env_multi"):
        m[name].USE_OCTAGON = use_octagon
    for name in ("env.sprite", "env.gaming_env", "env.gaming_env_multi"):
        m[name].BUFF_ON = buff_on
        m[name].DEBUFF_ON = debuff_on
    # F18: the zero-weight aiming ray-march (87 % of reference CPU time) can be
    # elided for long traces; test_reference_pin proves the streams are identical.
    sprite = m["env.sprite"]
    sprite.Tank._aiming_reward = (lambda self: None) if elide_aim else _STATE["orig_aim"]


class RefEnv:
    """One reference env + its ATRNG stream + its virtual clock.

    kind: "agent"      MultiAgentEnv(mode="agent")                       (BASELINE config 1)
          "bot_agent"  MultiAgentEnv(mode="bot_agent", bot_type, weakness)      (config 2)
          "arena"      MultiAgentEnv(mode="bot_vs_bot") driven like bot_arena.py:51-58 (config 4)
          "team"       MultiAgentTeamEnv(game_configs)                            (configs 3, 5)
    """

    def __init__(self, kind, seed, env_id, bot_type="smart", weakness=1.0, game_configs=None,
                 arena_bots=("defensive", "dodge")):
        st = load()
        self.m = st["mods"]
        self.kind = kind
        self.stream = Stream(seed, env_id)
        self.tick = 0
        self.arena_bots = arena_bots
        self.bots = None
        with self._ctx():
            if kind == "team":
                self.env = self.m["env.gym_env_multi"].MultiAgentTeamEnv(game_configs)
            elif kind == "arena":
                self.env = self.m["env.gym_env"].MultiAgentEnv(mode="bot_vs_bot")
            else:
                self.env = self.m["env.gym_env"].MultiAgentEnv(mode=kind, type="train", bot_type=bot_type,
                                                               weakness=weakness)
        self.game = self.env.game_env
        if kind == "arena":
            self._make_arena_bots()

    @contextlib.contextmanager
    def _ctx(self):
        RANDOM.stream = self.stream
        self.m["pygame"].time.now_ms = self.tick * 1000 // 60
        with contextlib.redirect_stdout(io.StringIO()):
            yield

    def _make_arena_bots(self):
        f = self.m["env.bots.bot_factory"].BotFactory
        self.bots = [f.create_bot(t, self.game.tanks[i]) for i, t in enumerate(self.arena_bots)]

    def reset(self):
        with self._ctx():
            obs, info = self.env.reset()
        if self.kind == "arena":
            self._make_arena_bots()
        return np.asarray(obs, dtype=np.float32).reshape(-1), info

    def step(self, actions=None):
        self.tick += 1
        with self._ctx():
            if self.kind == "arena":
                actions = [b.get_action() for b in self.bots]   # bot_arena.py:51-52 (pre-tick state)
                actions = [list(a) for a in actions]
            if actions is None:
                actions = []     # team config with no agent tanks (bot_team_configs)
            obs, rew, done, trunc, info = self.env.step([list(map(int, a)) for a in actions])
        return obs, rew, bool(done), info, actions

    # ------------------------------------------------------------ state dump
    def tank_state(self):
        """float64[n, 12]: x y angle speed alive hittingWall reward num_hit num_be_hit in_buff in_debuff max_bullets"""
        rows = []
        for t in self.game.tanks:
            rows.append([t.x, t.y, t.angle, t.speed, float(bool(t.alive)), float(bool(t.hittingWall)), t.reward,
                         t.num_hit, t.num_be_hit, float(bool(t.in_buff_zone)), float(bool(t.in_debuff_zone)),
                         t.max_bullets])
        return np.asarray(rows, dtype=np.float64)

    def bullet_state(self, cap):
        """(count, float64[cap, 9]): owner x y dx dy distance bounces pending cleared - in env.bullets order"""
        out = np.zeros((cap, 9), dtype=np.float64)
        bl = self.game.bullets
        assert len(bl) <= cap
        for i, b in enumerate(bl):
            out[i] = [self.game.tanks.index(b.owner), b.x, b.y, b.dx, b.dy, b.distance_traveled, b.bounces,
                      b.pending_penalty, float(b.penalty_cleared)]
        return len(bl), out

    def zones(self):
        z = list(self.game.buff_zones) + list(self.game.debuff_zones)
        return np.asarray(z, dtype=np.float64).reshape(-1)

    def maze(self):
        return np.asarray(self.game.maze, dtype=np.int32)


def record_trace(kind, seed, env_id, ticks, action_seed=None, **kw):
    """Run one reference env for ``ticks`` steps with ATRNG synthetic actions
    (stream 1+a, ctr = tick) and the trainers' reset-on-done policy
    (train_ppo_bot.py:121-129 minus the negative-reward rule).  Returns a dict
    of arrays; see tests/golden/gen_golden.py for the schema."""
    from oracle.atrng import synthetic_actions
    r = RefEnv(kind, seed, env_id, **kw)
    n = len(r.game.tanks)
    if kind == "team":
        n_act = r.env.num_agents
    elif kind == "arena":
        n_act = 0
    else:
        n_act = 2
    cap = 6 * n
    obs0, _ = r.reset()          # the trainers call reset() once after the constructor's own reset
    T = ticks
    rec = {
        "reset_obs": [obs0], "reset_tick": [0], "reset_tanks": [r.tank_state()], "reset_zones": [r.zones()],
        "obs": np.zeros((T, obs0.size), np.float32), "rewards": None, "done": np.zeros(T, np.uint8),
        "tanks": np.zeros((T, n, 12), np.float64), "nbul": np.zeros(T, np.int32),
        "bullets": np.zeros((T, cap, 9), np.float64), "actions": np.zeros((T, max(n_act, 2), 3), np.int32),
        "rng_ctr": np.zeros(T, np.int64), "maze": r.maze(),
    }
    aseed = seed if action_seed is None else action_seed
    for t in range(T):
        acts = None
        if n_act:
            acts = synthetic_actions(aseed, [env_id], t, n_act)[0]
        obs, rew, done, info, used = r.step(acts)
        if rec["rewards"] is None:
            rec["rewards"] = np.zeros((T, len(rew)), np.float32)
        rec["obs"][t] = obs
        rec["rewards"][t] = rew
        rec["done"][t] = done
        rec["tanks"][t] = r.tank_state()
        rec["nbul"][t], rec["bullets"][t] = r.bullet_state(cap)
        ua = np.asarray(used if used is not None else [], dtype=np.int32).reshape(-1, 3)
        rec["actions"][t, :ua.shape[0]] = ua
        rec["rng_ctr"][t] = r.stream.ctr
        if done:
            o, _ = r.reset()
            rec["reset_obs"].append(o)
            rec["reset_tick"].append(t + 1)
            rec["reset_tanks"].append(r.tank_state())
            rec["reset_zones"].append(r.zones())
    for k in ("reset_obs", "reset_tick", "reset_tanks", "reset_zones"):
        rec[k] = np.asarray(rec[k])
    return rec
