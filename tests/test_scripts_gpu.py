"""GPU (-m gpu): the batched counterparts of the reference's remaining callers of the env step:
bot_arena.py, train_ppo_ppo.py, train_ppo_cycle.py - each drives the CUDA path through the reference's gym API."""
import importlib.util
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
pytestmark = pytest.mark.gpu


def _script(name):
    spec = importlib.util.spec_from_file_location("at_script_" + name, os.path.join(REPO, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_masked_reset_after_the_step_equals_the_auto_reset_inside_it():
    """bot_arena.py / the cycle evaluation read info["winner"] at the terminal tick and reset afterwards (the
    reference's own order, bot_arena.py:61-70); that must be the same battle stream as the auto-reset inside the step."""
    from alphatank_b200 import MultiAgentEnv
    dev = torch.device("cuda:0")
    N = 1024
    kw = dict(mode="bot_vs_bot", bot_types=("aggressive", "dodge"), num_envs=N, device=dev, seed=7)
    a = MultiAgentEnv(auto_reset=True, **kw)
    b = MultiAgentEnv(auto_reset=False, **kw)
    oa, _ = a.reset()
    ob, _ = b.reset()
    assert torch.equal(oa, ob)
    ended = 0
    for t in range(500):
        oa, ra, da, _, _ = a.step()
        ob, rb, db, _, _ = b.step()
        assert torch.equal(da, db) and torch.equal(ra, rb), t
        ia, ib = a.core.get_info(terminal=True), b.core.get_info()          # b: not yet reset = the reference's info
        assert torch.equal(ia["latched"], da)
        for k in ("winner", "hits", "be_hits", "change_time"):
            assert torch.equal(ia[k], ib[k]), (t, k)
        ended += int(db.sum())
        ob = b.reset(mask=db)[0].view(N, -1)
        assert torch.equal(oa, ob), t
    assert ended > 100


def test_bot_arena_counts_every_match_once():
    """scripts/bot_arena.py: a stationary bot never fires, so it never wins; every finished match has exactly one
    outcome; the aggressive bot's hits are its wins (tanks die on the first hit when TERMINATE_TIME is None)."""
    arena = _script("bot_arena")
    r = arena.run_bot_matches("aggressive", "stationary", num_envs=2048, ticks=700, seed=3)
    assert r["episodes"] > 500
    assert r["bot1_wins"] + r["bot2_wins"] + r["draws"] == r["episodes"]
    assert r["bot2_wins"] == 0 and r["draws"] == 0 and r["bot2_hits"] == 0
    assert r["bot1_hits"] == r["bot1_wins"]
    assert 1.0 < r["mean_episode_ticks"] < 700
    with pytest.raises(ValueError):                                       # BotFactory.create_bot's error (bot_factory.py:29-31)
        arena.run_bot_matches("aggressive", "nosuchbot", num_envs=32, ticks=1)


def test_rollout_buffer_keeps_the_terminal_winner_of_every_tick():
    """RolloutBuffer(keep_winner=True): the graph-captured rollout stores info["winner"] of every tick; replaying the
    recorded actions through an env without auto-reset gives the same winners at the terminal ticks."""
    from alphatank_b200 import MultiAgentEnv
    from alphatank_b200.learner import PPOConfig, PPOLearner, RolloutBuffer, RolloutRunner
    from alphatank_b200.learner.ppo import SingleAgentView
    dev = torch.device("cuda:0")
    N, T = 512, 24
    env = MultiAgentEnv(mode="bot_agent", type="train", bot_type="aggressive", num_envs=N, device=dev, seed=5)
    ref = MultiAgentEnv(mode="bot_agent", type="train", bot_type="aggressive", num_envs=N, device=dev, seed=5, auto_reset=False)
    learner = PPOLearner(1, 388, (3, 3, 2), PPOConfig(num_steps=T), dev, seed=0)
    buf = RolloutBuffer(T, N, 1, 388, dev, keep_raw=False, keep_winner=True)
    runner = RolloutRunner(SingleAgentView(env), learner, buf)
    ref.reset()
    acts = torch.zeros((N, 2, 3), dtype=torch.int32, device=dev)
    ended = wins = 0
    for it in range(4):                                   # it 0: eager, it 1: capture + replay, then replays
        runner.run(first=(it == 0))
        for t in range(T):
            acts[:, 1] = buf.actions[t, :, 0]
            _, _, done, _, info = ref.step(acts)
            assert torch.equal(done, buf.dones[t + 1]), (it, t)
            d = done.bool()
            assert torch.equal(info["winner"][d], buf.winner[t][d]), (it, t)
            ended += int(d.sum())
            wins += int((buf.winner[t][d] == 1).sum())
            ref.reset(mask=done)
    assert ended > 100 and 0 < wins < ended, (ended, wins)


def test_ppo_cycle_rotates_bots_follows_the_weakness_schedule_and_saves(tmp_path):
    """scripts/train_ppo_cycle.py: fixed rotation over two bot types, linear weakness schedule applied in place
    (at_set_weakness), rollouts as CUDA graphs per bot type, evaluation with exact winners, reference checkpoint name."""
    cyc = _script("train_ppo_cycle")
    from alphatank_b200.learner import PPOAgentBot
    N, T, U = 512, 16, 6
    cfg = cyc.CycleTrainingConfig(bot_types=["aggressive", "random"], rotation_strategy="fixed", switch_every=N * T,
                                  eval_frequency=3 * N * T, num_steps=T, num_epochs=1, total_timesteps=U * N * T,
                                  weakness_start=0.2, weakness_end=0.8)
    assert abs(cfg.get_current_weakness(U * N * T) - 0.8) < 1e-12 and abs(cfg.get_current_weakness(0) - 0.2) < 1e-12
    lines = []
    learner, hist, path = cyc.train_cycle(cfg, num_envs=N, updates=U, seed=0, eval_envs=256, eval_ticks=150,
                                          ckpt_dir=str(tmp_path), log=lines.append)
    assert [h["bot"] for h in hist] == ["random", "aggressive", "random", "aggressive", "random", "aggressive"]
    ws = [h["weakness"] for h in hist]
    assert ws == sorted(ws) and abs(ws[-1] - 0.8) < 1e-12
    assert "eval" in hist[2] and set(hist[2]["eval"]) == {"aggressive", "random"}
    assert all(0.0 <= r["win_rate"] <= 1.0 for r in hist[5]["eval"].values())
    assert os.path.basename(path) == "ppo_agent_cycle.pt"
    sd = torch.load(path, map_location="cpu", weights_only=False)
    assert list(sd) == list(PPOAgentBot(388).state_dict())
    with pytest.raises(ValueError):
        cyc.CycleTrainingConfig(rotation_strategy="sometimes")
    with pytest.raises(ValueError):
        cyc.CycleTrainingConfig(weakness_schedule="cubic")
    s = cyc.CycleTrainingConfig(weakness_schedule="sigmoid", weakness_start=0.1, weakness_end=0.9, total_timesteps=100)
    assert abs(s.get_current_weakness(50) - 0.5) < 1e-12


def test_ppo_ppo_training_script_runs_and_saves_reference_layout(tmp_path):
    """scripts/train_ppo_ppo.py (the reference's 1v1 self-play trainer, batched): both tanks policy-driven, one CUDA
    graph per rollout, checkpoints ppo_agent_<i>.pt = state_dict of a PPOAgentPPO(388)."""
    from alphatank_b200.learner import PPOAgentPPO
    script = os.path.join(REPO, "scripts", "train_ppo_ppo.py")
    out = subprocess.run([sys.executable, script, "--num-envs", "1024", "--iterations", "3", "--num-steps", "16",
                          "--num-epochs", "1", "--ckpt-dir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "iter 2" in out.stdout and "Saved model for Agent 1" in out.stdout
    for i in range(2):
        sd = torch.load(str(tmp_path / ("ppo_agent_%d.pt" % i)), map_location="cpu", weights_only=False)
        assert list(sd) == list(PPOAgentPPO(388).state_dict())


def test_sac_ppo_training_script_runs_and_saves_reference_layout(tmp_path):
    """scripts/train_sac_ppo.py (the reference's PPO-vs-SAC trainer, batched): tank 0 learns by PPO from the on-policy
    rollout, tank 1 by SAC from the replay buffer; ppo_agent.pt / sac_agent.pt in the reference's layouts."""
    from alphatank_b200.learner import PPOAgentPPO
    script = os.path.join(REPO, "scripts", "train_sac_ppo.py")
    out = subprocess.run([sys.executable, script, "--num-envs", "512", "--iterations", "2", "--ppo-num-steps", "16",
                          "--ppo-num-epochs", "1", "--sac-update-every", "3", "--sac-batch-size", "128", "--capacity",
                          "20000", "--ckpt-dir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "iter 1" in out.stdout and "Saved models" in out.stdout
    sd = torch.load(str(tmp_path / "ppo_agent.pt"), map_location="cpu", weights_only=False)
    assert list(sd) == list(PPOAgentPPO(388).state_dict())
    sd = torch.load(str(tmp_path / "sac_agent.pt"), map_location="cpu", weights_only=False)
    assert set(sd) == {"actor", "critic1", "critic2", "critic1_target", "critic2_target", "log_alpha"}
    assert sd["actor"]["fc1.weight"].shape == (256, 388)
