"""The drop-in runners (shared + separated) drive a host vec-env exactly like the reference's MPERunner does."""
import pathlib

import numpy as np
import pytest
import torch

from oracle import mappo_oracle as O
from argsutil import make_args, make_spaces

pytestmark = pytest.mark.gpu


class FakeVecEnv:
    """Minimal stand-in for envs.env_wrappers.SubprocVecEnv over an MPE-shaped task (done at t == T)."""

    def __init__(self, cfg, seed=0, separated=False):
        self.cfg, self.rng, self.t = cfg, np.random.RandomState(seed), 0
        obs_s, share_s, act_s = make_spaces(cfg)
        M = cfg.num_agents
        self.observation_space, self.share_observation_space, self.action_space = [obs_s] * M, [share_s] * M, [act_s] * M
        self.last_actions = None

    def _obs(self):
        return self.rng.randn(self.cfg.n_rollout_threads, self.cfg.num_agents, self.cfg.obs_dim).astype(np.float32)

    def reset(self):
        self.t = 0
        return self._obs()

    def step(self, actions_env):
        a = np.asarray(actions_env)
        assert a.shape[:2] == (self.cfg.n_rollout_threads, self.cfg.num_agents)
        assert np.allclose(a.sum(-1), len(self.cfg.act_dims))            # one-hot per head
        self.last_actions = a
        self.t += 1
        N, M = self.cfg.n_rollout_threads, self.cfg.num_agents
        rew = np.repeat(self.rng.randn(N, 1, 1).astype(np.float32), M, 1)
        done = np.full((N, M), self.t % self.cfg.episode_length == 0)
        infos = [[{"individual_reward": float(rew[n, m, 0])} for m in range(M)] for n in range(N)]
        return self._obs(), rew, done, infos

    def render(self, mode="human"):
        self.render_calls = getattr(self, "render_calls", 0) + 1
        if mode == "rgb_array":                                           # [thread][viewer] frames like the MPE wrappers
            return [[np.full((4, 5, 3), self.t, dtype=np.uint8)] for _ in range(self.cfg.n_rollout_threads)]
        return None

    def close(self):
        pass


def _config(cfg, tmp_path, envs, **over):
    args = make_args(cfg)
    args.num_env_steps = cfg.episode_length * cfg.n_rollout_threads * 3
    args.log_interval = 1
    for k, v in over.items():
        setattr(args, k, v)
    return dict(all_args=args, envs=envs, eval_envs=None, num_agents=cfg.num_agents, device=torch.device("cuda:0"),
                run_dir=pathlib.Path(tmp_path))


@pytest.mark.parametrize("recurrent", [False, True])
def test_shared_mpe_runner_runs_and_checkpoints(tmp_path, recurrent):
    from onpolicy.runner.shared.mpe_runner import MPERunner
    cfg = O.PathConfig(episode_length=10, n_rollout_threads=4, num_agents=3, ppo_epoch=2, use_recurrent_policy=recurrent,
                       data_chunk_length=5, use_ReLU=False)
    envs = FakeVecEnv(cfg)
    runner = MPERunner(_config(cfg, tmp_path, envs))
    w0 = runner.policy.actor.flat.clone()
    runner.run()
    assert not torch.equal(w0, runner.policy.actor.flat)                  # it trained
    assert torch.isfinite(runner.policy.actor.flat).all() and torch.isfinite(runner.policy.critic.flat).all()
    assert runner.buffer.masks[0].sum().item() == 0                        # done at t == T carried into slot 0
    # checkpoints carry the reference's state_dict keys and reload into a fresh runner
    sd = torch.load(str(tmp_path / "models" / "actor.pt"))
    assert "base.mlp.fc1.0.weight" in sd and "act.action_out.linear.weight" in sd
    r2 = MPERunner(_config(cfg, tmp_path, FakeVecEnv(cfg), model_dir=str(tmp_path / "models")))
    assert torch.equal(r2.policy.actor.flat, runner.policy.actor.flat)
    runner.writter.export_scalars_to_json(str(tmp_path / "summary.json"))
    runner.writter.close()


def test_separated_mpe_runner_runs(tmp_path):
    from onpolicy.runner.separated.mpe_runner import MPERunner
    cfg = O.PathConfig(episode_length=8, n_rollout_threads=4, num_agents=2, obs_dim=6, share_obs_dim=12, act_dims=(3,),
                       ppo_epoch=2)
    envs = FakeVecEnv(cfg, separated=True)
    cfg_d = _config(cfg, tmp_path, envs, share_policy=False)
    runner = MPERunner(cfg_d)
    w0 = [p.actor.flat.clone() for p in runner.policy]
    runner.run()
    for a, p in zip(w0, runner.policy):
        assert not torch.equal(a, p.actor.flat)
        assert torch.isfinite(p.actor.flat).all()
    assert (tmp_path / "models" / "actor_agent1.pt").exists() and (tmp_path / "models" / "vnrom_agent0.pt").exists()


def test_shared_mpe_runner_drives_the_device_environment(tmp_path):
    """`DeviceSpreadVecEnv` (device-side simple_spread behind the reference's vec-env interface) is a drop-in for what
    make_train_env() returns: the unchanged runner loop collects, steps it with one-hot actions, inserts and trains."""
    from onpolicy.runner.shared.mpe_runner import MPERunner
    from mappo_b200.mpe_env import DeviceSpreadVecEnv
    cfg = O.PathConfig(episode_length=10, n_rollout_threads=8, num_agents=3, obs_dim=18, share_obs_dim=54, act_dims=(5,),
                       ppo_epoch=2, use_ReLU=False)
    envs = DeviceSpreadVecEnv(8, 3, 3, cfg.episode_length, device="cuda", seed=2)
    runner = MPERunner(_config(cfg, tmp_path, envs))
    w0 = runner.policy.actor.flat.clone()
    runner.run()
    assert not torch.equal(w0, runner.policy.actor.flat)
    assert torch.isfinite(runner.policy.actor.flat).all() and torch.isfinite(runner.policy.critic.flat).all()
    assert runner.buffer.masks[0].sum().item() == 0                        # world_length == episode_length: done at t == T
    r = runner.buffer.rewards
    assert torch.isfinite(r).all() and (r < 0).all()                       # distance + self-collision penalties
    assert torch.equal(r[:, :, 0], r[:, :, 1]) and torch.equal(r[:, :, 0], r[:, :, 2])      # shared reward
    obs = runner.buffer.obs[1:]                                            # share_obs = the thread's obs concatenated
    want = obs.reshape(obs.shape[0], obs.shape[1], 1, -1).expand(-1, -1, 3, -1)
    assert torch.equal(runner.buffer.share_obs[1:], want)
    runner.writter.close()


# --------------------------------------------------------------------------------------------
# SMAC runner: bookkeeping of insert against a NumPy restatement of reference smac_runner.py:133-151
# --------------------------------------------------------------------------------------------
class FakeSmacEnv:
    """SMAC-shaped host vec-env: per-agent deaths, env-level episode ends, available-action masks, bad_transition infos."""

    def __init__(self, cfg, n, seed=0):
        self.cfg, self.n, self.rng = cfg, n, np.random.RandomState(seed)
        obs_s, share_s, act_s = make_spaces(cfg)
        M = cfg.num_agents
        self.observation_space, self.share_observation_space, self.action_space = [obs_s] * M, [share_s] * M, [act_s] * M
        self.dead = np.zeros((n, M), bool)
        self.battles = np.zeros(n)
        self.log = []

    def _arrays(self):
        n, M, c = self.n, self.cfg.num_agents, self.cfg
        avail = (self.rng.rand(n, M, c.act_dims[0]) < 0.7).astype(np.float32)
        avail[..., 0] = 1.0
        return (self.rng.randn(n, M, c.obs_dim).astype(np.float32), self.rng.randn(n, M, c.share_obs_dim).astype(np.float32), avail)

    def reset(self):
        self.dead[:] = False
        return self._arrays()

    def step(self, actions):
        a = np.asarray(actions)
        n, M = self.n, self.cfg.num_agents
        assert a.shape == (n, M, 1) and (a >= 0).all() and (a < self.cfg.act_dims[0]).all()
        self.dead |= self.rng.rand(n, M) < 0.15
        end = self.rng.rand(n) < 0.2
        dones = self.dead | end[:, None]
        bad = self.rng.rand(n, M) < 0.1
        infos = [[{"bad_transition": bool(bad[i, m]), "battles_won": float(self.battles[i]), "battles_game": float(self.battles[i] + 1),
                   "won": bool(end[i])} for m in range(M)] for i in range(n)]
        self.battles += end
        rew = np.repeat(self.rng.randn(n, 1, 1).astype(np.float32), M, 1)
        obs, share, avail = self._arrays()
        self.log.append((dones.copy(), bad.copy()))
        self.dead[end] = False
        return obs, share, rew, dones, infos, avail

    def close(self):
        pass


def test_smac_runner_bookkeeping_and_training(tmp_path):
    from onpolicy.runner.shared.smac_runner import SMACRunner
    cfg = O.PathConfig(episode_length=12, n_rollout_threads=5, num_agents=3, obs_dim=30, share_obs_dim=48, act_dims=(9,), ppo_epoch=2,
                       use_recurrent_policy=True, data_chunk_length=4, use_value_active_masks=False)
    envs = FakeSmacEnv(cfg, 5, seed=1)
    c = _config(cfg, tmp_path, envs, env_name="StarCraft2", map_name="3m", use_eval=True, eval_interval=1, n_eval_rollout_threads=2,
                eval_episodes=3)
    ecfg = O.PathConfig(**{**cfg.to_dict(), "n_rollout_threads": 2, "act_dims": (9,)})
    c["eval_envs"] = FakeSmacEnv(ecfg, 2, seed=9)
    c["all_args"].num_env_steps = cfg.episode_length * cfg.n_rollout_threads * 2
    runner = SMACRunner(c)
    # one rollout by hand: compare the storage with the reference's NumPy bookkeeping (smac_runner.py:133-151)
    runner.warmup()
    T, N, M, H = cfg.episode_length, 5, 3, cfg.hidden_size
    want_masks, want_active, want_bad = np.ones((T + 1, N, M, 1), np.float32), np.ones((T + 1, N, M, 1), np.float32), np.ones((T + 1, N, M, 1), np.float32)
    for step in range(T):
        values, actions, lp, h_a, h_c = runner.collect(step)
        obs, share, rew, dones, infos, avail = envs.step(actions)
        runner.insert((obs, share, rew, dones, infos, avail, values, actions, lp, h_a, h_c))
        dones_env = np.all(dones, axis=1)
        m = np.ones((N, M, 1), np.float32); m[dones_env] = 0.0
        am = np.ones((N, M, 1), np.float32); am[dones] = 0.0; am[dones_env] = 1.0
        bm = np.array([[[0.0] if info[a]["bad_transition"] else [1.0] for a in range(M)] for info in infos], np.float32)
        want_masks[step + 1], want_active[step + 1], want_bad[step + 1] = m, am, bm
        assert torch.all(runner.buffer.rnn_states[step + 1][torch.from_numpy(dones_env)] == 0)
        np.testing.assert_array_equal(runner.buffer.available_actions[step + 1].cpu().numpy(), avail)
        np.testing.assert_array_equal(runner.buffer.share_obs[step + 1].cpu().numpy(), share)
    np.testing.assert_array_equal(runner.buffer.masks.cpu().numpy(), want_masks)
    np.testing.assert_array_equal(runner.buffer.active_masks.cpu().numpy(), want_active)
    np.testing.assert_array_equal(runner.buffer.bad_masks.cpu().numpy(), want_bad)
    assert 0 < want_active.mean() < 1 and want_masks.min() == 0          # the fake env exercised deaths and episode ends
    # and the whole loop: trains, logs dead_ratio / win rates, evaluates
    w0 = runner.policy.actor.flat.clone()
    runner.run()
    assert not torch.equal(w0, runner.policy.actor.flat) and torch.isfinite(runner.policy.actor.flat).all()
    sc = runner.writter.scalars if hasattr(runner.writter, "scalars") else None
    if sc is not None:
        assert any(k.startswith("dead_ratio") for k in sc) and any(k.startswith("eval_win_rate") for k in sc)
        assert any(k.startswith("incre_win_rate") for k in sc) and any(k.startswith("average_step_rewards") for k in sc)
    runner.writter.close()


@pytest.mark.parametrize("separated", [False, True])
def test_mpe_runners_evaluate(tmp_path, separated, capsys):
    """use_eval: deterministic rollouts on eval_envs, logged under the reference's keys (shared :141-183, separated :178-239)."""
    if separated:
        from onpolicy.runner.separated.mpe_runner import MPERunner
    else:
        from onpolicy.runner.shared.mpe_runner import MPERunner
    cfg = O.PathConfig(episode_length=6, n_rollout_threads=4, num_agents=2, obs_dim=6, share_obs_dim=12, act_dims=(3,), ppo_epoch=1)
    ecfg = O.PathConfig(**{**cfg.to_dict(), "n_rollout_threads": 3, "act_dims": (3,)})
    c = _config(cfg, tmp_path, FakeVecEnv(cfg, separated=separated), share_policy=not separated, use_eval=True, eval_interval=1,
                n_eval_rollout_threads=3)
    c["eval_envs"] = FakeVecEnv(ecfg, seed=5, separated=separated)
    c["all_args"].num_env_steps = cfg.episode_length * cfg.n_rollout_threads * 2
    runner = MPERunner(c)
    runner.run()
    out = capsys.readouterr().out
    assert "eval average episode rewards of agent" in out
    assert c["eval_envs"].last_actions is not None and c["eval_envs"].last_actions.shape[:2] == (3, 2)
    sc = getattr(runner.writter, "scalars", None)
    if sc is not None:
        assert any("eval_average_episode_rewards" in k for k in sc)
    runner.writter.close()


@pytest.mark.parametrize("separated", [False, True])
def test_mpe_runners_render(tmp_path, separated, capsys):
    """use_render (scripts/render/render_mpe.py): deterministic episodes, one frame per reset and per step, written under
    <run_dir>/gifs with --save_gifs (shared :185-245, separated :241-313); 'human' mode just calls envs.render every step."""
    if separated:
        from onpolicy.runner.separated.mpe_runner import MPERunner
    else:
        from onpolicy.runner.shared.mpe_runner import MPERunner
    cfg = O.PathConfig(episode_length=5, n_rollout_threads=1, num_agents=2, obs_dim=6, share_obs_dim=12, act_dims=(3,), ppo_epoch=1)
    for save in (True, False):
        env = FakeVecEnv(cfg, separated=separated)
        c = _config(cfg, tmp_path / ("gif" if save else "human"), env, share_policy=not separated, use_render=True, save_gifs=save,
                    render_episodes=2, ifi=0.0, n_render_rollout_threads=1)
        (tmp_path / ("gif" if save else "human")).mkdir(exist_ok=True)
        runner = MPERunner(c)
        runner.render()
        assert env.render_calls == 2 * (cfg.episode_length + 1)
        assert env.last_actions is not None and env.last_actions.shape[:2] == (1, 2)
        if save:
            gdir = tmp_path / "gif" / "gifs"
            files = sorted(f.name for f in gdir.iterdir())
            assert files and files[0] in ("render.gif", "render.npz")
            if files[0] == "render.npz":
                assert np.load(gdir / "render.npz")["frames"].shape == (2 * (cfg.episode_length + 1), 4, 5, 3)
    out = capsys.readouterr().out
    assert ("average episode rewards" in out)


def test_hanabi_forward_runner_matches_reference(tmp_path):
    """The turn-based runner against the UNMODIFIED reference runner driven on the same scripted environment
    (tests/golden/make_golden_hanabi.py): per-player collect on the chosen games, reward attribution to the player's NEXT move,
    chooseinsert, the one-slot reward shift, compute, train, chooseafter_update -- storage, weights, scores, step counts."""
    import ast
    import os
    from helpers import GOLDEN_DIR, assert_close
    from fake_hanabi import FakeHanabiVecEnv
    from onpolicy.runner.shared.hanabi_runner_forward import HanabiRunner
    z = np.load(os.path.join(GOLDEN_DIR, "hanabi_forward_runner.npz"), allow_pickle=False)
    S = ast.literal_eval(str(z["spec"]))
    cfg = O.PathConfig(episode_length=S["episode_length"], n_rollout_threads=S["n"], num_agents=S["players"], obs_dim=S["obs_dim"],
                       share_obs_dim=S["share_dim"], act_dims=(S["n_moves"],), ppo_epoch=S["ppo_epoch"], lr=7e-4, critic_lr=7e-4)
    envs = FakeHanabiVecEnv(S["n"], S["players"], S["obs_dim"], S["share_dim"], S["n_moves"], seed=S["env_seed"])
    c = _config(cfg, tmp_path, envs, env_name="Hanabi", hanabi_name="Hanabi-Scripted", save_interval=1000)
    c["all_args"].num_env_steps = S["episode_length"] * S["n"] * S["episodes"]
    runner = HanabiRunner(c)
    params = lambda pre: {k[len(pre):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(pre)}
    runner.policy.actor.load_state_dict(params("init/actor/"))
    runner.policy.critic.load_state_dict(params("init/critic/"))
    torch.manual_seed(S["run_seed"])
    runner.run()
    assert runner.true_total_num_steps == int(z["true_total_num_steps"]) and envs.n_steps == int(z["env_steps"])
    np.testing.assert_array_equal(np.array(runner.scores, dtype=np.float64), z["scores"])
    b = runner.buffer
    for nm in ("actions", "masks", "bad_masks", "active_masks", "available_actions", "rewards", "obs", "share_obs"):
        np.testing.assert_array_equal(getattr(b, nm).cpu().numpy(), z["buf/" + nm], err_msg=nm)      # data movement: exact
    assert_close(b.action_log_probs.cpu().numpy(), z["buf/action_log_probs"], 1e-4, 1e-5, "log-probs")
    assert_close(b.value_preds.cpu().numpy(), z["buf/value_preds"], 1e-4, 1e-5, "value_preds")
    assert_close(b.returns.cpu().numpy()[:-1], z["buf/returns"][:-1], 1e-4, 1e-4, "returns")
    for k, v in runner.policy.actor.state_dict().items():
        assert_close(v.cpu().numpy(), z[f"final/actor/{k}"], 2e-3, 2e-5, f"actor {k}")
    for k, v in runner.policy.critic.state_dict().items():
        assert_close(v.cpu().numpy(), z[f"final/critic/{k}"], 2e-3, 2e-5, f"critic {k}")
    assert_close(runner.trainer.value_normalizer.state.cpu().numpy(), z["valuenorm"], 1e-4, 1e-8, "valuenorm")
    runner.writter.close()
