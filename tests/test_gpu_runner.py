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
