"""B200: the device environment (mappo_mpe_spread_step, csrc/mpe_env.cu) against the reference-generated fixture and the
oracle, standalone and in the closed rollout loop of the engine."""
import os

import numpy as np
import pytest
import torch

from oracle import mappo_oracle as O
from oracle.mpe_oracle import SpreadVecEnv
import test_gpu_parity as TP

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpe_simple_spread.npz")
# float64 state, the reference's operation order, no FMA contraction: only exp / log1p of the contact term may differ in the
# last bits from glibc; observations / rewards are compared after the float32 cast the rollout storage applies.
RTOL, ATOL = 1e-6, 1e-6


def test_device_env_follows_the_reference_trajectories():
    from mappo_b200.mpe_env import DeviceSpreadEnv
    g = np.load(GOLD)
    N, T = g["obs0"].shape[0], g["actions"].shape[0]
    env = DeviceSpreadEnv(N, 3, 3, int(g["episode_length"]), device="cuda", seed=3)
    f = lambda *s: torch.zeros(*s, dtype=torch.float32, device="cuda")
    obs, share, rew, done = f(N * 3, 18), f(N * 3, 54), f(N * 3), f(N * 3)
    env.reset(obs, share, reset_states=g["resets"][:, 0])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(obs.cpu().numpy().reshape(N, 3, 18), g["obs0"].astype(np.float32))
    ep = np.zeros(N, dtype=np.int64)
    last = g["resets"].shape[1] - 1
    exact = 0
    for t in range(T):
        nxt = g["resets"][np.arange(N), np.minimum(ep + 1, last)]
        act = torch.from_numpy(g["actions"][t].reshape(-1).astype(np.float32)).cuda()
        env.step(act, obs, share, rew, done, reset_states=nxt)
        torch.cuda.synchronize()
        o = obs.cpu().numpy().reshape(N, 3, 18)
        np.testing.assert_allclose(o, g["obs"][t].astype(np.float32), rtol=RTOL, atol=ATOL, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy().reshape(N, 3, 1), g["rewards"][t].astype(np.float32), rtol=RTOL,
                                   atol=ATOL, err_msg=f"rewards t={t}")
        np.testing.assert_array_equal(done.cpu().numpy().reshape(N, 3) != 0, g["dones"][t])
        np.testing.assert_array_equal(share.cpu().numpy().reshape(N, 3, 54),
                                      np.repeat(o.reshape(N, 1, 54), 3, axis=1))         # mpe_runner.py:133-135
        exact += int(np.array_equal(o, g["obs"][t].astype(np.float32)))
        ep += g["dones"][t][:, 0]
    assert exact >= T - 8            # bit-identical after the fp32 cast except around contacts


def test_device_rng_resets_are_inside_the_reference_ranges():
    from mappo_b200.mpe_env import DeviceSpreadEnv
    env = DeviceSpreadEnv(512, 3, 3, 25, device="cuda", seed=7)
    obs = torch.zeros(512 * 3, 18, device="cuda")
    env.reset(obs)
    torch.cuda.synchronize()
    a, l = env.apos.cpu().numpy(), env.lpos.cpu().numpy()
    assert np.all(np.abs(a) < 1.0) and np.all(np.abs(l) < 0.8) and np.all(env.avel.cpu().numpy() == 0)
    assert abs(a.mean()) < 0.05 and abs(a.std() - 1 / np.sqrt(3)) < 0.03        # uniform(-1, 1)
    first = a.copy()
    env.reset(obs)
    torch.cuda.synchronize()
    assert not np.array_equal(first, env.apos.cpu().numpy())                    # the counter advanced


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("graph", [False, True])
def test_closed_loop_iteration_is_consistent_with_the_oracle_env(graph, fused, monkeypatch):
    """Engine in closed loop (policy_step -> device env -> insert, T times, then GAE + train): replaying the actions it
    stored through the oracle environment from the same episode starts reproduces the stored observations, share_obs,
    rewards and masks."""
    from mappo_b200.engine import RolloutEngine
    from mappo_b200.mpe_env import DeviceSpreadEnv
    monkeypatch.setenv("MAPPO_B200_PERSISTENT_ROLLOUT", "1" if fused else "0")
    cfg = O.PathConfig(episode_length=25, n_rollout_threads=15, num_agents=3, obs_dim=18, share_obs_dim=54,
                       act_dims=(5,), use_ReLU=False, ppo_epoch=2, num_mini_batch=1, lr=7e-4, critic_lr=7e-4)
    torch.manual_seed(1)
    args, policy, trainer, buf = TP.build(cfg)
    N, T = cfg.n_rollout_threads, cfg.episode_length
    env = DeviceSpreadEnv(N, 3, 3, T, device="cuda", seed=5)
    ref = SpreadVecEnv(N, 3, 3, T, seed=11)
    starts = ref.draw_reset_states(N)
    nxt = ref.draw_reset_states(N)
    eng = RolloutEngine(args, policy, trainer, buf, rng="device", seed=2, device_env=env)
    assert eng.closed_persistent == fused
    rs = torch.from_numpy(np.repeat(nxt[None], T, axis=0)).cuda()               # same restart state whenever an episode ends
    eng.env_reset_states = rs
    if graph:
        eng.reset_env(reset_states=starts)
        eng.capture(warmup=1)
    eng.reset_env(reset_states=starts)
    eng.step_resident()
    torch.cuda.synchronize()
    obs0 = ref.reset(starts)
    # storage after after_update(): slot 0 = last slot; replay from the stored actions
    acts = buf.actions.cpu().numpy().reshape(T, N, 3).astype(np.int64)
    want_obs, want_rew, want_done = [], [], []
    for t in range(T):
        o, r, d = ref.step(acts[t], nxt)
        want_obs.append(o); want_rew.append(r); want_done.append(d)
    got_obs = buf.obs.cpu().numpy()
    np.testing.assert_allclose(got_obs[1:], np.array(want_obs).astype(np.float32), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(buf.rewards.cpu().numpy(), np.array(want_rew).astype(np.float32), rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(buf.masks.cpu().numpy()[1:, :, :, 0], 1.0 - np.array(want_done).astype(np.float32))
    np.testing.assert_array_equal(buf.share_obs.cpu().numpy()[1:],
                                  np.repeat(got_obs[1:].reshape(T, N, 1, 54), 3, axis=2))
    assert want_done[-1].all()                                                  # world_length == episode_length
    assert np.isfinite(policy.actor.flat.cpu().numpy()).all()
    del obs0


def test_fused_and_per_step_closed_loops_agree_bit_for_bit(monkeypatch):
    """mappo_rollout_closed_loop (one launch) and T x [policy_step, mpe_spread_step, env_insert] run the same device
    functions with the same RNG counters: storage, world state and trained weights must be identical -- also with
    device-RNG resets (no injected episode starts) across two iterations."""
    from mappo_b200.engine import RolloutEngine
    from mappo_b200.mpe_env import DeviceSpreadEnv
    cfg = O.PathConfig(episode_length=25, n_rollout_threads=33, num_agents=3, obs_dim=18, share_obs_dim=54,
                       act_dims=(5,), use_ReLU=False, ppo_epoch=2, num_mini_batch=1, lr=7e-4, critic_lr=7e-4)
    outs = []
    for fused in (True, False):
        monkeypatch.setenv("MAPPO_B200_PERSISTENT_ROLLOUT", "1" if fused else "0")
        torch.manual_seed(1)
        args, policy, trainer, buf = TP.build(cfg)
        env = DeviceSpreadEnv(cfg.n_rollout_threads, 3, 3, cfg.episode_length, device="cuda", seed=9)
        eng = RolloutEngine(args, policy, trainer, buf, rng="device", seed=4, device_env=env)
        assert eng.closed_persistent == fused
        eng.reset_env()
        for _ in range(2):
            eng.step_resident()
        torch.cuda.synchronize()
        outs.append([t.clone() for t in (buf.obs, buf.share_obs, buf.rewards, buf.masks, buf.actions, buf.value_preds,
                                         buf.action_log_probs, env.apos, env.avel, env.lpos, env.step_count,
                                         env.rng_counter, policy.rng_offset, policy.actor.flat, policy.critic.flat)])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_device_reference_scenario_is_bit_exact_against_the_reference_trajectories():
    """mappo_mpe_reference_step against the trajectories of the unmodified reference `simple_reference` environment
    (make_golden_mpe.py main_reference): no exp / log in this scenario, so the float64 state is IEEE-exact and the float32
    outputs must be bit-identical.  Also checks share_obs and the device-RNG reset ranges."""
    from mappo_b200.mpe_env import DeviceReferenceEnv, DeviceReferenceVecEnv
    g = np.load(os.path.join(os.path.dirname(GOLD), "mpe_simple_reference.npz"))
    N, T = g["obs0"].shape[0], g["actions"].shape[0]
    env = DeviceReferenceEnv(N, int(g["episode_length"]), device="cuda", seed=3)
    f = lambda *s: torch.zeros(*s, dtype=torch.float32, device="cuda")
    obs, share, rew, done = f(N * 2, 21), f(N * 2, 42), f(N * 2), f(N * 2)
    env.reset(obs, share, reset_states=g["resets"][:, 0])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(obs.cpu().numpy().reshape(N, 2, 21), g["obs0"].astype(np.float32))
    ep = np.zeros(N, dtype=np.int64)
    last = g["resets"].shape[1] - 1
    for t in range(T):
        nxt = g["resets"][np.arange(N), np.minimum(ep + 1, last)]
        act = torch.from_numpy(np.ascontiguousarray(g["actions"][t].reshape(-1, 2).astype(np.float32))).cuda()
        env.step(act, obs, share, rew, done, reset_states=nxt)
        torch.cuda.synchronize()
        o = obs.cpu().numpy().reshape(N, 2, 21)
        np.testing.assert_array_equal(o, g["obs"][t].astype(np.float32), err_msg=f"obs t={t}")
        np.testing.assert_array_equal(rew.cpu().numpy().reshape(N, 2, 1), g["rewards"][t].astype(np.float32), err_msg=f"rewards t={t}")
        np.testing.assert_array_equal(done.cpu().numpy().reshape(N, 2) != 0, g["dones"][t])
        np.testing.assert_array_equal(share.cpu().numpy().reshape(N, 2, 42), np.repeat(o.reshape(N, 1, 42), 2, axis=1))
        ep += g["dones"][t][:, 0]
    # device RNG resets + the vec-env adapter with the runner's concatenated one-hot actions
    venv = DeviceReferenceVecEnv(256, 5, device="cuda", seed=11)
    o0 = venv.reset()
    e = venv.env
    assert np.all(np.abs(e.apos.cpu().numpy()) < 1) and np.all(np.abs(e.lpos.cpu().numpy()) < 0.8)
    gl = e.goal.cpu().numpy()
    assert gl.min() == 0 and gl.max() == 2 and set(np.unique(o0[..., 8:11])) == {0.25, 0.75} and o0[..., 11:].sum() == 0
    rng = np.random.RandomState(0)
    mv, sy = rng.randint(0, 5, (256, 2)), rng.randint(0, 10, (256, 2))
    onehot = np.concatenate([np.eye(5)[mv], np.eye(10)[sy]], axis=-1)
    o1, r1, d1, _ = venv.step(onehot)
    assert np.array_equal(o1[:, 0, 11:].argmax(-1), sy[:, 1]) and np.array_equal(o1[:, 1, 11:].argmax(-1), sy[:, 0])
    assert not d1.any() and np.all(r1 <= 0) and np.array_equal(r1[:, 0], r1[:, 1])
    for _ in range(4):
        o1, r1, d1, _ = venv.step(onehot)
    assert d1.all() and o1[..., 11:].sum() == 0 and np.all(o1[..., :2] == 0)          # auto-reset: silent, at rest
