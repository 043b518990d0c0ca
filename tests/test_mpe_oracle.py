"""CPU: the environment oracle (oracle/mpe_oracle.py) against the fixture produced by the unmodified reference MPE
simple_spread environment (tests/golden/make_golden_mpe.py) -- SURVEY.md section 8(f), row f1."""
import os

import numpy as np

from oracle.mpe_oracle import SpreadVecEnv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpe_simple_spread.npz")


def replay(env, g, on_step=None):
    N, T = g["obs0"].shape[0], g["actions"].shape[0]
    ep = np.zeros(N, dtype=np.int64)
    last = g["resets"].shape[1] - 1
    out = []
    for t in range(T):
        nxt = g["resets"][np.arange(N), np.minimum(ep + 1, last)]
        o, r, d = env.step(g["actions"][t], nxt)
        ep += d[:, 0]
        out.append((o, r, d))
    return out


def test_oracle_env_is_bit_exact_against_the_reference_environment():
    """64 steps x 6 worlds, two auto-resets each, ~20 agent-agent contacts: float64 observations, rewards and dones
    equal the reference's bit for bit (same NumPy operations in the same order)."""
    g = np.load(GOLD)
    env = SpreadVecEnv(g["obs0"].shape[0], 3, 3, int(g["episode_length"]))
    assert np.array_equal(env.reset(g["resets"][:, 0]), g["obs0"])
    for t, (o, r, d) in enumerate(replay(env, g)):
        assert np.array_equal(o, g["obs"][t]), t
        assert np.array_equal(r, g["rewards"][t]), t
        assert np.array_equal(d, g["dones"][t]), t


def test_fixture_exercises_contacts_and_resets():
    g = np.load(GOLD)
    assert g["dones"].any(axis=(1, 2)).sum() == 2                       # two episode ends inside the 64 steps
    env = SpreadVecEnv(g["obs0"].shape[0], 3, 3, int(g["episode_length"]))
    env.reset(g["resets"][:, 0])
    close = 0
    N, last = g["obs0"].shape[0], g["resets"].shape[1] - 1
    ep = np.zeros(N, dtype=np.int64)
    for t in range(g["actions"].shape[0]):
        for a in range(3):
            for b in range(a + 1, 3):
                close += int((np.linalg.norm(env.apos[:, a] - env.apos[:, b], axis=1) < 0.3).sum())
        _, _, d = env.step(g["actions"][t], g["resets"][np.arange(N), np.minimum(ep + 1, last)])
        ep += d[:, 0]
    assert close >= 5                                                   # the contact-force branch is exercised


def test_reward_counts_the_self_collision_like_the_reference():
    """simple_spread.py:80-84 loops over ALL agents including the agent itself: every agent pays -1 per step."""
    env = SpreadVecEnv(1, 3, 3, 25)
    far = np.array([[-0.9, -0.9, 0.0, 0.9, 0.9, -0.9, 0.5, 0.5, -0.5, 0.5, 0.0, -0.5]])
    env.reset(far)
    _, r, _ = env.step(np.zeros((1, 3), dtype=np.int64), far)
    dists = 0.0
    for l in range(3):
        dists += min(np.linalg.norm(env.apos[0, a] - env.lpos[0, l]) for a in range(3))
    assert np.allclose(r[0, :, 0], 3 * (-dists - 1.0))


def test_reference_scenario_oracle_is_bit_exact_against_the_reference_environment():
    """simple_reference (BASELINE c3's environment): MultiDiscrete move + symbol actions, goal colours, the other agent's
    communication state in the observation, reward = summed squared goal distances; 64 steps x 6 worlds, two auto-resets."""
    from oracle.mpe_oracle import ReferenceVecEnv
    g = np.load(os.path.join(os.path.dirname(GOLD), "mpe_simple_reference.npz"))
    env = ReferenceVecEnv(g["obs0"].shape[0], int(g["episode_length"]))
    assert np.array_equal(env.reset(g["resets"][:, 0]), g["obs0"])
    for t, (o, r, d) in enumerate(replay(env, g)):
        assert np.array_equal(o, g["obs"][t]), t
        assert np.array_equal(r, g["rewards"][t]), t
        assert np.array_equal(d, g["dones"][t]), t
    assert g["dones"].any(axis=(1, 2)).sum() == 2 and g["obs"][..., 11:].sum() > 0        # resets and symbols were exercised
