"""Generate tests/golden/mpe_simple_spread.npz by running the UNMODIFIED reference MPE simple_spread environment.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_mpe.py          (needs /root/reference)

SURVEY.md section 8(f), row f1 (device-side vectorised MPE stepping).  The reference environment
(onpolicy/envs/mpe/{core,environment}.py, scenarios/simple_spread.py) is pure NumPy, but imports three modules this image
does not have -- `gym`, `seaborn` (only used for colours) and the stdlib `imp` removed in Python 3.12.  They are shimmed
here (a few attribute-only classes); nothing of the reference is modified or copied.

N environments are stepped exactly like the reference's DummyVecEnv / SubprocVecEnv worker (env_wrappers.py:140-154,
672-685): one-hot actions in, auto-reset when every agent is done, the reset observation replacing the terminal one.
Stored: the initial state of every episode (agent / landmark positions drawn by scenario.reset_world), the integer
actions, and per step the observations, rewards and dones the reference produced -- all float64, as the reference has them.
"""
import importlib.util
import os
import sys
import types
from argparse import Namespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MAPPO_REFERENCE", "/root/reference")


def shim_missing_modules():
    gym = types.ModuleType("gym")
    gym.Env = type("Env", (), {})
    gym.Space = type("Space", (), {})
    spaces = types.ModuleType("gym.spaces")

    class _Space:
        def __init__(self, *a, **k):
            self.shape = k.get("shape")
            self.n = a[0] if a and isinstance(a[0], (int, np.integer)) else None

    for name in ("Box", "Discrete", "MultiBinary", "Tuple"):
        setattr(spaces, name, type(name, (_Space,), {}))
    gym.spaces = spaces
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")
    reg.EnvSpec = type("EnvSpec", (), {})
    envs.registration = reg
    gym.envs = envs
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs, "gym.envs.registration": reg})
    sys.modules["seaborn"] = types.ModuleType("seaborn")
    imp = types.ModuleType("imp")

    def load_source(name, pathname):
        spec = importlib.util.spec_from_file_location(name or "scenario", pathname)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    imp.load_source = load_source
    sys.modules["imp"] = imp


def main():
    shim_missing_modules()
    sys.path.insert(0, REF)
    from onpolicy.envs.mpe.MPE_env import MPEEnv

    N, T, M, L, EP = 6, 64, 3, 3, 25
    args = Namespace(scenario_name="simple_spread", episode_length=EP, num_agents=M, num_landmarks=L)
    act_rng = np.random.RandomState(1234)                    # actions come from their own stream
    envs = []
    for i in range(N):
        env = MPEEnv(args)
        env.seed(1 + i * 1000)                               # train_mpe.py:32 seeds env i with seed + rank * 1000
        envs.append(env)

    def state_of(env):
        return np.concatenate([a.state.p_pos for a in env.world.agents] + [l.state.p_pos for l in env.world.landmarks])

    n_ep = T // EP + 2
    resets = np.zeros((N, n_ep, 2 * (M + L)))
    ep = np.zeros(N, dtype=np.int64)
    obs0 = np.zeros((N, M, 18))
    for i, env in enumerate(envs):
        obs0[i] = np.array(env.reset())
        resets[i, 0] = state_of(env)
    actions = act_rng.randint(0, 5, size=(T, N, M))
    obs = np.zeros((T, N, M, 18))
    rew = np.zeros((T, N, M, 1))
    done = np.zeros((T, N, M), dtype=bool)
    for t in range(T):
        for i, env in enumerate(envs):
            onehot = np.eye(5)[actions[t, i]]                # mpe_runner.py:112-119: Discrete actions go in one-hot
            o, r, d, _ = env.step(list(onehot))
            if np.all(d):                                    # env_wrappers.py:150-152
                o = env.reset()
                ep[i] += 1
                resets[i, ep[i]] = state_of(env)
            obs[t, i], rew[t, i], done[t, i] = np.array(o), np.array(r), np.array(d)
    out = os.path.join(HERE, "mpe_simple_spread.npz")
    np.savez_compressed(out, episode_length=EP, obs0=obs0, resets=resets[:, :int(ep.max()) + 1], actions=actions, obs=obs,
                        rewards=rew, dones=done)
    print("wrote", out, "episodes per env", ep.tolist(), "reward range", float(rew.min()), float(rew.max()))


def main_reference():
    """simple_reference (c3's environment): MultiDiscrete([[0,4],[0,9]]) actions, goals + communication."""
    shim_missing_modules()
    sys.path.insert(0, REF)
    from onpolicy.envs.mpe.MPE_env import MPEEnv

    N, T, M, L, EP = 6, 64, 2, 3, 25
    args = Namespace(scenario_name="simple_reference", episode_length=EP, num_agents=M, num_landmarks=L)
    act_rng = np.random.RandomState(4321)
    envs = []
    for i in range(N):
        env = MPEEnv(args)
        env.seed(1 + i * 1000)
        envs.append(env)

    def state_of(env):
        w = env.world
        goals = [w.landmarks.index(w.agents[0].goal_b), w.landmarks.index(w.agents[1].goal_b)]
        return np.concatenate([np.array(goals, dtype=np.float64)] + [a.state.p_pos for a in w.agents] + [l.state.p_pos for l in w.landmarks])

    n_ep = T // EP + 2
    resets = np.zeros((N, n_ep, 2 + 2 * (M + L)))
    ep = np.zeros(N, dtype=np.int64)
    obs0 = np.zeros((N, M, 21))
    for i, env in enumerate(envs):
        obs0[i] = np.array(env.reset())
        resets[i, 0] = state_of(env)
    actions = np.stack([act_rng.randint(0, 5, size=(T, N, M)), act_rng.randint(0, 10, size=(T, N, M))], axis=-1)
    obs = np.zeros((T, N, M, 21))
    rew = np.zeros((T, N, M, 1))
    done = np.zeros((T, N, M), dtype=bool)
    for t in range(T):
        for i, env in enumerate(envs):
            # mpe_runner.py:112-119: MultiDiscrete actions go in as the heads' one-hots concatenated
            acts = [np.concatenate([np.eye(5)[actions[t, i, m, 0]], np.eye(10)[actions[t, i, m, 1]]]) for m in range(M)]
            o, r, d, _ = env.step(acts)
            if np.all(d):
                o = env.reset()
                ep[i] += 1
                resets[i, ep[i]] = state_of(env)
            obs[t, i], rew[t, i], done[t, i] = np.array(o), np.array(r), np.array(d)
    out = os.path.join(HERE, "mpe_simple_reference.npz")
    np.savez_compressed(out, episode_length=EP, obs0=obs0, resets=resets[:, :int(ep.max()) + 1], actions=actions, obs=obs,
                        rewards=rew, dones=done)
    print("wrote", out, "episodes per env", ep.tolist(), "reward range", float(rew.min()), float(rew.max()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["spread", "reference"]
    if "spread" in which:
        main()
    if "reference" in which:
        main_reference()
