"""A scripted turn-based vec-env with the interface of the reference's Hanabi ChooseSubprocVecEnv (envs/env_wrappers.py +
envs/hanabi/Hanabi_Env.py:289-331, 348-510): `reset(choose)` restarts the chosen games (zeros for the others), `step(actions)`
advances the games whose action is not -1 (zeros, done = None for the others), rewards are [n, players, 1], dones an object
array of True / False / None, infos carry the running 'score'.  Every output is drawn from ONE seeded RandomState in call
order, so two runners that follow the same protocol see the same game streams; an illegal (unavailable) action asserts."""
import numpy as np


class _Box:
    pass


def _space(shape):
    s = type("Box", (_Box,), {})()
    s.shape = shape
    return s


class FakeHanabiVecEnv:
    def __init__(self, n, players, obs_dim, share_dim, n_moves, seed=0, max_moves=9, p_end=0.12):
        self.n, self.players, self.Do, self.Ds, self.A = n, players, obs_dim, share_dim, n_moves
        self.rng = np.random.RandomState(seed)
        self.max_moves, self.p_end = max_moves, p_end
        act = type("Discrete", (_Box,), {})()
        act.n = n_moves
        self.observation_space = [_space((obs_dim,))] * players
        self.share_observation_space = [_space((share_dim,))] * players
        self.action_space = [act] * players
        self.moves = np.zeros(n, int)
        self.score = np.zeros(n)
        self.avail = np.zeros((n, n_moves), np.float32)
        self.n_steps = 0

    def _fresh(self):
        avail = (self.rng.rand(self.A) < 0.5).astype(np.float32)
        avail[self.rng.randint(self.A)] = 1.0
        return self.rng.randn(self.Do).astype(np.float32), self.rng.randn(self.Ds).astype(np.float32), avail

    def reset(self, choose):
        obs, share, avail = np.zeros((self.n, self.Do), np.float32), np.zeros((self.n, self.Ds), np.float32), np.zeros((self.n, self.A), np.float32)
        for i in range(self.n):
            if choose[i]:
                self.moves[i], self.score[i] = 0, 0.0
                obs[i], share[i], avail[i] = self._fresh()
                self.avail[i] = avail[i]
        return obs, share, avail

    def step(self, actions):
        a = np.asarray(actions).reshape(self.n, -1)
        obs, share, avail = np.zeros((self.n, self.Do), np.float32), np.zeros((self.n, self.Ds), np.float32), np.zeros((self.n, self.A), np.float32)
        rewards = np.zeros((self.n, self.players, 1), np.float32)
        dones, infos = np.empty(self.n, dtype=object), []
        for i in range(self.n):
            ai = int(a[i, 0])
            if ai == -1:                                    # Hanabi_Env.py:459-466: the game is not stepped
                dones[i] = None
                infos.append({"score": float(self.score[i])})
                continue
            assert 0 <= ai < self.A and self.avail[i, ai] == 1.0, f"illegal move {ai} in game {i}"
            self.n_steps += 1
            self.moves[i] += 1
            r = float(self.rng.randint(0, 3) - 1)
            self.score[i] += r
            rewards[i] = r
            done = bool(self.moves[i] >= self.max_moves or self.rng.rand() < self.p_end)
            dones[i] = done
            obs[i], share[i], avail[i] = self._fresh()
            self.avail[i] = avail[i]
            infos.append({"score": float(self.score[i])})
        return obs, share, rewards, dones, infos, avail

    def close(self):
        pass
