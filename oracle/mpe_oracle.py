"""CPU restatement of the reference's MPE `simple_spread` environment, vectorised over N environments.

TEST INFRASTRUCTURE ONLY (like oracle/mappo_oracle.py): imported by tests/ and bench.py's CPU legs, never by the product
path.  Pinned against tests/golden/mpe_simple_spread.npz, which tests/golden/make_golden_mpe.py produced by running the
unmodified reference environment.

Follows, line by line, in float64 and in the reference's order of operations:
  * action decoding          envs/mpe/environment.py:203-262 (_set_action: one-hot in, u = 5 * (a1 - a2, a3 - a4))
  * World.step               envs/mpe/core.py:207-226, apply_action_force :229-238, apply_environment_force :241-265,
                             integrate_state :267-281, get_entity_collision_force :293-323
  * reward / observation     envs/mpe/scenarios/simple_spread.py:72-103 (the self-"collision" of an agent with itself is
                             counted, as in the reference), shared reward = sum over agents environment.py:139-142
  * done                     environment.py:171-177 (current_step >= world_length)
  * auto-reset               envs/env_wrappers.py:146-152 (the reset observation replaces the terminal one)
  * reset_world              simple_spread.py:32-45 (uniform(-1, 1) agents, 0.8 * uniform(-1, 1) landmarks) -- the DRAWS are
                             an input here (`reset_states`), because NumPy's global Mersenne stream cannot be reproduced
                             on a device; without them a private RandomState draws in the reference's order.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

AGENT_SIZE = 0.15          # simple_spread.py:22
CONTACT_FORCE = 1e2        # core.py:128
CONTACT_MARGIN = 1e-3      # core.py:129
DAMPING = 0.25             # core.py:126
DT = 0.1                   # core.py:124
SENSITIVITY = 5.0          # environment.py:243


class SpreadVecEnv:
    """N independent simple_spread worlds with M agents and L landmarks (reference defaults M = L = 3)."""

    def __init__(self, n_envs: int, num_agents: int = 3, num_landmarks: int = 3, episode_length: int = 25, seed: int = 0):
        self.N, self.M, self.L, self.EP = n_envs, num_agents, num_landmarks, episode_length
        self.rng = np.random.RandomState(seed)
        self.apos = np.zeros((n_envs, num_agents, 2))
        self.avel = np.zeros((n_envs, num_agents, 2))
        self.lpos = np.zeros((n_envs, num_landmarks, 2))
        self.step_count = np.zeros(n_envs, dtype=np.int64)
        self.obs_dim = 4 + 2 * num_landmarks + 4 * (num_agents - 1)

    # -- reset --------------------------------------------------------------------------------------
    def draw_reset_states(self, n: int) -> np.ndarray:
        """[n, 2 (M + L)]: agent positions then landmark positions, drawn in reset_world's order."""
        out = np.zeros((n, 2 * (self.M + self.L)))
        for i in range(n):
            a = [self.rng.uniform(-1, +1, 2) for _ in range(self.M)]
            l = [0.8 * self.rng.uniform(-1, +1, 2) for _ in range(self.L)]
            out[i] = np.concatenate(a + l)
        return out

    def _set_states(self, idx, states):
        s = np.asarray(states, dtype=np.float64).reshape(len(idx), self.M + self.L, 2)
        self.apos[idx] = s[:, :self.M]
        self.lpos[idx] = s[:, self.M:]
        self.avel[idx] = 0.0
        self.step_count[idx] = 0

    def reset(self, reset_states: Optional[np.ndarray] = None) -> np.ndarray:
        if reset_states is None:
            reset_states = self.draw_reset_states(self.N)
        self._set_states(np.arange(self.N), reset_states)
        return self.observe()

    # -- observation / reward ---------------------------------------------------------------------------
    def observe(self) -> np.ndarray:
        """[N, M, obs_dim] = vel, pos, landmarks - pos, other agents - pos, other agents' (silent) comm."""
        N, M = self.N, self.M
        out = np.zeros((N, M, self.obs_dim))
        for m in range(M):
            parts = [self.avel[:, m], self.apos[:, m]]
            parts += [self.lpos[:, l] - self.apos[:, m] for l in range(self.L)]
            parts += [self.apos[:, o] - self.apos[:, m] for o in range(M) if o != m]
            parts += [np.zeros((N, 2)) for o in range(M) if o != m]
            out[:, m] = np.concatenate(parts, axis=1)
        return out

    def rewards(self) -> np.ndarray:
        """[N, M, 1]: every agent receives the sum of the individual rewards (world.collaborative)."""
        N, M = self.N, self.M
        indiv = np.zeros((N, M))
        for m in range(M):
            rew = np.zeros(N)
            for l in range(self.L):
                d = [np.sqrt(np.sum(np.square(self.apos[:, a] - self.lpos[:, l]), axis=1)) for a in range(M)]
                rew = rew - np.minimum.reduce(d)
            for a in range(M):                                   # includes a == m: dist 0 < 0.3 (reference quirk)
                delta = self.apos[:, a] - self.apos[:, m]
                dist = np.sqrt(np.sum(np.square(delta), axis=1))
                rew = rew - (dist < 2 * AGENT_SIZE).astype(np.float64)
            indiv[:, m] = rew
        total = indiv[:, 0]
        for m in range(1, M):
            total = total + indiv[:, m]                          # np.sum over 3 values = left-to-right
        return np.repeat(total[:, None, None], M, axis=1)

    # -- step ---------------------------------------------------------------------------------------
    def step(self, actions: np.ndarray, reset_states: Optional[np.ndarray] = None):
        """actions [N, M] integers in 0..4.  Returns obs [N, M, D], rewards [N, M, 1], dones [N, M] (bool).
        reset_states [N, 2 (M + L)]: the state an environment restarts from if this step ends its episode."""
        N, M = self.N, self.M
        a = np.asarray(actions).reshape(N, M).astype(np.int64)
        onehot = np.eye(5)[a]
        u = np.zeros((N, M, 2))
        u[:, :, 0] += onehot[:, :, 1] - onehot[:, :, 2]
        u[:, :, 1] += onehot[:, :, 3] - onehot[:, :, 4]
        u *= SENSITIVITY
        force = 1.0 * u + 0.0                                     # mass * u + noise (u_noise is None)
        for ia in range(M):
            for ib in range(ia + 1, M):                           # landmarks do not collide
                delta = self.apos[:, ia] - self.apos[:, ib]
                dist = np.sqrt(np.sum(np.square(delta), axis=1))
                k = CONTACT_MARGIN
                pen = np.logaddexp(0, -(dist - 2 * AGENT_SIZE) / k) * k
                f = CONTACT_FORCE * delta / dist[:, None] * pen[:, None]
                force[:, ia] = 1.0 * f + force[:, ia]
                force[:, ib] = -(1 / 1.0) * f + force[:, ib]
        self.avel = self.avel * (1 - DAMPING)
        self.avel = self.avel + (force / 1.0) * DT
        self.apos = self.apos + self.avel * DT
        self.step_count += 1
        rew = self.rewards()
        done_env = self.step_count >= self.EP
        if np.any(done_env):
            idx = np.nonzero(done_env)[0]
            if reset_states is None:
                rs = self.draw_reset_states(len(idx))
            else:
                rs = np.asarray(reset_states)[idx]
            self._set_states(idx, rs)
        obs = self.observe()
        return obs, rew, np.repeat(done_env[:, None], M, axis=1)


class ReferenceVecEnv:
    """N independent MPE `simple_reference` worlds (2 agents, 3 landmarks, 10 communication symbols), float64, in the reference's
    order of operations.  Pinned against tests/golden/mpe_simple_reference.npz (make_golden_mpe.py, unmodified reference env).

      * scenario               envs/mpe/scenarios/simple_reference.py:8-97: agents and landmarks do not collide; agent i wants
                               the OTHER agent at landmark goal[i]; reward_i = -|other.pos - landmark[goal_i].pos|^2, both
                               agents receive the sum (collaborative); observation = own velocity, landmark positions relative
                               to the agent, the colour of its goal landmark, the other agent's communication state
      * action decoding        envs/mpe/environment.py:184-250 with a MultiDiscrete([[0,4],[0,9]]) space: the runner sends the
                               two heads one-hot, concatenated (mpe_runner.py:112-119); u = 5 (a1 - a2, a3 - a4), c = one-hot
      * World.step             envs/mpe/core.py:207-226, :229-238, :267-281 (no contact forces), update_agent_state :283-290
                               (state.c = action.c, no noise)
      * done / auto-reset      as SpreadVecEnv
      * reset_world            :35-60: goal landmarks by np.random.choice, then positions -- the DRAWS are an input
                               (`reset_states` [N, 2 + 2 (2 + 3)]: goal_0, goal_1, agent positions, landmark positions)."""
    M, L, DIM_C = 2, 3, 10
    COLORS = np.array([[0.75, 0.25, 0.25], [0.25, 0.75, 0.25], [0.25, 0.25, 0.75]])        # simple_reference.py:46-48

    def __init__(self, n_envs: int, episode_length: int = 25, seed: int = 0):
        self.N, self.EP = n_envs, episode_length
        self.rng = np.random.RandomState(seed)
        self.apos, self.avel = np.zeros((n_envs, 2, 2)), np.zeros((n_envs, 2, 2))
        self.lpos = np.zeros((n_envs, 3, 2))
        self.goal = np.zeros((n_envs, 2), dtype=np.int64)
        self.comm = np.zeros((n_envs, 2, self.DIM_C))
        self.step_count = np.zeros(n_envs, dtype=np.int64)
        self.obs_dim = 2 + 2 * self.L + 3 + self.DIM_C           # 21

    def draw_reset_states(self, n: int) -> np.ndarray:
        out = np.zeros((n, 2 + 2 * (self.M + self.L)))
        for i in range(n):
            g = [self.rng.randint(0, self.L), self.rng.randint(0, self.L)]
            a = [self.rng.uniform(-1, +1, 2) for _ in range(self.M)]
            l = [0.8 * self.rng.uniform(-1, +1, 2) for _ in range(self.L)]
            out[i] = np.concatenate([np.array(g, dtype=np.float64)] + a + l)
        return out

    def _set_states(self, idx, states):
        s = np.asarray(states, dtype=np.float64).reshape(len(idx), -1)
        self.goal[idx] = s[:, :2].astype(np.int64)
        p = s[:, 2:].reshape(len(idx), self.M + self.L, 2)
        self.apos[idx], self.lpos[idx] = p[:, :self.M], p[:, self.M:]
        self.avel[idx] = 0.0
        self.comm[idx] = 0.0
        self.step_count[idx] = 0

    def reset(self, reset_states: Optional[np.ndarray] = None) -> np.ndarray:
        if reset_states is None:
            reset_states = self.draw_reset_states(self.N)
        self._set_states(np.arange(self.N), reset_states)
        return self.observe()

    def observe(self) -> np.ndarray:
        out = np.zeros((self.N, self.M, self.obs_dim))
        for m in range(self.M):
            parts = [self.avel[:, m]] + [self.lpos[:, l] - self.apos[:, m] for l in range(self.L)]
            parts += [self.COLORS[self.goal[:, m]], self.comm[:, 1 - m]]
            out[:, m] = np.concatenate(parts, axis=1)
        return out

    def rewards(self) -> np.ndarray:
        idx = np.arange(self.N)
        indiv = []
        for m in range(self.M):
            d = self.apos[:, 1 - m] - self.lpos[idx, self.goal[:, m]]
            indiv.append(-np.sum(np.square(d), axis=1))
        total = indiv[0] + indiv[1]
        return np.repeat(total[:, None, None], self.M, axis=1)

    def step(self, actions: np.ndarray, reset_states: Optional[np.ndarray] = None):
        """actions [N, 2, 2] integers (move in 0..4, symbol in 0..9).  Returns obs [N, 2, 21], rewards [N, 2, 1], dones [N, 2]."""
        N, M = self.N, self.M
        a = np.asarray(actions).reshape(N, M, 2).astype(np.int64)
        onehot = np.eye(5)[a[:, :, 0]]
        u = np.zeros((N, M, 2))
        u[:, :, 0] += onehot[:, :, 1] - onehot[:, :, 2]
        u[:, :, 1] += onehot[:, :, 3] - onehot[:, :, 4]
        u *= SENSITIVITY
        force = 1.0 * u + 0.0
        self.avel = self.avel * (1 - DAMPING)
        self.avel = self.avel + (force / 1.0) * DT
        self.apos = self.apos + self.avel * DT
        self.comm = np.eye(self.DIM_C)[a[:, :, 1]] + 0.0
        self.step_count += 1
        rew = self.rewards()
        done_env = self.step_count >= self.EP
        if np.any(done_env):
            idx = np.nonzero(done_env)[0]
            rs = self.draw_reset_states(len(idx)) if reset_states is None else np.asarray(reset_states)[idx]
            self._set_states(idx, rs)
        return self.observe(), rew, np.repeat(done_env[:, None], M, axis=1)
