"""onpolicy.utils.shared_buffer.SharedReplayBuffer resident in HBM.

Same constructor, attributes, methods and generator tuples as the reference (utils/shared_buffer.py:21-608),
but every array is a CUDA tensor with the reference's logical shape [T(+1), N, M, ...] and the arithmetic
(compute_returns' GAE scan, the minibatch gathers) runs in libmappo_b200 kernels.
"""
import ctypes as C

import numpy as np
import torch

from mappo_b200 import _lib
from mappo_b200.core import as_dev, check, obs_dim_of, ptr, require_cuda, stream_ptr
from onpolicy.utils.util import get_shape_from_act_space


def _flatten(T, N, x):
    return x.reshape(T * N, *x.shape[2:])


class SharedReplayBuffer(object):
    """Rollout storage (reference :21-88). `device` is an extension: default = current CUDA device."""

    _SEPARATED = False

    def __init__(self, args, num_agents, obs_space, cent_obs_space, act_space, device=None):
        self.device = require_cuda(device)
        _lib.load()
        self.episode_length = args.episode_length
        self.n_rollout_threads = args.n_rollout_threads
        self.hidden_size = args.hidden_size
        self.recurrent_N = args.recurrent_N
        self.gamma = args.gamma
        self.gae_lambda = args.gae_lambda
        self._use_gae = args.use_gae
        self._use_popart = args.use_popart
        self._use_valuenorm = args.use_valuenorm
        self._use_proper_time_limits = args.use_proper_time_limits
        self.algo = args.algorithm_name
        self.num_agents = num_agents
        if self.algo in ("mat", "mat_dec"):
            raise NotImplementedError("MAT is outside the B200 hot path (SURVEY 2.1 row 18)")
        if self._use_popart:
            raise NotImplementedError("use_popart: PopArt.update raises in the reference itself (SURVEY App. B-7); "
                                      "ValueNorm is the supported normaliser")
        T, N, M = self.episode_length, self.n_rollout_threads, num_agents
        Do, Ds = obs_dim_of(obs_space), obs_dim_of(cent_obs_space)
        lead = (N,) if self._SEPARATED else (N, M)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        o = lambda *s: torch.ones(*s, dtype=torch.float32, device=self.device)
        self.share_obs = z(T + 1, *lead, Ds)
        self.obs = z(T + 1, *lead, Do)
        self.rnn_states = z(T + 1, *lead, self.recurrent_N, self.hidden_size)
        self.rnn_states_critic = torch.zeros_like(self.rnn_states)
        self.value_preds = z(T + 1, *lead, 1)
        self.returns = torch.zeros_like(self.value_preds)
        self.advantages = z(T, *lead, 1)
        if act_space.__class__.__name__ == "Discrete":
            self.available_actions = o(T + 1, *lead, act_space.n)
        else:
            self.available_actions = None
        act_shape = get_shape_from_act_space(act_space)
        self.actions = z(T, *lead, act_shape)
        self.action_log_probs = z(T, *lead, act_shape)
        self.rewards = z(T, *lead, 1)
        self.masks = o(T + 1, *lead, 1)
        self.bad_masks = torch.ones_like(self.masks)
        self.active_masks = torch.ones_like(self.masks)
        self.step = 0
        # engine-side extras
        self._E = int(np.prod(lead))
        self._adv_stats = torch.zeros(3, dtype=torch.float64, device=self.device)
        self._adv_version = -1          # bumps when compute_returns produced advantages + statistics

    # ------------------------------------------------------------------------------------------
    def _put(self, dst, src):
        dst.copy_(as_dev(src, self.device).reshape(dst.shape), non_blocking=True)

    def insert(self, share_obs, obs, rnn_states_actor, rnn_states_critic, actions, action_log_probs,
               value_preds, rewards, masks, bad_masks=None, active_masks=None, available_actions=None):
        """reference :90-123."""
        s = self.step
        self._put(self.share_obs[s + 1], share_obs)
        self._put(self.obs[s + 1], obs)
        self._put(self.rnn_states[s + 1], rnn_states_actor)
        self._put(self.rnn_states_critic[s + 1], rnn_states_critic)
        self._put(self.actions[s], actions)
        self._put(self.action_log_probs[s], action_log_probs)
        self._put(self.value_preds[s], value_preds)
        self._put(self.rewards[s], rewards)
        self._put(self.masks[s + 1], masks)
        if bad_masks is not None:
            self._put(self.bad_masks[s + 1], bad_masks)
        if active_masks is not None:
            self._put(self.active_masks[s + 1], active_masks)
        if available_actions is not None:
            self._put(self.available_actions[s + 1], available_actions)
        self.step = (s + 1) % self.episode_length
        self._adv_version = -1

    def chooseinsert(self, share_obs, obs, rnn_states, rnn_states_critic, actions, action_log_probs,
                     value_preds, rewards, masks, bad_masks=None, active_masks=None, available_actions=None):
        """Turn-based (Hanabi) insert, reference :125-158."""
        s = self.step
        self._put(self.share_obs[s], share_obs)
        self._put(self.obs[s], obs)
        self._put(self.rnn_states[s + 1], rnn_states)
        self._put(self.rnn_states_critic[s + 1], rnn_states_critic)
        self._put(self.actions[s], actions)
        self._put(self.action_log_probs[s], action_log_probs)
        self._put(self.value_preds[s], value_preds)
        self._put(self.rewards[s], rewards)
        self._put(self.masks[s + 1], masks)
        if bad_masks is not None:
            self._put(self.bad_masks[s + 1], bad_masks)
        if active_masks is not None:
            self._put(self.active_masks[s], active_masks)
        if available_actions is not None:
            self._put(self.available_actions[s], available_actions)
        self.step = (s + 1) % self.episode_length
        self._adv_version = -1

    def after_update(self):
        """reference :160-170: slot T -> slot 0 (8 small device-to-device copies)."""
        for a in (self.share_obs, self.obs, self.rnn_states, self.rnn_states_critic, self.masks, self.bad_masks,
                  self.active_masks, self.available_actions):
            if a is not None:
                a[0].copy_(a[-1], non_blocking=True)

    def chooseafter_update(self):
        """reference :172-177."""
        for a in (self.rnn_states, self.rnn_states_critic, self.masks, self.bad_masks):
            a[0].copy_(a[-1], non_blocking=True)

    # ------------------------------------------------------------------------------------------
    def compute_returns(self, next_value, value_normalizer=None):
        """reference :179-262 (non-MAT branches) as ONE kernel: backward GAE / discounted scan with the
        ValueNorm denormalisation folded in; also leaves raw advantages + their masked statistics behind for
        R_MAPPO.train (r_mappo.py:179-187)."""
        lib = _lib.load()
        T, E = self.episode_length, self._E
        if self._use_gae:
            self._put(self.value_preds[-1], next_value)
        else:
            self._put(self.value_preds[-1], next_value)     # the kernel seeds returns[-1] from this slot
        vn = None
        if (self._use_popart or self._use_valuenorm) and value_normalizer is not None:
            vn = value_normalizer.state
        self._adv_stats.zero_()
        check(lib.mappo_compute_returns(ptr(self.rewards), ptr(self.value_preds), ptr(self.masks), ptr(self.bad_masks),
                                        ptr(self.active_masks), ptr(vn), T, E, float(self.gamma),
                                        float(self.gae_lambda), int(bool(self._use_gae)),
                                        int(bool(self._use_proper_time_limits)), ptr(self.returns),
                                        ptr(self.advantages), ptr(self._adv_stats), stream_ptr()))
        self._adv_version = id(value_normalizer) if value_normalizer is not None else 0

    # ------------------------------------------------------------------------------------------
    # minibatch generators: same 12-tuples as the reference, as CUDA tensors
    # ------------------------------------------------------------------------------------------
    def _tables(self, advantages):
        B = self.episode_length * self._E
        f = lambda a, n: a[:n].reshape(B, -1)
        T = self.episode_length
        H = self.hidden_size
        return dict(
            share_obs=f(self.share_obs, T), obs=f(self.obs, T),
            rnn_states=self.rnn_states[:T].reshape(B, self.recurrent_N * H),
            rnn_states_critic=self.rnn_states_critic[:T].reshape(B, self.recurrent_N * H),
            actions=f(self.actions, T), value_preds=f(self.value_preds, T), returns=f(self.returns, T),
            masks=f(self.masks, T), active_masks=f(self.active_masks, T),
            action_log_probs=f(self.action_log_probs, T),
            advantages=None if advantages is None else as_dev(advantages, self.device).reshape(B, 1),
            available_actions=None if self.available_actions is None else f(self.available_actions, T)), B

    _ORDER = ("share_obs", "obs", "rnn_states", "rnn_states_critic", "actions", "value_preds", "returns", "masks",
              "active_masks", "action_log_probs", "advantages", "available_actions")

    def _gather(self, src, rows):
        lib = _lib.load()
        src = src.contiguous()
        out = torch.empty(rows.numel(), src.shape[1], dtype=torch.float32, device=self.device)
        check(lib.mappo_gather_rows(ptr(src), ptr(rows), rows.numel(), src.shape[1], ptr(out), stream_ptr()))
        return out

    def _emit(self, tab, rows, first):
        out = []
        for name in self._ORDER:
            a = tab[name]
            if a is None:
                out.append(None)
            elif name in ("rnn_states", "rnn_states_critic"):
                out.append(self._gather(a, first).view(first.numel(), self.recurrent_N, self.hidden_size))
            else:
                out.append(self._gather(a, rows))
        factor = getattr(self, "factor", None)            # separated buffers: 13th element once update_factor() was called
        if factor is not None:                            # (reference utils/separated_buffer.py:197-227)
            out.append(self._gather(factor.reshape(self.episode_length * self._E, -1), rows))
        return tuple(out)

    def feed_forward_generator(self, advantages, num_mini_batch=None, mini_batch_size=None):
        """reference :340-400.  torch.randperm is drawn on the CPU generator exactly where the reference does."""
        batch_size = self.episode_length * self._E
        if mini_batch_size is None:
            assert batch_size >= num_mini_batch, (
                "PPO requires the number of processes ({}) * number of steps ({}) * number of agents ({}) = {} "
                "to be greater than or equal to the number of PPO mini batches ({}).".format(
                    self.n_rollout_threads, self.episode_length, self.num_agents, batch_size, num_mini_batch))
            mini_batch_size = batch_size // num_mini_batch
        rand = torch.randperm(batch_size).to(torch.int32).to(self.device, non_blocking=True)
        tab, _ = self._tables(advantages)
        for i in range(num_mini_batch):
            rows = rand[i * mini_batch_size:(i + 1) * mini_batch_size].contiguous()
            yield self._emit(tab, rows, rows)

    def _chunk_rows(self, chunks, L):
        lib = _lib.load()
        n = chunks.numel()
        rows = torch.empty(L * n, dtype=torch.int32, device=self.device)
        first = torch.empty(n, dtype=torch.int32, device=self.device)
        check(lib.mappo_chunk_rows(ptr(chunks), n, L, self.episode_length, self._E, ptr(rows), ptr(first),
                                   stream_ptr()))
        return rows, first

    def naive_recurrent_generator(self, advantages, num_mini_batch):
        """reference :402-497: whole trajectories, time-major [T, Nc], initial state = slot 0."""
        batch_size = self._E
        assert batch_size >= num_mini_batch, (
            "PPO requires the number of processes ({})* number of agents ({}) to be greater than or equal to the "
            "number of PPO mini batches ({}).".format(self.n_rollout_threads, self.num_agents, num_mini_batch))
        per = batch_size // num_mini_batch
        perm = torch.randperm(batch_size).to(torch.int32).to(self.device, non_blocking=True)
        tab, _ = self._tables(advantages)
        for start in range(0, batch_size, per):
            lanes = perm[start:start + per].contiguous()
            if lanes.numel() < per:
                break
            rows, first = self._chunk_rows(lanes, self.episode_length)
            yield self._emit(tab, rows, first)

    def recurrent_generator(self, advantages, num_mini_batch, data_chunk_length):
        """reference :499-608: (n,m,t)-ordered rows cut into L-step chunks (straddling trajectories when
        T % L != 0), chunks permuted, minibatch laid out time-major [L, Nc]."""
        batch_size = self.episode_length * self._E
        data_chunks = batch_size // data_chunk_length
        mini_batch_size = data_chunks // num_mini_batch
        rand = torch.randperm(data_chunks).to(torch.int32).to(self.device, non_blocking=True)
        tab, _ = self._tables(advantages)
        for i in range(num_mini_batch):
            chunks = rand[i * mini_batch_size:(i + 1) * mini_batch_size].contiguous()
            rows, first = self._chunk_rows(chunks, data_chunk_length)
            yield self._emit(tab, rows, first)
