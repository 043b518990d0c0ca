"""onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy.R_MAPPOPolicy on libmappo_b200
(reference: algorithms/r_mappo/algorithm/rMAPPOPolicy.py:6-127): same constructor, attributes and method
signatures; inputs may be NumPy or torch (any device), outputs are CUDA tensors."""
import os

import torch

from mappo_b200 import _lib
from mappo_b200.core import FusedAdam, as_dev, check, ptr, require_cuda, stream_ptr
from onpolicy.algorithms.r_mappo.algorithm.r_actor_critic import R_Actor, R_Critic
from onpolicy.utils.util import update_linear_schedule

import ctypes as C


class R_MAPPOPolicy:
    def __init__(self, args, obs_space, cent_obs_space, act_space, device=torch.device("cpu")):
        self.device = require_cuda(device if torch.device(device).type == "cuda" else None)
        self.lr = args.lr
        self.critic_lr = args.critic_lr
        self.opti_eps = args.opti_eps
        self.weight_decay = args.weight_decay
        self.obs_space = obs_space
        self.share_obs_space = cent_obs_space
        self.act_space = act_space
        # construction order = RNG consumption order of the reference (:27-28)
        self.actor = R_Actor(args, self.obs_space, self.act_space, self.device)
        self.critic = R_Critic(args, self.share_obs_space, self.device)
        self.actor_optimizer = FusedAdam(self.actor, lr=self.lr, eps=self.opti_eps, weight_decay=self.weight_decay)
        self.critic_optimizer = FusedAdam(self.critic, lr=self.critic_lr, eps=self.opti_eps,
                                          weight_decay=self.weight_decay)
        self._recurrent = bool(self.actor.desc.recurrent)
        self._H = int(args.hidden_size)
        self._recN = int(args.recurrent_N)
        self._use_policy_active_masks = bool(args.use_policy_active_masks)
        # sampling noise: "host" = Exp(1) drawn from torch's CPU generator exactly where Categorical.sample would
        # (seed-for-seed reproducible against the reference), "device" = Philox inside the kernel.
        self.rng_mode = os.environ.get("MAPPO_B200_RNG", "host")
        self.rng_seed = int(getattr(args, "seed", 1))
        self.rng_offset = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._img, self._img_key = None, None

    def lr_decay(self, episode, episodes):
        """reference :39-46."""
        update_linear_schedule(self.actor_optimizer, episode, episodes, self.lr)
        update_linear_schedule(self.critic_optimizer, episode, episodes, self.critic_lr)
        self.actor_optimizer.sync_lr()
        self.critic_optimizer.sync_lr()

    # ------------------------------------------------------------------------------------------
    def _noise(self, n_rows, deterministic):
        if deterministic or self.rng_mode != "host":
            return None
        parts = [torch.empty(n_rows, a).exponential_(1) for a in self.actor.head_dims]   # one draw per head
        return torch.cat(parts, 1).to(self.device, non_blocking=True).contiguous()

    def _step(self, cent_obs, obs, h_a, h_c, masks, avail, deterministic, want_actor, want_critic,
              out=None, exp_noise=None):
        lib = _lib.load()
        dev = self.device
        obs = as_dev(obs, dev) if want_actor else None
        cent = as_dev(cent_obs, dev) if want_critic else None
        n_rows = (obs if obs is not None else cent).shape[0]
        masks_d = as_dev(masks, dev)
        avail_d = as_dev(avail, dev) if (want_actor and avail is not None) else None
        h_a_d = as_dev(h_a, dev) if want_actor else None
        h_c_d = as_dev(h_c, dev) if want_critic else None
        a_s = len(self.actor.head_dims)
        o = out or {}
        values = o.get("values") if want_critic else None
        if want_critic and values is None:
            values = torch.empty(n_rows, 1, dtype=torch.float32, device=dev)
        actions = logp = actions_f = None
        if want_actor:
            actions = torch.empty(n_rows, a_s, dtype=torch.int64, device=dev)
            actions_f = o.get("actions")
            logp = o.get("logp")
            if logp is None:
                logp = torch.empty(n_rows, a_s, dtype=torch.float32, device=dev)
            if exp_noise is None:
                exp_noise = self._noise(n_rows, deterministic)
            else:
                exp_noise = as_dev(exp_noise, dev)
        h_a_out = h_c_out = None
        if self._recurrent:
            if want_actor:
                h_a_out = o.get("h_actor")
                if h_a_out is None:
                    h_a_out = torch.empty(n_rows, self._recN, self._H, dtype=torch.float32, device=dev)
            if want_critic:
                h_c_out = o.get("h_critic")
                if h_c_out is None:
                    h_c_out = torch.empty(n_rows, self._recN, self._H, dtype=torch.float32, device=dev)
        img_a = img_c = None
        gemm = _lib.GEMM_TF32 if os.environ.get("MAPPO_B200_GEMM", "fp32") == "tf32" else _lib.GEMM_FP32
        if True:
            # hand the kernel the packed weight image (one TMA bulk copy per CTA; feed-forward nets: the warp-per-row rollout path,
            # recurrent hidden-64 nets: the two-rows-per-warp path of rollout_gru.cuh).  Re-packed per call: the parameters may
            # have been stepped in between.
            # hidden >= 128 nets: the "image" is the workspace of the GEMM pipeline (packed weights + activations of n_rows)
            big = bool(lib.mappo_big_net(C.byref(self.actor.desc)))
            key = n_rows if big else 0
            if self._img is None or self._img_key != key:
                self._img = [torch.empty(int(lib.mappo_rollout_workspace_floats(C.byref(n.desc), max(n_rows, 1))),
                                         dtype=torch.float32, device=dev) for n in (self.actor, self.critic)]
                self._img_key = key
            if want_actor:
                img_a = self._img[0]
                check(lib.mappo_pack_rollout_weights_ex(C.byref(self.actor.desc), ptr(self.actor.flat), ptr(img_a), gemm, stream_ptr()))
            if want_critic:
                img_c = self._img[1]
                check(lib.mappo_pack_rollout_weights_ex(C.byref(self.critic.desc), ptr(self.critic.flat), ptr(img_c), gemm, stream_ptr()))
        check(lib.mappo_policy_step_ex(
            C.byref(self.actor.desc), ptr(self.actor.flat) if want_actor else None,
            C.byref(self.critic.desc), ptr(self.critic.flat) if want_critic else None,
            ptr(obs), ptr(cent), ptr(h_a_d), ptr(h_c_d), ptr(masks_d), ptr(avail_d), ptr(exp_noise),
            self.rng_seed, ptr(self.rng_offset), int(bool(deterministic)), n_rows,
            ptr(values), ptr(actions_f), ptr(actions), ptr(logp), ptr(h_a_out), ptr(h_c_out), ptr(img_a), ptr(img_c),
            gemm, stream_ptr()))
        if want_actor and exp_noise is None and not deterministic:
            check(lib.mappo_counter_add(ptr(self.rng_offset), n_rows, stream_ptr()))
        if not self._recurrent:                      # MLP policies hand the states back untouched (r_actor_critic.py:66-71)
            h_a_out, h_c_out = h_a_d, h_c_d
        return values, actions, logp, h_a_out, h_c_out

    def get_actions(self, cent_obs, obs, rnn_states_actor, rnn_states_critic, masks, available_actions=None,
                    deterministic=False):
        """reference :48-74 -> (values, actions[int64], action_log_probs, rnn_states_actor, rnn_states_critic)."""
        return self._step(cent_obs, obs, rnn_states_actor, rnn_states_critic, masks, available_actions, deterministic,
                          True, True)

    def get_values(self, cent_obs, rnn_states_critic, masks):
        """reference :76-86."""
        return self._step(cent_obs, None, None, rnn_states_critic, masks, None, True, False, True)[0]

    def act(self, obs, rnn_states_actor, masks, available_actions=None, deterministic=False):
        """reference :116-127."""
        _, actions, _, h_a, _ = self._step(None, obs, rnn_states_actor, None, masks, available_actions, deterministic,
                                           True, False)
        return actions, h_a

    def evaluate_actions(self, cent_obs, obs, rnn_states_actor, rnn_states_critic, action, masks,
                         available_actions=None, active_masks=None):
        """reference :88-114.  Gradient-free evaluation (values, log-probs, entropy) through the training kernels'
        forward half; R_MAPPO.ppo_update does not go through here -- it runs the fused forward+backward."""
        from onpolicy.algorithms.r_mappo.r_mappo import _evaluate_only
        return _evaluate_only(self, cent_obs, obs, rnn_states_actor, rnn_states_critic, action, masks,
                              available_actions, active_masks)
