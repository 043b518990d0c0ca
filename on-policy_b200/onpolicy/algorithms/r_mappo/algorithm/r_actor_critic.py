"""R_Actor / R_Critic stand-ins (reference: algorithms/r_mappo/algorithm/r_actor_critic.py:12-175).

The networks have no Python forward: parameters sit in one flat CUDA vector (mappo_b200.core.DeviceNet) that the
library's rollout and training kernels read directly.  These classes exist so `policy.actor.state_dict()`,
`load_state_dict`, `.parameters()`, `.train()/.eval()` keep working for the runner's save / restore
(runner/shared/base_runner.py:143-162).
"""
from mappo_b200.core import DeviceNet, act_heads_of, obs_dim_of


class R_Actor(DeviceNet):
    def __init__(self, args, obs_space, action_space, device=None):
        heads, multi = act_heads_of(action_space)
        self._multi = multi
        setattr(args, "_b200_multi_discrete", multi)
        super().__init__(args, obs_dim_of(obs_space), heads, is_critic=False, device=device)
        self.multi_discrete = multi
        self._keys = self._key_table()
        self.init_like_reference(args)
        self.algo = args.algorithm_name
        self._use_policy_active_masks = bool(args.use_policy_active_masks)

    def evaluate_actions(self, obs, rnn_states, action, masks, available_actions=None, active_masks=None):
        """reference :73-117 -> (action_log_probs [rows, as], dist_entropy); gradient free (the separated runner's factor
        bookkeeping, runner/separated/base_runner.py:145-179, is its only caller outside training)."""
        from onpolicy.algorithms.r_mappo.r_mappo import _evaluate_actor
        return _evaluate_actor(self, obs, rnn_states, action, masks, available_actions, active_masks,
                               self._use_policy_active_masks)


class R_Critic(DeviceNet):
    def __init__(self, args, cent_obs_space, device=None):
        if args.use_popart:
            raise NotImplementedError("use_popart: PopArt.update raises in the reference itself (SURVEY App. B-7)")
        setattr(args, "_b200_multi_discrete", False)
        super().__init__(args, obs_dim_of(cent_obs_space), [1], is_critic=True, device=device)
        self.multi_discrete = False
        self._keys = self._key_table()
        self.init_like_reference(args)
