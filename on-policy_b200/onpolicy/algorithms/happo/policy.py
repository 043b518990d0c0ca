"""onpolicy.algorithms.happo.policy.HAPPO_Policy (reference: algorithms/happo/policy.py -- the MAPPO policy wrapper under
another name: same actor / critic, optimisers, get_actions / get_values / evaluate_actions / act)."""
from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy


class HAPPO_Policy(R_MAPPOPolicy):
    def __init__(self, args, obs_space, cent_obs_space, act_space, device=None):
        import torch
        super().__init__(args, obs_space, cent_obs_space, act_space, device=device if device is not None else torch.device("cpu"))
        self.args = args
