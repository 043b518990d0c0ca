"""Build the argparse-Namespace the drop-in classes read (the reference's config.py defaults) from a PathConfig."""
from argparse import Namespace

from oracle import mappo_oracle as O


def make_args(cfg: O.PathConfig, algo=None, seed=1) -> Namespace:
    if algo is None:
        algo = "rmappo" if cfg.use_recurrent_policy else "mappo"
    d = cfg.to_dict()
    a = Namespace(algorithm_name=algo, experiment_name="check", seed=seed, cuda=True, n_training_threads=1,
                  env_name="MPE", use_popart=False, use_orthogonal=True, weight_decay=0, stacked_frames=1,
                  use_stacked_frames=False, use_centralized_V=True, use_obs_instead_of_state=False,
                  use_linear_lr_decay=False, num_env_steps=1e6, n_eval_rollout_threads=1, n_render_rollout_threads=1,
                  use_wandb=False, use_render=False, save_interval=1, use_eval=False, eval_interval=25,
                  log_interval=5, model_dir=None, share_policy=True, scenario_name="simple_spread")
    for k, v in d.items():
        setattr(a, k, v)
    return a


class _Space:
    pass


def make_spaces(cfg: O.PathConfig):
    import numpy as np
    Box = type("Box", (_Space,), {})
    obs, share = Box(), Box()
    obs.shape, share.shape = (cfg.obs_dim,), (cfg.share_obs_dim,)
    if cfg.multi_discrete:
        act = type("MultiDiscrete", (_Space,), {})()
        act.high = np.array([a - 1 for a in cfg.act_dims])
        act.low = np.zeros(len(cfg.act_dims), dtype=np.int64)
        act.shape = len(cfg.act_dims)
    else:
        act = type("Discrete", (_Space,), {})()
        act.n = cfg.act_dims[0]
    return obs, share, act
