"""onpolicy.runner.shared.base_runner.Runner on the B200 engine (reference: runner/shared/base_runner.py:12-187).

Same constructor contract (`config` dict with all_args / envs / eval_envs / num_agents / device / run_dir), same
attributes (`policy`, `trainer`, `buffer`, `writter`, `log_dir`, `save_dir`) and methods; the storage, the policy and
the trainer are the device-resident drop-ins, so `compute()` and `train()` issue kernels instead of NumPy/PyTorch loops.
"""
import json
import os

import numpy as np
import torch

from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy as Policy
from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO as TrainAlgo
from onpolicy.utils.shared_buffer import SharedReplayBuffer


def _t2n(x):
    """Tensor (any device) -> NumPy."""
    return x.detach().cpu().numpy()


class ScalarWriter:
    """Stand-in for tensorboardX.SummaryWriter when that package is absent: the three calls the reference makes
    (add_scalars :174,187; export_scalars_to_json / close, scripts/train/train_mpe.py:169-170)."""

    def __init__(self, log_dir):
        self.log_dir, self.scalars = log_dir, {}

    def add_scalars(self, main_tag, tag_scalar_dict, global_step=None):
        for tag, value in tag_scalar_dict.items():
            self.scalars.setdefault(f"{main_tag}/{tag}", []).append([int(global_step or 0), float(value)])

    def export_scalars_to_json(self, path):
        with open(path, "w") as f:
            json.dump(self.scalars, f)

    def close(self):
        pass


def make_writer(log_dir):
    try:
        from tensorboardX import SummaryWriter
        return SummaryWriter(log_dir)
    except Exception:
        return ScalarWriter(log_dir)


class Runner(object):
    def __init__(self, config):
        a = self.all_args = config["all_args"]
        self.envs, self.eval_envs = config["envs"], config["eval_envs"]
        self.device, self.num_agents = config["device"], config["num_agents"]
        if "render_envs" in config:
            self.render_envs = config["render_envs"]
        for name in ("env_name", "algorithm_name", "experiment_name", "use_centralized_V", "use_obs_instead_of_state",
                     "num_env_steps", "episode_length", "n_rollout_threads", "n_eval_rollout_threads",
                     "n_render_rollout_threads", "use_linear_lr_decay", "hidden_size", "use_wandb", "use_render",
                     "recurrent_N", "save_interval", "use_eval", "eval_interval", "log_interval", "model_dir"):
            setattr(self, name, getattr(a, name))
        if self.algorithm_name in ("mat", "mat_dec"):
            raise NotImplementedError("MAT is outside the B200 hot path (SURVEY 2.1 row 18)")
        if self.use_wandb:
            import wandb
            self.save_dir = self.run_dir = str(wandb.run.dir)
        else:
            self.run_dir = config["run_dir"]
            self.log_dir = str(self.run_dir / "logs")
            self.save_dir = str(self.run_dir / "models")
            os.makedirs(self.log_dir, exist_ok=True)
            os.makedirs(self.save_dir, exist_ok=True)
            self.writter = make_writer(self.log_dir)

        cent_space = self.envs.share_observation_space[0] if self.use_centralized_V else self.envs.observation_space[0]
        self.policy = Policy(a, self.envs.observation_space[0], cent_space, self.envs.action_space[0], device=self.device)
        if self.model_dir is not None:
            self.restore(self.model_dir)
        self.trainer = TrainAlgo(a, self.policy, device=self.device)
        self.buffer = SharedReplayBuffer(a, self.num_agents, self.envs.observation_space[0], cent_space,
                                         self.envs.action_space[0], device=self.policy.device)

    # hooks of the concrete runners
    def run(self):
        raise NotImplementedError

    def warmup(self):
        raise NotImplementedError

    def collect(self, step):
        raise NotImplementedError

    def insert(self, data):
        raise NotImplementedError

    def _rows(self, x):
        """[N, M, ...] slot -> [N*M, ...] rows (what np.concatenate does in the reference, :130-132)."""
        return x.reshape(-1, *x.shape[2:])

    @torch.no_grad()
    def compute(self):
        """reference :120-134: bootstrap value of slot T, then the GAE scan -- two kernels, no host round trip."""
        self.trainer.prep_rollout()
        b = self.buffer
        next_values = self.trainer.policy.get_values(self._rows(b.share_obs[-1]), self._rows(b.rnn_states_critic[-1]),
                                                     self._rows(b.masks[-1]))
        b.compute_returns(next_values.reshape(b.value_preds[-1].shape), self.trainer.value_normalizer)

    def train(self):
        """reference :136-141."""
        self.trainer.prep_training()
        train_infos = self.trainer.train(self.buffer)
        self.buffer.after_update()
        return train_infos

    def save(self, episode=0):
        """reference :143-151: actor.pt / critic.pt hold the reference's state_dict keys."""
        torch.save({k: v.cpu() for k, v in self.trainer.policy.actor.state_dict().items()}, str(self.save_dir) + "/actor.pt")
        torch.save({k: v.cpu() for k, v in self.trainer.policy.critic.state_dict().items()}, str(self.save_dir) + "/critic.pt")

    def restore(self, model_dir):
        """reference :153-162."""
        self.policy.actor.load_state_dict(torch.load(str(self.model_dir) + "/actor.pt", map_location="cpu"))
        if not self.all_args.use_render:
            self.policy.critic.load_state_dict(torch.load(str(self.model_dir) + "/critic.pt", map_location="cpu"))

    def _log(self, tag, value, step):
        if self.use_wandb:
            import wandb
            wandb.log({tag: value}, step=step)
        else:
            self.writter.add_scalars(tag, {tag: value}, step)

    def log_train(self, train_infos, total_num_steps):
        for k, v in train_infos.items():
            self._log(k, float(v), total_num_steps)

    def log_env(self, env_infos, total_num_steps):
        for k, v in env_infos.items():
            if len(v) > 0:
                self._log(k, float(np.mean(v)), total_num_steps)
