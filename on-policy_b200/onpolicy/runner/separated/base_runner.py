"""onpolicy.runner.separated.base_runner.Runner on the B200 engine (reference: runner/separated/base_runner.py:15-215).

One (policy, trainer, buffer) triple per agent; agents are trained sequentially in `torch.randperm(M)` order like the
reference (:142), with the HAPPO `factor` bookkeeping around each train() (`train_agents`).
"""
import os

import numpy as np
import torch

from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy as Policy
from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO as TrainAlgo
from onpolicy.runner.shared.base_runner import make_writer, _t2n
from onpolicy.utils.separated_buffer import SeparatedReplayBuffer


def train_agents(trainers, buffers, episode_length, n_rollout_threads):
    """The body of Runner.train (reference :135-183, MAPPO branch), also used by the parity tests: agents are trained in
    `torch.randperm(M)` order (:142); each gets the running HAPPO importance `factor` (update_factor, carried and unused by
    MAPPO -- r_mappo.py:108-111), which is the product over the agents trained so far of exp(new - old log-prob) of the
    whole buffer (two gradient-free actor evaluations around each train(), :145-179).  train_infos is APPENDED in training
    order, exactly like the reference (so `log_train` labels entry i "agent i" whatever agent produced it)."""
    train_infos = []
    dev = buffers[0].device
    factor = torch.ones(episode_length, n_rollout_threads, 1, dtype=torch.float32, device=dev)
    for agent_id in torch.randperm(len(trainers)):                # same RNG draw as the reference (:142)
        agent_id = int(agent_id)
        tr, b = trainers[agent_id], buffers[agent_id]
        tr.prep_training()
        b.update_factor(factor)
        avail = None if b.available_actions is None else b.available_actions[:-1].reshape(-1, b.available_actions.shape[-1])
        ev = lambda: tr.policy.actor.evaluate_actions(
            b.obs[:-1].reshape(-1, b.obs.shape[-1]), b.rnn_states[0:1].reshape(-1, *b.rnn_states.shape[2:]),
            b.actions.reshape(-1, b.actions.shape[-1]), b.masks[:-1].reshape(-1, 1), avail,
            b.active_masks[:-1].reshape(-1, 1))[0]
        old_lp = ev()
        train_infos.append(tr.train(b))
        new_lp = ev()
        factor = factor * torch.prod(torch.exp(new_lp - old_lp), dim=-1).reshape(episode_length, n_rollout_threads, 1)
        b.after_update()
    return train_infos


class Runner(object):
    def __init__(self, config):
        a = self.all_args = config["all_args"]
        self.envs, self.eval_envs = config["envs"], config["eval_envs"]
        self.device, self.num_agents = config["device"], config["num_agents"]
        if "render_envs" in config:
            self.render_envs = config["render_envs"]
        for name in ("env_name", "algorithm_name", "experiment_name", "use_centralized_V", "use_obs_instead_of_state",
                     "num_env_steps", "episode_length", "n_rollout_threads", "n_eval_rollout_threads",
                     "use_linear_lr_decay", "hidden_size", "use_render", "recurrent_N", "save_interval", "use_eval",
                     "eval_interval", "log_interval", "model_dir"):
            setattr(self, name, getattr(a, name))
        self.use_wandb = a.use_wandb
        Policy_, TrainAlgo_ = Policy, TrainAlgo
        if self.algorithm_name == "happo":                     # reference :69-71
            from onpolicy.algorithms.happo.happo_trainer import HAPPO as TrainAlgo_
            from onpolicy.algorithms.happo.policy import HAPPO_Policy as Policy_
        elif self.algorithm_name == "hatrpo":
            raise NotImplementedError("HATRPO (trust-region step with conjugate gradients) is outside the hot path (SURVEY 2.1)")
        if self.use_render:                                     # (imageio is imported where the gif is written: render())
            self.run_dir = config["run_dir"]
            self.gif_dir = str(self.run_dir / "gifs")
            os.makedirs(self.gif_dir, exist_ok=True)
        elif self.use_wandb:
            import wandb
            self.save_dir = self.run_dir = str(wandb.run.dir)
        else:
            self.run_dir = config["run_dir"]
            self.log_dir = str(self.run_dir / "logs")
            self.save_dir = str(self.run_dir / "models")
            os.makedirs(self.log_dir, exist_ok=True)
            os.makedirs(self.save_dir, exist_ok=True)
            self.writter = make_writer(self.log_dir)

        self.policy, self.trainer, self.buffer = [], [], []
        for agent_id in range(self.num_agents):
            cent = (self.envs.share_observation_space[agent_id] if self.use_centralized_V
                    else self.envs.observation_space[agent_id])
            self.policy.append(Policy_(a, self.envs.observation_space[agent_id], cent, self.envs.action_space[agent_id],
                                       device=self.device))
        if self.model_dir is not None:
            self.restore()
        for agent_id in range(self.num_agents):
            cent = (self.envs.share_observation_space[agent_id] if self.use_centralized_V
                    else self.envs.observation_space[agent_id])
            self.trainer.append(TrainAlgo_(a, self.policy[agent_id], device=self.device))
            self.buffer.append(SeparatedReplayBuffer(a, self.envs.observation_space[agent_id], cent,
                                                     self.envs.action_space[agent_id],
                                                     device=self.policy[agent_id].device))

    def run(self):
        raise NotImplementedError

    def warmup(self):
        raise NotImplementedError

    def collect(self, step):
        raise NotImplementedError

    def insert(self, data):
        raise NotImplementedError

    @torch.no_grad()
    def compute(self):
        """reference :125-133."""
        for agent_id in range(self.num_agents):
            self.trainer[agent_id].prep_rollout()
            b = self.buffer[agent_id]
            next_value = self.trainer[agent_id].policy.get_values(b.share_obs[-1], b.rnn_states_critic[-1], b.masks[-1])
            b.compute_returns(next_value, self.trainer[agent_id].value_normalizer)

    def train(self):
        """reference :135-183 (MAPPO branch)."""
        return train_agents(self.trainer, self.buffer, self.episode_length, self.n_rollout_threads)

    def save(self):
        """reference :185-193 (file names actor_agent{i}.pt / critic_agent{i}.pt / vnrom_agent{i}.pt)."""
        for agent_id in range(self.num_agents):
            pol = self.trainer[agent_id].policy
            torch.save({k: v.cpu() for k, v in pol.actor.state_dict().items()},
                       str(self.save_dir) + "/actor_agent" + str(agent_id) + ".pt")
            torch.save({k: v.cpu() for k, v in pol.critic.state_dict().items()},
                       str(self.save_dir) + "/critic_agent" + str(agent_id) + ".pt")
            if self.trainer[agent_id]._use_valuenorm:
                vn = self.trainer[agent_id].value_normalizer
                torch.save({k: v.cpu() for k, v in vn.state_dict().items()},
                           str(self.save_dir) + "/vnrom_agent" + str(agent_id) + ".pt")

    def restore(self):
        """reference :195-204."""
        for agent_id in range(self.num_agents):
            self.policy[agent_id].actor.load_state_dict(
                torch.load(str(self.model_dir) + "/actor_agent" + str(agent_id) + ".pt", map_location="cpu"))
            self.policy[agent_id].critic.load_state_dict(
                torch.load(str(self.model_dir) + "/critic_agent" + str(agent_id) + ".pt", map_location="cpu"))

    def _log(self, tag, value, step):
        if self.use_wandb:
            import wandb
            wandb.log({tag: value}, step=step)
        else:
            self.writter.add_scalars(tag, {tag: value}, step)

    def log_train(self, train_infos, total_num_steps):
        for agent_id in range(self.num_agents):
            for k, v in train_infos[agent_id].items():
                self._log("agent%i/" % agent_id + k, float(v), total_num_steps)

    def log_env(self, env_infos, total_num_steps):
        for k, v in env_infos.items():
            if len(v) > 0:
                self._log(k, float(np.mean(v)), total_num_steps)
