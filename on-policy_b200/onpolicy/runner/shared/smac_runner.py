"""onpolicy.runner.shared.smac_runner.SMACRunner on the B200 engine (reference: runner/shared/smac_runner.py:11-251).

Same loop, hooks and log keys as the reference.  The environment stays a host vec-env (StarCraft II is not on the GPU);
per step the env's obs / state / reward / done / available-action arrays go up in ONE device copy each, the sampled
actions come down, and the SMAC bookkeeping of `insert` (:129-151) -- env-level done -> rnn reset and masks, per-agent
death -> active_masks (forced back to 1 on an env-level done), bad_transition -> bad_masks -- is a handful of device-side
tensor operations on [N, M, 1] masks feeding SharedReplayBuffer.insert; values, log-probs and rnn states never leave HBM.
"""
import time
from functools import reduce

import numpy as np
import torch

from onpolicy.runner.shared.base_runner import Runner, _t2n


def smac_masks(dones, bad_transition, device):
    """The three mask arrays of reference :133-145 from the env's per-agent `dones` [N, M] (bool) and the infos'
    `bad_transition` flags [N, M] (bool): (dones_env [N] bool, masks, active_masks, bad_masks each [N, M, 1] fp32)."""
    d = torch.as_tensor(np.asarray(dones, dtype=bool), device=device)
    bad = torch.as_tensor(np.asarray(bad_transition, dtype=bool), device=device)
    dones_env = d.all(dim=1)
    one = torch.ones(*d.shape, 1, dtype=torch.float32, device=device)
    masks = one.clone()
    masks[dones_env] = 0.0                                   # :138-139
    active = one.clone()
    active[d] = 0.0                                          # :141-142 dead agents stop contributing ...
    active[dones_env] = 1.0                                  # :143     ... except on the step their episode ends
    bad_masks = one.clone()
    bad_masks[bad] = 0.0                                     # :145
    return dones_env, masks, active, bad_masks


class SMACRunner(Runner):
    """Runner class to perform training, evaluation and data collection for SMAC (reference :11-15)."""

    def __init__(self, config):
        super(SMACRunner, self).__init__(config)

    def run(self):
        """reference :17-101."""
        self.warmup()
        start = time.time()
        episodes = int(self.num_env_steps) // self.episode_length // self.n_rollout_threads
        last_battles_game = np.zeros(self.n_rollout_threads, dtype=np.float32)
        last_battles_won = np.zeros(self.n_rollout_threads, dtype=np.float32)
        for episode in range(episodes):
            if self.use_linear_lr_decay:
                self.trainer.policy.lr_decay(episode, episodes)
            for step in range(self.episode_length):
                values, actions, action_log_probs, rnn_states, rnn_states_critic = self.collect(step)
                obs, share_obs, rewards, dones, infos, available_actions = self.envs.step(actions)
                self.insert((obs, share_obs, rewards, dones, infos, available_actions, values, actions, action_log_probs,
                             rnn_states, rnn_states_critic))
            self.compute()
            train_infos = self.train()
            total_num_steps = (episode + 1) * self.episode_length * self.n_rollout_threads
            if episode % self.save_interval == 0 or episode == episodes - 1:
                self.save()
            if episode % self.log_interval == 0:
                end = time.time()
                print("\n Map {} Algo {} Exp {} updates {}/{} episodes, total num timesteps {}/{}, FPS {}.\n".format(
                    self.all_args.map_name, self.algorithm_name, self.experiment_name, episode, episodes, total_num_steps,
                    self.num_env_steps, int(total_num_steps / (end - start))))
                if self.env_name in ("StarCraft2", "SMACv2", "SMAC", "StarCraft2v2"):
                    battles_won, battles_game, incre_battles_won, incre_battles_game = [], [], [], []
                    for i, info in enumerate(infos):
                        if "battles_won" in info[0].keys():
                            battles_won.append(info[0]["battles_won"])
                            incre_battles_won.append(info[0]["battles_won"] - last_battles_won[i])
                        if "battles_game" in info[0].keys():
                            battles_game.append(info[0]["battles_game"])
                            incre_battles_game.append(info[0]["battles_game"] - last_battles_game[i])
                    incre_win_rate = np.sum(incre_battles_won) / np.sum(incre_battles_game) if np.sum(incre_battles_game) > 0 else 0.0
                    print("incre win rate is {}.".format(incre_win_rate))
                    self._log("incre_win_rate", incre_win_rate, total_num_steps)
                    last_battles_game, last_battles_won = battles_game, battles_won
                am = self.buffer.active_masks
                train_infos["dead_ratio"] = 1 - float(am.sum().item()) / reduce(lambda x, y: x * y, list(am.shape))
                self.log_train(train_infos, total_num_steps)
            if episode % self.eval_interval == 0 and self.use_eval:
                self.eval(total_num_steps)

    def _put0(self, dst, src):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32)).reshape(dst.shape))

    def warmup(self):
        """reference :103-113."""
        obs, share_obs, available_actions = self.envs.reset()
        if not self.use_centralized_V:
            share_obs = obs
        self._put0(self.buffer.share_obs[0], share_obs)
        self._put0(self.buffer.obs[0], obs)
        self._put0(self.buffer.available_actions[0], available_actions)

    @torch.no_grad()
    def collect(self, step):
        """reference :115-131.  Only the actions come down to the host (the env needs them); the rest stays on the device."""
        self.trainer.prep_rollout()
        b, N = self.buffer, self.n_rollout_threads
        value, action, logp, h_a, h_c = self.trainer.policy.get_actions(
            self._rows(b.share_obs[step]), self._rows(b.obs[step]), self._rows(b.rnn_states[step]),
            self._rows(b.rnn_states_critic[step]), self._rows(b.masks[step]), self._rows(b.available_actions[step]))
        split = lambda x: x.reshape(N, self.num_agents, *x.shape[1:])
        return split(value), _t2n(action).reshape(N, self.num_agents, -1), split(logp), split(h_a), split(h_c)

    def insert(self, data):
        """reference :133-151."""
        obs, share_obs, rewards, dones, infos, available_actions, values, actions, action_log_probs, rnn_states, rnn_states_critic = data
        dev = self.buffer.device
        bad = [[bool(info[agent_id]["bad_transition"]) for agent_id in range(self.num_agents)] for info in infos]
        dones_env, masks, active_masks, bad_masks = smac_masks(dones, bad, dev)
        rnn_states, rnn_states_critic = rnn_states.clone(), rnn_states_critic.clone()
        rnn_states[dones_env] = 0.0
        rnn_states_critic[dones_env] = 0.0
        if not self.use_centralized_V:
            share_obs = obs
        self.buffer.insert(share_obs, obs, rnn_states, rnn_states_critic, np.asarray(actions, dtype=np.float32), action_log_probs,
                           values, rewards, masks, bad_masks, active_masks, available_actions)

    def log_train(self, train_infos, total_num_steps):
        """reference :153-159."""
        train_infos["average_step_rewards"] = float(self.buffer.rewards.mean().item())
        for k, v in train_infos.items():
            self._log(k, float(v), total_num_steps)

    @torch.no_grad()
    def eval(self, total_num_steps):
        """reference :161-251: deterministic episodes on eval_envs until `eval_episodes` have finished; logs the mean episode
        reward and the win rate."""
        n, M = self.n_eval_rollout_threads, self.num_agents
        dev = self.buffer.device
        eval_battles_won, eval_episode = 0, 0
        eval_episode_rewards, one_episode_rewards = [], []
        eval_obs, eval_share_obs, eval_available_actions = self.eval_envs.reset()
        h = torch.zeros(n, M, self.recurrent_N, self.hidden_size, device=dev)
        masks = torch.ones(n, M, 1, device=dev)
        while True:
            self.trainer.prep_rollout()
            a, h_new = self.trainer.policy.act(np.concatenate(eval_obs), h.reshape(n * M, self.recurrent_N, self.hidden_size),
                                               masks.reshape(n * M, 1), np.concatenate(eval_available_actions), deterministic=True)
            eval_actions = _t2n(a).reshape(n, M, -1)
            h = h_new.reshape(n, M, self.recurrent_N, self.hidden_size).clone()
            eval_obs, eval_share_obs, eval_rewards, eval_dones, eval_infos, eval_available_actions = self.eval_envs.step(eval_actions)
            one_episode_rewards.append(eval_rewards)
            eval_dones_env = np.all(eval_dones, axis=1)
            de = torch.from_numpy(eval_dones_env).to(dev)
            h[de] = 0.0
            masks = torch.ones(n, M, 1, device=dev)
            masks[de] = 0.0
            for eval_i in range(n):
                if eval_dones_env[eval_i]:
                    eval_episode += 1
                    eval_episode_rewards.append(np.sum(one_episode_rewards, axis=0))
                    one_episode_rewards = []
                    if eval_infos[eval_i][0]["won"]:
                        eval_battles_won += 1
            if eval_episode >= self.all_args.eval_episodes:
                # (two threads finishing in the same step leave a 0-d entry next to [n, M, 1] ones -- the reference's list is
                #  ragged in exactly the same way, which NumPy >= 1.24 refuses to stack: flatten before averaging)
                flat = np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in eval_episode_rewards])
                self.log_env({"eval_average_episode_rewards": flat}, total_num_steps)
                eval_win_rate = eval_battles_won / eval_episode
                print("eval win rate is {}.".format(eval_win_rate))
                self._log("eval_win_rate", eval_win_rate, total_num_steps)
                break
