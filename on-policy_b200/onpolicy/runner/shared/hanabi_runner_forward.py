"""onpolicy.runner.shared.hanabi_runner_forward.HanabiRunner on the B200 engine
(reference: runner/shared/hanabi_runner_forward.py:14-328).

Hanabi is turn based: in one "step" every player moves once, in turn, in the games that are still running, and the reward of
a move is only known after all the other players have moved.  The reference keeps one row of pending per-player data per
rollout thread (`turn_*`, [N, M, ...]) and writes it into the storage with `chooseinsert` once per step, shifting the rewards
by one slot at the start of the next episode (:52-66).  Here the `turn_*` rows are DEVICE tensors and every masked update of
`collect` (:138-214) is a device-side indexed assignment; per player move the host sees only what the environment needs
(the chosen games' actions) and what it returns (obs / share_obs / available actions / rewards / dones of N games).
Same attribute names, loop structure, RNG consumption (one get_actions per player per step on the chosen rows) and log keys
as the reference; `tests/golden/make_golden_hanabi.py` runs the unmodified reference runner on the same scripted environment.
"""
import time

import numpy as np
import torch

from onpolicy.runner.shared.base_runner import Runner, _t2n


class HanabiRunner(Runner):
    """Runner class to perform training, evaluation and data collection for Hanabi (reference :14-18)."""

    def __init__(self, config):
        super(HanabiRunner, self).__init__(config)
        self.true_total_num_steps = 0

    # -- helpers --------------------------------------------------------------------------------
    def _dev(self, a, dtype=torch.float32):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float32)), dtype=dtype, device=self.buffer.device)

    def _mask(self, m):
        return torch.as_tensor(np.asarray(m, dtype=bool), device=self.buffer.device)

    def run(self):
        """reference :21-126."""
        b, N, dev = self.buffer, self.n_rollout_threads, self.buffer.device
        z = lambda a: torch.zeros(N, *a.shape[2:], dtype=torch.float32, device=dev)
        self.turn_obs, self.turn_share_obs = z(b.obs), z(b.share_obs)
        self.turn_available_actions = z(b.available_actions)
        self.turn_values, self.turn_actions, self.turn_action_log_probs = z(b.value_preds), z(b.actions), z(b.action_log_probs)
        self.turn_rnn_states, self.turn_rnn_states_critic = z(b.rnn_states), z(b.rnn_states)
        self.turn_masks = torch.ones(N, *b.masks.shape[2:], dtype=torch.float32, device=dev)
        self.turn_active_masks, self.turn_bad_masks = torch.ones_like(self.turn_masks), torch.ones_like(self.turn_masks)
        self.turn_rewards = z(b.rewards)
        self.turn_rewards_since_last_action = torch.zeros_like(self.turn_rewards)

        self.warmup()
        start = time.time()
        episodes = int(self.num_env_steps) // self.episode_length // self.n_rollout_threads
        train_infos = {}
        for episode in range(episodes):
            if self.use_linear_lr_decay:
                self.trainer.policy.lr_decay(episode, episodes)
            self.scores = []
            for step in range(self.episode_length):
                self.reset_choose = np.zeros(N) == 1.0
                self.collect(step)
                if step == 0 and episode > 0:
                    # the data of the last index of the previous episode (:52-57), the one-slot reward shift (:59-63), then train
                    b.share_obs[-1].copy_(self.turn_share_obs)
                    b.obs[-1].copy_(self.turn_obs)
                    b.available_actions[-1].copy_(self.turn_available_actions)
                    b.active_masks[-1].copy_(self.turn_active_masks)
                    b.rewards[0:self.episode_length - 1] = b.rewards[1:].clone()
                    b.rewards[-1].copy_(self.turn_rewards)
                    b._adv_version = -1
                    self.compute()
                    train_infos = self.train()
                b.chooseinsert(self.turn_share_obs, self.turn_obs, self.turn_rnn_states, self.turn_rnn_states_critic,
                               self.turn_actions, self.turn_action_log_probs, self.turn_values, self.turn_rewards,
                               self.turn_masks, self.turn_bad_masks, self.turn_active_masks, self.turn_available_actions)
                obs, share_obs, available_actions = self.envs.reset(self.reset_choose)
                share_obs = share_obs if self.use_centralized_V else obs
                rc = self.reset_choose
                self.use_obs[rc] = obs[rc]
                self.use_share_obs[rc] = share_obs[rc]
                self.use_available_actions[rc] = available_actions[rc]

            total_num_steps = (episode + 1) * self.episode_length * self.n_rollout_threads
            if episode % self.save_interval == 0 or episode == episodes - 1:
                self.save()
            if episode % self.log_interval == 0 and episode > 0:
                end = time.time()
                print("\n Env {} Algo {} Exp {} updates {}/{} episodes, total num timesteps {}/{}, FPS {}.\n".format(
                    self.all_args.hanabi_name, self.algorithm_name, self.experiment_name, episode, episodes, total_num_steps,
                    self.num_env_steps, int(total_num_steps / (end - start))))
                if self.env_name == "Hanabi":
                    average_score = np.mean(self.scores) if len(self.scores) > 0 else 0.0
                    print("average score is {}.".format(average_score))
                    self._log("average_score", average_score, self.true_total_num_steps)
                train_infos["average_step_rewards"] = float(b.rewards.mean().item())
                self.log_train(train_infos, self.true_total_num_steps)
            if episode % self.eval_interval == 0 and self.use_eval:
                self.eval(self.true_total_num_steps)

    def warmup(self):
        """reference :128-139.  `use_*` = what the player to move sees; host arrays (they come from / index the host env)."""
        self.reset_choose = np.ones(self.n_rollout_threads) == 1.0
        obs, share_obs, available_actions = self.envs.reset(self.reset_choose)
        share_obs = share_obs if self.use_centralized_V else obs
        self.use_obs = np.array(obs, dtype=np.float32)
        self.use_share_obs = np.array(share_obs, dtype=np.float32)
        self.use_available_actions = np.array(available_actions, dtype=np.float32)

    @torch.no_grad()
    def collect(self, step):
        """reference :141-214."""
        N, M = self.n_rollout_threads, self.num_agents
        for current_agent_id in range(M):
            env_actions = np.ones((N, *self.buffer.actions.shape[3:]), dtype=np.float32) * (-1.0)
            choose = np.any(self.use_available_actions == 1, axis=1)
            if ~np.any(choose):
                self.reset_choose = np.ones(N) == 1.0
                break
            ch = self._mask(choose)
            self.trainer.prep_rollout()
            value, action, action_log_prob, rnn_state, rnn_state_critic = self.trainer.policy.get_actions(
                self.use_share_obs[choose], self.use_obs[choose], self.turn_rnn_states[ch, current_agent_id],
                self.turn_rnn_states_critic[ch, current_agent_id], self.turn_masks[ch, current_agent_id],
                self.use_available_actions[choose])
            self.turn_obs[ch, current_agent_id] = self._dev(self.use_obs[choose])
            self.turn_share_obs[ch, current_agent_id] = self._dev(self.use_share_obs[choose])
            self.turn_available_actions[ch, current_agent_id] = self._dev(self.use_available_actions[choose])
            self.turn_values[ch, current_agent_id] = value
            self.turn_actions[ch, current_agent_id] = action.to(torch.float32)
            env_actions[choose] = _t2n(action)                      # the one device -> host copy of the move
            self.turn_action_log_probs[ch, current_agent_id] = action_log_prob
            self.turn_rnn_states[ch, current_agent_id] = rnn_state
            self.turn_rnn_states_critic[ch, current_agent_id] = rnn_state_critic

            obs, share_obs, rewards, dones, infos, available_actions = self.envs.step(env_actions)
            self.true_total_num_steps += int((choose == True).sum())
            share_obs = share_obs if self.use_centralized_V else obs
            self.use_obs = np.array(obs, dtype=np.float32)
            self.use_share_obs = np.array(share_obs, dtype=np.float32)
            self.use_available_actions = np.array(available_actions, dtype=np.float32)

            # rearrange reward: what a player earned shows up when it moves again (the reward of step 0 is thrown away)
            self.turn_rewards[ch, current_agent_id] = self.turn_rewards_since_last_action[ch, current_agent_id].clone()
            self.turn_rewards_since_last_action[ch, current_agent_id] = 0.0
            self.turn_rewards_since_last_action[ch] += self._dev(np.asarray(rewards, dtype=np.float32)[choose])

            dones = np.asarray(dones, dtype=object)
            done_t, done_f = (dones == True), (dones == False)      # (None = the game was not stepped)
            dt, df = self._mask(done_t), self._mask(done_f)
            self.reset_choose[done_t] = True
            # all agents of a finished game
            self.use_available_actions[done_t] = 0.0
            self.turn_masks[dt] = 0.0
            self.turn_rnn_states[dt] = 0.0
            self.turn_rnn_states_critic[dt] = 0.0
            # the current agent
            self.turn_active_masks[dt, current_agent_id] = 1.0
            # the agents that did not get to move in this round
            left = current_agent_id + 1
            self.turn_active_masks[dt, left:] = 0.0
            self.turn_rewards[dt, left:] = self.turn_rewards_since_last_action[dt, left:]
            self.turn_rewards_since_last_action[dt, left:] = 0.0
            self.turn_values[dt, left:] = 0.0
            self.turn_obs[dt, left:] = 0.0
            self.turn_share_obs[dt, left:] = 0.0
            # games that go on
            self.turn_masks[df, current_agent_id] = 1.0
            self.turn_active_masks[df, current_agent_id] = 1.0
            for done, info in zip(dones, infos):
                if done:
                    if "score" in info.keys():
                        self.scores.append(info["score"])

    def train(self):
        """reference :216-220."""
        self.trainer.prep_training()
        train_infos = self.trainer.train(self.buffer)
        self.buffer.chooseafter_update()
        return train_infos

    def _play(self, eval_envs, n):
        """One batch of deterministic games on `eval_envs` (reference :224-268) -> list of final scores."""
        dev = self.buffer.device
        scores = []
        reset_choose = np.ones(n) == 1.0
        eval_obs, eval_share_obs, eval_available_actions = eval_envs.reset(reset_choose)
        eval_available_actions = np.array(eval_available_actions, dtype=np.float32)
        h = torch.zeros(n, *self.buffer.rnn_states.shape[2:], device=dev)
        masks = torch.ones(n, self.num_agents, 1, device=dev)
        finish = False
        while not finish:
            for agent_id in range(self.num_agents):
                eval_actions = np.ones((n, 1), dtype=np.float32) * (-1.0)
                choose = np.any(eval_available_actions == 1, axis=1)
                if ~np.any(choose):
                    finish = True
                    break
                ch = self._mask(choose)
                self.trainer.prep_rollout()
                a, hs = self.trainer.policy.act(np.asarray(eval_obs, dtype=np.float32)[choose], h[ch, agent_id], masks[ch, agent_id],
                                                eval_available_actions[choose], deterministic=True)
                eval_actions[choose] = _t2n(a)
                h[ch, agent_id] = hs
                eval_obs, eval_share_obs, eval_rewards, eval_dones, eval_infos, eval_available_actions = eval_envs.step(eval_actions)
                eval_available_actions = np.array(eval_available_actions, dtype=np.float32)
                eval_dones = np.asarray(eval_dones, dtype=object)
                eval_available_actions[eval_dones == True] = 0.0
                for d, info in zip(eval_dones, eval_infos):
                    if d:
                        if "score" in info.keys():
                            scores.append(info["score"])
        return scores

    @torch.no_grad()
    def eval(self, total_num_steps):
        """reference :222-276."""
        eval_average_score = np.mean(self._play(self.eval_envs, self.n_eval_rollout_threads))
        print("eval average score is {}.".format(eval_average_score))
        self._log("eval_average_score", eval_average_score, total_num_steps)

    @torch.no_grad()
    def eval_100k(self, eval_games=100000):
        """reference :279-328."""
        trials = int(eval_games / self.n_eval_rollout_threads)
        eval_scores = []
        for trial in range(trials):
            print("trail is {}".format(trial))
            eval_scores += self._play(self.eval_envs, self.n_eval_rollout_threads)
        eval_average_score = np.mean(eval_scores)
        print("eval average score is {}.".format(eval_average_score))
