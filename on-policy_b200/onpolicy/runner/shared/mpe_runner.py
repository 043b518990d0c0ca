"""onpolicy.runner.shared.mpe_runner.MPERunner on the B200 engine (reference: runner/shared/mpe_runner.py:11-277).

The environment stays a host (CPU) vec-env: per step the only host<->device traffic is the env's obs/reward/done
going up and the sampled actions coming down; everything else stays in HBM.
"""
import time

import numpy as np
import torch

from onpolicy.runner.shared.base_runner import Runner, _t2n


class MPERunner(Runner):
    def __init__(self, config):
        super(MPERunner, self).__init__(config)

    def run(self):
        """reference :16-79."""
        self.warmup()
        start = time.time()
        episodes = int(self.num_env_steps) // self.episode_length // self.n_rollout_threads
        for episode in range(episodes):
            if self.use_linear_lr_decay:
                self.trainer.policy.lr_decay(episode, episodes)
            for step in range(self.episode_length):
                values, actions, action_log_probs, rnn_states, rnn_states_critic, actions_env = self.collect(step)
                obs, rewards, dones, infos = self.envs.step(actions_env)
                self.insert((obs, rewards, dones, infos, values, actions, action_log_probs, rnn_states,
                             rnn_states_critic))
            self.compute()
            train_infos = self.train()
            total_num_steps = (episode + 1) * self.episode_length * self.n_rollout_threads
            if episode % self.save_interval == 0 or episode == episodes - 1:
                self.save()
            if episode % self.log_interval == 0:
                fps = int(total_num_steps / (time.time() - start))
                print("\n Scenario {} Algo {} Exp {} updates {}/{} episodes, total num timesteps {}/{}, FPS {}.\n".format(
                    self.all_args.scenario_name, self.algorithm_name, self.experiment_name, episode, episodes,
                    total_num_steps, self.num_env_steps, fps))
                env_infos = {}
                if self.env_name == "MPE":
                    for agent_id in range(self.num_agents):
                        env_infos["agent%i/individual_rewards" % agent_id] = [
                            info[agent_id]["individual_reward"] for info in infos
                            if "individual_reward" in info[agent_id].keys()]
                train_infos["average_episode_rewards"] = float(self.buffer.rewards.mean().item()) * self.episode_length
                print("average episode rewards is {}".format(train_infos["average_episode_rewards"]))
                self.log_train(train_infos, total_num_steps)
                self.log_env(env_infos, total_num_steps)
            if episode % self.eval_interval == 0 and self.use_eval:
                self.eval(total_num_steps)

    def _share(self, obs):
        """Centralised observation = all agents' obs concatenated, repeated per agent (reference :84-88, :133-135)."""
        if not self.use_centralized_V:
            return obs
        flat = obs.reshape(obs.shape[0], -1)
        return np.repeat(flat[:, None, :], self.num_agents, axis=1)

    def warmup(self):
        """reference :81-93."""
        obs = self.envs.reset()
        self.buffer.share_obs[0].copy_(torch.from_numpy(np.ascontiguousarray(self._share(obs), dtype=np.float32)))
        self.buffer.obs[0].copy_(torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float32)))

    def _one_hot_actions(self, actions):
        """Env-facing one-hot encoding (reference :110-121)."""
        space = self.envs.action_space[0]
        kind = space.__class__.__name__
        if kind == "MultiDiscrete":
            parts = [np.eye(space.high[i] + 1)[actions[:, :, i]] for i in range(space.shape)]
            return np.concatenate(parts, axis=2)
        if kind == "Discrete":
            return np.squeeze(np.eye(space.n)[actions], 2)
        raise NotImplementedError

    @torch.no_grad()
    def collect(self, step):
        """reference :95-123.  values / log-probs / rnn states stay on the device."""
        self.trainer.prep_rollout()
        b, N = self.buffer, self.n_rollout_threads
        avail = None
        value, action, logp, h_a, h_c = self.trainer.policy.get_actions(
            self._rows(b.share_obs[step]), self._rows(b.obs[step]), self._rows(b.rnn_states[step]),
            self._rows(b.rnn_states_critic[step]), self._rows(b.masks[step]), avail)
        split = lambda x: x.reshape(N, self.num_agents, *x.shape[1:])
        actions = _t2n(action).reshape(N, self.num_agents, -1)          # the one device->host copy of the step
        return split(value), actions, split(logp), split(h_a), split(h_c), self._one_hot_actions(actions)

    def insert(self, data):
        """reference :125-139."""
        obs, rewards, dones, infos, values, actions, action_log_probs, rnn_states, rnn_states_critic = data
        dev = self.buffer.device
        done = torch.from_numpy(np.asarray(dones, dtype=bool)).to(dev)
        rnn_states = rnn_states.clone()
        rnn_states_critic = rnn_states_critic.clone()
        rnn_states[done] = 0.0
        rnn_states_critic[done] = 0.0
        masks = torch.ones(self.n_rollout_threads, self.num_agents, 1, dtype=torch.float32, device=dev)
        masks[done] = 0.0
        self.buffer.insert(self._share(obs), obs, rnn_states, rnn_states_critic, actions.astype(np.float32),
                           action_log_probs, values, rewards, masks)

    @torch.no_grad()
    def eval(self, total_num_steps):
        """reference :141-183: deterministic rollouts on eval_envs, logs eval_average_episode_rewards."""
        n = self.n_eval_rollout_threads
        dev = self.buffer.device
        eval_obs = self.eval_envs.reset()
        h = torch.zeros(n * self.num_agents, self.recurrent_N, self.hidden_size, device=dev)
        masks = torch.ones(n * self.num_agents, 1, device=dev)
        rewards_log = []
        self.trainer.prep_rollout()
        for _ in range(self.episode_length):
            action, h = self.trainer.policy.act(eval_obs.reshape(n * self.num_agents, -1), h, masks, deterministic=True)
            actions = _t2n(action).reshape(n, self.num_agents, -1)
            eval_obs, eval_rewards, eval_dones, _ = self.eval_envs.step(self._one_hot_actions(actions))
            rewards_log.append(eval_rewards)
            done = torch.from_numpy(np.asarray(eval_dones, dtype=bool)).to(dev).reshape(-1)
            h = h.clone()
            h[done] = 0.0
            masks = torch.ones(n * self.num_agents, 1, device=dev)
            masks[done] = 0.0
        avg = float(np.mean(np.sum(np.array(rewards_log), axis=0)))
        print("eval average episode rewards of agent: " + str(avg))
        self.log_env({"eval_average_episode_rewards": [avg]}, total_num_steps)

    @torch.no_grad()
    def render(self):
        """reference :185-245: `render_episodes` deterministic episodes on self.envs; frames from envs.render('rgb_array') go to
        <run_dir>/gifs/render.gif when --save_gifs (imageio, as in the reference; without it the frames are kept as render.npz),
        otherwise envs.render('human') is called every step.  Prints the reference's average-episode-reward line."""
        envs, a = self.envs, self.all_args
        n, dev = self.n_rollout_threads, self.buffer.device
        frames = []

        def show():
            if a.save_gifs:
                frames.append(envs.render("rgb_array")[0][0])
            else:
                envs.render("human")

        for _ in range(a.render_episodes):
            obs = envs.reset()
            show()
            h = torch.zeros(n * self.num_agents, self.recurrent_N, self.hidden_size, device=dev)
            masks = torch.ones(n * self.num_agents, 1, device=dev)
            episode_rewards = []
            for _step in range(self.episode_length):
                t0 = time.time()
                self.trainer.prep_rollout()
                action, h = self.trainer.policy.act(np.asarray(obs).reshape(n * self.num_agents, -1), h, masks, deterministic=True)
                actions = _t2n(action).reshape(n, self.num_agents, -1)
                obs, rewards, dones, _ = envs.step(self._one_hot_actions(actions))
                episode_rewards.append(rewards)
                done = torch.from_numpy(np.asarray(dones, dtype=bool)).to(dev).reshape(-1)
                h = h.clone()
                h[done] = 0.0
                masks = torch.ones(n * self.num_agents, 1, device=dev)
                masks[done] = 0.0
                show()
                if a.save_gifs:
                    elapsed = time.time() - t0
                    if elapsed < a.ifi:
                        time.sleep(a.ifi - elapsed)
            print("average episode rewards is: " + str(np.mean(np.sum(np.array(episode_rewards), axis=0))))
        if a.save_gifs:
            self._save_frames(frames, a.ifi)

    def _save_frames(self, frames, ifi):
        import os
        gif_dir = getattr(self, "gif_dir", None) or str(self.run_dir / "gifs")
        os.makedirs(gif_dir, exist_ok=True)
        try:
            import imageio
            imageio.mimsave(gif_dir + "/render.gif", frames, duration=ifi)
        except ImportError:
            np.savez_compressed(gif_dir + "/render.npz", frames=np.asarray(frames))

