"""onpolicy.runner.separated.mpe_runner.MPERunner on the B200 engine (reference: runner/separated/mpe_runner.py)."""
import time

import numpy as np
import torch

from onpolicy.runner.separated.base_runner import Runner, _t2n


class MPERunner(Runner):
    def __init__(self, config):
        super(MPERunner, self).__init__(config)

    def run(self):
        self.warmup()
        start = time.time()
        episodes = int(self.num_env_steps) // self.episode_length // self.n_rollout_threads
        for episode in range(episodes):
            if self.use_linear_lr_decay:
                for agent_id in range(self.num_agents):
                    self.trainer[agent_id].policy.lr_decay(episode, episodes)
            for step in range(self.episode_length):
                values, actions, logps, h_a, h_c, actions_env = self.collect(step)
                obs, rewards, dones, infos = self.envs.step(actions_env)
                self.insert((obs, rewards, dones, infos, values, actions, logps, h_a, h_c))
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
                if self.env_name == "MPE":
                    for agent_id in range(self.num_agents):
                        train_infos[agent_id]["average_episode_rewards"] = \
                            float(self.buffer[agent_id].rewards.mean().item()) * self.episode_length
                self.log_train(train_infos, total_num_steps)
            if episode % self.eval_interval == 0 and self.use_eval:
                self.eval(total_num_steps)

    def _share(self, obs):
        return np.stack([np.concatenate(list(o)) for o in obs])          # [N, sum of agents' obs]

    def warmup(self):
        obs = self.envs.reset()
        share = self._share(obs)
        for agent_id in range(self.num_agents):
            b = self.buffer[agent_id]
            mine = np.stack([np.asarray(o[agent_id], dtype=np.float32) for o in obs])
            b.share_obs[0].copy_(torch.from_numpy(np.ascontiguousarray(share if self.use_centralized_V else mine,
                                                                       dtype=np.float32)))
            b.obs[0].copy_(torch.from_numpy(mine))

    @torch.no_grad()
    def collect(self, step):
        values, actions, logps, h_as, h_cs, env_parts = [], [], [], [], [], []
        for agent_id in range(self.num_agents):
            self.trainer[agent_id].prep_rollout()
            b = self.buffer[agent_id]
            v, a, lp, h_a, h_c = self.trainer[agent_id].policy.get_actions(
                b.share_obs[step], b.obs[step], b.rnn_states[step], b.rnn_states_critic[step], b.masks[step])
            a_np = _t2n(a)
            space = self.envs.action_space[agent_id]
            if space.__class__.__name__ == "MultiDiscrete":
                env_parts.append(np.concatenate([np.eye(space.high[i] + 1)[a_np[:, i]] for i in range(space.shape)], 1))
            elif space.__class__.__name__ == "Discrete":
                env_parts.append(np.squeeze(np.eye(space.n)[a_np], 1))
            else:
                raise NotImplementedError
            values.append(v); actions.append(a_np); logps.append(lp); h_as.append(h_a); h_cs.append(h_c)
        actions_env = [[env_parts[m][n] for m in range(self.num_agents)] for n in range(self.n_rollout_threads)]
        return values, actions, logps, h_as, h_cs, actions_env

    def insert(self, data):
        obs, rewards, dones, infos, values, actions, logps, h_as, h_cs = data
        dones = np.asarray(dones, dtype=bool)
        share = self._share(obs)
        for agent_id in range(self.num_agents):
            b = self.buffer[agent_id]
            d = torch.from_numpy(dones[:, agent_id]).to(b.device)
            h_a, h_c = h_as[agent_id].clone(), h_cs[agent_id].clone()
            h_a[d] = 0.0
            h_c[d] = 0.0
            masks = torch.ones(self.n_rollout_threads, 1, dtype=torch.float32, device=b.device)
            masks[d] = 0.0
            mine = np.stack([np.asarray(o[agent_id], dtype=np.float32) for o in obs])
            b.insert(share if self.use_centralized_V else mine, mine, h_a, h_c, actions[agent_id].astype(np.float32),
                     logps[agent_id], values[agent_id], np.asarray(rewards)[:, agent_id].reshape(-1, 1), masks)

    @torch.no_grad()
    def eval(self, total_num_steps):
        """reference :178-239: deterministic rollouts on eval_envs with every agent's own policy; logs each agent's
        eval_average_episode_rewards (per-agent reward column summed over the episode, averaged over the eval threads)."""
        n, M = self.n_eval_rollout_threads, self.num_agents
        dev = self.buffer[0].device
        eval_obs = self.eval_envs.reset()
        h = [torch.zeros(n, self.recurrent_N, self.hidden_size, device=dev) for _ in range(M)]
        masks = torch.ones(n, M, 1, device=dev)
        rewards_log = []
        for _ in range(self.episode_length):
            parts = []
            for agent_id in range(M):
                self.trainer[agent_id].prep_rollout()
                mine = np.stack([np.asarray(o[agent_id], dtype=np.float32) for o in eval_obs])
                a, h[agent_id] = self.trainer[agent_id].policy.act(mine, h[agent_id], masks[:, agent_id], deterministic=True)
                a_np = _t2n(a)
                space = self.eval_envs.action_space[agent_id]
                if space.__class__.__name__ == "MultiDiscrete":
                    parts.append(np.concatenate([np.eye(space.high[i] + 1)[a_np[:, i]] for i in range(space.shape)], 1))
                elif space.__class__.__name__ == "Discrete":
                    parts.append(np.squeeze(np.eye(space.n)[a_np], 1))
                else:
                    raise NotImplementedError
            actions_env = [[parts[m][i] for m in range(M)] for i in range(n)]
            eval_obs, eval_rewards, eval_dones, _ = self.eval_envs.step(actions_env)
            rewards_log.append(eval_rewards)
            done = torch.from_numpy(np.asarray(eval_dones, dtype=bool)).to(dev)
            masks = torch.ones(n, M, 1, device=dev)
            masks[done] = 0.0
            for agent_id in range(M):
                h[agent_id] = h[agent_id].clone()
                h[agent_id][done[:, agent_id]] = 0.0
        rew = np.array(rewards_log)                                   # [T, n, M, 1]
        infos = []
        for agent_id in range(M):
            avg = float(np.mean(np.sum(rew[:, :, agent_id], axis=0)))
            infos.append({"eval_average_episode_rewards": avg})
            print("eval average episode rewards of agent%i: " % agent_id + str(avg))
        self.log_train(infos, total_num_steps)

    @torch.no_grad()
    def render(self):
        """reference :241-313: like the shared runner's render, every agent acting with its own policy."""
        import time
        envs, a = self.envs, self.all_args
        n, M = self.n_rollout_threads, self.num_agents
        dev = self.buffer[0].device
        frames = []

        def show():
            if a.save_gifs:
                frames.append(envs.render("rgb_array")[0][0])
            else:
                envs.render("human")

        for _ in range(a.render_episodes):
            obs = envs.reset()
            show()
            h = [torch.zeros(n, self.recurrent_N, self.hidden_size, device=dev) for _ in range(M)]
            masks = torch.ones(n, M, 1, device=dev)
            episode_rewards = []
            for _step in range(self.episode_length):
                t0 = time.time()
                parts = []
                for agent_id in range(M):
                    self.trainer[agent_id].prep_rollout()
                    mine = np.stack([np.asarray(o[agent_id], dtype=np.float32) for o in obs])
                    act, h[agent_id] = self.trainer[agent_id].policy.act(mine, h[agent_id], masks[:, agent_id], deterministic=True)
                    a_np = _t2n(act)
                    space = envs.action_space[agent_id]
                    if space.__class__.__name__ == "MultiDiscrete":
                        parts.append(np.concatenate([np.eye(space.high[i] + 1)[a_np[:, i]] for i in range(space.shape)], 1))
                    elif space.__class__.__name__ == "Discrete":
                        parts.append(np.squeeze(np.eye(space.n)[a_np], 1))
                    else:
                        raise NotImplementedError
                obs, rewards, dones, _ = envs.step([[parts[m][i] for m in range(M)] for i in range(n)])
                episode_rewards.append(rewards)
                done = torch.from_numpy(np.asarray(dones, dtype=bool)).to(dev)
                masks = torch.ones(n, M, 1, device=dev)
                masks[done] = 0.0
                for agent_id in range(M):
                    h[agent_id] = h[agent_id].clone()
                    h[agent_id][done[:, agent_id]] = 0.0
                show()
                if a.save_gifs:
                    elapsed = time.time() - t0
                    if elapsed < a.ifi:
                        time.sleep(a.ifi - elapsed)
            rew = np.array(episode_rewards)
            for agent_id in range(M):
                print("eval average episode rewards of agent%i: " % agent_id + str(np.mean(np.sum(rew[:, :, agent_id], axis=0))))
        if a.save_gifs:
            import os
            gif_dir = getattr(self, "gif_dir", None) or str(self.run_dir / "gifs")
            os.makedirs(gif_dir, exist_ok=True)
            try:
                import imageio
                imageio.mimsave(gif_dir + "/render.gif", frames, duration=a.ifi)
            except ImportError:
                np.savez_compressed(gif_dir + "/render.npz", frames=np.asarray(frames))

