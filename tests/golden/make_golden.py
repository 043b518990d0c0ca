"""Generate the golden fixtures in this directory by running the UNMODIFIED reference on CPU.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py          (needs /root/reference)

The reference (marlbenchmark/on-policy) has no tests or golden vectors for the hot path
(SURVEY.md section 4), so parity is pinned on outputs of the reference itself: for each case this script
builds the reference's R_MAPPOPolicy / R_MAPPO / SharedReplayBuffer with duck-typed spaces,
drives `iters` iterations of the loop in runner/shared/mpe_runner.py:26-40 +
base_runner.py:120-141 on synthetic env data, and stores inputs (initial weights, env feed, the
Exp(1) sampling noise and the minibatch permutations the reference's CPU RNG produced) together with
every output (buffer contents, advantages, first-update gradients, train_info, final weights,
ValueNorm state).  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MAPPO_REFERENCE", "/root/reference")

from oracle import mappo_oracle as O  # noqa: E402


class _Space:
    pass


def make_spaces(cfg: O.PathConfig):
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


def ref_args(cfg: O.PathConfig, algo: str) -> Namespace:
    sys.path.insert(0, REF)
    from onpolicy.config import get_config
    a = get_config().parse_known_args([])[0]
    for k, v in cfg.to_dict().items():
        if hasattr(a, k):
            setattr(a, k, v)
    a.algorithm_name = algo
    a.use_popart = False
    return a


CASES = {
    # c1 of BASELINE.json exactly (train_mpe_spread.sh: mappo, Tanh, ppo_epoch 10, lr 7e-4)
    "c1_mlp_discrete": dict(
        cfg=O.PathConfig(episode_length=25, n_rollout_threads=8, num_agents=3, obs_dim=18, share_obs_dim=54,
                         act_dims=(5,), use_ReLU=False, ppo_epoch=10, lr=7e-4, critic_lr=7e-4),
        algo="mappo", feed="mpe", iters=2, seed=1),
    # c3-shaped (train_mpe_reference.sh: rmappo, GRU, L=10, MultiDiscrete [5,10]) at N=4; T%L != 0 straddle
    "c3_gru_multidiscrete": dict(
        cfg=O.PathConfig(episode_length=25, n_rollout_threads=4, num_agents=2, obs_dim=21, share_obs_dim=42,
                         act_dims=(5, 10), multi_discrete=True, use_recurrent_policy=True, data_chunk_length=10,
                         ppo_epoch=3, lr=7e-4, critic_lr=7e-4),
        algo="rmappo", feed="mpe", iters=2, seed=2),
    # c4-shaped (train_smac_3m.sh: rmappo, avail + active masks, unmasked value loss) with 2 minibatches
    "c4_gru_smac": dict(
        cfg=O.PathConfig(episode_length=20, n_rollout_threads=4, num_agents=3, obs_dim=30, share_obs_dim=48,
                         act_dims=(9,), use_recurrent_policy=True, data_chunk_length=10, ppo_epoch=2,
                         num_mini_batch=2, use_value_active_masks=False),
        algo="rmappo", feed="smac", iters=2, seed=3),
    # c5-shaped MLP (hanabi: layer_N 2, avail masks, entropy 0.015, critic_lr 1e-3) small, 2 minibatches,
    # with the non-default loss switches flipped
    "c5_mlp_switches": dict(
        cfg=O.PathConfig(episode_length=12, n_rollout_threads=6, num_agents=2, obs_dim=40, share_obs_dim=50,
                         act_dims=(20,), layer_N=2, ppo_epoch=2, num_mini_batch=2, entropy_coef=0.015,
                         lr=7e-4, critic_lr=1e-3, use_clipped_value_loss=False, use_huber_loss=False,
                         use_policy_active_masks=False, use_max_grad_norm=False),
        algo="mappo", feed="smac", iters=1, seed=4),
    # c5 nets at full width (train_hanabi_forward.sh: hidden 512, layer_N 2, obs 658+2, share 783+2, 20 actions, ReLU,
    # entropy 0.015, critic_lr 1e-3) on a small batch; big tensors are stored subsampled (see `put`), the initial
    # weights as seed + checksums (the engine's initialiser is seed-identical to the reference's)
    "c5_h512_hanabi": dict(
        cfg=O.PathConfig(episode_length=6, n_rollout_threads=12, num_agents=2, obs_dim=660, share_obs_dim=785,
                         act_dims=(20,), layer_N=2, hidden_size=512, ppo_epoch=2, num_mini_batch=1, entropy_coef=0.015,
                         lr=7e-4, critic_lr=1e-3),
        algo="mappo", feed="smac", iters=1, seed=6, compact=True),
    # c2 of BASELINE.json exactly (the benchmark size, N = 128): one iteration
    "c2_mlp_n128": dict(
        cfg=O.PathConfig(episode_length=25, n_rollout_threads=128, num_agents=3, obs_dim=18, share_obs_dim=54,
                         act_dims=(5,), use_ReLU=False, ppo_epoch=10, lr=7e-4, critic_lr=7e-4),
        algo="mappo", feed="mpe", iters=1, seed=7, compact=True),
    # non-default return modes: proper time limits + GAE, naive recurrent generator
    "naive_rnn_ptl": dict(
        cfg=O.PathConfig(episode_length=8, n_rollout_threads=4, num_agents=2, obs_dim=10, share_obs_dim=20,
                         act_dims=(4,), use_naive_recurrent_policy=True, ppo_epoch=2, num_mini_batch=2,
                         use_proper_time_limits=True),
        algo="mappo", feed="smac", iters=1, seed=5),
}


def put(out, key, arr, compact):
    """Store a tensor; in compact cases a matrix with more than 65536 elements keeps every 16th row + its Frobenius norm."""
    arr = np.asarray(arr)
    if compact and arr.ndim == 2 and arr.size > 65536:
        idx = np.arange(0, arr.shape[0], 16, dtype=np.int32)
        out[key + "@rows"] = idx
        out[key + "@norm"] = np.array(np.sqrt((arr.astype(np.float64) ** 2).sum()))
        out[key] = arr[idx].copy()
    else:
        out[key] = arr.copy()


def run_case(name, spec):
    cfg, algo, seed = spec["cfg"], spec["algo"], spec["seed"]
    compact = bool(spec.get("compact", False))
    sys.path.insert(0, REF)
    from onpolicy.utils.shared_buffer import SharedReplayBuffer
    from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO
    from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy

    torch.set_num_threads(1)
    torch.manual_seed(seed)
    np.random.seed(seed)
    args = ref_args(cfg, algo)
    # train_mpe.py:68-80 derives the recurrence flags from algorithm_name; mirror the effective values
    args.use_recurrent_policy = cfg.use_recurrent_policy
    args.use_naive_recurrent_policy = cfg.use_naive_recurrent_policy
    obs_s, share_s, act_s = make_spaces(cfg)
    dev = torch.device("cpu")
    policy = R_MAPPOPolicy(args, obs_s, share_s, act_s, device=dev)
    trainer = R_MAPPO(args, policy, device=dev)
    buf = SharedReplayBuffer(args, cfg.num_agents, obs_s, share_s, act_s)

    out = {"cfg_json": np.array(repr(cfg.to_dict()))}
    if compact:
        out["init_seed"] = np.array(seed)
        for net, sd in (("actor", policy.actor.state_dict()), ("critic", policy.critic.state_dict())):
            for k, v in sd.items():
                a = v.numpy().astype(np.float64)
                out[f"init_checksum/{net}/{k}"] = np.array([a.sum(), (a * a).sum()])
    else:
        for k, v in policy.actor.state_dict().items():
            out[f"init/actor/{k}"] = v.numpy().copy()
        for k, v in policy.critic.state_dict().items():
            out[f"init/critic/{k}"] = v.numpy().copy()

    T, N, M = cfg.episode_length, cfg.n_rollout_threads, cfg.num_agents
    E, sumA = N * M, sum(cfg.act_dims)
    first_grads = {}
    orig_update = trainer.ppo_update

    def recording_update(sample, update_actor=True):
        r = orig_update(sample, update_actor)
        if not first_grads:
            for k, p in policy.actor.named_parameters():
                first_grads[f"actor/{k}"] = p.grad.numpy().copy()
            for k, p in policy.critic.named_parameters():
                first_grads[f"critic/{k}"] = p.grad.numpy().copy()
            first_grads["norms"] = np.array([float(r[4]), float(r[1])], dtype=np.float64)
            first_grads["losses"] = np.array([float(r[0]), float(r[2]), float(r[3]), float(r[5].mean())])
        return r

    trainer.ppo_update = recording_update

    for it in range(spec["iters"]):
        feed = O.make_feed(cfg, seed=100 * seed + it, kind=spec["feed"])
        if it > 0 and spec["feed"] == "mpe":
            pass
        # what the CPU generator will hand out, in the order the reference consumes it (App. B-8):
        # T x one exponential_ per head [E, A_k], then per epoch one randperm
        rng_state = torch.get_rng_state()
        noise = np.zeros((T, E, sumA), np.float32)
        for t in range(T):
            off = 0
            for A in cfg.act_dims:
                noise[t, :, off:off + A] = torch.empty(E, A).exponential_(1).numpy()
                off += A
        perms = np.stack([torch.randperm(O.perm_length(cfg)).numpy() for _ in range(cfg.ppo_epoch)])
        torch.set_rng_state(rng_state)

        if it == 0:   # warmup
            buf.obs[0], buf.share_obs[0] = feed.obs[0].copy(), feed.share_obs[0].copy()
            if feed.available_actions is not None:
                buf.available_actions[0] = feed.available_actions[0].copy()
        else:
            # slot 0 holds last iteration's slot T (after_update); the feed continues from there
            feed.obs[0], feed.share_obs[0] = buf.obs[0].copy(), buf.share_obs[0].copy()
            if feed.available_actions is not None:
                feed.available_actions[0] = buf.available_actions[0].copy()
        trainer.prep_rollout()
        with torch.no_grad():
            for t in range(T):
                avail = None if feed.available_actions is None else np.concatenate(buf.available_actions[t])
                v, a, lp, ha, hc = policy.get_actions(np.concatenate(buf.share_obs[t]), np.concatenate(buf.obs[t]),
                                                      np.concatenate(buf.rnn_states[t]),
                                                      np.concatenate(buf.rnn_states_critic[t]),
                                                      np.concatenate(buf.masks[t]), avail)
                sp = lambda x: np.array(np.split(x.detach().cpu().numpy(), N))
                v, a, lp, ha, hc = sp(v), sp(a), sp(lp), sp(ha), sp(hc)
                d = feed.dones[t]
                ha[d] = 0.0
                hc[d] = 0.0
                masks = np.ones((N, M, 1), np.float32)
                masks[d] = 0.0
                buf.insert(feed.share_obs[t + 1], feed.obs[t + 1], ha, hc, a, lp, v, feed.rewards[t], masks,
                           active_masks=None if feed.active_masks is None else feed.active_masks[t],
                           available_actions=None if feed.available_actions is None
                           else feed.available_actions[t + 1])
            nv = policy.get_values(np.concatenate(buf.share_obs[-1]), np.concatenate(buf.rnn_states_critic[-1]),
                                   np.concatenate(buf.masks[-1]))
            buf.compute_returns(np.array(np.split(nv.numpy(), N)), trainer.value_normalizer)
        # advantages as R_MAPPO.train computes them (r_mappo.py:179-187), for the record
        vn = trainer.value_normalizer
        adv = buf.returns[:-1] - vn.denormalize(buf.value_preds[:-1])
        c = adv.copy()
        c[buf.active_masks[:-1] == 0.0] = np.nan
        adv = (adv - np.nanmean(c)) / (np.nanstd(c) + 1e-5)
        pre = f"it{it}/"
        for nm in ("obs", "share_obs", "rewards"):
            out[pre + "feed/" + nm] = getattr(feed, nm).copy()
        out[pre + "feed/dones"] = feed.dones.copy()
        if feed.active_masks is not None:
            out[pre + "feed/active_masks"] = feed.active_masks.copy()
        if feed.available_actions is not None:
            out[pre + "feed/available_actions"] = feed.available_actions.copy()
        out[pre + "noise"], out[pre + "perms"] = noise, perms
        for nm in ("actions", "action_log_probs", "value_preds", "returns", "rnn_states", "rnn_states_critic",
                   "masks", "active_masks"):
            if compact and nm.startswith("rnn_states") and not cfg.recurrent:
                continue                                  # all zeros for MLP policies
            out[pre + "buf/" + nm] = getattr(buf, nm).copy()
        out[pre + "advantages"] = adv.astype(np.float32)

        trainer.prep_training()
        info = trainer.train(buf)
        buf.after_update()
        out[pre + "train_info"] = np.array([float(info[k]) for k in
                                            ("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm",
                                             "critic_grad_norm", "ratio")], dtype=np.float64)
        if it == 0:
            for k, v in first_grads.items():
                put(out, f"it0/first_update/{k}", v, compact)
        for k, v in policy.actor.state_dict().items():
            put(out, pre + f"actor/{k}", v.numpy(), compact)
        for k, v in policy.critic.state_dict().items():
            put(out, pre + f"critic/{k}", v.numpy(), compact)
        out[pre + "valuenorm"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(),
                                           vn.debiasing_term.item()], dtype=np.float32)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, train_info(last)={out[pre + 'train_info']}")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: fixtures can only be regenerated where the reference is mounted")
    sys.dont_write_bytecode = True
    for n, s in CASES.items():
        if len(sys.argv) > 1 and n not in sys.argv[1:]:
            continue
        run_case(n, s)
