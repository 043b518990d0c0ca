"""world_size-2 gloo test (CPU): sharding rollout threads over ranks with GLOBAL normalisers and a gradient
all-reduce reproduces the single-process gradient -- the host-side logic of the multi-GPU path."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mappo_oracle as O


def _full_problem():
    cfg = O.PathConfig(episode_length=6, n_rollout_threads=4, num_agents=2, obs_dim=7, share_obs_dim=14, act_dims=(4,),
                       ppo_epoch=1, use_valuenorm=False, use_max_grad_norm=False)
    rng = np.random.RandomState(0)
    store = O.RolloutStore(cfg)
    for nm in ("share_obs", "obs", "value_preds", "returns", "action_log_probs", "rewards"):
        a = getattr(store, nm)
        a[:] = rng.randn(*a.shape) * 0.3
    store.actions[:] = rng.randint(0, 4, size=store.actions.shape)
    store.active_masks[:] = (rng.rand(*store.active_masks.shape) > 0.3)
    adv = rng.randn(6, 4, 2, 1).astype(np.float32)
    pa, pc = O.init_params(cfg, False, seed=1), O.init_params(cfg, True, seed=2)
    return cfg, store, adv, pa, pc


def _shard(cfg, store, adv, lo, hi):
    c = O.PathConfig(**{**cfg.to_dict(), "n_rollout_threads": hi - lo, "act_dims": tuple(cfg.act_dims)})
    s = O.RolloutStore(c)
    for nm in ("share_obs", "obs", "rnn_states", "rnn_states_critic", "value_preds", "returns", "actions",
               "action_log_probs", "rewards", "masks", "active_masks", "available_actions"):
        getattr(s, nm)[:] = getattr(store, nm)[:, lo:hi]
    return c, s, adv[:, lo:hi]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "on-policy_b200"))
    from mappo_b200 import dist as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg, store, adv, pa, pc = _full_problem()
    lo, hi = D.shard_of_threads(cfg.n_rollout_threads, world, rank)
    c, s, a = _shard(cfg, store, adv, lo, hi)
    B_loc = c.episode_length * c.n_rollout_threads * c.num_agents
    # statistics collective: [sum active, sum R, sum R^2, rows] + advantage stats
    act = s.active_masks[:-1]
    local = torch.tensor([[act.sum(), s.returns[:-1].sum(), (s.returns[:-1] ** 2).sum(), B_loc]], dtype=torch.float64)
    flat = D.pack_stats(local, torch.zeros(3, dtype=torch.float64))
    D.allreduce_sum_(flat)
    per_update, _ = D.unpack_stats(flat, 1)
    # local masked-mean gradients re-weighted to their share of the global mean, then summed over ranks
    learner = O.Learner(c, pa, pc)
    sample = next(O.minibatches(s, a, np.arange(B_loc)))
    g = learner.ppo_update(sample, keep_grads=True)
    w = D.loss_weight(float(act.sum()), B_loc, float(per_update[0, 0]), int(per_update[0, 3]), True)
    grads = torch.cat([v.reshape(-1) for v in g["actor_grads"].values()] +
                      [v.reshape(-1) for v in g["critic_grads"].values()]) * w
    D.allreduce_sum_(grads)
    if rank == 0:
        np.save(out, np.concatenate([grads.numpy(), per_update.numpy().reshape(-1)]))
    dist.destroy_process_group()


def test_sharded_gradients_equal_single_process(tmp_path):
    out = str(tmp_path / "g.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    cfg, store, adv, pa, pc = _full_problem()
    B = cfg.episode_length * cfg.n_rollout_threads * cfg.num_agents
    learner = O.Learner(cfg, pa, pc)
    g = learner.ppo_update(next(O.minibatches(store, adv, np.arange(B))), keep_grads=True)
    want = torch.cat([v.reshape(-1) for v in g["actor_grads"].values()] +
                     [v.reshape(-1) for v in g["critic_grads"].values()]).numpy()
    np.testing.assert_allclose(got[:-4], want, rtol=2e-4, atol=1e-7)
    act = store.active_masks[:-1]
    np.testing.assert_allclose(got[-4:], [act.sum(), store.returns[:-1].sum(), (store.returns[:-1] ** 2).sum(), B],
                               rtol=1e-6)


def test_shard_of_threads():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "on-policy_b200"))
    from mappo_b200 import dist as D
    assert D.shard_of_threads(128, 8, 3) == (48, 64)
    try:
        D.shard_of_threads(10, 4, 0)
        assert False
    except ValueError:
        pass
