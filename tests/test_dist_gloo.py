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


def _worker_mb(rank, world, port, out):
    """num_mini_batch = 2: every rank draws the reference's GLOBAL permutation (same seed), keeps the rows it owns."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "on-policy_b200"))
    from mappo_b200 import dist as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg, store, adv, pa, pc = _full_problem()
    T, N, M = cfg.episode_length, cfg.n_rollout_threads, cfg.num_agents
    lo, hi = D.shard_of_threads(N, world, rank)
    c, s, a = _shard(cfg, store, adv, lo, hi)
    torch.manual_seed(123)
    perm = torch.randperm(T * N * M)                      # identical on every rank
    mb = T * N * M // 2
    mine = D.local_rows_of_global(perm[:mb], N, M, lo, hi).numpy().astype(np.int64)       # first minibatch
    tab, _ = O._flat_tables(s, a)
    act_rows = tab["active_masks"][mine]
    local = torch.tensor([[act_rows.sum(), tab["returns"][mine].sum(), (tab["returns"][mine] ** 2).sum(), len(mine)]], dtype=torch.float64)
    flat = D.pack_stats(local, torch.zeros(3, dtype=torch.float64))
    D.allreduce_sum_(flat)
    per_update, _ = D.unpack_stats(flat, 1)
    learner = O.Learner(c, pa, pc)
    sample = tuple(None if tab[nm] is None else tab[nm][mine] for nm in O._GEN_FIELDS)
    g = learner.ppo_update(sample, keep_grads=True)
    w = D.loss_weight(float(act_rows.sum()), len(mine), float(per_update[0, 0]), int(per_update[0, 3]), True)
    grads = torch.cat([v.reshape(-1) for v in g["actor_grads"].values()] + [v.reshape(-1) for v in g["critic_grads"].values()]) * w
    D.allreduce_sum_(grads)
    if rank == 0:
        np.save(out, np.concatenate([grads.numpy(), [float(per_update[0, 3])]]))
    dist.destroy_process_group()


def test_global_minibatch_partition_equals_single_process(tmp_path):
    out = str(tmp_path / "g.npy")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_mb, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    cfg, store, adv, pa, pc = _full_problem()
    T, N, M = cfg.episode_length, cfg.n_rollout_threads, cfg.num_agents
    torch.manual_seed(123)
    perm = torch.randperm(T * N * M).numpy()
    learner = O.Learner(cfg, pa, pc)
    tab, _ = O._flat_tables(store, adv)
    rows = perm[:T * N * M // 2]                         # the first of two minibatches (shared_buffer.py:358-361)
    g = learner.ppo_update(tuple(None if tab[nm] is None else tab[nm][rows] for nm in O._GEN_FIELDS), keep_grads=True)
    want = torch.cat([v.reshape(-1) for v in g["actor_grads"].values()] + [v.reshape(-1) for v in g["critic_grads"].values()]).numpy()
    assert got[-1] == len(rows)
    np.testing.assert_allclose(got[:-1], want, rtol=2e-4, atol=1e-7)


def test_local_rows_of_global_is_the_ownership_filter():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "on-policy_b200"))
    from mappo_b200 import dist as D
    T, N, M, world = 5, 8, 3, 4
    g = torch.arange(T * N * M).reshape(T, N, M)
    perm = torch.randperm(T * N * M)
    seen = []
    for rank in range(world):
        lo, hi = D.shard_of_threads(N, world, rank)
        loc = D.local_rows_of_global(perm, N, M, lo, hi).long()
        local_storage = g[:, lo:hi].reshape(-1)                # what the rank's own flattened storage holds (global row ids)
        picked = local_storage[loc]
        want = perm[(perm // M % N >= lo) & (perm // M % N < hi)]
        assert torch.equal(picked, want)                       # the owned rows, in permutation order
        seen.append(picked)
    assert torch.equal(torch.sort(torch.cat(seen)).values, torch.arange(T * N * M))
