"""2-GPU data parallelism over rollout threads reproduces the reference's single-process iteration (golden c1):
each rank owns half of the threads, normalisers are global, gradients are all-reduced."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (root, os.path.join(root, "on-policy_b200"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from oracle import mappo_oracle as O
    from helpers import Golden
    from argsutil import make_args, make_spaces
    from mappo_b200.dist import shard_of_threads
    from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy
    from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO
    from onpolicy.utils.shared_buffer import SharedReplayBuffer
    import test_gpu_parity as TP

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = Golden("c1_mlp_discrete")
    cfg = g.cfg
    lo, hi = shard_of_threads(cfg.n_rollout_threads, world, rank)
    c = O.PathConfig(**{**cfg.to_dict(), "n_rollout_threads": hi - lo, "act_dims": tuple(cfg.act_dims)})
    args = make_args(c)
    obs_s, share_s, act_s = make_spaces(c)
    dev = torch.device("cuda", rank)
    policy = R_MAPPOPolicy(args, obs_s, share_s, act_s, device=dev)
    policy.actor.load_state_dict(g.init_params("actor"))
    policy.critic.load_state_dict(g.init_params("critic"))
    trainer = R_MAPPO(args, policy, device=dev)
    buf = SharedReplayBuffer(args, c.num_agents, obs_s, share_s, act_s)
    feed = g.feed(0)
    sh = O.SyntheticFeed(feed.obs[:, lo:hi], feed.share_obs[:, lo:hi], feed.rewards[:, lo:hi], feed.dones[:, lo:hi])
    M, A = cfg.num_agents, sum(cfg.act_dims)
    noise = g.get("it0/noise").reshape(cfg.episode_length, cfg.n_rollout_threads, M, A)[:, lo:hi]
    noise = np.ascontiguousarray(noise).reshape(cfg.episode_length, (hi - lo) * M, A)
    TP.warm(buf, sh)
    # patch helpers that assume cuda:0
    TP.collect_and_returns.__globals__["torch"] = torch
    with torch.cuda.device(rank):
        TP.collect_and_returns(c, policy, trainer, buf, sh, noise)
        info = trainer.train(buf)
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out_path, actor=policy.actor.flat.cpu().numpy(), critic=policy.critic.flat.cpu().numpy(),
                 info=np.array([info[k] for k in TP.INFO_KEYS]), vn=trainer.value_normalizer.state.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_iteration_matches_reference(tmp_path):
    import torch.multiprocessing as mp
    from helpers import Golden, INFO_KEYS, assert_close
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, out), nprocs=2, join=True)
    z = np.load(out)
    g = Golden("c1_mlp_discrete")
    want = np.concatenate([v.numpy().reshape(-1) for v in g.params("it0/actor/").values()])
    # flat layout == state_dict order for this net (feature norm, fc1, ln1, fc2, ln2, head)
    assert_close(z["actor"], want, 2e-3, 2e-5, "actor weights after 2-rank train")
    wantc = np.concatenate([v.numpy().reshape(-1) for v in g.params("it0/critic/").values()])
    assert_close(z["critic"], wantc, 2e-3, 2e-5, "critic weights after 2-rank train")
    assert_close(z["info"], g.get("it0/train_info"), 2e-3, 2e-5, "train_info")
    assert_close(z["vn"], g.get("it0/valuenorm"), 1e-4, 1e-8, "valuenorm")
