"""Data-parallel plumbing over rollout threads: which threads a rank owns, and the two collectives
of a `train()` -- the statistics all-reduce that makes every normaliser global and the gradient all-reduce.
Device-agnostic on purpose (works with gloo on CPU tensors) so the host logic is testable without GPUs."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_of_threads(n_rollout_threads_global: int, world_size: int, rank: int):
    """Contiguous block of rollout threads owned by `rank` (axis 1 of every storage array)."""
    if n_rollout_threads_global % world_size:
        raise ValueError(f"n_rollout_threads {n_rollout_threads_global} not divisible by world size {world_size}")
    per = n_rollout_threads_global // world_size
    return rank * per, (rank + 1) * per


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    ws, _ = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def pack_stats(per_update_stats: torch.Tensor, adv_stats: torch.Tensor) -> torch.Tensor:
    """[n_updates, 4] (sum active, sum R, sum R^2, rows) + [3] (sum adv, sum adv^2, count) -> one flat fp64 payload."""
    return torch.cat([per_update_stats.reshape(-1).double(), adv_stats.reshape(-1).double()])


def unpack_stats(flat: torch.Tensor, n_updates: int):
    return flat[:4 * n_updates].reshape(n_updates, 4), flat[4 * n_updates:4 * n_updates + 3]


def loss_weight(active_local_sum: float, rows_local: int, active_global_sum: float, rows_global: int, use_active: bool):
    """Factor turning a rank-local masked mean into its share of the global masked mean."""
    return (active_local_sum / active_global_sum) if use_active else (rows_local / rows_global)
