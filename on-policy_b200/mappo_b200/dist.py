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


def local_rows_of_global(rows_global: torch.Tensor, n_threads_global: int, num_agents: int, lo: int, hi: int) -> torch.Tensor:
    """The rows of a GLOBAL minibatch (indices into the flattened [T, N_global, M] storage, row = (t N + n) M + m, as drawn by
    the reference's `torch.randperm(B_global)`, shared_buffer.py:360) that live on the rank owning rollout threads [lo, hi),
    renumbered for that rank's local [T, hi - lo, M] storage.  Order preserved (minibatch partitioning)."""
    r = rows_global.to(torch.int64)
    m = r % num_agents
    tn = r // num_agents
    n = tn % n_threads_global
    t = tn // n_threads_global
    keep = (n >= lo) & (n < hi)
    local = (t * (hi - lo) + (n - lo)) * num_agents + m
    return local[keep].to(torch.int32)


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


class P2PReducer:
    """All-reduce through libmappo_b200's one-shot peer-memory kernel (mappo_p2p_allreduce_*): buffers come from
    torch.distributed._symmetric_memory (plumbing: allocation + exchange of the peer mappings), the reduction itself is
    our kernel -- no NCCL call, no host work per collective, so a multi-GPU iteration stays ONE CUDA graph.

    Layout of the symmetric buffer (floats): [grad half 0][grad half 1][stats (fp64)][loss (fp64)]."""

    def __init__(self, device, n_grad: int, n_stats: int):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.lib, self.C = _lib.load(), C
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 8:
            raise RuntimeError("P2PReducer: one NVSwitch box (<= 8 ranks)")
        pad = lambda n: (n + 3) & ~3
        self.n_grad = pad(n_grad)
        self.n_stats = n_stats
        self.off_grad = [0, self.n_grad]
        self.off_stats = 2 * self.n_grad
        self.off_loss = self.off_stats + pad(2 * n_stats)
        total = self.off_loss + 16
        self.sym = symm_mem.empty(total, dtype=torch.float32, device=device)
        self.sig = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.sym.zero_()
        self.sig.zero_()
        torch.cuda.synchronize()
        h_buf = symm_mem.rendezvous(self.sym, dist.group.WORLD)
        h_sig = symm_mem.rendezvous(self.sig, dist.group.WORLD)
        self.bufs = (C.c_void_p * self.world)(*[int(p) for p in h_buf.buffer_ptrs])
        self.sigs = (C.c_void_p * self.world)(*[int(p) for p in h_sig.buffer_ptrs])
        self._keep = (h_buf, h_sig)
        self.round = torch.zeros(4, dtype=torch.int32, device=device)      # {round, CTA ticket, error flag, -}
        dist.barrier()
        torch.cuda.synchronize()

    def check_error(self):
        """Host read of the kernel's error flag (a peer's arrival timed out): synchronises; raises if set."""
        flag = int(self.round[2].item())
        if flag:
            raise RuntimeError(f"peer-memory all-reduce: rank {self.rank} timed out waiting for rank {flag - 1} "
                               "(dead peer, or the two per-net reducers could not run concurrently on some rank)")

    def grad_half(self, parity: int) -> torch.Tensor:
        o = self.off_grad[parity & 1]
        return self.sym[o:o + self.n_grad]

    def _stream(self):
        return self.C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def allreduce_grad(self, parity: int, out: torch.Tensor, sumsq_part: torch.Tensor = None) -> int:
        """out[:] = sum over ranks of grad_half(parity)  (out is an ordinary local tensor).  With `sumsq_part` the
        kernel also leaves per-CTA sums of squares of the result there; returns how many."""
        nb = self.C.c_int32(0)
        rc = self.lib.mappo_p2p_allreduce_f32(self.bufs, self.sigs, self.world, self.rank, 4 * self.off_grad[parity & 1],
                                              out.numel(), out.data_ptr(), self.round.data_ptr(),
                                              None if sumsq_part is None else sumsq_part.data_ptr(),
                                              self.C.byref(nb), self._stream())
        _check(self.lib, rc)
        return nb.value if sumsq_part is not None else 0

    def allreduce_f64_(self, t: torch.Tensor, region: str):
        """In-place sum over ranks of a small fp64 tensor through the 'stats' or 'loss' region."""
        off = self.off_stats if region == "stats" else self.off_loss
        n = t.numel()
        self.sym[off:off + 2 * n].view(torch.float64).copy_(t.reshape(-1))
        rc = self.lib.mappo_p2p_allreduce_f64(self.bufs, self.sigs, self.world, self.rank, 4 * off, n, t.data_ptr(),
                                              self.round.data_ptr(), self._stream())
        _check(self.lib, rc)
        return t


def _check(lib, rc):
    if rc != 0:
        raise RuntimeError(f"libmappo_b200: status {rc}: {lib.mappo_last_error().decode()}")
