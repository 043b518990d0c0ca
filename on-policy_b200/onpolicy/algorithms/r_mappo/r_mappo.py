"""onpolicy.algorithms.r_mappo.r_mappo.R_MAPPO on libmappo_b200 (reference: algorithms/r_mappo/r_mappo.py:8-232).

`train(buffer)` never materialises a minibatch: the permutation the reference would have drawn is uploaded as
an index list and the fused kernels read the rollout storage through it (gather fused into the loads).
Per optimiser step and net: ONE forward+loss+backward launch, one slot reduction, one clip+Adam launch; all loss
scalars stay on the device until the single read at the end of train().
"""
import ctypes as C
import os

import numpy as np
import torch

from mappo_b200 import _lib
from mappo_b200.core import (Batch, UpdateWorkspace, as_dev, check, launch_grads, launch_step, launch_update, launch_update_p2p,
                             make_loss_cfg, ptr, require_cuda, stream_ptr)
from onpolicy.utils.valuenorm import ValueNorm

INFO_KEYS = ("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio")


def _dist_allreduce():
    """Multi-GPU data parallelism over rollout threads: sum across ranks when a process group
    with more than one rank exists, else None."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return None


class R_MAPPO():
    def __init__(self, args, policy, device=torch.device("cpu")):
        self.device = policy.device
        self.tpdv = dict(dtype=torch.float32, device=self.device)
        self.policy = policy
        self.args = args
        self.clip_param = args.clip_param
        self.ppo_epoch = args.ppo_epoch
        self.num_mini_batch = args.num_mini_batch
        self.data_chunk_length = args.data_chunk_length
        self.value_loss_coef = args.value_loss_coef
        self.entropy_coef = args.entropy_coef
        self.max_grad_norm = args.max_grad_norm
        self.huber_delta = args.huber_delta
        self._use_recurrent_policy = args.use_recurrent_policy
        self._use_naive_recurrent = args.use_naive_recurrent_policy
        self._use_max_grad_norm = args.use_max_grad_norm
        self._use_clipped_value_loss = args.use_clipped_value_loss
        self._use_huber_loss = args.use_huber_loss
        self._use_popart = args.use_popart
        self._use_valuenorm = args.use_valuenorm
        self._use_value_active_masks = args.use_value_active_masks
        self._use_policy_active_masks = args.use_policy_active_masks
        assert (self._use_popart and self._use_valuenorm) == False, (
            "self._use_popart and self._use_valuenorm can not be set True simultaneously")
        if self._use_popart:
            raise NotImplementedError("use_popart: PopArt.update raises in the reference itself (SURVEY App. B-7)")
        self.value_normalizer = ValueNorm(1, device=self.device) if self._use_valuenorm else None
        self._happo = False               # set by onpolicy.algorithms.happo.happo_trainer.HAPPO (same kernels, three switches)
        # GEMM engine of the update kernels: "tf32" = tcgen05 tensor cores (fp32 accumulate), "fp32" = exact FFMA tiles
        self.gemm_mode = {"fp32": _lib.GEMM_FP32, "tf32": _lib.GEMM_TF32}[os.environ.get("MAPPO_B200_GEMM", "fp32")]
        # actor and critic are independent nets: their update chains run on two streams (fork / join with events,
        # captured as parallel branches of the CUDA graph).  Off when a process group is active (one collective order).
        self.overlap_nets = os.environ.get("MAPPO_B200_OVERLAP", "1") == "1"
        self._side = None
        # joint gradient vector [actor | critic]: what a multi-GPU run all-reduces in one collective per optimiser step
        na, nc = policy.actor.n_params, policy.critic.n_params
        pad = (-na) % 4
        self._joint_grad = torch.zeros(na + pad + nc, dtype=torch.float32, device=self.device)
        policy.actor.grad = self._joint_grad[:na]
        policy.critic.grad = self._joint_grad[na + pad:]
        self._crit_off = na + pad
        self._p2p, self._p2p_tried, self._parity = None, False, 0
        self._p2p_nets, self._par = None, [0, 0]
        self._ws = {}
        self._loss_out = torch.zeros(6, dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------------------------------
    def _workspaces(self, n_rows):
        key = int(n_rows)
        if key not in self._ws:
            self._ws[key] = (UpdateWorkspace(self.policy.actor, key, self.gemm_mode),
                             UpdateWorkspace(self.policy.critic, key, self.gemm_mode))
        return self._ws[key]

    def _invalidate_images(self):
        """The weights may have changed since the last fused optimiser tail (checkpoint load, another trainer): the first
        update of a train() / ppo_update() packs its weight image again."""
        for pair in self._ws.values():
            for ws in pair:
                ws.image_ready = 0

    def _one_update(self, batch, n_rows, norm_stats, adv_stats, loss_out, update_actor, allreduce, only=None, prepared=False):
        """One optimiser step of both nets (`only`: just the "actor" / "critic" chain).  `prepared`: the workspaces already
        hold the normalised input rows of this batch (hidden >= 128 nets, later epochs over the same rows)."""
        pol = self.policy
        ws_a, ws_c = self._workspaces(n_rows)
        loss_a = make_loss_cfg(self.args, update_actor)
        loss_c = make_loss_cfg(self.args, update_actor)
        loss_a.inputs_prepared = loss_c.inputs_prepared = int(bool(prepared))
        loss_a.happo = loss_c.happo = int(self._happo)
        vn = self.value_normalizer.state if self.value_normalizer is not None else None

        # data parallel with the peer-memory kernel: every net has its OWN reducer (buffer, signal pads, round counter),
        # so the actor and the critic chain stay independent across ranks too -- one all-reduce per optimiser step of
        # each optimiser, issued inside that net's chain
        per_net = allreduce is not None and only is not None and self._p2p_nets is not None

        def p2p_chain(k, net, ws, loss, a_stats, vn_state, opt, slot):
            r, par = self._p2p_nets[k], self._par[k]
            if ws.fused_tail:             # update kernel + ONE tail launch with the exchange inside
                launch_update_p2p(net, ws, batch, loss, norm_stats, a_stats, vn_state, loss_out, opt, self.max_grad_norm,
                                  self._use_max_grad_norm, slot, r, par)
                self._par[k] ^= 1
                return
            launch_grads(net, ws, batch, loss, norm_stats, a_stats, vn_state, loss_out,
                         grad_out=r.grad_half(par)[:net.n_params])       # local gradient straight into the peer-visible half
            nb = r.allreduce_grad(par, net.grad, ws.sumsq_part)      # + the partial sums of squares clip_adam needs
            self._par[k] ^= 1
            launch_step(net, ws, loss_out, opt, self.max_grad_norm, self._use_max_grad_norm, slot, nb)

        def actor_chain():          # backward, clip, step (reference :141-153)
            if per_net:
                return p2p_chain(0, pol.actor, ws_a, loss_a, adv_stats, None, pol.actor_optimizer, 3)
            launch_update(pol.actor, ws_a, batch, loss_a, norm_stats, adv_stats, None, loss_out, pol.actor_optimizer,
                          self.max_grad_norm, self._use_max_grad_norm, 3, allreduce)

        def critic_chain():         # ValueNorm.update(return_batch) BEFORE the value loss (reference :65), then :156-167
            if vn is not None and not self._happo:      # (happo_trainer.py:56-66 never updates its normaliser)
                check(_lib.load().mappo_valuenorm_update(ptr(vn), ptr(norm_stats), stream_ptr()))
            if per_net:
                return p2p_chain(1, pol.critic, ws_c, loss_c, None, vn, pol.critic_optimizer, 4)
            launch_update(pol.critic, ws_c, batch, loss_c, norm_stats, None, vn, loss_out, pol.critic_optimizer,
                          self.max_grad_norm, self._use_max_grad_norm, 4, allreduce)

        if only == "actor":
            actor_chain()
        elif only == "critic":
            critic_chain()
        elif self.overlap_nets and allreduce is None:      # a single ppo_update(): fork / join around this update
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                critic_chain()
            actor_chain()
            main.wait_stream(self._side)
        elif allreduce is not None:
            # data parallel: both nets' local gradients first, ONE collective on the joint gradient vector, then both
            # optimiser steps (actor and critic are independent, so this is the reference's order up to commuting)
            p2p = self._p2p
            ga = gc = None
            if p2p is not None:           # write the local gradients straight into the peer-visible half
                half = p2p.grad_half(self._parity)
                ga, gc = half[:pol.actor.n_params], half[self._crit_off:self._crit_off + pol.critic.n_params]
            two_streams = self.overlap_nets and p2p is not None      # our collective is an ordinary kernel on `main`
            main = torch.cuda.current_stream()
            if two_streams and self._side is None:
                self._side = torch.cuda.Stream(device=self.device)

            def critic_grads():
                if vn is not None and not self._happo:
                    check(_lib.load().mappo_valuenorm_update(ptr(vn), ptr(norm_stats), stream_ptr()))
                launch_grads(pol.critic, ws_c, batch, loss_c, norm_stats, None, vn, loss_out, grad_out=gc)

            def critic_step():
                launch_step(pol.critic, ws_c, loss_out, pol.critic_optimizer, self.max_grad_norm,
                            self._use_max_grad_norm, 4, 0)

            if two_streams:
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    critic_grads()
            launch_grads(pol.actor, ws_a, batch, loss_a, norm_stats, adv_stats, None, loss_out, grad_out=ga)
            if two_streams:
                main.wait_stream(self._side)
            else:
                critic_grads()
            if p2p is not None:
                p2p.allreduce_grad(self._parity, self._joint_grad)
                self._parity ^= 1
            else:
                allreduce(self._joint_grad)
            if two_streams:
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    critic_step()
            launch_step(pol.actor, ws_a, loss_out, pol.actor_optimizer, self.max_grad_norm, self._use_max_grad_norm, 3, 0)
            if two_streams:
                main.wait_stream(self._side)
            else:
                critic_step()
        else:
            actor_chain()
            critic_chain()

    def _restore_parity(self, k):
        """The reducers alternate two halves of their buffer (each round is a full barrier).  A captured graph replays a
        FIXED parity sequence, so an odd number of reductions per train() is padded with one 4-float round."""
        if self._p2p_nets is not None and self._par[k] & 1:
            self._p2p_nets[k].allreduce_grad(self._par[k], self._par_scratch)
            self._par[k] ^= 1

    def _storage_batch(self, buffer, adv, rows, first, seq_len, n_rows=None):
        b = Batch()
        T = buffer.episode_length
        b.obs, b.share_obs = ptr(buffer.obs), ptr(buffer.share_obs)
        b.actions, b.old_logp = ptr(buffer.actions), ptr(buffer.action_log_probs)
        b.value_preds, b.returns, b.advantages = ptr(buffer.value_preds), ptr(buffer.returns), ptr(adv)
        b.masks, b.active_masks = ptr(buffer.masks), ptr(buffer.active_masks)
        b.avail = ptr(buffer.available_actions)
        b.h0_actor, b.h0_critic = ptr(buffer.rnn_states), ptr(buffer.rnn_states_critic)
        if self._happo and getattr(buffer, "factor", None) is not None:
            b.factor = ptr(buffer.factor)                         # [T, N, 1] in storage row order (separated_buffer.py:62-63)
        b.rows, b.seq_first = ptr(rows), ptr(first)             # None: storage order
        b.n_rows = rows.numel() if rows is not None else int(n_rows)
        b.seq_len = seq_len
        b.n_seq = b.n_rows // seq_len
        return b

    # ------------------------------------------------------------------------------------------
    def _ensure_p2p(self, n_stats):
        """Peer-memory all-reduce (our kernel over NVLink) when the process group allows it; NCCL through
        torch.distributed otherwise (MAPPO_B200_P2P=0 forces NCCL)."""
        if self._p2p_tried:
            return
        self._p2p_tried = True
        if os.environ.get("MAPPO_B200_P2P", "1") != "1":
            return
        import torch.distributed as dist
        ok, err = 1, None
        try:
            from mappo_b200.dist import P2PReducer
            self._p2p = P2PReducer(self.device, self._joint_grad.numel(), n_stats)
            # per-net reducers keep the actor and the critic chain independent across ranks; their two spin-wait kernels must
            # be able to run CONCURRENTLY on every rank, so a serialising environment uses the single joint collective
            joint = os.environ.get("MAPPO_B200_P2P_JOINT", "0") == "1" or os.environ.get("CUDA_LAUNCH_BLOCKING", "0") == "1"
            if not joint:
                self._p2p_nets = (P2PReducer(self.device, self.policy.actor.n_params, 4),
                                  P2PReducer(self.device, self.policy.critic.n_params, 4))
                self._par = [0, 0]
                self._par_scratch = torch.zeros(4, dtype=torch.float32, device=self.device)
        except (RuntimeError, ImportError, AttributeError, NotImplementedError) as e:     # no symmetric memory here
            ok, err = 0, e
        # every rank must take the same path: agree on the outcome (a rank falling back to NCCL alone would leave its peers
        # spinning in the peer-memory kernel)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            import sys
            if dist.get_rank() == 0:
                print(f"[mappo_b200] peer-memory all-reduce unavailable on some rank ({type(err).__name__ if err else 'peer'}: {err}); "
                      "using NCCL", file=sys.stderr)
            self._p2p, self._p2p_nets = None, None

    def check_collectives(self):
        """Host check (synchronises): no peer-memory all-reduce of this trainer timed out."""
        for r in [self._p2p] + list(self._p2p_nets or ()):
            if r is not None:
                r.check_error()

    def _host_permutation(self, n):
        """torch.randperm on the CPU generator, exactly where the reference's generators draw it
        (utils/shared_buffer.py:360, 415, 511), uploaded as int32."""
        return torch.randperm(n).to(torch.int32).to(self.device, non_blocking=True)

    def perm_length(self, buffer):
        B = buffer.episode_length * buffer._E
        if self._use_recurrent_policy:
            return B // self.data_chunk_length
        if self._use_naive_recurrent:
            return buffer._E
        return B

    def launch_train(self, buffer, update_actor=True, draw_perm=None, loss_out=None, allreduce="auto"):
        """All launches of reference :171-224 without any host synchronisation (CUDA-graph capturable when
        `draw_perm` is).  Leaves the SUMS of the six train_info terms in `loss_out` (device, float64)."""
        lib = _lib.load()
        st = stream_ptr()
        dev = self.device
        T, E = buffer.episode_length, buffer._E
        B = T * E
        self._invalidate_images()
        if allreduce == "auto":
            allreduce = _dist_allreduce()
            if allreduce is not None:
                self._ensure_p2p(n_stats=self.ppo_epoch * self.num_mini_batch * 4 + 4)
        p2p = self._p2p if allreduce is not None else None
        # (with ONE minibatch per epoch -- every BASELINE config -- the update is permutation independent and each rank simply
        # permutes its own rows; device-drawn permutations (the engine's graph) keep per-rank permutations in every case)
        global_minibatches = (allreduce is not None and draw_perm is None and self.num_mini_batch > 1
                              and not (self._use_recurrent_policy or self._use_naive_recurrent))
        draw_perm = draw_perm or self._host_permutation
        loss_out = self._loss_out if loss_out is None else loss_out
        vn = self.value_normalizer.state if self.value_normalizer is not None else None

        # advantages + masked statistics (reference :179-187): reuse what compute_returns left when still valid
        want = id(self.value_normalizer) if self.value_normalizer is not None else 0
        adv, adv_stats = buffer.advantages, buffer._adv_stats
        if self._happo:                  # happo_trainer.py:181-184: value predictions are denormalised only under use_popart
            adv_stats.zero_()
            check(lib.mappo_advantages(ptr(buffer.returns), ptr(buffer.value_preds), ptr(buffer.active_masks), None,
                                       B, ptr(adv), ptr(adv_stats), st))
        elif buffer._adv_version != want:
            adv_stats.zero_()
            check(lib.mappo_advantages(ptr(buffer.returns), ptr(buffer.value_preds), ptr(buffer.active_masks), ptr(vn),
                                       B, ptr(adv), ptr(adv_stats), st))

        # minibatch plans: one permutation per epoch, drawn where the reference's generators would
        n_updates = self.ppo_epoch * self.num_mini_batch
        plans = []
        for _ in range(self.ppo_epoch):
            if self._use_recurrent_policy or self._use_naive_recurrent:
                L = self.data_chunk_length if self._use_recurrent_policy else T
                units = (B // L) if self._use_recurrent_policy else E
                mb = units // self.num_mini_batch
                perm = draw_perm(units)
                for i in range(self.num_mini_batch):
                    rows, first = buffer._chunk_rows(perm[i * mb:(i + 1) * mb].contiguous(), L)
                    plans.append((rows, first, L))
            elif global_minibatches:
                # data parallel with several minibatches: the reference's GLOBAL permutation (every rank draws it from the
                # same CPU generator state), each rank keeps the rows of its own rollout threads
                import torch.distributed as dist
                from mappo_b200.dist import local_rows_of_global, shard_of_threads
                ws, rk = dist.get_world_size(), dist.get_rank()
                n_loc, M = buffer.n_rollout_threads, buffer.num_agents
                lo, hi = shard_of_threads(n_loc * ws, ws, rk)
                perm_g = torch.randperm(B * ws)
                mbg = (B * ws) // self.num_mini_batch
                for i in range(self.num_mini_batch):
                    rows = local_rows_of_global(perm_g[i * mbg:(i + 1) * mbg], n_loc * ws, M, lo, hi).to(dev).contiguous()
                    plans.append((rows, rows, 1))
            else:
                mb = B // self.num_mini_batch
                perm = draw_perm(B)
                for i in range(self.num_mini_batch):
                    rows = perm[i * mb:(i + 1) * mb].contiguous()
                    plans.append((rows, rows, 1))

        # per-update statistics (sum active, sum R, sum R^2, rows): known up front -> one collective for all of them
        stats = torch.zeros(n_updates * 4 + 4, dtype=torch.float64, device=dev)
        addr = [pl[0].data_ptr() for pl in plans]
        step = addr[1] - addr[0] if n_updates > 1 else 0
        n_each = plans[0][0].numel()
        if (n_updates > 1 and step > 0 and step % 4 == 0 and all(pl[0].numel() == n_each for pl in plans)
                and all(addr[u] == addr[0] + u * step for u in range(n_updates))):
            # equally spaced slices of one permutation buffer: every minibatch's statistics in ONE launch
            check(lib.mappo_minibatch_stats_batch(ptr(buffer.returns), ptr(buffer.active_masks), ptr(plans[0][0]),
                                                  step // 4, n_each, n_updates, ptr(stats), st))
        else:
            for u, (rows, _, _) in enumerate(plans):
                check(lib.mappo_minibatch_stats(ptr(buffer.returns), ptr(buffer.active_masks), ptr(rows), rows.numel(),
                                                C.c_void_p(stats.data_ptr() + 32 * u), st))
        if allreduce is not None:
            stats[n_updates * 4:n_updates * 4 + 3].copy_(adv_stats)
            if p2p is not None:
                p2p.allreduce_f64_(stats, "stats")
            else:
                allreduce(stats)
            adv_stats = stats[n_updates * 4:n_updates * 4 + 3]

        loss_out.zero_()
        # hidden >= 128 nets with ONE minibatch per epoch (every BASELINE config): an epoch is a sum over all rows, so the
        # rows are processed in storage order (no gather; the permutation is still drawn where the reference draws it) and
        # the normalised inputs computed by the first epoch are reused by the later ones (the buffer does not change)
        in_order = (bool(lib.mappo_big_net(C.byref(self.policy.actor.desc))) and self.num_mini_batch == 1 and
                    not (self._use_recurrent_policy or self._use_naive_recurrent) and
                    os.environ.get("MAPPO_B200_BIG_IN_ORDER", "1") == "1")
        batches = [self._storage_batch(buffer, adv, None if in_order else rows, None if in_order else first, seq_len,
                                       n_rows=rows.numel()) for rows, first, seq_len in plans]

        def updates(only):
            for u, (rows, _, _) in enumerate(plans):
                self._one_update(batches[u], rows.numel(), stats[4 * u:4 * u + 4], adv_stats, loss_out, update_actor,
                                 allreduce, only, prepared=in_order and u > 0)

        if self.overlap_nets and (allreduce is None or self._p2p_nets is not None):
            # actor and critic never read each other's state inside train(): their whole update sequences are two
            # independent chains (fork once, join once).  Free-running, the two 75-CTA update kernels interleave on
            # the 148 SMs instead of colliding in lock step (150 CTAs = two waves).  Multi-GPU: each chain carries its
            # own peer-memory all-reduce (per-net reducers), so the chains stay independent across ranks as well.
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                updates("critic")
                self._restore_parity(1)
            updates("actor")
            self._restore_parity(0)
            main.wait_stream(self._side)
        else:
            updates(None)
            if p2p is not None and self._parity & 1:          # joint reducer: same fixed-parity rule under graph replay
                p2p.allreduce_grad(self._parity, self._joint_grad[:4].clone())
                self._parity ^= 1
        if allreduce is not None:
            # loss scalars are partial sums over local rows (global normalisers); norms are already global
            part = loss_out.clone()
            part[3:5] = 0
            if p2p is not None:
                p2p.allreduce_f64_(part, "loss")
            else:
                allreduce(part)
            loss_out[0:3] = part[0:3]
            loss_out[5] = part[5]
        self._keepalive = plans
        # the advantages / statistics cached by compute_returns were formed with the ValueNorm state of BEFORE this train():
        # a second train() on the same buffer must recompute them (reference r_mappo.py:179-187 recomputes every time)
        buffer._adv_version = -1
        return loss_out, n_updates

    def train(self, buffer, update_actor=True):
        """reference :171-224.  One host read at the very end."""
        self.policy.actor_optimizer.sync_lr()
        self.policy.critic_optimizer.sync_lr()
        loss_out, n_updates = self.launch_train(buffer, update_actor)
        vals = (loss_out / float(n_updates)).cpu().tolist()
        return dict(zip(INFO_KEYS, vals))

    # ------------------------------------------------------------------------------------------
    def _sample_batch(self, sample):
        dev = self.device
        (share_obs, obs, h_a, h_c, actions, v_old, ret, masks, active, lp_old, adv, avail) = \
            [as_dev(x, dev) for x in sample[:12]]
        n_rows = obs.shape[0]
        n_seq = h_a.shape[0] if (self._use_recurrent_policy or self._use_naive_recurrent) else n_rows
        b = Batch()
        b.obs, b.share_obs, b.actions, b.old_logp = ptr(obs), ptr(share_obs), ptr(actions), ptr(lp_old)
        b.value_preds, b.returns, b.advantages = ptr(v_old), ptr(ret), ptr(adv)
        b.masks, b.active_masks, b.avail = ptr(masks), ptr(active), ptr(avail)
        b.h0_actor = ptr(h_a.reshape(h_a.shape[0], -1).contiguous())
        b.h0_critic = ptr(h_c.reshape(h_c.shape[0], -1).contiguous())
        b.rows = b.seq_first = None
        b.n_rows, b.n_seq, b.seq_len = n_rows, n_seq, n_rows // n_seq
        factor = as_dev(sample[12], dev) if (self._happo and len(sample) > 12) else None      # 13-tuples of separated buffers
        b.factor = ptr(factor)
        keep = (share_obs, obs, h_a, h_c, actions, v_old, ret, masks, active, lp_old, adv, avail, factor)
        return b, keep, ret, active

    def ppo_update(self, sample, update_actor=True):
        """reference :91-169 on a materialised sample (12- or 13-tuple as the generators yield them).
        Returns (value_loss, critic_grad_norm, policy_loss, dist_entropy, actor_grad_norm, imp_weights) as device
        scalars; `imp_weights` is the mean ratio as a 1-element tensor (its only use is `.mean()`, :217)."""
        lib = _lib.load()
        batch, keep, ret, active = self._sample_batch(sample)
        self._invalidate_images()
        allreduce = _dist_allreduce()
        stats = torch.zeros(4, dtype=torch.float64, device=self.device)
        check(lib.mappo_minibatch_stats(ptr(ret), ptr(active), None, batch.n_rows, ptr(stats), stream_ptr()))
        if allreduce is not None:
            allreduce(stats)
        out = torch.zeros(6, dtype=torch.float64, device=self.device)
        self.policy.actor_optimizer.sync_lr()
        self.policy.critic_optimizer.sync_lr()
        self._one_update(batch, batch.n_rows, stats, None, out, update_actor, allreduce)
        o = out.to(torch.float32)
        return o[0], o[4], o[1], o[2], o[3], o[5:6]

    def cal_value_loss(self, values, value_preds_batch, return_batch, active_masks_batch):
        """reference :52-89: clipped / Huber value loss of `values` against the (ValueNorm-normalised) returns, INCLUDING the
        `value_normalizer.update(return_batch)` side effect (:65).  Device tensors in, a device scalar out.  The training
        path does not come through here (the loss is fused into the critic's update kernel, net_tiles.cuh row_loss_pre);
        this is the reference's public method kept for callers that evaluate the loss on their own tensors -- a handful
        of elementwise torch ops on the device (no autograd graph: values produced by the engine carry none)."""
        dev = self.device
        values, v_old = as_dev(values, dev), as_dev(value_preds_batch, dev)
        ret, active = as_dev(return_batch, dev), as_dev(active_masks_batch, dev)
        v_clip = v_old + (values - v_old).clamp(-self.clip_param, self.clip_param)
        if self._use_valuenorm:
            if not self._happo:
                self.value_normalizer.update(ret)
            target = self.value_normalizer.normalize(ret)
        else:
            target = ret
        e_c, e_o = target - v_clip, target - values
        if self._use_huber_loss:
            d = self.huber_delta
            hub = lambda e: torch.where(e.abs() <= d, e * e / 2, d * (e.abs() - d / 2))
            l_c, l_o = hub(e_c), hub(e_o)
        else:
            l_c, l_o = e_c * e_c / 2, e_o * e_o / 2
        vl = torch.max(l_o, l_c) if self._use_clipped_value_loss else l_o
        if self._use_value_active_masks:
            return (vl * active).sum() / active.sum()
        return vl.mean()

    def prep_training(self):
        self.policy.actor.train()
        self.policy.critic.train()

    def prep_rollout(self):
        self.policy.actor.eval()
        self.policy.critic.eval()


def _evaluate_actor(actor, obs, h_a, action, masks, available_actions, active_masks, use_policy_active_masks=True):
    """R_Actor.evaluate_actions (r_actor_critic.py:73-117): log-probs [rows, as] and the (masked) mean entropy; no gradients."""
    lib = _lib.load()
    dev = actor.device
    obs, action, masks = as_dev(obs, dev), as_dev(action, dev), as_dev(masks, dev)
    n_rows = obs.shape[0]
    active = as_dev(active_masks, dev) if active_masks is not None else torch.ones(n_rows, 1, device=dev)
    recurrent = bool(actor.desc.recurrent)
    h_a = as_dev(h_a, dev)
    n_seq = h_a.shape[0] if recurrent else n_rows
    b = Batch()
    b.obs, b.actions, b.masks, b.active_masks = ptr(obs), ptr(action), ptr(masks), ptr(active)
    avail_d = as_dev(available_actions, dev) if available_actions is not None else None
    b.avail = ptr(avail_d)
    h_a2 = h_a.reshape(h_a.shape[0], -1).contiguous()
    b.h0_actor = ptr(h_a2)
    b.n_rows, b.n_seq, b.seq_len = n_rows, n_seq, n_rows // n_seq
    stats = torch.tensor([float(active.sum().item()), 0.0, 0.0, float(n_rows)], dtype=torch.float64, device=dev)
    out = torch.zeros(6, dtype=torch.float64, device=dev)
    args_like = type("A", (), dict(clip_param=0.2, entropy_coef=0.0, value_loss_coef=1.0, huber_delta=10.0,
                                   use_clipped_value_loss=True, use_huber_loss=True, use_value_active_masks=True,
                                   use_policy_active_masks=(active_masks is not None and use_policy_active_masks),
                                   use_valuenorm=False, use_popart=False))
    loss = make_loss_cfg(args_like)
    logp = torch.empty(n_rows, len(actor.head_dims), dtype=torch.float32, device=dev)
    ws = UpdateWorkspace(actor, n_rows)
    check(lib.mappo_evaluate_actions(C.byref(actor.desc), ptr(actor.flat), C.byref(b), C.byref(loss), ptr(stats), ptr(logp),
                                     ptr(out), ptr(ws.workspace), stream_ptr()))
    return logp, out[2].to(torch.float32)


def _evaluate_only(policy, cent_obs, obs, h_a, h_c, action, masks, available_actions, active_masks):
    """policy.evaluate_actions (rMAPPOPolicy.py:88-114): values, log-probs, entropy; no gradients."""
    lib = _lib.load()
    dev = policy.device
    obs, cent, action, masks = as_dev(obs, dev), as_dev(cent_obs, dev), as_dev(action, dev), as_dev(masks, dev)
    n_rows = obs.shape[0]
    active = as_dev(active_masks, dev) if active_masks is not None else torch.ones(n_rows, 1, device=dev)
    h_a, h_c = as_dev(h_a, dev), as_dev(h_c, dev)
    recurrent = bool(policy.actor.desc.recurrent)
    n_seq = h_a.shape[0] if recurrent else n_rows
    b = Batch()
    b.obs, b.share_obs, b.actions, b.masks, b.active_masks = ptr(obs), ptr(cent), ptr(action), ptr(masks), ptr(active)
    avail_d = as_dev(available_actions, dev) if available_actions is not None else None     # keep alive
    b.avail = ptr(avail_d)
    h_a2, h_c2 = h_a.reshape(h_a.shape[0], -1).contiguous(), h_c.reshape(h_c.shape[0], -1).contiguous()
    b.h0_actor, b.h0_critic = ptr(h_a2), ptr(h_c2)
    b.n_rows, b.n_seq, b.seq_len = n_rows, n_seq, n_rows // n_seq
    stats = torch.tensor([float(active.sum().item()), 0.0, 0.0, float(n_rows)], dtype=torch.float64, device=dev)
    out = torch.zeros(6, dtype=torch.float64, device=dev)
    args_like = type("A", (), dict(clip_param=0.2, entropy_coef=0.0, value_loss_coef=1.0, huber_delta=10.0,
                                   use_clipped_value_loss=True, use_huber_loss=True, use_value_active_masks=True,
                                   use_policy_active_masks=(active_masks is not None and policy._use_policy_active_masks),
                                   use_valuenorm=False,
                                   use_popart=False))
    loss = make_loss_cfg(args_like)
    logp = torch.empty(n_rows, len(policy.actor.head_dims), dtype=torch.float32, device=dev)
    values = torch.empty(n_rows, 1, dtype=torch.float32, device=dev)
    ws_a = UpdateWorkspace(policy.actor, n_rows)
    ws_c = UpdateWorkspace(policy.critic, n_rows)
    st = stream_ptr()
    check(lib.mappo_evaluate_actions(C.byref(policy.actor.desc), ptr(policy.actor.flat), C.byref(b), C.byref(loss),
                                     ptr(stats), ptr(logp), ptr(out), ptr(ws_a.workspace), st))
    check(lib.mappo_evaluate_actions(C.byref(policy.critic.desc), ptr(policy.critic.flat), C.byref(b), C.byref(loss),
                                     ptr(stats), ptr(values), ptr(out), ptr(ws_c.workspace), st))
    return values, logp, out[2].to(torch.float32)
