"""Device-side building blocks behind the drop-in `onpolicy.*` classes.

Everything here is plumbing around libmappo_b200.so: PyTorch allocates CUDA memory and provides
streams; all arithmetic of the hot path runs in the library's kernels.  There is deliberately no
fallback: without the library (or without a CUDA device) these objects cannot be constructed.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from collections import OrderedDict
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import Batch, LossCfg, NetDesc, NetLayout, check, ptr


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("mappo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def as_dev(x, device, dtype=torch.float32):
    """numpy / torch (any device) -> contiguous CUDA tensor of `dtype` (reference `check`, algorithms/utils/util.py:16-18)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not torch.is_tensor(x):
        x = torch.as_tensor(x)
    return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


# ---------------------------------------------------------------------------------------------
# spaces (duck typed exactly like the reference: utils/util.py:31-51, algorithms/utils/act.py:18-42)
# ---------------------------------------------------------------------------------------------
def obs_dim_of(space) -> int:
    name = space.__class__.__name__
    if name == "Box":
        shape = space.shape
    elif name == "list":
        shape = space
    else:
        raise NotImplementedError(f"observation space {name}")
    if type(shape[-1]) == list:                     # shared_buffer.py:48-52
        shape = shape[:1]
    if len(shape) != 1:
        raise NotImplementedError("image observations (CNNBase) are outside the B200 hot path (SURVEY 2.1 row 8)")
    return int(shape[0])


def act_heads_of(space):
    """-> (head_dims, multi_discrete)."""
    name = space.__class__.__name__
    if name == "Discrete":
        return [int(space.n)], False
    if name == "MultiDiscrete":
        dims = [int(h - l + 1) for h, l in zip(space.high, space.low)]
        return dims, True
    raise NotImplementedError(f"action space {name}: only Discrete / MultiDiscrete are on the B200 hot path")


# ---------------------------------------------------------------------------------------------
# one network = flat fp32 parameter vector + the reference's state_dict view of it
# ---------------------------------------------------------------------------------------------
class DeviceNet:
    """R_Actor / R_Critic stand-in (algorithms/r_mappo/algorithm/r_actor_critic.py:12-175).

    Parameters live in ONE flat CUDA vector laid out by mappo_net_layout(); `state_dict()` exposes views
    under the reference's key names so checkpoints interchange with the reference.
    """

    def __init__(self, args, in_dim: int, head_dims, is_critic: bool, device):
        self.device = require_cuda(device)
        lib = _lib.load()
        recurrent = bool(args.use_recurrent_policy or args.use_naive_recurrent_policy)
        if recurrent and int(args.recurrent_N) != 1:
            raise NotImplementedError("recurrent_N != 1 is not built (every reference script uses 1)")
        d = NetDesc()
        d.in_dim, d.hidden, d.layer_n = int(in_dim), int(args.hidden_size), int(args.layer_N)
        d.use_feature_norm = int(bool(args.use_feature_normalization))
        d.use_relu = int(bool(args.use_ReLU))
        d.recurrent = int(recurrent)
        d.n_heads = len(head_dims)
        for k, a in enumerate(head_dims):
            d.head_dim[k] = int(a)
        d.is_critic = int(is_critic)
        self.desc = d
        self.layout = NetLayout()
        check(lib.mappo_net_layout(C.byref(d), C.byref(self.layout)))
        self.head_dims = list(head_dims)
        self.multi_discrete = len(head_dims) > 1 or getattr(args, "_b200_multi_discrete", False)
        self.hidden = d.hidden
        self.n_params = int(self.layout.total)
        self.flat = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros_like(self.flat)
        self._keys = self._key_table()
        self.training = True

    # -- state_dict key table -------------------------------------------------------------------
    def _key_table(self):
        L, d = self.layout, self.desc
        H, I = d.hidden, d.in_dim
        t = []
        if d.use_feature_norm:
            t += [("base.feature_norm.weight", L.fn_w, (I,)), ("base.feature_norm.bias", L.fn_b, (I,))]
        t += [("base.mlp.fc1.0.weight", L.fc1_w, (H, I)), ("base.mlp.fc1.0.bias", L.fc1_b, (H,)),
              ("base.mlp.fc1.2.weight", L.ln1_w, (H,)), ("base.mlp.fc1.2.bias", L.ln1_b, (H,))]
        for i in range(d.layer_n):
            t += [(f"base.mlp.fc2.{i}.0.weight", L.fc2_w[i], (H, H)), (f"base.mlp.fc2.{i}.0.bias", L.fc2_b[i], (H,)),
                  (f"base.mlp.fc2.{i}.2.weight", L.ln2_w[i], (H,)), (f"base.mlp.fc2.{i}.2.bias", L.ln2_b[i], (H,))]
        if d.recurrent:
            t += [("rnn.rnn.weight_ih_l0", L.gru_wih, (3 * H, H)), ("rnn.rnn.weight_hh_l0", L.gru_whh, (3 * H, H)),
                  ("rnn.rnn.bias_ih_l0", L.gru_bih, (3 * H,)), ("rnn.rnn.bias_hh_l0", L.gru_bhh, (3 * H,)),
                  ("rnn.norm.weight", L.rnn_ln_w, (H,)), ("rnn.norm.bias", L.rnn_ln_b, (H,))]
        if d.is_critic:
            t += [("v_out.weight", L.head_w, (1, H)), ("v_out.bias", L.head_b, (1,))]
        elif self.multi_discrete:
            off = 0
            for k, a in enumerate(self.head_dims):
                t += [(f"act.action_outs.{k}.linear.weight", L.head_w + off * H, (a, H)),
                      (f"act.action_outs.{k}.linear.bias", L.head_b + off, (a,))]
                off += a
        else:
            a = self.head_dims[0]
            t += [("act.action_out.linear.weight", L.head_w, (a, H)), ("act.action_out.linear.bias", L.head_b, (a,))]
        return t

    def _view(self, base, off, shape):
        n = int(np.prod(shape))
        return base[off:off + n].view(*shape)

    # -- nn.Module-like surface the reference's runner uses --------------------------------------
    def state_dict(self):
        return OrderedDict((k, self._view(self.flat, o, s)) for k, o, s in self._keys)

    def load_state_dict(self, sd, strict=True):
        want = {k for k, _, _ in self._keys}
        if strict and set(sd.keys()) != want:
            raise RuntimeError(f"state_dict keys mismatch: missing {sorted(want - set(sd))}, "
                               f"unexpected {sorted(set(sd) - want)}")
        for k, o, s in self._keys:
            if k in sd:
                src = sd[k]
                src = torch.from_numpy(src) if isinstance(src, np.ndarray) else src
                self._view(self.flat, o, s).copy_(src.to(dtype=torch.float32).reshape(s))

    def named_parameters(self):
        return [(k, self._view(self.flat, o, s)) for k, o, s in self._keys]

    def parameters(self):
        return [p for _, p in self.named_parameters()]

    def named_grads(self):
        return OrderedDict((k, self._view(self.grad, o, s)) for k, o, s in self._keys)

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, *a, **k):
        return self

    def init_like_reference(self, args):
        self.load_state_dict(reference_init_state_dict(args, self.desc.in_dim, self.head_dims,
                                                       bool(self.desc.is_critic), self.multi_discrete))


def reference_init_state_dict(args, in_dim, head_dims, is_critic, multi_discrete):
    """Initial weights drawn like the reference draws them (host only, no CUDA needed).

    Builds throw-away torch.nn layers in the construction order of R_Actor / R_Critic
    (r_actor_critic.py:32-41, 139-152; mlp.py:8-24, 46-49; rnn.py:13-22; distributions.py:56-62) so that the global
    CPU generator is consumed identically: `torch.manual_seed(s)` then yields the reference's initial weights bit
    for bit."""
    import torch.nn as nn
    H, I = int(args.hidden_size), int(in_dim)
    use_relu = bool(args.use_ReLU)
    recurrent = bool(args.use_recurrent_policy or args.use_naive_recurrent_policy)
    ortho = bool(args.use_orthogonal)
    winit = nn.init.orthogonal_ if ortho else nn.init.xavier_uniform_
    gain = nn.init.calculate_gain("relu" if use_relu else "tanh")
    sd = OrderedDict()

    def lin(key, i, o, g):
        m = nn.Linear(i, o)
        winit(m.weight.data, gain=g)
        nn.init.constant_(m.bias.data, 0)
        sd[key + ".weight"], sd[key + ".bias"] = m.weight.data, m.bias.data

    def ln(key, n):
        sd[key + ".weight"], sd[key + ".bias"] = torch.ones(n), torch.zeros(n)

    if args.use_feature_normalization:
        ln("base.feature_norm", I)
    lin("base.mlp.fc1.0", I, H, gain)
    ln("base.mlp.fc1.2", H)
    for i in range(int(args.layer_N)):
        lin(f"base.mlp.fc2.{i}.0", H, H, gain)
        ln(f"base.mlp.fc2.{i}.2", H)
    if recurrent:
        gru = nn.GRU(H, H, num_layers=int(args.recurrent_N))
        for name, p in gru.named_parameters():
            if "bias" in name:
                nn.init.constant_(p, 0)
            elif "weight" in name:
                (nn.init.orthogonal_ if ortho else nn.init.xavier_uniform_)(p)
            sd["rnn.rnn." + name] = p.data
        ln("rnn.norm", H)
    if is_critic:
        lin("v_out", H, 1, 1)
    elif multi_discrete:
        for k, a in enumerate(head_dims):
            lin(f"act.action_outs.{k}.linear", H, a, args.gain)
    else:
        lin("act.action_out.linear", H, head_dims[0], args.gain)
    return sd


# ---------------------------------------------------------------------------------------------
# ValueNorm (utils/valuenorm.py:8-79) with its three running scalars resident on the device
# ---------------------------------------------------------------------------------------------
class DeviceValueNorm:
    def __init__(self, input_shape=1, device=None):
        self.device = require_cuda(device)
        self.state = torch.zeros(3, dtype=torch.float32, device=self.device)   # running_mean, running_mean_sq, debias
        self._stats = torch.zeros(4, dtype=torch.float64, device=self.device)

    # reference attribute names
    @property
    def running_mean(self):
        return self.state[0:1]

    @property
    def running_mean_sq(self):
        return self.state[1:2]

    @property
    def debiasing_term(self):
        return self.state[2]

    def running_mean_var(self):
        d = self.state[2].clamp(min=1e-5)
        mean = self.state[0:1] / d
        var = (self.state[1:2] / d - mean ** 2).clamp(min=1e-2)
        return mean, var

    @torch.no_grad()
    def update(self, input_vector):
        x = as_dev(input_vector, self.device).reshape(-1)
        lib = _lib.load()
        self._stats.zero_()
        ones = torch.ones_like(x)
        check(lib.mappo_minibatch_stats(ptr(x), ptr(ones), None, x.numel(), ptr(self._stats), stream_ptr()))
        check(lib.mappo_valuenorm_update(ptr(self.state), ptr(self._stats), stream_ptr()))

    def normalize(self, input_vector):
        x = as_dev(input_vector, self.device)
        mean, var = self.running_mean_var()
        return (x - mean) / torch.sqrt(var)

    def denormalize(self, input_vector):
        was_np = isinstance(input_vector, np.ndarray)
        x = as_dev(input_vector, self.device)
        mean, var = self.running_mean_var()
        out = x * torch.sqrt(var) + mean
        return out.cpu().numpy() if was_np else out            # the reference returns NumPy (valuenorm.py:77)

    def state_dict(self):
        return OrderedDict(running_mean=self.state[0:1].clone(), running_mean_sq=self.state[1:2].clone(),
                           debiasing_term=self.state[2].clone())

    def load_state_dict(self, sd):
        self.state[0] = float(torch.as_tensor(sd["running_mean"]).reshape(-1)[0])
        self.state[1] = float(torch.as_tensor(sd["running_mean_sq"]).reshape(-1)[0])
        self.state[2] = float(torch.as_tensor(sd["debiasing_term"]))


# ---------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam stand-in, rMAPPOPolicy.py:31-37) over the flat vectors
# ---------------------------------------------------------------------------------------------
class FusedAdam:
    def __init__(self, net: DeviceNet, lr: float, eps: float, weight_decay: float = 0.0):
        if weight_decay:
            raise NotImplementedError("weight_decay != 0 (config.py:230 default 0; no reference script sets it)")
        self.net = net
        self.exp_avg = torch.zeros_like(net.flat)
        self.exp_avg_sq = torch.zeros_like(net.flat)
        self.step_dev = torch.zeros(2, dtype=torch.int32, device=net.device)      # {step, ticket}
        # {0.9^t, 0.999^t, t}: bias-correction powers cached by the kernel (tag -1 = empty; self-validating)
        self.beta_pow = torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64, device=net.device)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=net.device)
        self._lr_pinned = torch.empty(1, dtype=torch.float32).pin_memory()
        self.param_groups = [dict(lr=float(lr), eps=float(eps), betas=(0.9, 0.999), weight_decay=0.0,
                                  params=net.parameters())]
        self._lr_on_dev = float(lr)
        self.sumsq_part = torch.zeros((net.n_params + 31) // 32, dtype=torch.float32, device=net.device)

    def sync_lr(self):
        """lr_decay writes param_groups[0]['lr'] (utils/util.py:17-21); mirror it to the device scalar."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_dev:
            self._lr_pinned[0] = lr
            self.lr_dev.copy_(self._lr_pinned, non_blocking=True)
            self._lr_on_dev = lr

    def zero_grad(self):
        self.net.grad.zero_()

    def apply(self, max_grad_norm, use_max_grad_norm, grad_norm_out, n_sumsq_blocks=0):
        """clip_grad_norm_ + Adam on net.grad.  `grad_norm_out`: device double* accumulating the pre-clip norm.
        n_sumsq_blocks > 0: `sumsq_part` already holds that many partial sums of squares (from mappo_grad_reduce)."""
        lib = _lib.load()
        st = stream_ptr()
        if n_sumsq_blocks <= 0:
            nb = C.c_int32(0)
            check(lib.mappo_grad_sumsq(ptr(self.net.grad), self.net.n_params, ptr(self.sumsq_part), C.byref(nb), st))
            n_sumsq_blocks = nb.value
        check(lib.mappo_clip_adam(ptr(self.net.flat), ptr(self.net.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                  self.net.n_params, ptr(self.sumsq_part), int(n_sumsq_blocks), ptr(self.lr_dev),
                                  ptr(self.step_dev), float(self.param_groups[0]["eps"]), float(max_grad_norm),
                                  int(bool(use_max_grad_norm)), grad_norm_out, ptr(self.beta_pow), st))

    def step(self):
        self.sync_lr()
        self.apply(0.0, False, None)

    def state_dict(self):
        return dict(step=int(self.step_dev[0].item()), exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(),
                    lr=self.param_groups[0]["lr"])

    def load_state_dict(self, sd):
        self.step_dev.zero_()
        self.step_dev[0] = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0]["lr"] = sd["lr"]
        self.sync_lr()


# ---------------------------------------------------------------------------------------------
# update launch helper: one net, one minibatch
# ---------------------------------------------------------------------------------------------
class UpdateWorkspace:
    """Per-net scratch for mappo_update_fwd_bwd (gradient slots, recurrent activations)."""

    def __init__(self, net: DeviceNet, max_rows: int, gemm_mode: int = 0):
        lib = _lib.load()
        self.net = net
        self.max_rows = int(max_rows)
        self.gemm_mode = int(gemm_mode) if lib.mappo_tf32_supported(C.byref(net.desc)) else _lib.GEMM_FP32
        self.n_slots = int(lib.mappo_update_grad_slots(C.byref(net.desc), self.max_rows, self.gemm_mode))
        if self.n_slots <= 0:
            raise RuntimeError("mappo_update_grad_slots failed: " + lib.mappo_last_error().decode())
        self.slot_floats = int(lib.mappo_update_slot_floats(C.byref(net.desc), self.gemm_mode))
        self.grad_part = torch.empty(self.n_slots * self.slot_floats, dtype=torch.float32, device=net.device)
        wf = int(lib.mappo_update_workspace_floats(C.byref(net.desc), self.max_rows, self.gemm_mode))
        if wf < 0:
            raise RuntimeError("mappo_update_workspace_floats failed: " + lib.mappo_last_error().decode())
        self.workspace = torch.empty(max(wf, 1), dtype=torch.float32, device=net.device)
        self.sumsq_part = torch.zeros((net.n_params + 31) // 32, dtype=torch.float32, device=net.device)
        # hidden-64 tcgen05 path: the optimiser tail (slot sum, unfold, clip + Adam, next weight image) as one cluster launch
        # (mappo_update_tail).  image_ready: the workspace holds the image of the CURRENT weights -- callers clear it whenever
        # the parameters may have changed behind the tail's back (R_MAPPO.train() does at its start).
        self.fused_tail = (self.gemm_mode == _lib.GEMM_TF32 and not lib.mappo_big_net(C.byref(net.desc)) and not net.desc.recurrent
                           and os.environ.get("MAPPO_B200_FUSED_TAIL", "0") == "1")
        self.image_ready = 0


def _tail(net, ws, opt, stages, n_part, max_grad_norm, use_max_grad_norm, gn_ptr, grad, reducer=None, parity=0):
    lib = _lib.load()
    n_slots = ws._n_slots_last
    peers = (None, None, 0, 0, 0, None)
    if reducer is not None:
        peers = (reducer.bufs, reducer.sigs, reducer.world, reducer.rank, 4 * reducer.off_grad[parity & 1], ptr(reducer.round))
    check(lib.mappo_update_tail(C.byref(net.desc), ptr(net.flat), ptr(ws.grad_part), n_slots, ptr(grad),
                                None if opt is None else ptr(opt.exp_avg), None if opt is None else ptr(opt.exp_avg_sq),
                                ptr(ws.sumsq_part), int(n_part), None if opt is None else ptr(opt.lr_dev),
                                None if opt is None else ptr(opt.step_dev),
                                0.0 if opt is None else float(opt.param_groups[0]["eps"]), float(max_grad_norm),
                                int(bool(use_max_grad_norm)), gn_ptr, None if opt is None else ptr(opt.beta_pow),
                                ptr(ws.workspace), int(stages), *peers, stream_ptr()))


def make_loss_cfg(args, update_actor=True) -> LossCfg:
    c = LossCfg()
    c.clip_param, c.entropy_coef = float(args.clip_param), float(args.entropy_coef)
    c.value_loss_coef, c.huber_delta = float(args.value_loss_coef), float(args.huber_delta)
    c.use_clipped_value_loss = int(bool(args.use_clipped_value_loss))
    c.use_huber_loss = int(bool(args.use_huber_loss))
    c.use_value_active_masks = int(bool(args.use_value_active_masks))
    c.use_policy_active_masks = int(bool(args.use_policy_active_masks))
    c.use_valuenorm = int(bool(args.use_valuenorm or args.use_popart))
    c.update_actor = int(bool(update_actor))
    c.gemm_mode = _lib.GEMM_FP32
    return c


def launch_grads(net: DeviceNet, ws: UpdateWorkspace, batch: Batch, loss: LossCfg, norm_stats, adv_stats, vn_state,
                 loss_out, grad_out=None, finish=True):
    """forward + loss + backward -> slot reduction: leaves the (local) flat gradient in net.grad.
    Returns the number of sum-of-squares partials left in the optimiser's scratch (valid until an all-reduce)."""
    lib = _lib.load()
    st = stream_ptr()
    n_rows = int(batch.n_rows)
    n_slots = min(ws.n_slots, int(lib.mappo_update_grad_slots(C.byref(net.desc), n_rows, ws.gemm_mode)))
    loss.gemm_mode = ws.gemm_mode
    loss.image_ready = ws.image_ready if ws.fused_tail else 0
    ws.image_ready = 0                   # whatever follows changes the weights; only a completed fused tail re-arms it
    ws._n_slots_last = n_slots
    check(lib.mappo_update_fwd_bwd(C.byref(net.desc), ptr(net.flat), C.byref(batch), C.byref(loss), ptr(norm_stats),
                                   None if adv_stats is None else ptr(adv_stats),
                                   None if vn_state is None else ptr(vn_state), ptr(ws.grad_part), n_slots,
                                   ptr(loss_out), ptr(ws.workspace), st))
    if finish is False:
        return 0
    if ws.fused_tail:
        _tail(net, ws, None, 1, 0, 0.0, False, None, net.grad if grad_out is None else grad_out)
        return 12
    nb = C.c_int32(0)
    check(lib.mappo_update_finish(C.byref(net.desc), ptr(net.flat), ptr(ws.grad_part), n_slots, ws.gemm_mode,
                                  ptr(net.grad if grad_out is None else grad_out), ptr(ws.sumsq_part), C.byref(nb),
                                  ptr(ws.workspace), st))
    return nb.value


def launch_step(net: DeviceNet, ws: UpdateWorkspace, loss_out, opt: FusedAdam, max_grad_norm, use_max_grad_norm,
                grad_norm_slot: int, n_sumsq_blocks: int):
    """clip_grad_norm_ + Adam on net.grad (n_sumsq_blocks == 0: re-derive the norm, e.g. after an all-reduce)."""
    gn_ptr = C.c_void_p(loss_out.data_ptr() + 8 * grad_norm_slot)
    if ws.fused_tail and n_sumsq_blocks > 0:          # clip + Adam + the next step's weight image in one launch
        _tail(net, ws, opt, 2, n_sumsq_blocks, max_grad_norm, use_max_grad_norm, gn_ptr, net.grad)
        ws.image_ready = 1
        return
    if n_sumsq_blocks > 0:
        opt.sumsq_part = ws.sumsq_part
    opt.apply(max_grad_norm, use_max_grad_norm, gn_ptr, n_sumsq_blocks=n_sumsq_blocks)


def launch_update_p2p(net: DeviceNet, ws: UpdateWorkspace, batch: Batch, loss: LossCfg, norm_stats, adv_stats, vn_state,
                      loss_out, opt: FusedAdam, max_grad_norm, use_max_grad_norm, grad_norm_slot: int, reducer, parity: int):
    """Data-parallel optimiser step of one hidden-64 tcgen05 net as TWO launches: the update kernel, then the fused tail with
    the peer-memory exchange inside it (slot sum -> unfold into the symmetric buffer -> signals over NVLink -> sum of all
    ranks' gradients -> clip + Adam -> next weight image)."""
    assert ws.fused_tail
    launch_grads(net, ws, batch, loss, norm_stats, adv_stats, vn_state, loss_out, finish=False)
    _tail(net, ws, opt, 7, 0, max_grad_norm, use_max_grad_norm, C.c_void_p(loss_out.data_ptr() + 8 * grad_norm_slot),
          net.grad, reducer, parity)
    ws.image_ready = 1


def launch_update(net: DeviceNet, ws: UpdateWorkspace, batch: Batch, loss: LossCfg, norm_stats, adv_stats, vn_state,
                  loss_out, opt: FusedAdam, max_grad_norm, use_max_grad_norm, grad_norm_slot: int,
                  allreduce=None):
    """forward+loss+backward -> slot reduction -> [all-reduce] -> clip + Adam for one net."""
    if ws.fused_tail and allreduce is None:           # the whole tail as one cluster launch
        launch_grads(net, ws, batch, loss, norm_stats, adv_stats, vn_state, loss_out, finish=False)
        _tail(net, ws, opt, 3, 0, max_grad_norm, use_max_grad_norm,
              C.c_void_p(loss_out.data_ptr() + 8 * grad_norm_slot), net.grad)
        ws.image_ready = 1
        return
    nb = launch_grads(net, ws, batch, loss, norm_stats, adv_stats, vn_state, loss_out)
    if allreduce is not None:
        allreduce(net.grad)
        nb = 0                            # the norm must be re-derived from the all-reduced gradient
    launch_step(net, ws, loss_out, opt, max_grad_norm, use_max_grad_norm, grad_norm_slot, nb)
