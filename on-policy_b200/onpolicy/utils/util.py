"""Host helpers with the reference's names (onpolicy/utils/util.py): lr schedule and space shapes.
The loss helpers (huber_loss / mse_loss, :23-29) and get_gard_norm (:9-15) live inside the CUDA kernels."""
import numpy as np
import torch


def check(value):
    """utils/util.py:5-7 returns None for non-ndarray input; the trainer uses the variant of
    algorithms/utils/util.py:16-18 (ndarray -> tensor, anything else unchanged), which is what this is."""
    return torch.from_numpy(value) if type(value) == np.ndarray else value


def update_linear_schedule(optimizer, epoch, total_num_epochs, initial_lr):
    """utils/util.py:17-21."""
    lr = initial_lr - (initial_lr * (epoch / float(total_num_epochs)))
    for group in optimizer.param_groups:
        group["lr"] = lr


def get_shape_from_obs_space(obs_space):
    kind = obs_space.__class__.__name__
    if kind == "Box":
        return obs_space.shape
    if kind == "list":
        return obs_space
    raise NotImplementedError


def get_shape_from_act_space(act_space):
    kind = act_space.__class__.__name__
    if kind == "Discrete":
        return 1
    if kind == "MultiDiscrete":
        return act_space.shape
    if kind in ("Box", "MultiBinary"):
        return act_space.shape[0]
    return act_space[0].shape[0] + 1
