"""algorithms/utils/util.py:16-18."""
import numpy as np
import torch


def check(value):
    return torch.from_numpy(value) if type(value) == np.ndarray else value
