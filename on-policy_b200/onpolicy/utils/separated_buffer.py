"""onpolicy.utils.separated_buffer.SeparatedReplayBuffer in HBM (reference: utils/separated_buffer.py:12-424).

Per-agent storage [T(+1), N, ...] is the M == 1 case of the shared layout, so every kernel is reused as is.
"""
import numpy as np
import torch

from mappo_b200.core import as_dev
from onpolicy.utils.shared_buffer import SharedReplayBuffer


class SeparatedReplayBuffer(SharedReplayBuffer):
    _SEPARATED = True

    def __init__(self, args, obs_space, share_obs_space, act_space, device=None):
        super().__init__(args, 1, obs_space, share_obs_space, act_space, device=device)
        self.factor = None

    def update_factor(self, factor):
        """reference :62-63 (HAPPO importance factor; carried, unused by MAPPO -- r_mappo.py:108-111).  Once set, every
        generator yields it as the 13th tuple element (reference :197-227), gathered with the same rows."""
        self.factor = as_dev(factor, self.device).clone()

    def recurrent_generator(self, advantages, num_mini_batch, data_chunk_length):
        # The reference stacks chunks on axis 0 here and then flattens them as if time-major
        # (separated_buffer.py:386-420): the time/chunk axes come out scrambled.  That path is
        # not reachable from any BASELINE config; the coherent shared layout is provided instead.
        return super().recurrent_generator(advantages, num_mini_batch, data_chunk_length)
