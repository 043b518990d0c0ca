"""onpolicy.algorithms.happo.happo_trainer.HAPPO on libmappo_b200 (reference: algorithms/happo/happo_trainer.py:9-232).

HAPPO is MAPPO's update with three differences, all switches of the SAME kernels:
  * actor loss (:129-141): one importance weight per row -- the product over the action heads of exp(logp - old_logp) -- and
    the row's `factor` (the running product of the previously trained agents' ratios, separated_buffer.py:62-63, handed over
    by the separated runner) multiplies min(surr1, surr2): `mappo_loss_cfg_t.happo` + `mappo_batch_t.factor`
    (csrc/net_tiles.cuh row_loss_pre);
  * cal_value_loss (:42-84) normalises the returns with the ValueNorm state AS IS -- it never calls `update`;
  * train() (:181-184) subtracts the raw (normalised) value predictions from the returns when forming advantages
    (denormalisation only under use_popart).
Both quirks of the last two bullets are the reference's behaviour and are reproduced, not repaired."""
import torch

from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO


class HAPPO(R_MAPPO):
    def __init__(self, args, policy, device=torch.device("cpu")):
        super().__init__(args, policy, device=device)
        self._happo = True
