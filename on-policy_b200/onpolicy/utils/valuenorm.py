"""onpolicy.utils.valuenorm.ValueNorm on the device (reference: utils/valuenorm.py:8-79)."""
from mappo_b200.core import DeviceValueNorm


class ValueNorm(DeviceValueNorm):
    def __init__(self, input_shape, norm_axes=1, beta=0.99999, per_element_update=False, epsilon=1e-5, device=None):
        if input_shape != 1 or norm_axes != 1 or per_element_update or beta != 0.99999 or epsilon != 1e-5:
            raise NotImplementedError("only the configuration R_MAPPO constructs (ValueNorm(1), r_mappo.py:48) is built")
        super().__init__(input_shape, device=device)
