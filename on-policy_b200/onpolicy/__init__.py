"""Drop-in `onpolicy` package: the module paths and class names of marlbenchmark/on-policy's MAPPO hot path
(SURVEY.md section 8b), implemented on libmappo_b200.so (B200, sm_100a).  Put `on-policy_b200/` ahead of the
reference on PYTHONPATH; environments, config and scripts keep coming from the reference tree."""
import os as _os
from pkgutil import extend_path as _extend_path

# let `onpolicy.envs`, `onpolicy.config`, `onpolicy.scripts` ... resolve to a reference checkout further down
# sys.path: only the hot-path modules are provided here.
__path__ = _extend_path(__path__, __name__)
