"""nerf_models.envmap: not on the hot path (only ever constructed and checkpointed, never sampled by the renderer),
so the reference's own file is executed under this name."""
from ibl_nerf_b200.factory import load_reference_module as _load

_ref = _load("nerf_models/envmap.py", "ibl_nerf_b200._reference_envmap")
EnvironmentMap = _ref.EnvironmentMap
direction_to_canonical = _ref.direction_to_canonical
