"""create_IBLNeRF for the drop-in `nerf_models.ibl_nerf` module.

Model / optimizer construction, checkpoint reload and the render-kwargs schema are control-plane code outside the hot
path (out of scope), so this package does not restate them: the reference's OWN `create_IBLNeRF`
(nerf_models/ibl_nerf.py:255-428) runs, from the reference checkout the launcher was given, with exactly three names
substituted in its module namespace -- `IBLNeRF`, `run_network`, `batchify` (the hot-path types this package
provides; `get_embedder` already resolves to the drop-in) -- and afterwards the anonymous `network_query_fn` lambda
(:327-329) is replaced by the equivalent `NetworkQuery` object so the renderer can recognise it and use the fused
ray-march kernels.  Same signature, same return tuple, same kwargs keys and checkpoint keys by construction.
"""
import importlib.util
import os
import sys

from ._lib import IblnError
from .mlp import get_embedder
from .model import IBLNeRF, NetworkQuery, batchify, run_network

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = os.environ.get("IBLN_REFERENCE_SRC")        # launcher.install() sets this
_loaded = {}


def reference_file(relpath):
    """Path of `relpath` (e.g. 'nerf_models/ibl_nerf.py') inside the reference's src/ directory."""
    roots = [REFERENCE_SRC] if REFERENCE_SRC else []
    roots += [p for p in sys.path if p and not os.path.abspath(p).startswith(_HERE)]
    for root in roots:
        cand = os.path.join(root, relpath)
        if os.path.isfile(cand) and os.path.isfile(os.path.join(root, "config_parser.py")):
            return cand
    raise IblnError("create_IBLNeRF / EnvironmentMap are the reference's own control-plane code: run through "
                    "`python -m ibl_nerf_b200.launcher <IBL-NeRF/src> ...` or set IBLN_REFERENCE_SRC to the reference's "
                    "src/ directory (looked for %s)" % relpath)


def load_reference_module(relpath, alias, substitutions=None):
    """Execute a reference source file under a private module name (it is shadowed by the drop-in package under its
    own name) and overwrite the given globals afterwards."""
    path = reference_file(relpath)
    key = (path, alias)
    if key not in _loaded:
        spec = importlib.util.spec_from_file_location(alias, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for name, obj in (substitutions or {}).items():
            setattr(mod, name, obj)
        _loaded[key] = mod
    return _loaded[key]


def create_IBLNeRF(args):
    ref = load_reference_module("nerf_models/ibl_nerf.py", "ibl_nerf_b200._reference_ibl_nerf",
                                dict(IBLNeRF=IBLNeRF, run_network=run_network, batchify=batchify))
    out = ref.create_IBLNeRF(args)
    query = NetworkQuery(get_embedder(args.multires, args.i_embed)[0], get_embedder(args.multires_views, args.i_embed)[0],
                         args.netchunk)
    for kw in (out[0], out[1]):
        kw["network_query_fn"] = query
    use_fused_adam(out[-1])
    return out


def use_fused_adam(optimizer):
    """Switch the torch.optim.Adam the reference built (ibl_nerf.py:336) to torch's fused implementation, in place.

    The drivers make CUDA the default tensor type (train.py:76, test.py:76), so the default implementation creates its
    per-parameter `step` counters on the device and reads every one of them back with `.item()` twice per
    `optimizer.step()`: 184 stream synchronisations = ~4 ms of a 14.7 ms iteration with the GPU idle (torch.profiler,
    tools/prof_dropin.py).  The fused implementation keeps the counters on the device.  Same class, same state_dict
    keys, same update rule; the learning-rate writes of train.py:483-498 into `param_groups` keep working."""
    import torch
    if not isinstance(optimizer, torch.optim.Adam):
        return optimizer
    params = [p for g in optimizer.param_groups for p in g["params"]]
    if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
        return optimizer
    for g in optimizer.param_groups:
        g["fused"], g["foreach"] = True, False
    optimizer.defaults["fused"], optimizer.defaults["foreach"] = True, False
    for p, st in optimizer.state.items():           # counters restored from a checkpoint (ibl_nerf.py:362)
        if "step" in st:
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=p.device)
    return optimizer
