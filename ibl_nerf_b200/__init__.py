"""iblnerf-b200: B200-native implementation of IBL-NeRF's per-ray volumetric hot path.

Python host code over a C-ABI CUDA library (libiblnerf_b200.so, sm_100a).  Import is cheap and works
without a GPU; the first kernel call loads the library and fails loudly if it is missing.
"""
from . import _lib, ops, mlp, model, renderer, helper, factory          # noqa: F401
from .model import IBLNeRF, NetworkQuery, run_network, batchify         # noqa: F401
from .mlp import get_embedder, set_default_precision                    # noqa: F401
from .renderer import (render_decomp, render_decomp_path, render_rays, raw2outputs, raw2outputs_simple,  # noqa: F401
                       raw2outputs_depth, batchify_rays, rgb_to_srgb)
from .helper import sample_pdf                                          # noqa: F401
from .factory import create_IBLNeRF                                     # noqa: F401

__version__ = "0.1.0"
