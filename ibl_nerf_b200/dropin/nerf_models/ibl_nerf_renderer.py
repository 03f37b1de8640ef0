from nerf_models.nerf_renderer_helper import *          # noqa: F401,F403  (same star-import surface as the reference)
import os                                                # noqa: F401
import time                                              # noqa: F401
from ibl_nerf_b200.renderer import (render_decomp, render_decomp_path, render_rays, raw2outputs, raw2outputs_simple,  # noqa: F401
                                    raw2outputs_depth, batchify_rays, rgb_to_srgb, tonemap_reinherd, gamma, epsilon_srgb)
from nerf_models.microfacet import fresnel_schlick_roughness   # noqa: F401
