import torch                                             # noqa: F401
import torch.nn as nn                                    # noqa: F401
import torch.nn.functional as F                          # noqa: F401
import numpy as np                                       # noqa: F401
from ibl_nerf_b200.helper import (img2mse, mse2psnr, to8b, get_rays, get_rays_few, get_rays_patch_few, get_rays_np,  # noqa: F401
                                  get_rays_few_np, sample_pdf)
