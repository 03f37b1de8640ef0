import os                                                # noqa: F401
import numpy as np                                       # noqa: F401
import torch                                             # noqa: F401
import torch.nn as nn                                    # noqa: F401
import torch.nn.functional as F                          # noqa: F401
from ibl_nerf_b200.model import IBLNeRF, run_network, batchify, NetworkQuery   # noqa: F401
from ibl_nerf_b200.factory import create_IBLNeRF                               # noqa: F401
from nerf_models.positional_embedder import get_embedder                       # noqa: F401
