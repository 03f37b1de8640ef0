import torch                                             # noqa: F401
from ibl_nerf_b200.mlp import get_embedder, Embedder     # noqa: F401
