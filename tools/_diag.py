"""Tuning tools import this first: it points the package at the diagnostics build of the library
(libiblnerf_b200_diag.so = same sources + -DIBLN_DIAGNOSTICS; `python -m ibl_nerf_b200.build --diag`)."""
import os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_p = os.path.join(R, "ibl_nerf_b200", "libiblnerf_b200_diag.so")
if "IBLN_LIB" not in os.environ:
    if not os.path.exists(_p):
        raise SystemExit("diagnostics library missing: run `python -m ibl_nerf_b200.build --diag` first")
    os.environ["IBLN_LIB"] = _p
