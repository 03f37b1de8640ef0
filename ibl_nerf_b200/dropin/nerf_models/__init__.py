"""Drop-in `nerf_models` package: put ibl_nerf_b200/dropin ahead of the reference's src/ on sys.path and
src/train.py / src/test.py run unchanged on the B200-native kernels."""
