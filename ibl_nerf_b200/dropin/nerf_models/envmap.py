from ibl_nerf_b200.factory import EnvironmentMap         # noqa: F401
