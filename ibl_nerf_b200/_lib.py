"""ctypes binding of libiblnerf_b200.so (the C ABI declared in include/iblnerf_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# IBLN_LIB selects another build of the same ABI (tools/ use it for the diagnostics build, libiblnerf_b200_diag.so)
LIB_PATH = os.environ.get("IBLN_LIB") or os.path.join(_HERE, "libiblnerf_b200.so")

c_int, c_i64, c_f, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes (device + stream are appended automatically)
_SIGS = {
    "ibln_stratified_z": [c_p, c_p, c_p, c_int, c_int, c_int, c_p],
    "ibln_sample_pdf": [c_p, c_i64, c_p, c_i64, c_p, c_int, c_int, c_int, c_p],
    "ibln_inverse_cdf": [c_p, c_p, c_p, c_int, c_int, c_int, c_p, c_p],
    "ibln_hierarchical_sample": [c_p, c_p, c_p, c_int, c_int, c_int, c_p, c_p],
    "ibln_merge_sort_z": [c_p, c_p, c_int, c_int, c_int, c_p],
    "ibln_composite_fwd": [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p],
    "ibln_composite_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "ibln_composite_simple_fwd": [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p],
    "ibln_depth_fwd": [c_p, c_p, c_p, c_int, c_int, c_int, c_p, c_p, c_p],
    "ibln_normal_eps_points": [c_p, c_p, c_p, c_int, c_int, c_f, c_p],
    "ibln_normal_eps_finish": [c_p, c_p, c_int, c_f, c_p, c_p, c_p, c_p, c_int, c_p],
    "ibln_shade_fwd_maps": [c_p] * 6 + [c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p],
    "ibln_shade_bwd_maps": [c_p] * 6 + [c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p],
    "ibln_shade_fwd": [c_p] * 10 + [c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p],
    "ibln_shade_bwd": [c_p] * 10 + [c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_p],
    "ibln_encode": [c_p, c_i64, c_int, c_p, c_i64],
    "ibln_encode_dirs": [c_p, c_i64, c_int, c_int, c_p, c_i64],
    "ibln_sgemm": [c_p, c_i64, c_p, c_i64, c_int, c_p, c_p, c_i64, c_i64, c_int, c_int, c_int, c_int, c_p, c_i64],
    "ibln_sgemm_wgrad": [c_p, c_i64, c_p, c_i64, c_i64, c_int, c_int, c_p, c_i64, c_p, c_int, c_p],
    "ibln_mlp_pack_weights": [c_p, c_p],
    "ibln_mlp_fwd": [c_p, c_int, c_p, c_p, c_p, c_p, c_i64, c_int, c_f, c_int, c_p, c_p],
    "ibln_mlp_bwd": [c_p, c_p, c_p, c_i64, c_p, c_p, c_int],
    "ibln_image_losses": [c_p] * 7 + [c_int] + [c_f] * 6 + [c_p, c_p, c_p],
    "ibln_sample_rays": [c_p, c_p, c_int, c_int, c_int, c_f, c_f, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_int],
    "ibln_pack_u8": [c_p, c_p, c_p, c_p, c_int, c_p],
    "ibln_depth_to_normal": [c_p, c_int, c_int, c_f, c_f, c_f, c_f, c_p, c_p],
    "ibln_adam_step": [c_p, c_p, c_p, c_p, c_i64, c_f, c_f, c_f, c_f, c_int, c_f],
    "ibln_adam_step_pack": [c_p, c_p, c_p, c_p, c_int, c_f, c_f, c_f, c_f, c_int, c_f, c_p],
    "ibln_adam_allreduce_step": [c_p, c_p, c_p, c_int, c_p, c_p, c_i64, c_f, c_f, c_f, c_f, c_int, c_f],
    "ibln_adam_allreduce_step_pack": [c_p, c_p, c_p, c_int, c_p, c_p, c_int, c_f, c_f, c_f, c_f, c_int, c_f, c_p],
    "ibln_zero": [c_p, c_i64],
    "ibln_umma_selftest": [c_p, c_p, c_p, c_int, c_int, c_int],
    "ibln_umma_mn_selftest": [c_p, c_p, c_p, c_int],
    "ibln_umma_pair_selftest": [c_p, c_p, c_p, c_int],
}
# include/iblnerf_b200_diag.h: bound only when the loaded library is the diagnostics build
_SIGS_DIAG = {
    "ibln_store_probe": [c_p, c_i64, c_int, c_int],
    "ibln_tmem_probe": [c_p, c_int, c_int, c_int],
}
_PLAIN_DIAG = {
    "ibln_debug_set": ([c_int], c_int),
    "ibln_debug_timeline": ([c_p], c_int),
}
_PLAIN = {  # no device/stream tail
    "ibln_abi_version": ([], c_int),
    "ibln_error_string": ([c_int], ctypes.c_char_p),
    "ibln_wgrad_workspace_bytes": ([c_int, c_int], c_i64),
    "ibln_mlp_packed_bytes": ([], c_i64),
    "ibln_mlp_saved_bytes": ([c_i64], c_i64),
    "ibln_mlp_bwd_workspace_bytes": ([c_i64], c_i64),
}

ABI_VERSION = 4      # include/iblnerf_b200.h: IBLN_ABI_VERSION
_lib = None
# kernels launched per entry point (for bench.py's gpu_launches); default 1
KERNELS_PER_CALL = {"ibln_sgemm_wgrad": 2, "ibln_mlp_pack_weights": 3, "ibln_mlp_bwd": 2, "ibln_adam_step_pack": 2, "ibln_adam_allreduce_step_pack": 2, "ibln_zero": 0}
# bench.py sets this to {} to collect per-entry launch counts, CUDA-event pairs and algorithmic FLOPs;
# PROFILE_EVENTS (None = every entry) limits the CUDA-event bracketing to the named entry points
PROFILE = None
PROFILE_EVENTS = None


class IblnError(RuntimeError):
    pass


def exported_names():
    return sorted(list(_SIGS) + list(_PLAIN))


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IblnError("%s not found: build it with `python -m ibl_nerf_b200.build` (nvcc, sm_100a). "
                            "There is no CPU fallback." % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        h.ibln_abi_version.restype = c_int
        if h.ibln_abi_version() != ABI_VERSION:
            raise IblnError("%s has ABI version %d, this package binds version %d: rebuild it (python -m ibl_nerf_b200.build --force)"
                            % (LIB_PATH, h.ibln_abi_version(), ABI_VERSION))
        for name, at in _SIGS.items():
            fn = getattr(h, name)
            fn.argtypes = at + [c_int, c_p]
            fn.restype = c_int
        for name, (at, rt) in _PLAIN.items():
            fn = getattr(h, name)
            fn.argtypes = at
            fn.restype = rt
        if hasattr(h, "ibln_debug_set"):
            for name, at in _SIGS_DIAG.items():
                fn = getattr(h, name)
                fn.argtypes = at + [c_int, c_p]
                fn.restype = c_int
            for name, (at, rt) in _PLAIN_DIAG.items():
                fn = getattr(h, name)
                fn.argtypes = at
                fn.restype = rt
        _lib = h
    return _lib


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise IblnError("iblnerf_b200 kernels need CUDA tensors (got %s); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise IblnError("non-contiguous tensor passed to a kernel")
    return c_p(t.data_ptr())


def call(name, device, *args, flops=0.0):
    """Invoke an entry point on torch's current stream of `device`; raise on a non-zero status."""
    h = lib()
    dev = device.index if device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    if PROFILE is not None:
        rec = PROFILE.setdefault(name, {"launches": 0, "events": [], "flops": 0.0})
        rec["launches"] += KERNELS_PER_CALL.get(name, 1)
        rec["flops"] += flops
        if PROFILE_EVENTS is None or name in PROFILE_EVENTS:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(dev))
            rc = getattr(h, name)(*args, dev, c_p(stream))
            e1.record(torch.cuda.current_stream(dev))
            rec["events"].append((e0, e1))
        else:
            rc = getattr(h, name)(*args, dev, c_p(stream))
    else:
        rc = getattr(h, name)(*args, dev, c_p(stream))
    if rc != 0:
        raise IblnError("%s failed: %s (%d)" % (name, h.ibln_error_string(rc).decode(), rc))


def f32c(t):
    """Contiguous fp32 view/copy."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
