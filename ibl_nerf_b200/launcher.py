"""Run the reference's own drivers (src/train.py, src/test.py) unchanged on the B200-native hot path.

    python -m ibl_nerf_b200.launcher /path/to/IBL-NeRF/src train --config ../configs/IBL-NeRF/kitchen/IBL-NeRF.txt

install() puts the drop-in `nerf_models` package ahead of the reference's src/ on sys.path, adds stand-ins for the
Python dependencies that are missing in this image (only if the real ones are absent) and neutralises the
import-time side effect of miscellaneous/test_dataset_speed.py (it sets CUDA_VISIBLE_DEVICES=5).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def install(reference_src):
    reference_src = os.path.abspath(reference_src)
    shims = os.path.join(HERE, "shims")
    for mod in ("configargparse", "imageio", "matplotlib"):
        try:
            importlib.import_module(mod)
        except ImportError:
            if shims not in sys.path:
                sys.path.append(shims)          # after site-packages: real packages win when present
            importlib.import_module(mod)
    for p in (reference_src, os.path.join(HERE, "dropin"), os.path.dirname(HERE)):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, reference_src)
    sys.path.insert(0, os.path.join(HERE, "dropin"))    # shadows src/nerf_models
    sys.path.insert(0, os.path.dirname(HERE))
    for name in list(sys.modules):
        if name == "nerf_models" or name.startswith("nerf_models."):
            del sys.modules[name]
    from . import factory
    factory.REFERENCE_SRC = reference_src           # create_IBLNeRF runs the reference's own control-plane code from here
    # harmless replacement of the scratch module that train.py / test.py star-import (train.py:21)
    stub = types.ModuleType("miscellaneous.test_dataset_speed")
    stub.__all__ = []
    sys.modules.setdefault("miscellaneous.test_dataset_speed", stub)
    return reference_src


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2 or argv[1] not in ("train", "test"):
        raise SystemExit("usage: python -m ibl_nerf_b200.launcher <reference src dir> train|test [reference options]")
    src = install(argv[0])
    os.chdir(src)                                    # the reference uses paths relative to src/ (../data, ../configs)
    sys.argv = [argv[1] + ".py"] + argv[2:]
    mod = importlib.import_module(argv[1])
    if os.environ.get("IBLN_DEVICE_SAMPLER", "1") != "0" and hasattr(mod, "sample_generator_single_image"):
        # the driver star-imported the host-numpy generator (train.py:23): rebind it to the device one
        from . import sampling
        mod.sample_generator_single_image = sampling.sample_generator_single_image
    args = prepare_args(mod, mod.recursive_config_parser().parse_args())
    getattr(mod, argv[1])(args)


def prepare_args(mod, args):
    """The prologue of the drivers' own `__main__` blocks (train.py:530-541, test.py:160-174), which importing
    the module does not execute: the device, the experiment name derived from the config file name, and (test.py)
    the export directory."""
    args.device = mod.device
    if args.expname is None:
        args.expname = args.config.split("/")[-1].split(".")[0]
    if hasattr(args, "export_basedir") and args.export_basedir is None and mod.__name__ == "test":
        args.export_basedir = args.basedir.replace("logs", "logs_eval")
    return args


if __name__ == "__main__":
    main()
