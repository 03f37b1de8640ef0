"""create_IBLNeRF: drop-in for reference nerf_models/ibl_nerf.py:255-428 (models, Adam, checkpoint
reload, render kwargs).  Same signature, same return tuple, same kwargs keys, same checkpoint keys."""
import logging
import os

import torch

from .mlp import get_embedder
from .model import IBLNeRF, NetworkQuery


def _logger(name):
    try:
        from utils.logging_utils import load_logger     # reference utility when running as a drop-in
        return load_logger(name)
    except Exception:
        return logging.getLogger(name)


class EnvironmentMap(torch.nn.Module):
    """envmap.py: only ever constructed and checkpointed (never sampled by the renderer)."""

    def __init__(self, n=16):
        super().__init__()
        self.emission = torch.nn.Parameter(torch.ones(1, 3, n, 2 * n))


def _aux(kind, **kw):
    from networks.MLP import PositionMLP, PositionDirectionMLP   # reference aux MLPs, opaque nn.Modules (off in shipped configs)
    return {"pos": PositionMLP, "posdir": PositionDirectionMLP}[kind](**kw)


def create_IBLNeRF(args):
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    logger = _logger("IBL-NeRF Loader")
    skips = [4]
    mk = lambda: IBLNeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, input_ch_views=input_ch_views, skips=skips,
                         use_illumination_feature_layer=args.use_illumination_feature_layer,
                         use_instance_feature_layer=args.use_instance_feature_layer,
                         coarse_radiance_number=args.coarse_radiance_number,
                         is_color_independent_to_direction=args.color_independent_to_direction).to(args.device)
    model = mk()
    logger.info(model)
    grad_vars = [{'params': model.parameters(), 'name': 'coarse'}]
    model_fine = None
    if args.N_importance > 0:
        model_fine = mk()
        logger.info("NeRFDecomp fine model")
        logger.info(model_fine)
        grad_vars.append({'params': model_fine.parameters(), 'name': 'fine'})

    common = dict(D=args.netdepth, W=args.netwidth, input_ch=input_ch, skips=skips)
    depth_mlp = visibility_mlp = normal_mlp = albedo_mlp = roughness_mlp = irradiance_mlp = None
    if args.infer_depth:
        depth_mlp = _aux("posdir", input_ch_views=input_ch_views, out_ch=1, **common)
        grad_vars.append({'params': depth_mlp.parameters(), 'name': 'depth_mlp'})
    if args.infer_visibility:
        visibility_mlp = _aux("posdir", input_ch_views=input_ch_views, out_ch=1, **common)
        grad_vars.append({'params': depth_mlp.parameters(), 'name': 'visibility_mlp'})      # sic: ibl_nerf.py:304
    if args.infer_normal:
        normal_mlp = _aux("pos", out_ch=3, **common)
        grad_vars.append({'params': normal_mlp.parameters(), 'name': 'normal_mlp'})
    if args.infer_albedo_separate:
        albedo_mlp = _aux("pos", out_ch=3, **common)
        grad_vars.append({'params': albedo_mlp.parameters(), 'name': 'albedo_mlp'})
    if args.infer_roughness_separate:
        roughness_mlp = _aux("pos", out_ch=1, **common)
        grad_vars.append({'params': roughness_mlp.parameters(), 'name': 'roughness_mlp'})
    if args.infer_irradiance_separate:
        irradiance_mlp = _aux("pos", out_ch=1, **common)
        grad_vars.append({'params': irradiance_mlp.parameters(), 'name': 'irradiance_mlp'})

    network_query_fn = NetworkQuery(embed_fn, embeddirs_fn, args.netchunk)

    env_map = None
    if args.use_environment_map:
        env_map = EnvironmentMap(n=args.N_envmap_size)
        grad_vars.append({'params': env_map.emission, 'name': 'env_map', 'lr': args.lrate_env_map})

    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start, elapsed_time = 0, 0
    basedir, expname = args.basedir, args.expname
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    elif args.target_load_N_iter > 0:
        ckpts = [os.path.join(basedir, expname, '{:06d}.tar'.format(args.target_load_N_iter))]
    else:
        ckpts = [os.path.join(basedir, expname, f) for f in sorted(os.listdir(os.path.join(basedir, expname))) if 'tar' in f]
    logger.info('Found ckpts: ' + str(ckpts))
    if len(ckpts) > 0 and not args.no_reload:
        ckpt_path = ckpts[-1]
        logger.info('Reloading from ' + str(ckpt_path))
        ckpt = torch.load(ckpt_path, weights_only=False)
        start = ckpt['global_step']
        elapsed_time = ckpt.get('elapsed_time', 0)
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if args.infer_depth:
            depth_mlp.load_state_dict(ckpt['depth_mlp'])
        if args.infer_normal:
            normal_mlp.load_state_dict(ckpt['normal_mlp'])
        if args.infer_albedo_separate and 'albedo_mlp' in ckpt:
            albedo_mlp.load_state_dict(ckpt['albedo_mlp'])
        if args.infer_roughness_separate and 'roughness_mlp' in ckpt:
            roughness_mlp.load_state_dict(ckpt['roughness_mlp'])
        if args.infer_irradiance_separate and 'irradiance_mlp' in ckpt:
            irradiance_mlp.load_state_dict(ckpt['irradiance_mlp'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])
        if args.use_environment_map:
            env_map.emission.data = ckpt['env_map']

    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
        'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': model, 'use_viewdirs': args.use_viewdirs,
        'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std, 'ndc': False, 'lindisp': args.lindisp,
        "depth_mlp": depth_mlp, "visibility_mlp": visibility_mlp, "normal_mlp": normal_mlp, "albedo_mlp": albedo_mlp,
        "roughness_mlp": roughness_mlp, "irradiance_mlp": irradiance_mlp, "infer_depth": args.infer_depth,
        "infer_visibility": args.infer_visibility, "infer_normal": args.infer_normal,
        "infer_normal_at_surface": args.infer_normal_at_surface, "coarse_radiance_number": args.coarse_radiance_number,
        "use_monte_carlo_integration": args.use_monte_carlo_integration,
        "use_gradient_for_incident_radiance": args.use_gradient_for_incident_radiance,
        "use_radiance_linear": args.use_radiance_linear, "gamma_correct": args.gamma_correct,
        "monte_carlo_integration_method": args.monte_carlo_integration_method,
        'use_environment_map': args.use_environment_map, "env_map": env_map, "lut_coefficient": args.lut_coefficient,
        "depth_map_from_ground_truth": args.depth_map_from_ground_truth,
        "target_normal_map_for_radiance_calculation": args.calculating_normal_type,
        "calculate_albedo_from_gt": args.calculate_albedo_from_gt,
        "calculate_roughness_from_gt": args.calculate_roughness_from_gt,
        "calculate_irradiance_from_gt": args.calculate_irradiance_from_gt,
        "epsilon": args.epsilon_for_numerical_normal, "epsilon_direction": args.epsilon_direction_for_numerical_normal,
        "N_hemisphere_sample_sqrt": args.N_hemisphere_sample_sqrt, "roughness_exp_coefficient": args.roughness_exp_coefficient,
        "albedo_multiplier": args.albedo_multiplier,
        "correct_depth_for_prefiltered_radiance_infer": args.correct_depth_for_prefiltered_radiance_infer,
    }
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0
    return render_kwargs_train, render_kwargs_test, start, elapsed_time, grad_vars, optimizer
