"""Usage:
    wayne -p <parameter_file> [--gpus <n>] [--max-exposures <k>]

Runs a visit from a YAML parameter file -- the ``wayne`` command of the
reference (wayne/run_visit.py:1-324, setup.py:56-60) with the same file format
(sections general / target / observation / trends, e.g. the reference's
examples/hd209458b_12181_simulation_parameters.yml).  Keys this build adds, all
optional: ``target.stellar_temperature`` (black-body star when no stellar
spectrum file is given; the reference reads it from the Open Exoplanet
Catalogue, which is not available), ``general.rng`` ('philox' | 'numpy').

Exposures are generated on the GPU(s) by wayne_b200.observation.Observation;
with ``--gpus n`` (one child process per GPU, spawned here) or under torchrun the
visit is partitioned exposure-wise.  A visit without ``general.seed`` gets ONE
seed for all ranks (drawn by the parent / by rank 0), so the frames and the
RANDSEED headers do not depend on the GPU count.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

import numpy as np
import yaml

from . import detector, grism, observation, params, tools
from . import units as u
from .trend_generators import scan_speed_varations


class WFC3SimConfigError(Exception):
    pass


def _get(cfg, section, key, default=None):
    try:
        v = cfg[section][key]
    except KeyError:
        return default
    return v


def shared_visit_seed(seed=None):
    """The visit's seed, identical on every rank: the parameter file's, else the one
    the ``--gpus`` parent handed down (WAYNE_B200_VISIT_SEED), else drawn by rank 0 and
    broadcast (a two-line gloo exchange under torchrun; the data path has no collective)."""
    if seed:
        return int(seed)
    env = os.environ.get('WAYNE_B200_VISIT_SEED')
    if env:
        return int(env)
    s = int.from_bytes(os.urandom(4), 'little') & 0x7fffffff
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and os.environ.get('MASTER_ADDR'):
        import torch
        import torch.distributed as dist
        made = not dist.is_initialized()
        if made:
            dist.init_process_group('gloo')
        t = torch.tensor([s], dtype=torch.int64)
        dist.broadcast(t, 0)
        s = int(t.item())
        if made:
            dist.destroy_process_group()
    return s


def build_observation(cfg, base_dir='.'):
    """Wire an Observation from a parsed parameter file (run_visit.py:47-314)."""
    def path(p):
        return p if os.path.isabs(p) else os.path.join(base_dir, p)

    outdir = path(cfg['general']['outdir'])
    os.makedirs(outdir, exist_ok=True)
    seed = shared_visit_seed(_get(cfg, 'general', 'seed'))
    np.random.seed(seed)
    params.seed = seed
    if _get(cfg, 'general', 'rng'):
        params.rng = cfg['general']['rng']

    grisms = {'G141': grism.G141, 'G102': grism.G102}
    chosen_grism = grisms[cfg['observation']['grism']]()
    det = detector.WFC3_IR()
    rebin_resolution = _get(cfg, 'target', 'rebin_resolution')
    lo, hi = (float(u.value_in(v, u.micron)) for v in chosen_grism.wl_limits)

    planet_file = _get(cfg, 'target', 'planet_spectrum_file')
    name = cfg['target']['name']
    tgt = {k: _get(cfg, 'target', k) for k in ('transit_time', 'period', 'rp', 'sma', 'stellar_radius',
                                               'inclination', 'eccentricity', 'ldcoeffs', 'periastron')}
    planet = observation.Planet(name=name)
    if planet_file:
        wl_planet, depth_planet = tools.load_and_sort_spectrum(path(planet_file))
        wl_planet, depth_planet = (np.array(a) for a in tools.crop_spectrum(0.9, 1.8, wl_planet, depth_planet))
        if rebin_resolution:
            new_wl = tools.wl_at_resolution(rebin_resolution, lo, hi)
            depth_planet = tools.rebin_spec(wl_planet, depth_planet, new_wl)
            wl_planet = new_wl
    else:
        wl_planet = depth_planet = None

    stellar_file = _get(cfg, 'target', 'stellar_spectrum_file')
    if stellar_file and os.path.isfile(path(stellar_file)):
        wl_star, flux_star = tools.load_pheonix_stellar_grid_fits(path(stellar_file))
        if planet_file:
            flux_star = tools.rebin_spec(wl_star, flux_star, wl_planet)
        elif rebin_resolution:
            new_wl = tools.wl_at_resolution(rebin_resolution, lo, hi)
            flux_star = tools.rebin_spec(wl_star, flux_star, new_wl)
            wl_star = new_wl
    else:
        temperature = _get(cfg, 'target', 'stellar_temperature')
        if not temperature:
            raise WFC3SimConfigError(
                "stellar spectrum file '{}' not found and target.stellar_temperature not given"
                "".format(stellar_file))
        wl_star = wl_planet if planet_file else tools.wl_at_resolution(rebin_resolution or 130 * 40, lo, hi)
        flux_star = tools.blackbody_lambda(wl_star, float(temperature))
    stellar_flux_scaled = np.asarray(flux_star) * cfg['target']['flux_scale']
    wl = (wl_planet if planet_file else wl_star) * u.micron

    ob = cfg['observation']

    def maybe_file(v):
        return np.loadtxt(path(v)) if isinstance(v, str) else v

    x_ref, y_ref = maybe_file(ob['x_ref']), maybe_file(ob['y_ref'])
    sky = maybe_file(ob['sky_background']) * u.count / u.s
    start_JD = (ob['start_JD'] or 0) * u.day
    exp_start_times = ob.get('exp_start_times', False)
    if exp_start_times:
        exp_start_times = np.loadtxt(path(exp_start_times)) * u.day
    spatial_scan = ob['spatial_scan']
    sample_rate = ob['sample_rate'] * u.ms if spatial_scan else False
    scan_speed = ob['scan_speed'] * (u.pixel / u.s) if spatial_scan else False
    ssv_classes = {'sine': scan_speed_varations.SSVSine, 'mod-sine': scan_speed_varations.SSVModulatedSine}
    ssv_gen = None
    if ob.get('ssv_type'):
        if ob['ssv_type'] not in ssv_classes:
            raise WFC3SimConfigError("Invalid ssv_type given")
        ssv_gen = ssv_classes[ob['ssv_type']](*ob['ssv_coeffs'])

    obs = observation.Observation(outdir)
    obs.setup_detector(det, ob['NSAMP'], ob['SAMPSEQ'], ob['SUBARRAY'])
    obs.setup_grism(chosen_grism)
    obs.setup_target(planet if planet_file else name, wl, depth_planet, stellar_flux_scaled,
                     tgt['transit_time'], tgt['ldcoeffs'], tgt['period'], tgt['rp'], tgt['sma'],
                     tgt['inclination'], tgt['eccentricity'], tgt['periastron'], tgt['stellar_radius'])
    obs.setup_visit(start_JD, ob['num_orbits'], exp_start_times)
    obs.setup_reductions(ob['add_dark'], ob['add_flat'], ob['add_gain_variations'], ob['add_non_linear'],
                         ob['add_initial_bias'])
    obs.setup_observation(x_ref, y_ref, spatial_scan, scan_speed)
    obs.setup_simulator(sample_rate, ob['clip_values_det_limits'], _get(cfg, 'general', 'threads', 2))
    obs.setup_trends(ssv_gen, ob['x_shifts'], ob['x_jitter'], ob['y_shifts'], ob['y_jitter'])
    obs.setup_noise_sources(sky, ob['cosmic_rate'], ob['add_read_noise'], ob['add_stellar_noise'])
    obs.setup_gaussian_noise(ob['noise_mean'], ob['noise_std'])
    coeffs = _get(cfg, 'trends', 'visit_trend_coeffs')
    if coeffs:
        obs.setup_visit_trend(coeffs)
    return obs


def run(argv=None):
    ap = argparse.ArgumentParser(prog='wayne', description=__doc__.split('\n\n')[1])
    ap.add_argument('-p', dest='parameter_file', required=True)
    ap.add_argument('--max-exposures', type=int, default=None, help='stop after this many exposures')
    ap.add_argument('--gpus', type=int, default=None,
                    help='partition the visit exposure-wise over this many GPUs (one process each)')
    args = ap.parse_args(argv)
    with open(args.parameter_file) as f:
        cfg = yaml.safe_load(f)
    if args.gpus and args.gpus > 1 and 'RANK' not in os.environ:
        return _spawn_ranks(args, cfg)
    base = os.path.dirname(os.path.abspath(args.parameter_file))
    obs = build_observation(cfg, base)
    shutil.copy2(args.parameter_file, os.path.join(obs.outdir, os.path.basename(args.parameter_file)))
    if args.max_exposures:
        obs.exp_start_times = obs.exp_start_times[:args.max_exposures]
    import torch
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    if torch.cuda.is_available():
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        if world > 1:
            from .engine import bind_to_gpu_numa_node
            bind_to_gpu_numa_node(local)
    out = obs.run_observation(shard=(rank, world))
    sys.stdout.write('rank {}: wrote {} exposures to {}\n'.format(rank, len(out), obs.outdir))
    return out


def _spawn_ranks(args, cfg):
    """``--gpus n``: one child process per GPU, each generating its strided share of the
    exposures; the parent only hands down the rank layout and the shared seed."""
    import subprocess
    seed = shared_visit_seed(_get(cfg, 'general', 'seed'))
    cmd = [sys.executable, '-m', 'wayne_b200.run_visit', '-p', args.parameter_file]
    if args.max_exposures:
        cmd += ['--max-exposures', str(args.max_exposures)]
    procs = []
    for r in range(args.gpus):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(args.gpus),
                   WAYNE_B200_VISIT_SEED=str(seed))
        env.pop('MASTER_ADDR', None)
        procs.append(subprocess.Popen(cmd, env=env))
    rcs = [p.wait() for p in procs]
    if any(rcs):
        raise RuntimeError("rank exit codes {}".format(rcs))
    return {}


if __name__ == '__main__':
    run()
