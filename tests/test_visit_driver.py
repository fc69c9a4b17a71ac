"""Visit driver: planner, spectrum preparation, parameter
file wiring on the CPU; a short pipelined visit on the GPU."""
import os

import numpy as np
import pytest
import yaml

from wayne import detector, observation, run_visit, tools, visit_planner
from wayne import units as u

LD = [0.800627, -0.757066, 0.897268, -0.384804]


def test_visit_planner():
    det = detector.WFC3_IR()
    vp = visit_planner.VisitPlanner(det, 5, 'SPARS10', 256, num_orbits=2)
    t = vp['exp_times'].to(u.min).value
    assert vp['num_exp'] == len(t) and vp['orbit_start_index'][0] == 0
    o1 = vp['orbit_start_index'][1]
    assert t[0] == 6.0 and abs(t[o1] - (95.0 + 5.0)) < 1e-9            # guide-star acquisition 6 / 5 min
    step = det.exptime(5, 256, 'SPARS10').to(u.min).value + 1.0
    gaps = np.diff(t[:o1])
    assert np.all((np.abs(gaps - step) < 1e-9) | (np.abs(gaps - step - 5.8) < 1e-9))   # buffer dumps
    assert t[o1 - 1] < 54.0 and np.all(np.diff(t) > 0)
    assert observation.detect_orbits(np.array([1.001, 1.002, 1.032]) * u.day) == [0, 2]   # test_tools.py:84-89


def test_rebin_and_blackbody():
    wl = np.linspace(0.9, 1.9, 4001)
    sp = 2.0 + np.sin(9 * wl)
    new = tools.wl_at_resolution(130, 0.988, 1.777)
    r = tools.rebin_spec(wl, sp, new)
    edges = tools.bin_centers_to_edges(new)
    exact = np.diff(2.0 * edges - np.cos(9 * edges) / 9) / np.diff(edges)
    assert np.abs(r - exact).max() < 1e-6                        # bin averages of the input
    assert np.allclose(tools.rebin_spec(wl, np.full_like(wl, 3.3), new), 3.3)
    b = tools.blackbody_lambda(np.array([0.5, 1.0, 2.0]), 6065.0)
    assert b[0] > b[1] > b[2] > 0                                 # Wien peak at 0.48 micron
    assert abs(b[1] / 1.2251e6 - 1) < 1e-3


def _write_visit(tmp_path, n_exp=3, scan=True):
    d = tmp_path
    wl = np.linspace(0.85, 1.85, 700)
    np.savetxt(d / 'planet.dat', np.c_[wl, 0.0146 * (1 + 0.01 * np.sin(11 * wl))], fmt='%.8f')
    jd = 2456196.28836 - 0.05 + np.arange(n_exp) * 0.0011
    np.savetxt(d / 'jd.txt', jd, fmt='%.8f')
    np.savetxt(d / 'xref.txt', 404.0 + 0.1 * np.arange(n_exp))
    np.savetxt(d / 'yref.txt', 457.3 + 0.05 * np.arange(n_exp))
    np.savetxt(d / 'sky.txt', 5.0 + 0.2 * np.arange(n_exp))
    cfg = {
        'general': {'oec_location': False, 'outdir': 'out', 'seed': 1963, 'threads': 2},
        'target': {'name': 'HD 209458 b', 'planet_spectrum_file': 'planet.dat', 'rebin_resolution': False,
                   'stellar_spectrum_file': False, 'stellar_temperature': 6065, 'flux_scale': 2.5e-19,
                   'period': 3.524746, 'sma': 0.047309, 'stellar_radius': 1.155, 'inclination': 86.71,
                   'eccentricity': 0.0, 'periastron': 0.0, 'transit_time': 2456196.28836, 'ldcoeffs': LD},
        'observation': {'detector': 'WFC3IR', 'grism': 'G141', 'x_ref': 'xref.txt', 'y_ref': 'yref.txt',
                        'NSAMP': 5, 'SAMPSEQ': 'SPARS10', 'SUBARRAY': 256, 'start_JD': False,
                        'exp_start_times': 'jd.txt', 'num_orbits': 1, 'sample_rate': 100,
                        'spatial_scan': scan, 'scan_speed': 7.4325, 'ssv_type': 'sine',
                        'ssv_coeffs': [1.5, 1.1, 0], 'x_shifts': 0, 'x_jitter': 0.025, 'y_shifts': 0,
                        'y_jitter': 1e-15, 'noise_mean': False, 'noise_std': False, 'add_dark': True,
                        'add_flat': True, 'add_gain_variations': True, 'add_non_linear': True,
                        'add_read_noise': True, 'add_initial_bias': True, 'add_stellar_noise': True,
                        'sky_background': 'sky.txt', 'cosmic_rate': 11, 'clip_values_det_limits': True},
        'trends': {'visit_trend_coeffs': [0.005, 0.0011, 400, 2456196.28836]},
    }
    with open(d / 'params.yml', 'w') as f:
        yaml.safe_dump(cfg, f)
    return str(d / 'params.yml')


def test_parameter_file_wiring(tmp_path):
    pfile = _write_visit(tmp_path)
    with open(pfile) as f:
        obs = run_visit.build_observation(yaml.safe_load(f), str(tmp_path))
    assert obs.NSAMP == 5 and obs.SUBARRAY == 256 and obs.spatial_scan and obs.transmission_spectroscopy
    assert len(obs.exp_start_times) == 3 and obs.visit_plan['orbit_start_index'] == [0]
    assert len(obs.wl) == len(obs.stellar_flux) == len(obs.planet_spectrum)
    assert 0.9 <= obs.wl.value.min() and obs.wl.value.max() <= 1.8        # run_visit.py:151-152 crop
    assert obs._visit_trend is not False and len(obs._visit_trend.scale_factors) == 3
    # light curves: dense reference-style evaluation == the Chebyshev form the device path uses
    t = obs.exp_start_times[1] + np.linspace(0, 20, 7) * u.s
    dense = 1 - obs.generate_lightcurves(t.to(u.day))
    cheb = obs._planet_signal(t.to(u.day)).to_array()
    assert dense.shape == cheb.shape == (7, len(obs.wl))
    assert np.abs(dense - cheb).max() < 1e-9 and dense.max() > 1e-3


@pytest.mark.gpu
def test_short_visit_on_gpu(tmp_path, calb_dir):
    from wayne import fitsio
    pfile = _write_visit(tmp_path, n_exp=4)
    out = run_visit.run(['-p', pfile])
    assert sorted(out) == [1, 2, 3, 4]
    outdir = os.path.join(str(tmp_path), 'out')
    files = sorted(os.listdir(outdir))
    assert files == ['0000_flt.fits', '0001_raw.fits', '0002_raw.fits', '0003_raw.fits', '0004_raw.fits',
                     'params.yml']
    f = fitsio.open(os.path.join(outdir, '0003_raw.fits'))
    assert len(f) == 1 + 5 * 5 and f[0].header['FILENAME'] == '0003_raw.fits'
    assert abs(f[0].header['SKY-LVL'] - 5.4) < 1e-9 and f[0].header['RNG'] == 'philox'
    last, zero = f[1].data, f[21].data
    assert last.shape == (266, 266) and np.isfinite(last).all()
    sig = (last - zero)[5:-5, 5:-5]
    assert sig.sum() > 1e6 and sig[100:, :].sum() > 20 * abs(sig[:40, :].sum())    # the scan is on the frame
    # the file a GPU exposure was written to, read back by the strict parser written from the FITS
    # standard (tests/fits_standard.py shares nothing with wayne_b200.fitsio): layout of the reference's
    # writer (wayne/exposure.py:133-214) and the same pixels
    from tests import fits_standard as FS
    hdus = FS.read(os.path.join(outdir, '0003_raw.fits'))
    assert [h['EXTNAME'] for h, _, _ in hdus[1:]] == ['SCI', 'ERR', 'DQ', 'SAMP', 'TIME'] * 5
    assert hdus[1][0]['BITPIX'] == -64 and hdus[1][0]['SAMPNUM'] == 4 and hdus[21][0]['SAMPNUM'] == 0
    assert np.array_equal(hdus[1][2], last) and np.array_equal(hdus[21][2], zero)
    # same visit, exposure-wise over two "ranks": identical frames (keys do not depend on the partition)
    from wayne_b200 import params
    with open(pfile) as fh:
        cfg = yaml.safe_load(fh)
    cfg['general']['outdir'] = 'out2'
    obs = run_visit.build_observation(cfg, str(tmp_path))
    a = obs.run_observation(shard=(1, 2), write_fits=False)
    assert sorted(a) == [2, 4] and params.seed == 1963
    g = fitsio.open(os.path.join(outdir, '0002_raw.fits'))
    assert np.array_equal(a[2].reads[-1][0], g[1].data)


@pytest.mark.gpu
def test_device_transit_kernel_matches_host_model():
    from wayne_b200 import lightcurve as lc
    from wayne_b200.engine import DeviceEngine
    orb = dict(period=3.524746, a=8.81, e=0.0, inc_deg=86.71, w_deg=0.0, t0=2456196.28836)
    depth = 0.0146 * (1 + 0.02 * np.sin(np.linspace(0, 9, 500)))
    for t in (orb['t0'] + np.linspace(-0.09, 0.09, 333),          # whole transit incl. contacts
              orb['t0'] + 1.2 + np.linspace(0, 0.01, 17)):        # out of transit
        host = lc.planet_signal(t, depth, LD, **orb)
        dev = lc.planet_signal_device(DeviceEngine.get(), t, depth, LD, **orb)
        assert hasattr(dev.coef, 'is_cuda')
        npt = np.abs(dev.coef.cpu().numpy() - host.coef).max()
        assert npt < 1e-13, npt
        assert np.abs(dev.to_array() - host.to_array()).max() < 1e-13
    e = dict(orb, e=0.2, w_deg=60.0)
    t = orb['t0'] + np.linspace(-0.1, 0.1, 101)
    assert np.abs(lc.planet_signal_device(DeviceEngine.get(), t, depth, LD, **e).to_array()
                  - lc.planet_signal(t, depth, LD, **e).to_array()).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("ssv_type,coeffs", [('mod-sine', [10, 1.1, 100]), ('sine', [1.5, 1.1, 'rand'])])
def test_sharded_visit_with_stochastic_ssv_is_partition_independent(tmp_path, calb_dir, ssv_type, coeffs):
    """Scan-speed-variation generators draw from numpy's global stream (scan_speed_varations.py:45,
    100-132).  With the visit spread over ranks each rank has drawn a different number of values
    before a given exposure; the frames must not depend on that."""
    pfile = _write_visit(tmp_path, n_exp=4)
    with open(pfile) as fh:
        cfg = yaml.safe_load(fh)
    cfg['observation']['ssv_type'] = ssv_type
    cfg['observation']['ssv_coeffs'] = coeffs
    frames = {}
    for tag, shards in (('one', [(0, 1)]), ('two', [(0, 2), (1, 2)])):
        cfg['general']['outdir'] = 'out_' + tag
        got = {}
        for shard in shards:
            obs = run_visit.build_observation(cfg, str(tmp_path))       # reseeds numpy like a fresh rank
            got.update(obs.run_observation(shard=shard, write_fits=False))
        frames[tag] = {n: np.array([r[0] for r in e.reads]) for n, e in got.items()}
    assert sorted(frames['one']) == sorted(frames['two']) == [1, 2, 3, 4]
    for n in (1, 2, 3, 4):
        assert np.array_equal(frames['one'][n], frames['two'][n]), n
    assert not np.array_equal(frames['one'][1][-1], frames['one'][2][-1])


@pytest.mark.gpu
def test_compat_visit_equals_the_oracle_exposure_by_exposure(tmp_path, calb_dir, monkeypatch):
    """The batched visit driver against the ORACLE, not against itself: a two-exposure visit in the
    reference's stream mode (rng='numpy': one sequential numpy stream across the visit,
    run_visit.py:73-77) must give, exposure by exposure, the reads of the numpy restatement of
    ExposureGenerator.scanning_frame fed what Observation._generate_exposure hands it
    (wayne/observation.py:415-504): per-exposure x_ref / y_ref / sky from the lists plus the linear
    shifts (:446-453), the visit-trend factor of that exposure (:455-458, trend recomputed by the
    oracle's hook_and_long_term_ramp), the sample times of the mode, 1 - light curves as the planet
    signal (:441-443), and ONE RandomState consumed in exposure order."""
    from oracle import exposure_oracle as E
    from tests import harness
    from wayne_b200 import params
    from wayne.exposure_generator import ExposureGenerator
    monkeypatch.setattr(params, 'rng', params.rng)                    # build_observation sets it globally
    pfile = _write_visit(tmp_path, n_exp=2)
    with open(pfile) as fh:
        cfg = yaml.safe_load(fh)
    cfg['general']['rng'] = 'numpy'
    cfg['observation'].update(x_shifts=0.3, y_shifts=-0.2)
    obs = run_visit.build_observation(cfg, str(tmp_path))            # np.random.seed(1963)
    got = obs.run_observation(write_fits=False)
    assert sorted(got) == [1, 2]

    cal = harness.oracle_calibration()
    rs = np.random.RandomState(1963)
    eg = ExposureGenerator(obs.detector, obs.grism, obs.NSAMP, obs.SAMPSEQ, obs.SUBARRAY, None, rng='numpy')
    read_times = np.asarray(u.value_in(eg.read_times, u.s), dtype=float)
    _, mid, dur, ri = eg._gen_scanning_sample_times(obs.sample_rate)
    jd = np.array([float(u.value_in(t, u.day)) for t in obs.exp_start_times])
    a1, b1, b2, to = cfg['trends']['visit_trend_coeffs']
    t0 = E.gen_orbit_start_times_per_exp(jd, obs.visit_plan['orbit_start_index'])
    trend = E.hook_and_long_term_ramp(jd, t0, a1, b1, b2, to)
    xs, ys, sky = (np.loadtxt(str(tmp_path / f)) for f in ('xref.txt', 'yref.txt', 'sky.txt'))
    wl = np.asarray(u.value_in(obs.wl, u.micron), dtype=float)
    flux = np.asarray(getattr(obs.stellar_flux, 'value', obs.stellar_flux), dtype=float)
    for i in range(2):
        t = (mid + obs.exp_start_times[i]).to(u.day)
        depth = 1.0 - obs.generate_lightcurves(t)                     # [n_samples][n_wl], observation.py:441-443
        o = E.scanning_frame(cal, 'G141', 256, read_times, wl, flux, depth, xs[i] + 0.3 * i, ys[i] - 0.2 * i,
                             0.025, 1e-15, 7.4325e-3, 100.0, rs, ssv=(1.5, 1.1, 0), cosmic_rate=11.,
                             sky_background=float(sky[i]), scale_factor=float(trend[i]), threads=2)
        reads = got[i + 1].reads
        assert len(reads) == len(o['reads']) == 5
        for r in range(5):
            err = np.abs(reads[r][0] - o['reads'][r]).max()
            assert err <= 1e-9 * max(1.0, np.abs(o['reads'][r]).max()), (i, r, err)
    assert np.abs(got[1].reads[-1][0] - got[2].reads[-1][0]).max() > 1.0   # the exposures do differ
