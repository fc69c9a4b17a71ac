"""More GPU parity cases of the exposure path against the CPU oracle, in the
deterministic ('numpy' compat) mode: config 3 (G102, 1024 full frame, NSAMP=15,
scan + SSV + cosmics + trend), other subarrays, switched-off terms and edge
cases (empty spectrum, trace leaving the frame)."""
import numpy as np
import pytest

from oracle import exposure_oracle as E
from tests import harness

pytestmark = pytest.mark.gpu


def _pair(calb_dir, grism_name, sub, nsamp, seq, seed, wl, flux, depth, x_ref, y_ref, scan, rate_ms,
          level_kw=None, threads=2, **kw):
    """Run product (compat) and oracle on the same inputs; returns (exposure, oracle dict, generator)."""
    from wayne import detector, grism
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    from wayne.trend_generators.scan_speed_varations import SSVSine
    from wayne_b200 import calibration
    calibration.write_synthetic_calibration(calb_dir, modes=((sub, seq),))
    cal = harness.oracle_calibration(grism_name, dark_mode=(sub, seq), nsamp=nsamp)
    g = grism.G141() if grism_name == 'G141' else grism.G102()
    eg = ExposureGenerator(detector.WFC3_IR(), g, nsamp, seq, sub, None, rng='numpy')
    ssv = kw.pop('ssv', None)
    okw = dict(kw)
    if 'sky_background' in kw:
        kw['sky_background'] = kw['sky_background'] * u.count / u.s
    np.random.seed(seed)
    exp = eg.scanning_frame(x_ref, y_ref, 0.02, 0.02, wl * u.micron, flux, depth, scan * u.pixel / u.s,
                            rate_ms * u.ms, ssv_generator=SSVSine(*ssv) if ssv else None,
                            threads=threads, **kw)
    o = E.scanning_frame(cal, grism_name, sub, eg.read_times.to(u.s).value, wl, flux, depth, x_ref, y_ref,
                         0.02, 0.02, scan * 0.001, rate_ms, np.random.RandomState(seed), ssv=ssv,
                         threads=threads, **okw)
    return exp, o, eg


def _check(exp, o, eg, tol=1e-9):
    assert eg.photons == o['photons']
    peak = max(1.0, max(np.abs(r).max() for r in o['reads']))
    assert len(exp.reads) == len(o['reads'])
    for r, want in enumerate(o['reads']):
        got = exp.reads[r][0]
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= tol * peak, r


def test_config3_g102_full_frame(calb_dir):
    """G102, SUBARRAY 1024 (1014 light-sensitive), NSAMP=15 RAPID, scan, SSV,
    cosmics, visit trend, all reductions; B1/B2/B3 quirks with their defined behaviour."""
    wl, flux, planet = harness.spectrum(n_wl=500, lo=0.7, hi=1.25, level=1.5e-14)
    from wayne import detector
    from wayne import units as u
    rt = detector.WFC3_IR().get_read_times(15, 1024, 'RAPID').to(u.s).value
    n = len(E.gen_scanning_sample_times(rt, 700.0)[1])
    depth = np.tile(planet, (n, 1)) * np.linspace(0, 1, n)[:, None]
    exp, o, eg = _pair(calb_dir, 'G102', 1024, 15, 'RAPID', 31, wl, flux, depth, 350.3, 120.7, 18.0, 700.0,
                       ssv=(1.5, 1.1, 0), cosmic_rate=11., sky_background=3.0, scale_factor=0.9991)
    assert exp.reads[0][0].shape == (1024, 1024) and len(exp.reads) == 15
    assert o['photons'] > 2e6
    _check(exp, o, eg)
    # scan direction +y: the last interval's light is below (higher y than) the first's
    first = exp.reads[1][0] - exp.reads[0][0]
    last = exp.reads[14][0] - exp.reads[13][0]
    rows = np.arange(1024)[:, None]
    cy = lambda a: float((np.clip(a, 0, None)[200:900] * rows[200:900]).sum() / np.clip(a, 0, None)[200:900].sum())  # noqa: E731
    assert cy(last) > cy(first)


@pytest.mark.parametrize("sub,seq,nsamp", [(64, 'RAPID', 4), (128, 'RAPID', 6), (512, 'SPARS25', 3)])
def test_other_subarrays(calb_dir, sub, seq, nsamp):
    wl, flux, planet = harness.spectrum(level=3.0e-14)
    # put the first-order trace on the subarray: sub_scale = 507 - sub/2
    x_ref = 507 - sub // 2 + 0.15 * sub - 20.0
    y_ref = 507 - sub // 2 + 0.5 * sub
    exp, o, eg = _pair(calb_dir, 'G141', sub, nsamp, seq, 5, wl, flux, None, x_ref, y_ref, 0.5, 400.0,
                       cosmic_rate=11., sky_background=2.0)
    assert exp.reads[0][0].shape == (sub + 10, sub + 10)
    _check(exp, o, eg)


def test_switched_off_terms(calb_dir):
    wl, flux, planet = harness.spectrum(level=2.0e-14)
    exp, o, eg = _pair(calb_dir, 'G141', 256, 5, 'SPARS10', 9, wl, flux, None, 404.5, 457.4, 7.4325, 500.0,
                       add_dark=False, add_flat=False, cosmic_rate=None, sky_background=0,
                       add_gain_variations=False, add_non_linear=False, clip_values_det_limits=False,
                       add_read_noise=False, add_stellar_noise=False, add_initial_bias=False)
    assert o['photons'] > 1e6
    for r, want in enumerate(o['reads']):
        assert np.array_equal(exp.reads[r][0], want)          # nothing stochastic but the electrons
    assert np.all(exp.reads[0][0] == 0)
    # electrons / 2.35 are conserved up to the few that leave the frame
    total = exp.reads[-1][0].sum() * 2.35
    assert 0.97 * o['photons'] < total <= o['photons'] + 1e-6


def test_empty_spectrum_and_offframe_trace(calb_dir):
    wl, flux, planet = harness.spectrum(level=2.0e-14)
    # no light at all: reads are bias + noise terms only, still equal to the oracle
    exp, o, eg = _pair(calb_dir, 'G141', 256, 5, 'SPARS10', 3, wl, flux * 0.0, None, 404.5, 457.4, 7.4325,
                       800.0, cosmic_rate=11., sky_background=1.0)
    assert o['photons'] == 0
    _check(exp, o, eg)
    # source near the right edge: most of the trace falls off the subarray (electrons dropped,
    # strict 0 < x < nr), and the scan carries it over the top edge
    exp, o, eg = _pair(calb_dir, 'G141', 256, 5, 'SPARS10', 4, wl, flux, None, 560.0, 600.0, 7.4325, 800.0,
                       cosmic_rate=None, sky_background=0, add_dark=False, add_read_noise=False,
                       add_non_linear=False)
    lit = (o['reads'][-1] - o['reads'][0])[5:-5, 5:-5].sum() * 2.35
    assert 0 < lit < 0.6 * o['photons']
    for r, want in enumerate(o['reads']):
        assert np.array_equal(exp.reads[r][0], want)


def test_output_options_and_missing_dark(calb_dir):
    """float32 output, device-resident result, and the reference's "no dark file ->
    warn and switch dark off" behaviour (exposure_generator.py:414-423)."""
    import warnings

    import torch
    from wayne import detector, grism
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator, WFC3SimNoDarkFileWarning
    wl, flux, planet = harness.spectrum(level=2.0e-14)
    args = (404.5, 457.4, 0.02, 0.02, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s, 500 * u.ms)
    kw = dict(cosmic_rate=11., sky_background=2.0 * u.count / u.s, rng_key=(5, 6))

    def gen(sub=256, seq='SPARS10', nsamp=5):
        return ExposureGenerator(detector.WFC3_IR(), grism.G141(), nsamp, seq, sub, None, rng='philox')

    ref = gen().scanning_frame(*args, **kw)
    f32 = gen().scanning_frame(*args, out_dtype=np.float32, **kw)
    assert f32.reads[-1][0].dtype == np.float32
    assert np.array_equal(f32.reads[-1][0], ref.reads[-1][0].astype(np.float32))
    dev = gen().scanning_frame(*args, device_result=True, **kw)
    assert isinstance(dev.device_reads, torch.Tensor) and dev.device_reads.is_cuda
    assert np.array_equal(dev.device_reads.cpu().numpy()[-1], ref.reads[-1][0])
    assert float(u.value_in(ref.exp_info['sim_time'], u.s)) > 0
    # SUBARRAY 512 / SPARS25 has no super-dark in the synthetic set written for this session
    from wayne_b200 import params
    import os
    name = detector.WFC3_IR()._dark_file(512, 'SPARS25')
    path = os.path.join(params._calb_dir, name)
    if os.path.exists(path):
        os.remove(path)
    eg = gen(512, 'SPARS25', 3)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        exp = eg.scanning_frame(650.0, 500.0, 0.02, 0.02, wl * u.micron, flux, None, 1.0 * u.pixel / u.s,
                                2000 * u.ms, **kw)
        assert len(exp.reads) == 3
    assert any(issubclass(x.category, WFC3SimNoDarkFileWarning) for x in w)
    assert exp.exp_info['add_dark'] is False


def test_direct_path_full_frame_matches_gather_path(calb_dir, monkeypatch):
    """Native mode at SUBARRAY 1024 / G102 (flat index wrap, offset -5): fused
    tile-flush accumulation == windows + ordered gather, same Philox electrons."""
    from wayne import detector, grism
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    from wayne_b200 import calibration, params
    calibration.write_synthetic_calibration(calb_dir, modes=((1024, 'RAPID'),))
    wl, flux, planet = harness.spectrum(n_wl=400, lo=0.7, hi=1.25, level=2.0e-14)
    out = {}
    for direct in (True, False):
        monkeypatch.setattr(params, 'direct_accumulation', direct)
        eg = ExposureGenerator(detector.WFC3_IR(), grism.G102(), 4, 'RAPID', 1024, None, rng='philox')
        exp = eg.scanning_frame(8.0, 12.0, 0.02, 0.02, wl * u.micron, flux, None, 25.0 * u.pixel / u.s,
                                400 * u.ms, add_dark=False, cosmic_rate=None,
                                sky_background=0 * u.count / u.s, add_non_linear=False,
                                add_read_noise=False, rng_key=(9, 9))
        out[direct] = (np.array([r[0] for r in exp.reads]), eg.photons)
    assert out[True][1] == out[False][1] > 1e6
    a, b = out[True][0], out[False][0]
    assert b.max() > 10 and np.abs(a - b).max() / b.max() < 1e-7
    # the source sits in the corner: light reaches the first rows/columns of the light-sensitive area
    assert b[-1][5:40, 5:200].sum() > 0


def test_ssv_modulated_sine_exposure(calb_dir):
    """SSVModulatedSine (scan_speed_varations.py:63-171) through the exposure path: it replaces
    the durations AND the read indexes (exposure_generator.py:262-273) and comes out one
    sub-sample short, so the last sub-sample has zero duration (:340-342) and sub-samples after
    the last read index never reach a read (:361).  Compat mode: the generator draws from the
    numpy stream BEFORE the per-sub-sample seeds (A.7), which the oracle side replays."""
    from wayne import detector, grism
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    from wayne.trend_generators.scan_speed_varations import SSVModulatedSine
    from wayne_b200 import calibration
    calibration.write_synthetic_calibration(calb_dir, modes=((256, 'SPARS10'),))
    cal = harness.oracle_calibration('G141', dark_mode=(256, 'SPARS10'), nsamp=5)
    wl, flux, planet = harness.spectrum(level=4.0e-14)
    eg = ExposureGenerator(detector.WFC3_IR(), grism.G141(), 5, 'SPARS10', 256, None, rng='numpy')
    rate = 50.0
    np.random.seed(77)
    exp = eg.scanning_frame(404.5, 457.4, 0.02, 0.02, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s,
                            rate * u.ms, ssv_generator=SSVModulatedSine(10, 1.1, 100), cosmic_rate=11.,
                            sky_background=2.0 * u.count / u.s, threads=2)
    # oracle side: the same generator call on the same stream, then the oracle's own loop
    rt = eg.read_times.to(u.s).value
    _, mid, dur0, _ = E.gen_scanning_sample_times(rt, rate)
    np.random.seed(77)
    dur, ri = SSVModulatedSine(10, 1.1, 100).get_subsample_exposure_times(None, None, eg.read_times, rate * u.ms)
    dur = np.asarray(u.value_in(dur, u.ms))
    assert len(dur) <= len(mid) and ri != list(E.gen_scanning_sample_times(rt, rate)[3])
    dur = np.concatenate([dur, np.zeros(len(mid) - len(dur))])
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    o = E.scanning_frame(cal, 'G141', 256, rt, wl, flux, None, 404.5, 457.4, 0.02, 0.02, 7.4325 * 0.001, rate,
                         rs, cosmic_rate=11., sky_background=2.0, threads=2,
                         sample_times=(mid, dur, [int(i) for i in ri]))
    assert o['photons'] > 1e6
    _check(exp, o, eg)
    # native mode: the same generator object works there too, and both accumulation paths drop
    # the sub-samples after the last read index
    out = {}
    from wayne_b200 import params
    for direct in (True, False):
        params.direct_accumulation = direct
        try:
            np.random.seed(78)
            eg2 = ExposureGenerator(detector.WFC3_IR(), grism.G141(), 5, 'SPARS10', 256, None, rng='philox')
            e2 = eg2.scanning_frame(404.5, 457.4, 0.02, 0.02, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s,
                                    rate * u.ms, ssv_generator=SSVModulatedSine(10, 1.1, 100), cosmic_rate=None,
                                    sky_background=0 * u.count / u.s, add_dark=False, add_non_linear=False,
                                    add_read_noise=False, rng_key=(3, 4))
            out[direct] = np.array([r[0] for r in e2.reads])
        finally:
            params.direct_accumulation = True
    assert out[False].max() > 10 and np.abs(out[True] - out[False]).max() / out[False].max() < 1e-7
