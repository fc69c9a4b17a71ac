"""The exposure-level C interface (include/wayne_b200.h: wb200_ctx_* / wb200_exposure_run):
one call per exposure on a resident context must give the SAME reads, bit for bit, as the
stage-by-stage calls driven from Python (which the parity suite pins to the oracle), for
every form of the planet signal, every switch combination and both output dtypes."""
import numpy as np
import pytest

from tests import harness

pytestmark = pytest.mark.gpu


def _gen(sub=256, seq='SPARS10', nsamp=5, grism_name='G141'):
    from wayne import detector, grism
    from wayne.exposure_generator import ExposureGenerator
    g = grism.G141() if grism_name == 'G141' else grism.G102()
    return ExposureGenerator(detector.WFC3_IR(), g, nsamp, seq, sub, None, rng='philox')


def _both(monkeypatch, make, **kw):
    """Run the same exposure through the context and through the staged calls."""
    from wayne_b200 import params
    out = {}
    for use in (True, False):
        monkeypatch.setattr(params, 'use_context', use)
        eg = make()
        exp = eg.scanning_frame(**kw)
        out[use] = (np.array([r[0] for r in exp.reads]), eg.photons, type(eg._run).__name__)
    assert out[True][2] == 'ContextRun' and out[False][2] == 'ExposureRun'
    return out


def test_single_call_equals_staged_path_bit_for_bit(calb_dir, monkeypatch):
    from wayne import units as u
    from wayne.trend_generators.scan_speed_varations import SSVSine
    wl, flux, planet = harness.spectrum(level=3.0e-14)
    n = len(_gen()._gen_scanning_sample_times(100 * u.ms)[3]) and \
        len(np.asarray(u.value_in(_gen()._gen_scanning_sample_times(100 * u.ms)[1], u.ms)))
    depth = np.tile(planet, (n, 1)) * np.linspace(0.1, 1.0, n)[:, None]
    kw = dict(x_ref=404.5, y_ref=457.4, x_jitter=0.02, y_jitter=0.02, wl=wl * u.micron, stellar_flux=flux,
              planet_signal=depth, scan_speed=7.4325 * u.pixel / u.s, sample_rate=100 * u.ms,
              ssv_generator=SSVSine(1.5, 1.1, 0), cosmic_rate=11., sky_background=5.5 * u.count / u.s,
              scale_factor=0.9991, noise_mean=0.5, noise_std=0.1, rng_key=(1963, 5))
    out = _both(monkeypatch, _gen, **kw)
    assert out[True][1] == out[False][1] > 1e6
    assert out[True][0].shape == (5, 266, 266)
    assert np.array_equal(out[True][0], out[False][0])
    assert np.abs(out[True][0][0]).max() > 0                       # SUBARRAY 256: the initial bias is in


@pytest.mark.parametrize("switch", ["add_flat", "add_dark", "add_gain_variations", "add_non_linear",
                                    "clip_values_det_limits", "add_read_noise", "add_stellar_noise",
                                    "add_initial_bias"])
def test_each_switch_off(calb_dir, monkeypatch, switch):
    from wayne import units as u
    wl, flux, planet = harness.spectrum(level=2.0e-14)
    kw = dict(x_ref=404.5, y_ref=457.4, x_jitter=0.02, y_jitter=0.02, wl=wl * u.micron, stellar_flux=flux,
              planet_signal=None, scan_speed=7.4325 * u.pixel / u.s, sample_rate=400 * u.ms,
              cosmic_rate=11., sky_background=2.0 * u.count / u.s, rng_key=(7, 8))
    kw[switch] = False
    out = _both(monkeypatch, _gen, **kw)
    assert np.array_equal(out[True][0], out[False][0]) and out[True][1] == out[False][1]


def test_planet_signal_forms(calb_dir, monkeypatch):
    """Dense host array, CUDA tensor, SeparableSignal (bit-identical to the dense product) and
    ChebyshevSignal (host and device coefficients) through the context."""
    import torch
    from wayne import units as u
    from wayne_b200 import lightcurve as lc
    wl, flux, planet = harness.spectrum(level=2.0e-14)
    eg = _gen()
    _, mid, dur, ri = eg._gen_scanning_sample_times(150 * u.ms)
    n = len(ri) and len(np.asarray(u.value_in(mid, u.ms)))
    curve = 0.5 * (1 + np.tanh(np.linspace(-2, 2, n)))
    dense = planet[None, :] * curve[:, None]
    kw = dict(x_ref=404.5, y_ref=457.4, x_jitter=0.02, y_jitter=0.02, wl=wl * u.micron, stellar_flux=flux,
              scan_speed=7.4325 * u.pixel / u.s, sample_rate=150 * u.ms, cosmic_rate=None,
              sky_background=1.0 * u.count / u.s, rng_key=(11, 12))
    ref = _both(monkeypatch, _gen, planet_signal=dense, **kw)
    assert np.array_equal(ref[True][0], ref[False][0])
    from wayne_b200 import params
    monkeypatch.setattr(params, 'use_context', True)

    def reads(signal):
        g = _gen()
        e = g.scanning_frame(planet_signal=signal, **kw)
        return np.array([r[0] for r in e.reads]), g.photons

    sep = reads(lc.SeparableSignal(curve, planet))
    assert np.array_equal(sep[0], ref[True][0]) and sep[1] == ref[True][1]
    dev = reads(torch.from_numpy(dense).cuda())
    assert np.array_equal(dev[0], ref[True][0])
    # Chebyshev form: order 2 in x = normalised depth reproduces the separable signal to rounding,
    # so the expected counts agree to ~1e-16 and the frames are statistically the same exposure
    mid0, half0 = 0.5 * (planet.max() + planet.min()), 0.5 * (planet.max() - planet.min())
    cheb = lc.ChebyshevSignal(np.c_[curve * mid0, curve * half0], (planet - mid0) / half0)
    a = reads(cheb)
    b = reads(lc.ChebyshevSignal(torch.from_numpy(cheb.coef).cuda(), cheb.x))
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    assert abs(a[1] - ref[True][1]) < 5 * np.sqrt(ref[True][1])
    monkeypatch.setattr(params, 'use_context', False)
    c = reads(cheb)
    assert np.array_equal(a[0], c[0])


def test_full_frame_float32_output_and_reuse(calb_dir, monkeypatch):
    """SUBARRAY 1024 (offset -5), float32 output, and back-to-back exposures on one context:
    the interval planes are handed back zeroed by the ramp pass (no memset), so a second
    exposure must not see the first one's electrons."""
    from wayne import units as u
    from wayne_b200 import calibration
    calibration.write_synthetic_calibration(calb_dir, modes=((1024, 'RAPID'),))
    wl, flux, planet = harness.spectrum(n_wl=400, lo=0.95, hi=1.8, level=2.0e-14)
    kw = dict(x_ref=330.0, y_ref=110.0, x_jitter=0.02, y_jitter=0.02, wl=wl * u.micron, stellar_flux=flux,
              planet_signal=None, scan_speed=20.0 * u.pixel / u.s, sample_rate=300 * u.ms, cosmic_rate=11.,
              sky_background=3.0 * u.count / u.s, out_dtype=np.float32)
    mk = lambda: _gen(1024, 'RAPID', 6)                                         # noqa: E731
    first = _both(monkeypatch, mk, rng_key=(1, 1), **kw)
    assert first[True][0].dtype == np.float32 and np.array_equal(first[True][0], first[False][0])
    from wayne_b200 import params
    monkeypatch.setattr(params, 'use_context', True)
    seq = []
    for key in ((2, 2), (1, 1), (2, 2)):
        g = mk()
        e = g.scanning_frame(rng_key=key, **kw)
        seq.append(np.array([r[0] for r in e.reads]))
    assert np.array_equal(seq[1], first[True][0])                   # same key -> same frames, whatever ran before
    assert np.array_equal(seq[0], seq[2]) and not np.array_equal(seq[0], seq[1])


def test_context_refuses_bad_input(calb_dir):
    import ctypes as C
    from wayne import detector, grism
    from wayne_b200 import _lib
    from wayne_b200.engine import DeviceEngine
    eng = DeviceEngine.get()
    ctx = eng.exposure_context(grism.G141(), detector.WFC3_IR(), 256, 'SPARS10')
    a = _lib.ExposureArgs()
    a.n_samples, a.n_bins, a.n_reads, a.count_mode = 4, 16, 4, _lib.COUNT_NONE
    rc = _lib.lib.wb200_exposure_run(ctx._h, C.byref(a), C.c_void_p(8), None)
    assert rc == -1 and b"count_mode" in _lib.lib.wb200_ctx_last_error(ctx._h)
    bad = np.array([0.1 + 1e-12], dtype=np.float64)                 # not float32-representable
    rc = _lib.lib.wb200_ctx_upload_plane(ctx._h, _lib.PLANE_SKY, C.c_void_p(bad.ctypes.data), _lib.F64, 1)
    assert rc == -1 and b"float32" in _lib.lib.wb200_ctx_last_error(ctx._h)


def test_stated_limits_are_enforced_and_accounted_for(calb_dir):
    """(1) more than 65535 sub-samples per exposure (grid.y of the kernels) is refused with a
    message, not truncated; (2) bins whose position is beyond the +-2^22 px range of the
    thrower's magic-add floor (or NaN) throw nothing and every one of their electrons is
    accounted as dropped: binned + dropped == thrown still holds exactly."""
    from wayne import detector, grism
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    from wayne_b200 import _lib
    wl, flux, planet = harness.spectrum(level=2.0e-14)
    eg = ExposureGenerator(detector.WFC3_IR(), grism.G141(), 5, 'SPARS10', 256, None, rng='philox')
    with pytest.raises(_lib.WayneB200Error, match="65535"):
        eg.scanning_frame(404.5, 457.4, 0.02, 0.02, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s,
                          0.3 * u.ms, cosmic_rate=None)                      # 74 000 sub-samples
    for x_ref in (2.0e7, -9.0e6, float('nan')):
        eg = ExposureGenerator(detector.WFC3_IR(), grism.G141(), 5, 'SPARS10', 256, None, rng='philox')
        exp = eg.scanning_frame(x_ref, 457.4, 0.0, 0.0, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s,
                                400 * u.ms, cosmic_rate=None, sky_background=0 * u.count / u.s, add_dark=False,
                                add_non_linear=False, add_read_noise=False, add_initial_bias=False, rng_key=(3, 3))
        stats = eg._run.stats.cpu().numpy()
        if np.isnan(x_ref):
            assert stats[1] == 0 and stats[1] + stats[2] == stats[0]         # NaN positions: nothing binned
        else:
            assert stats[0] > 1e6 and stats[1] == 0 and stats[2] == stats[0]
        assert np.all(np.array([r[0] for r in exp.reads]) == 0)


def test_overlapped_exposures_equal_serial_ones(calb_dir):
    """Consecutive exposures overlap on the device (the next exposure's tables and counts are made
    on the context's own stream, into the other scratch lane, beside the current electron throw;
    include/wayne_b200.h).  Eight different exposures queued back to back -- host planet signals,
    a device-resident one handed over repeatedly, none -- must give, bit for bit, the frames of
    the same exposures run one at a time with everything on the caller's stream."""
    import torch
    from wayne import units as u
    from wayne_b200.engine import DeviceEngine
    from wayne_b200.lightcurve import SeparableSignal
    eng = DeviceEngine.get()
    wl, flux, planet = harness.spectrum(level=3.0e-14)
    n = len(np.asarray(u.value_in(_gen()._gen_scanning_sample_times(100 * u.ms)[1], u.ms)))
    dense = np.tile(planet, (n, 1)) * np.linspace(0.1, 1.0, n)[:, None]
    dense_dev = torch.from_numpy(dense).to(eng.device)
    torch.cuda.synchronize()

    def signal(i):
        return (None, dense * (1 + 0.1 * i), dense_dev,
                SeparableSignal(np.linspace(0.2, 1.0, n) * (1 + 0.05 * i), planet))[i % 4]

    def queue(i):
        eg = _gen()
        exp = eg.scanning_frame(x_ref=404.5 + 0.3 * i, y_ref=457.4 - 0.2 * i, x_jitter=0.02, y_jitter=0.02,
                                wl=wl * u.micron, stellar_flux=flux * (1 + 0.02 * i), planet_signal=signal(i),
                                scan_speed=7.4325 * u.pixel / u.s, sample_rate=100 * u.ms, cosmic_rate=11.,
                                sky_background=5.5 * u.count / u.s, scale_factor=1.0 - 1e-3 * i,
                                rng_key=(1963, 100 + i))
        return eg, exp

    assert not eng.profile
    handles = [queue(i) for i in range(8)]                         # all eight in flight, overlapped
    overlapped = [(np.array([r[0] for r in exp.reads]), eg.photons) for eg, exp in handles]
    serial = []
    eng.profile = True                                             # per-stage timing keeps an exposure on one stream
    try:
        for i in range(8):
            eg, exp = queue(i)
            serial.append((np.array([r[0] for r in exp.reads]), eg.photons))
            torch.cuda.synchronize()
    finally:
        eng.profile = False
        eng.stage_times()
    for i in range(8):
        assert overlapped[i][1] == serial[i][1] > 1e6, i
        assert np.array_equal(overlapped[i][0], serial[i][0]), i
    assert not np.array_equal(overlapped[0][0], overlapped[4][0])


def test_tile_placement_from_chunk_spans_equals_the_scan(calb_dir, monkeypatch):
    """The native thrower places each CTA's tile from the chunk's first / last populated bins
    (k_chunk_spans in stage 1) instead of scanning the chunk's counts and positions.  Where the
    tile sits only decides which electrons take the slow path, never where they land: the frames
    with WB200_THROW_SCAN=1 (the scan) must be bit-identical."""
    from wayne import units as u
    from wayne.trend_generators.scan_speed_varations import SSVSine
    wl, flux, planet = harness.spectrum(level=3.0e-14)
    flux = np.array(flux, dtype=float)
    flux[: len(flux) // 3] = 0.0                 # a dark stretch: the span starts inside the first chunk
    kw = dict(x_ref=404.5, y_ref=457.4, x_jitter=0.02, y_jitter=0.02, wl=wl * u.micron, stellar_flux=flux,
              planet_signal=None, scan_speed=7.4325 * u.pixel / u.s, sample_rate=100 * u.ms,
              ssv_generator=SSVSine(1.5, 1.1, 0), cosmic_rate=11., sky_background=5.5 * u.count / u.s,
              rng_key=(1963, 77))
    out = {}
    for scan in (False, True):
        if scan:
            monkeypatch.setenv('WB200_THROW_SCAN', '1')
        eg = _gen()
        exp = eg.scanning_frame(**kw)
        out[scan] = (np.array([r[0] for r in exp.reads]), eg.photons)
    assert out[False][1] == out[True][1] > 1e5
    assert np.array_equal(out[False][0], out[True][0])
