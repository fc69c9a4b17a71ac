"""GPU parity of the whole exposure path through the reference-facing API
(ExposureGenerator.scanning_frame / staring_frame) against the CPU oracle
(oracle/exposure_oracle.py + oracle/psf_oracle.c).

Deterministic ('numpy' compat) mode: both sides consume the same numpy
RandomState in the reference's order and the same rand_r electron streams, so
the comparison is exact: int32 histograms bit-exact; float64 reads compared
bit-for-bit where stated, else to 1e-12 relative.  Stage-1 math: <= 1e-6
relative (north_star), in practice a few ulp.
"""
import numpy as np
import pytest

from oracle import exposure_oracle as E
from tests import harness

pytestmark = pytest.mark.gpu

X_REF, Y_REF = 404.497, 457.427


def _gen(grism_name='G141', nsamp=5, seq='SPARS10', sub=256, rng='numpy'):
    from wayne import detector, grism
    from wayne.exposure_generator import ExposureGenerator
    g = grism.G141() if grism_name == 'G141' else grism.G102()
    return ExposureGenerator(detector.WFC3_IR(), g, nsamp, seq, sub, None, rng=rng)


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def test_stage1_tables_and_trace(calb_dir):
    """Trace / dispersion / sensitivity / expected counts vs the float64 oracle: 1e-6 relative."""
    from wayne import units as u
    cal = harness.oracle_calibration()
    wl, flux, planet = harness.spectrum()
    eg = _gen(rng='numpy')
    np.random.seed(5)
    depth = np.tile(planet, (64, 1)) * np.linspace(0.2, 1.0, 64)[:, None]
    eg.scanning_frame(X_REF, Y_REF, 0.02, 0.02, wl * u.micron, flux, depth, 7.4325 * u.pixel / u.s,
                      400 * u.ms, add_dark=False, cosmic_rate=None, threads=2)
    run = eg._run
    rs = np.random.RandomState(5)
    o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, depth, X_REF, Y_REF,
                         0.02, 0.02, 7.4325 * 0.001, 400.0, rs, add_dark=False, threads=2, keep=True)
    i0, i1 = o['crop']
    ratio, sigl, sigh, sens, dwl = E.bin_tables(wl[i0:i1], cal['sens_wl_um'], cal['sens_val'])
    for got, want in zip(run.tables_host(), (ratio, sigl, sigh, sens, dwl)):
        assert _rel(got, want) < 1e-12
    tr = run.trace_host()
    xp, yp = run.positions_host()
    for s in (0, 17, run.N - 1):
        t = E.Trace(X_REF + o['jitter'][0][s], o['s_y_refs'][s] + o['jitter'][1][s], E.G141_TRACE, E.G141_WLSOL)
        want = np.array([t.x_ref, t.y_ref, t.m_t, t.c_t, t.m_w, t.c_w, t.m_wl, t.c_wl])
        assert _rel(tr[s], want) < 1e-9          # m_wl / c_wl difference of nearly equal numbers
        assert np.max(np.abs(xp[s] - (t.wl_to_x(wl[i0:i1]) - 379))) < 1e-6
        assert np.max(np.abs(yp[s] - (t.wl_to_y(wl[i0:i1]) - 379))) < 1e-6
    exp = run.expected_host()
    for s in (0, 33, run.N - 1):
        want = E.expected_counts(flux[i0:i1], depth[s][i0:i1], sens, dwl, o['durations'][s], None)
        assert _rel(exp[s], want) < 1e-12


@pytest.mark.parametrize("threads", [1, 2])
def test_staring_deterministic_bit_exact(calb_dir, threads):
    """Config 2: G141 staring, 256 subarray, NSAMP=5, every term on."""
    from wayne import units as u
    cal = harness.oracle_calibration()
    wl, flux, planet = harness.spectrum(level=2.0e-15)
    eg = _gen()
    _, mid, dur, ri = eg._gen_scanning_sample_times(1 * u.year)
    depth = np.tile(planet, (len(ri), 1))
    np.random.seed(1963)
    exp = eg.staring_frame(X_REF, Y_REF, 0.01, 0.01, wl * u.micron, flux, depth, mid, dur, ri,
                           0.5, 0.1, True, True, 11., 1.5 * u.count / u.s, 0.998, True, True, True,
                           True, True, True, None, threads)
    rs = np.random.RandomState(1963)
    o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, depth, X_REF, Y_REF,
                         0.01, 0.01, 0.0, 365.25 * 86400e3, rs, noise_mean=0.5, noise_std=0.1,
                         add_dark=True, cosmic_rate=11., sky_background=1.5, scale_factor=0.998,
                         threads=threads)
    assert len(exp.reads) == 5 and eg.photons == o['photons'] and o['photons'] > 1e5
    for r in range(5):
        got, want = exp.reads[r][0], o['reads'][r]
        assert got.shape == (266, 266)
        assert np.max(np.abs(got - want)) <= 1e-9 * max(1.0, np.abs(want).max()), r
    assert np.array_equal(exp.reads[0][0], o['reads'][0])       # zero read: bias + read noise, exact


def test_scan_deterministic_no_noise_exact(calb_dir):
    """Scan with every stochastic term off except the electrons: reads are a
    deterministic function of the int32 histograms -> compare bit for bit."""
    from wayne import units as u
    from wayne.trend_generators.scan_speed_varations import SSVSine
    cal = harness.oracle_calibration()
    wl, flux, planet = harness.spectrum(level=1.5e-14)
    eg = _gen()
    np.random.seed(7)
    exp = eg.scanning_frame(X_REF, Y_REF, 0.025, 0.025, wl * u.micron, flux, None,
                            7.4325 * u.pixel / u.s, 250 * u.ms, ssv_generator=SSVSine(1.5, 1.1, 0),
                            add_dark=False, cosmic_rate=None, sky_background=0 * u.count / u.s,
                            add_non_linear=False, add_read_noise=False, add_stellar_noise=False,
                            threads=3)
    rs = np.random.RandomState(7)
    o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, None, X_REF, Y_REF,
                         0.025, 0.025, 7.4325 * 0.001, 250.0, rs, ssv=(1.5, 1.1, 0), add_dark=False,
                         sky_background=0, add_non_linear=False, add_read_noise=False,
                         add_stellar_noise=False, threads=3)
    assert eg.photons == o['photons'] > 1e7
    for r in range(5):
        assert np.array_equal(exp.reads[r][0], o['reads'][r]), r


def test_scan_full_chain_compat(calb_dir):
    """Scan, all reductions on (sky, cosmics, gain, dark, non-linearity with the
    reference's global Newton stopping rule, clip, bias, read noise)."""
    from wayne import units as u
    from wayne.trend_generators.scan_speed_varations import SSVSine
    cal = harness.oracle_calibration()
    wl, flux, planet = harness.spectrum(level=2.5e-14)
    eg = _gen()
    _, mid, dur, ri = eg._gen_scanning_sample_times(300 * u.ms)
    depth = np.tile(planet, (len(mid), 1)) * np.linspace(0., 1., len(mid))[:, None]
    np.random.seed(42)
    exp = eg.scanning_frame(X_REF, Y_REF, 0.025, 0.025, wl * u.micron, flux, depth,
                            7.4325 * u.pixel / u.s, 300 * u.ms, mid, dur, ri,
                            ssv_generator=SSVSine(1.5, 1.1, 0), cosmic_rate=11.,
                            sky_background=5.5 * u.count / u.s, scale_factor=1.0012, threads=2)
    rs = np.random.RandomState(42)
    o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, depth, X_REF, Y_REF,
                         0.025, 0.025, 7.4325 * 0.001, 300.0, rs, ssv=(1.5, 1.1, 0), cosmic_rate=11.,
                         sky_background=5.5, scale_factor=1.0012, threads=2)
    assert eg.photons == o['photons']
    peak = max(np.abs(r).max() for r in o['reads'])
    assert min(o['newton_iters']) >= 2      # the non-linearity solve really iterates
    assert list(eg._run.newton_iters.cpu().numpy()[:4]) == o['newton_iters']
    for r in range(5):
        got, want = exp.reads[r][0], o['reads'][r]
        assert np.max(np.abs(got - want)) <= 1e-9 * peak, r
    hdr = exp.reads[2][1]
    assert abs(hdr['SAMPTIME'] - 7.624) < 1e-9 and hdr['CRPIX1'] == 0


def test_philox_mode_statistics(calb_dir):
    """Native mode: per-pixel mean / variance of the final read vs the oracle
    (which draws from numpy + rand_r) over seeds: within 3 sigma."""
    from wayne import units as u
    cal = harness.oracle_calibration()
    wl, flux, planet = harness.spectrum(level=4.0e-15, n_wl=300)
    n_seeds = 40
    kw = dict(add_dark=True, cosmic_rate=None, sky_background=2.0, threads=1)
    g_stack, o_stack = [], []
    for s in range(n_seeds):
        eg = _gen(rng='philox')
        _, mid, dur, ri = eg._gen_scanning_sample_times(1 * u.year)
        exp = eg.scanning_frame(X_REF, Y_REF, 0.01, 0.01, wl * u.micron, flux, None, 0 * u.pixel / u.s,
                                1 * u.year, mid, dur, ri, add_dark=True, cosmic_rate=None,
                                sky_background=2.0 * u.count / u.s, rng_key=(1963, s))
        g_stack.append(exp.reads[-1][0][5:-5, 5:-5])
        rs = np.random.RandomState(1000 + s)
        o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, None, X_REF, Y_REF,
                             0.01, 0.01, 0.0, 365.25 * 86400e3, rs, **kw)
        o_stack.append(o['reads'][-1][5:-5, 5:-5])
    g_stack, o_stack = np.array(g_stack), np.array(o_stack)
    # same exposure twice with the same key -> identical; different key -> different
    eg = _gen(rng='philox')
    _, mid, dur, ri = eg._gen_scanning_sample_times(1 * u.year)
    again = eg.scanning_frame(X_REF, Y_REF, 0.01, 0.01, wl * u.micron, flux, None, 0 * u.pixel / u.s,
                              1 * u.year, mid, dur, ri, add_dark=True, cosmic_rate=None,
                              sky_background=2.0 * u.count / u.s, rng_key=(1963, 0)).reads[-1][0]
    assert np.array_equal(again[5:-5, 5:-5], g_stack[0])
    # bin 8x8 blocks: mean and variance agree within 3 sigma of the sampling error
    def blocks(a):
        return a.reshape(a.shape[0], 32, 8, 32, 8).sum(axis=(2, 4))
    gb, ob = blocks(g_stack), blocks(o_stack)
    gm, om = gb.mean(0), ob.mean(0)
    gv, ov = gb.var(0, ddof=1), ob.var(0, ddof=1)
    se = np.sqrt((gv + ov) / n_seeds)
    z = (gm - om) / se
    assert np.abs(z).max() < 4.5 and abs(z.mean()) < 0.3 and (np.abs(z) > 3).mean() < 0.01
    # variance: compare on the ~25 trace blocks that carry the signal; each sample
    # variance (40 seeds) has a relative error of sqrt(2/39) = 0.23
    sig = om - np.median(om)
    bright = sig > 0.2 * sig.max()
    lr = np.log(gv[bright] / ov[bright])
    assert abs(lr.mean()) < 4 * np.sqrt(2 * 2.0 / 39 / bright.sum()) + 0.02, (lr.mean(), bright.sum())


def test_direct_accumulation_equals_window_gather_path(calb_dir, monkeypatch):
    """Native mode: the fused tile-flush accumulation (int64 fixed point) gives the
    same interval planes as the per-sub-sample windows + ordered gather, to the
    fixed-point resolution (2^-24 electron per add); Philox draws are identical."""
    from wayne import units as u
    from wayne_b200 import params
    wl, flux, planet = harness.spectrum(level=3.0e-14)
    out = {}
    for direct in (True, False):
        monkeypatch.setattr(params, 'direct_accumulation', direct)
        eg = _gen(rng='philox')
        exp = eg.scanning_frame(X_REF, Y_REF, 0.02, 0.02, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s,
                                200 * u.ms, add_dark=False, cosmic_rate=None,
                                sky_background=0 * u.count / u.s, add_non_linear=False,
                                add_read_noise=False, add_initial_bias=False, rng_key=(7, 7))
        out[direct] = (np.array([r[0] for r in exp.reads]), eg.photons)
    assert out[True][1] == out[False][1] > 1e6
    a, b = out[True][0], out[False][0]
    assert np.abs(a - b).max() < 1e-3          # DN; typical pixel values are 1e2..1e4
    assert np.abs(a - b).max() / b.max() < 1e-7
    # twice the same key -> bit-identical (integer atomics commute)
    monkeypatch.setattr(params, 'direct_accumulation', True)
    eg = _gen(rng='philox')
    exp = eg.scanning_frame(X_REF, Y_REF, 0.02, 0.02, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s,
                            200 * u.ms, add_dark=False, cosmic_rate=None, sky_background=0 * u.count / u.s,
                            add_non_linear=False, add_read_noise=False, add_initial_bias=False,
                            rng_key=(7, 7))
    assert np.array_equal(np.array([r[0] for r in exp.reads]), a)


def test_stochastic_mode_ks_and_moments_100_seeds(calb_dir):
    """north_star's stochastic criterion: per pixel, over 100 seeds, the native
    (Philox) exposures and the reference-stream oracle exposures come from the
    same distribution -- two-sample KS test per pixel, and mean / variance within
    3 sigma of their sampling errors."""
    from scipy import stats
    from wayne import units as u
    cal = harness.oracle_calibration()
    wl, flux, planet = harness.spectrum(level=1.5e-15, n_wl=256)
    n_seeds = 100
    G, O_ = [], []
    for s in range(n_seeds):
        eg = _gen(rng='philox')
        _, mid, dur, ri = eg._gen_scanning_sample_times(1 * u.year)
        exp = eg.scanning_frame(X_REF, Y_REF, 0.01, 0.01, wl * u.micron, flux, None, 0 * u.pixel / u.s,
                                1 * u.year, mid, dur, ri, cosmic_rate=None,
                                sky_background=2.0 * u.count / u.s, rng_key=(20170410, s))
        G.append(exp.reads[-1][0][70:95, 40:230].copy())      # the trace and its wings
        o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, None, X_REF, Y_REF,
                             0.01, 0.01, 0.0, 365.25 * 86400e3, np.random.RandomState(5000 + s),
                             cosmic_rate=None, sky_background=2.0, threads=1)
        O_.append(o['reads'][-1][70:95, 40:230])
    G, O_ = np.array(G), np.array(O_)
    npix = G[0].size
    # mean and variance per pixel
    gm, om = G.mean(0), O_.mean(0)
    gv, ov = G.var(0, ddof=1), O_.var(0, ddof=1)
    z_mean = (gm - om) / np.sqrt((gv + ov) / n_seeds)
    assert np.abs(z_mean).max() < 5.0 and (np.abs(z_mean) > 3).mean() < 0.01 and abs(z_mean.mean()) < 0.15
    # log variance ratio: sampling error sqrt(2/(n-1)) per estimate (Gaussian-ish pixels)
    z_var = np.log(gv / ov) / np.sqrt(2 * 2.0 / (n_seeds - 1))
    assert np.abs(z_var).max() < 5.5 and (np.abs(z_var) > 3).mean() < 0.02 and abs(z_var.mean()) < 0.2
    # two-sample KS per pixel: p-values must look uniform
    p = np.array([stats.ks_2samp(G[:, i, j], O_[:, i, j]).pvalue
                  for i in range(G.shape[1]) for j in range(G.shape[2])])
    assert p.size == npix
    assert (p < 0.01).mean() < 0.03 and (p < 1e-4).mean() < 0.002 and p.mean() > 0.40


def test_chebyshev_planet_signal_equals_array_signal(calb_dir):
    """The planet signal as a per-sub-sample Chebyshev expansion evaluated inside
    k_counts gives the same expected counts (and, with rounding instead of
    Poisson, the same reads) as the materialised [n_samples][n_wl] array."""
    from wayne import units as u
    from wayne_b200 import lightcurve as lc
    wl, flux, planet = harness.spectrum(level=2.0e-14)
    eg = _gen(rng='philox')
    _, mid, dur, ri = eg._gen_scanning_sample_times(250 * u.ms)
    t = 2456196.28836 - 0.0002 + np.asarray(u.value_in(mid, u.ms)) / 86400e3
    sig = lc.planet_signal(t, planet, [0.800627, -0.757066, 0.897268, -0.384804], 3.524746, 8.81, 0.0,
                           86.71, 0.0, 2456196.28836)
    arr = sig.to_array()
    assert arr.max() > 0.01
    outs = []
    for ps in (sig, arr):
        eg = _gen(rng='philox')
        exp = eg.scanning_frame(X_REF, Y_REF, 0.02, 0.02, wl * u.micron, flux, ps, 7.4325 * u.pixel / u.s,
                                250 * u.ms, mid, dur, ri, add_dark=False, cosmic_rate=None,
                                sky_background=0 * u.count / u.s, add_read_noise=False,
                                add_stellar_noise=False, rng_key=(3, 4))
        outs.append((np.array([r[0] for r in exp.reads]), eg.photons))
    assert outs[0][1] == outs[1][1] > 1e6
    assert np.array_equal(outs[0][0], outs[1][0])


def test_config2_explicit_photon_list_bit_exact(calb_dir):
    """BASELINE configs[1]: G141 staring, 256 subarray, NSAMP=5, the photon list
    supplied explicitly (counts from the shared numpy stream, normals as the A
    table of PSF()): int32 histograms and all five reads bit-exact."""
    from wayne import units as u
    cal = harness.oracle_calibration()
    wl, flux, planet = harness.spectrum(level=2.5e-15, n_wl=4096, lo=0.95, hi=1.8)
    eg = _gen()
    _, mid, dur, ri = eg._gen_scanning_sample_times(1 * u.year)

    def normals(i, n):
        return np.random.default_rng(300 + i).standard_normal(2 * n)

    kw = dict(add_dark=False, cosmic_rate=None, add_non_linear=False, add_read_noise=False)
    np.random.seed(2)
    exp = eg.scanning_frame(X_REF, Y_REF, 0.0, 0.0, wl * u.micron, flux, None, 0 * u.pixel / u.s, 1 * u.year,
                            mid, dur, ri, sky_background=0 * u.count / u.s, electron_normals=normals, **kw)
    o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, None, X_REF, Y_REF, 0.0, 0.0,
                         0.0, 365.25 * 86400e3, np.random.RandomState(2), sky_background=0, psf='normals',
                         normals=normals, **kw)
    assert eg.photons == o['photons'] > 5e5
    for r in range(5):
        assert np.array_equal(exp.reads[r][0], o['reads'][r]), r
    # every stochastic term on, supplied as host-drawn planes: still equal to the oracle
    np.random.seed(2)
    exp = eg.scanning_frame(X_REF, Y_REF, 0.0, 0.0, wl * u.micron, flux, None, 0 * u.pixel / u.s, 1 * u.year,
                            mid, dur, ri, sky_background=1.2 * u.count / u.s, cosmic_rate=11.,
                            electron_normals=normals)
    o = E.scanning_frame(cal, 'G141', 256, eg.read_times.to(u.s).value, wl, flux, None, X_REF, Y_REF, 0.0, 0.0,
                         0.0, 365.25 * 86400e3, np.random.RandomState(2), sky_background=1.2, cosmic_rate=11.,
                         psf='normals', normals=normals)
    for r in range(5):
        assert np.max(np.abs(exp.reads[r][0] - o['reads'][r])) <= 1e-9 * 4e4, r


def test_native_reads_kernel_equals_generic_kernel_full_chain(calb_dir, monkeypatch):
    """k_reads_native (shared-memory sky CDF window, FMA Newton with an fp32-seeded
    reciprocal, magic int64 -> double) against the generic k_reads<0, FAST> on one full
    native exposure with every term on except the sky (its draws are compared kernel to
    kernel in test_rng_gpu.py: the ramp's read intervals differ by a millisecond, where the
    throughput kernel re-uses its window plus a remainder draw): the same Philox draws, so
    the reads agree to the Newton stopping tolerance (both stop when a step is < 1e-3 DN)."""
    from wayne import units as u
    wl, flux, planet = harness.spectrum(level=3.0e-14)
    out = {}
    for generic in (False, True):
        if generic:
            monkeypatch.setenv('WB200_GENERIC_READS', '1')
        eg = _gen(rng='philox')
        exp = eg.scanning_frame(X_REF, Y_REF, 0.02, 0.02, wl * u.micron, flux, None, 7.4325 * u.pixel / u.s,
                                200 * u.ms, cosmic_rate=500., sky_background=0 * u.count / u.s,
                                rng_key=(9, 11))
        out[generic] = np.array([r[0] for r in exp.reads])
    monkeypatch.delenv('WB200_GENERIC_READS')
    a, b = out[False], out[True]
    assert b.max() > 1e3 and np.isfinite(a).all()
    assert np.abs(a - b).max() < 2e-3, np.abs(a - b).max()
