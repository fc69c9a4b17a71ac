"""oracle/ and the host mirror against golden vectors produced by executing the
REFERENCE'S OWN function bodies (tests/golden/make_reference_goldens.py: ast
extraction from /root/reference + Python-2 division shim).  This pins the
trace / dispersion math, the flat-field expression (incl. its float32 storage),
the Newton non-linearity solve, the helper functions, the cosmic-ray generator's
draw order, SSVSine, the visit trend and the sample timing."""
import os

import numpy as np
import numpy.testing as npt
import pytest

from oracle import exposure_oracle as E
from wayne import detector, grism, tools
from wayne import units as u
from wayne.exposure_generator import ExposureGenerator
from wayne.trend_generators import cosmic_rays, scan_speed_varations, visit_trends

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_goldens.npz'))


@pytest.fixture(scope='module')
def planes():
    """The generator script's random planes, redrawn in the same order."""
    rng = np.random.default_rng(int(G['flat_seed']))
    flat = [(1 + 0.01 * rng.standard_normal((1014, 1014))).astype('>f4')]
    flat += [(0.005 * rng.standard_normal((1014, 1014))).astype('>f4') for _ in range(3)]
    rng.integers(60, 110, 400)
    rng.integers(30, 230, 400)
    rng.uniform(0.85, 1.9, 300)
    c2 = (6.4e-7 * (1 + 0.05 * rng.standard_normal((1024, 1024)))).astype('>f4')
    c3 = (1e-12 * rng.standard_normal((1024, 1024))).astype('>f4')
    zero = np.zeros((1024, 1024), '>f4')
    return dict(flat=flat, nl=(zero, c2, c3, zero))


def test_trace_and_dispersion():
    wl = G['trace_wl']
    rows = G['trace_rows']
    sets = [(E.G141_TRACE, E.G141_WLSOL, grism.G141_Trace)] * 4 + [(E.G102_TRACE, E.G102_WLSOL, grism.G102_Trace)] * 4
    for row, (a, b, mirror) in zip(rows, sets):
        x, y = row[0], row[1]
        for tr in (E.Trace(x, y, a, b), mirror(x, y)):
            got = np.concatenate([[x, y, tr.m_t, tr.c_t, tr.m_w, tr.c_w, tr.m_wl, tr.c_wl],
                                  tr.wl_to_x(wl), tr.wl_to_y(wl)])
            npt.assert_allclose(got, row, rtol=2e-12, atol=0)


def test_flat_field_expression(planes):
    idx = (G['flat_idx_r'].astype(int), G['flat_idx_c'].astype(int))
    frame = np.zeros((256, 256))
    frame[idx] = 1
    cal = {'flat': planes['flat'], 'flat_wmin': 9880.0, 'flat_wmax': 17770.0}
    got = E.flat_field_at_hits(404.497, 457.427, 256, frame, cal, E.G141_TRACE, E.G141_WLSOL)
    assert np.array_equal(got[idx], G['flat_values'])          # bit for bit, float32 storage included
    assert float(G['flat_off_pixel']) == 1.0 and got[0, 0] == 0.0
    g = grism.G141()
    g._flat = {'wmin': 9880.0, 'wmax': 17770.0, 'f': tuple(planes['flat'])}
    ff = g.get_flat_field(404.497, 457.427, 256, idx)
    assert ff.shape == (256, 256) and ff[0, 0] == 1.0
    # the mirror keeps float64; the reference stores float32 (the device kernel rounds, flat_f32)
    npt.assert_allclose(ff[idx], G['flat_values'], rtol=1e-7)
    full = g.get_flat_field(404.497, 457.427, 256, None)
    npt.assert_allclose(full[::17, ::19], G['flat_full_sample'], rtol=1e-7)


def test_tools_helpers():
    wl = G['tools_wl']
    for fn in (E.crop_spectrum_ind, tools.crop_spectrum_ind):
        assert tuple(int(v) for v in fn(0.988, 1.777, wl)) == tuple(int(v) for v in G['tools_crop_ind'])
    for fn in (E.bin_centers_to_widths, tools.bin_centers_to_widths):
        assert np.array_equal(fn(wl), G['tools_widths'])
    assert np.array_equal(tools.bin_centers_to_edges(wl), G['tools_edges'])
    a = np.arange(66 * 66, dtype=float).reshape(66, 66)
    for fn in (E.crop_central_box, tools.crop_central_box):
        assert np.array_equal(fn(a, 32), G['tools_crop_box'])
    from wayne import observation
    assert observation.detect_orbits(np.array([1.0, 1.01, 1.05, 1.06, 1.12, 1.121])) == list(G['tools_orbits'])


def test_non_linearity_and_detector_geometry(planes):
    px = G['nl_input']
    got, iters = E.apply_non_linearity(px, {'nl': planes['nl']})
    assert np.array_equal(got, G['nl_output']) and iters >= 2
    det = detector.WFC3_IR()
    det._nl = planes['nl']
    assert np.array_equal(det.apply_non_linearity(px), G['nl_output'])
    # the seven float64 planes handed to the CUDA Newton solve reproduce the same iteration
    b0, c2, c3, c4, d2, d3, d4 = (p[464:560, 464:560] for p in det.non_linear_planes(1024))
    u0 = px
    for _ in range(iters):
        u0 = u0 - ((-px + u0 * (b0 + u0 * (c2 + u0 * (c3 + c4 * u0)))) /
                   (b0 + d2 * u0 + d3 * u0 * u0 + d4 * u0 * u0 * u0))
    assert np.array_equal(u0, G['nl_output'])
    assert np.array_equal(det.add_bias_pixels(np.arange(64 * 64, dtype=float).reshape(64, 64)), G['bias_pixels'])
    shapes = [det.gen_pixel_array(s, ls).shape for s in (1024, 512, 256) for ls in (True, False)]
    assert np.array_equal(np.array(shapes), G['pixel_array_shapes'])
    assert [det.num_exp_per_buffer(n, s) for n, s in ((5, 256), (15, 1024), (16, 64), (3, 512))] == \
        list(G['exp_per_buffer'])


def test_cosmic_generator_draw_order():
    want = G['cosmic_frame_seed1963_t20_s256']
    assert want.sum() > 0
    assert np.array_equal(E.cosmic_frame(np.random.RandomState(1963), 11., 20.0, 256), want)
    gen = cosmic_rays.MinMaxPossionCosmicGenerator(11., 10000, 35000)
    np.random.seed(1963)
    assert np.array_equal(gen.cosmic_frame(20.0, 256), want)
    np.random.seed(1963)
    rows, cols, en = gen.cosmic_hits(20.0, 256)
    dense = np.zeros((256, 256))
    np.add.at(dense, (rows, cols), en)
    assert np.array_equal(dense, want)
    assert gen._rate_full_frame_to_size(11., 64) == float(G['cosmic_rate_64'])


def test_ssv_and_visit_trend():
    y, dur = G['ssv_y'], G['ssv_dur']
    assert np.array_equal(E.ssv_sine(y, dur, 1.5, 1.1, 0), G['ssv_out'])
    assert np.array_equal(E.ssv_sine(y, dur, 2.5, 0.7, 1.3), G['ssv_out_phase'])
    out = scan_speed_varations.SSVSine(1.5, 1.1, 0).get_subsample_exposure_times(y, dur * u.ms, None, None)
    assert np.array_equal(np.asarray(u.value_in(out, u.ms)), G['ssv_out'])
    t, t0 = G['trend_t'], G['trend_t0']
    assert np.array_equal(E.gen_orbit_start_times_per_exp(t, [0, 13, 27]), t0)
    assert np.array_equal(visit_trends.gen_orbit_start_times_per_exp(t, [0, 13, 27]), t0)
    assert np.array_equal(E.hook_and_long_term_ramp(t, t0, 0.005, 0.0011, 400, 2456196.28836), G['trend_factors'])
    assert np.array_equal(visit_trends.HookAndLongTermRamp.ramp_model(t, t0, 0.005, 0.0011, 400, 2456196.28836),
                          G['trend_factors'])


@pytest.mark.parametrize("tag,nsamp,seq,sub,rate", [('c1', 5, 'SPARS10', 256, 10.0),
                                                    ('rapid1024', 15, 'RAPID', 1024, 10.0),
                                                    ('staring', 5, 'SPARS10', 256, 365.25 * 86400e3)])
def test_sample_timing(tag, nsamp, seq, sub, rate):
    rt = G['times_%s_rt' % tag]
    _, mid, dur, ri = E.gen_scanning_sample_times(rt, rate)
    assert np.array_equal(mid, G['times_%s_mid' % tag]) and np.array_equal(dur, G['times_%s_dur' % tag])
    assert list(ri) == list(G['times_%s_ri' % tag])
    npt.assert_allclose(457.4 + mid * (7.4325 * 0.001), G['times_%s_yref' % tag], rtol=1e-15)
    eg = ExposureGenerator(detector.WFC3_IR(), grism.G141(), nsamp, seq, sub, None)
    if np.allclose(eg.read_times.to(u.s).value, rt):      # the mode table agrees with the golden's read times
        _, m, d, r = eg._gen_scanning_sample_times(rate * u.ms)
        npt.assert_allclose(u.value_in(m, u.ms), G['times_%s_mid' % tag], rtol=1e-14)
        npt.assert_allclose(u.value_in(d, u.ms), G['times_%s_dur' % tag], rtol=1e-12, atol=1e-9)
        assert list(r) == list(G['times_%s_ri' % tag])


@pytest.mark.parametrize("tag", ['a', 'blip', 'c'])
def test_ssv_modulated_sine(tag):
    """SSVModulatedSine.get_subsample_exposure_times (scan_speed_varations.py:83-171) against the
    executed reference body: same numpy RandomState seed -> the same sub-sample durations to the
    last bit, the same read indexes, and the stream left at the same position (i.e. the mirror
    consumes numpy's global state draw for draw, blips and the 1 us redistribution included)."""
    seed, amp, per, blip, rate = G['ssvm_%s_args' % tag]
    rt = G['ssvm_%s_rt' % tag]
    np.random.seed(int(seed))
    gen = scan_speed_varations.SSVModulatedSine(amp, per, blip)
    dur, ri = gen.get_subsample_exposure_times(None, None, rt * u.s, rate * u.s)
    assert np.array_equal(np.asarray(u.value_in(dur, u.ms)), G['ssvm_%s_dur_ms' % tag])
    assert [int(i) for i in ri] == [int(i) for i in G['ssvm_%s_ri' % tag]]
    assert np.random.random() == float(G['ssvm_%s_next_random' % tag])
    # what the body promises: the exposure time is kept to the microsecond
    assert abs(np.sum(G['ssvm_%s_dur_ms' % tag]) * 1e-3 - rt[-1]) < 2e-6


@pytest.mark.parametrize("tag", ['256', '512', '64'])
def test_direct_image_values(tag):
    """ExposureGenerator.direct_image (exposure_generator.py:83-144) against the executed
    reference body: zero read F x F of zeros, second read the S x S Gaussian (B11: no border),
    CRPIX1 -5, exp_info switches."""
    sub, xr, yr = G['direct_%s_args' % tag]
    sub = int(sub)
    seq = {256: 'SPARS10', 512: 'SPARS25', 64: 'RAPID'}[sub]
    eg = ExposureGenerator(detector.WFC3_IR(), grism.G141(), 3, seq, sub, None)
    ex = eg.direct_image(float(xr), float(yr))
    assert len(ex.reads) == 2
    zero, img = ex.reads[0][0], ex.reads[1][0]
    assert tuple(zero.shape) == tuple(G['direct_%s_zero_shape' % tag]) and np.abs(zero).sum() == 0.0
    assert float(G['direct_%s_zero_sum' % tag]) == 0.0
    assert np.array_equal(img, G['direct_%s_image' % tag])
    assert img.shape == (sub, sub) and img.max() > 5000
    assert ex.reads[1][1]['CRPIX1'] == int(G['direct_%s_crpix1' % tag]) == -5
    info = eg.exp_info
    assert [info['NSAMP'], info['cosmic_rate'], info['scale_factor']] == list(G['direct_%s_info' % tag])
    assert info['OBSTYPE'] + '|' + info['SAMP-SEQ'] == str(G['direct_%s_obstype' % tag])
