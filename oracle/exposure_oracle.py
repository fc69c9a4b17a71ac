"""TEST INFRASTRUCTURE ONLY -- float64 numpy restatement of the reference's
Python exposure path (the parity oracle; never imported by wayne_b200/).

The reference's Python layer cannot be imported here (Python-2 syntax plus
astropy / matplotlib / pandas-.ix), so this module restates it
function by function, each citing the reference lines it follows; astropy units
become explicit factors.  Calibration planes are passed in as plain
arrays IN THE DTYPE THE FITS FILES HOLD (float32), because several reference
expressions round through float32 (noted where they do).  The electron thrower
is NOT restated here: it is oracle/psf.py (C restatement, pinned bit-for-bit to
the unmodified reference kernel).

Pinning:
 (1) every KAT the reference's own tests hold for the path (trace / dispersion,
     bin widths, read times, visit trend: tests/test_grism.py, test_tools.py,
     test_detector.py, test_visit_trends.py) -- tests/test_oracle_kats.py;
 (2) golden vectors produced by EXECUTING the reference's own function bodies
     under Python 3 (ast extraction from /root/reference + Python-2 division
     shim; tests/golden/make_reference_goldens.py -> reference_goldens.npz):
     _SpectrumTrace / wavelength_calibration_coeffs, G141.get_flat_field (bit
     for bit, float32 storage included), WFC3_IR.apply_non_linearity (bit for
     bit), tools helpers, MinMaxPossionCosmicGenerator draw order, SSVSine,
     HookAndLongTermRamp, _gen_scanning_sample_times -- tests/test_reference_goldens.py.
 Not executable here and therefore pinned only by this restatement ("parity
 unpinned"): the astropy unit algebra of _gen_subsample / _flux_to_counts, the
 order of operations of _add_read_reductions and _post_exposure_reductions
 (their numpy-1.14 casting is written out explicitly below).

numpy version note: the reference pins numpy 1.14 (README.md:26), i.e. legacy
value-based casting.  Where that changes an expression's working precision it
is written out explicitly below instead of relying on the numpy in this image.
"""
import time as _time

import numpy as np

from . import psf as _psf

G141_TRACE = (1.96882, 9.09159E-5, -1.93260E-3, 1.04275E-2, -7.96978E-6,
              -2.49607E-6, 1.45963E-9, 1.39757E-8, 4.8494E-10)           # grism.py:756-758
G141_WLSOL = (8.95431E3, 9.35925E-2, 0, 4.51423E1, 3.17239E-4,
              2.17055E-3, -7.42504E-7, 3.48639E-7, 3.09213E-7)           # grism.py:769-770
G102_TRACE = (-3.55018E-1, 3.28722E-5, -1.44571E-3, 1.42852E-2, -7.20713E-6,
              -2.42542E-6, 1.18294E-9, 1.19634E-8, 6.17274E-10)          # grism.py:762-764
G102_WLSOL = (6.38738E3, 4.55507E-2, 0, 2.35716E1, 3.60396E-4,
              1.58739E-3, -4.25234E-7, -6.53726E-8, 0.)                  # grism.py:773-774
PSF_RATIO = (-0.25063428, 0.8332488, -0.80546074, 0.39896516)            # grism.py:85-86
PSF_SIGL = (0.69245668, -2.1043046, 2.22284446, -0.29689335)             # grism.py:87-88
PSF_SIGH = (2.90366189, -8.81859432, 8.96049229, 2.254503)               # grism.py:89-90
WL_LIMITS = {'G141': (0.988, 1.777), 'G102': (0.75, 1.2)}                # grism.py:94, 464
CONSTANT_GAIN = 2.35                                                     # detector.py:29
READ_NOISE = 14.1 / 2.35                                                 # detector.py:33
MIN_COUNTS, MAX_COUNTS = -20, 78000                                      # detector.py:26-27


# --------------------------------------------------------------------------
# trace / dispersion (grism.py:779-803, 491-506, 525-537, 553-602, 635-669)
# --------------------------------------------------------------------------
def wavelength_calibration_coeffs(x_ref, y_ref, a, b):
    m_t = a[3] + a[4] * x_ref + a[5] * y_ref + a[6] * x_ref ** 2 + a[7] * x_ref * y_ref + a[8] * y_ref ** 2
    c_t = a[0] + a[1] * x_ref + a[2] * y_ref
    m_w = b[3] + b[4] * x_ref + b[5] * y_ref + b[6] * x_ref ** 2 + b[7] * x_ref * y_ref + b[8] * y_ref ** 2
    c_w = (b[0] + b[1] * x_ref) + b[2] * y_ref
    return m_t, c_t, m_w, c_w


class Trace(object):
    """_SpectrumTrace for one source position."""

    def __init__(self, x_ref, y_ref, a, b):
        self.x_ref, self.y_ref = x_ref, y_ref
        self.m_t, self.c_t, self.m_w, self.c_w = wavelength_calibration_coeffs(x_ref, y_ref, a, b)
        x = np.array([x_ref + 10, x_ref + 20])                            # grism.py:591
        y = self.x_to_y(x)
        d = np.sqrt((y - y_ref) ** 2 + (x - x_ref) ** 2)
        wl = (self.m_w * d + self.c_w) * 1e-4                             # angstrom -> micron (:596-597)
        self.m_wl = (wl[1] - wl[0]) / (x[1] - x[0])
        self.c_wl = wl[0] - self.m_wl * x[0]

    def x_to_y(self, x):
        return self.m_t * (x - self.x_ref) + self.c_t + self.y_ref        # grism.py:537

    def wl_to_x(self, wl_um):
        return (wl_um - self.c_wl) / self.m_wl                            # grism.py:651

    def wl_to_y(self, wl_um):
        return self.x_to_y((wl_um - self.c_wl) / self.m_wl)               # grism.py:667-669


def pixel_wl(x_ref, y_ref, x, y, a, b):
    """G141.get_pixel_wl (grism.py:152-161): wavelength [angstrom] of pixel (x, y)."""
    m_t, _, m_w, c_w = wavelength_calibration_coeffs(x_ref, y_ref, a, b)
    a_t_i = 1 / m_t
    arr = y_ref - y + a_t_i * x_ref - a_t_i * x
    d = np.sqrt((arr * arr) / (a_t_i * a_t_i + 1))
    return m_w * d + c_w


# --------------------------------------------------------------------------
# helpers (tools.py:46-77, 106-128, 317-324)
# --------------------------------------------------------------------------
def crop_spectrum_ind(min_wl, max_wl, wl):
    wl = np.asarray(wl, dtype=float)
    d = wl - min_wl
    d[d < 0] = d.max()
    imin = d.argmin()
    d = wl - max_wl
    d[d > 0] = d.min()
    imax = d.argmax() + 1
    return imin, imax


def bin_centers_to_widths(centers):
    c = np.asarray(centers, dtype=float)
    left = (c - np.roll(c, 1)) / 2.
    left[0] = left[1]
    right = np.roll(left, -1)
    right[-1] = left[-1]
    return left + right


def crop_central_box(array, size):
    """tools.py:317-324; the reference's slice is EMPTY for size >= len (SURVEY
    B1: every 1024 full-frame call with sky / gain / flat raises there).  Defined
    behaviour shared with the product: identity when size >= len."""
    n = len(array)
    if size >= n:
        return array
    i = (n - size) // 2
    return array[i:-i, i:-i]


# --------------------------------------------------------------------------
# sample timing (exposure_generator.py:531-579, 517-529), SSV, visit trend
# --------------------------------------------------------------------------
def gen_scanning_sample_times(read_times_s, sample_rate_ms):
    read_times = np.asarray(read_times_s, dtype=float) * 1000.0
    read_index, starts_all, i, prev = [], [], -1, 0.
    for t in read_times:
        starts = np.arange(prev, t, sample_rate_ms)
        starts_all.append(starts)
        i += len(starts)
        read_index.append(i)
        prev = t
    starts = np.concatenate(starts_all)
    ends = np.roll(starts, -1)
    ends[-1] = read_times[-1]
    dur = ends - starts
    mid = starts + (dur / 2)
    return starts, mid, dur, read_index


def ssv_sine(y_mid_points, durations, stddev, period, phase):
    """SSVSine.get_subsample_exposure_times (scan_speed_varations.py:33-60), numeric phase."""
    zeroed = np.asarray(y_mid_points) - y_mid_points[0]
    scaling = (stddev / 100.) * np.sin((period * zeroed) + phase) + 1.
    return durations * scaling


def hook_and_long_term_ramp(t, t_0, a1, b1, b2, to):
    """visit_trends.py:43-57."""
    t = np.asarray(t, dtype=float)
    return (1 - a1 * (t - to)) * (1 - b1 * np.exp(-b2 * (t - t_0)))


# --------------------------------------------------------------------------
# per-sub-sample pieces (exposure_generator.py:581-647, 649-687; grism.py:111-118)
# --------------------------------------------------------------------------
def bin_tables(wl_um, sens_wl_um, sens_val):
    wl = np.asarray(wl_um, dtype=float)
    return (np.polyval(PSF_RATIO, wl), np.polyval(PSF_SIGL, wl), np.polyval(PSF_SIGH, wl),
            np.interp(wl, sens_wl_um, sens_val), bin_centers_to_widths(wl))


def expected_counts(flux, depth, sens, dwl_um, dur_ms, scale):
    """One sub-sample: exposure_generator.py:344-348 then :602-623 with the unit
    conversions as explicit factors in the order astropy applies them."""
    f = flux * (1. - depth) if depth is not None else flux
    rate = f * sens                       # ph / s / angstrom          (:602-603)
    rate = rate * dwl_um                  # x micron                   (:682)
    rate = rate * 1e4                     # .to(ph / s)                (:684)
    counts = rate * dur_ms                # x ms                       (:613)
    counts = counts * 1e-3                # .to(photon)
    if scale is not None:
        counts = counts * scale           # :620-621
    return counts


def flat_field_at_hits(x_ref, y_ref, subarray, frame, cal, a, b):
    """get_flat_field(..., indices=np.where(frame > 0)) applied to the frame
    (grism.py:359-385, 406-407; exposure_generator.py:641-645).

    The reference writes the flat values into np.ones_like(flat_f0), i.e. into
    an array of the FLAT FILE's dtype (float32): values are rounded to float32
    before they multiply the float64 frame.  For SUBARRAY 1024 the central crop
    is undefined in the reference (B1); defined behaviour: the value computed for
    hit pixel (r, c) multiplies pixel (r, c)."""
    f0, f1, f2, f3 = cal['flat']
    n = len(f0)
    rr, cc = np.where(frame > 0)
    off = (1014 - subarray) // 2                       # py2 floor division (:361-363)
    Y, X = rr + off, cc + off
    Y = np.where(Y < 0, Y + n, Y)                      # numpy negative-index wrap
    X = np.where(X < 0, X + n, X)
    m_t, _, m_w, c_w = wavelength_calibration_coeffs(x_ref, y_ref, a, b)
    a_t_i = 1 / m_t
    arr = y_ref - Y + a_t_i * x_ref - a_t_i * X
    d = np.sqrt((arr * arr) / (a_t_i * a_t_i + 1))
    wl = m_w * d + c_w
    w = (wl - cal['flat_wmin']) / (cal['flat_wmax'] - cal['flat_wmin'])
    w2 = w * w
    w3 = w2 * w
    val = f0[Y, X] + (f1[Y, X] * w) + (f2[Y, X] * w2) + (f3[Y, X] * w3)
    val = val.astype(f0.dtype.newbyteorder('=')).astype(np.float64)   # ones_like(flat_f0) storage
    out = frame.astype(np.float64).copy()
    out[rr, cc] *= val
    return out


# --------------------------------------------------------------------------
# per-read and post-exposure chains
# --------------------------------------------------------------------------
def master_sky_scaled(cal, L, bg_count):
    """grism.py:411-423 + exposure_generator.py:489-493: `master_sky *= bg_count`
    is an in-place multiply of the FITS float32 plane: under the reference's
    numpy (legacy casting) the scalar is taken to float32 and the product is
    float32."""
    sky = crop_central_box(cal['sky'], L)
    sky = np.asarray(sky, dtype=np.float32)
    return (sky * np.float32(bg_count)).astype(np.float64)


def gain_plane(cal, subarray):
    """detector.py:200-209: 2.35 / pfl[5:-5, 5:-5] in the file's float32."""
    pfl = np.asarray(cal['pfl'], dtype=np.float32)[5:-5, 5:-5]
    g = np.float32(CONSTANT_GAIN) / pfl
    return crop_central_box(g, subarray)


def add_bias_pixels(px):
    full = np.zeros((len(px) + 10, len(px) + 10))
    full[5:-5, 5:-5] = px
    return full


def apply_non_linearity(p, cal):
    """detector.py:318-350 (global stopping rule)."""
    n = len(p)
    half = len(cal['nl'][0]) // 2
    lo, hi = half - n // 2, half + n // 2
    c1, c2, c3, c4 = (np.asarray(c, dtype=np.float32)[lo:hi, lo:hi] for c in cal['nl'])
    u0 = p
    u1 = u0 * 0
    iters = 0
    for _ in range(10000):
        u1 = u0 - ((-p + u0 * (1 + c1 + u0 * (c2 + u0 * (c3 + c4 * u0)))) /
                   (1 + c1 + 2 * c2 * u0 + 3 * c3 * u0 * u0 + 4 * c4 * u0 * u0 * u0))
        iters += 1
        if (np.abs(u1 - u0) < 10 ** (-3)).all():
            break
        u0 = u1
    return u1, iters


def reset_reference_pixels(a):
    m = np.ones_like(a, dtype=bool)
    m[5:-5, 5:-5] = False
    a = a.copy()
    a[m] = 0.
    return a


def cosmic_frame(rs, rate, time, size, min_count=10000, max_count=35000):
    """MinMaxPossionCosmicGenerator(rate).cosmic_frame(time, size)
    (cosmic_rays.py:33-44, 70-139): Poisson number of hits at rate/1024^2 per
    pixel per second, uniform integer energies, uniform positions (rows first)."""
    n_hits = rs.poisson(rate / (1024. * 1024.) * (size * size) * time)
    energies = rs.randint(min_count, max_count, n_hits)
    rows = rs.randint(0, size, n_hits)
    cols = rs.randint(0, size, n_hits)
    frame = np.zeros((size, size))
    for k in range(n_hits):
        frame[rows[k], cols[k]] += energies[k]
    return frame


def gen_orbit_start_times_per_exp(time_array, obs_start_index):
    """visit_trends.py:60-73."""
    t = np.asarray(time_array, dtype=float)
    idx = list(obs_start_index) + [len(t)]
    t0 = np.zeros(len(t))
    for a, b in zip(idx[:-1], idx[1:]):
        t0[a:b] = t[a]
    return t0


def scanning_frame(cal, grism, subarray, read_times_s, wl_um, stellar_flux, planet_signal,
                   x_ref, y_ref, x_jitter, y_jitter, scan_speed_px_per_ms, sample_rate_ms,
                   rs, ssv=None, noise_mean=False, noise_std=False, add_dark=True, add_flat=True,
                   cosmic_rate=None, sky_background=1.0, scale_factor=None,
                   add_gain_variations=True, add_non_linear=True, clip_values_det_limits=True,
                   add_read_noise=True, add_stellar_noise=True, add_initial_bias=True, threads=2,
                   sample_times=None, psf='port', normals=None, keep=False):
    """ExposureGenerator.scanning_frame (exposure_generator.py:178-405) followed by
    _post_exposure_reductions (:407-444).

    rs: np.random.RandomState standing in for numpy's global state.
    ssv: None or (stddev, period, phase) of an SSVSine.
    psf: 'port' (C restatement) or 'reference' (unmodified kernel) with the
         rand_r stream, or 'normals' with normals = callable(sample, ssum) -> A[2*ssum].
    Returns dict(reads=[NSAMP float64 F x F], ...diagnostics).
    """
    a, b = (G141_TRACE, G141_WLSOL) if grism == 'G141' else (G102_TRACE, G102_WLSOL)
    S = subarray
    L = 1014 if S == 1024 else S                                 # detector.py:116-119
    F = min(S + 10, 1024)                                        # detector.py:121-124
    read_times_s = np.asarray(read_times_s, dtype=float)
    if sample_times is None:
        _, mid, dur, read_index = gen_scanning_sample_times(read_times_s, sample_rate_ms)
    else:
        mid, dur, read_index = sample_times
    s_y_refs = y_ref + mid * scan_speed_px_per_ms               # :258, :527
    if ssv is not None:
        dur = ssv_sine(s_y_refs, dur, *ssv)                     # :262-273

    zero_read = np.zeros((F, F))                                 # :446-466
    if S == 256 and add_initial_bias:
        zero_read = zero_read + cal['bias256']
    reads = [zero_read.copy()]
    cumulative = np.zeros((F, F))
    pixel_array = np.zeros((L, L))

    N = len(mid)
    seeds = rs.randint(0, 100000, N)                             # :327
    jx = rs.normal(0, x_jitter, N)                               # :328
    jy = rs.normal(0, y_jitter, N)                               # :329

    lim = WL_LIMITS[grism]
    i0, i1 = crop_spectrum_ind(lim[0], lim[1], wl_um)            # :332-334
    s_wl = np.asarray(wl_um, dtype=float)[i0:i1]
    ratio, sigl, sigh, sens, dwl = bin_tables(s_wl, cal['sens_wl_um'], cal['sens_val'])
    sub_scale = 507 - (S // 2)                                   # :630 (py2 int division)
    psf_fn = _psf.psf_reference if psf == 'reference' else _psf.psf_port

    all_counts = np.zeros((N, len(s_wl)), dtype=np.int64) if keep else None
    photons = 0
    read_num, prev_t = 0, 0.0
    timing = {'subsamples': 0.0, 'reads': 0.0, 'post': 0.0}
    for i in range(N):
        _t0 = _time.perf_counter()
        depth = planet_signal[i][i0:i1] if planet_signal is not None else None
        sx, sy = x_ref + jx[i], s_y_refs[i] + jy[i]              # :355-356
        tr = Trace(sx, sy, a, b)
        x_pos, y_pos = tr.wl_to_x(s_wl), tr.wl_to_y(s_wl)        # :593-595
        counts = expected_counts(stellar_flux[i0:i1], depth, sens, dwl, dur[i], scale_factor)
        counts = rs.poisson(counts) if add_stellar_noise else np.round(counts)   # :625-628
        if keep:
            all_counts[i] = counts
        icounts = np.asarray(counts).astype(np.int32)            # pyparallel.pyx:24-25 (C int)
        photons += int(icounts.sum())
        x_sub, y_sub = x_pos - sub_scale, y_pos - sub_scale      # :631-632
        if psf == 'normals':
            A = normals(i, int(icounts.sum()))
            frame = _psf.bin_electrons(icounts, x_sub, y_sub, ratio, sigl, sigh, L, L, A)
        else:
            frame = psf_fn(icounts, x_sub, y_sub, ratio, sigl, sigh, L, L, int(seeds[i]), threads)
        frame = frame.astype(np.float64)                         # pyparallel.pyx:31-34
        if add_flat:
            frame = flat_field_at_hits(sx, sy, S, frame, cal, a, b)   # :641-645
        pixel_array += frame                                     # :359
        timing['subsamples'] += _time.perf_counter() - _t0

        if i in read_index:                                      # :361
            _t0 = _time.perf_counter()
            dt = read_times_s[read_num] - prev_t                 # :363-365
            px = pixel_array
            if noise_mean and noise_std:                         # :477-484
                px = px + rs.normal(noise_mean * dt, noise_std * dt, (L, L))
            if sky_background:                                   # :488-495
                px = px + rs.poisson(master_sky_scaled(cal, L, sky_background * dt))
            if cosmic_rate is not None:                          # :498-505
                px = px + cosmic_frame(rs, cosmic_rate, dt, L)
            if add_gain_variations:                              # :507-511
                px = px / gain_plane(cal, S)
            else:
                px = px / CONSTANT_GAIN
            cumulative = cumulative + add_bias_pixels(px)        # :513, :378
            reads.append(cumulative.copy())                      # :381-382
            prev_t = read_times_s[read_num]
            read_num += 1
            pixel_array = np.zeros((L, L))                       # :388
            timing['reads'] += _time.perf_counter() - _t0
    _t0 = _time.perf_counter()
    assert len(reads) == len(read_times_s) + 1                   # :397

    newton_iters = []
    if add_dark:                                                 # exposure.py:70-80, detector.py:151-191
        for r in range(1, len(reads)):
            dark, err = cal['dark'][r + 1]                       # read_NSAMP = i + 1 -> ext -(NSAMP)*5
            err = np.where(err > 0, err, np.float32(0.00001))
            reads[r] = reads[r] + rs.normal(dark, err)
    if add_non_linear:                                           # exposure.py:49-59
        for r in range(1, len(reads)):
            reads[r], it = apply_non_linearity(reads[r], cal)
            newton_iters.append(it)
    if clip_values_det_limits:                                   # exposure.py:82-92
        reads = [np.clip(r, MIN_COUNTS, MAX_COUNTS) for r in reads]
    reads = [reset_reference_pixels(r) for r in reads]           # exposure.py:122-131
    for r in range(1, len(reads)):                               # exposure.py:94-104
        reads[r] = reads[r] + reads[0]
    if add_read_noise:                                           # exposure.py:61-68, detector.py:193-198
        reads = [rs.normal(r, READ_NOISE) for r in reads]
    timing['post'] = _time.perf_counter() - _t0
    return dict(reads=reads, timing=timing, seeds=seeds, jitter=(jx, jy), durations=dur, s_y_refs=s_y_refs,
                read_index=read_index, counts=all_counts, photons=photons,
                newton_iters=newton_iters, wl=s_wl, crop=(i0, i1))
