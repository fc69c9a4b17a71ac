"""Distribution tests of the native (Philox) random streams, through the C ABI:
the Poisson sampler of wb200_counts / wb200_reads, the fp32 Box-Muller of the
photon kernel and the fp64 normals of the per-pixel pass.  Stochastic-mode
equivalence with the reference rests on these being the SAME distributions as
np.random.poisson / np.random.normal / the reference's rand_r Box-Muller."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


def _poisson_draws(lam, n_samples=256, n_bins=4096, key=(7, 9), dur_scale=None, flat=True):
    import torch
    from wayne_b200 import _lib
    dev = torch.device('cuda', 0)
    f = torch.full((n_bins,), float(lam), dtype=torch.float64, device=dev)
    one = torch.ones((n_bins,), dtype=torch.float64, device=dev)
    dwl = torch.full((n_bins,), 1e-4, dtype=torch.float64, device=dev)     # x 1e4 -> 1
    dur = torch.full((n_samples,), 1000.0, dtype=torch.float64, device=dev)  # x 1e-3 -> 1
    if dur_scale is not None:
        dur = dur * torch.as_tensor(np.asarray(dur_scale, dtype=np.float64), device=dev)
    counts = torch.empty((n_samples, n_bins), dtype=torch.int32, device=dev)
    totals = torch.empty((n_samples,), dtype=torch.int64, device=dev)
    _lib.check(_lib.lib.wb200_counts(n_samples, n_bins, _p(f), None, 0, _p(one), _p(dwl), _p(dur), 1.0,
                                     _lib.COUNT_POISSON, key[0], key[1], None, _p(counts), _p(totals),
                                     None), 'wb200_counts')
    c = counts.cpu().numpy()
    assert np.array_equal(c.sum(axis=1), totals.cpu().numpy())
    return c.ravel() if flat else c


@pytest.mark.parametrize("lam", [0.05, 0.7, 3.0, 9.99, 10.0, 14.7, 39.9, 40.0, 60.0, 243.0, 4000.0, 2.5e5])
def test_poisson_sampler_distribution(lam):
    x = _poisson_draws(lam)
    n = x.size
    assert abs(x.mean() - lam) < 5 * np.sqrt(lam / n)
    # var of the sample variance of a Poisson: (lam + 2 lam^2 (n/(n-1))) / n
    assert abs(x.var() - lam) < 5 * np.sqrt((lam + 2 * lam * lam) / n)
    # chi-square against the exact pmf on bins holding >= 50 expected draws
    lo = int(max(0, np.floor(lam - 6 * np.sqrt(lam) - 2)))
    hi = int(np.ceil(lam + 6 * np.sqrt(lam) + 3))
    step = max(1, (hi - lo) // 200)
    edges = np.arange(lo, hi + step, step)
    obs = np.histogram(x, bins=np.append(edges, edges[-1] + step) - 0.5)[0][:-1]
    cdf = stats.poisson.cdf(edges - 1, lam)
    exp = np.diff(np.append(cdf, stats.poisson.cdf(edges[-1] + step - 1, lam)))[: len(obs)] * n
    sel = exp >= 50
    chi2 = ((obs[sel] - exp[sel]) ** 2 / exp[sel]).sum()
    dof = sel.sum() - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)


def test_poisson_independent_of_geometry_and_key():
    a = _poisson_draws(37.0, 64, 1024, key=(1, 2))
    b = _poisson_draws(37.0, 64, 1024, key=(1, 2))
    c = _poisson_draws(37.0, 64, 1024, key=(1, 3))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # a sub-block of a bigger launch draws the same numbers (counter = (bin, sample))
    big = _poisson_draws(37.0, 128, 1024, key=(1, 2)).reshape(128, 1024)
    assert np.array_equal(big[:64].ravel(), a)


def test_photon_normals_distribution():
    """Marginal x / y distributions of electrons thrown from one bin: truncated
    normal binned at integer pixels, chi-square against the exact cell
    probabilities; narrow/wide mixture fractions respected."""
    from wayne_b200 import pyparallel
    n = 4_000_000
    sig_l, sig_h, ratio = 9.0, 31.0, 0.25
    x0, y0 = 512.3, 500.7
    frame = pyparallel.psf_frame([n], [x0], [y0], [ratio], [sig_l], [sig_h], 1014, 1014, test=12345,
                                 rng='philox')
    assert frame.sum() > 0.999 * n
    for axis, c0 in ((0, x0), (1, y0)):
        obs = frame.sum(axis=axis).astype(float)
        k = np.arange(1015)
        cdf = ratio * stats.norm.cdf(k, c0, sig_h) + (1 - ratio) * stats.norm.cdf(k, c0, sig_l)
        exp = np.diff(cdf) * n
        sel = exp >= 100
        chi2 = ((obs[sel] - exp[sel]) ** 2 / exp[sel]).sum()
        dof = sel.sum() - 1
        assert chi2 < dof + 6 * np.sqrt(2 * dof), (axis, chi2, dof)
    # x and y of one electron are independent: correlation of the 2-D histogram ~ 0
    yy, xx = np.mgrid[0:1014, 0:1014]
    w = frame / frame.sum()
    mx, my = (w * xx).sum(), (w * yy).sum()
    cov = (w * (xx - mx) * (yy - my)).sum()
    sx = np.sqrt((w * (xx - mx) ** 2).sum())
    sy = np.sqrt((w * (yy - my) ** 2).sum())
    assert abs(cov / (sx * sy)) < 5 / np.sqrt(n)


def _reads_noise_only(read_noise=0.0, sky_rate=0.0, noise=(0.0, 0.0), F=266, R=3, key=(3, 4), fast=0,
                      dts=(1.0, 2.0, 4.0), sky_level=1.0):
    import torch
    from wayne_b200 import _lib
    dev = torch.device('cuda', 0)
    a = _lib.ReadsArgs()
    a.n_reads, a.F, a.border = R, F, 5
    a.add_read_noise = 1 if read_noise else 0
    a.add_sky = 1 if sky_rate else 0
    a.add_noise = 1 if noise[1] else 0
    a.noise_mean, a.noise_std = noise
    a.sky_rate, a.sky_f32 = sky_rate, 1
    a.const_gain = 1.0
    a.read_noise = read_noise
    a.fast_math = fast
    a.key0, a.key1 = key
    dt = torch.tensor(list(dts)[:R], dtype=torch.float64, device=dev)
    acc = torch.zeros((R, F, F), dtype=torch.float64, device=dev)
    sky = torch.full((F, F), float(sky_level), dtype=torch.float64, device=dev)
    out = torch.empty((R + 1, F, F), dtype=torch.float64, device=dev)
    iters = torch.zeros((16,), dtype=torch.int32, device=dev)
    a.d_dt, a.d_acc, a.d_sky, a.d_out = dt.data_ptr(), acc.data_ptr(), sky.data_ptr(), out.data_ptr()
    a.d_newton_iters = iters.data_ptr()
    _lib.check(_lib.lib.wb200_reads(C.byref(a), None), 'wb200_reads')
    return out.cpu().numpy()


@pytest.mark.parametrize("fast", [0, 1])
def test_read_noise_normals(fast):
    out = _reads_noise_only(read_noise=6.0, fast=fast)
    for r in range(4):
        z = out[r] / 6.0
        assert stats.kstest(z.ravel(), 'norm').pvalue > 1e-4
        # the two pixels of a thread share one Philox call: still independent
        assert abs(np.corrcoef(z[:, 0::2].ravel(), z[:, 1::2].ravel())[0, 1]) < 5 / np.sqrt(z.size / 2)
    assert abs(np.corrcoef(out[0].ravel(), out[1].ravel())[0, 1]) < 0.02


@pytest.mark.parametrize("fast", [0, 1])
def test_sky_poisson_and_background_noise(fast):
    out = _reads_noise_only(sky_rate=14.7, noise=(0.5, 0.25), fast=fast)
    inner = out[:, 5:-5, 5:-5]
    d1 = inner[1]                       # first interval: Poisson(14.7 * 1) + N(0.5, 0.25)
    n = d1.size
    assert abs(d1.mean() - (14.7 + 0.5)) < 5 * np.sqrt((14.7 + 0.0625) / n)
    assert abs(d1.var() - (14.7 + 0.0625)) < 0.35
    d3 = inner[3] - inner[2]            # third interval, dt = 4
    assert abs(d3.mean() - (14.7 * 4 + 2.0)) < 5 * np.sqrt((58.8 + 1.0) / n)
    assert abs(d3.var() - (58.8 + 1.0)) < 1.5
    assert np.all(out[:, :5, :] == 0) and np.all(out[:, :, -5:] == 0)   # reference pixels stay 0


@pytest.mark.parametrize("fast", [0, 1])
def test_sky_poisson_has_no_outliers(fast):
    """30 M sky draws at lam = 14.7: nothing beyond the 1e-9 quantile region
    (guards the fp32 inversion's u -> 1 and saturated-sum corner cases)."""
    worst = 0.0
    for key in range(6):
        out = _reads_noise_only(sky_rate=14.7, F=1024, R=3, key=(11, key), fast=fast)
        inner = out[1, 5:-5, 5:-5]                       # first interval: Poisson(14.7)
        worst = max(worst, float(inner.max()))
        assert inner.min() >= 0
    assert worst <= 14.7 + 8.5 * np.sqrt(14.7), worst    # P(X > 47) ~ 1e-11 per draw


def _chi2_poisson(x, lam):
    n = x.size
    lo = int(max(0, np.floor(lam - 6 * np.sqrt(lam) - 2)))
    hi = int(np.ceil(lam + 6 * np.sqrt(lam) + 3))
    edges = np.arange(lo, hi + 1)
    obs = np.histogram(x, bins=np.append(edges, edges[-1] + 1) - 0.5)[0][:-1]
    cdf = stats.poisson.cdf(edges - 1, lam)
    exp = np.diff(np.append(cdf, stats.poisson.cdf(edges[-1], lam)))[: len(obs)] * n
    sel = exp >= 50
    chi2 = ((obs[sel] - exp[sel]) ** 2 / exp[sel]).sum()
    dof = sel.sum() - 1
    return chi2, dof


@pytest.mark.parametrize("lam", [0.3, 4.0, 14.7, 23.9, 24.1, 58.8, 131.0])
def test_native_reads_sky_poisson_pmf(lam):
    """k_reads_native's shared-memory CDF window (one draw below a mean of 24, the sum of
    m window draws above): chi-square of 1 M draws per read interval against the exact pmf,
    for equal, changing and nearly equal interval lengths (the window is rebuilt when the mean
    drops or grows by more than 0.25, and serves means up to 0.25 larger together with a
    Poisson draw of the remainder)."""
    for dts in ((1.0, 1.0, 1.0), (1.0, 0.5, 1.0), (1.0, 1.01, 1.012)):
        out = _reads_noise_only(sky_rate=lam, F=1034, R=3, key=(5, 6), fast=1, dts=dts, sky_level=1.0)
        inner = out[:, 5:-5, 5:-5]
        for r in range(3):
            x = (inner[r + 1] - inner[r]).ravel()
            assert np.all(x == np.rint(x)) and x.min() >= 0
            chi2, dof = _chi2_poisson(x, lam * dts[r])
            assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (lam, dts, r, chi2, dof)
        # draws of different read intervals are independent
        a, b = (inner[1] - inner[0]).ravel(), (inner[2] - inner[1]).ravel()
        assert abs(np.corrcoef(a, b)[0, 1]) < 5 / np.sqrt(a.size)


def test_native_reads_equal_generic_fast_kernel(monkeypatch):
    """The throughput kernel keeps the Philox counters of k_reads<0, FAST>: same sky draws
    (window search == serial walk, for every draw) and the same normals."""
    kw = dict(read_noise=6.0, sky_rate=14.7, noise=(0.5, 0.25), F=522, R=3, key=(21, 22), fast=1,
              dts=(1.0, 1.0, 1.5))          # means below 24: one window draw == the serial walk
    new = _reads_noise_only(**kw)
    monkeypatch.setenv('WB200_GENERIC_READS', '1')
    old = _reads_noise_only(**kw)
    monkeypatch.delenv('WB200_GENERIC_READS')
    assert np.max(np.abs(new - old)) < 1e-4          # ftz / non-ftz SFU forms only
    sky_only = dict(sky_rate=9.3, F=522, R=3, key=(23, 24), fast=1, dts=(1.0, 2.0, 2.5))
    new = _reads_noise_only(**sky_only)
    monkeypatch.setenv('WB200_GENERIC_READS', '1')
    old = _reads_noise_only(**sky_only)
    assert np.array_equal(new, old)


@pytest.mark.parametrize("lam", [0.4, 7.0, 59.0, 99.0])
def test_count_window_sampler_with_drifting_and_jumping_means(lam):
    """k_counts_window re-uses one tabulated CDF window over neighbouring sub-samples of a
    bin (window draw + a Poisson draw of the remainder) and rebuilds it when the mean
    leaves [lam_t, lam_t + 0.25]: means that drift both ways by parts in 1e4 (a transit in
    progress) and means that jump by 50 % between sub-samples keep the exact pmf."""
    n_s = 256
    drift = 1.0 + 2e-4 * np.sin(np.arange(n_s) / 7.0)
    c = _poisson_draws(lam, n_s, 4096, key=(31, 32), dur_scale=drift, flat=False)
    chi2, dof = _chi2_poisson(c.ravel(), lam)            # pooled: all means within 2e-4 of lam
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
    m = c.mean(axis=1)
    assert np.abs(m - lam * drift).max() < 5.5 * np.sqrt(lam / 4096)
    jump = np.where(np.arange(n_s) % 2 == 0, 1.0, 1.5)
    c = _poisson_draws(lam, n_s, 4096, key=(33, 34), dur_scale=jump, flat=False)
    for par, mult in ((0, 1.0), (1, 1.5)):
        chi2, dof = _chi2_poisson(c[par::2].ravel(), lam * mult)
        assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (par, chi2, dof)
    # neighbouring sub-samples of one bin share a Philox call (different words): independent
    a, b = c[0::2].ravel().astype(float), c[1::2].ravel().astype(float)
    assert abs(np.corrcoef(a, b)[0, 1]) < 5 / np.sqrt(a.size)


def test_count_window_sampler_tails():
    """Draws that leave the tabulated window (below lam - 3.7 sigma, above its 64th term)
    are finished by the recurrence: 64 M draws at lam = 59 reproduce the far-tail masses."""
    lam = 59.0
    lo_k, hi_k = 32, 90                      # P(X <= 32) = 8.8e-5, P(X >= 90) = 1.05e-4; the window is [29, 93)
    n_lo = n_hi = n = 0
    for key in range(16):
        x = _poisson_draws(lam, 1024, 4096, key=(41, key))
        n += x.size
        n_lo += int((x <= lo_k).sum())
        n_hi += int((x >= hi_k).sum())
    for got, p in ((n_lo, stats.poisson.cdf(lo_k, lam)), (n_hi, stats.poisson.sf(hi_k - 1, lam))):
        assert abs(got - n * p) < 5 * np.sqrt(n * p) + 0.01 * n * p, (got, n * p)


def test_photon_2d_cell_probabilities_narrow_psf():
    """The thrower's electrons (four per Philox call: 16-bit radius and angle fields, far tail
    refined from a second call) against the exact cell probabilities of a narrow double
    Gaussian -- sigma = 0.55 / 1.9 px, the regime of the real PSF -- at 5e7 electrons:
    chi-square over every cell with >= 200 expected, and the x / y marginals."""
    from wayne_b200 import pyparallel
    n = 50_000_000
    sig_l, sig_h, ratio = 0.55, 1.9, 0.3
    x0, y0 = 100.37, 90.81
    frame = pyparallel.psf_frame([n], [x0], [y0], [ratio], [sig_l], [sig_h], 200, 200, test=99,
                                 rng='philox').astype(float)
    assert frame.sum() == n
    k = np.arange(201)
    nh = int(n * ratio)

    def cells(sig):
        px = np.diff(stats.norm.cdf(k, x0, sig))
        py = np.diff(stats.norm.cdf(k, y0, sig))
        return np.outer(py, px)                      # frame[y, x]
    exp = nh * cells(sig_h) + (n - nh) * cells(sig_l)
    sel = exp >= 200
    chi2 = ((frame[sel] - exp[sel]) ** 2 / exp[sel]).sum()
    dof = sel.sum() - 1
    assert dof > 150
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
    for axis in (0, 1):
        o, e = frame.sum(axis=axis), exp.sum(axis=axis)
        s1 = e >= 200
        c2 = ((o[s1] - e[s1]) ** 2 / e[s1]).sum()
        assert c2 < s1.sum() + 6 * np.sqrt(2 * s1.sum()), (axis, c2)


def test_photon_far_tail_and_isotropy():
    """Radius fields below 16 (beyond 4.08 sigma) take their radius uniform from a second
    Philox call: the marginal tail masses beyond 3.5 / 4.5 / 5.2 sigma match the normal
    law (6.4e7 electrons), and the octants around the centre hold equal counts."""
    from wayne_b200 import pyparallel
    n, sig = 64_000_000, 20.0
    c0 = 507.0                                       # on a pixel corner: the octants are mirror images
    frame = pyparallel.psf_frame([n], [c0], [c0], [0.0], [sig], [sig], 1014, 1014, test=7,
                                 rng='philox').astype(float)
    assert frame.sum() == n
    for axis in (0, 1):
        prof = frame.sum(axis=axis)
        for z in (3.5, 4.5, 5.2):
            lo, hi = int(np.floor(c0 - z * sig)), int(np.ceil(c0 + z * sig))
            got = prof[:lo].sum() + prof[hi:].sum()
            p = stats.norm.cdf(lo, c0, sig) + stats.norm.sf(hi, c0, sig)
            assert abs(got - n * p) < 5 * np.sqrt(n * p), (axis, z, got, n * p)
    yy, xx = np.mgrid[0:1014, 0:1014]
    dx, dy = xx + 0.5 - c0, yy + 0.5 - c0            # pixel centres
    octant = (dx > 0).astype(int) + 2 * (dy > 0) + 4 * (np.abs(dx) > np.abs(dy))
    diag = np.abs(dx) == np.abs(dy)
    tot = np.array([frame[(octant == o) & ~diag].sum() for o in range(8)])
    assert np.abs(tot - tot.mean()).max() < 5 * np.sqrt(tot.mean()), tot
