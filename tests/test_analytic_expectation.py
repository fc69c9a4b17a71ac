"""The analytic per-pixel expectation (tests/analytic.py) that the bench-shape GPU
test leans on is itself pinned here, on the CPU, to the reference's electron
thrower (the C oracle == the unmodified pyparallel_menu.c): many seeds of PSF()
on one sub-sample's inputs average to the expectation, pixel by pixel."""
import numpy as np

from oracle import psf as O
from tests import analytic


def test_expectation_matches_the_reference_thrower():
    case = O.psf_case(seed=3, n_bins=384, mean_count=400.0, frame=256, x0=30.0, x1=210.0, y0=118.0)
    tot = np.zeros((256, 256))
    n_seeds = 24
    for seed in range(n_seeds):
        tot += O.psf_port(test=1000 + 37 * seed, threads=1 + seed % 3, **case)
    exp, var = analytic.expected_interval_images(
        case['counts'][None, :], case['x'][None, :], case['y'][None, :], case['ratio'], case['sigl'],
        case['sigh'], [0], 256, device='cpu', with_var=True)
    exp, var = exp[0] * n_seeds, var[0] * n_seeds
    # the reference tests x against nr and y against nc with strict > 0: nothing lands in row / column 0
    assert tot[0].sum() == 0 and tot[:, 0].sum() == 0 and exp[0].sum() == 0 and exp[:, 0].sum() == 0
    assert abs(tot.sum() - exp.sum()) < 5 * np.sqrt(exp.sum())
    m = exp > 30
    z = (tot[m] - exp[m]) / np.sqrt(exp[m])
    assert m.sum() > 3000
    assert abs(z.mean()) < 5 / np.sqrt(m.sum())
    assert 0.93 < z.std() < 1.05        # multinomial: slightly below 1 ...
    zv = (tot[m] - exp[m]) / np.sqrt(var[m])
    assert 0.97 < zv.std() < 1.03       # ... and 1 against the exact multinomial variance
    assert np.all(var <= exp + 1e-9) and var[m].min() > 0
    assert np.abs(z).max() < 6.0
    # the far halo (expectation below one electron per pixel) holds what it should in total
    halo = exp < 1.0
    assert abs(tot[halo].sum() - exp[halo].sum()) < 6 * np.sqrt(exp[halo].sum() + 1)


def test_flat_value_is_the_oracles():
    """analytic.flat_value == what oracle.flat_field_at_hits multiplies a hit pixel by."""
    from oracle import exposure_oracle as E
    rng = np.random.default_rng(5)
    cal = {'flat': tuple((1 + 0.01 * rng.standard_normal((1014, 1014))).astype(np.float32) if i == 0 else
                         (0.005 * rng.standard_normal((1014, 1014))).astype(np.float32) for i in range(4)),
           'flat_wmin': 10000.0, 'flat_wmax': 17000.0}
    for sub in (256, 1024):
        L = 1014 if sub == 1024 else sub
        frame = np.zeros((L, L))
        rr, cc = rng.integers(0, L, 500), rng.integers(0, L, 500)
        frame[rr, cc] = 1.0
        out = E.flat_field_at_hits(404.3, 457.9, sub, frame, cal, E.G141_TRACE, E.G141_WLSOL)
        got = analytic.flat_value(cal, 'G141', sub, 404.3, 457.9, rr, cc)
        assert np.array_equal(out[rr, cc], got)


def test_flat_weighted_expectation_uses_that_flat_value():
    """expected_interval_images(cal=...) == sum over sub-samples of (image without flat) x flat_value."""
    rng = np.random.default_rng(8)
    cal = {'flat': tuple((1 + 0.01 * rng.standard_normal((1014, 1014))).astype(np.float32) if i == 0 else
                         (0.005 * rng.standard_normal((1014, 1014))).astype(np.float32) for i in range(4)),
           'flat_wmin': 10000.0, 'flat_wmax': 17000.0}
    case = O.psf_case(seed=4, n_bins=16, mean_count=50.0, frame=1014, x0=3.0, x1=30.0, y0=6.0)
    counts = np.stack([case['counts'], case['counts'][::-1]])
    xs = np.stack([case['x'], case['x'] + 0.3])
    ys = np.stack([case['y'], case['y'] + 2.7])
    refs = np.array([[330.1, 110.2], [330.2, 112.9]])
    args = (case['ratio'], case['sigl'], case['sigh'], [1], 1014)
    got = analytic.expected_interval_images(counts, xs, ys, *args, cal=cal, subarray=1024, refs=refs,
                                            batch=2, device='cpu')[0]
    want = np.zeros((1014, 1014))
    rows, cols = np.mgrid[0:1014, 0:1014]
    for s in range(2):
        img = analytic.expected_interval_images(counts[s:s + 1], xs[s:s + 1], ys[s:s + 1], case['ratio'],
                                                case['sigl'], case['sigh'], [0], 1014, device='cpu')[0]
        want += img * analytic.flat_value(cal, 'G141', 1024, refs[s, 0], refs[s, 1], rows, cols)
    assert want.sum() > 1000 and np.max(np.abs(got - want)) < 1e-12 * want.max()
