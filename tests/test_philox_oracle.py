"""Known-answer pins of the generator under the native mode's random streams.

CPU: oracle/philox_oracle.py against the Random123 known-answer vectors of Philox4x32-10
(kat_vectors of the Random123 distribution) and splitmix64's first output.
GPU: the device functions (through the wb200_philox_words test hook) against the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import philox_oracle as P

KATS = [
    ([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_oracle_philox_known_answers():
    for c, k, want in KATS:
        assert P.philox4x32_10(c, k) == want


def test_oracle_splitmix_known_answer():
    # splitmix64 seeded with 0: first output 0xE220A8397B1DCDAF
    hy, hw = P.throw_keys(0, 0)
    assert (hw << 32 | hy) == 0xE220A8397B1DCDAF
    assert P.throw_keys(1963, 7) != P.throw_keys(1963, 8)


def test_oracle_thrower_fields():
    k, t, u1, th = P.thrower_fields(0x0000ffff)
    assert (k, t) == (0, 65535) and u1 == 0.5 / 65536 and abs(th - (np.pi - np.pi / 65536)) < 1e-12
    k, t, u1, th = P.thrower_fields(0xffff0000)
    assert (k, t) == (65535, 0) and u1 == 1 - 0.5 / 65536 and abs(th + np.pi - np.pi / 65536) < 1e-12


def _device_words(which, c, k):
    from wayne_b200 import _lib
    cc = (C.c_uint32 * 4)(*[int(x) & 0xffffffff for x in c])
    kk = (C.c_uint32 * 2)(*[int(x) & 0xffffffff for x in k])
    out = (C.c_uint32 * 6)()
    _lib.check(_lib.lib.wb200_philox_words(which, cc, kk, out), 'wb200_philox_words')
    return list(out)


@pytest.mark.gpu
def test_device_philox_known_answers():
    for c, k, want in KATS:
        assert _device_words(0, c, k)[:4] == want
    rng = np.random.default_rng(5)
    for _ in range(50):
        c = rng.integers(0, 2 ** 32, 4)
        k = rng.integers(0, 2 ** 32, 2)
        assert _device_words(0, c, k)[:4] == P.philox4x32_10(c, k)


@pytest.mark.gpu
def test_device_thrower_stream_equals_oracle():
    rng = np.random.default_rng(6)
    for _ in range(50):
        unit, sample, b = int(rng.integers(0, 1 << 20)), int(rng.integers(0, 5000)), int(rng.integers(0, 8192))
        key = (int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 2 ** 32)))
        for stream in (P.STREAM_PHOTONS, P.STREAM_PHOTON_TAIL):
            got = _device_words(1, [unit, sample, b, stream], key)
            assert got[:4] == P.thrower_words(unit, sample, b, key, stream)
            assert tuple(got[4:]) == P.throw_keys(*key)


@pytest.mark.gpu
@pytest.mark.parametrize("lam", [0.4, 7.0, 59.0])
def test_count_sampler_is_the_poisson_quantile_of_its_uniforms(lam):
    """k_counts_window is inversion of ONE uniform per term: with the Philox words restated by
    the oracle, every drawn count equals  Q(lam_t; u) + Q(lam - lam_t; u2)  (Q = the exact
    Poisson quantile function, lam_t = the tabulated mean 4.5 % below lam), except where a
    uniform falls within float32 rounding of a CDF step."""
    import torch
    from scipy import stats
    from wayne_b200 import _lib
    n_s, n_b, key = 64, 1024, (77, 99)
    dev = torch.device('cuda', 0)
    f = torch.full((n_b,), float(lam), dtype=torch.float64, device=dev)
    one = torch.ones((n_b,), dtype=torch.float64, device=dev)
    dwl = torch.full((n_b,), 1e-4, dtype=torch.float64, device=dev)
    dur = torch.full((n_s,), 1000.0, dtype=torch.float64, device=dev)
    counts = torch.empty((n_s, n_b), dtype=torch.int32, device=dev)
    totals = torch.empty((n_s,), dtype=torch.int64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())          # noqa: E731
    _lib.check(_lib.lib.wb200_counts(n_s, n_b, p(f), None, 0, p(one), p(dwl), p(dur), 1.0,
                                     _lib.COUNT_POISSON, key[0], key[1], None, p(counts), p(totals),
                                     None), 'wb200_counts')
    got = counts.cpu().numpy()
    # the kernel's arithmetic, restated: mean in float64 (same multiplication order), then float32
    e = np.float64(lam) * 1.0 * 1e-4 * 1e4 * 1000.0 * 1e-3 * 1.0
    lam32 = np.float32(e)
    lam_t = np.float32(lam32 * np.float32(np.float32(1.0) - np.float32(0.045)))
    rest = np.float32(lam32 - lam_t)
    assert 0 < rest <= 0.09 * lam_t + 0.25
    want = np.empty_like(got)
    for s in range(n_s):
        for w in range(n_b):
            words = P.philox4x32_10([0x80000000, w, s >> 1, 1], key)
            wu, wu2 = (words[2], words[3]) if (s & 1) else (words[0], words[1])
            u = [np.float32(np.float64(np.float32(x)) * 2.0 ** -32 + 2.0 ** -33) for x in (wu, wu2)]
            u = [min(float(x), 0.99999994) for x in u]
            want[s, w] = stats.poisson.ppf(u[0], float(lam_t)) + stats.poisson.ppf(u[1], float(rest))
    diff = got.astype(np.int64) - want
    assert (diff != 0).mean() < 1e-3, (diff != 0).mean()
    assert np.abs(diff).max() <= 1


@pytest.mark.gpu
def test_native_sky_draws_are_the_poisson_quantile_of_their_uniforms():
    """k_reads_native's sky term, pixel by pixel and read by read: the count equals the exact
    Poisson quantile of the uniform the oracle derives from the Philox call
    (1, pixel pair, read, WB_STREAM_SKY) -- word x for the even pixel, y for the odd one."""
    from scipy import stats
    from tests.test_rng_gpu import _reads_noise_only
    F, R, key, rate = 138, 3, (3, 4), 14.7
    out = _reads_noise_only(sky_rate=rate, F=F, R=R, key=key, fast=1, dts=(1.0, 1.0, 1.0))
    lam = float(np.float32(np.float32(1.0) * np.float32(rate * 1.0)))
    B = 5
    bad = n = 0
    for r in range(R):
        got = (out[r + 1] - out[r])[B:F - B, B:F - B]
        assert np.all(got == np.rint(got))
        for Y in range(B, F - B):
            for X0 in range(B - (B & 1), F - B, 2):
                words = P.philox4x32_10([1, Y * F + X0, r, 4], key)
                for h in (0, 1):
                    X = X0 + h
                    if X < B or X >= F - B:
                        continue
                    u = float(np.float32((np.float64(words[h]) + 0.5) * 2.3283064365386963e-10))
                    want = stats.poisson.ppf(min(u, 0.99999994), lam)
                    n += 1
                    bad += int(got[Y - B, X - B] != want)
    assert n == R * (F - 2 * B) ** 2
    assert bad < 1e-3 * n, (bad, n)


@pytest.mark.gpu
def test_thrower_electrons_land_where_the_oracle_puts_them():
    """The native thrower against an independent restatement of its whole recipe: Philox words
    (oracle), 16-bit radius / angle fields, tail refinement from the second stream for radius
    fields < 16, wide electrons first in units of four, Box-Muller in float64 with exact
    functions.  Every electron must land in the oracle's pixel, except the few that fall
    within float32 / SFU rounding (~1e-5 px) of a pixel edge."""
    import math
    from wayne_b200 import pyparallel
    rng = np.random.default_rng(11)
    nb, NR = 48, 128
    counts = rng.integers(0, 4000, nb).astype(np.int32)
    counts[5] = 0
    counts[6] = 1
    counts[7] = 5
    x = 40.0 + rng.random(nb) * 48.0
    y = 40.0 + rng.random(nb) * 48.0
    ratio = rng.random(nb) * 0.6
    sigl = 0.4 + rng.random(nb) * 0.5
    sigh = 1.2 + rng.random(nb) * 1.0
    test = 4321
    got = pyparallel.psf_frame(counts, x, y, ratio, sigl, sigh, NR, NR, test=test, rng='philox')
    assert got.sum() == counts.sum()                       # 8.2 sigma_h stays inside the frame
    key = (test, 0x57415945)                               # wb200_psf_host's key pair
    c32 = np.float32(2 * np.pi / 65536)
    d32 = np.float32(-float(c32) * (2 ** 23 + 32768) + float(c32) / 2)
    want = np.zeros((NR, NR), dtype=np.int64)
    for w in range(nb):
        cnt = int(counts[w])
        nh = max(0, min(int(float(cnt) * ratio[w]), cnt))  # (int)(counts*ratio), pyparallel_menu.c:89
        nl = cnt - nh
        uh = (nh + 3) // 4
        for j in range(uh + (nl + 3) // 4):
            wide = j < uh
            rem = (nh - 4 * j) if wide else (nl - 4 * (j - uh))
            sigma = float(np.float32(sigh[w] if wide else sigl[w]))
            words = P.thrower_words(4 * j, 0, w, key)            # first counter word = the unit's first electron
            tail = None
            for h in range(min(4, rem)):
                k, t = words[h] >> 16, words[h] & 0xffff
                if k < 16:
                    if tail is None:
                        tail = P.thrower_words(4 * j, 0, w, key, P.STREAM_PHOTON_TAIL)
                    u1 = (k + (tail[h] + 0.5) / 2.0 ** 32) / 65536.0
                else:
                    u1 = (k + 0.5) / 65536.0
                theta = float(np.float32(float(c32) * (2 ** 23 + t) + float(d32)))   # the device's fused form
                r = math.sqrt(-2.0 * math.log(u1)) * sigma
                px, py = x[w] + r * math.cos(theta), y[w] + r * math.sin(theta)
                want[int(math.floor(py)), int(math.floor(px))] += 1
    moved = np.abs(got.astype(np.int64) - want).sum() / 2
    assert moved <= 5e-5 * counts.sum() + 3, (moved, counts.sum())


def test_thrower_streams_of_distinct_exposures_do_not_overlap():
    """The thrower's Philox key is fixed; an exposure enters through the counter: word 1 =
    hy + sub-sample, word 3 = hw ^ stream, (hy, hw) = splitmix64(exposure key).  Two exposures
    share random blocks only if their hw agree up to the stream ids (2 photons, 7 tail) AND their
    hy lie within the sub-sample range of each other (photons.cuh argues ~2^-49 per pair).
    Checked exhaustively for the key sets the drivers use: the 8 x 128 exposures of BASELINE
    configs[4] as bench.py keys them, and the file-name keys of a 1000-exposure visit."""
    import zlib
    import numpy as np
    from oracle import philox_oracle as P
    keys = [(1963 + v, v * 128 + k) for v in range(8) for k in range(128)]                    # bench.py multi_visit
    keys += [(1963, zlib.crc32(('%04d_raw.fits' % n).encode()) & 0xffffffff) for n in range(1, 1001)]  # Observation
    keys += [(s, i) for s in (0, 1, 2 ** 32 - 1) for i in range(64)]
    hy, hw = (np.array(v, dtype=np.uint64) for v in zip(*(P.throw_keys(*k) for k in keys)))
    assert len(set(zip(hy.tolist(), hw.tolist()))) == len(keys)
    x = hw[:, None] ^ hw[None, :]
    same_stream_word = (x == 0) | (x == (P.STREAM_PHOTONS ^ P.STREAM_PHOTON_TAIL))
    d = (hy[:, None].astype(np.int64) - hy[None, :].astype(np.int64)) % (1 << 32)
    near = (d < 65536) | (d > (1 << 32) - 65536)                     # sub-samples per exposure <= 65535
    clash = same_stream_word & near
    np.fill_diagonal(clash, False)
    assert not clash.any()
    # the hashed words look uniform: no two exposures even share hw
    assert len(np.unique(hw)) == len(keys)
