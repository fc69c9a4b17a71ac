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
