"""CPU checks of the literal constants of the native thrower's random recipe
(wayne_b200/csrc/photons.cuh): the 16-bit field decoding must give the lattice
DESIGN.md §4.1b describes.  The literals are read from the source, the float32
arithmetic of the device code is redone with numpy."""
import os
import re

import numpy as np

SRC = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        'wayne_b200', 'csrc', 'photons.cuh')).read()
f32 = np.float32


def _byte_perm(x, y, sel):
    """CUDA __byte_perm: result byte i = byte (sel >> 4i) & 7 of the 8 bytes {x, y}."""
    b = [(x >> (8 * i)) & 0xff for i in range(4)] + [(y >> (8 * i)) & 0xff for i in range(4)]
    return sum(b[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))


def _as_float(bits):
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def _fmaf(a, b, c):
    """fp32 fused multiply-add: exact product and sum (float64 holds them here), one rounding."""
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def _literals(func):
    body = SRC[SRC.index(func):]
    body = body[:body.index('}')]
    return body


def test_field_extraction_selectors():
    hi = int(re.search(r'hi16_biased.*?0x4B000000u, (0x[0-9a-fA-F]+)\)', SRC, re.S).group(1), 16)
    lo = int(re.search(r'lo16_biased.*?0x4B000000u, (0x[0-9a-fA-F]+)\)', SRC, re.S).group(1), 16)
    for w in (0x00000000, 0xffffffff, 0x12345678, 0x8000ffff, 0x0001fffe):
        assert _as_float(_byte_perm(w, 0x4B000000, hi)) == f32(2 ** 23 + (w >> 16))
        assert _as_float(_byte_perm(w, 0x4B000000, lo)) == f32(2 ** 23 + (w & 0xffff))


def test_radius_uniform_is_exact_midpoint_lattice():
    m = re.search(r'fmaf\(hi16_biased\(w\), ([0-9.e+-]+)f, (-[0-9.e+-]+)f\)', _literals('float throw_u1(uint32_t w)'))
    a, b = f32(float(m.group(1))), f32(float(m.group(2)))
    assert a == f32(2.0 ** -16)
    for k in (0, 1, 15, 16, 12345, 65535):
        u1 = _fmaf(f32(2 ** 23 + k), a, b)
        assert u1 == f32((k + 0.5) / 65536.0), k        # exactly representable
    assert _fmaf(f32(2 ** 23), a, b) > 0


def test_angle_lattice_covers_the_circle_once():
    m = re.search(r'fmaf\(lo16_biased\(w\), ([0-9.e+-]+)f, (-[0-9.e+-]+)f\)', _literals('float throw_theta(uint32_t w)'))
    c, d = f32(float(m.group(1))), f32(float(m.group(2)))
    assert c == f32(2 * np.pi / 65536)
    t = np.arange(65536, dtype=np.float64)
    th = (np.float64(c) * (2 ** 23 + t) + np.float64(d)).astype(np.float32).astype(np.float64)
    step = 2 * np.pi / 65536
    # uniform lattice (to float32 rounding of the result), total span 2 pi minus one step,
    # centred on 0 up to a constant rotation well below one lattice step ... plus margin
    assert np.abs(np.diff(th) - step).max() < 3e-7
    assert abs((th[-1] - th[0]) - (2 * np.pi - step)) < 1e-6
    assert abs(th[0] + th[-1]) < 2 * step
    assert th[0] > -np.pi - step and th[-1] < np.pi + step   # inside sin/cos.approx's accurate range


def test_sigma_prefactor_and_tail_threshold():
    k = float(re.search(r'WB_SQRT_2LN2 = ([0-9.]+)f', SRC).group(1))
    assert abs(k * k - 2 * np.log(2)) < 1e-6       # sqrt(-2 ln u) = sqrt(2 ln 2) sqrt(-lg2 u)
    tail = re.search(r'WB_TAIL_WORD = (\d+)u << (\d+)', SRC)
    assert (int(tail.group(1)), int(tail.group(2))) == (16, 16)
    # radius fields below 16: u1 < 16/65536, i.e. everything beyond sqrt(-2 ln(16/65536)) sigma
    assert abs(np.sqrt(-2 * np.log(16 / 65536.0)) - 4.08) < 0.01
    zmax = float(re.search(r'WB_ZMAX_THROW = ([0-9.]+)f', SRC).group(1))
    assert np.sqrt(2 * 49 * np.log(2)) < zmax < 8.5
    from wayne_b200 import engine
    assert engine.ZMAX[engine._lib.RNG_PHILOX] >= np.sqrt(2 * 49 * np.log(2))
