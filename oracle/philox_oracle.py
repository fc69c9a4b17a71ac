"""CPU restatement of the counter-based random functions of the native mode -- TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py): the product never imports this.

* ``philox4x32_10`` -- Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw,
  "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 constants).
  Pinned to the Random123 known-answer vectors in tests/test_philox_oracle.py.
  (The reference itself has no counter-based generator: its streams are numpy's
  global Mersenne Twister and glibc rand_r, restated in exposure_oracle.py /
  psf_oracle.c for the compat mode.  The native mode's claim is the DISTRIBUTION,
  and this oracle pins the generator underneath it.)
* ``throw_keys`` / ``thrower_words`` -- the native thrower's use of it
  (wayne_b200/csrc/photons.cuh): splitmix64 hash of the exposure key, fixed Philox key,
  counter (e0, hy + sub-sample, bin, hw ^ stream) with e0 = 4 * unit = the index of the unit's
  first electron in the bin's list (wide electrons first, each width padded to a multiple of 4).
* ``thrower_fields`` -- the 16-bit radius / angle fields of a word and the uniforms they
  stand for (exact rationals; the device evaluates them in fp32 and then uses the SFU).
"""
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xffffffff
THROW_KEY = (0xA4093822, 0x299F31D0)
STREAM_PHOTONS, STREAM_PHOTON_TAIL = 2, 7


def philox4x32_10(counter, key, rounds=10):
    c = [int(x) & MASK for x in counter]
    k0, k1 = (int(x) & MASK for x in key)
    for _ in range(rounds):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def throw_keys(key0, key1):
    """splitmix64 of (key1 << 32 | key0): (hy, hw) = (low, high) words."""
    m64 = (1 << 64) - 1
    z = (((int(key1) & MASK) << 32 | (int(key0) & MASK)) + 0x9E3779B97F4A7C15) & m64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m64
    z ^= z >> 31
    return z & MASK, z >> 32


def thrower_words(e0, sample, bin_index, key, stream=STREAM_PHOTONS):
    """The four words of one unit (= four electrons); e0 = 4 * unit index within the bin."""
    hy, hw = throw_keys(*key)
    return philox4x32_10([e0, (hy + sample) & MASK, bin_index, hw ^ stream], THROW_KEY)


def thrower_fields(word):
    """(k, t, u1, theta): radius field, angle field, the radius uniform (k + 1/2) 2^-16 and the
    angle (t - 32768 + 1/2) 2 pi 2^-16 they stand for (k < 16 is refined from the tail stream)."""
    import math
    k, t = word >> 16, word & 0xffff
    return k, t, (k + 0.5) / 65536.0, (t - 32768 + 0.5) * 2 * math.pi / 65536.0
