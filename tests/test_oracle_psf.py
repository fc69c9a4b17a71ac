"""Pins the C restatement of PSF() (oracle/psf_oracle.c):
  * wo_rand_r against the C library's rand_r,
  * wo_psf bit-for-bit against the UNMODIFIED reference kernel compiled from
    /root/reference (oracle/_ref, when it was built in this container),
  * wo_psf against the committed golden histograms (tests/golden/psf_golden.npz,
    made by tests/golden/make_psf_golden.py from the unmodified reference).
"""
import os

import numpy as np
import pytest

from oracle import psf as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "psf_golden.npz")


def test_rand_r_matches_libc():
    for seed in (0, 1, 25234, 25234 + 17 * 3 + 99999, 2 ** 32 - 1):
        assert O.port_rand_r(seed, 1000) == O.libc_rand_r(seed, 1000)


@pytest.mark.skipif(not O.have_reference(), reason="reference kernel not built here")
@pytest.mark.parametrize("test,threads", [(0, 1), (7, 1), (7, 2), (99999, 3), (123, 4), (5, 8)])
def test_port_equals_unmodified_reference(test, threads):
    case = O.psf_case(seed=test, n_bins=700, mean_count=35.0)
    a = O.psf_port(test=test, threads=threads, **case)
    b = O.psf_reference(test=test, threads=threads, **case)
    assert a.sum() > 0
    assert np.array_equal(a, b)


@pytest.mark.skipif(not O.have_reference(), reason="reference kernel not built here")
def test_port_equals_reference_edges():
    # trace hanging over the frame edges, zero-count bins, strict >0 bound
    case = O.psf_case(seed=3, n_bins=300, mean_count=20.0, frame=64, x0=-5.0, x1=70.0, y0=1.0)
    case["counts"][::7] = 0
    for threads in (1, 2, 5):
        a = O.psf_port(test=11, threads=threads, **case)
        b = O.psf_reference(test=11, threads=threads, **case)
        assert np.array_equal(a, b)
        assert a[0, :].sum() == 0 and a[:, 0].sum() == 0      # row/column 0 never filled


def test_port_equals_golden():
    g = np.load(GOLD)
    n = int(g["n_cases"])
    assert n >= 4
    for i in range(n):
        seed, test, threads, n_bins, frame = (int(v) for v in g["case%d_meta" % i])
        case = O.psf_case(seed=seed, n_bins=n_bins, mean_count=30.0, frame=frame)
        a = O.psf_port(test=test, threads=threads, **case)
        assert np.array_equal(np.flatnonzero(a), g["case%d_idx" % i])
        assert np.array_equal(a.ravel()[g["case%d_idx" % i]], g["case%d_val" % i])


def test_normal_table_layout():
    # A[i] x-normal, A[i+ssum] y-normal; threads chunk the electrons
    A1 = O.fill_normals(1000, 5, 1)
    A2 = O.fill_normals(1000, 5, 2)
    assert np.array_equal(A1[:500], A2[:500])          # thread 0 has the same seed/chunk start
    assert not np.array_equal(A1[500:1000], A2[500:1000])
    # thread t of test == thread t-1 of test+17 (seed collision)
    A3 = O.fill_normals(1000, 5 + 17, 2)
    assert np.array_equal(A2[500:1000], A3[:500])


def test_empty_and_ragged():
    case = O.psf_case(seed=1, n_bins=64)
    case["counts"][:] = 0
    assert O.psf_port(test=0, threads=2, **case).sum() == 0
    case = O.psf_case(seed=1, n_bins=1, mean_count=1000.0)
    f = O.psf_port(test=0, threads=3, **case)
    assert 0 < f.sum() <= case["counts"].sum()
