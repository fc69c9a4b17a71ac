"""Generates tests/golden/psf_golden.npz from the UNMODIFIED reference kernel
(oracle/_ref/libwayne_ref_psf.so, compiled by oracle/Makefile from
/root/reference/wayne/pyparallel_menu.c).  Run here, where the reference exists:

    python tests/golden/make_psf_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import psf as O  # noqa: E402

CASES = [  # seed, test, threads, n_bins, frame
    (0, 0, 1, 400, 256),
    (1, 4242, 2, 400, 256),
    (2, 99999, 4, 256, 128),
    (3, 17, 8, 512, 256),
    (4, 31337, 3, 64, 64),
]
out = {"n_cases": np.int64(len(CASES))}
for i, (seed, test, threads, n_bins, frame) in enumerate(CASES):
    case = O.psf_case(seed=seed, n_bins=n_bins, mean_count=30.0, frame=frame)
    f = O.psf_reference(test=test, threads=threads, **case).ravel()
    idx = np.flatnonzero(f)
    out["case%d_meta" % i] = np.array([seed, test, threads, n_bins, frame], dtype=np.int64)
    out["case%d_idx" % i] = idx.astype(np.int32)
    out["case%d_val" % i] = f[idx].astype(np.int32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "psf_golden.npz"), **out)
print("wrote", len(CASES), "cases")
