"""``wayne.pyparallel`` drop-in: the electron thrower behind the same signature.

Reference: the Cython wrapper ``apply_psf`` (wayne/pyparallel.pyx:14-38) around
the OpenMP C function ``PSF`` (wayne/pyparallel_menu.c:10-113).  Here the same
call goes to the ``PSF`` symbol of libwayne_b200.so, which runs the photon
kernel on the current CUDA device and reproduces the reference's rand_r stream
for the given ``(test, threads)`` bit for bit.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib

_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def apply_psf(counts, pos_x, pos_y, ratio_psf, sigmal_psf, sigmah_psf, NR, NC, test, threads):
    """Throw ``counts[i]`` electrons per wavelength bin through the double-Gaussian
    PSF and histogram them on an NR x NC frame.  Returns float64 ndarray [NR*NC]
    (the caller reshapes, wayne/exposure_generator.py:636-639)."""
    ccounts = np.ascontiguousarray(np.asarray(counts), dtype=np.int32)  # C int truncation of the pyx loop
    n = len(ccounts)
    px, py, r, sl, sh = (_f64(a) for a in (pos_x, pos_y, ratio_psf, sigmal_psf, sigmah_psf))
    for a in (px, py, r, sl, sh):
        if len(a) != n:
            raise ValueError("all per-bin arrays must have len(counts) elements")
    ptr = lib.PSF(ccounts.ctypes.data_as(_lib.IP), n, px.ctypes.data_as(_lib.DP),
                  py.ctypes.data_as(_lib.DP), r.ctypes.data_as(_lib.DP),
                  sl.ctypes.data_as(_lib.DP), sh.ctypes.data_as(_lib.DP),
                  int(NR), int(NC), int(test), int(threads))
    if not ptr:
        raise _lib.WayneB200Error("PSF failed: " + lib.wb200_last_error().decode())
    try:
        out = np.ctypeslib.as_array(ptr, shape=(int(NR) * int(NC),)).astype(np.float64)
    finally:
        _libc.free(ptr)
    return out


def psf_frame(counts, pos_x, pos_y, ratio_psf, sigmal_psf, sigmah_psf, NR, NC, test=0, threads=1,
              rng='randr', normals=None):
    """int32 [NR][NC] histogram with an explicit RNG mode: 'randr' (reference
    stream), 'host' (caller-supplied normal table A[2*ssum]) or 'philox'."""
    mode = {'philox': _lib.RNG_PHILOX, 'randr': _lib.RNG_RANDR, 'host': _lib.RNG_HOST}[rng]
    ccounts = np.ascontiguousarray(np.asarray(counts), dtype=np.int32)
    n = len(ccounts)
    px, py, r, sl, sh = (_f64(a) for a in (pos_x, pos_y, ratio_psf, sigmal_psf, sigmah_psf))
    frame = np.empty(int(NR) * int(NC), dtype=np.int32)
    nrm = None
    if normals is not None:
        nrm = _f64(normals)
    _lib.check(lib.wb200_psf_host(
        ccounts.ctypes.data_as(_lib.IP), n, px.ctypes.data_as(_lib.DP), py.ctypes.data_as(_lib.DP),
        r.ctypes.data_as(_lib.DP), sl.ctypes.data_as(_lib.DP), sh.ctypes.data_as(_lib.DP),
        int(NR), int(NC), int(test), int(threads), mode,
        nrm.ctypes.data_as(_lib.DP) if nrm is not None else None,
        frame.ctypes.data_as(_lib.IP)), "wb200_psf_host")
    return frame.reshape(int(NR), int(NC))
