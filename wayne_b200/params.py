"""Locations of bundled data and of the instrument calibration set.

The reference computes these at import time and downloads the calibration zip
from a UCL server when it is missing (wayne/params.py:26-56).  There is no
network here and no download is attempted: the calibration directory is
``$WAYNE_CALB_DIR`` (or ``~/.wayne/calibration``); when a file is missing the
grism / detector classes raise an error naming it.  A synthetic, seeded
stand-in set can be written with :func:`wayne_b200.calibration.write_synthetic_calibration`.
"""
import os

_ROOT = os.path.abspath(os.path.dirname(__file__))
_data_dir = os.path.join(_ROOT, "data")
_calb_dir = os.environ.get("WAYNE_CALB_DIR",
                           os.path.join(os.path.expanduser("~"), ".wayne", "calibration"))

seed = None  # set by the visit driver (wayne/params.py:60, run_visit.py:69-77)

# default random-stream mode of ExposureGenerator: 'philox' (native, counter
# based) or 'numpy' (compat: the reference's numpy + rand_r streams)
rng = os.environ.get("WAYNE_B200_RNG", "philox")

# native mode: fuse flat field + accumulation into the photon kernel's tile flush
# (int64 fixed-point interval planes).  Off = per-sub-sample windows + ordered
# gather, the path the parity mode always uses.
direct_accumulation = os.environ.get("WAYNE_B200_DIRECT", "1") != "0"


# native mode: one wb200_exposure_run call per exposure on a resident context
# (include/wayne_b200.h, "Exposure-level interface").  Off = the stage-by-stage calls driven
# from Python (engine.ExposureRun), kept as the A/B reference and used by the parity mode.
use_context = os.environ.get("WAYNE_B200_CONTEXT", "1") != "0"


def set_calibration_dir(path):
    """Point the package at a calibration directory (affects objects built afterwards)."""
    global _calb_dir
    _calb_dir = os.path.abspath(path)
    return _calb_dir


class CalibrationFileMissing(IOError):
    pass


def calb_path(name):
    path = os.path.join(_calb_dir, name)
    if not os.path.isfile(path):
        raise CalibrationFileMissing(
            "calibration file '{}' not found in '{}' (set WAYNE_CALB_DIR, or write a synthetic "
            "set with wayne_b200.calibration.write_synthetic_calibration)".format(name, _calb_dir))
    return path
