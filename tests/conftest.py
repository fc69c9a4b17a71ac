import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def calb_dir(tmp_path_factory):
    """Synthetic, seeded calibration set (the real one cannot be downloaded here)."""
    from wayne_b200 import calibration, params
    d = str(tmp_path_factory.mktemp("calb"))
    calibration.write_synthetic_calibration(d, modes=((256, 'SPARS10'),))
    params.set_calibration_dir(d)
    return d
