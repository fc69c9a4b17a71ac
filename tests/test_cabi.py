"""The C-ABI library loads on a machine without a GPU and exports every symbol
include/wayne_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "wayne_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(wb200_[a-z_0-9]+|PSF)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from wayne_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert "PSF" in names and "wb200_throw_photons" in names and len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_argument_errors_are_reported_not_thrown():
    from wayne_b200 import _lib
    lib = _lib.lib
    assert lib.wb200_version() >= 100
    rc = lib.wb200_throw_photons(None, None)
    assert rc == -1 and b"null args" in lib.wb200_last_error()
    rc = lib.wb200_reads(None, None)
    assert rc == -1


def test_struct_layouts_match_header():
    # sizes computed by hand from include/wayne_b200.h (LP64)
    from wayne_b200 import _lib
    assert ctypes.sizeof(_lib.PhotonArgs) == 9 * 4 + 4 + 8 + 8 + 17 * 8
    assert ctypes.sizeof(_lib.GatherArgs) == 14 * 4 + 2 * 8 + 5 * 8 + 4 * 8 + 8
    assert ctypes.sizeof(_lib.ReadsArgs) == 12 * 4 + 8 + 3 * 8 + 8 + 8 + 4 * 8 + 7 * 8 + 7 * 8 + 4 * 8 + 4 * 8 + 2 * 8


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "wayne_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
