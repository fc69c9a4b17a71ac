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


def test_struct_layouts_match_header(tmp_path):
    """Every ctypes Structure against the C compiler's own sizeof / offsetof of the header's
    struct (the header compiled as plain C: the boundary is a C ABI)."""
    import subprocess
    from wayne_b200 import _lib
    pairs = [("wb200_photon_args", _lib.PhotonArgs), ("wb200_counts_args", _lib.CountsArgs),
             ("wb200_gather_args", _lib.GatherArgs), ("wb200_reads_args", _lib.ReadsArgs),
             ("wb200_instrument", _lib.Instrument), ("wb200_exposure_args", _lib.ExposureArgs)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "wayne_b200.h"', 'int main(void) {']
    for cname, st in pairs:
        lines.append('printf("%s.sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in st._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = str(tmp_path / 'layout')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', exe],
                   check=True)
    got = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True)
               .stdout.strip().splitlines())
    for cname, st in pairs:
        assert int(got[cname + '.sizeof']) == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got['%s.%s' % (cname, fname)]) == getattr(st, fname).offset, (cname, fname)


def test_context_calls_report_errors_without_a_gpu():
    """The exposure-level entry points: integer status + message, nothing thrown, NULL-safe."""
    from wayne_b200 import _lib
    lib = _lib.lib
    assert lib.wb200_ctx_destroy(None) == 0
    assert lib.wb200_exposure_run(None, None, None, None) == -1
    assert b"null context" in lib.wb200_last_error()
    assert lib.wb200_ctx_set_instrument(None, None) == -1
    assert lib.wb200_ctx_upload_plane(None, 0, None, 0, 0) == -1
    h = ctypes.c_void_p()
    rc = lib.wb200_ctx_create(0, ctypes.byref(h))
    import torch
    if not torch.cuda.is_available():
        assert rc < 0 and not h.value and lib.wb200_last_error()
    else:
        assert rc == 0 and h.value
        assert lib.wb200_exposure_run(h, None, None, None) == -1
        assert b"set_instrument" in lib.wb200_ctx_last_error(h)
        assert lib.wb200_ctx_destroy(h) == 0


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "wayne_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
