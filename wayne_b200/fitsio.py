"""A small FITS reader / writer (numpy only).

The reference does all of its file I/O on the path through astropy.io.fits:
calibration planes (wayne/grism.py:66-106, wayne/detector.py:31-67, 172-191,
200-209), the initial bias (wayne/exposure_generator.py:456-458) and the
output (wayne/exposure.py:133-214).  astropy is not a dependency of this
package, so the subset of FITS that those files use is implemented here:
primary + IMAGE extensions (BITPIX 8/16/32/64/-32/-64, BSCALE/BZERO ignored
unless trivial) and fixed-width BINTABLE columns (L, B, I, J, K, E, D, A).

``open`` returns a list-like of HDUs with ``.header`` (dict-like, ordered) and
``.data`` (numpy array, big-endian on disk exactly as astropy presents it, so
float32 planes stay float32 and downstream dtype promotion matches the
reference), and ``hdu.field(name)`` for tables.
"""
from __future__ import annotations

import builtins
import collections
import os

import numpy as np

BLOCK = 2880
_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_DTYPE_BITPIX = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}
_TFORM = {"L": "i1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8"}


class Header(collections.OrderedDict):
    """Ordered keyword -> value map; comments kept on the side."""

    def __init__(self, *a, **k):
        super(Header, self).__init__(*a, **k)
        self.comments = {}

    def set(self, key, value, comment=None):
        self[key] = value
        if comment is not None:
            self.comments[key] = comment

    def __setitem__(self, key, value):
        if isinstance(value, tuple) and len(value) == 2:
            self.comments[key] = value[1]
            value = value[0]
        super(Header, self).__setitem__(key, value)


class HDU(object):
    def __init__(self, data=None, header=None, name=None):
        self.data = data
        self.header = header if header is not None else Header()
        if name is not None:
            self.header["EXTNAME"] = name
        self.columns = None

    @property
    def name(self):
        return self.header.get("EXTNAME", "PRIMARY")

    def field(self, name):
        return self.data[name]


class HDUList(list):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


def _parse_value(raw):
    raw = raw.strip()
    if not raw:
        return None
    if raw.startswith("'"):
        end = 1
        out = []
        while end < len(raw):
            if raw[end] == "'":
                if end + 1 < len(raw) and raw[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(raw[end])
            end += 1
        return "".join(out).rstrip()
    raw = raw.split("/")[0].strip()
    if raw in ("T", "F"):
        return raw == "T"
    try:
        return int(raw)
    except ValueError:
        pass
    try:
        return float(raw.replace("D", "E"))
    except ValueError:
        return raw


def _read_header(buf, pos):
    hdr = Header()
    while True:
        block = buf[pos:pos + BLOCK]
        if len(block) < BLOCK:
            raise IOError("truncated FITS header")
        pos += BLOCK
        done = False
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if not key or key in ("COMMENT", "HISTORY") or card[8:10] != "= ":
                continue
            hdr[key] = _parse_value(card[10:])
        if done:
            return hdr, pos


def _padded(n):
    return (n + BLOCK - 1) // BLOCK * BLOCK


def open(path):  # noqa: A001 - mirrors astropy.io.fits.open
    """Read every HDU of ``path``."""
    with builtins.open(path, "rb") as f:
        buf = f.read()
    hdus = HDUList()
    pos = 0
    while pos < len(buf):
        hdr, pos = _read_header(buf, pos)
        naxis = int(hdr.get("NAXIS", 0))
        shape = [int(hdr["NAXIS%d" % (i + 1)]) for i in range(naxis)]
        bitpix = int(hdr.get("BITPIX", 8))
        pcount = int(hdr.get("PCOUNT", 0))
        gcount = int(hdr.get("GCOUNT", 1))
        nelem = int(np.prod(shape)) if naxis else 0
        nbytes = abs(bitpix) // 8 * gcount * (pcount + nelem)
        raw = buf[pos:pos + nbytes]
        pos += _padded(nbytes)
        hdu = HDU(None, hdr)
        xt = hdr.get("XTENSION", "IMAGE" if naxis else None)
        if nelem and xt in ("IMAGE", None):
            arr = np.frombuffer(raw, dtype=_BITPIX_DTYPE[bitpix], count=nelem)
            arr = arr.reshape(shape[::-1])
            bscale, bzero = hdr.get("BSCALE", 1), hdr.get("BZERO", 0)
            if bscale != 1 or bzero != 0:
                arr = arr * bscale + bzero
            hdu.data = arr
        elif nelem and xt == "BINTABLE":
            nfields = int(hdr["TFIELDS"])
            names, formats = [], []
            for i in range(1, nfields + 1):
                tform = str(hdr["TFORM%d" % i]).strip()
                rep = ""
                while tform and tform[0].isdigit():
                    rep += tform[0]
                    tform = tform[1:]
                code = tform[0]
                rep = int(rep) if rep else 1
                names.append(str(hdr.get("TTYPE%d" % i, "col%d" % i)).strip())
                if code == "A":
                    formats.append("S%d" % rep)
                elif rep == 1:
                    formats.append(_TFORM[code])
                else:
                    formats.append((_TFORM[code], (rep,)))
            dt = np.dtype({"names": names, "formats": formats})
            hdu.data = np.frombuffer(raw, dtype=dt, count=shape[1])
            hdu.columns = names
        hdus.append(hdu)
    return hdus


# ---------------------------------------------------------------------------
# writing
# ---------------------------------------------------------------------------
def _card(key, value, comment=None):
    if key == "":
        body = str(value) if value is not None else ""
        return ("        " + body)[:80].ljust(80)
    if isinstance(value, bool):
        v = "T" if value else "F"
        s = "{:<8}= {:>20}".format(key, v)
    elif isinstance(value, (int, np.integer)):
        s = "{:<8}= {:>20d}".format(key, int(value))
    elif isinstance(value, (float, np.floating)):
        r = repr(float(value)).upper()
        if "E" not in r and "." not in r and "N" not in r:
            r += "."
        s = "{:<8}= {:>20}".format(key, r)
    else:
        v = str(value).replace("'", "''")
        s = "{:<8}= '{:<8}'".format(key, v)
    if comment:
        s += " / " + str(comment)
    return s[:80].ljust(80)


def _header_bytes(cards):
    txt = "".join(cards) + "END".ljust(80)
    txt = txt.ljust(_padded(len(txt)))
    return txt.encode("ascii", "replace")


def _image_cards(data, primary, extra):
    cards = []
    if primary:
        cards.append(_card("SIMPLE", True, "conforms to FITS standard"))
    else:
        cards.append(_card("XTENSION", "IMAGE", "Image extension"))
    if data is None:
        cards += [_card("BITPIX", 8), _card("NAXIS", 0)]
    else:
        bitpix = _DTYPE_BITPIX[data.dtype.str[1:]]
        cards += [_card("BITPIX", bitpix), _card("NAXIS", data.ndim)]
        for i, n in enumerate(data.shape[::-1]):
            cards.append(_card("NAXIS%d" % (i + 1), n))
    if primary:
        cards.append(_card("EXTEND", True))
    else:
        cards += [_card("PCOUNT", 0), _card("GCOUNT", 1)]
    reserved = {"SIMPLE", "XTENSION", "BITPIX", "NAXIS", "EXTEND", "PCOUNT", "GCOUNT", "END"}
    if extra is not None:
        comments = getattr(extra, "comments", {})
        for k, v in extra.items():
            ku = str(k).upper()
            if ku in reserved or ku.startswith("NAXIS"):
                continue
            cards.append(_card(ku, v, comments.get(k)))
    return cards


def _data_bytes(data):
    if data is None:
        return b""
    be = np.ascontiguousarray(data, dtype=data.dtype.newbyteorder(">"))
    raw = be.tobytes()
    return raw + b"\0" * (_padded(len(raw)) - len(raw))


def table_hdu(columns, name=None, header=None):
    """Binary-table HDU from ``{name: 1-D array}`` (float32/float64/int columns)."""
    names = list(columns)
    n = len(columns[names[0]])
    fmts, codes = [], []
    for k in names:
        a = np.asarray(columns[k])
        kind = a.dtype.str[1:]
        code = {"f4": "E", "f8": "D", "i4": "J", "i8": "K", "i2": "I"}[kind]
        fmts.append(">" + kind)
        codes.append(code)
    dt = np.dtype({"names": names, "formats": fmts})
    rec = np.zeros(n, dtype=dt)
    for k in names:
        rec[k] = columns[k]
    hdu = HDU(rec, header if header is not None else Header(), name)
    hdu.columns = names
    hdu._tform = codes
    return hdu


def _table_bytes(hdu):
    rec = hdu.data
    cards = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8),
             _card("NAXIS", 2), _card("NAXIS1", rec.dtype.itemsize), _card("NAXIS2", len(rec)),
             _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(hdu.columns))]
    for i, (k, c) in enumerate(zip(hdu.columns, hdu._tform), 1):
        cards.append(_card("TTYPE%d" % i, k))
        cards.append(_card("TFORM%d" % i, c))
    for k, v in hdu.header.items():
        cards.append(_card(str(k).upper(), v))
    raw = rec.tobytes()
    return _header_bytes(cards) + raw + b"\0" * (_padded(len(raw)) - len(raw))


def writeto(path, hdus, overwrite=True):
    """Write a list of HDUs; the first is the primary."""
    if os.path.exists(path) and not overwrite:
        raise IOError("{} exists".format(path))
    with builtins.open(path, "wb") as f:
        for i, h in enumerate(hdus):
            if h.columns is not None and i > 0:
                f.write(_table_bytes(h))
                continue
            data = None if h.data is None else np.asarray(h.data)
            f.write(_header_bytes(_image_cards(data, i == 0, h.header)))
            f.write(_data_bytes(data))
