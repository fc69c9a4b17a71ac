"""TEST INFRASTRUCTURE: a strict FITS reader written from the FITS standard
(v4.0: sections 3.1 blocks, 4.1-4.2 card images and fixed-format values, 4.4.1
mandatory keywords, 5.2-5.3 big-endian two's-complement / IEEE-754 data, 7.1
IMAGE extensions) and sharing NOTHING with wayne_b200/fitsio.py, so that a file
the product writes is checked by something other than the product's own reader.

It is deliberately unforgiving: every deviation from the standard raises
FitsFormatError with the byte offset.
"""
import re
import struct

import numpy as np

BLOCK = 2880
CARD = 80


class FitsFormatError(Exception):
    pass


_KEY_RE = re.compile(r'^[A-Z0-9_-]{1,8} *$')


def _fail(offset, msg):
    raise FitsFormatError('byte %d: %s' % (offset, msg))


def _parse_card(card, offset):
    """(keyword, value, comment) of one 80-character card image."""
    if len(card) != CARD:
        _fail(offset, 'card is not 80 characters')
    if any(ord(c) < 32 or ord(c) > 126 for c in card):
        _fail(offset, 'card holds a character outside printable ASCII (4.1.2.3)')
    key = card[:8]
    if key.strip() == '':
        return '', None, card[8:].rstrip()
    if not _KEY_RE.match(key) or key[0] == ' ':
        _fail(offset, 'bad keyword field %r (4.1.2.1: left-justified, A-Z 0-9 _ -)' % key)
    key = key.rstrip()
    if key in ('COMMENT', 'HISTORY'):
        return key, None, card[8:].rstrip()
    if key == 'END':
        if card[3:].strip():
            _fail(offset, 'END card must be blank after the keyword (4.4.1)')
        return key, None, None
    if card[8:10] != '= ':
        _fail(offset, 'keyword %s: value indicator "= " missing in columns 9-10 (4.1.2.2)' % key)
    body = card[10:]
    # character string: quote in column 11, closing quote at or after column 20 (fixed format 4.2.1.1)
    if body.startswith("'"):
        i, out = 1, []
        while True:
            if i >= len(body):
                _fail(offset, 'keyword %s: unterminated string' % key)
            if body[i] == "'":
                if i + 1 < len(body) and body[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(body[i])
            i += 1
        if i < 9:
            _fail(offset, 'keyword %s: closing quote before column 20 (4.2.1.1 fixed format)' % key)
        rest = body[i + 1:]
        value = ''.join(out).rstrip()
    else:
        field, sep, rest2 = body.partition('/')
        rest = (sep + rest2) if sep else ''
        fixed = body[:20]
        tok = field.strip()
        if tok in ('T', 'F'):
            if fixed[19] != tok:
                _fail(offset, 'keyword %s: logical value must sit in column 30 (4.2.2)' % key)
            value = tok == 'T'
        elif re.match(r'^[+-]?\d+$', tok):
            if fixed.strip() != tok or fixed[19] == ' ':
                _fail(offset, 'keyword %s: integer must be right-justified in columns 11-30 (4.2.3)' % key)
            value = int(tok)
        elif re.match(r'^[+-]?(\d+\.?\d*|\.\d+)([ED][+-]?\d+)?$', tok):
            if len(field.rstrip()) > 20 and fixed.strip() != tok:
                pass                      # free format allowed for reals longer than the fixed field
            value = float(tok.replace('D', 'E'))
        elif tok == '':
            value = None
        else:
            _fail(offset, 'keyword %s: unparsable value field %r' % (key, tok))
    rest = rest.strip()
    comment = None
    if rest:
        if not rest.startswith('/'):
            _fail(offset, 'keyword %s: text after the value must start with "/" (4.1.2.3)' % key)
        comment = rest[1:].strip()
    return key, value, comment


def _read_header(buf, pos):
    cards, seen_end = [], False
    start = pos
    while not seen_end:
        if pos + BLOCK > len(buf):
            _fail(pos, 'header runs past the end of the file')
        block = buf[pos:pos + BLOCK]
        try:
            text = block.decode('ascii')
        except UnicodeDecodeError:
            _fail(pos, 'non-ASCII byte in a header block')
        for i in range(0, BLOCK, CARD):
            card = text[i:i + CARD]
            if seen_end:
                if card.strip():
                    _fail(pos + i, 'non-blank card after END (4.4.1: fill with ASCII blanks)')
                continue
            k, v, c = _parse_card(card, pos + i)
            if k == 'END':
                seen_end = True
            else:
                cards.append((k, v, c))
        pos += BLOCK
    return cards, pos, start


def _check_mandatory(cards, primary, offset):
    keys = [k for k, _, _ in cards]
    d = {k: v for k, v, _ in cards if k}
    if primary:
        if keys[0] != 'SIMPLE' or d['SIMPLE'] is not True:
            _fail(offset, 'primary header must start with SIMPLE = T (4.4.1.1)')
    else:
        if keys[0] != 'XTENSION' or not isinstance(d['XTENSION'], str):
            _fail(offset, 'extension header must start with XTENSION (4.4.1.2)')
    if keys[1] != 'BITPIX' or d['BITPIX'] not in (8, 16, 32, 64, -32, -64):
        _fail(offset, 'BITPIX must be the second keyword and one of 8 16 32 64 -32 -64')
    if keys[2] != 'NAXIS' or not (0 <= d['NAXIS'] <= 999):
        _fail(offset, 'NAXIS must be the third keyword')
    n = d['NAXIS']
    for i in range(n):
        if keys[3 + i] != 'NAXIS%d' % (i + 1) or d['NAXIS%d' % (i + 1)] < 0:
            _fail(offset, 'NAXIS%d must follow NAXIS in order (4.4.1)' % (i + 1))
    if not primary:
        if keys[3 + n] != 'PCOUNT' or keys[4 + n] != 'GCOUNT':
            _fail(offset, 'PCOUNT and GCOUNT must follow the NAXISn of an extension (4.4.1.2)')
        if d['XTENSION'].strip() == 'IMAGE' and (d['PCOUNT'] != 0 or d['GCOUNT'] != 1):
            _fail(offset, 'IMAGE extension needs PCOUNT = 0 and GCOUNT = 1 (7.1.1)')
    dup = {k for k in keys if k and k not in ('COMMENT', 'HISTORY') and keys.count(k) > 1}
    if dup:
        _fail(offset, 'keyword(s) appear more than once: %s' % sorted(dup))
    return d


_DTYPES = {8: '>u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


def read(path):
    """[(header dict, ordered cards, ndarray or None)] for every HDU of the file."""
    with open(path, 'rb') as f:
        buf = f.read()
    if len(buf) % BLOCK:
        raise FitsFormatError('file length %d is not a multiple of 2880 (3.1)' % len(buf))
    out, pos, primary = [], 0, True
    while pos < len(buf):
        cards, pos, start = _read_header(buf, pos)
        d = _check_mandatory(cards, primary, start)
        if primary and len(buf) > pos and d.get('EXTEND') is not True:
            _fail(start, 'extensions follow but EXTEND = T is missing from the primary header')
        n = d['NAXIS']
        shape = [d['NAXIS%d' % (i + 1)] for i in range(n)]
        nelem = int(np.prod(shape)) if n else 0
        nbytes = abs(d['BITPIX']) // 8 * d.get('GCOUNT', 1) * (d.get('PCOUNT', 0) + nelem)
        data = None
        if nbytes:
            if pos + nbytes > len(buf):
                _fail(pos, 'data unit runs past the end of the file')
            padded = (nbytes + BLOCK - 1) // BLOCK * BLOCK
            if any(buf[pos + nbytes:pos + padded]):
                _fail(pos + nbytes, 'data unit must be padded with zero bytes (3.3.2)')
            if primary or d['XTENSION'].strip() == 'IMAGE':
                data = np.frombuffer(buf, dtype=_DTYPES[d['BITPIX']], count=nelem, offset=pos).reshape(shape[::-1])
            pos += padded
        out.append((d, cards, data))
        primary = False
    return out


def f64_be(x):
    """The eight big-endian IEEE-754 bytes of a double (5.3), for spot checks."""
    return struct.pack('>d', float(x))
