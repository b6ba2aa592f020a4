"""
tess_b200.fitsio -- the I/O edge of the path: S/N-map in, segmentation
map out, as single-HDU FITS images.

The reference reads and writes its maps with astropy (scripts/demo_iso_sn_accretor.py:26,
tess/tests/test_pixel_accretion.py:35: `astropy.io.fits.writeto("iso_sn.fits", segmap)`), which this
image does not have.  This is a minimal reader/writer for exactly that use: the PRIMARY HDU of a
file, BITPIX 8/16/32/64/-32/-64, big-endian data, BSCALE/BZERO applied on read when present.  No
extensions, no compression, no WCS handling (header cards are returned verbatim for the caller).
Host-side only; nothing here touches the GPU.
"""
import numpy as np

_BLOCK = 2880
_DTYPES = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_BITPIX = {"uint8": 8, "int16": 16, "int32": 32, "int64": 64, "float32": -32, "float64": -64}


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % int(value)
    elif isinstance(value, (float, np.floating)):
        v = "%20s" % ("%.16G" % float(value))
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
    s = "%-8s= %s" % (key, v)
    if comment:
        s += " / " + comment
    return ("%-80s" % s)[:80]


def _parse_value(s):
    s = s.split("/")[0].strip() if not s.strip().startswith("'") else s.strip()
    if s.startswith("'"):
        end = s.find("'", 1)
        while end != -1 and s[end:end + 2] == "''":
            end = s.find("'", end + 2)
        return s[1:end if end != -1 else None].replace("''", "'").rstrip()
    if s in ("T", "F"):
        return s == "T"
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace("D", "E"))
        except ValueError:
            return s


def read_image(path, raw=False):
    """(data, header) of the primary HDU.  `data` is a native-endian C-contiguous array shaped
    (NAXISn, ..., NAXIS1) -- rows are y, columns x, as astropy returns it; `header` is an ordered
    dict of the cards.  BSCALE/BZERO are applied (result float64) unless `raw`."""
    with open(path, "rb") as f:
        buf = f.read()
    header, pos, done = {}, 0, False
    while not done:
        block = buf[pos:pos + _BLOCK]
        if len(block) < _BLOCK:
            raise ValueError("%s: truncated FITS header" % path)
        pos += _BLOCK
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                header[key] = _parse_value(card[10:])
    if header.get("SIMPLE") is not True:
        raise ValueError("%s: not a standard FITS file (SIMPLE != T)" % path)
    bitpix, naxis = int(header["BITPIX"]), int(header["NAXIS"])
    if bitpix not in _DTYPES:
        raise ValueError("%s: unsupported BITPIX %d" % (path, bitpix))
    shape = tuple(int(header["NAXIS%d" % (i + 1)]) for i in range(naxis))[::-1]
    n = int(np.prod(shape)) if naxis else 0
    dt = np.dtype(_DTYPES[bitpix])
    if len(buf) < pos + n * dt.itemsize:
        raise ValueError("%s: truncated FITS data" % path)
    data = np.frombuffer(buf, dtype=dt, count=n, offset=pos).reshape(shape)
    data = data.astype(dt.newbyteorder("="))
    if not raw and ("BSCALE" in header or "BZERO" in header):
        bs, bz = float(header.get("BSCALE", 1.0)), float(header.get("BZERO", 0.0))
        if bs != 1.0 or bz != 0.0:
            data = data.astype(np.float64) * bs + bz
    return np.ascontiguousarray(data), header


def write_image(path, data, header=None):
    """Write `data` (2-D or n-D; uint8, int16/32/64, float32/64; bool and other ints are widened) as
    the primary HDU of a new file, as `astropy.io.fits.writeto(path, data)` does for the reference's
    segmentation maps.  `header`: extra {KEY: value} cards."""
    a = np.asarray(data)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype.name not in _BITPIX:
        if np.issubdtype(a.dtype, np.integer):
            a = a.astype(np.int64)
        elif np.issubdtype(a.dtype, np.floating):
            a = a.astype(np.float64)
        else:
            raise TypeError("write_image: unsupported dtype %s" % a.dtype)
    bitpix = _BITPIX[a.dtype.name]
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", bitpix, "array data type"),
             _card("NAXIS", a.ndim, "number of array dimensions")]
    for i, n in enumerate(a.shape[::-1]):
        cards.append(_card("NAXIS%d" % (i + 1), n))
    for k, v in (header or {}).items():
        k = str(k).upper()[:8]
        if k in ("SIMPLE", "BITPIX", "END") or k.startswith("NAXIS"):
            continue
        cards.append(_card(k, v))
    cards.append("%-80s" % "END")
    head = "".join(cards).encode("ascii")
    head += b" " * (-len(head) % _BLOCK)
    body = np.ascontiguousarray(a, dtype=np.dtype(_DTYPES[bitpix])).tobytes()
    body += b"\0" * (-len(body) % _BLOCK)
    with open(path, "wb") as f:
        f.write(head)
        f.write(body)
