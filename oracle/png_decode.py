"""PNG -> 8-bit RGBA as the reference's reader produces it (includes/utils.hpp:49-66 through libpng's simplified
API), restated without libpng.  TEST INFRASTRUCTURE: chunk parsing and zlib inflate here, scanline reconstruction and
libpng's gamma / 16->8 bit semantics in oracle/png_restate.c (which cites the libpng routines it follows).  Pinned
against the real libpng of this image by tests/test_png_restatement.py."""
from __future__ import annotations

import ctypes
import os
import struct
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class Unsupported(ValueError):
    """The file uses a PNG feature outside the restated set (interlace, palette, < 8 bits, tRNS, sBIT, iCCP)."""


def _c():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        _lib.png_restate_unfilter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_int]
        _lib.png_restate_unfilter.restype = ctypes.c_int
        _lib.png_restate_to_rgba8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int32, ctypes.c_void_p]
        _lib.png_restate_to_rgba8.restype = ctypes.c_int
        _lib.png_restate_16to8_table.argtypes = [ctypes.c_void_p, ctypes.c_int32]
        _lib.png_restate_16to8_table.restype = None
    return _lib


def chunks(data: bytes):
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos = 8
    while pos + 12 <= len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        if zlib.crc32(tag + body) & 0xffffffff != crc:
            raise ValueError(f"bad CRC in {tag!r} chunk")
        yield tag, body
        pos += 12 + n
        if tag == b"IEND":
            return


def gamma_16to8_table(gamma_val: int = 219998) -> np.ndarray:
    """libpng's "16 to 8" gamma table (2048 entries indexed by the top 11 bits of a 16-bit sample; each entry is the
    8-bit result times 257).  219998 = reciprocal(reciprocal2(100000, 220000)): linear 16-bit input, sRGB output."""
    t = np.zeros(2048, dtype=np.uint16)
    _c().png_restate_16to8_table(t.ctypes.data, gamma_val)
    return t


def decode_rgba8(src) -> np.ndarray:
    """`src`: path or bytes.  Returns uint8 [H, W, 4] == pngio.read_rgba8(path)."""
    data = src if isinstance(src, (bytes, bytearray)) else open(src, "rb").read()
    ihdr, idat, file_gamma = None, [], 0
    for tag, body in chunks(bytes(data)):
        if tag == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat.append(body)
        elif tag == b"gAMA" and not file_gamma:          # an sRGB chunk, if present, wins (PNG spec 11.3.3.5)
            (file_gamma,) = struct.unpack(">I", body)
        elif tag == b"sRGB":
            file_gamma = 45455
        elif tag in (b"tRNS", b"sBIT", b"iCCP", b"PLTE"):
            raise Unsupported(f"{tag.decode()} chunk")
    if ihdr is None or not idat:
        raise ValueError("IHDR or IDAT missing")
    w, h, depth, ctype, _, _, interlace = ihdr
    if interlace or depth not in (8, 16) or ctype not in (0, 2, 4, 6):
        raise Unsupported(f"bit depth {depth}, colour type {ctype}, interlace {interlace}")
    channels = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    bpp = channels * depth // 8
    rowbytes = w * bpp
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8).copy()
    if raw.size != h * (rowbytes + 1):
        raise ValueError("IDAT holds the wrong number of bytes")
    lib = _c()
    if lib.png_restate_unfilter(raw.ctypes.data, h, rowbytes, bpp):
        raise ValueError("invalid filter type")
    out = np.empty((h, w, 4), dtype=np.uint8)
    if lib.png_restate_to_rgba8(raw.ctypes.data, w, h, depth, ctype, file_gamma, out.ctypes.data):
        raise Unsupported("pixel format")
    return out
