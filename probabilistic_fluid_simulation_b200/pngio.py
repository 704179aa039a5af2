"""libpng-exact PNG decode/encode, mirroring includes/utils.hpp:32-150 of the reference.

The reference reads every PNG through libpng's *simplified API* forced to 8-bit RGBA
(utils.hpp:53-66).  For 16-bit files without a gAMA chunk libpng gamma-encodes the samples on the
way down to 8 bits, so PIL/OpenCV decode those files to different bytes.  This
module calls the same libpng entry points through ctypes on the libpng16 that ships inside the
``pillow.libs`` wheel directory, with a hand-declared ``png_image`` struct.

Used by scripts/make_golden.py (fixture generation, in the build container) and by the
``fluidsim`` Python driver.  Not part of the GPU hot path.
"""
from __future__ import annotations

import ctypes
import glob
import os
import sysconfig
import zlib

import numpy as np

PNG_IMAGE_VERSION = 1
PNG_FORMAT_RGBA = 0x03


class PngImage(ctypes.Structure):
    _fields_ = [("opaque", ctypes.c_void_p), ("version", ctypes.c_uint32), ("width", ctypes.c_uint32),
                ("height", ctypes.c_uint32), ("format", ctypes.c_uint32), ("flags", ctypes.c_uint32),
                ("colormap_entries", ctypes.c_uint32), ("warning_or_error", ctypes.c_uint32),
                ("message", ctypes.c_char * 64)]


_lib = None


def libpng():
    global _lib
    if _lib is None:
        site = sysconfig.get_paths()["purelib"]
        cands = sorted(glob.glob(os.path.join(site, "pillow.libs", "libpng16*.so*")))
        cands += ["libpng16.so.16", "libpng16.so"]
        err = None
        for c in cands:
            try:
                _lib = ctypes.CDLL(c)
                break
            except OSError as e:  # pragma: no cover
                err = e
        if _lib is None:
            raise OSError(f"no libpng16 found ({err})")
        _lib.png_image_begin_read_from_file.argtypes = [ctypes.POINTER(PngImage), ctypes.c_char_p]
        _lib.png_image_begin_read_from_file.restype = ctypes.c_int
        _lib.png_image_finish_read.argtypes = [ctypes.POINTER(PngImage), ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_int32, ctypes.c_void_p]
        _lib.png_image_finish_read.restype = ctypes.c_int
        _lib.png_image_write_to_file.argtypes = [ctypes.POINTER(PngImage), ctypes.c_char_p, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
        _lib.png_image_write_to_file.restype = ctypes.c_int
        _lib.png_image_free.argtypes = [ctypes.POINTER(PngImage)]
        _lib.png_image_free.restype = None
    return _lib


def read_rgba8(path: str) -> np.ndarray:
    """utils.hpp:49-66: begin_read, force PNG_FORMAT_RGBA, finish_read.  Returns uint8 [H,W,4]."""
    lib = libpng()
    img = PngImage()
    img.version = PNG_IMAGE_VERSION
    if not lib.png_image_begin_read_from_file(ctypes.byref(img), path.encode()):
        raise IOError(f"{path}: {img.message.decode(errors='replace')}")
    img.format = PNG_FORMAT_RGBA
    buf = np.empty((img.height, img.width, 4), dtype=np.uint8)
    if not lib.png_image_finish_read(ctypes.byref(img), None, buf.ctypes.data, 0, None):
        raise IOError(f"{path}: {img.message.decode(errors='replace')}")
    return buf


def write_rgba8(path: str, rgba: np.ndarray) -> None:
    """utils.hpp:134: png_image_write_to_file of an 8-bit RGBA buffer."""
    lib = libpng()
    assert rgba.dtype == np.uint8 and rgba.ndim == 3 and rgba.shape[2] == 4
    rgba = np.ascontiguousarray(rgba)
    img = PngImage()
    img.version = PNG_IMAGE_VERSION
    img.height, img.width = rgba.shape[0], rgba.shape[1]
    img.format = PNG_FORMAT_RGBA
    if not lib.png_image_write_to_file(ctypes.byref(img), path.encode(), 0, rgba.ctypes.data, 0, None):
        raise IOError(f"{path}: {img.message.decode(errors='replace')}")


def crc32(buf: np.ndarray) -> str:
    return f"{zlib.crc32(np.ascontiguousarray(buf).tobytes()) & 0xffffffff:08x}"


if __name__ == "__main__":
    import sys
    for p in sys.argv[1:]:
        a = read_rgba8(p)
        print(p, a.shape, crc32(a))
