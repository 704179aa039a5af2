"""The C++ driver's PNG boundary (host/png_io.hpp; reference includes/utils.hpp:32-150) without a GPU: a small
program built from that header must decode to the floats the Python mirror (pngio.py + oracle.bytes_to_unit_float)
gives -- including a 16-bit file, which libpng gamma-encodes on the way down to 8 bits -- and must
write the bytes utils.hpp:129-131 prescribes (truncation, not rounding)."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import oracle
from probabilistic_fluid_simulation_b200 import fixtures, pngio

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pngio") / "png_io_check")
    cxx = os.environ.get("CXX", "g++")
    # the same run-time hint host/Makefile compiles into the driver: where the python on PATH keeps Pillow's libpng16
    import sysconfig
    hint = os.path.join(sysconfig.get_paths()["purelib"], "pillow.libs", "libpng16*.so*")
    r = subprocess.run([cxx, "-std=c++17", "-O1", f'-DPFS_LIBPNG_HINT="{hint}"', os.path.join(HERE, "png_io_check.cpp"),
                        "-o", out, "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return out


def _cpp_read(exe, png, tmp):
    out = os.path.join(tmp, "out.bin")
    r = subprocess.run([exe, "read", png, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    w, h = struct.unpack("<ii", raw[:8])
    return np.frombuffer(raw[8:], dtype=np.float32).reshape(h, w, 4)


def _png16_rgb(path, rgb16):
    """A minimal 16-bit RGB PNG (colour type 2, no gAMA/sRGB chunk, filter 0) like the bundled velocity fields."""
    h, w, _ = rgb16.shape

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    rows = b"".join(b"\x00" + rgb16[j].astype(">u2").tobytes() for j in range(h))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(rows, 6)) + chunk(b"IEND", b"")
    open(path, "wb").write(png)


def test_cpp_reader_equals_python_mirror_8bit(exe, tmp_path):
    rgba = np.random.default_rng(3).integers(0, 256, (37, 53, 4), dtype=np.uint8)
    p = str(tmp_path / "a.png")
    pngio.write_rgba8(p, rgba)
    got = _cpp_read(exe, p, str(tmp_path))
    want = oracle.bytes_to_unit_float(pngio.read_rgba8(p))
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_cpp_reader_equals_python_mirror_16bit(exe, tmp_path):
    """Every 16-bit sample value once (256x256 pixels, R = value, G = reversed, B = 0): the 16 -> 8 bit conversion is
    libpng's own in both readers; what is checked is the call sequence and the hand-declared png_image ABI in C++."""
    v = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    rgb = np.stack([v, v[::-1, ::-1], np.zeros_like(v)], axis=-1)
    p = str(tmp_path / "v16.png")
    _png16_rgb(p, rgb)
    got = _cpp_read(exe, p, str(tmp_path))
    b = pngio.read_rgba8(p)
    assert b.shape == (256, 256, 4) and (b[..., 3] == 255).all()
    # not a plain >> 8: 16-bit samples without a gAMA chunk are taken as linear light and gamma-encoded
    assert not np.array_equal(b[..., 0], (v >> 8).astype(np.uint8))
    assert np.all(np.diff(b[..., 0].reshape(-1).astype(np.int32)) >= 0)          # monotone in the sample value
    want = oracle.bytes_to_unit_float(b)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_cpp_writer_truncates_like_the_reference(exe, tmp_path):
    rng = np.random.default_rng(4)
    x = rng.random((19, 31, 4), dtype=np.float32)
    x[0, :8, 0] = np.float32([0.0, 1.0, 0.999999, 0.5, 127.5 / 255, 128 / 255, 254.999 / 255, 1 / 255])
    binf, out = str(tmp_path / "x.bin"), str(tmp_path / "x.png")
    with open(binf, "wb") as f:
        f.write(struct.pack("<ii", x.shape[1], x.shape[0]))
        f.write(x.tobytes())
    r = subprocess.run([exe, "write", binf, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(pngio.read_rgba8(out), fixtures.unit_float_to_bytes(x))      # utils.hpp:129-131


def test_cpp_reader_rejects_what_the_reference_rejects(exe, tmp_path):
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    for p in (str(bad), str(tmp_path / "missing.png")):
        assert subprocess.run([exe, "read", p, str(tmp_path / "o.bin")], capture_output=True).returncode == 1
