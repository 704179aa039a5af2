"""oracle/png_decode.py + oracle/png_restate.c restate what the reference's PNG reader does (includes/utils.hpp:49-66
through libpng's simplified API, 8-bit RGBA output).  Here the restatement is pinned against the real libpng of this
image (pngio.read_rgba8 -- the same entry points the reference calls): on every sample value of a 16-bit file, on
synthetic files of every restated pixel format with every scanline filter, with and without gAMA / sRGB chunks, and
-- where the reference tree is mounted -- on every input the reference ships."""
import glob
import os
import struct
import zlib

import numpy as np
import pytest

from oracle import png_decode
from probabilistic_fluid_simulation_b200 import pngio

REF_INPUTS = "/root/reference/inputs"
CHANNELS = {0: 1, 2: 3, 4: 2, 6: 4}


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)


def _filter_row(ftype, row, prev, bpp):
    """Forward filters of the PNG specification (9.2): `row`, `prev` are the raw bytes of this and the previous line."""
    out = bytearray(len(row))
    for i in range(len(row)):
        a = row[i - bpp] if i >= bpp else 0
        b = prev[i]
        c = prev[i - bpp] if i >= bpp else 0
        if ftype == 0:
            pred = 0
        elif ftype == 1:
            pred = a
        elif ftype == 2:
            pred = b
        elif ftype == 3:
            pred = (a + b) >> 1
        else:
            p = a + b - c
            pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
            pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
        out[i] = (row[i] - pred) & 0xff
    return bytes(out)


def encode_png(samples, depth, ctype, extra_chunks=(), filters=(0, 1, 2, 3, 4)):
    """samples: [H, W, channels] of uint8 / uint16.  Row j uses filter type filters[j % len(filters)]."""
    h, w, ch = samples.shape
    assert ch == CHANNELS[ctype]
    bpp = ch * depth // 8
    raw_rows = [samples[j].astype(">u2" if depth == 16 else "u1").tobytes() for j in range(h)]
    prev = bytes(len(raw_rows[0]))
    body = bytearray()
    for j, row in enumerate(raw_rows):
        f = filters[j % len(filters)]
        body += bytes([f]) + _filter_row(f, row, prev, bpp)
        prev = row
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    for tag, data in extra_chunks:
        png += _chunk(tag, data)
    return png + _chunk(b"IDAT", zlib.compress(bytes(body), 6)) + _chunk(b"IEND", b"")


def _both(png_bytes, tmp_path, name="t.png"):
    p = str(tmp_path / name)
    with open(p, "wb") as f:
        f.write(png_bytes)
    return png_decode.decode_rgba8(p), pngio.read_rgba8(p)


def test_every_16bit_sample_value(tmp_path):
    """R = v, G = 65535 - v, B = v with bytes swapped, A = v: all 65536 values through the colour path (gamma table of
    linear 16-bit input) and through the alpha path (plain scaling)."""
    v = np.arange(65536, dtype=np.uint32).reshape(256, 256)
    rgba = np.stack([v, 65535 - v, ((v & 0xff) << 8) | (v >> 8), v], axis=-1).astype(np.uint16)
    got, want = _both(encode_png(rgba, 16, 6, filters=(0,)), tmp_path)
    assert np.array_equal(got, want)
    assert np.array_equal(want[..., 3].reshape(-1), np.round(np.arange(65536) * 255 / 65535).astype(np.uint8))
    # known answer: CRC-32 of the 65536 red bytes as the libpng 1.6.53 / 1.6.55 / 1.6.56 builds of this image decode them
    assert zlib.crc32(want[..., 0].tobytes()) & 0xffffffff == 1178597


def test_table_is_the_published_algorithm():
    t = png_decode.gamma_16to8_table()
    assert t.shape == (2048,) and t[0] == 0 and t[-1] == 65535 and np.all(np.diff(t.astype(np.int64)) >= 0)
    assert np.all(t[:-1] % 257 == 0) or np.all(t[t < 65535] % 257 == 0)     # 8-bit values replicated into both bytes
    # top 11 bits only: 32 consecutive samples share an entry; the result is a gamma-1/2.2 encode within one level
    x = (np.arange(2048) + 0.5) / 2048.0
    approx = 255.0 * x ** (1 / 2.2)
    assert np.max(np.abs((t >> 8).astype(np.float64) - approx)[8:]) <= 1.5


@pytest.mark.parametrize("depth", [8, 16])
@pytest.mark.parametrize("ctype", [0, 2, 4, 6])
def test_formats_and_filters(depth, ctype, tmp_path):
    rng = np.random.default_rng(100 * depth + ctype)
    hi = 256 if depth == 8 else 65536
    s = rng.integers(0, hi, (23, 31, CHANNELS[ctype])).astype(np.uint16 if depth == 16 else np.uint8)
    s[0, :4] = 0
    s[1, :4] = hi - 1
    got, want = _both(encode_png(s, depth, ctype), tmp_path)
    assert got.shape == (23, 31, 4) and np.array_equal(got, want)


@pytest.mark.parametrize("depth", [8, 16])
@pytest.mark.parametrize("gamma", [45455, 100000, 50000, 47000, 22727, 250000])
def test_gamma_chunk(depth, gamma, tmp_path):
    """gAMA decides the correction: for 8-bit files 45455 (and anything within 5 %) is a no-op, for 16-bit files
    100000 is the default; other values build other tables."""
    rng = np.random.default_rng(gamma + depth)
    hi = 256 if depth == 8 else 65536
    s = rng.integers(0, hi, (40, 64, 4)).astype(np.uint16 if depth == 16 else np.uint8)
    s[:2] = (np.arange(128) * (hi // 128)).reshape(2, 64, 1)
    got, want = _both(encode_png(s, depth, 6, extra_chunks=[(b"gAMA", struct.pack(">I", gamma))]), tmp_path)
    assert np.array_equal(got, want)


def test_srgb_chunk_means_no_correction_for_16bit(tmp_path):
    s = np.random.default_rng(9).integers(0, 65536, (16, 16, 3)).astype(np.uint16)
    got, want = _both(encode_png(s, 16, 2, extra_chunks=[(b"sRGB", b"\x00")]), tmp_path)
    assert np.array_equal(got, want)
    assert np.array_equal(want[..., :3], np.round(s.astype(np.float64) * 255 / 65535).astype(np.uint8))


def test_unsupported_features_are_refused_not_guessed(tmp_path):
    s = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(png_decode.Unsupported):
        png_decode.decode_rgba8(encode_png(s, 8, 2, extra_chunks=[(b"tRNS", b"\x00\x00\x00\x00\x00\x00")]))
    bad = bytearray(encode_png(s, 8, 2))
    bad[28] = 1                                                     # interlace method 1 (CRC now wrong as well)
    with pytest.raises(ValueError):
        png_decode.decode_rgba8(bytes(bad))
    with pytest.raises(ValueError):
        png_decode.decode_rgba8(b"not a png")


@pytest.mark.skipif(not os.path.isdir(REF_INPUTS), reason="reference tree not mounted")
def test_every_input_the_reference_ships():
    files = sorted(glob.glob(os.path.join(REF_INPUTS, "**", "*.png"), recursive=True))
    assert len(files) >= 31
    for f in files:
        assert np.array_equal(png_decode.decode_rgba8(f), pngio.read_rgba8(f)), f


def test_committed_fixtures_still_decode_the_same():
    """tests/golden/png_*.npz hold libpng's decode of eleven bundled inputs and golden.json their CRC-32
    (scripts/make_golden.py); where the reference tree is mounted the restatement must reproduce both from the files."""
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    index = json.load(open(os.path.join(here, "golden", "golden.json")))["png_inputs"]
    assert len(index) >= 11
    if not os.path.isdir(REF_INPUTS):
        pytest.skip("reference tree not mounted")
    for stem, meta in index.items():
        rgba = png_decode.decode_rgba8(os.path.join(REF_INPUTS, meta["png"]))
        assert list(rgba.shape) == meta["shape"] and pngio.crc32(rgba) == meta["crc32"], stem
        assert np.array_equal(rgba, np.load(os.path.join(here, "golden", stem + ".npz"))["rgba"]), stem
