"""Driver-side initial conditions (src/main.cpp:143-195) and synthetic input generators.

Host logic only (numpy): how the reference driver turns decoded PNG bytes into the float buffers
the fluid step consumes, plus the synthetic grids named by BASELINE.json's configs.  Integer
formulas and seeded generators so that every box produces byte-identical inputs.
"""
from __future__ import annotations

import numpy as np


def bytes_to_unit_float(b: np.ndarray) -> np.ndarray:
    """includes/utils.hpp:82-84: x = (float)byte / 255.0 (double divide, rounded to float)."""
    return (b.astype(np.float64) / 255.0).astype(np.float32)


def velocity_from_bytes(b: np.ndarray) -> np.ndarray:
    """utils.hpp:82-84 then main.cpp:170-179: v = (float)((double)v * 2.0 - 1.0) on all 4 channels."""
    return (bytes_to_unit_float(b).astype(np.float64) * 2.0 - 1.0).astype(np.float32)


def initial_vtmp(h: int, w: int) -> np.ndarray:
    """main.cpp:188-195: the temporary velocity buffer starts as (-1,-1,-1,+1) per cell; its
    channel 2 is the pressure warm start of the first step."""
    t = np.full((h, w, 4), -1.0, dtype=np.float32)
    t[..., 3] = 1.0
    return t


def unit_float_to_bytes(x: np.ndarray) -> np.ndarray:
    """utils.hpp:129-131: (png_byte)(x * 255.0) -- truncation toward zero."""
    return (x.astype(np.float64) * 255.0).astype(np.uint8)


# ---- synthetic inputs ---------------------------------------------------------------------------

def formula_velocity_bytes(h: int, w: int) -> np.ndarray:
    """R=(4i+2j)&255, G=(3i+5j+64)&255, B=0, A=255 (i = column, j = row)."""
    i = np.arange(w, dtype=np.int64)[None, :]
    j = np.arange(h, dtype=np.int64)[:, None]
    b = np.zeros((h, w, 4), np.uint8)
    b[..., 0] = (4 * i + 2 * j) & 255
    b[..., 1] = (3 * i + 5 * j + 64) & 255
    b[..., 3] = 255
    return b


def formula_image_bytes(h: int, w: int) -> np.ndarray:
    """R=(i^j)&255, G=(2i+j)&255, B=(i+2j)&255, A=255."""
    i = np.arange(w, dtype=np.int64)[None, :]
    j = np.arange(h, dtype=np.int64)[:, None]
    b = np.zeros((h, w, 4), np.uint8)
    b[..., 0] = (i ^ j) & 255
    b[..., 1] = (2 * i + j) & 255
    b[..., 2] = (i + 2 * j) & 255
    b[..., 3] = 255
    return b


def _triangle(t: np.ndarray, period: int) -> np.ndarray:
    """Integer triangle wave in [0, 255] with the given period (exact integer arithmetic)."""
    ph = np.mod(t, period)
    half = period // 2
    up = np.where(ph < half, ph, period - ph)          # 0 .. half
    return (up * 255) // max(half, 1)


def smooth_velocity_bytes(h: int, w: int, rows: tuple[int, int] | None = None) -> np.ndarray:
    """Sum of four integer-phase triangle waves per component (coherent flow: neighbouring cells
    have neighbouring departure points).  Periodic in both axes when h, w are multiples of 64.
    `rows = (r0, r1)` returns only rows r0..r1-1 of the h x w field (multi-GPU bands)."""
    r0, r1 = rows if rows is not None else (0, h)
    i = np.arange(w, dtype=np.int64)[None, :]
    j = np.arange(r0, r1, dtype=np.int64)[:, None]
    px, py = max(w // 4, 2), max(h // 4, 2)
    u = (_triangle(i + 0 * j, px) + _triangle(j + 0 * i, py) + _triangle(i + j, max(w // 2, 2)) + _triangle(3 * i - j + 7 * w, max(w // 8, 2))) // 4
    v = (_triangle(j + 0 * i + py // 3, py) + _triangle(i + 0 * j + px // 5, px) + _triangle(2 * j - i + 5 * h, max(h // 2, 2)) + _triangle(i + 3 * j, max(h // 8, 2))) // 4
    b = np.zeros((r1 - r0, w, 4), np.uint8)
    b[..., 0] = np.clip(u, 0, 255)
    b[..., 1] = np.clip(v, 0, 255)
    b[..., 3] = 255
    return b


def hash_bytes(h: int, w: int, channels: int, seed: int, rows: tuple[int, int] | None = None) -> np.ndarray:
    """Position-hashed pseudo-random bytes [rows, w, channels]: value(i, j, k) depends only on the
    cell, so any band of rows equals the same rows of the whole field (multi-GPU inputs)."""
    r0, r1 = rows if rows is not None else (0, h)
    i = np.arange(w, dtype=np.uint32)[None, :, None]
    j = np.arange(r0, r1, dtype=np.uint32)[:, None, None]
    k = np.arange(channels, dtype=np.uint32)[None, None, :]
    x = i * np.uint32(0x9E3779B1) + j * np.uint32(0x85EBCA77) + k * np.uint32(0xC2B2AE3D) + np.uint32(seed & 0xFFFFFFFF)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x2C1B3C6D)
    x ^= x >> np.uint32(12)
    x *= np.uint32(0x297A2D39)
    x ^= x >> np.uint32(15)
    return (x & np.uint32(255)).astype(np.uint8)


def random_velocity_bytes(h: int, w: int, seed: int = 1234) -> np.ndarray:
    """Seeded white-noise bytes in R,G; B=0, A=255 (as every bundled velocity PNG)."""
    rng = np.random.default_rng(seed)
    b = np.zeros((h, w, 4), np.uint8)
    b[..., :2] = rng.integers(0, 256, size=(h, w, 2), dtype=np.uint8)
    b[..., 3] = 255
    return b


def random_image_bytes(h: int, w: int, seed: int = 4321) -> np.ndarray:
    rng = np.random.default_rng(seed)
    b = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    b[..., 3] = 255
    return b


def make_state(vel_bytes: np.ndarray, img_bytes: np.ndarray | None = None):
    """(vp, vtmp, image, itmp) float32 host arrays exactly as main.cpp builds them before the loop."""
    h, w, _ = vel_bytes.shape
    vp = velocity_from_bytes(vel_bytes)
    vtmp = initial_vtmp(h, w)
    if img_bytes is None:
        return vp, vtmp, None, None
    image = bytes_to_unit_float(img_bytes)
    itmp = np.zeros_like(image)          # main.cpp:186 leaves it uninitialised; it is fully overwritten
    return vp, vtmp, image, itmp
