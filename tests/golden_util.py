"""Rebuilds the inputs of a tests/golden/golden.json case (shared by scripts/make_golden.py, the CPU
oracle tests and the GPU parity tests)."""
from __future__ import annotations

import json
import os

import numpy as np

from probabilistic_fluid_simulation_b200 import fixtures

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden() -> dict:
    with open(os.path.join(GOLD, "golden.json")) as f:
        return json.load(f)


def build_inputs(spec: dict):
    """-> (velocity bytes [H,W,4], image bytes [Hi,Wi,4] or None)."""
    kind = spec["kind"]
    if kind == "png":
        vel = np.load(os.path.join(GOLD, spec["velocity"]))["rgba"]
        img = np.load(os.path.join(GOLD, spec["image"]))["rgba"]
        return vel, img
    vh, vw = spec["vel_hw"]
    if kind == "formula":
        vel = fixtures.formula_velocity_bytes(vh, vw)
    elif kind == "smooth":
        vel = fixtures.smooth_velocity_bytes(vh, vw)
    elif kind == "random":
        vel = fixtures.random_velocity_bytes(vh, vw, spec.get("seed", 1234))
    else:
        raise ValueError(kind)
    img = None
    if spec.get("img_hw"):
        ih, iw = spec["img_hw"]
        img = (fixtures.formula_image_bytes(ih, iw) if spec.get("img_kind", "formula") == "formula"
               else fixtures.random_image_bytes(ih, iw, spec.get("img_seed", 4321)))
    return vel, img


def case_state(case: dict):
    vel, img = build_inputs(case["inputs"])
    return fixtures.make_state(vel, img)


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bit_equal(got: np.ndarray, want: np.ndarray, what: str = ""):
    g, w = bits(got), bits(want)
    if g.shape != w.shape:
        raise AssertionError(f"{what}: shape {g.shape} != {w.shape}")
    bad = g != w
    if bad.any():
        idx = np.argwhere(bad)
        first = tuple(idx[0])
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.size} words differ; first at {first}: "
                             f"got {got[first]!r} want {want[first]!r}; max |diff| {d.max():.3e}")
