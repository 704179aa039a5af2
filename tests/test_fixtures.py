"""Driver-side initial conditions (main.cpp:170-195, utils.hpp:82-84,129-131): the package's numpy
host logic against the oracle's C restatement of the same lines."""
import ctypes

import numpy as np

import oracle
from golden_util import assert_bit_equal
from probabilistic_fluid_simulation_b200 import fixtures


def test_byte_to_float_all_values():
    b = np.arange(256, dtype=np.uint8)
    want = np.empty(256, np.float32)
    L = oracle.Oracle.lib()
    L.oracle_bytes_to_unit_float(b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                 want.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 256)
    assert_bit_equal(fixtures.bytes_to_unit_float(b), want, "byte/255.0")
    vel = want.copy()
    L.oracle_init_velocity_from_unit(vel.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 256)
    assert_bit_equal(fixtures.velocity_from_bytes(b), vel, "v*2-1")
    back = np.empty(256, np.uint8)
    L.oracle_unit_float_to_bytes(want.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                 back.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 256)
    assert np.array_equal(back, b)                       # exact round trip
    assert np.array_equal(fixtures.unit_float_to_bytes(want), b)


def test_initial_vtmp():
    t = fixtures.initial_vtmp(3, 5)
    want = np.empty(3 * 5 * 4, np.float32)
    oracle.Oracle.lib().oracle_init_vtmp(want.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), want.size)
    assert_bit_equal(t, want.reshape(3, 5, 4), "vtmp")


def test_generators_are_deterministic_and_shaped():
    a = fixtures.random_velocity_bytes(16, 24, 5)
    b = fixtures.random_velocity_bytes(16, 24, 5)
    assert np.array_equal(a, b) and a.shape == (16, 24, 4)
    assert (a[..., 2] == 0).all() and (a[..., 3] == 255).all()   # B=0, A=max
    s = fixtures.smooth_velocity_bytes(64, 128)
    assert s.shape == (64, 128, 4) and s[..., 0].max() <= 255
    # periodic continuity of the smooth field: wrap-around neighbours differ little
    assert np.abs(s[:, 0, 0].astype(int) - s[:, -1, 0].astype(int)).max() < 48
