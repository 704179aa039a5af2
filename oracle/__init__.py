"""CPU parity checker for the fluid-step hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product (``probabilistic_fluid_simulation_b200``) never does.

Two back ends with the same Python surface, and the input boundary:

* :class:`Oracle` -- ``oracle/liboracle.so``, the plain-C restatement in ``fluid_oracle.c`` (run-time
  sweep counts).
* ``oracle.png_decode`` -- the reference's PNG reader (libpng simplified API, 8-bit RGBA) restated without libpng
  (``png_restate.c``, also in ``liboracle.so``); pinned against the real libpng by tests/test_png_restatement.py.
* :class:`Reference` -- ``oracle/_ref/libfluid_ref_<N>.so``, the UNMODIFIED reference
  ``/root/reference/src/fluid.cpp`` compiled by ``oracle/Makefile`` through ``ref_wrap.cpp``
  (compile-time sweep count N, one library per N).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


class _Field(ctypes.Structure):
    # includes/fluid.hpp:17-22
    _fields_ = [("x", ctypes.c_int), ("y", ctypes.c_int), ("z", ctypes.c_int),
                ("data", ctypes.POINTER(ctypes.c_float))]


def build(quiet: bool = True) -> None:
    """Run ``make`` in oracle/ (restatement always; reference only if /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _as_field(arr: np.ndarray) -> _Field:
    assert arr.dtype == np.float32 and arr.ndim == 3 and arr.flags["C_CONTIGUOUS"], (arr.dtype, arr.shape)
    h, w, c = arr.shape
    return _Field(w, h, c, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))


def _addr(f: _Field) -> int:
    return ctypes.cast(f.data, ctypes.c_void_p).value


class _Pair:
    """Two numpy buffers behind two vp_field structs whose data pointers the C code may swap."""

    def __init__(self, a: np.ndarray, b: np.ndarray):
        self.bufs = {a.ctypes.data: a, b.ctypes.data: b}
        self.fa, self.fb = _as_field(a), _as_field(b)

    def resolve(self):
        return self.bufs[_addr(self.fa)], self.bufs[_addr(self.fb)]


def channel_hash(arr: np.ndarray, k: int) -> str:
    """64-bit FNV-1a over the 32-bit words of channel k, row-major."""
    words = np.ascontiguousarray(arr[..., k]).view(np.uint32).ravel()
    lib = Oracle.lib()
    plane = np.ascontiguousarray(words.view(np.float32))
    h = lib.oracle_channel_hash(plane.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                ctypes.c_size_t(plane.size), 1, 0)
    return f"{h:016x}"


def field_hashes(arr: np.ndarray) -> list[str]:
    return [channel_hash(arr, k) for k in range(arr.shape[-1])]


class Oracle:
    """ctypes view of oracle/liboracle.so (fluid_oracle.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(_HERE, "liboracle.so")
            if not os.path.exists(path):
                build()
            lib = ctypes.CDLL(path)
            F = ctypes.POINTER(_Field)
            f32 = ctypes.c_float
            lib.oracle_advect.argtypes = [F, F, f32]
            lib.oracle_advect_color.argtypes = [F, F, F, f32]
            lib.oracle_diffuse.argtypes = [F, F, f32, f32, ctypes.c_int]
            lib.oracle_compute_pressure.argtypes = [F, F, f32, ctypes.c_int]
            lib.oracle_subtract_pressure_gradient.argtypes = [F, F, f32]
            lib.oracle_simulate_fluid_step.argtypes = [F, F, f32, f32, ctypes.c_int, ctypes.c_int]
            lib.oracle_advect_color_step.argtypes = [F, F, F, f32]
            lib.oracle_run_steps.argtypes = [F, F, F, F, f32, f32, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            lib.oracle_add_forces_stochastic.argtypes = [F, f32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int]
            lib.oracle_add_forces_stochastic.restype = None
            lib.oracle_simulate_fluid_step_stochastic.argtypes = [F, F, f32, f32, ctypes.c_int, ctypes.c_int, f32,
                                                                  ctypes.c_uint64, ctypes.c_uint32]
            lib.oracle_simulate_fluid_step_stochastic.restype = None
            lib.oracle_add_forces.argtypes = [F, ctypes.POINTER(f32)]
            lib.oracle_add_forces.restype = None
            lib.oracle_simulate_fluid_step_forced.argtypes = [F, F, f32, f32, ctypes.c_int, ctypes.c_int, ctypes.POINTER(f32)]
            lib.oracle_simulate_fluid_step_forced.restype = None
            lib.oracle_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
            lib.oracle_philox4x32_10.restype = None
            lib.oracle_channel_hash.argtypes = [ctypes.POINTER(f32), ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
            lib.oracle_channel_hash.restype = ctypes.c_uint64
            lib.oracle_init_velocity_from_unit.argtypes = [ctypes.POINTER(f32), ctypes.c_size_t]
            lib.oracle_init_vtmp.argtypes = [ctypes.POINTER(f32), ctypes.c_size_t]
            lib.oracle_bytes_to_unit_float.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(f32), ctypes.c_size_t]
            lib.oracle_unit_float_to_bytes.argtypes = [ctypes.POINTER(f32), ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t]
            for name in ("oracle_advect", "oracle_advect_color", "oracle_diffuse", "oracle_compute_pressure",
                         "oracle_subtract_pressure_gradient", "oracle_simulate_fluid_step",
                         "oracle_advect_color_step", "oracle_run_steps", "oracle_init_velocity_from_unit",
                         "oracle_init_vtmp", "oracle_bytes_to_unit_float", "oracle_unit_float_to_bytes"):
                getattr(lib, name).restype = None
            cls._lib = lib
        return cls._lib

    def __init__(self, n_diffuse: int = 30, n_pressure: int | None = None):
        self.n_diffuse = int(n_diffuse)
        self.n_pressure = int(n_diffuse if n_pressure is None else n_pressure)
        self.L = self.lib()

    # --- single operators (arrays are modified in place; returns (vp, out) after pointer swaps) ---
    def advect(self, vp, out, dt):
        p = _Pair(vp, out)
        self.L.oracle_advect(p.fa, p.fb, dt)
        return p.resolve()

    def diffuse(self, vp, out, viscosity, dt, n=None):
        p = _Pair(vp, out)
        self.L.oracle_diffuse(p.fa, p.fb, viscosity, dt, self.n_diffuse if n is None else n)
        return p.resolve()

    def compute_pressure(self, vp, out, dt, n=None):
        p = _Pair(vp, out)
        self.L.oracle_compute_pressure(p.fa, p.fb, dt, self.n_pressure if n is None else n)
        return p.resolve()

    def subtract_pressure_gradient(self, vp, out, dt):
        p = _Pair(vp, out)
        self.L.oracle_subtract_pressure_gradient(p.fa, p.fb, dt)
        return p.resolve()

    def advect_color(self, image, itmp, vp, dt):
        p = _Pair(image, itmp)
        fv = _as_field(vp)
        self.L.oracle_advect_color(p.fa, p.fb, fv, dt)
        return p.resolve()

    def simulate_fluid_step(self, vp, tmp, dt, viscosity):
        p = _Pair(vp, tmp)
        self.L.oracle_simulate_fluid_step(p.fa, p.fb, dt, viscosity, self.n_diffuse, self.n_pressure)
        return p.resolve()

    def advect_color_step(self, image, itmp, vp, dt):
        p = _Pair(image, itmp)
        fv = _as_field(vp)
        self.L.oracle_advect_color_step(p.fa, p.fb, fv, dt)
        return p.resolve()

    # --- opt-in stochastic forcing (extension; the reference has no random term) ---
    def add_forces_stochastic(self, vp, sigma, seed, step, row0=0):
        self.L.oracle_add_forces_stochastic(_as_field(vp), sigma, seed, step, row0)
        return vp

    def simulate_fluid_step_stochastic(self, vp, tmp, dt, viscosity, sigma, seed, step):
        p = _Pair(vp, tmp)
        self.L.oracle_simulate_fluid_step_stochastic(p.fa, p.fb, dt, viscosity, self.n_diffuse, self.n_pressure, sigma,
                                                     seed, step)
        return p.resolve()

    # --- external force at the addForces slot (the reference body is empty: parity unpinned by construction) ---
    def add_forces(self, vp, forces):
        assert forces.dtype == np.float32 and forces.shape == vp.shape and forces.flags["C_CONTIGUOUS"]
        self.L.oracle_add_forces(_as_field(vp), forces.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        return vp

    def simulate_fluid_step_forced(self, vp, tmp, dt, viscosity, forces):
        assert forces.dtype == np.float32 and forces.shape == vp.shape and forces.flags["C_CONTIGUOUS"]
        p = _Pair(vp, tmp)
        self.L.oracle_simulate_fluid_step_forced(p.fa, p.fb, dt, viscosity, self.n_diffuse, self.n_pressure,
                                                 forces.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        return p.resolve()

    def run_steps(self, vp, vtmp, image, itmp, dt, viscosity, n_steps):
        """main.cpp:219-240 loop.  Returns (vp, vtmp, image, itmp) as the structs point at the end."""
        pv = _Pair(vp, vtmp)
        if image is not None:
            pi = _Pair(image, itmp)
            self.L.oracle_run_steps(pv.fa, pv.fb, pi.fa, pi.fb, dt, viscosity, self.n_diffuse, self.n_pressure, n_steps)
            return (*pv.resolve(), *pi.resolve())
        self.L.oracle_run_steps(pv.fa, pv.fb, None, None, dt, viscosity, self.n_diffuse, self.n_pressure, n_steps)
        return (*pv.resolve(), None, None)


class Reference:
    """ctypes view of oracle/_ref/libfluid_ref_<N>.so -- the reference's own fluid.cpp, unmodified."""

    _libs: dict[int, ctypes.CDLL] = {}

    @staticmethod
    def available(n: int = 30) -> bool:
        return os.path.exists(os.path.join(REF_DIR, f"libfluid_ref_{n}.so"))

    @classmethod
    def _load(cls, n: int):
        if n not in cls._libs:
            path = os.path.join(REF_DIR, f"libfluid_ref_{n}.so")
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path}: run `make -C oracle` where /root/reference exists")
            lib = ctypes.CDLL(path)
            F = ctypes.POINTER(_Field)
            f32 = ctypes.c_float
            lib.ref_num_jacobi_iters.restype = ctypes.c_int
            lib.ref_advect.argtypes = [F, F, f32]
            lib.ref_advect_color.argtypes = [F, F, F, f32]
            lib.ref_diffuse.argtypes = [F, F, f32, f32]
            lib.ref_compute_pressure.argtypes = [F, F, f32]
            lib.ref_subtract_pressure_gradient.argtypes = [F, F, f32]
            lib.ref_simulate_fluid_step.argtypes = [F, F, f32, f32]
            lib.ref_advect_color_step.argtypes = [F, F, F, f32]
            lib.ref_run_steps.argtypes = [F, F, F, F, f32, f32, ctypes.c_int]
            for name in ("ref_advect", "ref_advect_color", "ref_diffuse", "ref_compute_pressure",
                         "ref_subtract_pressure_gradient", "ref_simulate_fluid_step", "ref_advect_color_step",
                         "ref_run_steps"):
                getattr(lib, name).restype = None
            assert lib.ref_num_jacobi_iters() == n
            cls._libs[n] = lib
        return cls._libs[n]

    def __init__(self, n_iters: int = 30, n_pressure: int | None = None):
        if n_pressure is not None and n_pressure != n_iters:
            raise ValueError("the reference uses one NUM_JACOBI_ITERS for both loops (fluid.hpp:11)")
        self.n_diffuse = self.n_pressure = int(n_iters)
        self.L = self._load(int(n_iters))

    def advect(self, vp, out, dt):
        p = _Pair(vp, out)
        self.L.ref_advect(p.fa, p.fb, dt)
        return p.resolve()

    def diffuse(self, vp, out, viscosity, dt, n=None):
        assert n is None or n == self.n_diffuse
        p = _Pair(vp, out)
        self.L.ref_diffuse(p.fa, p.fb, viscosity, dt)
        return p.resolve()

    def compute_pressure(self, vp, out, dt, n=None):
        assert n is None or n == self.n_pressure
        p = _Pair(vp, out)
        self.L.ref_compute_pressure(p.fa, p.fb, dt)
        return p.resolve()

    def subtract_pressure_gradient(self, vp, out, dt):
        p = _Pair(vp, out)
        self.L.ref_subtract_pressure_gradient(p.fa, p.fb, dt)
        return p.resolve()

    def advect_color(self, image, itmp, vp, dt):
        p = _Pair(image, itmp)
        self.L.ref_advect_color(p.fa, p.fb, _as_field(vp), dt)
        return p.resolve()

    def simulate_fluid_step(self, vp, tmp, dt, viscosity):
        p = _Pair(vp, tmp)
        self.L.ref_simulate_fluid_step(p.fa, p.fb, dt, viscosity)
        return p.resolve()

    def advect_color_step(self, image, itmp, vp, dt):
        p = _Pair(image, itmp)
        self.L.ref_advect_color_step(p.fa, p.fb, _as_field(vp), dt)
        return p.resolve()

    def run_steps(self, vp, vtmp, image, itmp, dt, viscosity, n_steps):
        pv = _Pair(vp, vtmp)
        if image is not None:
            pi = _Pair(image, itmp)
            self.L.ref_run_steps(pv.fa, pv.fb, pi.fa, pi.fb, dt, viscosity, n_steps)
            return (*pv.resolve(), *pi.resolve())
        self.L.ref_run_steps(pv.fa, pv.fb, None, None, dt, viscosity, n_steps)
        return (*pv.resolve(), None, None)


# ---- driver-side initial conditions (src/main.cpp:170-195, includes/utils.hpp:82-84) ----

def bytes_to_unit_float(b: np.ndarray) -> np.ndarray:
    """utils.hpp:82-84: (float)byte / 255.0 (double divide, rounded to float)."""
    return (b.astype(np.float64) / 255.0).astype(np.float32)


def velocity_from_bytes(b: np.ndarray) -> np.ndarray:
    """utils.hpp:82-84 then main.cpp:170-179: v = (float)((double)v*2.0 - 1.0) on all 4 channels."""
    f = bytes_to_unit_float(b)
    return (f.astype(np.float64) * 2.0 - 1.0).astype(np.float32)


def initial_vtmp(h: int, w: int) -> np.ndarray:
    """main.cpp:188-195: (-1,-1,-1,+1) per cell."""
    t = np.full((h, w, 4), -1.0, dtype=np.float32)
    t[..., 3] = 1.0
    return t


def unit_float_to_bytes(x: np.ndarray) -> np.ndarray:
    """utils.hpp:129-131: (png_byte)(x*255.0) -- truncation."""
    return (x.astype(np.float64) * 255.0).astype(np.uint8)
