"""Host-side mirror of the reference operator interface ``includes/fluid.hpp`` on top of the C-ABI.

Same names, argument order and buffer semantics as the reference:

    vp_field                               fluid.hpp:17-22
    advect(vp, vp_out, dt)                 fluid.hpp:32   / fluid.cpp:24-70
    advect_color(image, out, vp, dt)       fluid.hpp:45   / fluid.cpp:72-127
    diffuse(vp, vp_out, viscosity, dt)     fluid.hpp:58   / fluid.cpp:129-196
    addForces(vp, forces)                  fluid.hpp:70   / fluid.cpp:198-208 (empty in the reference)
    computePressure(vp, vp_out, dt)        fluid.hpp:81   / fluid.cpp:210-267
    subtractPressureGradient(vp, out, dt)  fluid.hpp:92   / fluid.cpp:269-296
    simulate_fluid_step(vp, tmp, dt, visc) fluid.hpp:109  / fluid.cpp:298-305
    advect_color_step(image, itmp, vp, dt) fluid.hpp:118  / fluid.cpp:312-320

``vp_field.data`` is either a CUDA ``torch.Tensor`` of shape [H, W, 4] float32 (the reference's
USE_CUDA build: caller-owned device buffers) or a C-contiguous ``numpy`` array of the same shape
(the reference's CPU build: host buffers; the call then uploads, computes on the GPU and downloads).
Like the reference, operators exchange the ``data`` members of the two structs they are given.
The sweep counts, compile-time ``NUM_JACOBI_ITERS`` in the reference (fluid.hpp:11), are keyword
arguments defaulting to 30.

Every call goes to hand-written CUDA kernels through ``libpfs_b200.so``; nothing here computes.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from ._cabi import Field, check

NUM_JACOBI_ITERS = 30  # fluid.hpp:11


class vp_field:  # noqa: N801  (reference spelling)
    """x = width, y = height, z = channels (4), data = interleaved [H, W, 4] float32 buffer."""

    __slots__ = ("x", "y", "z", "data")

    def __init__(self, data):
        if data.ndim != 3 or data.shape[2] != 4:
            raise ValueError(f"expected [H, W, 4], got {tuple(data.shape)}")
        self.y, self.x, self.z = int(data.shape[0]), int(data.shape[1]), int(data.shape[2])
        self.data = data

    @property
    def on_device(self) -> bool:
        return not isinstance(self.data, np.ndarray)


def _is_torch(t) -> bool:
    return not isinstance(t, np.ndarray)


def _check_buf(f: vp_field, name: str):
    d = f.data
    if _is_torch(d):
        import torch
        if not (d.is_cuda and d.dtype == torch.float32 and d.is_contiguous()):
            raise ValueError(f"{name}.data must be a contiguous float32 CUDA tensor")
    else:
        if not (d.dtype == np.float32 and d.flags["C_CONTIGUOUS"]):
            raise ValueError(f"{name}.data must be a C-contiguous float32 array")
    if tuple(d.shape) != (f.y, f.x, f.z):
        raise ValueError(f"{name}: data shape {tuple(d.shape)} does not match (y, x, z) = {(f.y, f.x, f.z)}")


def _stream_of(t) -> ctypes.c_void_p:
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _dev_guard(t):
    import torch
    return torch.cuda.device(t.device)


class _Handles:
    """float** arguments for a set of device tensors; after the call, writes the (possibly
    exchanged) buffers back into the structs."""

    def __init__(self, *fields: vp_field):
        self.fields = fields
        self.by_ptr = {f.data.data_ptr(): f.data for f in fields}
        self.slots = [ctypes.c_void_p(f.data.data_ptr()) for f in fields]

    def ref(self, i):
        return ctypes.byref(self.slots[i])

    def commit(self):
        for f, s in zip(self.fields, self.slots):
            f.data = self.by_ptr[s.value]


def _host_struct(f: vp_field) -> Field:
    return Field(f.x, f.y, f.z, f.data.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))


def _commit_host(pairs):
    by_ptr = {f.data.ctypes.data: f.data for f, _ in pairs}
    for f, st in pairs:
        f.data = by_ptr[ctypes.cast(st.data, ctypes.c_void_p).value]


# ------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------

def advect(vp: vp_field, vp_out: vp_field, dt: float) -> None:
    L = _cabi.lib()
    _check_buf(vp, "vp"); _check_buf(vp_out, "vp_out")
    _require_device("advect", vp, vp_out)
    with _dev_guard(vp.data):
        check(L.pfs_advect(vp.data.data_ptr(), vp_out.data.data_ptr(), dt, vp.x, vp.y, vp.z, _stream_of(vp.data)))


def advect_color(image: vp_field, out: vp_field, vp: vp_field, dt: float) -> None:
    L = _cabi.lib()
    for f, n in ((image, "image"), (out, "out"), (vp, "vp")):
        _check_buf(f, n)
    _require_device("advect_color", image, out, vp)
    with _dev_guard(vp.data):
        check(L.pfs_advect_color(image.data.data_ptr(), out.data.data_ptr(), vp.data.data_ptr(), dt,
                                 image.x, image.y, image.z, vp.x, vp.y, vp.z, _stream_of(vp.data)))


def diffuse(vp: vp_field, vp_out: vp_field, viscosity: float, dt: float, n_sweeps: int = NUM_JACOBI_ITERS) -> None:
    L = _cabi.lib()
    _check_buf(vp, "vp"); _check_buf(vp_out, "vp_out")
    _require_device("diffuse", vp, vp_out)
    h = _Handles(vp, vp_out)
    with _dev_guard(vp.data):
        check(L.pfs_diffuse(h.ref(0), h.ref(1), viscosity, dt, vp.x, vp.y, vp.z, n_sweeps, _stream_of(vp.data)))
    h.commit()


def addForces(vp: vp_field, forces=None) -> None:  # noqa: N802
    """forces: None (the reference's empty body, fluid.cpp:198-208) or a CUDA tensor shaped like vp.data whose channels
    0,1 are added to the velocity (pfs_add_forces)."""
    L = _cabi.lib()
    if forces is not None and tuple(forces.shape) != tuple(vp.data.shape):
        raise ValueError("forces must be shaped like vp.data")
    _check_buf(vp, "vp")
    _require_device("addForces", vp)
    with _dev_guard(vp.data):
        check(L.pfs_add_forces(vp.data.data_ptr(), None if forces is None else forces.data_ptr(), vp.x, vp.y, vp.z,
                               _stream_of(vp.data)))


def computePressure(vp: vp_field, vp_out: vp_field, dt: float, n_sweeps: int = NUM_JACOBI_ITERS) -> None:  # noqa: N802
    L = _cabi.lib()
    _check_buf(vp, "vp"); _check_buf(vp_out, "vp_out")
    _require_device("computePressure", vp, vp_out)
    h = _Handles(vp, vp_out)
    with _dev_guard(vp.data):
        check(L.pfs_compute_pressure(h.ref(0), h.ref(1), dt, vp.x, vp.y, vp.z, n_sweeps, _stream_of(vp.data)))
    h.commit()


def computePressureAdaptive(vp: vp_field, vp_out: vp_field, dt: float, tol: float, max_sweeps: int,  # noqa: N802
                            check_every: int = 16) -> tuple[int, float]:
    """computePressure with the sweep count chosen at run time (pfs_compute_pressure_adaptive; not in the reference,
    which fixes NUM_JACOBI_ITERS).  -> (sweeps done, rms of the last update).  The buffers hold exactly what
    computePressure(n_sweeps=sweeps done) would have left."""
    L = _cabi.lib()
    _check_buf(vp, "vp"); _check_buf(vp_out, "vp_out")
    _require_device("computePressureAdaptive", vp, vp_out)
    h = _Handles(vp, vp_out)
    n, rms = ctypes.c_int(0), ctypes.c_double(0.0)
    with _dev_guard(vp.data):
        check(L.pfs_compute_pressure_adaptive(h.ref(0), h.ref(1), dt, vp.x, vp.y, vp.z, tol, max_sweeps, check_every,
                                              ctypes.byref(n), ctypes.byref(rms), _stream_of(vp.data)))
    h.commit()
    return n.value, rms.value


def computePressureSOR(vp: vp_field, vp_out: vp_field, dt: float, omega: float, tol: float, max_sweeps: int,  # noqa: N802
                       check_every: int = 8) -> tuple[int, float]:
    """Red-black SOR instead of the reference's Jacobi sweeps (pfs_compute_pressure_sor): NOT a parity path.  Divergence into
    channel 3 of both buffers, relaxed pressure into channel 2 of vp_out.  -> (sweeps done, rms of the last sweep's update)."""
    L = _cabi.lib()
    _check_buf(vp, "vp"); _check_buf(vp_out, "vp_out")
    _require_device("computePressureSOR", vp, vp_out)
    n, rms = ctypes.c_int(0), ctypes.c_double(0.0)
    with _dev_guard(vp.data):
        check(L.pfs_compute_pressure_sor(vp.data.data_ptr(), vp_out.data.data_ptr(), dt, vp.x, vp.y, vp.z, omega, tol, max_sweeps,
                                         check_every, ctypes.byref(n), ctypes.byref(rms), _stream_of(vp.data)))
    return n.value, rms.value


def subtractPressureGradient(vp: vp_field, vp_out: vp_field, dt: float) -> None:  # noqa: N802
    L = _cabi.lib()
    _check_buf(vp, "vp"); _check_buf(vp_out, "vp_out")
    _require_device("subtractPressureGradient", vp, vp_out)
    with _dev_guard(vp.data):
        check(L.pfs_subtract_pressure_gradient(vp.data.data_ptr(), vp_out.data.data_ptr(), dt, vp.x, vp.y, vp.z,
                                               _stream_of(vp.data)))


# ------------------------------------------------------------------------------------------------
# the two entry points the reference driver calls (main.cpp:222,225 / :236,239)
# ------------------------------------------------------------------------------------------------

def add_forces_stochastic(vp: vp_field, sigma: float, seed: int, step: int) -> None:
    """Opt-in extension (not in the reference): vp ch0,1 += sigma * N(0,1), Philox-4x32-10 keyed by
    `seed`, counter (cell, step).  See include/pfs_b200.h."""
    L = _cabi.lib()
    _check_buf(vp, "vp")
    _require_device("add_forces_stochastic", vp)
    with _dev_guard(vp.data):
        check(L.pfs_add_forces_stochastic(vp.data.data_ptr(), sigma, seed, step, vp.x, vp.y, vp.z, _stream_of(vp.data)))


def simulate_fluid_step(vp: vp_field, tmp: vp_field, dt: float, viscosity: float,
                        n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None,
                        sigma: float = 0.0, seed: int = 0, step: int = 0, forces=None) -> None:
    """sigma != 0 (device buffers only) adds the opt-in stochastic forcing at the addForces slot; `forces` (a CUDA tensor
    shaped like vp.data) applies addForces(vp, forces) there (fluid.cpp:302; pfs_simulate_fluid_step_forced)."""
    L = _cabi.lib()
    n_pressure = n_diffuse if n_pressure is None else n_pressure
    _check_buf(vp, "vp"); _check_buf(tmp, "tmp")
    if (vp.x, vp.y) != (tmp.x, tmp.y):
        raise ValueError("vp and tmp must have the same shape")
    if vp.on_device != tmp.on_device:
        raise ValueError("vp and tmp must both be device tensors or both be host arrays")
    if vp.on_device:
        h = _Handles(vp, tmp)
        with _dev_guard(vp.data):
            if forces is not None:
                if sigma != 0.0:
                    raise ValueError("forces and sigma cannot be combined in one call")
                if tuple(forces.shape) != tuple(vp.data.shape):
                    raise ValueError("forces must be shaped like vp.data")
                check(L.pfs_simulate_fluid_step_forced(h.ref(0), h.ref(1), dt, viscosity, vp.x, vp.y, vp.z, n_diffuse,
                                                       n_pressure, forces.data_ptr(), _stream_of(vp.data)))
            elif sigma != 0.0:
                check(L.pfs_simulate_fluid_step_stochastic(h.ref(0), h.ref(1), dt, viscosity, vp.x, vp.y, vp.z,
                                                           n_diffuse, n_pressure, sigma, seed, step,
                                                           _stream_of(vp.data)))
            else:
                check(L.pfs_simulate_fluid_step(h.ref(0), h.ref(1), dt, viscosity, vp.x, vp.y, vp.z,
                                                n_diffuse, n_pressure, _stream_of(vp.data)))
        h.commit()
    else:
        if sigma != 0.0 or forces is not None:
            raise ValueError("the stochastic / external forcing is available on device buffers only")
        sv, st = _host_struct(vp), _host_struct(tmp)
        pairs = [(vp, sv), (tmp, st)]
        check(L.pfs_simulate_fluid_step_host(ctypes.byref(sv), ctypes.byref(st), dt, viscosity, n_diffuse, n_pressure))
        _commit_host(pairs)


class FluidContext:
    """Persistent-state context (pfs_ctx_*): the fields live on the device in the library's planar layout between steps;
    the interleaved [H, W, 4] form exists only in upload() and download().  n steps leave behind, bit for bit, what n
    calls of simulate_fluid_step + advect_color_step leave in the caller's buffers.

        ctx = FluidContext(vx, vy, ix, iy); ctx.upload(vp, vtmp, image); ctx.step(100, dt, nu); vp, vtmp, image = ctx.download()
    """

    def __init__(self, vx: int, vy: int, ix: int = 0, iy: int = 0, device=None):
        import torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.vx, self.vy, self.ix, self.iy = vx, vy, ix, iy
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(_cabi.lib().pfs_ctx_create(ctypes.byref(self._h), vx, vy, ix, iy))

    def close(self):
        if self._h:
            _cabi.lib().pfs_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t, shape):
        if t is None:
            return None
        import torch
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape):
            raise ValueError(f"expected a contiguous float32 CUDA tensor of shape {shape}")
        return t.data_ptr()

    def upload(self, vp=None, vtmp=None, image=None) -> None:
        """CUDA tensors [vy, vx, 4] / [iy, ix, 4]; None leaves that buffer's state alone."""
        v, i = (self.vy, self.vx, 4), (self.iy, self.ix, 4)
        check(_cabi.lib().pfs_ctx_upload(self._h, self._ptr(vp, v), self._ptr(vtmp, v), self._ptr(image, i), self._stream()))

    def download(self, image_only: bool = False):
        """-> (vp, vtmp, image) as new CUDA tensors (image None without one)."""
        import torch
        vp = vtmp = image = None
        if not image_only:
            vp = torch.empty((self.vy, self.vx, 4), dtype=torch.float32, device=self.device)
            vtmp = torch.empty_like(vp)
        if self.ix > 0:
            image = torch.empty((self.iy, self.ix, 4), dtype=torch.float32, device=self.device)
        check(_cabi.lib().pfs_ctx_download(self._h, None if vp is None else vp.data_ptr(), None if vtmp is None else vtmp.data_ptr(),
                                           None if image is None else image.data_ptr(), self._stream()))
        return vp, vtmp, image

    def step(self, n_steps: int, dt: float, viscosity: float, n_diffuse: int = NUM_JACOBI_ITERS,
             n_pressure: int | None = None) -> None:
        n_pressure = n_diffuse if n_pressure is None else n_pressure
        check(_cabi.lib().pfs_ctx_step(self._h, n_steps, dt, viscosity, n_diffuse, n_pressure, self._stream()))

    def simulate_fluid_step(self, dt: float, viscosity: float, n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None,
                            sigma: float = 0.0, seed: int = 0, step: int = 0, forces=None) -> None:
        n_pressure = n_diffuse if n_pressure is None else n_pressure
        L = _cabi.lib()
        if forces is not None:
            check(L.pfs_ctx_simulate_fluid_step_forced(self._h, dt, viscosity, n_diffuse, n_pressure,
                                                       self._ptr(forces, (self.vy, self.vx, 4)), self._stream()))
        elif sigma != 0.0:
            check(L.pfs_ctx_simulate_fluid_step_stochastic(self._h, dt, viscosity, n_diffuse, n_pressure, sigma, seed, step,
                                                           self._stream()))
        else:
            check(L.pfs_ctx_simulate_fluid_step(self._h, dt, viscosity, n_diffuse, n_pressure, self._stream()))

    def advect_color_step(self, dt: float, frame_out=None) -> None:
        """frame_out: optional CUDA uint8 tensor of iy*ix*4 bytes; the advecting kernel then also stores the frame of the new
        image, (png_byte)(x*255.0) per channel (pfs_ctx_advect_color_step_rgba8)."""
        if frame_out is None:
            check(_cabi.lib().pfs_ctx_advect_color_step(self._h, dt, self._stream()))
        else:
            check(_cabi.lib().pfs_ctx_advect_color_step_rgba8(self._h, dt, frame_out.data_ptr(), self._stream()))


def advect_color_step(image: vp_field, itmp: vp_field, vp: vp_field, dt: float, frame_out=None) -> None:
    """frame_out (device fields only): CUDA uint8 tensor of iy*ix*4 bytes that receives the frame of the new image from the
    same kernel (pfs_advect_color_step_rgba8)."""
    L = _cabi.lib()
    for f, n in ((image, "image"), (itmp, "itmp"), (vp, "vp")):
        _check_buf(f, n)
    if (image.x, image.y) != (itmp.x, itmp.y):
        raise ValueError("image and itmp must have the same shape")
    if not (image.on_device == itmp.on_device == vp.on_device):
        raise ValueError("image, itmp and vp must all be device tensors or all be host arrays")
    if vp.on_device:
        h = _Handles(image, itmp, vp)
        with _dev_guard(vp.data):
            if frame_out is None:
                check(L.pfs_advect_color_step(h.ref(0), h.ref(1), h.ref(2), dt, image.x, image.y, image.z,
                                              vp.x, vp.y, vp.z, _stream_of(vp.data)))
            else:
                check(L.pfs_advect_color_step_rgba8(h.ref(0), h.ref(1), h.ref(2), dt, image.x, image.y, image.z,
                                                    vp.x, vp.y, vp.z, frame_out.data_ptr(), _stream_of(vp.data)))
        h.commit()
    elif frame_out is not None:
        raise ValueError("frame_out needs device fields")
    else:
        si, st, sv = _host_struct(image), _host_struct(itmp), _host_struct(vp)
        pairs = [(image, si), (itmp, st)]
        check(L.pfs_advect_color_step_host(ctypes.byref(si), ctypes.byref(st), ctypes.byref(sv), dt))
        _commit_host(pairs)


def timestep_host(vp: vp_field, vtmp: vp_field, image: vp_field, itmp: vp_field, dt: float, viscosity: float,
                  n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None) -> None:
    """One iteration of the reference driver loop (main.cpp:236-239) on HOST buffers, with the
    uploads, kernels and downloads of the two halves overlapped (pfs_timestep_host)."""
    L = _cabi.lib()
    n_pressure = n_diffuse if n_pressure is None else n_pressure
    fs = (vp, vtmp, image, itmp)
    for f, n in zip(fs, ("vp", "vtmp", "image", "itmp")):
        _check_buf(f, n)
        if f.on_device:
            raise ValueError("timestep_host takes host (numpy) buffers")
    structs = [_host_struct(f) for f in fs]
    check(L.pfs_timestep_host(*(ctypes.byref(s) for s in structs), dt, viscosity, n_diffuse, n_pressure))
    _commit_host(list(zip(fs[:2], structs[:2])))
    _commit_host(list(zip(fs[2:], structs[2:])))


def image_to_rgba8(image: vp_field):
    """Device-side (png_byte)(x * 255.0) of an image (utils.hpp:129-131) -> uint8 CUDA tensor [H, W, 4]."""
    import torch
    L = _cabi.lib()
    _check_buf(image, "image")
    _require_device("image_to_rgba8", image)
    out = torch.empty((image.y, image.x, 4), dtype=torch.uint8, device=image.data.device)
    with _dev_guard(image.data):
        check(L.pfs_image_to_rgba8(image.data.data_ptr(), out.data_ptr(), image.x, image.y, image.z, _stream_of(image.data)))
    return out


def step_norms(vp: vp_field, tmp: vp_field) -> dict:
    """Diagnostics of the step that produced (vp, tmp): L2 norm of the divergence, of the last Jacobi update
    p_N - p_{N-1}, of the projected velocity, and the maximum speed component (pfs_step_norms)."""
    L = _cabi.lib()
    _check_buf(vp, "vp"); _check_buf(tmp, "tmp")
    _require_device("step_norms", vp, tmp)
    out = (ctypes.c_double * 4)()
    with _dev_guard(vp.data):
        check(L.pfs_step_norms(vp.data.data_ptr(), tmp.data.data_ptr(), vp.x, vp.y, vp.z, out, _stream_of(vp.data)))
    return {"div_l2": out[0], "pressure_update_l2": out[1], "velocity_l2": out[2], "speed_max": out[3]}


def _require_device(op: str, *fields: vp_field) -> None:
    for f in fields:
        if not f.on_device:
            raise ValueError(f"{op}: single operators take device tensors; use simulate_fluid_step / "
                             "advect_color_step / timestep_host for host buffers")


# ------------------------------------------------------------------------------------------------
# library controls
# ------------------------------------------------------------------------------------------------

def kernel_launch_count() -> int:
    return int(_cabi.lib().pfs_kernel_launch_count())


def set_fuse_depth(depth: int) -> None:
    check(_cabi.lib().pfs_set_fuse_depth(depth))


def get_fuse_depth() -> int:
    return int(_cabi.lib().pfs_get_fuse_depth())


def phase_timing(enable: bool) -> None:
    check(_cabi.lib().pfs_phase_timing_enable(1 if enable else 0))


def phase_times(reset: bool = True):
    """-> ({phase: ms}, {phase: launches}) accumulated since the last reset."""
    ms = (ctypes.c_float * len(_cabi.PHASES))()
    ln = (ctypes.c_uint64 * len(_cabi.PHASES))()
    check(_cabi.lib().pfs_phase_times(ms, ln, 1 if reset else 0))
    return ({p: float(ms[i]) for i, p in enumerate(_cabi.PHASES)},
            {p: int(ln[i]) for i, p in enumerate(_cabi.PHASES)})


def pinned_empty(shape) -> np.ndarray:
    """float32 numpy array backed by pinned host memory from pfs_host_alloc (cudaMallocHost, as the
    reference's CUDA build does for its PNG buffers, utils.hpp:69-76).  Freed with the array."""
    n = int(np.prod(shape))
    p = ctypes.c_void_p()
    check(_cabi.lib().pfs_host_alloc(ctypes.byref(p), n * 4))
    buf = (ctypes.c_float * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.float32).reshape(shape)
    _PINNED[arr.ctypes.data] = p.value
    return arr


_PINNED: dict[int, int] = {}


def pinned_free(arr: np.ndarray) -> None:
    p = _PINNED.pop(arr.ctypes.data, None)
    if p is not None:
        check(_cabi.lib().pfs_host_free(ctypes.c_void_p(p)))
