"""ctypes binding of libpfs_b200.so (include/pfs_b200.h).  Fails loudly if the library is missing:
there is no Python/CPU fallback for any operator."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpfs_b200.so")

PFS_OK = 0
STATUS_NAMES = {0: "PFS_OK", 1: "PFS_EINVAL", 2: "PFS_ECUDA", 3: "PFS_ENOMEM", 4: "PFS_ENODEVICE", 5: "PFS_ESTATE"}
PHASES = ("advect", "diffuse", "divergence", "pressure", "project", "advect_color")


class PfsError(RuntimeError):
    def __init__(self, status: int, message: str):
        self.status = status
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")


class Field(ctypes.Structure):
    """pfs_field == the reference's vp_field (includes/fluid.hpp:17-22)."""
    _fields_ = [("x", ctypes.c_int), ("y", ctypes.c_int), ("z", ctypes.c_int),
                ("data", ctypes.POINTER(ctypes.c_float))]


# name -> (restype, argtypes); must list every symbol declared in include/pfs_b200.h
_f32 = ctypes.c_float
_int = ctypes.c_int
_vp = ctypes.c_void_p
_pp = ctypes.POINTER(ctypes.c_void_p)   # float** passed as void**
_F = ctypes.POINTER(Field)
SIGNATURES = {
    "pfs_version": (_int, []),
    "pfs_last_error": (ctypes.c_char_p, []),
    "pfs_kernel_launch_count": (ctypes.c_uint64, []),
    "pfs_shutdown": (_int, []),
    "pfs_set_fuse_depth": (_int, [_int]),
    "pfs_get_fuse_depth": (_int, []),
    "pfs_diffuse_division_ops": (_int, [_f32, _f32]),
    "pfs_host_alloc": (_int, [_pp, ctypes.c_size_t]),
    "pfs_host_free": (_int, [_vp]),
    "pfs_simulate_fluid_step": (_int, [_pp, _pp, _f32, _f32, _int, _int, _int, _int, _int, _vp]),
    "pfs_advect_color_step": (_int, [_pp, _pp, _pp, _f32, _int, _int, _int, _int, _int, _int, _vp]),
    "pfs_advect_color_step_rgba8": (_int, [_pp, _pp, _pp, _f32, _int, _int, _int, _int, _int, _int, _vp, _vp]),
    "pfs_advect": (_int, [_vp, _vp, _f32, _int, _int, _int, _vp]),
    "pfs_diffuse": (_int, [_pp, _pp, _f32, _f32, _int, _int, _int, _int, _vp]),
    "pfs_add_forces": (_int, [_vp, _vp, _int, _int, _int, _vp]),
    "pfs_compute_pressure": (_int, [_pp, _pp, _f32, _int, _int, _int, _int, _vp]),
    "pfs_subtract_pressure_gradient": (_int, [_vp, _vp, _f32, _int, _int, _int, _vp]),
    "pfs_advect_color": (_int, [_vp, _vp, _vp, _f32, _int, _int, _int, _int, _int, _int, _vp]),
    "pfs_add_forces_stochastic": (_int, [_vp, _f32, ctypes.c_uint64, ctypes.c_uint32, _int, _int, _int, _vp]),
    "pfs_simulate_fluid_step_stochastic": (_int, [_pp, _pp, _f32, _f32, _int, _int, _int, _int, _int, _f32,
                                                   ctypes.c_uint64, ctypes.c_uint32, _vp]),
    "pfs_simulate_fluid_step_forced": (_int, [_pp, _pp, _f32, _f32, _int, _int, _int, _int, _int, _vp, _vp]),
    "pfs_ctx_create": (_int, [_pp, _int, _int, _int, _int]),
    "pfs_ctx_destroy": (_int, [_vp]),
    "pfs_ctx_upload": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "pfs_ctx_download": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "pfs_ctx_simulate_fluid_step": (_int, [_vp, _f32, _f32, _int, _int, _vp]),
    "pfs_ctx_simulate_fluid_step_forced": (_int, [_vp, _f32, _f32, _int, _int, _vp, _vp]),
    "pfs_ctx_simulate_fluid_step_stochastic": (_int, [_vp, _f32, _f32, _int, _int, _f32, ctypes.c_uint64, ctypes.c_uint32, _vp]),
    "pfs_ctx_advect_color_step": (_int, [_vp, _f32, _vp]),
    "pfs_ctx_advect_color_step_rgba8": (_int, [_vp, _f32, _vp, _vp]),
    "pfs_ctx_step": (_int, [_vp, _int, _f32, _f32, _int, _int, _vp]),
    "pfs_ctx_image": (_int, [_vp, _pp]),
    "pfs_simulate_fluid_step_host": (_int, [_F, _F, _f32, _f32, _int, _int]),
    "pfs_advect_color_step_host": (_int, [_F, _F, _F, _f32]),
    "pfs_timestep_host": (_int, [_F, _F, _F, _F, _f32, _f32, _int, _int]),
    "pfs_slab_partition": (_int, [_int, _int, _int, _int] + [ctypes.POINTER(_int)] * 4),
    "pfs_slab_create": (_int, [_pp, _int, _int, _int, _int, _int, _int]),
    "pfs_slab_destroy": (_int, [_vp]),
    "pfs_slab_rows": (_int, [_vp] + [ctypes.POINTER(_int)] * 4),
    "pfs_slab_connect_local": (_int, [_pp, _int]),
    "pfs_slab_nccl_unique_id": (_int, [ctypes.c_char_p]),
    "pfs_slab_connect_nccl": (_int, [_vp, ctypes.c_char_p]),
    "pfs_slab_simulate_fluid_step": (_int, [_pp, _int, _pp, _pp, _f32, _f32, _int, _int, _pp]),
    "pfs_slab_advect_color_step": (_int, [_pp, _int, _pp, _pp, _pp, _f32, _pp]),
    "pfs_slab_simulate_fluid_step_forced": (_int, [_pp, _int, _pp, _pp, _f32, _f32, _int, _int, _pp, _pp]),
    "pfs_slab_upload": (_int, [_pp, _int, _pp, _pp, _pp, _pp]),
    "pfs_slab_step": (_int, [_pp, _int, _int, _f32, _f32, _int, _int, _pp]),
    "pfs_slab_download": (_int, [_pp, _int, _pp, _pp, _pp, _pp]),
    "pfs_slab_step_fluid": (_int, [_pp, _int, _f32, _f32, _int, _int, _pp]),
    "pfs_slab_step_color": (_int, [_pp, _int, _f32, _pp]),
    "pfs_slab_check": (_int, [_pp, _int]),
    "pfs_image_to_rgba8": (_int, [_vp, _vp, _int, _int, _int, _vp]),
    "pfs_step_norms": (_int, [_vp, _vp, _int, _int, _int, ctypes.POINTER(ctypes.c_double), _vp]),
    "pfs_compute_pressure_adaptive": (_int, [_pp, _pp, _f32, _int, _int, _int, _f32, _int, _int, ctypes.POINTER(_int),
                                             ctypes.POINTER(ctypes.c_double), _vp]),
    "pfs_slab_compute_pressure_adaptive": (_int, [_pp, _int, _pp, _pp, _f32, _f32, _int, _int, ctypes.POINTER(_int),
                                                  ctypes.POINTER(ctypes.c_double), _pp]),
    "pfs_compute_pressure_sor": (_int, [_vp, _vp, _f32, _int, _int, _int, _f32, _f32, _int, _int, ctypes.POINTER(_int),
                                        ctypes.POINTER(ctypes.c_double), _vp]),
    "pfs_slab_transport": (ctypes.c_char_p, [_vp]),
    "pfs_slab_step_norms": (_int, [_pp, _int, _pp, _pp, ctypes.POINTER(ctypes.c_double), _pp]),
    "pfs_phase_timing_enable": (_int, [_int]),
    "pfs_phase_times": (_int, [ctypes.POINTER(_f32), ctypes.POINTER(ctypes.c_uint64), _int]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libpfs_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C probabilistic_fluid_simulation_b200/csrc`).  There is no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)      # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(status: int) -> None:
    if status != PFS_OK:
        raise PfsError(status, lib().pfs_last_error().decode(errors="replace"))
