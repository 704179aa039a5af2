"""probabilistic_fluid_simulation_b200 -- B200-native (sm_100a) per-timestep fluid update of
mdushkoff/Probabilistic_Fluid_Simulation, behind the reference's own operator interface.

    from probabilistic_fluid_simulation_b200 import vp_field, simulate_fluid_step, advect_color_step

The compute lives in ``lib/libpfs_b200.so`` (hand-written CUDA, C-ABI in ``include/pfs_b200.h``);
this package is the thin host-side mirror of ``includes/fluid.hpp``.  Importing the package does
not load the library; the first operator call does, and raises if it is missing (no CPU fallback).
"""
from .fluid import (NUM_JACOBI_ITERS, FluidContext, add_forces_stochastic, addForces, advect, advect_color, advect_color_step, computePressure, computePressureAdaptive, computePressureSOR,
                    diffuse, get_fuse_depth, image_to_rgba8, kernel_launch_count, phase_times, phase_timing, pinned_empty,
                    pinned_free, set_fuse_depth, simulate_fluid_step, step_norms, subtractPressureGradient, timestep_host,
                    vp_field)
from ._cabi import PfsError, LIB_PATH

__all__ = ["NUM_JACOBI_ITERS", "FluidContext", "vp_field", "advect", "advect_color", "diffuse", "addForces", "computePressure", "computePressureAdaptive", "computePressureSOR",
           "subtractPressureGradient", "add_forces_stochastic", "simulate_fluid_step", "advect_color_step", "timestep_host",
           "step_norms", "image_to_rgba8", "kernel_launch_count", "set_fuse_depth", "get_fuse_depth", "phase_timing", "phase_times",
           "pinned_empty", "pinned_free", "PfsError", "LIB_PATH"]
