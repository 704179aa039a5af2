/*
 * ref_cu_wrap.cu -- builds the UNMODIFIED reference CUDA backend (src/fluid.cu, written for sm_75) for
 * sm_100a into oracle/_ref/libfluid_refcu_<N>.so, as a TIMED BASELINE only.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY.  It is NOT a correctness reference: for even sweep counts fluid.cu
 * deviates from fluid.cpp, and nothing in the product calls it.
 * Same trick as ref_wrap.cpp: pre-define the header guard so `#include "../includes/fluid.hpp"` becomes a
 * no-op, set the sweep count with -DPFS_REF_ITERS, then #include the reference source where it lies.
 */
#define FLUID_HPP_
#ifndef PFS_REF_ITERS
#define PFS_REF_ITERS 30
#endif
#define NUM_JACOBI_ITERS (PFS_REF_ITERS)

#include PFS_REF_FLUID_CU

extern "C" {
int refcu_num_jacobi_iters(void) { return NUM_JACOBI_ITERS; }
/* One timestep of the reference CUDA driver loop (main.cpp:222,225) on device buffers. */
void refcu_timestep(float **vp, float **tmp, float **image, float **itmp, float dt, float viscosity, int vx, int vy,
                    int ix, int iy)
{
    simulate_fluid_step(vp, tmp, dt, viscosity, vx, vy, 4);
    advect_color_step(image, itmp, vp, dt, ix, iy, 4, vx, vy, 4);
}
}
