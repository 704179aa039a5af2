/*
 * ref_wrap.cpp -- builds the UNMODIFIED reference CPU backend (src/fluid.cpp) into a shared
 * object with C entry points, without copying any reference source into this repository.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/fluid_oracle.c header).  Output goes to oracle/_ref/.
 *
 * How: the reference header includes/fluid.hpp hard-codes `#define NUM_JACOBI_ITERS (30)`
 * unguarded (fluid.hpp:11), so -D cannot override it.  We therefore pre-define the header's
 * include guard (FLUID_HPP_, fluid.hpp:7-8), which turns `#include "../includes/fluid.hpp"`
 * (fluid.cpp:10) into a no-op, declare the one type the source needs (vp_field, fluid.hpp:17-22)
 * and the sweep count ourselves, and then #include the reference .cpp where it lies
 * (-DPFS_REF_FLUID_CPP='"/root/reference/src/fluid.cpp"').  Every operator is compiled from the
 * reference's own text; only the iteration count is a build parameter (-DPFS_REF_ITERS=N).
 */
#define FLUID_HPP_
#ifndef PFS_REF_ITERS
#define PFS_REF_ITERS 30
#endif
#define NUM_JACOBI_ITERS (PFS_REF_ITERS)

typedef struct {
    int x;
    int y;
    int z;
    float *data;
} vp_field;

#include PFS_REF_FLUID_CPP

extern "C" {
int ref_num_jacobi_iters(void) { return NUM_JACOBI_ITERS; }
void ref_advect(vp_field *vp, vp_field *out, float dt) { advect(vp, out, dt); }
void ref_advect_color(vp_field *image, vp_field *itmp, vp_field *vp, float dt) { advect_color(image, itmp, vp, dt); }
void ref_diffuse(vp_field *vp, vp_field *out, float viscosity, float dt) { diffuse(vp, out, viscosity, dt); }
void ref_compute_pressure(vp_field *vp, vp_field *out, float dt) { computePressure(vp, out, dt); }
void ref_subtract_pressure_gradient(vp_field *vp, vp_field *out, float dt) { subtractPressureGradient(vp, out, dt); }
void ref_simulate_fluid_step(vp_field *vp, vp_field *tmp, float dt, float viscosity) { simulate_fluid_step(vp, tmp, dt, viscosity); }
void ref_advect_color_step(vp_field *image, vp_field *itmp, vp_field *vp, float dt) { advect_color_step(image, itmp, vp, dt); }
/* main.cpp:219-240 loop without the PNG writes */
void ref_run_steps(vp_field *vp, vp_field *vtmp, vp_field *image, vp_field *itmp, float dt, float viscosity, int n_steps)
{
    for (int s = 0; s < n_steps; s++) {
        simulate_fluid_step(vp, vtmp, dt, viscosity);
        if (image && itmp) advect_color_step(image, itmp, vp, dt);
    }
}
}
