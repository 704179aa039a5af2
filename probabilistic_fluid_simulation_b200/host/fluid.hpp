/*
 * fluid.hpp -- drop-in replacement for the reference's includes/fluid.hpp (operator API of
 * mdushkoff/Probabilistic_Fluid_Simulation), backed by libpfs_b200.so.
 *
 * Same type, same function names, same argument order and meaning as the reference header:
 *   - without USE_CUDA: the CPU-build signatures taking vp_field* with HOST data pointers
 *     (reference fluid.hpp:32,45,58,70,81,92,109,118).  Each call uploads, runs the CUDA kernels,
 *     downloads, and exchanges the structs' data pointers exactly as src/fluid.cpp does.
 *   - with USE_CUDA: the two host entry points of the CUDA build, simulate_fluid_step /
 *     advect_color_step on caller-owned DEVICE buffers (reference fluid.hpp:107,116) -- the only two
 *     functions the reference driver calls (main.cpp:222,225).  The reference also declares its
 *     seven __global__ kernels in this branch; they are implementation details of its fluid.cu and
 *     are not part of what a driver can link against, so they are not declared here.
 *   Semantics follow the CPU implementation src/fluid.cpp (the reference fluid.cu
 *   deviates from it for even sweep counts).
 *
 * Differences a caller can observe: none in results (bit-identical to fluid.cpp).  Failures
 * (no GPU, CUDA error, bad shape) print pfs_last_error() and abort(); the reference ignores them.
 * NUM_JACOBI_ITERS may be overridden at compile time (-DNUM_JACOBI_ITERS=50): it is passed to the
 * library at run time, whereas the reference bakes it into its loops.
 */
#ifndef FLUID_HPP_
#define FLUID_HPP_

#ifndef NUM_JACOBI_ITERS
#define NUM_JACOBI_ITERS (30)  // reference fluid.hpp:11
#endif

typedef struct {
    int x;
    int y;
    int z;
    float *data;  // interleaved RGBA-order rows, as in the reference (fluid.hpp:17-22)
} vp_field;

#ifdef USE_CUDA
void simulate_fluid_step(float **vp, float **tmp, float dt, float viscosity, int vx, int vy, int vz);
void advect_color_step(float **image, float **itmp, float **vp, float dt, int ix, int iy, int iz, int vx, int vy,
                       int vz);
#else
void advect(vp_field *vp, vp_field *vp_out, float dt);
void advect_color(vp_field *image, vp_field *out, vp_field *vp, float dt);
void diffuse(vp_field *vp, vp_field *vp_out, float viscosity, float dt);
void addForces(vp_field *vp, float *forces);
void computePressure(vp_field *vp, vp_field *vp_out, float dt);
void subtractPressureGradient(vp_field *vp, vp_field *vp_out, float dt);
void simulate_fluid_step(vp_field *vp, vp_field *tmp, float dt, float viscosity);
void advect_color_step(vp_field *image, vp_field *itmp, vp_field *vp, float dt);
#endif  // USE_CUDA

#endif  // FLUID_HPP_
