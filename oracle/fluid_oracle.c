/*
 * fluid_oracle.c -- CPU restatement of the reference's per-timestep fluid update.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product library (libpfs_b200.so) never links, loads or calls anything in oracle/.
 *
 * What it restates: /root/reference/src/fluid.cpp (the single-threaded CPU backend behind
 * includes/fluid.hpp), function by function, with the ONE generalisation that the sweep counts
 * are run-time arguments (n_diffuse, n_pressure) instead of the compile-time NUM_JACOBI_ITERS
 * (includes/fluid.hpp:11).  With n_diffuse == n_pressure == NUM_JACOBI_ITERS every function is
 * operation-for-operation the reference's arithmetic, in IEEE binary32, including the buffer
 * pointer choreography (the data pointers of the two structs are swapped between sweeps).
 *
 * Pinning: oracle/Makefile compiles the UNMODIFIED reference source (where it lies under
 * /root/reference) into oracle/_ref/libfluid_ref_<N>.so; tests/test_oracle_vs_ref.py checks this
 * restatement against it bit for bit, and tests/test_oracle_golden.py checks it against the
 * golden hashes of SURVEY.md 4.4 and tests/golden/ (generated from the compiled reference by
 * scripts/make_golden.py).  Parity status: PINNED (bit-exact on every vector).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: the reference
 * results are flag-independent only as long as a*b+c is never fused).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* includes/fluid.hpp:17-22 -- same fields, same order */
typedef struct {
    int x;        /* width  */
    int y;        /* height */
    int z;        /* channels (always 4 in the reference driver, main.cpp:25) */
    float *data;  /* interleaved rows, idx = (j*W + i)*C + k */
} pfs_oracle_field;

/* src/fluid.cpp:15-17 */
static inline int as_idx(int i, int j, int k, int width, int channels)
{
    return (j * width * channels) + (i * channels) + k;
}

/* src/fluid.cpp:19-21 -- left-to-right float adds, one multiply, true division */
static inline float jacobi_update(float xl, float xr, float xt, float xb, float alpha, float beta, float b)
{
    return (xl + xr + xt + xb + alpha * b) / beta;
}

/* src/fluid.cpp:48-49, 105-106 -- double fmod wrap; <math.h> in C++ resolves fmod(float,float)
 * to the float overload, i.e. fmodf, and the "+ extent" add is a binary32 add. */
static inline float wrap_coord(float x, float extent)
{
    return fmodf(fmodf(x, extent) + extent, extent);
}

/* src/fluid.cpp:24-70 */
void oracle_advect(pfs_oracle_field *vp, pfs_oracle_field *vp_out, float dt)
{
    int width = vp->x, height = vp->y, depth = vp->z;
    const float *field = vp->data;
    float *new_field = vp_out->data;
    float fwidth = (float)width, fheight = (float)height;

    for (int j = 0; j < height; j++) {
        for (int i = 0; i < width; i++) {
            float u = field[as_idx(i, j, 0, width, depth)];
            float v = field[as_idx(i, j, 1, width, depth)];
            float x_prev = (float)i - dt * u / fwidth;   /* :39  ((dt*u)/fW) */
            float y_prev = (float)j - dt * v / fheight;  /* :41 */
            x_prev = wrap_coord(x_prev, fwidth);         /* :48 */
            y_prev = wrap_coord(y_prev, fheight);        /* :49 */
            int i0 = (int)x_prev, j0 = (int)y_prev;      /* :55 */
            int i1 = (i0 + 1) % width, j1 = (j0 + 1) % height; /* :57 */
            float sx = x_prev - (float)i0, sy = y_prev - (float)j0; /* :58 */
            for (int k = 0; k < 2; k++) {                /* :61-66 */
                new_field[as_idx(i, j, k, width, depth)] =
                    (1 - sx) * (1 - sy) * field[as_idx(i0, j0, k, width, depth)] +
                    sx * (1 - sy) * field[as_idx(i1, j0, k, width, depth)] +
                    (1 - sx) * sy * field[as_idx(i0, j1, k, width, depth)] +
                    sx * sy * field[as_idx(i1, j1, k, width, depth)];
            }
        }
    }
}

/* src/fluid.cpp:72-127 */
void oracle_advect_color(pfs_oracle_field *image, pfs_oracle_field *itmp, pfs_oracle_field *vp, float dt)
{
    int iwidth = image->x, iheight = image->y, idepth = image->z;
    int vwidth = vp->x, vdepth = vp->z;
    float fiwidth = (float)iwidth, fiheight = (float)iheight;
    float fvwidth = (float)vp->x, fvheight = (float)vp->y;
    float viw = fvwidth / fiwidth;   /* :82 */
    float vih = fvheight / fiheight; /* :83 */
    const float *img = image->data;
    float *out = itmp->data;

    for (int j = 0; j < iheight; j++) {
        for (int i = 0; i < iwidth; i++) {
            int vi = (int)((float)i * viw);  /* :89 float multiply, then truncate */
            int vj = (int)((float)j * vih);  /* :90 */
            float u = vp->data[as_idx(vi, vj, 0, vwidth, vdepth)];
            float v = vp->data[as_idx(vi, vj, 1, vwidth, vdepth)];
            float x_prev = (float)i - (dt / viw) * u / fiwidth;   /* :97 */
            float y_prev = (float)j - (dt / vih) * v / fiheight;  /* :98 */
            x_prev = wrap_coord(x_prev, fiwidth);                 /* :105 */
            y_prev = wrap_coord(y_prev, fiheight);                /* :106 */
            int i0 = (int)x_prev, j0 = (int)y_prev;
            int i1 = (i0 + 1) % iwidth, j1 = (j0 + 1) % iheight;  /* :113 */
            float sx = x_prev - (float)i0, sy = y_prev - (float)j0;
            for (int k = 0; k < idepth; k++) {                     /* :117-124 */
                out[as_idx(i, j, k, iwidth, idepth)] =
                    (1 - sx) * (1 - sy) * img[as_idx(i0, j0, k, iwidth, idepth)] +
                    sx * (1 - sy) * img[as_idx(i1, j0, k, iwidth, idepth)] +
                    (1 - sx) * sy * img[as_idx(i0, j1, k, iwidth, idepth)] +
                    sx * sy * img[as_idx(i1, j1, k, iwidth, idepth)];
            }
        }
    }
}

/* src/fluid.cpp:129-196, sweep count generalised */
void oracle_diffuse(pfs_oracle_field *vp, pfs_oracle_field *vp_out, float viscosity, float dt, int n_sweeps)
{
    float alpha = viscosity * dt;            /* :144 */
    float beta = 1.0 + 4.0 * alpha;          /* :145 double arithmetic, rounded to float */
    int w = vp->x, h = vp->y, c = vp->z;
    float *data = vp->data, *data_out = vp_out->data;

    for (int iter = 0; iter < n_sweeps; iter++) {
        for (int j = 0; j < h; j++) {
            for (int i = 0; i < w; i++) {
                int iminus = ((i - 1) % w + w) % w;
                int iplus = (i + 1) % w;
                int jminus = ((j - 1) % h + h) % h;
                int jplus = (j + 1) % h;
                for (int k = 0; k < 2; k++) {
                    float left = data[as_idx(iminus, j, k, w, c)];
                    float right = data[as_idx(iplus, j, k, w, c)];
                    float top = data[as_idx(i, jminus, k, w, c)];
                    float bottom = data[as_idx(i, jplus, k, w, c)];
                    float u_n = data[as_idx(i, j, k, w, c)];
                    data_out[as_idx(i, j, k, w, c)] =
                        jacobi_update(alpha * left, alpha * right, alpha * top, alpha * bottom, 1.0f, beta, u_n);
                }
            }
        }
        if (iter != (n_sweeps - 1)) {        /* :188-194 pointer swap between sweeps */
            float *tp = vp_out->data;
            vp_out->data = vp->data;
            vp->data = tp;
            data = vp->data;
            data_out = vp_out->data;
        }
    }
}

/* src/fluid.cpp:210-267, sweep count generalised */
void oracle_compute_pressure(pfs_oracle_field *vp, pfs_oracle_field *vp_out, float dt, int n_sweeps)
{
    int w = vp->x, h = vp->y, d = vp->z;
    float *data_in = vp->data;
    float alpha = 1.0f, beta = 4.0f;         /* :215-216 */
    float gamma = -1.0 / dt;                 /* :218 double divide, rounded to float */

    for (int j = 0; j < h; j++) {            /* :221-237 divergence into ch3 of BOTH buffers */
        for (int i = 0; i < w; i++) {
            int iminus = ((i - 1) % w + w) % w;
            int iplus = (i + 1) % w;
            int jminus = ((j - 1) % h + h) % h;
            int jplus = (j + 1) % h;
            float uR = data_in[as_idx(iplus, j, 0, w, d)] - data_in[as_idx(iminus, j, 0, w, d)];
            float vT = data_in[as_idx(i, jplus, 1, w, d)] - data_in[as_idx(i, jminus, 1, w, d)];
            data_in[as_idx(i, j, 3, w, d)] = gamma * (uR + vT);
            vp_out->data[as_idx(i, j, 3, w, d)] = gamma * (uR + vT);
        }
    }

    for (int iter = 0; iter < n_sweeps; iter++) {  /* :239-266 */
        for (int j = 0; j < h; j++) {
            for (int i = 0; i < w; i++) {
                int iminus = ((i - 1) % w + w) % w;
                int iplus = (i + 1) % w;
                int jminus = ((j - 1) % h + h) % h;
                int jplus = (j + 1) % h;
                float pL = data_in[as_idx(iminus, j, 2, w, d)];
                float pR = data_in[as_idx(iplus, j, 2, w, d)];
                float pT = data_in[as_idx(i, jminus, 2, w, d)];
                float pB = data_in[as_idx(i, jplus, 2, w, d)];
                float b = data_in[as_idx(i, j, 3, w, d)];
                vp_out->data[as_idx(i, j, 2, w, d)] = jacobi_update(pL, pR, pT, pB, alpha, beta, b);
            }
        }
        if (iter != (n_sweeps - 1)) {
            float *tp = vp_out->data;
            vp_out->data = vp->data;
            vp->data = tp;
            data_in = vp->data;
        }
    }
}

/* src/fluid.cpp:269-296 */
void oracle_subtract_pressure_gradient(pfs_oracle_field *vp, pfs_oracle_field *vp_out, float dt)
{
    int w = vp->x, h = vp->y, d = vp->z;
    const float *data_in = vp->data;
    for (int j = 0; j < h; j++) {
        for (int i = 0; i < w; i++) {
            int iminus = ((i - 1) % w + w) % w;
            int iplus = (i + 1) % w;
            int jminus = ((j - 1) % h + h) % h;
            int jplus = (j + 1) % h;
            float pL = data_in[as_idx(iminus, j, 2, w, d)];
            float pR = data_in[as_idx(iplus, j, 2, w, d)];
            float pT = data_in[as_idx(i, jminus, 2, w, d)];
            float pB = data_in[as_idx(i, jplus, 2, w, d)];
            float gradX = (pR - pL) * dt / 2.0f;   /* :288 */
            float gradY = (pB - pT) * dt / 2.0f;   /* :289 */
            vp_out->data[as_idx(i, j, 0, w, d)] = data_in[as_idx(i, j, 0, w, d)] - gradX;
            vp_out->data[as_idx(i, j, 1, w, d)] = data_in[as_idx(i, j, 1, w, d)] - gradY;
        }
    }
}

/* src/fluid.cpp:298-305 (addForces is commented out there, :302) */
void oracle_simulate_fluid_step(pfs_oracle_field *vp, pfs_oracle_field *tmp, float dt, float viscosity,
                                int n_diffuse, int n_pressure)
{
    oracle_advect(vp, tmp, dt);
    oracle_diffuse(tmp, vp, viscosity, dt, n_diffuse);
    oracle_compute_pressure(vp, tmp, dt, n_pressure);
    oracle_subtract_pressure_gradient(tmp, vp, dt);
}

/* src/fluid.cpp:312-320 */
void oracle_advect_color_step(pfs_oracle_field *image, pfs_oracle_field *itmp, pfs_oracle_field *vp, float dt)
{
    oracle_advect_color(image, itmp, vp, dt);
    float *tp = image->data;
    image->data = itmp->data;
    itmp->data = tp;
}

/* Driver-side initial conditions, src/main.cpp:170-195: v = v*2.0 - 1.0 in double on all four
 * channels; vtmp = (-1,-1,-1,+1) per cell. */
void oracle_init_velocity_from_unit(float *vp, size_t n_floats)
{
    for (size_t i = 0; i < n_floats; i++) {
        float v = vp[i];
        vp[i] = (v * 2.0) - 2.0 / 2.0;
    }
}

void oracle_init_vtmp(float *vtmp, size_t n_floats)
{
    for (size_t i = 0; i < n_floats; i++) vtmp[i] = ((i % 4) == 3) ? 1.0f : -1.0f;
}

/* includes/utils.hpp:82-84 (byte -> float) and :129-131 (float -> byte, truncating) */
void oracle_bytes_to_unit_float(const uint8_t *bytes, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = (float)bytes[i] / 255.0;
}

void oracle_unit_float_to_bytes(const float *x, uint8_t *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)(x[i] * 255.0);
}

/* 64-bit FNV-1a over the 32-bit words of one channel, row-major cell order. */
uint64_t oracle_channel_hash(const float *data, size_t n_cells, int channels, int k)
{
    uint64_t h = 1469598103934665603ULL;
    for (size_t c = 0; c < n_cells; c++) {
        union { float f; uint32_t u; } cv;
        cv.f = data[c * (size_t)channels + (size_t)k];
        h = (h ^ (uint64_t)cv.u) * 1099511628211ULL;
    }
    return h;
}

/* Whole-loop harness mirroring main.cpp:219-240 (no PNG writes): n_steps of
 * simulate_fluid_step + advect_color_step on caller buffers.  The data pointers inside the four
 * structs are updated exactly as the reference updates them. */
void oracle_run_steps(pfs_oracle_field *vp, pfs_oracle_field *vtmp, pfs_oracle_field *image,
                      pfs_oracle_field *itmp, float dt, float viscosity, int n_diffuse, int n_pressure,
                      int n_steps)
{
    for (int s = 0; s < n_steps; s++) {
        oracle_simulate_fluid_step(vp, vtmp, dt, viscosity, n_diffuse, n_pressure);
        if (image && itmp) oracle_advect_color_step(image, itmp, vp, dt);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Opt-in stochastic forcing at the reference's addForces slot (src/fluid.cpp:198-208, call site
 * commented out at :302).  THE REFERENCE HAS NO STOCHASTIC TERM: this restates the
 * extension implemented by libpfs_b200.so (csrc/kernels_basic.cu: stochastic_force_kernel) so that the
 * CUDA path can be checked bit for bit; "reference parity" for sigma > 0 is UNPINNED by construction.
 * sigma == 0 leaves the deterministic path untouched.
 *
 * Noise: Philox-4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi), counter =
 * (cell_lo, cell_hi, step, draw) with cell = j*W + i (global index) and draw = 0,1 -> 8 words = 16
 * uniform 16-bit integers; component u sums the first 8, v the last 8 (Irwin-Hall, n = 8):
 *   g = (float)(2*S - 8*65535) * norm,  norm = (float)(1/sqrt(4 * 8 * (65536^2 - 1)/12))  -> mean 0, var 1
 *   u += sigma * g        (binary32: one multiply by norm, one by sigma, one add)
 * Everything before the three float operations is integer arithmetic, so CPU and GPU agree exactly.
 * ------------------------------------------------------------------------------------------------ */
static inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

float oracle_stochastic_norm(void)
{
    return (float)(1.0 / sqrt(4.0 * 8.0 * (65536.0 * 65536.0 - 1.0) / 12.0));
}

static inline void gaussian_pair(uint64_t cell, uint64_t seed, uint32_t step, float norm, float *gu, float *gv)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[8];
    for (uint32_t d = 0; d < 2; d++) {
        uint32_t ctr[4] = {(uint32_t)cell, (uint32_t)(cell >> 32), step, d};
        philox4x32_10(ctr, key, w + 4 * d);
    }
    int32_t su = 0, sv = 0;
    for (int k = 0; k < 4; k++) {
        su += (int32_t)(w[k] & 0xffffu) + (int32_t)(w[k] >> 16);
        sv += (int32_t)(w[4 + k] & 0xffffu) + (int32_t)(w[4 + k] >> 16);
    }
    *gu = (float)(2 * su - 8 * 65535) * norm;
    *gv = (float)(2 * sv - 8 * 65535) * norm;
}

/* vp ch0,1 += sigma * N(0,1), independently per cell and component.  row0 = global row of vp's first
 * row, gw = width (cell index = (row0 + j)*gw + i), so a band of a larger grid draws the same numbers. */
void oracle_add_forces_stochastic(pfs_oracle_field *vp, float sigma, uint64_t seed, uint32_t step, int row0)
{
    const float norm = oracle_stochastic_norm();
    for (int j = 0; j < vp->y; j++) {
        for (int i = 0; i < vp->x; i++) {
            float gu, gv;
            gaussian_pair((uint64_t)(row0 + j) * (uint64_t)vp->x + (uint64_t)i, seed, step, norm, &gu, &gv);
            float *c = vp->data + as_idx(i, j, 0, vp->x, vp->z);
            c[0] = c[0] + sigma * gu;
            c[1] = c[1] + sigma * gv;
        }
    }
}

/* simulate_fluid_step with the forcing where fluid.cpp:302 would call addForces: on struct `vp` after
 * diffuse, before computePressure. */
void oracle_simulate_fluid_step_stochastic(pfs_oracle_field *vp, pfs_oracle_field *tmp, float dt, float viscosity,
                                           int n_diffuse, int n_pressure, float sigma, uint64_t seed, uint32_t step)
{
    oracle_advect(vp, tmp, dt);
    oracle_diffuse(tmp, vp, viscosity, dt, n_diffuse);
    if (sigma != 0.0f) oracle_add_forces_stochastic(vp, sigma, seed, step, 0);
    oracle_compute_pressure(vp, tmp, dt, n_pressure);
    oracle_subtract_pressure_gradient(tmp, vp, dt);
}

/* ---------------------------------------------------------------------------------------------
 * External force at the addForces slot.  fluid.cpp:198-208 declares addForces(vp_field *vp, float *forces) with an
 * EMPTY triple loop over (x, y, z) ("TODO: Perform force addition") and fluid.cpp:302 leaves its call commented out,
 * so there is nothing in the reference to restate: PARITY OF THIS TERM IS UNPINNED BY CONSTRUCTION.  What is restated
 * here is the definition the B200 library gives the loop body (include/pfs_b200.h, pfs_add_forces): `forces` is shaped
 * like vp ("force values for each pixel in the grid", fluid.hpp:62-66), the velocity channels take it with one
 * rounded addition each, the pressure and divergence channels are left alone.  forces == NULL is the reference.
 * --------------------------------------------------------------------------------------------- */
void oracle_add_forces(pfs_oracle_field *vp, const float *forces)
{
    if (!forces) return;
    const int w = vp->x, h = vp->y, c = vp->z;
    for (int x = 0; x < w; x++) {
        for (int y = 0; y < h; y++) {
            for (int z = 0; z < c && z < 2; z++) {
                const int idx = as_idx(x, y, z, w, c);
                vp->data[idx] = vp->data[idx] + forces[idx];
            }
        }
    }
}

/* simulate_fluid_step with addForces(vp, forces) where fluid.cpp:302 has it commented out. */
void oracle_simulate_fluid_step_forced(pfs_oracle_field *vp, pfs_oracle_field *tmp, float dt, float viscosity,
                                       int n_diffuse, int n_pressure, const float *forces)
{
    oracle_advect(vp, tmp, dt);
    oracle_diffuse(tmp, vp, viscosity, dt, n_diffuse);
    oracle_add_forces(vp, forces);
    oracle_compute_pressure(vp, tmp, dt, n_pressure);
    oracle_subtract_pressure_gradient(tmp, vp, dt);
}
