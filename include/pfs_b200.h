/*
 * pfs_b200.h -- C-ABI of libpfs_b200.so: the B200-native (sm_100a) implementation of the
 * per-timestep fluid update of mdushkoff/Probabilistic_Fluid_Simulation.
 *
 * Scope: the hot path that the reference's src/fluid.cpp runs on the CPU (simulate_fluid_step +
 * advect_color_step and the six operators they call), behind the reference's own operator
 * interface includes/fluid.hpp.  Every entry point below names the reference declaration it
 * replaces.  Plain C types only; no torch, no C++ types.  There is NO CPU fallback: every call
 * either runs hand-written CUDA kernels on the current device or returns an error.
 *
 * Data layout at the boundary (unchanged from the reference, fluid.cpp:15-17, main.cpp:25):
 *   row-major, interleaved, 4 floats per cell:  idx = (j*W + i)*4 + k
 *   velocity/pressure field:  k = 0 u, 1 v, 2 pressure, 3 divergence
 *   image:                    k = R,G,B,A in [0,1]
 * Internally the library works on SoA planes it owns; callers never see them.
 *
 * Sweep counts: the reference hard-codes NUM_JACOBI_ITERS = 30 for both loops (fluid.hpp:11);
 * here they are run-time arguments.  n_diffuse == n_pressure == 30 reproduces the reference.
 * For any pair of counts the buffer-pointer choreography of fluid.cpp (data pointers swapped
 * after every sweep but the last, fluid.cpp:188-194, 260-265) is reproduced literally: which of
 * the two caller buffers ends up holding which iterate, and whether the caller's two pointers
 * end up exchanged, is exactly what the reference's loops would produce.
 *
 * Threading: one host thread and one in-flight step per device at a time (the library keeps per-device
 * scratch planes, a per-device guard-flag buffer and a cache of step graphs; none of it is locked).
 * Streams: every device-pointer entry point takes a `stream` (a cudaStream_t passed as void*;
 * NULL = the legacy default stream, which is what the reference driver uses) and only enqueues
 * work on it -- no hidden synchronisation.  Host-buffer entry points synchronise before return.
 *
 * Errors: every function returns PFS_OK (0) or a pfs_status; pfs_last_error() gives the message
 * (thread-local).  The reference returns void and ignores CUDA errors (main.cpp:45-49); the
 * fluid.hpp-compatible C++ shim (host/fluid_shim.cpp) aborts on a non-zero status instead.
 */
#ifndef PFS_B200_H_
#define PFS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFS_B200_VERSION 200 /* major*100 + minor */

/* Largest supported height of a grid or image (rows are the y dimension of the kernels' launch grids: 65535 blocks of
 * 4 rows).  Wider-than-tall shapes up to the reference's 2^28-cell index limit (fluid.cpp:15-17) are unaffected. */
#define PFS_MAX_ROWS 262140

typedef enum pfs_status {
    PFS_OK = 0,
    PFS_EINVAL = 1,      /* bad argument (null pointer, non-positive size, channels != 4, ...) */
    PFS_ECUDA = 2,       /* a CUDA runtime call or kernel launch failed */
    PFS_ENOMEM = 3,      /* device or pinned-host allocation failed */
    PFS_ENODEVICE = 4,   /* no sm_100-class CUDA device visible: there is no CPU fallback */
    PFS_ESTATE = 5       /* multi-GPU slab context used before its peers were attached, etc. */
} pfs_status;

/* Same fields, same order as the reference's vp_field (includes/fluid.hpp:17-22). */
typedef struct pfs_field {
    int x;        /* width  */
    int y;        /* height */
    int z;        /* channels, must be 4 */
    float *data;
} pfs_field;

/* ---- library ------------------------------------------------------------------------------ */
int pfs_version(void);
const char *pfs_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py gpu_launches). */
uint64_t pfs_kernel_launch_count(void);
/* Frees all per-device scratch, cached step graphs and internal streams.  Optional: the driver releases
 * everything at process exit anyway. */
int pfs_shutdown(void);
/* Tuning knob for tests/benchmarks: maximum number of Jacobi sweeps fused per kernel launch
 * (temporal blocking depth).  0 = library default.  1 = one sweep per launch.  Results are
 * bit-identical for every value. */
int pfs_set_fuse_depth(int max_sweeps_per_launch);
int pfs_get_fuse_depth(void);
/* Diagnostic, host arithmetic only (works without a GPU): how many instructions the fused diffusion passes spend on the
 * division by beta = (float)(1.0 + 4.0*(double)(viscosity*dt)) of fluid.cpp:144-145,182 -- 2 if the two-instruction
 * constant division was verified for this beta (all 2^23 numerator significands tried against the true quotient),
 * else 3 (the FMA-corrected reciprocal multiply, proven for every divisor), 0 if the parameters take the scalar
 * kernels.  Every variant returns the correctly rounded quotient; the choice never changes a result bit. */
int pfs_diffuse_division_ops(float viscosity, float dt);

/* Pinned host memory (the reference's CUDA build reads PNGs into cudaMallocHost memory,
 * includes/utils.hpp:69-76). */
int pfs_host_alloc(void **ptr, size_t bytes);
int pfs_host_free(void *ptr);

/* ---- device-pointer step API: replaces the USE_CUDA branch of includes/fluid.hpp ----------- */

/* Replaces: void simulate_fluid_step(float **vp, float **tmp, float dt, float viscosity,
 *                                    int vx, int vy, int vz)            (fluid.hpp:107)
 * with the semantics of the CPU implementation (fluid.cpp:298-305):
 *   advect(vp->tmp); diffuse(tmp->vp); computePressure(vp->tmp); subtractPressureGradient(tmp->vp)
 * *vp and *tmp are caller-owned DEVICE buffers of vx*vy*vz floats (vz == 4).  On return *vp points
 * at [u_projected, v_projected, p_{N-1}, divergence] and *tmp at [u_diffused, v_diffused, p_N,
 * divergence]; the two pointers are exchanged iff the reference's loops would exchange them
 * (never when n_diffuse and n_pressure have the same parity).  tmp's channel 2 is the warm start
 * of the pressure solve (main.cpp:188-195 initialises it to -1).
 * Launch behaviour: the call only enqueues work on `stream`.  When it is repeated with identical arguments
 * (the reference driver's loop, main.cpp:219-226) the library captures the kernel sequence into a CUDA graph
 * on the second call and replays that graph on `stream` afterwards (PFS_STEP_GRAPH=0 disables this); a
 * caller that is itself capturing `stream` gets plain launches, so the step can be part of a user graph. */
int pfs_simulate_fluid_step(float **vp, float **tmp, float dt, float viscosity,
                            int vx, int vy, int vz, int n_diffuse, int n_pressure, void *stream);

/* Replaces: void advect_color_step(float **image, float **itmp, float **vp, float dt,
 *                                  int ix, int iy, int iz, int vx, int vy, int vz) (fluid.hpp:116)
 * (fluid.cpp:312-320: advect_color(image->itmp) then exchange *image and *itmp). */
int pfs_advect_color_step(float **image, float **itmp, float **vp, float dt,
                          int ix, int iy, int iz, int vx, int vy, int vz, void *stream);
/* pfs_advect_color_step whose kernel also stores the frame bytes of the new image (see pfs_ctx_advect_color_step_rgba8). */
int pfs_advect_color_step_rgba8(float **image, float **itmp, float **vp, float dt, int ix, int iy, int iz,
                                int vx, int vy, int vz, unsigned char *rgba8_out, void *stream);

/* ---- device-pointer operator API: the six operators of includes/fluid.hpp ------------------ */
/* Each one has the reference CPU operator's exact effect on the caller's interleaved buffers
 * (which channels are written, which are left untouched, how the pointers end up). */

/* advect (fluid.hpp:32, fluid.cpp:24-70): vp_out ch0,1 <- semi-Lagrangian advection of vp ch0,1. */
int pfs_advect(const float *vp, float *vp_out, float dt, int vx, int vy, int vz, void *stream);
/* diffuse (fluid.hpp:58, fluid.cpp:129-196): n_sweeps smoothing sweeps on ch0,1, ping-ponging
 * between the two buffers; *vp_out ends on the buffer written last, *vp on the other. */
int pfs_diffuse(float **vp, float **vp_out, float viscosity, float dt,
                int vx, int vy, int vz, int n_sweeps, void *stream);
/* addForces (fluid.hpp:70, fluid.cpp:198-208): the reference body is an empty loop over every cell and channel and its
 * call site is commented out (fluid.cpp:302).  forces == NULL reproduces that (a validated no-op).  With a force field
 * (device buffer shaped like vp) the velocity channels take it: vp[..,0:2] += forces[..,0:2]; see
 * pfs_simulate_fluid_step_forced for the definition and its parity status. */
int pfs_add_forces(float *vp, const float *forces, int vx, int vy, int vz, void *stream);
/* computePressure (fluid.hpp:81, fluid.cpp:210-267): divergence of *vp ch0,1 into ch3 of BOTH
 * buffers, then n_sweeps Jacobi sweeps on ch2 starting from *vp ch2; pointer rule as pfs_diffuse. */
int pfs_compute_pressure(float **vp, float **vp_out, float dt,
                         int vx, int vy, int vz, int n_sweeps, void *stream);
/* subtractPressureGradient (fluid.hpp:92, fluid.cpp:269-296): vp_out ch0,1 <- vp ch0,1 - grad(vp ch2)*dt/2 */
int pfs_subtract_pressure_gradient(const float *vp, float *vp_out, float dt,
                                   int vx, int vy, int vz, void *stream);
/* advect_color (fluid.hpp:45, fluid.cpp:72-127): itmp <- image advected through vp ch0,1. */
int pfs_advect_color(const float *image, float *itmp, const float *vp, float dt,
                     int ix, int iy, int iz, int vx, int vy, int vz, void *stream);

/* ---- external force at the addForces slot --------------------------------------------------- */
/* pfs_simulate_fluid_step with addForces(vp, forces) (fluid.hpp:62-71) applied where fluid.cpp:302 would call it
 * (after diffuse, before computePressure): `forces` is a device buffer shaped like the velocity field (vx*vy*4 floats,
 * "force values for each pixel in the grid"); channels 0,1 are added to (u, v) of the field struct `vp` points at after
 * diffuse, one rounded addition each, channels 2,3 are ignored.  forces == NULL is pfs_simulate_fluid_step.  The
 * reference's addForces body is an empty loop (fluid.cpp:198-208), so reference parity of the force term is UNPINNED BY
 * CONSTRUCTION; oracle/fluid_oracle.c restates this definition and the GPU path is checked against it bit for bit.
 * The addition is fused into the store of the last diffusion pass (no extra trip through memory). */
int pfs_simulate_fluid_step_forced(float **vp, float **tmp, float dt, float viscosity, int vx, int vy, int vz,
                                   int n_diffuse, int n_pressure, const float *forces, void *stream);

/* ---- persistent-state contexts (pfs_ctx_create / upload / step / download / destroy) ---------- */
/* The stateless calls above re-read and re-write the caller's interleaved buffers every step (the caller may have
 * changed them in between).  A context owns the state instead, in the planar layout the kernels use, so a step moves
 * only what the algorithm needs; the interleaved form exists only in upload and download.  What n steps leave behind is
 * bit for bit what n calls of pfs_simulate_fluid_step + pfs_advect_color_step leave in the caller's buffers (same
 * pointer choreography for every sweep-count parity).  The reference driver's loop (main.cpp:219-246) maps onto it as:
 * upload once, pfs_ctx_step per timestep, pfs_ctx_image / pfs_image_to_rgba8 per frame, download at the end.
 * One context per field; a context is bound to the device that was current at creation; calls on one context are
 * serialised internally.  Repeated steps with the same parameters replay CUDA graphs (PFS_STEP_GRAPH=0 disables). */
typedef struct pfs_ctx pfs_ctx;
int pfs_ctx_create(pfs_ctx **out, int vx, int vy, int ix, int iy);      /* ix = iy = 0: no image */
int pfs_ctx_destroy(pfs_ctx *c);
/* Device pointers to interleaved buffers (vx*vy*4 / ix*iy*4 floats); NULL = leave that buffer's state as it is (the
 * first upload needs all of them).  `tmp` is the reference's vtmp: its channel 2 is the first pressure guess. */
int pfs_ctx_upload(pfs_ctx *c, const float *vp, const float *tmp, const float *image, void *stream);
int pfs_ctx_download(pfs_ctx *c, float *vp, float *tmp, float *image, void *stream);   /* NULL = skip */
/* simulate_fluid_step / advect_color_step (fluid.hpp:107,116) on the context's state. */
int pfs_ctx_simulate_fluid_step(pfs_ctx *c, float dt, float viscosity, int n_diffuse, int n_pressure, void *stream);
int pfs_ctx_simulate_fluid_step_forced(pfs_ctx *c, float dt, float viscosity, int n_diffuse, int n_pressure,
                                       const float *forces, void *stream);
int pfs_ctx_simulate_fluid_step_stochastic(pfs_ctx *c, float dt, float viscosity, int n_diffuse, int n_pressure,
                                           float sigma, uint64_t seed, uint32_t step, void *stream);
int pfs_ctx_advect_color_step(pfs_ctx *c, float dt, void *stream);
/* The same step, its kernel also storing the frame of the NEW image as bytes (ix*iy*4, device memory, 4-byte aligned): what
 * write_png_from_array (utils.hpp:129-131) would form, (png_byte)(x*255.0) per channel -- so a frame costs no extra pass
 * over the image and only bytes cross PCIe. */
int pfs_ctx_advect_color_step_rgba8(pfs_ctx *c, float dt, unsigned char *rgba8_out, void *stream);
/* n_steps iterations of the driver loop: simulate_fluid_step + advect_color_step (the latter only with an image). */
int pfs_ctx_step(pfs_ctx *c, int n_steps, float dt, float viscosity, int n_diffuse, int n_pressure, void *stream);
/* Device pointer of the current image (interleaved RGBA floats), e.g. for pfs_image_to_rgba8; valid until the next step. */
int pfs_ctx_image(pfs_ctx *c, const float **image_dev);

/* ---- opt-in stochastic forcing at the addForces slot (NOT in the reference) ------------------- */
/* The reference's addForces is an empty stub whose call is commented out (fluid.cpp:198-208, :302) and the
 * project has no random term anywhere.  These two entry points fill that slot with an
 * additive Gaussian forcing of (u, v): u += sigma*g, g ~ N(0,1) approximated by an Irwin-Hall sum of 8
 * uniform 16-bit integers drawn from Philox-4x32-10 keyed by `seed` with counter (cell, step).  Counter
 * based, so any decomposition draws the same numbers; restated in oracle/fluid_oracle.c and checked bit
 * for bit.  sigma == 0 is exactly the deterministic path. */
int pfs_add_forces_stochastic(float *vp, float sigma, uint64_t seed, uint32_t step, int vx, int vy, int vz,
                              void *stream);
/* pfs_simulate_fluid_step with the forcing applied where fluid.cpp:302 would call addForces (after
 * diffuse, before computePressure). */
int pfs_simulate_fluid_step_stochastic(float **vp, float **tmp, float dt, float viscosity, int vx, int vy, int vz,
                                       int n_diffuse, int n_pressure, float sigma, uint64_t seed, uint32_t step,
                                       void *stream);

/* ---- host-buffer API: replaces the non-USE_CUDA branch of includes/fluid.hpp --------------- */
/* Same semantics as the device-pointer calls, but the pfs_field structs hold HOST pointers, as
 * in the reference CPU build (fluid.hpp:109,118; main.cpp:236,239).  Each call copies its inputs
 * to the device, runs the kernels and copies every buffer the reference would have modified
 * back into the caller's memory, exchanging the structs' data pointers as the reference does.
 * Pinned host memory (pfs_host_alloc) makes the copies asynchronous and ~2x faster. */
int pfs_simulate_fluid_step_host(pfs_field *vp, pfs_field *tmp, float dt, float viscosity,
                                 int n_diffuse, int n_pressure);
int pfs_advect_color_step_host(pfs_field *image, pfs_field *itmp, pfs_field *vp, float dt);
/* One whole timestep of the reference driver loop (main.cpp:236-239) on host buffers:
 * simulate_fluid_step + advect_color_step, with the uploads, kernels and downloads of the two
 * halves overlapped on separate streams. */
int pfs_timestep_host(pfs_field *vp, pfs_field *vtmp, pfs_field *image, pfs_field *itmp,
                      float dt, float viscosity, int n_diffuse, int n_pressure);

/* ---- multi-GPU: one periodic grid as a ring of row slabs (BASELINE configs[3]) ---------------- */
/* Not on the reference's surface (the reference is single-device); same operators, same
 * results bit for bit.  Rank r of R owns a contiguous band of velocity rows and the band of image rows
 * whose velocity look-up (fluid.cpp:89-90) falls into it; the caller's interleaved buffers are split the
 * same way (band rows x width x 4 floats each).  32 halo rows move between ring neighbours about every 30 sweeps
 * (a fused pass also recomputes the rows just outside its band, csrc/slab.cu).  Transports: NCCL, one process per GPU (pfs_slab_connect_nccl), or direct
 * copies between slabs living in one process on any devices (pfs_slab_connect_local). */
typedef struct pfs_slab pfs_slab;
/* Pure host arithmetic: the bands of rank `rank` (works without a GPU). */
int pfs_slab_partition(int rank, int nranks, int gh, int ih, int *row0, int *rows, int *irow0, int *irows);
/* Creates the context of one rank on the CURRENT device: gw x gh velocity grid, iw x ih image (0 x 0: none). */
int pfs_slab_create(pfs_slab **out, int rank, int nranks, int gw, int gh, int iw, int ih);
int pfs_slab_destroy(pfs_slab *s);
int pfs_slab_rows(const pfs_slab *s, int *row0, int *rows, int *irow0, int *irows);
int pfs_slab_connect_local(pfs_slab *const *slabs, int n);         /* all n ranks, in rank order */
int pfs_slab_nccl_unique_id(char id[128]);                          /* rank 0; ship the bytes to every rank */
int pfs_slab_connect_nccl(pfs_slab *s, const char id[128]);         /* collective over all ranks */
/* Which transport carries this slab's halo rows: "p2p" (rows stored straight into the ring neighbours' memory
 * through CUDA IPC mappings, set up by pfs_slab_connect_nccl when every rank can map its neighbours;
 * PFS_SLAB_TRANSPORT=nccl turns it off), "nccl" (send/recv), "local" (slabs of one process), "unconnected". */
const char *pfs_slab_transport(const pfs_slab *s);
/* simulate_fluid_step / advect_color_step (fluid.hpp:107,116) on the bands.  slabs/vp/tmp/streams are
 * arrays over the LOCAL slabs (all ranks for the in-process transport, exactly one under NCCL); vp[k] and
 * tmp[k] are exchanged exactly as pfs_simulate_fluid_step exchanges *vp and *tmp. */
int pfs_slab_simulate_fluid_step(pfs_slab *const *slabs, int n_local, float **vp, float **tmp, float dt,
                                 float viscosity, int n_diffuse, int n_pressure, void *const *streams);
int pfs_slab_advect_color_step(pfs_slab *const *slabs, int n_local, float **image, float **itmp,
                               float *const *vp, float dt, void *const *streams);
/* pfs_simulate_fluid_step_forced on the bands: forces[k] = the k-th local slab's band of the force field. */
int pfs_slab_simulate_fluid_step_forced(pfs_slab *const *slabs, int n_local, float **vp, float **tmp, float dt,
                                        float viscosity, int n_diffuse, int n_pressure, const float *const *forces,
                                        void *const *streams);
/* Resident state on the slabs (the multi-GPU counterpart of pfs_ctx_*): upload the bands once, step, download.  Between
 * steps the bands live in each slab's planes, so a step moves no interleaved data; n steps leave behind, bit for bit, what n
 * calls of pfs_slab_simulate_fluid_step + pfs_slab_advect_color_step leave in the caller's bands.  vp/tmp/image: arrays
 * over the local slabs of device pointers to the bands.  The first upload needs all of them (image may be NULL when the
 * slabs were created without one); later uploads may replace the velocity pair (vp AND tmp) or the image alone (NULL
 * arrays are left as they are); download skips NULL arrays and NULL entries.  The gather depths of both advections are guessed from the previous
 * step and verified by the kernels; a wrong guess is repaired before anything that depends on it is overwritten
 * (pfs_slab_download and pfs_slab_check complete that verification for the last step). */
int pfs_slab_upload(pfs_slab *const *slabs, int n_local, const float *const *vp, const float *const *tmp,
                    const float *const *image, void *const *streams);
int pfs_slab_step(pfs_slab *const *slabs, int n_local, int n_steps, float dt, float viscosity, int n_diffuse,
                  int n_pressure, void *const *streams);
/* The two halves of a resident step on their own (simulate_fluid_step / advect_color_step). */
int pfs_slab_step_fluid(pfs_slab *const *slabs, int n_local, float dt, float viscosity, int n_diffuse, int n_pressure,
                        void *const *streams);
int pfs_slab_step_color(pfs_slab *const *slabs, int n_local, float dt, void *const *streams);
int pfs_slab_download(pfs_slab *const *slabs, int n_local, float *const *vp, float *const *tmp, float *const *image,
                      void *const *streams);
/* Synchronises, completes deferred verification and reports an internal halo overflow (a bug, never expected). */
int pfs_slab_check(pfs_slab *const *slabs, int n_local);

/* ---- frame packing: the float -> byte half of write_png_from_array (includes/utils.hpp:129-131) -------- */
/* out[i] = (png_byte)(image[i] * 255.0) for all ix*iy*iz floats -- the reference's DOUBLE multiply and
 * truncation, done on the device so that a frame costs ix*iy*4 bytes of PCIe traffic instead of 16 bytes per
 * pixel (the reference copies the float image back every frame, main.cpp:231).  `out` is a device buffer. */
int pfs_image_to_rgba8(const float *image, unsigned char *out, int ix, int iy, int iz, void *stream);

/* ---- step diagnostics (not in the reference: it runs a fixed sweep count, no convergence test) ---- */
/* From the two post-state buffers of pfs_simulate_fluid_step (*vp = [u, v, p_{N-1}, div], *tmp = [.., p_N, ..]):
 * out = { ||div||_2, ||p_N - p_{N-1}||_2 (the last Jacobi update, a residual proxy), ||(u,v)||_2, max(|u|,|v|) }.
 * Warp-shuffle + fixed-order block reduction in double; synchronises `stream`.  The slab variant all-reduces
 * over the ring (NCCL) and returns the norms of the whole grid on every rank. */
int pfs_step_norms(const float *vp, const float *tmp, int vx, int vy, int vz, double out[4], void *stream);

/* ---- pressure solve with a run-time sweep count (not in the reference: computePressure always runs
 *      NUM_JACOBI_ITERS sweeps with no residual test, fluid.cpp:239-266) ---- */
/* As pfs_compute_pressure, but the number of sweeps N is chosen while running: sweeps go in batches of
 * `check_every` (>= 2; each batch is fused passes), and after each batch rms(p_N - p_{N-1}) =
 * ||p_N - p_{N-1}||_2 / sqrt(vx*vy) is read back; the solve stops at the first batch where it is <= tol, or at
 * max_sweeps.  The buffers then hold EXACTLY what pfs_compute_pressure(n_sweeps = *sweeps_out) leaves in them
 * (same bits, same pointer exchange), so a run can be replayed with the fixed-count entry point.
 * *update_rms_out = the last rms.  Synchronises `stream` once per batch; cannot be captured into a graph. */
int pfs_compute_pressure_adaptive(float **vp, float **vp_out, float dt, int vx, int vy, int vz, float tol,
                                  int max_sweeps, int check_every, int *sweeps_out, double *update_rms_out,
                                  void *stream);
int pfs_slab_step_norms(pfs_slab *const *slabs, int n_local, float *const *vp, float *const *tmp, double out[4],
                        void *const *streams);
/* The same run-time sweep count on a ring of slabs: vp[k] / vp_out[k] are the bands; the per-rank sums of squared updates are
 * all-reduced (NCCL) after every batch, so every rank takes the same decision; one host synchronisation per batch, none
 * inside it.  The bands end up bit-identical to pfs_compute_pressure with n_sweeps = *sweeps_out on the whole grid. */
int pfs_slab_compute_pressure_adaptive(pfs_slab *const *slabs, int n_local, float **vp, float **vp_out, float dt, float tol,
                                       int max_sweeps, int check_every, int *sweeps_out, double *update_rms_out,
                                       void *const *streams);
/* Red-black successive over-relaxation instead of Jacobi sweeps.  NOT the reference's solver and NOT a parity path: an opt-in
 * alternative whose sweeps-to-tolerance is reported beside the Jacobi solve (profiles/).  Divergence as computePressure forms
 * it (channel 3 of both buffers); the pressure is relaxed in place from channel 2 of vp with factor omega (0 < omega < 2)
 * until the rms update of a full sweep is <= tol (looked at every check_every sweeps) or max_sweeps; result in channel 2
 * of vp_out; pointers are not exchanged.  Single device. */
int pfs_compute_pressure_sor(const float *vp, float *vp_out, float dt, int vx, int vy, int vz, float omega, float tol,
                             int max_sweeps, int check_every, int *sweeps_out, double *update_rms_out, void *stream);

/* ---- phase timing (diagnostics for bench.py; not on the reference's surface) --------------- */
/* When enabled, pfs_simulate_fluid_step / pfs_advect_color_step bracket each phase with CUDA
 * events on the caller's stream.  pfs_phase_times() synchronises those events and returns the
 * accumulated milliseconds and launch counts since the last reset. */
#define PFS_PHASE_ADVECT 0
#define PFS_PHASE_DIFFUSE 1
#define PFS_PHASE_DIVERGENCE 2
#define PFS_PHASE_PRESSURE 3
#define PFS_PHASE_PROJECT 4
#define PFS_PHASE_ADVECT_COLOR 5
#define PFS_NUM_PHASES 6
int pfs_phase_timing_enable(int on);
int pfs_phase_times(float ms_out[PFS_NUM_PHASES], uint64_t launches_out[PFS_NUM_PHASES], int reset);

#ifdef __cplusplus
}
#endif
#endif /* PFS_B200_H_ */
