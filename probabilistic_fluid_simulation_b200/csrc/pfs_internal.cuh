// pfs_internal.cuh -- shared declarations of libpfs_b200.so (not installed; the public surface
// is include/pfs_b200.h).
//
// Arithmetic contract (DESIGN.md "Parity"): every floating-point operation on the path is written
// with a round-to-nearest intrinsic (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn), which nvcc never
// contracts into FMA, in exactly the association order of the reference expression cited next to
// it.  The translation units are additionally compiled with -fmad=false.  The reference CPU
// build has no FMA, so results are bit-identical, not merely close.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pfs_b200.h"

namespace pfs {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PFS_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return ::pfs::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define PFS_TRY(call)                       \
    do {                                    \
        int s__ = (call);                   \
        if (s__ != PFS_OK) return s__;      \
    } while (0)

// Multiprocessors of the current device (cached per device; 148 on a full B200, fewer on MIG slices and other bins).
// The fused passes size their chunks to one resident wave of warps, so the count has to be the device's own.
int sm_count();

extern unsigned long long g_launches;   // kernels launched by this library (host counter)
extern unsigned long long g_passes;     // same, minus the guard-repair launches (what pfs_phase_times reports)
int check_launch(const char *kernel, const char *file, int line);

#define PFS_LAUNCH(kernel, grid, block, smem, stream, ...)                     \
    do {                                                                       \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
        ++::pfs::g_launches;                                                   \
        ++::pfs::g_passes;                                                     \
        PFS_TRY(::pfs::check_launch(#kernel, __FILE__, __LINE__));             \
    } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched this way may have its CTAs placed on SMs the previous kernel of
// the stream has already vacated, before that kernel has finished everywhere; they block in pdl_wait() until it has
// completed and its writes are visible.  The fused sweep passes are one-wave kernels that follow each other 30+ times a
// step, so the launch latency between two of them is otherwise paid on an idle machine.  Rules kept by every kernel
// launched with PFS_LAUNCH_PDL: pdl_wait() is the first thing EVERY thread does after pdl_launch_dependents() -- before any
// global access (reads of the previous kernel's output, and writes to planes it may still be reading) and before any
// early exit, so that this kernel's completion implies the previous kernel's.  PFS_PDL=0 launches them the ordinary way.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    // not inside a stream capture: graph replays showed no gain from programmatic edges and one slow outlier (profiles/r02_tuning.md)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    bool capturing = cap != cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) {
        (void)cudaGetLastError();
        capturing = true;
    } else
        capturing = cap != cudaStreamCaptureStatusNone;
    cfg.numAttrs = (pdl_enabled() && !capturing) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#define PFS_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                \
    do {                                                                                      \
        (void)::pfs::launch_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
        ++::pfs::g_launches;                                                                  \
        ++::pfs::g_passes;                                                                    \
        PFS_TRY(::pfs::check_launch(#kernel, __FILE__, __LINE__));                            \
    } while (0)

// Brackets one phase of a step with CUDA events on `s` when pfs_phase_timing_enable(1) is set
// (pfs_api.cu); a no-op otherwise.
struct PhaseScope {
    int phase;
    cudaStream_t s;
    cudaEvent_t a = nullptr;
    unsigned long long l0 = 0;
    PhaseScope(int phase, cudaStream_t s);
    ~PhaseScope();
    PhaseScope(const PhaseScope &) = delete;
    PhaseScope &operator=(const PhaseScope &) = delete;
};

// ---------------------------------------------------------------------------------------------
// exact-arithmetic device helpers (each cites the reference expression it reproduces)
// ---------------------------------------------------------------------------------------------

// fluid.cpp:48-49,105-106: fmod(fmod(x, ext) + ext, ext) on floats.  CUDA fmodf is exact (0 ulp) but
// costs tens of instructions, and both calls are trivial for almost every cell:
//   fmod(x, ext) = x            for 0 <= x < ext (also for x = -0, which stays -0),
//   fmod(t, ext) = t            for 0 <= t < ext,
//   fmod(t, ext) = t - ext      for ext <= t < 2*ext, and that subtraction is exact (Sterbenz' lemma),
// so only cells whose departure point left the domain call fmodf.  Same bits in every case.
__device__ __forceinline__ float wrap_coord(float x, float ext)
{
    const float r = (x >= 0.0f && x < ext) ? x : fmodf(x, ext);
    const float t = __fadd_rn(r, ext);
    if (t >= ext && t < __fadd_rn(ext, ext)) return __fsub_rn(t, ext);
    if (t >= 0.0f && t < ext) return t;
    return fmodf(t, ext);
}

// a / ext, correctly rounded, for an extent with correctly rounded reciprocal rext = RN(1/ext) (formed once on the host): the 3-instruction FMA
// division whose exactness is established in sweeps_packed.cu (div_const_fast2) and
// tests/exact_div_check.c; numerators outside its safe range (zero, tiny, huge, non-finite) take __fdiv_rn.
__device__ __forceinline__ float div_extent(float a, float ext, float rext)
{
    const float aa = fabsf(a);
    if (aa >= 0x1p-96f && aa <= 0x1p96f) {
        const float q0 = __fmul_rn(a, rext);
        const float e = __fmaf_rn(-ext, q0, a);
        return __fmaf_rn(e, rext, q0);
    }
    return __fdiv_rn(a, ext);
}

// Back-traced coordinate of fluid.cpp:39-47 / 97-105: wrap(fi - a / ext), a = dt * velocity.  Almost every cell has a
// displacement inside the division's safe range and a position within one period of the grid; for those the whole
// thing is eight straight-line instructions (for |x| < ext, fmodf(x, ext) is x itself, so the first wrap_coord branch
// is the identity, and x + ext >= 0).  Everything else -- tiny, huge or non-finite displacements, positions further out
// -- goes through the general code out of line, so the common path carries no divergence bookkeeping.
static __device__ __noinline__ float backtrace_coord_general(float fi, float a, float ext, float rext)
{
    return wrap_coord(__fsub_rn(fi, div_extent(a, ext, rext)), ext);
}

__device__ __forceinline__ float backtrace_coord(float fi, float a, float ext, float rext)
{
    const float aa = fabsf(a);
    const float q0 = __fmul_rn(a, rext);
    const float e = __fmaf_rn(-ext, q0, a);
    const float q = __fmaf_rn(e, rext, q0);        // a == 0 gives a zero here too (its sign cannot matter: fi >= +0)
    const float x = __fsub_rn(fi, q);
    const float t = __fadd_rn(x, ext);
    const bool ok = aa <= 0x1p96f && (aa >= 0x1p-96f || aa == 0.0f) && fabsf(x) < ext && t < __fadd_rn(ext, ext);
    if (!ok) return backtrace_coord_general(fi, a, ext, rext);
    return t >= ext ? __fsub_rn(t, ext) : t;
}

// fluid.cpp:19-21 with alpha = 1, beta = 4 (fluid.cpp:215-216,255): (((pL+pR)+pT)+pB + 1.0f*b)/4.0f.
// Division by 4 and multiplication by 0.25 round identically (same real value, one rounding).
__device__ __forceinline__ float pressure_update(float pl, float pr, float pt, float pb, float b)
{
    return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(pl, pr), pt), pb), b), 0.25f);
}

// fluid.cpp:175-182: jacobi(alpha*L, alpha*R, alpha*T, alpha*B, 1.0f, beta, u_n)
//   = ((((aL + aR) + aT) + aB) + u_n) / beta, IEEE division.
__device__ __forceinline__ float diffuse_update(float l, float r, float t, float b, float c,
                                                float alpha, float beta)
{
    float s = __fadd_rn(__fmul_rn(alpha, l), __fmul_rn(alpha, r));
    s = __fadd_rn(s, __fmul_rn(alpha, t));
    s = __fadd_rn(s, __fmul_rn(alpha, b));
    s = __fadd_rn(s, c);
    return __fdiv_rn(s, beta);
}

// fluid.cpp:230-235: gamma * ((uR - uL) + (vB - vT))
__device__ __forceinline__ float divergence_value(float ur, float ul, float vb, float vt, float gamma)
{
    return __fmul_rn(gamma, __fadd_rn(__fsub_rn(ur, ul), __fsub_rn(vb, vt)));
}

// fluid.cpp:288-293: u - ((pR - pL) * dt) / 2.0f   (x/2 == x*0.5 exactly)
__device__ __forceinline__ float project_component(float vel, float p_hi, float p_lo, float dt)
{
    return __fsub_rn(vel, __fmul_rn(__fmul_rn(__fsub_rn(p_hi, p_lo), dt), 0.5f));
}

// fluid.cpp:55-66 / :112-124: cell indices and the four bilinear weights of a wrapped departure
// point.  Weights are formed first ((1-sx)*(1-sy) etc.), each then multiplies its sample, and the
// four terms are added left to right.
struct Bilinear {
    int i0, i1, j0, j1;
    float w00, w10, w01, w11;
};

__device__ __forceinline__ Bilinear make_bilinear(float xp, float yp, int w, int h)
{
    Bilinear b;
    b.i0 = (int)xp;
    b.j0 = (int)yp;
    b.i1 = b.i0 + 1;
    if (b.i1 >= w) b.i1 -= w;   // (i0 + 1) % width, i0 in [0, w)
    b.j1 = b.j0 + 1;
    if (b.j1 >= h) b.j1 -= h;
    float sx = __fsub_rn(xp, (float)b.i0), sy = __fsub_rn(yp, (float)b.j0);
    float ox = __fsub_rn(1.0f, sx), oy = __fsub_rn(1.0f, sy);
    b.w00 = __fmul_rn(ox, oy);
    b.w10 = __fmul_rn(sx, oy);
    b.w01 = __fmul_rn(ox, sy);
    b.w11 = __fmul_rn(sx, sy);
    return b;
}

__device__ __forceinline__ float bilerp(const Bilinear &b, float f00, float f10, float f01, float f11)
{
    float s = __fadd_rn(__fmul_rn(b.w00, f00), __fmul_rn(b.w10, f10));
    s = __fadd_rn(s, __fmul_rn(b.w01, f01));
    return __fadd_rn(s, __fmul_rn(b.w11, f11));
}

// ---------------------------------------------------------------------------------------------
// host-side launch wrappers (kernels_basic.cu, sweeps_fused.cu, sweeps_packed.cu)
//
// Internal layout: the velocity lives in "(u,v) planes" -- row-major, (u, v) interleaved, 8 bytes per cell,
// 2*w floats per row -- because the diffusion sweeps advance u and v together in packed-FP32 register pairs; pressure
// and divergence are scalar planes of w floats per row.  The caller-visible buffers stay interleaved [u, v, p, div].
// ---------------------------------------------------------------------------------------------

struct SweepParams {
    int w, h;
    float alpha, beta;   // diffusion only (fluid.cpp:144-145)
    // Row map of the planes.  Single GPU: y_base = 0, wrap = 1 (row j-1 of row 0 is row h-1, by index).
    // Slab of a multi-GPU grid: planes carry halo rows, interior row j lives at plane row y_base + j,
    // rows -1 and h are real halo rows, wrap = 0.
    int y_base = 0;
    int wrap = 1;
};

// External force of the addForces slot (fluid.cpp:198-208, :302): an interleaved [rows][w][4] field like the velocity
// buffers; channels 0,1 are added to (u, v), channels 2,3 are ignored.  Row 0 of the field is row `skip_rows` of the
// plane rows a pass writes (a slab pass also recomputes rows outside its band, which get no force here).
struct ForceField {
    const float *aos;
    int skip_rows, rows;
};

// ---- shared by the stateless entry points (pfs_api.cu), the persistent contexts (pfs_ctx.cu) and the slabs (slab.cu) ----
int check_dims(const char *fn, int x, int y, int z);
int check_ptr(const char *fn, const char *name, const void *p);
int check_sweeps(const char *fn, int n);
SweepParams diffuse_params(int w, int h, float viscosity, float dt);     // alpha, beta as fluid.cpp:144-145 forms them
int fuse_depth();                                                         // pfs_set_fuse_depth (0 = default)
bool phase_timing_on();
// n sweeps from iterate 0 in plane a, ping-ponging a <-> b, iterate n-1 into `extra` when a fused pass can store it;
// *last = plane of iterate n, *prev = plane of iterate n-1 (the one the reference leaves in its other buffer)
int run_diffuse(float *a_uv, float *b_uv, float *extra_uv, const SweepParams &p, int n, float **last, float **prev,
                cudaStream_t s, const ForceField *force);
int run_pressure(float *a, float *b, float *extra, const float *rhs, const SweepParams &p, int n, float **last,
                 float **prev, cudaStream_t s);

// interleaved [u,v,p,div] <-> (u,v) plane + scalar planes.  Null pointers are skipped (channels left alone).
int launch_unpack(const float *aos, float *uv, float *p, float *div, int w, int h, cudaStream_t s);
int launch_pack(float *aos, const float *uv, const float *p, const float *div, int w, int h, cudaStream_t s);

// advect (fluid.cpp:24-70).  src: the field to advect AND to gather from, cells `src_stride` floats apart with (u,v) first
// (4: an interleaved buffer, 2: a (u,v) plane); dst likewise (2: a (u,v) plane, 4: channels 0,1 of an interleaved buffer).
int launch_advect(const float *src, int src_stride, float *dst, int dst_stride, float dt, int w, int h, cudaStream_t s);

// n one-sweep launches, ping-ponging a <-> b; *flips = n.  Reference point and remainder path of the fused passes.
int launch_pressure_basic(float *a, float *b, const float *rhs, const SweepParams &p, int n, int *flips, cudaStream_t s);
int launch_diffuse_basic(float *a_uv, float *b_uv, const SweepParams &p, int n, int *flips, cudaStream_t s);

// Temporally blocked pressure sweeps: up to `depth` sweeps fused per launch, bit-identical results.
// *flips = a<->b hops taken (the result is in b if it is odd).  prev (optional): the pass that reaches sweep n also
// stores iterate n-1 there (the reference keeps it in its other buffer); *prev_written says whether it did (not when the
// last hop is a single plain sweep -- iterate n-1 is then simply the plane that sweep read).
int launch_pressure_fused(float *a, float *b, const float *rhs, const SweepParams &p, int n, int depth, int *flips,
                          cudaStream_t s, float *prev = nullptr, int *prev_written = nullptr);
bool fused_sweeps_supported(int w, int h);
int pick_chunk_rows(int h, int columns_of_items, long long slots, int forced_rows);   // sweeps_packed.cu

// Packed-FP32 (f32x2) temporally blocked diffusion on (u,v) planes (sweeps_packed.cu); same contract.  `force`: added to
// iterate n as the last pass stores it.
bool packed_diffuse_supported(const SweepParams &p);
int launch_diffuse_packed(float *a_uv, float *b_uv, const SweepParams &p, int n, int depth, int *flips, cudaStream_t s,
                          float *prev_uv = nullptr, int *prev_written = nullptr, const ForceField *force = nullptr);
int default_diffuse_depth();              // sweeps fused per launch when the caller does not say (PFS_DIFFUSE_DEPTH)
int packed_division_ops(float beta);

// dst[cell*stride + {0,1}] += force[cell*4 + {0,1}] over `rows` rows of w cells (both pointers at their first row).
int launch_add_forces(float *dst, int dst_stride, const float *force_aos, int w, int rows, cudaStream_t s);

// divergence (fluid.cpp:221-237) of a (u,v) plane into plane div; optionally also extracts channel 2 of an
// interleaved buffer into plane p0 (the pressure warm start) in the same pass.
int launch_divergence(const float *uv, float *div, const float *p0_src_aos, float *p0, float dt, int w, int h,
                      cudaStream_t s, int y_base = 0, int wrap = 1);

// End of simulate_fluid_step: subtract the gradient of p_n from (u, v) (fluid.cpp:269-296) and write
// BOTH interleaved post-state buffers with full-cell stores:
//   out_q = [u - gx, v - gy, p_prev, div]      out_p = [u, v, p_n, div]
int launch_project_pack(const float *uv, const float *p_n, const float *p_prev, const float *div, float *out_q,
                        float *out_p, float dt, int w, int h, cudaStream_t s, int y_base = 0, int wrap = 1);
// The same subtraction for state that stays in planes (pfs_ctx): uv_out <- uv - grad(p_n)*dt/2.  uv_out has no halo rows
// of its own concern: it is indexed like uv.  vmax_out (optional): atomic max of |v| of the result (bounds the next gathers).
int launch_project_uv(const float *uv, const float *p_n, float *uv_out, float dt, int w, int h, cudaStream_t s,
                      int y_base = 0, int wrap = 1, float *vmax_out = nullptr);

// subtractPressureGradient as a stand-alone operator on interleaved buffers (writes ch0,1 of out).
int launch_subtract_gradient_aos(const float *vp_aos, float *out_aos, float dt, int w, int h, cudaStream_t s);

// Interleaved float image -> RGBA8 bytes with the reference writer's conversion (utils.hpp:129-131).
int launch_pack_rgba8(const float *image_aos, unsigned char *out, size_t pixels, cudaStream_t s);

// Step diagnostics from the two post-state buffers (kernels_basic.cu): out4 = {sum div^2, sum (p_N - p_{N-1})^2,
// sum (u^2 + v^2), max(|u|,|v|)} in device memory; `partials` holds 4 doubles per block (<= max_blocks blocks).
int launch_plane_diff_norms(const float *a, const float *b, size_t cells, double *partials, int max_blocks, double *out4,
                            cudaStream_t s);
int launch_step_norms(const float *vp_aos, const float *tmp_aos, size_t cells, double *partials, int max_blocks,
                      double *out4, cudaStream_t s);

// Red-black SOR of the pressure equation, one full sweep in place (not a parity path; kernels_basic.cu).
int sor_partial_blocks(int w, int h);
int launch_sor_sweep(float *p, const float *rhs, int w, int h, float omega, double *partials, double *out4, cudaStream_t s);

// Opt-in Gaussian forcing of (u, v) at the addForces slot (kernels_basic.cu).  stride 2 = a (u,v) plane, 4 = interleaved.
int launch_stochastic_force(float *u, float *v, int stride, float sigma, unsigned long long seed, unsigned step, int w,
                            int h, int row0, int y_base, cudaStream_t s);

// advect_color (fluid.cpp:72-127) on interleaved image buffers; the velocity is point-sampled from cells `vel_stride`
// floats apart (4: interleaved buffer, 2: (u,v) plane).  rgba8 != nullptr: the same kernel also stores the frame bytes of the
// new image (iw*ih*4 bytes, 4-byte aligned), (png_byte)(x*255.0) per channel as includes/utils.hpp:129-131.
int launch_advect_color(const float *image, float *out, const float *vel, int vel_stride, float dt, int iw, int ih, int vw,
                        int vh, cudaStream_t s, unsigned char *rgba8 = nullptr);

}  // namespace pfs
