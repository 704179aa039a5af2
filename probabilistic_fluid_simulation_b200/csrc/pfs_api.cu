// pfs_api.cu -- the C-ABI of libpfs_b200.so (include/pfs_b200.h): argument validation, per-device
// planar scratch, the buffer-pointer choreography of the reference (fluid.cpp:188-194, 260-265,
// 298-305) and the sequencing of the kernels in kernels_basic.cu / sweeps_fused.cu.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "pfs_internal.cuh"

namespace pfs {

// ---------------------------------------------------------------------------------------------
// errors, launch accounting
// ---------------------------------------------------------------------------------------------
static thread_local char t_error[512] = "";
unsigned long long g_launches = 0;
unsigned long long g_passes = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return PFS_ENOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice)
        return PFS_ENODEVICE;
    return PFS_ECUDA;
}

int sm_count()
{
    static std::mutex m;
    static std::map<int, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 148;
    }
    std::lock_guard<std::mutex> lock(m);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = 148;
    }
    cache[dev] = n;
    return n;
}

bool pdl_enabled()
{
    static const bool on = !(getenv("PFS_PDL") && getenv("PFS_PDL")[0] == '0');
    return on;
}

int check_launch(const char *kernel, const char *file, int line)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return PFS_OK;
    set_error("launch of %s failed at %s:%d: %s", kernel, file, line, cudaGetErrorString(e));
    if (e == cudaErrorNoKernelImageForDevice || e == cudaErrorNoDevice) return PFS_ENODEVICE;
    return PFS_ECUDA;
}

// ---------------------------------------------------------------------------------------------
// per-device scratch: seven planes (u,v x2, p x2, divergence) sized for the largest grid seen,
// plus staging buffers of the host API.
// ---------------------------------------------------------------------------------------------
// Scratch of the stateless entry points, in units of plane_cells floats: three (u,v) planes of two units each
// (ping, pong, diffusion iterate n-1), three pressure planes (ping, pong, iterate n-1), the divergence.
// (The iterates n-1 are what the reference leaves in its other buffer.)
constexpr size_t N_PLANES = 10;

struct DeviceScratch {
    size_t plane_cells = 0;
    float *planes = nullptr;          // N_PLANES * plane_cells floats, one allocation
    float *stage[4] = {nullptr, nullptr, nullptr, nullptr};   // host-API device copies: vp, tmp, image, itmp
    size_t stage_floats[4] = {0, 0, 0, 0};
    cudaStream_t streams[2] = {nullptr, nullptr};             // host-API streams
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double *norm_dev = nullptr;       // reduction partials + 4 results (pfs_compute_pressure_adaptive)
    double *norm_host = nullptr;      // pinned, 4 doubles
    float *uv(int k) const { return planes + (size_t)(2 * k) * plane_cells; }      // k = 0..2
    float *p(int k) const { return planes + (size_t)(6 + k) * plane_cells; }       // k = 0..2
    float *div() const { return planes + (size_t)9 * plane_cells; }
};

static std::mutex g_mutex;
static std::map<int, DeviceScratch> g_scratch;
static int g_fuse_depth = 0;     // 0 = default

static void free_scratch(DeviceScratch &sc)
{
    if (sc.planes) cudaFree(sc.planes);
    for (int i = 0; i < 4; i++)
        if (sc.stage[i]) cudaFree(sc.stage[i]);
    for (int i = 0; i < 2; i++)
        if (sc.streams[i]) cudaStreamDestroy(sc.streams[i]);
    for (int i = 0; i < 4; i++)
        if (sc.ev[i]) cudaEventDestroy(sc.ev[i]);
    if (sc.norm_dev) cudaFree(sc.norm_dev);
    if (sc.norm_host) cudaFreeHost(sc.norm_host);
    sc = DeviceScratch();
}

static int current_device(int *dev)
{
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) {
        set_error("no usable CUDA device (%s); libpfs_b200 has no CPU fallback", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return PFS_ENODEVICE;
    }
    return PFS_OK;
}

// Planes are padded so every plane starts 256-byte aligned whatever the cell count.
static size_t padded_cells(size_t cells) { return (cells + 63) & ~(size_t)63; }

static int get_scratch(size_t cells, DeviceScratch **out)
{
    int dev;
    PFS_TRY(current_device(&dev));
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceScratch &sc = g_scratch[dev];
    size_t need = padded_cells(cells);
    if (sc.plane_cells < need) {
        if (sc.planes) {
            PFS_CUDA(cudaDeviceSynchronize());   // earlier steps may still be using the old planes
            PFS_CUDA(cudaFree(sc.planes));
            sc.planes = nullptr;
            sc.plane_cells = 0;
        }
        PFS_CUDA(cudaMalloc((void **)&sc.planes, N_PLANES * need * sizeof(float)));
        sc.plane_cells = need;
    }
    *out = &sc;
    return PFS_OK;
}

// ---------------------------------------------------------------------------------------------
// phase timing
// ---------------------------------------------------------------------------------------------
struct PhaseSpan {
    int phase;
    cudaEvent_t a, b;
    unsigned long long launches;
};
static bool g_phase_timing = false;
bool phase_timing_on() { return g_phase_timing; }
static std::mutex g_phase_mutex;                 // guards g_spans and g_event_pool (worker threads may step too)
static std::vector<PhaseSpan> g_spans;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t take_event()                  // g_phase_mutex held
{
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

PhaseScope::PhaseScope(int phase_, cudaStream_t s_) : phase(phase_), s(s_)
{
    if (!g_phase_timing) return;
    {
        std::lock_guard<std::mutex> lock(g_phase_mutex);
        a = take_event();
    }
    l0 = g_passes;
    cudaEventRecord(a, s);
}

PhaseScope::~PhaseScope()
{
    if (!a) return;
    std::lock_guard<std::mutex> lock(g_phase_mutex);
    cudaEvent_t b = take_event();
    cudaEventRecord(b, s);
    g_spans.push_back({phase, a, b, g_passes - l0});
}

// ---------------------------------------------------------------------------------------------
// argument checks
// ---------------------------------------------------------------------------------------------
int check_dims(const char *fn, int x, int y, int z)
{
    if (x <= 0 || y <= 0) {
        set_error("%s: width and height must be positive (got %d x %d)", fn, x, y);
        return PFS_EINVAL;
    }
    if (z != 4) {
        set_error("%s: channel count must be 4 (interleaved RGBA / u,v,p,div as in main.cpp:25), got %d", fn, z);
        return PFS_EINVAL;
    }
    if ((size_t)x * (size_t)y > ((size_t)1 << 28)) {
        // the reference's int indexing (fluid.cpp:15-17) tops out at 2^30 floats = 2^28 cells
        set_error("%s: %d x %d exceeds 2^28 cells (the reference's int32 index limit)", fn, x, y);
        return PFS_EINVAL;
    }
    if (y > PFS_MAX_ROWS) {
        // rows are the y dimension of the launch grids (4 rows per block, 65535 blocks)
        set_error("%s: height %d exceeds the supported maximum of %d rows", fn, y, PFS_MAX_ROWS);
        return PFS_EINVAL;
    }
    return PFS_OK;
}

int check_ptr(const char *fn, const char *name, const void *p)
{
    if (p == nullptr) {
        set_error("%s: %s is null", fn, name);
        return PFS_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) {
        set_error("%s: %s must be 16-byte aligned (one interleaved cell)", fn, name);
        return PFS_EINVAL;
    }
    return PFS_OK;
}

int check_sweeps(const char *fn, int n)
{
    if (n < 1) {
        set_error("%s: sweep count must be >= 1 (got %d)", fn, n);
        return PFS_EINVAL;
    }
    return PFS_OK;
}

// ---------------------------------------------------------------------------------------------
// n sweeps on planes with the reference's "n-1 swaps" semantics.
//
// Start: iterate 0 in plane a.  The reference loop writes sweep k to the "other" buffer and swaps, so the caller
// needs BOTH iterate n (the result) and iterate n-1 (left behind in the other buffer, fluid.cpp:188-194, 260-265).
// The fused pass that reaches sweep n also stores iterate n-1 (into `extra`), so no separate last sweep -- and no
// extra trip through HBM -- is needed to have both.  On return *last is the plane of iterate n, *prev that of n-1.
// ---------------------------------------------------------------------------------------------
int fuse_depth() { return g_fuse_depth; }

int run_diffuse(float *a, float *b, float *extra, const SweepParams &p, int n, float **last, float **prev, cudaStream_t s,
                const ForceField *force)
{
    int flips = 0, prev_written = 0;
    if (g_fuse_depth != 1 && packed_diffuse_supported(p)) {
        PFS_TRY(launch_diffuse_packed(a, b, p, n, g_fuse_depth, &flips, s, extra, &prev_written, force));
    } else {
        PFS_TRY(launch_diffuse_basic(a, b, p, n, &flips, s));
        if (force)
            PFS_TRY(launch_add_forces(((flips & 1) ? b : a) + (size_t)(p.y_base + force->skip_rows) * 2 * p.w, 2, force->aos,
                                      p.w, force->rows, s));
    }
    *last = (flips & 1) ? b : a;
    *prev = prev_written ? extra : ((flips & 1) ? a : b);     // else: the plane the last single sweep read
    return PFS_OK;
}

int run_pressure(float *a, float *b, float *extra, const float *rhs, const SweepParams &p, int n, float **last,
                 float **prev, cudaStream_t s)
{
    int flips = 0, prev_written = 0;
    if (g_fuse_depth != 1 && fused_sweeps_supported(p.w, p.h))
        PFS_TRY(launch_pressure_fused(a, b, rhs, p, n, g_fuse_depth, &flips, s, extra, &prev_written));
    else
        PFS_TRY(launch_pressure_basic(a, b, rhs, p, n, &flips, s));
    *last = (flips & 1) ? b : a;
    *prev = prev_written ? extra : ((flips & 1) ? a : b);
    return PFS_OK;
}

SweepParams diffuse_params(int w, int h, float viscosity, float dt)
{
    SweepParams p;
    p.w = w;
    p.h = h;
    p.alpha = viscosity * dt;                               // fluid.cpp:144 (binary32 product)
    p.beta = (float)(1.0 + 4.0 * (double)p.alpha);          // fluid.cpp:145 (double, rounded to float)
    return p;
}

}  // namespace pfs

using namespace pfs;

static std::mutex g_graph_mutex;                        // guards the step-graph cache, g_prev_key and the capture streams: a second host
                                                        // thread stepping another field must not see a half-built entry
static void drop_step_graphs();                         // step-graph cache, defined with the step API below
static void destroy_capture_streams();

// =============================================================================================
// library
// =============================================================================================
extern "C" int pfs_version(void) { return PFS_B200_VERSION; }

extern "C" const char *pfs_last_error(void) { return t_error; }

extern "C" uint64_t pfs_kernel_launch_count(void) { return g_launches; }

extern "C" int pfs_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    {
        std::lock_guard<std::mutex> glock(g_graph_mutex);
        drop_step_graphs();
        destroy_capture_streams();
    }
    int cur = 0;
    bool have = (cudaGetDevice(&cur) == cudaSuccess);
    for (auto &kv : g_scratch) {
        if (cudaSetDevice(kv.first) == cudaSuccess) {
            cudaDeviceSynchronize();
            free_scratch(kv.second);
        }
    }
    g_scratch.clear();
    std::lock_guard<std::mutex> plock(g_phase_mutex);
    for (auto &sp : g_spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    g_spans.clear();
    for (auto e : g_event_pool) cudaEventDestroy(e);
    g_event_pool.clear();
    if (have) cudaSetDevice(cur);
    (void)cudaGetLastError();
    return PFS_OK;
}

extern "C" int pfs_set_fuse_depth(int d)
{
    if (d < 0) {
        set_error("pfs_set_fuse_depth: depth must be >= 0");
        return PFS_EINVAL;
    }
    g_fuse_depth = d;
    return PFS_OK;
}

extern "C" int pfs_get_fuse_depth(void) { return g_fuse_depth; }

extern "C" int pfs_diffuse_division_ops(float viscosity, float dt)
{
    SweepParams p = diffuse_params(4, 1, viscosity, dt);
    if (!packed_diffuse_supported(p)) return 0;
    return packed_division_ops(p.beta);
}

extern "C" int pfs_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) {
        set_error("pfs_host_alloc: ptr is null");
        return PFS_EINVAL;
    }
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__);
    return PFS_OK;
}

extern "C" int pfs_host_free(void *ptr)
{
    if (ptr) PFS_CUDA(cudaFreeHost(ptr));
    return PFS_OK;
}

extern "C" int pfs_phase_timing_enable(int on)
{
    g_phase_timing = (on != 0);
    return PFS_OK;
}

extern "C" int pfs_phase_times(float ms_out[PFS_NUM_PHASES], uint64_t launches_out[PFS_NUM_PHASES], int reset)
{
    for (int i = 0; i < PFS_NUM_PHASES; i++) {
        if (ms_out) ms_out[i] = 0.f;
        if (launches_out) launches_out[i] = 0;
    }
    std::lock_guard<std::mutex> lock(g_phase_mutex);
    for (auto &sp : g_spans) {
        PFS_CUDA(cudaEventSynchronize(sp.b));
        float ms = 0.f;
        PFS_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
        if (ms_out) ms_out[sp.phase] += ms;
        if (launches_out) launches_out[sp.phase] += sp.launches;
    }
    if (reset) {
        for (auto &sp : g_spans) {
            g_event_pool.push_back(sp.a);
            g_event_pool.push_back(sp.b);
        }
        g_spans.clear();
    }
    return PFS_OK;
}

// =============================================================================================
// device-pointer operator API
// =============================================================================================
extern "C" int pfs_advect(const float *vp, float *vp_out, float dt, int vx, int vy, int vz, void *stream)
{
    PFS_TRY(check_dims("pfs_advect", vx, vy, vz));
    PFS_TRY(check_ptr("pfs_advect", "vp", vp));
    PFS_TRY(check_ptr("pfs_advect", "vp_out", vp_out));
    return launch_advect(vp, 4, vp_out, 4, dt, vx, vy, (cudaStream_t)stream);
}

extern "C" int pfs_add_forces(float *vp, const float *forces, int vx, int vy, int vz, void *stream)
{
    // fluid.cpp:198-208: a loop over every cell and channel with an EMPTY body ("TODO: Perform force addition"); the
    // call site is commented out (fluid.cpp:302).  forces == NULL keeps exactly that: nothing happens.  With a force
    // field the loop gets the body its signature and comment announce (fluid.hpp:62-71, "force values for each pixel in
    // the grid"): the velocity channels take the force, vp[..,0:2] += forces[..,0:2], one rounded addition each; the
    // pressure and divergence channels are left alone.  Reference parity of that body is unpinned by construction.
    const char *fn = "pfs_add_forces";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_ptr(fn, "vp", vp));
    if (forces == nullptr) return PFS_OK;
    PFS_TRY(check_ptr(fn, "forces", forces));
    return launch_add_forces(vp, 4, forces, vx, vy, (cudaStream_t)stream);
}

extern "C" int pfs_diffuse(float **vp, float **vp_out, float viscosity, float dt, int vx, int vy, int vz, int n_sweeps,
                           void *stream)
{
    const char *fn = "pfs_diffuse";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_sweeps(fn, n_sweeps));
    if (!vp || !vp_out) {
        set_error("%s: vp / vp_out handle is null", fn);
        return PFS_EINVAL;
    }
    PFS_TRY(check_ptr(fn, "*vp", *vp));
    PFS_TRY(check_ptr(fn, "*vp_out", *vp_out));
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScratch *sc;
    PFS_TRY(get_scratch((size_t)vx * vy, &sc));
    float *in0 = *vp, *out0 = *vp_out;
    float *last = nullptr, *prev = nullptr;
    PFS_TRY(launch_unpack(in0, sc->uv(0), nullptr, nullptr, vx, vy, s));
    PFS_TRY(run_diffuse(sc->uv(0), sc->uv(1), sc->uv(2), diffuse_params(vx, vy, viscosity, dt), n_sweeps, &last, &prev, s,
                        nullptr));
    // Sweep k writes the original vp_out buffer when k is odd and the original vp buffer when k is
    // even (fluid.cpp:188-194).  Only channels 0,1 are ever written.
    float *buf_last = (n_sweeps & 1) ? out0 : in0;
    float *buf_prev = (n_sweeps & 1) ? in0 : out0;
    PFS_TRY(launch_pack(buf_last, last, nullptr, nullptr, vx, vy, s));
    if (n_sweeps >= 2) PFS_TRY(launch_pack(buf_prev, prev, nullptr, nullptr, vx, vy, s));
    *vp_out = buf_last;
    *vp = buf_prev;
    return PFS_OK;
}

extern "C" int pfs_compute_pressure(float **vp, float **vp_out, float dt, int vx, int vy, int vz, int n_sweeps,
                                    void *stream)
{
    const char *fn = "pfs_compute_pressure";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_sweeps(fn, n_sweeps));
    if (!vp || !vp_out) {
        set_error("%s: vp / vp_out handle is null", fn);
        return PFS_EINVAL;
    }
    PFS_TRY(check_ptr(fn, "*vp", *vp));
    PFS_TRY(check_ptr(fn, "*vp_out", *vp_out));
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScratch *sc;
    PFS_TRY(get_scratch((size_t)vx * vy, &sc));
    float *in0 = *vp, *out0 = *vp_out;
    float *uv = sc->uv(0), *div = sc->div();
    float *last = nullptr, *prev = nullptr;
    PFS_TRY(launch_unpack(in0, uv, nullptr, nullptr, vx, vy, s));
    PFS_TRY(launch_divergence(uv, div, in0, sc->p(0), dt, vx, vy, s));
    SweepParams p{vx, vy, 1.0f, 4.0f};
    PFS_TRY(run_pressure(sc->p(0), sc->p(1), sc->p(2), div, p, n_sweeps, &last, &prev, s));
    float *buf_last = (n_sweeps & 1) ? out0 : in0;
    float *buf_prev = (n_sweeps & 1) ? in0 : out0;
    // channel 3 of both buffers <- divergence (fluid.cpp:235-236); channel 2 <- the iterate each
    // buffer was last written with.  With one sweep the input buffer keeps its pressure.
    PFS_TRY(launch_pack(buf_last, nullptr, last, div, vx, vy, s));
    PFS_TRY(launch_pack(buf_prev, nullptr, (n_sweeps >= 2) ? prev : nullptr, div, vx, vy, s));
    *vp_out = buf_last;
    *vp = buf_prev;
    return PFS_OK;
}

// Run-time choice of the sweep count (the reference fixes it, fluid.cpp:239).  Batches of
// `check_every` sweeps; after each, rms(p_N - p_{N-1}) from the two iterates the fused pass leaves behind.
extern "C" int pfs_compute_pressure_adaptive(float **vp, float **vp_out, float dt, int vx, int vy, int vz, float tol,
                                             int max_sweeps, int check_every, int *sweeps_out, double *update_rms_out,
                                             void *stream)
{
    const char *fn = "pfs_compute_pressure_adaptive";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_sweeps(fn, max_sweeps));
    if (!vp || !vp_out) {
        set_error("%s: vp / vp_out handle is null", fn);
        return PFS_EINVAL;
    }
    if (check_every < 2 || !(tol >= 0.0f)) {
        set_error("%s: check_every must be >= 2 and tol >= 0 (got %d, %g)", fn, check_every, (double)tol);
        return PFS_EINVAL;
    }
    PFS_TRY(check_ptr(fn, "*vp", *vp));
    PFS_TRY(check_ptr(fn, "*vp_out", *vp_out));
    cudaStream_t s = (cudaStream_t)stream;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    PFS_CUDA(cudaStreamIsCapturing(s, &cap));
    if (cap != cudaStreamCaptureStatusNone) {
        set_error("%s: reads a norm back after every batch, cannot be captured into a graph", fn);
        return PFS_EINVAL;
    }
    DeviceScratch *sc;
    const size_t cells = (size_t)vx * vy;
    PFS_TRY(get_scratch(cells, &sc));
    float *in0 = *vp, *out0 = *vp_out;
    float *uv = sc->uv(0), *div = sc->div();
    float *cur = sc->p(0), *oth = sc->p(1), *last = cur, *prev = oth;
    float *const extra = sc->p(2);
    PFS_TRY(launch_unpack(in0, uv, nullptr, nullptr, vx, vy, s));
    PFS_TRY(launch_divergence(uv, div, in0, cur, dt, vx, vy, s));
    SweepParams p{vx, vy, 1.0f, 4.0f};
    constexpr int kBlocks = 1184;
    if (!sc->norm_dev) PFS_CUDA(cudaMalloc((void **)&sc->norm_dev, (4 * (size_t)kBlocks + 4) * sizeof(double)));
    if (!sc->norm_host) PFS_CUDA(cudaMallocHost((void **)&sc->norm_host, 4 * sizeof(double)));
    double *scratch = sc->norm_dev, *host = sc->norm_host;
    int done = 0, rc = PFS_OK;
    double rms = 0.0;
    while (done < max_sweeps) {
        int n = std::min(check_every, max_sweeps - done);
        if (max_sweeps - done - n == 1) n += 1;             // never leave a batch of one sweep (it keeps no p_{N-1})
        rc = run_pressure(cur, oth, extra, div, p, n, &last, &prev, s);
        if (rc != PFS_OK) break;
        done += n;
        rc = launch_plane_diff_norms(last, prev, cells, scratch, kBlocks, scratch + 4 * (size_t)kBlocks, s);
        if (rc != PFS_OK) break;
        cudaError_t e = cudaMemcpyAsync(host, scratch + 4 * (size_t)kBlocks, 4 * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, fn, __FILE__, __LINE__);
            break;
        }
        rms = sqrt(host[0] / (double)cells);
        if (rms <= (double)tol) break;
        if (done < max_sweeps) {
            // next batch starts from iterate `done`; the other plane of the ping-pong pair is free again
            // (p_{done-1} is only needed if this was the final batch)
            oth = (last == sc->p(0)) ? sc->p(1) : sc->p(0);
            cur = last;
        }
    }
    if (rc != PFS_OK) return rc;
    float *buf_last = (done & 1) ? out0 : in0;
    float *buf_prev = (done & 1) ? in0 : out0;
    PFS_TRY(launch_pack(buf_last, nullptr, last, div, vx, vy, s));
    PFS_TRY(launch_pack(buf_prev, nullptr, prev, div, vx, vy, s));
    *vp_out = buf_last;
    *vp = buf_prev;
    if (sweeps_out) *sweeps_out = done;
    if (update_rms_out) *update_rms_out = rms;
    return PFS_OK;
}

// Red-black SOR variant of computePressure: NOT the reference's solver and not a parity path -- reported
// beside the Jacobi solve as sweeps-to-tolerance.  Divergence as the reference forms it into channel 3 of both buffers; the
// pressure is relaxed in place starting from channel 2 of vp until the rms update of a full sweep is <= tol (checked every
// `check_every` sweeps) or max_sweeps is reached; the result goes to channel 2 of vp_out.  No pointer exchange.
extern "C" int pfs_compute_pressure_sor(const float *vp, float *vp_out, float dt, int vx, int vy, int vz, float omega, float tol,
                                        int max_sweeps, int check_every, int *sweeps_out, double *update_rms_out, void *stream)
{
    const char *fn = "pfs_compute_pressure_sor";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_sweeps(fn, max_sweeps));
    PFS_TRY(check_ptr(fn, "vp", vp));
    PFS_TRY(check_ptr(fn, "vp_out", vp_out));
    if (!(omega > 0.0f && omega < 2.0f) || check_every < 1 || !(tol >= 0.0f)) {
        set_error("%s: need 0 < omega < 2, check_every >= 1, tol >= 0 (got %g, %d, %g)", fn, (double)omega, check_every, (double)tol);
        return PFS_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScratch *sc;
    const size_t cells = (size_t)vx * vy;
    PFS_TRY(get_scratch(cells, &sc));
    float *uv = sc->uv(0), *div = sc->div(), *p = sc->p(0);
    PFS_TRY(launch_unpack(vp, uv, p, nullptr, vx, vy, s));
    PFS_TRY(launch_divergence(uv, div, nullptr, nullptr, dt, vx, vy, s));
    const int nb = sor_partial_blocks(vx, vy);
    double *scratch = nullptr;
    PFS_CUDA(cudaMalloc((void **)&scratch, (4 * (size_t)nb + 4) * sizeof(double)));
    int done = 0, rc = PFS_OK;
    double rms = 0.0;
    while (done < max_sweeps && rc == PFS_OK) {
        const int batch = std::min(check_every, max_sweeps - done);
        for (int k = 0; k < batch && rc == PFS_OK; k++)
            rc = launch_sor_sweep(p, div, vx, vy, omega, scratch, (k == batch - 1) ? scratch + 4 * (size_t)nb : nullptr, s);
        if (rc != PFS_OK) break;
        done += batch;
        double host = 0.0;
        cudaError_t e = cudaMemcpyAsync(&host, scratch + 4 * (size_t)nb, sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, fn, __FILE__, __LINE__);
            break;
        }
        rms = sqrt(host / (double)cells);
        if (rms <= (double)tol) break;
    }
    cudaFree(scratch);
    if (rc != PFS_OK) return rc;
    PFS_TRY(launch_pack(vp_out, nullptr, p, div, vx, vy, s));
    PFS_TRY(launch_pack(const_cast<float *>(vp), nullptr, nullptr, div, vx, vy, s));     // channel 3 of the input buffer too (fluid.cpp:235-236)
    if (sweeps_out) *sweeps_out = done;
    if (update_rms_out) *update_rms_out = rms;
    return PFS_OK;
}

extern "C" int pfs_subtract_pressure_gradient(const float *vp, float *vp_out, float dt, int vx, int vy, int vz,
                                              void *stream)
{
    const char *fn = "pfs_subtract_pressure_gradient";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_ptr(fn, "vp", vp));
    PFS_TRY(check_ptr(fn, "vp_out", vp_out));
    return launch_subtract_gradient_aos(vp, vp_out, dt, vx, vy, (cudaStream_t)stream);
}

extern "C" int pfs_advect_color(const float *image, float *itmp, const float *vp, float dt, int ix, int iy, int iz,
                                int vx, int vy, int vz, void *stream)
{
    const char *fn = "pfs_advect_color";
    PFS_TRY(check_dims(fn, ix, iy, iz));
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_ptr(fn, "image", image));
    PFS_TRY(check_ptr(fn, "itmp", itmp));
    PFS_TRY(check_ptr(fn, "vp", vp));
    return launch_advect_color(image, itmp, vp, 4, dt, ix, iy, vx, vy, (cudaStream_t)stream);
}

// =============================================================================================
// device-pointer step API
// =============================================================================================
// The kernel sequence of one simulate_fluid_step on stream s (no argument checks, no pointer exchange).
// wait_tmp: optional event the stream must wait for before the first read of Y (its channel 2 is the
// pressure warm start, first touched by the divergence kernel) -- lets pfs_timestep_host upload tmp while
// advect and the diffusion sweeps already run on vp.
static int enqueue_fluid_step(DeviceScratch *sc, float *X, float *Y, float dt, float viscosity, int vx, int vy,
                              int n_diffuse, int n_pressure, float sigma, unsigned long long seed, unsigned step,
                              const float *forces, cudaStream_t s, cudaEvent_t wait_tmp)
{
    float *d_last = nullptr, *d_prev = nullptr, *p_last = nullptr, *p_prev = nullptr;
    float *div = sc->div();

    // advect(vp -> tmp)  (fluid.cpp:299): X.uv gathered, result kept in a (u,v) plane (iterate 0 of diffusion)
    {
        PhaseScope ph(PFS_PHASE_ADVECT, s);
        PFS_TRY(launch_advect(X, 4, sc->uv(0), 2, dt, vx, vy, s));
    }
    // diffuse(tmp -> vp)  (fluid.cpp:300); addForces slot (fluid.cpp:302, commented out in the reference): the external
    // force, if any, goes to the field struct `vp` points at after diffuse, i.e. diffusion iterate n_diffuse -- added as
    // the last fused pass stores it
    {
        PhaseScope ph(PFS_PHASE_DIFFUSE, s);
        ForceField ff{forces, 0, vy};
        PFS_TRY(run_diffuse(sc->uv(0), sc->uv(1), sc->uv(2), diffuse_params(vx, vy, viscosity, dt), n_diffuse, &d_last,
                            &d_prev, s, forces ? &ff : nullptr));
    }
    // After diffuse the struct `vp` points at the buffer written last: sweep k writes X for odd k,
    // Y for even k (sweep 1 writes vp_out = the original vp buffer X).
    float *Bv = (n_diffuse & 1) ? X : Y;     // holds iterate n_diffuse in ch0,1; its ch2 is the warm start
    float *Bo = (n_diffuse & 1) ? Y : X;     // holds iterate n_diffuse-1 in ch0,1
    // optional stochastic forcing at the same slot
    if (sigma != 0.0f) {
        PhaseScope ph(PFS_PHASE_DIFFUSE, s);
        PFS_TRY(launch_stochastic_force(d_last, d_last + 1, 2, sigma, seed, step, vx, vy, 0, 0, s));
    }
    // computePressure(vp -> tmp)  (fluid.cpp:303): divergence of iterate n_diffuse, p_0 = Bv.ch2
    if (wait_tmp) PFS_CUDA(cudaStreamWaitEvent(s, wait_tmp, 0));
    {
        PhaseScope ph(PFS_PHASE_DIVERGENCE, s);
        PFS_TRY(launch_divergence(d_last, div, Bv, sc->p(0), dt, vx, vy, s));
    }
    {
        PhaseScope ph(PFS_PHASE_PRESSURE, s);
        SweepParams pp{vx, vy, 1.0f, 4.0f};
        PFS_TRY(run_pressure(sc->p(0), sc->p(1), sc->p(2), div, pp, n_pressure, &p_last, &p_prev, s));
    }
    // Pressure sweep k writes Bo for odd k, Bv for even k; struct `tmp` ends on the buffer with p_N.
    float *Bp = (n_pressure & 1) ? Bo : Bv;
    float *Bq = (n_pressure & 1) ? Bv : Bo;
    // subtractPressureGradient(tmp -> vp)  (fluid.cpp:304) reads ch0,1 of the buffer holding p_N,
    // which carries diffusion iterate n_diffuse if that buffer is Bv, else iterate n_diffuse-1.
    const float *uv_p = (Bp == Bv) ? d_last : d_prev;
    {
        PhaseScope ph(PFS_PHASE_PROJECT, s);
        PFS_TRY(launch_project_pack(uv_p, p_last, p_prev, div, Bq, Bp, dt, vx, vy, s));
    }
    return PFS_OK;
}

// ---------------------------------------------------------------------------------------------
// Step graphs.  A timestep is ~50 short launches; at 4096^2 the gaps between them are 5 % of the step and
// at 1024^2 30 %.  The reference driver calls simulate_fluid_step with the same buffers and parameters every
// timestep (main.cpp:219-226), so the second identical call is captured into a CUDA graph (on an internal
// stream) and every later identical call replays it on the caller's stream.  Anything that changes the
// kernel sequence or its arguments is part of the key.  PFS_STEP_GRAPH=0 disables it.
// ---------------------------------------------------------------------------------------------
struct StepKey {
    int dev = -1;
    const float *X = nullptr, *Y = nullptr, *planes = nullptr, *forces = nullptr;
    size_t plane_cells = 0;
    int vx = 0, vy = 0, nd = 0, np = 0, fuse = 0;
    unsigned dt_bits = 0, visc_bits = 0;
    bool operator==(const StepKey &o) const
    {
        return dev == o.dev && X == o.X && Y == o.Y && planes == o.planes && forces == o.forces && plane_cells == o.plane_cells && vx == o.vx && vy == o.vy && nd == o.nd &&
               np == o.np && fuse == o.fuse && dt_bits == o.dt_bits && visc_bits == o.visc_bits;
    }
};
struct CachedStep {
    StepKey key;
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches = 0, passes = 0;
};
static std::vector<CachedStep> g_step_graphs;
static StepKey g_prev_key;
static std::map<int, cudaStream_t> g_capture_streams;
static int g_step_graph_mode = -1;      // -1 unread, 0 off, 1 on

static void drop_step_graphs()
{
    for (auto &c : g_step_graphs)
        if (c.exec) cudaGraphExecDestroy(c.exec);
    g_step_graphs.clear();
    g_prev_key = StepKey();
}

static void destroy_capture_streams()
{
    for (auto &kv : g_capture_streams)
        if (kv.second) cudaStreamDestroy(kv.second);
    g_capture_streams.clear();
}

static bool step_graphs_enabled()
{
    if (g_step_graph_mode < 0) {
        const char *e = getenv("PFS_STEP_GRAPH");
        g_step_graph_mode = (e && e[0] == '0') ? 0 : 1;
    }
    return g_step_graph_mode == 1;
}

static unsigned float_bits(float f)
{
    unsigned u;
    memcpy(&u, &f, sizeof(u));
    return u;
}

static int simulate_fluid_step_impl(const char *fn, float **vp, float **tmp, float dt, float viscosity, int vx, int vy,
                                    int vz, int n_diffuse, int n_pressure, float sigma, unsigned long long seed,
                                    unsigned step, void *stream, cudaEvent_t wait_tmp = nullptr, const float *forces = nullptr)
{
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_sweeps(fn, n_diffuse));
    PFS_TRY(check_sweeps(fn, n_pressure));
    if (!vp || !tmp) {
        set_error("%s: vp / tmp handle is null", fn);
        return PFS_EINVAL;
    }
    PFS_TRY(check_ptr(fn, "*vp", *vp));
    PFS_TRY(check_ptr(fn, "*tmp", *tmp));
    if (*vp == *tmp) {
        set_error("%s: vp and tmp must be distinct buffers", fn);
        return PFS_EINVAL;
    }
    if (forces != nullptr) PFS_TRY(check_ptr(fn, "forces", forces));
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScratch *sc;
    PFS_TRY(get_scratch((size_t)vx * vy, &sc));
    float *X = *vp, *Y = *tmp;

    // the buffer-pointer outcome is pure bookkeeping (see enqueue_fluid_step for the reasoning)
    float *Bv = (n_diffuse & 1) ? X : Y, *Bo = (n_diffuse & 1) ? Y : X;
    float *Bp = (n_pressure & 1) ? Bo : Bv, *Bq = (n_pressure & 1) ? Bv : Bo;

    bool use_graph = step_graphs_enabled() && sigma == 0.0f && wait_tmp == nullptr && !g_phase_timing;
    if (use_graph) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
            (void)cudaGetLastError();
            use_graph = false;          // the caller is building a graph of their own: just enqueue
        }
    }
    if (use_graph) {
        std::lock_guard<std::mutex> glock(g_graph_mutex);
        StepKey key;
        PFS_CUDA(cudaGetDevice(&key.dev));
        key.X = X; key.Y = Y; key.planes = sc->planes; key.plane_cells = sc->plane_cells; key.forces = forces;
        key.vx = vx; key.vy = vy; key.nd = n_diffuse; key.np = n_pressure; key.fuse = g_fuse_depth;
        key.dt_bits = float_bits(dt); key.visc_bits = float_bits(viscosity);
        for (auto &c : g_step_graphs) {
            if (c.key == key) {
                PFS_CUDA(cudaGraphLaunch(c.exec, s));
                g_launches += c.launches;
                g_passes += c.passes;
                *vp = Bq;
                *tmp = Bp;
                return PFS_OK;
            }
        }
        if (key == g_prev_key) {
            // second identical call in a row: capture (nothing allocates now, the first call did)
            cudaStream_t cs = g_capture_streams[key.dev];
            if (!cs) {
                PFS_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
                g_capture_streams[key.dev] = cs;
            }
            const unsigned long long l0 = g_launches, p0 = g_passes;
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                const int rc = enqueue_fluid_step(sc, X, Y, dt, viscosity, vx, vy, n_diffuse, n_pressure, 0.0f, 0ull, 0u,
                                                  forces, cs, nullptr);
                ok = (cudaStreamEndCapture(cs, &graph) == cudaSuccess) && rc == PFS_OK && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                if (g_step_graphs.size() >= 8) {
                    cudaGraphExecDestroy(g_step_graphs.front().exec);
                    g_step_graphs.erase(g_step_graphs.begin());
                }
                g_step_graphs.push_back({key, exec, g_launches - l0, g_passes - p0});
                PFS_CUDA(cudaGraphLaunch(exec, s));     // the launches counted during capture are this replay's
                *vp = Bq;
                *tmp = Bp;
                return PFS_OK;
            }
            (void)cudaGetLastError();                   // capture not possible here: stay eager from now on
            g_launches = l0;
            g_passes = p0;
            g_step_graph_mode = 0;
        }
        g_prev_key = key;
    }
    PFS_TRY(enqueue_fluid_step(sc, X, Y, dt, viscosity, vx, vy, n_diffuse, n_pressure, sigma, seed, step, forces, s, wait_tmp));
    *vp = Bq;
    *tmp = Bp;
    return PFS_OK;
}

extern "C" int pfs_simulate_fluid_step(float **vp, float **tmp, float dt, float viscosity, int vx, int vy, int vz,
                                       int n_diffuse, int n_pressure, void *stream)
{
    return simulate_fluid_step_impl("pfs_simulate_fluid_step", vp, tmp, dt, viscosity, vx, vy, vz, n_diffuse, n_pressure,
                                    0.0f, 0ull, 0u, stream);
}

extern "C" int pfs_simulate_fluid_step_stochastic(float **vp, float **tmp, float dt, float viscosity, int vx, int vy,
                                                  int vz, int n_diffuse, int n_pressure, float sigma, uint64_t seed,
                                                  uint32_t step, void *stream)
{
    return simulate_fluid_step_impl("pfs_simulate_fluid_step_stochastic", vp, tmp, dt, viscosity, vx, vy, vz, n_diffuse,
                                    n_pressure, sigma, seed, step, stream);
}

extern "C" int pfs_simulate_fluid_step_forced(float **vp, float **tmp, float dt, float viscosity, int vx, int vy, int vz,
                                              int n_diffuse, int n_pressure, const float *forces, void *stream)
{
    return simulate_fluid_step_impl("pfs_simulate_fluid_step_forced", vp, tmp, dt, viscosity, vx, vy, vz, n_diffuse,
                                    n_pressure, 0.0f, 0ull, 0u, stream, nullptr, forces);
}

extern "C" int pfs_add_forces_stochastic(float *vp, float sigma, uint64_t seed, uint32_t step, int vx, int vy, int vz,
                                         void *stream)
{
    const char *fn = "pfs_add_forces_stochastic";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_ptr(fn, "vp", vp));
    if (sigma == 0.0f) return PFS_OK;
    return launch_stochastic_force(vp, vp + 1, 4, sigma, seed, step, vx, vy, 0, 0, (cudaStream_t)stream);
}

extern "C" int pfs_image_to_rgba8(const float *image, unsigned char *out, int ix, int iy, int iz, void *stream)
{
    const char *fn = "pfs_image_to_rgba8";
    PFS_TRY(check_dims(fn, ix, iy, iz));
    PFS_TRY(check_ptr(fn, "image", image));
    if (!out || (reinterpret_cast<uintptr_t>(out) & 3u)) {
        set_error("%s: out must be a non-null, 4-byte aligned device buffer", fn);
        return PFS_EINVAL;
    }
    return launch_pack_rgba8(image, out, (size_t)ix * iy, (cudaStream_t)stream);
}

extern "C" int pfs_step_norms(const float *vp, const float *tmp, int vx, int vy, int vz, double out[4], void *stream)
{
    const char *fn = "pfs_step_norms";
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_ptr(fn, "vp", vp));
    PFS_TRY(check_ptr(fn, "tmp", tmp));
    if (!out) {
        set_error("%s: out is null", fn);
        return PFS_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int kBlocks = 1184;
    double *scratch = nullptr;
    PFS_CUDA(cudaMalloc((void **)&scratch, (4 * (size_t)kBlocks + 4) * sizeof(double)));
    int rc = launch_step_norms(vp, tmp, (size_t)vx * vy, scratch, kBlocks, scratch + 4 * (size_t)kBlocks, s);
    double host[4] = {0, 0, 0, 0};
    cudaError_t e = cudaSuccess;
    if (rc == PFS_OK) e = cudaMemcpyAsync(host, scratch + 4 * (size_t)kBlocks, sizeof(host), cudaMemcpyDeviceToHost, s);
    if (rc == PFS_OK && e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (rc != PFS_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "pfs_step_norms", __FILE__, __LINE__);
    out[0] = sqrt(host[0]);
    out[1] = sqrt(host[1]);
    out[2] = sqrt(host[2]);
    out[3] = host[3];
    return PFS_OK;
}

extern "C" int pfs_advect_color_step(float **image, float **itmp, float **vp, float dt, int ix, int iy, int iz, int vx,
                                     int vy, int vz, void *stream)
{
    const char *fn = "pfs_advect_color_step";
    if (!image || !itmp || !vp) {
        set_error("%s: image / itmp / vp handle is null", fn);
        return PFS_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    {
        PhaseScope ph(PFS_PHASE_ADVECT_COLOR, s);
        PFS_TRY(pfs_advect_color(*image, *itmp, *vp, dt, ix, iy, iz, vx, vy, vz, stream));
    }
    float *t = *image;     // fluid.cpp:317-319
    *image = *itmp;
    *itmp = t;
    return PFS_OK;
}

// advect_color_step whose kernel also stores the frame the reference's PNG writer would form from the new image
extern "C" int pfs_advect_color_step_rgba8(float **image, float **itmp, float **vp, float dt, int ix, int iy, int iz, int vx,
                                           int vy, int vz, unsigned char *rgba8_out, void *stream)
{
    const char *fn = "pfs_advect_color_step_rgba8";
    if (!image || !itmp || !vp) {
        set_error("%s: image / itmp / vp handle is null", fn);
        return PFS_EINVAL;
    }
    PFS_TRY(check_dims(fn, ix, iy, iz));
    PFS_TRY(check_dims(fn, vx, vy, vz));
    PFS_TRY(check_ptr(fn, "image", *image));
    PFS_TRY(check_ptr(fn, "itmp", *itmp));
    PFS_TRY(check_ptr(fn, "vp", *vp));
    if (!rgba8_out || (reinterpret_cast<uintptr_t>(rgba8_out) & 3u)) {
        set_error("%s: rgba8_out must be a 4-byte aligned device pointer", fn);
        return PFS_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    {
        PhaseScope ph(PFS_PHASE_ADVECT_COLOR, s);
        PFS_TRY(launch_advect_color(*image, *itmp, *vp, 4, dt, ix, iy, vx, vy, s, rgba8_out));
    }
    float *t = *image;     // fluid.cpp:317-319
    *image = *itmp;
    *itmp = t;
    return PFS_OK;
}

// =============================================================================================
// host-buffer API
// =============================================================================================
namespace pfs {

static int host_ctx(DeviceScratch **out)
{
    int dev;
    PFS_TRY(current_device(&dev));
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceScratch &sc = g_scratch[dev];
    for (int i = 0; i < 2; i++)
        if (!sc.streams[i]) PFS_CUDA(cudaStreamCreateWithFlags(&sc.streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++)
        if (!sc.ev[i]) PFS_CUDA(cudaEventCreateWithFlags(&sc.ev[i], cudaEventDisableTiming));
    *out = &sc;
    return PFS_OK;
}

static int ensure_stage(DeviceScratch *sc, int k, size_t floats)
{
    if (sc->stage_floats[k] >= floats) return PFS_OK;
    if (sc->stage[k]) {
        PFS_CUDA(cudaDeviceSynchronize());
        PFS_CUDA(cudaFree(sc->stage[k]));
        sc->stage[k] = nullptr;
        sc->stage_floats[k] = 0;
    }
    PFS_CUDA(cudaMalloc((void **)&sc->stage[k], floats * sizeof(float)));
    sc->stage_floats[k] = floats;
    return PFS_OK;
}

static int check_field(const char *fn, const char *name, const pfs_field *f)
{
    if (!f || !f->data) {
        set_error("%s: %s (or its data pointer) is null", fn, name);
        return PFS_EINVAL;
    }
    return check_dims(fn, f->x, f->y, f->z);
}

}  // namespace pfs

extern "C" int pfs_simulate_fluid_step_host(pfs_field *vp, pfs_field *tmp, float dt, float viscosity, int n_diffuse,
                                            int n_pressure)
{
    const char *fn = "pfs_simulate_fluid_step_host";
    PFS_TRY(check_field(fn, "vp", vp));
    PFS_TRY(check_field(fn, "tmp", tmp));
    if (vp->x != tmp->x || vp->y != tmp->y) {
        set_error("%s: vp and tmp must have the same shape", fn);
        return PFS_EINVAL;
    }
    DeviceScratch *sc;
    PFS_TRY(host_ctx(&sc));
    size_t n = (size_t)vp->x * vp->y * 4;
    PFS_TRY(ensure_stage(sc, 0, n));
    PFS_TRY(ensure_stage(sc, 1, n));
    cudaStream_t s = sc->streams[0];
    float *dvp = sc->stage[0], *dtmp = sc->stage[1];
    PFS_CUDA(cudaMemcpyAsync(dvp, vp->data, n * sizeof(float), cudaMemcpyHostToDevice, s));
    PFS_CUDA(cudaMemcpyAsync(dtmp, tmp->data, n * sizeof(float), cudaMemcpyHostToDevice, s));
    float *a = dvp, *b = dtmp;
    PFS_TRY(pfs_simulate_fluid_step(&a, &b, dt, viscosity, vp->x, vp->y, 4, n_diffuse, n_pressure, s));
    // mirror the pointer exchange on the caller's structs, then bring both post-state buffers back
    float *hvp = vp->data, *htmp = tmp->data;
    if (a != dvp) {
        vp->data = htmp;
        tmp->data = hvp;
    }
    PFS_CUDA(cudaMemcpyAsync(vp->data, a, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    PFS_CUDA(cudaMemcpyAsync(tmp->data, b, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    PFS_CUDA(cudaStreamSynchronize(s));
    return PFS_OK;
}

extern "C" int pfs_advect_color_step_host(pfs_field *image, pfs_field *itmp, pfs_field *vp, float dt)
{
    const char *fn = "pfs_advect_color_step_host";
    PFS_TRY(check_field(fn, "image", image));
    PFS_TRY(check_field(fn, "itmp", itmp));
    PFS_TRY(check_field(fn, "vp", vp));
    if (image->x != itmp->x || image->y != itmp->y) {
        set_error("%s: image and itmp must have the same shape", fn);
        return PFS_EINVAL;
    }
    DeviceScratch *sc;
    PFS_TRY(host_ctx(&sc));
    size_t nv = (size_t)vp->x * vp->y * 4, ni = (size_t)image->x * image->y * 4;
    PFS_TRY(ensure_stage(sc, 0, nv));
    PFS_TRY(ensure_stage(sc, 2, ni));
    PFS_TRY(ensure_stage(sc, 3, ni));
    cudaStream_t s = sc->streams[0];
    PFS_CUDA(cudaMemcpyAsync(sc->stage[0], vp->data, nv * sizeof(float), cudaMemcpyHostToDevice, s));
    PFS_CUDA(cudaMemcpyAsync(sc->stage[2], image->data, ni * sizeof(float), cudaMemcpyHostToDevice, s));
    PFS_TRY(pfs_advect_color(sc->stage[2], sc->stage[3], sc->stage[0], dt, image->x, image->y, 4, vp->x, vp->y, 4, s));
    // fluid.cpp:316-319: the result is written to itmp's buffer, then the two data pointers swap
    PFS_CUDA(cudaMemcpyAsync(itmp->data, sc->stage[3], ni * sizeof(float), cudaMemcpyDeviceToHost, s));
    PFS_CUDA(cudaStreamSynchronize(s));
    float *t = image->data;
    image->data = itmp->data;
    itmp->data = t;
    return PFS_OK;
}

extern "C" int pfs_timestep_host(pfs_field *vp, pfs_field *vtmp, pfs_field *image, pfs_field *itmp, float dt,
                                 float viscosity, int n_diffuse, int n_pressure)
{
    const char *fn = "pfs_timestep_host";
    PFS_TRY(check_field(fn, "vp", vp));
    PFS_TRY(check_field(fn, "vtmp", vtmp));
    PFS_TRY(check_field(fn, "image", image));
    PFS_TRY(check_field(fn, "itmp", itmp));
    if (vp->x != vtmp->x || vp->y != vtmp->y || image->x != itmp->x || image->y != itmp->y) {
        set_error("%s: vp/vtmp and image/itmp must pairwise have the same shape", fn);
        return PFS_EINVAL;
    }
    DeviceScratch *sc;
    PFS_TRY(host_ctx(&sc));
    size_t nv = (size_t)vp->x * vp->y * 4, ni = (size_t)image->x * image->y * 4;
    PFS_TRY(ensure_stage(sc, 0, nv));
    PFS_TRY(ensure_stage(sc, 1, nv));
    PFS_TRY(ensure_stage(sc, 2, ni));
    PFS_TRY(ensure_stage(sc, 3, ni));
    cudaStream_t s0 = sc->streams[0], s1 = sc->streams[1];
    float *dvp = sc->stage[0], *dtmp = sc->stage[1], *dimg = sc->stage[2], *ditmp = sc->stage[3];

    // PCIe is the bottleneck of this call (3 buffers up, 3 down), so the copies are ordered to keep both
    // directions busy and the kernels hidden behind them:
    //   stream 0: vp up | advect + diffusion (need only vp) | wait for vtmp | divergence .. project | vp, vtmp down
    //   stream 1:          vtmp up | image up                | wait for the step | advect_color | image down
    // The first read of vtmp (pressure warm start) waits for its upload through an event.
    PFS_CUDA(cudaMemcpyAsync(dvp, vp->data, nv * sizeof(float), cudaMemcpyHostToDevice, s0));
    PFS_CUDA(cudaEventRecord(sc->ev[2], s0));
    PFS_CUDA(cudaStreamWaitEvent(s1, sc->ev[2], 0));          // keep the H2D engine order vp, vtmp, image
    PFS_CUDA(cudaMemcpyAsync(dtmp, vtmp->data, nv * sizeof(float), cudaMemcpyHostToDevice, s1));
    PFS_CUDA(cudaEventRecord(sc->ev[0], s1));
    PFS_CUDA(cudaMemcpyAsync(dimg, image->data, ni * sizeof(float), cudaMemcpyHostToDevice, s1));
    float *a = dvp, *b = dtmp;
    PFS_TRY(simulate_fluid_step_impl(fn, &a, &b, dt, viscosity, vp->x, vp->y, 4, n_diffuse, n_pressure, 0.0f, 0ull, 0u,
                                     s0, sc->ev[0]));
    PFS_CUDA(cudaEventRecord(sc->ev[1], s0));
    float *hvp = vp->data, *htmp = vtmp->data;
    if (a != dvp) {
        vp->data = htmp;
        vtmp->data = hvp;
    }
    // stream 1: advect_color once the image is up and the projected velocity exists, image down.
    PFS_CUDA(cudaStreamWaitEvent(s1, sc->ev[1], 0));
    float *di = dimg, *dt2 = ditmp, *dv = a;
    PFS_TRY(pfs_advect_color_step(&di, &dt2, &dv, dt, image->x, image->y, 4, vp->x, vp->y, 4, s1));
    // stream 0: velocity post-state down (starts while the image is still going up / being advected)
    PFS_CUDA(cudaMemcpyAsync(vp->data, a, nv * sizeof(float), cudaMemcpyDeviceToHost, s0));
    PFS_CUDA(cudaMemcpyAsync(vtmp->data, b, nv * sizeof(float), cudaMemcpyDeviceToHost, s0));
    PFS_CUDA(cudaMemcpyAsync(itmp->data, di, ni * sizeof(float), cudaMemcpyDeviceToHost, s1));
    PFS_CUDA(cudaStreamSynchronize(s0));
    PFS_CUDA(cudaStreamSynchronize(s1));
    float *t = image->data;     // fluid.cpp:317-319
    image->data = itmp->data;
    itmp->data = t;
    return PFS_OK;
}
