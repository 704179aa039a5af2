// pfs_ctx.cu -- persistent-state contexts (include/pfs_b200.h "pfs_ctx_*", SURVEY.md 8b "Interface").
//
// The stateless entry points (pfs_api.cu) must assume that the caller touches the interleaved buffers between two
// calls, so every step reads them (advect gathers from 16-byte cells to use 8 bytes, the divergence kernel walks a
// whole buffer to extract channel 2) and writes both of them back in full: 64 of the ~110 bytes per cell that the
// one-pass kernels of a step move.  A context owns the state instead, in the layout the kernels want -- (u,v) planes,
// scalar pressure and divergence planes, the image ping-pong -- and the interleaved form only exists in pfs_ctx_upload
// and pfs_ctx_download.  What a step leaves behind is exactly what fluid.cpp:298-320 leaves in its two buffers (the
// logical buffers `vp` and `tmp` below), so a download after n steps equals n stateless steps bit for bit.
//
// Logical state between steps (which plane holds which channel of which reference buffer):
//   vp  = [ uv[uvX], p[pX], div[dX] ]     tmp = [ uv[uvY], p[pY], div[dY] ]
// One step (the reference's pointer choreography for any sweep-count parity, see pfs_api.cu enqueue_fluid_step):
//   advect            uv[uvX] -> A                      (A, B, C: the three (u,v) planes that are not uvX)
//   diffuse           A <-> B, iterate n-1 -> C         -> d_last, d_prev
//   divergence        d_last -> div[0]
//   pressure          warm start = the p plane of the buffer struct `vp` points at after diffuse (tmp's for an even
//                     sweep count, vp's for an odd one), ping-pong with the other two p planes
//   project           (d_last or d_prev, p_last) -> uv[uvX]   (the old projected field is dead once advect has read it)
//   new state         vp = [uvX, p_prev, div0]   tmp = [uv_p, p_last, div0]
//   advect_color      img[cur] -> img[cur^1] through uv[uvX]; cur ^= 1
// Steps whose plane roles repeat (they cycle with a period of at most six) are captured into CUDA graphs and replayed.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "pfs_internal.cuh"

using namespace pfs;

struct pfs_ctx {
    int device = 0;
    int vx = 0, vy = 0, ix = 0, iy = 0;
    size_t cells = 0, unit = 0;           // unit = cells padded to 64 (256-byte aligned planes)
    float *mem = nullptr;                 // 4 (u,v) planes (2 units each) + 3 p + 2 div = 13 units
    float *img[2] = {nullptr, nullptr};   // interleaved RGBA image ping-pong
    float *vmax = nullptr;                // device scalar: max|v| of the projected field (by-product of project)
    int uvX = 0, uvY = 1, pX = 0, pY = 1, dX = 0, dY = 1, cur = 0;
    bool loaded = false;
    unsigned long long steps = 0;
    std::mutex m;                         // one step at a time per context
    struct Graph {
        int warm, cur, nd, np, fuse;
        unsigned dt_bits, visc_bits;
        cudaGraphExec_t exec;
        unsigned long long launches, passes;
        // the roles the step leaves behind
        int uvY, pX, pY;
    };
    std::vector<Graph> graphs;
    cudaStream_t capture_stream = nullptr;
    int graph_mode = -1;                  // -1 unread, 0 off, 1 on

    float *uv(int k) const { return mem + (size_t)(2 * k) * unit; }
    float *p(int k) const { return mem + (size_t)(8 + k) * unit; }
    float *div(int k) const { return mem + (size_t)(11 + k) * unit; }
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

unsigned float_bits(float f)
{
    unsigned u;
    memcpy(&u, &f, sizeof(u));
    return u;
}

int check_ctx(const char *fn, const pfs_ctx *c, bool need_state)
{
    if (!c) {
        set_error("%s: context is null", fn);
        return PFS_EINVAL;
    }
    if (need_state && !c->loaded) {
        set_error("%s: the context holds no state yet (call pfs_ctx_upload first)", fn);
        return PFS_ESTATE;
    }
    return PFS_OK;
}

struct Roles {
    int uvY, pX, pY;
};

// The kernel sequence of simulate_fluid_step on the context's planes (no locking, no role update).
int enqueue_fluid(pfs_ctx *c, float dt, float viscosity, int nd, int np, float sigma, unsigned long long seed, unsigned step,
                  const float *forces, cudaStream_t s, Roles *out)
{
    const int vx = c->vx, vy = c->vy;
    int free_uv[3], q = 0;
    for (int k = 0; k < 4; k++)
        if (k != c->uvX) free_uv[q++] = k;
    float *A = c->uv(free_uv[0]), *B = c->uv(free_uv[1]), *C = c->uv(free_uv[2]);
    float *d_last = nullptr, *d_prev = nullptr, *p_last = nullptr, *p_prev = nullptr;
    {
        PhaseScope ph(PFS_PHASE_ADVECT, s);
        PFS_TRY(launch_advect(c->uv(c->uvX), 2, A, 2, dt, vx, vy, s));
    }
    {
        PhaseScope ph(PFS_PHASE_DIFFUSE, s);
        ForceField ff{forces, 0, vy};
        PFS_TRY(run_diffuse(A, B, C, diffuse_params(vx, vy, viscosity, dt), nd, &d_last, &d_prev, s, forces ? &ff : nullptr));
        if (sigma != 0.0f) PFS_TRY(launch_stochastic_force(d_last, d_last + 1, 2, sigma, seed, step, vx, vy, 0, 0, s));
    }
    // struct `vp` points at buffer Bv after diffuse: the original vp buffer for an odd sweep count, else tmp's
    const bool bv_is_x = (nd & 1) != 0;
    const int warm = bv_is_x ? c->pX : c->pY;
    const int oth = (warm + 1) % 3, ext = (warm + 2) % 3;
    {
        PhaseScope ph(PFS_PHASE_DIVERGENCE, s);
        PFS_TRY(launch_divergence(d_last, c->div(0), nullptr, nullptr, dt, vx, vy, s));
    }
    {
        PhaseScope ph(PFS_PHASE_PRESSURE, s);
        SweepParams pp{vx, vy, 1.0f, 4.0f};
        PFS_TRY(run_pressure(c->p(warm), c->p(oth), c->p(ext), c->div(0), pp, np, &p_last, &p_prev, s));
    }
    // struct `tmp` ends on the buffer holding p_N: Bv for an even pressure count, the other buffer for an odd one;
    // the gradient is subtracted from that buffer's (u,v): diffusion iterate n if it is Bv, else iterate n-1
    const bool bp_is_bv = (np & 1) == 0;
    float *uv_p = bp_is_bv ? d_last : d_prev;
    {
        PhaseScope ph(PFS_PHASE_PROJECT, s);
        PFS_TRY(launch_project_uv(uv_p, p_last, c->uv(c->uvX), dt, vx, vy, s, 0, 1, nullptr));
    }
    auto uv_index = [&](const float *ptr) { return (int)((ptr - c->mem) / (ptrdiff_t)(2 * c->unit)); };
    auto p_index = [&](const float *ptr) { return (int)((ptr - c->p(0)) / (ptrdiff_t)c->unit); };
    out->uvY = uv_index(uv_p);
    out->pX = p_index(p_prev);
    out->pY = p_index(p_last);
    return PFS_OK;
}

int enqueue_color(pfs_ctx *c, float dt, cudaStream_t s, unsigned char *rgba8 = nullptr)
{
    PhaseScope ph(PFS_PHASE_ADVECT_COLOR, s);
    return launch_advect_color(c->img[c->cur], c->img[c->cur ^ 1], c->uv(c->uvX), 2, dt, c->ix, c->iy, c->vx, c->vy, s, rgba8);
}

void apply_roles(pfs_ctx *c, const Roles &r)
{
    c->uvY = r.uvY;
    c->pX = r.pX;
    c->pY = r.pY;
    c->dX = c->dY = 0;
}

bool graphs_enabled(pfs_ctx *c)
{
    if (c->graph_mode < 0) {
        const char *e = getenv("PFS_STEP_GRAPH");
        c->graph_mode = (e && e[0] == '0') ? 0 : 1;
    }
    return c->graph_mode == 1;
}

}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" int pfs_ctx_create(pfs_ctx **out, int vx, int vy, int ix, int iy)
{
    const char *fn = "pfs_ctx_create";
    if (!out) {
        set_error("%s: out is null", fn);
        return PFS_EINVAL;
    }
    PFS_TRY(check_dims(fn, vx, vy, 4));
    if (ix != 0 || iy != 0) PFS_TRY(check_dims(fn, ix, iy, 4));
    pfs_ctx *c = new pfs_ctx();
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) {
        delete c;
        set_error("%s: no usable CUDA device (%s); libpfs_b200 has no CPU fallback", fn, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return PFS_ENODEVICE;
    }
    c->vx = vx; c->vy = vy; c->ix = ix; c->iy = iy;
    c->cells = (size_t)vx * vy;
    c->unit = (c->cells + 63) & ~(size_t)63;
    int st = PFS_OK;
    if ((e = cudaMalloc((void **)&c->mem, 13 * c->unit * sizeof(float))) != cudaSuccess) st = cuda_fail(e, "cudaMalloc planes", __FILE__, __LINE__);
    const size_t ibytes = (size_t)ix * iy * 4 * sizeof(float);
    for (int k = 0; k < 2 && st == PFS_OK && ibytes > 0; k++)
        if ((e = cudaMalloc((void **)&c->img[k], ibytes)) != cudaSuccess) st = cuda_fail(e, "cudaMalloc image", __FILE__, __LINE__);
    if (st == PFS_OK && (e = cudaMalloc((void **)&c->vmax, 4 * sizeof(float))) != cudaSuccess) st = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
    if (st == PFS_OK && (e = cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking)) != cudaSuccess) st = cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
    if (st != PFS_OK) {
        pfs_ctx_destroy(c);
        return st;
    }
    *out = c;
    return PFS_OK;
}

extern "C" int pfs_ctx_destroy(pfs_ctx *c)
{
    if (!c) return PFS_OK;
    {
        DeviceGuard g(c->device);
        cudaDeviceSynchronize();
        for (auto &gr : c->graphs)
            if (gr.exec) cudaGraphExecDestroy(gr.exec);
        if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
        if (c->mem) cudaFree(c->mem);
        if (c->img[0]) cudaFree(c->img[0]);
        if (c->img[1]) cudaFree(c->img[1]);
        if (c->vmax) cudaFree(c->vmax);
        (void)cudaGetLastError();
    }
    delete c;
    return PFS_OK;
}

extern "C" int pfs_ctx_upload(pfs_ctx *c, const float *vp, const float *tmp, const float *image, void *stream)
{
    const char *fn = "pfs_ctx_upload";
    PFS_TRY(check_ctx(fn, c, false));
    if (!c->loaded && (!vp || !tmp || (c->ix > 0 && !image))) {
        set_error("%s: the first upload needs vp, tmp%s", fn, c->ix > 0 ? " and image" : "");
        return PFS_EINVAL;
    }
    if (vp) PFS_TRY(check_ptr(fn, "vp", vp));
    if (tmp) PFS_TRY(check_ptr(fn, "tmp", tmp));
    if (image) PFS_TRY(check_ptr(fn, "image", image));
    if (image && c->ix == 0) {
        set_error("%s: the context was created without an image", fn);
        return PFS_EINVAL;
    }
    std::lock_guard<std::mutex> lock(c->m);
    DeviceGuard g(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    // vp and tmp get planes of their own again (after a step both structs share the divergence plane)
    if (vp || tmp) {
        if (!c->loaded) {
            c->uvX = 0; c->uvY = 1; c->pX = 0; c->pY = 1; c->dX = 0; c->dY = 1;
        } else {
            // after a step both structs share one divergence plane: the replaced buffer takes the spare plane (about to be
            // overwritten), the other one keeps the shared plane; (u,v) and pressure planes are never shared
            if (c->dX == c->dY) {
                const int spare = 1 - c->dX;
                if (vp) c->dX = spare; else c->dY = spare;
            }
        }
        if (vp) PFS_TRY(launch_unpack(vp, c->uv(c->uvX), c->p(c->pX), c->div(c->dX), c->vx, c->vy, s));
        if (tmp) PFS_TRY(launch_unpack(tmp, c->uv(c->uvY), c->p(c->pY), c->div(c->dY), c->vx, c->vy, s));
    }
    if (image)
        PFS_CUDA(cudaMemcpyAsync(c->img[c->cur], image, (size_t)c->ix * c->iy * 4 * sizeof(float), cudaMemcpyDefault, s));
    c->loaded = true;
    return PFS_OK;
}

extern "C" int pfs_ctx_download(pfs_ctx *c, float *vp, float *tmp, float *image, void *stream)
{
    const char *fn = "pfs_ctx_download";
    PFS_TRY(check_ctx(fn, c, true));
    if (vp) PFS_TRY(check_ptr(fn, "vp", vp));
    if (tmp) PFS_TRY(check_ptr(fn, "tmp", tmp));
    if (image) PFS_TRY(check_ptr(fn, "image", image));
    if (image && c->ix == 0) {
        set_error("%s: the context was created without an image", fn);
        return PFS_EINVAL;
    }
    std::lock_guard<std::mutex> lock(c->m);
    DeviceGuard g(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (vp) PFS_TRY(launch_pack(vp, c->uv(c->uvX), c->p(c->pX), c->div(c->dX), c->vx, c->vy, s));
    if (tmp) PFS_TRY(launch_pack(tmp, c->uv(c->uvY), c->p(c->pY), c->div(c->dY), c->vx, c->vy, s));
    if (image)
        PFS_CUDA(cudaMemcpyAsync(image, c->img[c->cur], (size_t)c->ix * c->iy * 4 * sizeof(float), cudaMemcpyDefault, s));
    return PFS_OK;
}

extern "C" int pfs_ctx_image(pfs_ctx *c, const float **image_dev)
{
    PFS_TRY(check_ctx("pfs_ctx_image", c, true));
    if (!image_dev || c->ix == 0) {
        set_error("pfs_ctx_image: no image in this context (or null argument)");
        return PFS_EINVAL;
    }
    *image_dev = c->img[c->cur];
    return PFS_OK;
}

static int ctx_fluid_step(const char *fn, pfs_ctx *c, float dt, float viscosity, int nd, int np, float sigma,
                          unsigned long long seed, unsigned step, const float *forces, void *stream)
{
    PFS_TRY(check_ctx(fn, c, true));
    PFS_TRY(check_sweeps(fn, nd));
    PFS_TRY(check_sweeps(fn, np));
    if (forces) PFS_TRY(check_ptr(fn, "forces", forces));
    std::lock_guard<std::mutex> lock(c->m);
    DeviceGuard g(c->device);
    Roles r;
    PFS_TRY(enqueue_fluid(c, dt, viscosity, nd, np, sigma, seed, step, forces, (cudaStream_t)stream, &r));
    apply_roles(c, r);
    return PFS_OK;
}

extern "C" int pfs_ctx_simulate_fluid_step(pfs_ctx *c, float dt, float viscosity, int n_diffuse, int n_pressure, void *stream)
{
    return ctx_fluid_step("pfs_ctx_simulate_fluid_step", c, dt, viscosity, n_diffuse, n_pressure, 0.0f, 0ull, 0u, nullptr, stream);
}

extern "C" int pfs_ctx_simulate_fluid_step_forced(pfs_ctx *c, float dt, float viscosity, int n_diffuse, int n_pressure,
                                                  const float *forces, void *stream)
{
    return ctx_fluid_step("pfs_ctx_simulate_fluid_step_forced", c, dt, viscosity, n_diffuse, n_pressure, 0.0f, 0ull, 0u, forces,
                          stream);
}

extern "C" int pfs_ctx_simulate_fluid_step_stochastic(pfs_ctx *c, float dt, float viscosity, int n_diffuse, int n_pressure,
                                                      float sigma, uint64_t seed, uint32_t step, void *stream)
{
    return ctx_fluid_step("pfs_ctx_simulate_fluid_step_stochastic", c, dt, viscosity, n_diffuse, n_pressure, sigma, seed, step,
                          nullptr, stream);
}

extern "C" int pfs_ctx_advect_color_step(pfs_ctx *c, float dt, void *stream)
{
    const char *fn = "pfs_ctx_advect_color_step";
    PFS_TRY(check_ctx(fn, c, true));
    if (c->ix == 0) {
        set_error("%s: the context was created without an image", fn);
        return PFS_EINVAL;
    }
    std::lock_guard<std::mutex> lock(c->m);
    DeviceGuard g(c->device);
    PFS_TRY(enqueue_color(c, dt, (cudaStream_t)stream));
    c->cur ^= 1;                                             // fluid.cpp:317-319
    return PFS_OK;
}

extern "C" int pfs_ctx_advect_color_step_rgba8(pfs_ctx *c, float dt, unsigned char *rgba8_out, void *stream)
{
    const char *fn = "pfs_ctx_advect_color_step_rgba8";
    PFS_TRY(check_ctx(fn, c, true));
    if (c->ix == 0) {
        set_error("%s: the context was created without an image", fn);
        return PFS_EINVAL;
    }
    if (!rgba8_out || (reinterpret_cast<uintptr_t>(rgba8_out) & 3u)) {
        set_error("%s: rgba8_out must be a 4-byte aligned device pointer", fn);
        return PFS_EINVAL;
    }
    std::lock_guard<std::mutex> lock(c->m);
    DeviceGuard g(c->device);
    {
        PhaseScope ph(PFS_PHASE_ADVECT_COLOR, (cudaStream_t)stream);
        PFS_TRY(enqueue_color(c, dt, (cudaStream_t)stream, rgba8_out));
    }
    c->cur ^= 1;                                             // fluid.cpp:317-319
    return PFS_OK;
}

// n_steps iterations of the reference driver loop (main.cpp:236-239): simulate_fluid_step + advect_color_step.
extern "C" int pfs_ctx_step(pfs_ctx *c, int n_steps, float dt, float viscosity, int n_diffuse, int n_pressure, void *stream)
{
    const char *fn = "pfs_ctx_step";
    PFS_TRY(check_ctx(fn, c, true));
    PFS_TRY(check_sweeps(fn, n_diffuse));
    PFS_TRY(check_sweeps(fn, n_pressure));
    if (n_steps < 0) {
        set_error("%s: n_steps must be >= 0", fn);
        return PFS_EINVAL;
    }
    std::lock_guard<std::mutex> lock(c->m);
    DeviceGuard g(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    bool use_graph = graphs_enabled(c) && !phase_timing_on();
    if (use_graph) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
            (void)cudaGetLastError();
            use_graph = false;          // the caller is building a graph of their own: just enqueue
        }
    }
    for (int it = 0; it < n_steps; it++) {
        const int warm = (n_diffuse & 1) ? c->pX : c->pY;
        pfs_ctx::Graph *hit = nullptr;
        if (use_graph) {
            for (auto &gr : c->graphs)
                if (gr.warm == warm && gr.cur == c->cur && gr.nd == n_diffuse && gr.np == n_pressure && gr.fuse == fuse_depth() &&
                    gr.dt_bits == float_bits(dt) && gr.visc_bits == float_bits(viscosity))
                    hit = &gr;
        }
        if (hit) {
            PFS_CUDA(cudaGraphLaunch(hit->exec, s));
            g_launches += hit->launches;
            g_passes += hit->passes;
            apply_roles(c, Roles{hit->uvY, hit->pX, hit->pY});
            if (c->ix > 0) c->cur ^= 1;
            c->steps++;
            continue;
        }
        Roles r;
        if (use_graph && c->steps >= 1) {
            // capture this role assignment (the first step of a context ran eagerly: host-side caches are warm)
            const unsigned long long l0 = g_launches, p0 = g_passes;
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            cudaStream_t cs = c->capture_stream;
            bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                int rc = enqueue_fluid(c, dt, viscosity, n_diffuse, n_pressure, 0.0f, 0ull, 0u, nullptr, cs, &r);
                if (rc == PFS_OK && c->ix > 0) rc = enqueue_color(c, dt, cs);
                ok = (cudaStreamEndCapture(cs, &graph) == cudaSuccess) && rc == PFS_OK && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                if (c->graphs.size() >= 12) {
                    cudaGraphExecDestroy(c->graphs.front().exec);
                    c->graphs.erase(c->graphs.begin());
                }
                c->graphs.push_back({warm, c->cur, n_diffuse, n_pressure, fuse_depth(), float_bits(dt), float_bits(viscosity), exec,
                                     g_launches - l0, g_passes - p0, r.uvY, r.pX, r.pY});
                PFS_CUDA(cudaGraphLaunch(exec, s));       // the launches counted during capture are this replay's
                apply_roles(c, r);
                if (c->ix > 0) c->cur ^= 1;
                c->steps++;
                continue;
            }
            (void)cudaGetLastError();                     // capture not possible here: stay eager from now on
            g_launches = l0;
            g_passes = p0;
            c->graph_mode = 0;
            use_graph = false;
        }
        PFS_TRY(enqueue_fluid(c, dt, viscosity, n_diffuse, n_pressure, 0.0f, 0ull, 0u, nullptr, s, &r));
        apply_roles(c, r);
        if (c->ix > 0) {
            PFS_TRY(enqueue_color(c, dt, s));
            c->cur ^= 1;
        }
        c->steps++;
    }
    return PFS_OK;
}
