// slab.cu -- the fluid step on a row-slab decomposition of one periodic grid over several GPUs
// (BASELINE.json configs[3]: 16384^2 over 2/4/8 B200).
//
// Decomposition: rank r of R owns a contiguous band of rows of the velocity grid and the band of image
// rows whose velocity look-up (fluid.cpp:89-90) falls into it; x stays whole, so a halo row is one
// contiguous run of W floats per plane and the x wrap stays local.  The domain is periodic, so the
// ranks form a ring (R = 1 wraps onto itself).  Planes carry `halo` rows above and below the band; the
// kernels are the single-GPU kernels with the row map (y_base = halo, wrap = 0).
//
// What moves between neighbours each timestep (rows are W*4 bytes per plane):
//   advect        D rows of (u,v) each way, D = ceil(max|dt*v/H|) + 2 from an all-reduce(max) on the first step,
//                 guessed from the previous step and verified by a device flag afterwards (rerun if too shallow)
//   diffusion     `halo` (32) rows x 2 planes; a pass of depth t then also recomputes the halo-t rows next to
//                 the band (an "extended interior"), so the next passes find valid halos without another
//                 exchange: one exchange per 32/t passes, ~1.5 % redundant rows on a 2048-row band
//   divergence    nothing if the last diffusion sweep left >= 1 valid halo row; then `halo` rows of the
//                 divergence (the fused pressure passes need the right-hand side in their halo trapezoid)
//   pressure      as diffusion, 1 plane; the gradient needs 1 valid halo row of p_N
//   advect_color  D_i rows of the image each way
// Results are bit-identical to the single-GPU path (same per-cell arithmetic, global indices).
//
// Transports, one process per GPU: peer stores into the neighbours' planes through CUDA IPC mappings (one kernel
// per exchange, flag handshakes; set up over an NCCL communicator, which also carries the all-reduces), or NCCL
// send/recv when a rank cannot map its neighbours (NCCL is resolved with dlopen, so the single-GPU library has no
// NCCL dependency).  Slabs that live in one process use direct copies (any devices; used by the tests to run
// R slabs on one GPU and by single-process multi-GPU callers).
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "pfs_internal.cuh"

namespace pfs {

constexpr int MAX_HALO = 32;   // halo rows per side: one exchange feeds MAX_HALO / depth fused passes
constexpr int MIN_HALO = 8;    // >= the deepest fused pass
constexpr int N_SLAB_PLANES = 13;   // units of plane_floats: (u,v) ping [0,1] | (u,v) pong [2,3] | p x2 [4][5] | div [6] |
                                    // iterate n-1 of the last pass: (u,v) [7,8] | p [9] | resident state: projected (u,v) [10,11],
                                    // second divergence plane [12] (only until the first step after an upload).  A (u,v) plane is two units wide.
constexpr int UV_A = 0, UV_B = 2, UV_X = 7, P_A = 4, P_B = 5, P_X = 9, DIV = 6, UV_S = 10, DIV2 = 12;
constexpr int P2P_GATHER_ROWS = 64; // capacity (rows per side) of the peer-written gather halos; deeper gathers go through NCCL
constexpr int P2P_FLAG_INTS = 16;

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen.  The handful of types and enum values the calls below need are declared here (they are part of
// NCCL's stable ABI, nccl.h 2.x), so that the library builds on a box without NCCL headers; every function is resolved
// at run time and the single-GPU paths never touch any of it.
// ---------------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;          // NCCL_UNIQUE_ID_BYTES
typedef int ncclResult_t;
constexpr ncclResult_t ncclSuccess = 0;
typedef int ncclDataType_t;
constexpr ncclDataType_t ncclUint8 = 1, ncclFloat = 7, ncclDouble = 8;
typedef int ncclRedOp_t;
constexpr ncclRedOp_t ncclSum = 0, ncclMax = 2, ncclMin = 3;

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;   // optional
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool loaded = false;
};

static NcclApi &nccl()
{
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    void *h = nullptr;
    const char *env = getenv("PFS_NCCL_LIB");
    for (const char *name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return a;
#define PFS_NCCL_SYM(field, sym) a.field = (decltype(a.field))dlsym(h, sym)
    PFS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    PFS_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    PFS_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    PFS_NCCL_SYM(GroupStart, "ncclGroupStart");
    PFS_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    PFS_NCCL_SYM(Send, "ncclSend");
    PFS_NCCL_SYM(Recv, "ncclRecv");
    PFS_NCCL_SYM(AllReduce, "ncclAllReduce");
    PFS_NCCL_SYM(Broadcast, "ncclBroadcast");
    PFS_NCCL_SYM(AllGather, "ncclAllGather");
    PFS_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef PFS_NCCL_SYM
    a.loaded = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv &&
               a.AllReduce && a.Broadcast && a.GetErrorString;
    return a;
}

#define PFS_NCCL(call)                                                                            \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) {                                                                 \
            set_error("NCCL error %d (%s) at %s:%d: %s", (int)r__, nccl().GetErrorString(r__), __FILE__, __LINE__, #call); \
            return PFS_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// partition (pure host logic; exported for the CPU tests)
// ---------------------------------------------------------------------------------------------
static void band(int total, int parts, int r, int *first, int *count)
{
    const int base = total / parts, rem = total % parts;
    *first = r * base + std::min(r, rem);
    *count = base + (r < rem ? 1 : 0);
}

// image rows j whose velocity row (int)((float)j * vih) lies in [row0, row0+rows)  (fluid.cpp:83,90)
static void image_band(int ih, int gh, int row0, int rows, int *irow0, int *irows)
{
    const float vih = (float)gh / (float)ih;
    int first = ih, last = -1;
    for (int j = 0; j < ih; j++) {
        const int vj = (int)((float)j * vih);
        if (vj >= row0 && vj < row0 + rows) {
            first = std::min(first, j);
            last = std::max(last, j);
        }
    }
    if (last < first) {
        *irow0 = 0;
        *irows = 0;
    } else {
        *irow0 = first;
        *irows = last - first + 1;
    }
}

}  // namespace pfs

// ---------------------------------------------------------------------------------------------
// slab context
// ---------------------------------------------------------------------------------------------
struct pfs_slab {
    int rank = 0, nranks = 1;
    int gw = 0, gh = 0, row0 = 0, rows = 0;
    int iw = 0, ih = 0, irow0 = 0, irows = 0;
    int device = 0;
    int halo = pfs::MIN_HALO;             // halo rows per side of every plane (multiple of 8, <= the thinnest band)
    size_t plane_floats = 0;
    float *planes = nullptr;              // N_SLAB_PLANES planes of (rows + 2*halo) x gw
    // gather sources of advect / advect_color: D halo rows received from each ring neighbour (interleaved
    // cells, [above | below]); or -- when D exceeds a band -- a copy of the whole field (all-gather)
    float4 *vhalo = nullptr, *ihalo = nullptr;
    size_t vhalo_rows = 0, ihalo_rows = 0;         // capacity in rows (both sides together)
    float4 *vwhole = nullptr, *iwhole = nullptr;
    size_t vwhole_rows = 0, iwhole_rows = 0;
    float *d_scalars = nullptr;           // [0] max|v| (advect), [1] max|v| (advect_color), [2] sticky error word (as int: gather
                                          // overflow under an exact bound 1/2/3, peer-transport timeout 9; read by pfs_slab_check),
                                          // [3] setup scratch, [4] max|v| and [5] "a departure row was missing" (as float) of a
                                          // speculative advect, [6] the same flag as the kernel raises it (int; reset every step),
                                          // [8] "row missing" flag of a speculative advect_color (int), [9] the same as float (all-reduced),
                                          // [10] max|v| of the projected field of the last resident step (by-product of project)
    float *h_scalars = nullptr;           // pinned mirror
    pfs::ncclComm_t comm = nullptr;
    // peer transport (one process per GPU, CUDA IPC): the ring neighbours' planes / gather halos / flags mapped here
    struct PeerLink {
        float *planes = nullptr;
        float4 *vhalo = nullptr, *ihalo = nullptr;
        int *flags = nullptr;
        size_t plane_floats = 0;
        int rows = 0, irows = 0;
    };
    bool p2p = false;
    PeerLink link_up, link_down;
    int *flags = nullptr;                 // [0] ready<-up [1] ready<-down [2] pushed<-up [3] pushed<-down [4] CTA counter
    float4 *vhalo_p2p = nullptr, *ihalo_p2p = nullptr;   // [above: P2P_GATHER_ROWS rows | below: same], written by the neighbours
    // Exchanges over the peer transport are numbered (the same on every rank); a push waits for / publishes its number in
    // the neighbours' flag words.  The number is formed ON THE DEVICE as flags[5] + ordinal, so that a captured graph of
    // pushes can be replayed: xord counts the pushes enqueued since flags[5] was last advanced (flush_seq).
    int xord = 0;
    // the sweeps of a resident step (diffusion, divergence, pressure and their halo exchanges) as replayable graphs
    struct SegGraph {
        int p_warm, nd, np, fuse, halo;
        unsigned dt_bits, visc_bits;
        cudaGraphExec_t exec;
        unsigned long long launches, passes;
        int dl, dpv, pl_last, pl_prev, p_valid;
    };
    std::vector<SegGraph> seg_graphs;
    std::vector<SegGraph> seg_seen;       // keys launched one by one so far: a key is captured the second time it turns up
    cudaStream_t capture_stream = nullptr;
    int graph_mode = -1;                  // -1 unread, 0 off, 1 on (PFS_SLAB_GRAPH=0 disables)
    std::vector<void *> ipc_opened;
    std::vector<pfs_slab *> group;        // in-process transport: all ranks, indexed by rank (empty under NCCL)
    cudaStream_t stream = nullptr;        // stream of the call in flight
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_meas = nullptr, ev_fork = nullptr;
    cudaStream_t side = nullptr;          // carries the all-reduce + read-back of a speculative advect's (max|v|, flag)
    float vbound = -1.f;                  // max|v| of the field the last fluid step started from (< 0: not known yet)
    // resident state (pfs_slab_upload / pfs_slab_step / pfs_slab_download): which plane unit holds which channel of the
    // reference's two buffers -- vp = [uv UV_S, p pX, div dX], tmp = [uv uvY, p pY, div dY] -- and the image ping-pong
    bool resident = false;
    int uvY = pfs::UV_A, pX = pfs::P_A, pY = pfs::P_B, dX = pfs::DIV, dY = pfs::DIV2;
    float *img[2] = {nullptr, nullptr};
    int img_cur = 0;
    bool pending_color = false;           // a speculative advect_color whose "row missing" flag has not been looked at yet
    float pending_dt = 0.f;
    cudaEvent_t ev_color = nullptr;
    float *plane(int k) const { return planes + (size_t)k * plane_floats; }
    float *interior(int k, size_t row_floats) const { return plane(k) + (size_t)halo * row_floats; }
    int up() const { return (rank + nranks - 1) % nranks; }
    int down() const { return (rank + 1) % nranks; }
};

namespace pfs {

namespace {

struct Guard {   // cudaSetDevice for the scope of one slab's work
    int prev = -1;
    explicit Guard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~Guard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- small kernels of the slab path -----------------------------------------------------------
// flag -> out[1] as a float; vmax (optional: the running maximum a project kernel left) -> out[0]
__global__ void flag_to_float_kernel(const int *flag, float *out, const float *vmax)
{
    out[1] = (*flag != 0) ? 1.f : 0.f;
    if (vmax != nullptr) out[0] = *vmax;
}

template <int CF>     // floats per cell: 4 = interleaved [u,v,p,div], 2 = (u,v) plane
__global__ void __launch_bounds__(256) max_abs_v_kernel(const float *__restrict__ vp, size_t n, float *out)
{
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = fabsf(__ldg(vp + i * CF + 1));
        m = (v > m || v != v) ? v : m;           // NaN propagates: the host then takes the all-gather path
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, m, o);
        m = (t > m || t != t) ? t : m;
    }
    if ((threadIdx.x & 31) == 0) {
        if (m != m) m = __int_as_float(0x7f800000);   // NaN -> +inf, so that every rank takes the same decision
        atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));   // non-negative floats order like their bits
    }
}

// Rows of a periodic field of `total` rows: the local band plus d halo rows received from each ring neighbour
// (or the whole field: band0 = 0, band_n = total, d = 0).  Rows are `row_bytes` apart; the cell type is the kernel's business.
struct RowSource {
    const char *band, *above, *below;
    int band0, band_n, d, total;
    size_t row_bytes;
};

// The same lookup without branches, for the two bilinear rows of every cell: the three sources are one affine map each
// (base + r * row_bytes with the base shifted so that r is the row relative to the band), chosen by two compares.  A row
// outside all three raises *miss (the caller flags it once) and reads the band's first row, so nothing is out of bounds.
__device__ __forceinline__ const char *source_row_select(const RowSource &S, int grow, bool *miss)
{
    int r = grow - S.band0;
    r += (r < -S.d) ? S.total : 0;
    r -= (r >= S.band_n + S.d) ? S.total : 0;
    const bool out = (r < -S.d) || (r >= S.band_n + S.d);
    *miss = *miss || out;
    r = out ? 0 : r;
    const char *base = (r < 0) ? S.above + (size_t)S.d * S.row_bytes
                               : ((r >= S.band_n) ? S.below - (size_t)S.band_n * S.row_bytes : S.band);
    return base + (ptrdiff_t)r * (ptrdiff_t)S.row_bytes;
}

// advect (fluid.cpp:24-70) with GLOBAL indices: this rank produces rows [row0, row0+rows) of the gh-row
// grid from a field whose cells are CF floats apart with (u,v) first (an interleaved buffer or a (u,v) plane).
constexpr int GATHER_ROWS = 2;      // rows per thread of the two gather kernels (rows jl and jl + 4 of an 8-row tile), staged so
                                    // that the two cells' memory round trips overlap (kernels_basic.cu: advect_kernel)

template <int CF>
__global__ void __launch_bounds__(256)
    advect_slab_kernel(const RowSource S, float2 *__restrict__ uv_out, float dt, int w, int gh,
                       int row0, int rows, int y_base, int *overflow, float *vmax_out, float rfw, float rfh)
{
    constexpr int R = GATHER_ROWS;
    const int i = blockIdx.x * 64 + threadIdx.x, jb = blockIdx.y * (4 * R) + threadIdx.y;
    const float fw = (float)w, fh = (float)gh;
    auto cell = [](const char *row, int x) { return reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(row) + (size_t)x * CF); };
    bool live[R];
    int jl[R];
    float2 uv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        jl[r] = jb + 4 * r;
        live[r] = (i < w && jl[r] < rows);
        // the cell's own row is a row of the band (row0 <= j < row0 + rows, and the band source starts at band0 <= row0)
        uv[r] = live[r] ? __ldg(cell(S.band + (size_t)(row0 + jl[r] - S.band0) * S.row_bytes, i)) : make_float2(0.f, 0.f);
    }
    if (vmax_out != nullptr) {
        // by-product: max|v| of the field being advected (what bounds the row displacement), NaN -> +inf.
        // Reduced over the whole CTA first (every thread gets here, dead ones with 0): ONE look at the running
        // maximum per CTA, and an atomic only when the CTA would raise it -- a look per warp was a million loads of
        // one address per step, all served by the same L2 slice.
        __shared__ float warp_max[8];
        float m = 0.f;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float a = fabsf(uv[r].y);
            m = (a > m || a != a) ? a : m;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float t = __shfl_xor_sync(0xffffffffu, m, o);
            m = (t > m || t != t) ? t : m;
        }
        const int tid = threadIdx.y * 64 + threadIdx.x;
        if ((tid & 31) == 0) warp_max[tid >> 5] = m;
        __syncthreads();
        if (tid < 32) {
            m = (tid < 8) ? warp_max[tid] : 0.f;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const float t = __shfl_xor_sync(0xffffffffu, m, o);
                m = (t > m || t != t) ? t : m;
            }
            if (tid == 0) {
                if (m != m) m = __int_as_float(0x7f800000);
                if (m > *reinterpret_cast<volatile float *>(vmax_out))
                    atomicMax(reinterpret_cast<int *>(vmax_out), __float_as_int(m));
            }
        }
    }
    Bilinear b[R];
    const char *r0[R], *r1[R];
    bool miss = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float xp = backtrace_coord((float)i, __fmul_rn(dt, uv[r].x), fw, rfw);
        const float yp = backtrace_coord((float)(row0 + (live[r] ? jl[r] : 0)), __fmul_rn(dt, uv[r].y), fh, rfh);
        b[r] = make_bilinear(xp, yp, w, gh);
        bool mr = false;
        r0[r] = source_row_select(S, b[r].j0, &mr);
        r1[r] = source_row_select(S, b[r].j1, &mr);
        miss = miss || (mr && live[r]);
    }
    if (miss) atomicExch(overflow, 1);       // cannot happen if D was computed correctly; a guessed D is verified through this word
    float2 f00[R], f10[R], f01[R], f11[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int x0 = live[r] ? b[r].i0 : 0, x1 = live[r] ? b[r].i1 : 0;     // dead threads read cell 0 of a valid row
        f00[r] = __ldg(cell(r0[r], x0));
        f10[r] = __ldg(cell(r0[r], x1));
        f01[r] = __ldg(cell(r1[r], x0));
        f11[r] = __ldg(cell(r1[r], x1));
    }
#pragma unroll
    for (int r = 0; r < R; r++)
        if (live[r])
            uv_out[(size_t)(y_base + jl[r]) * w + i] =
                make_float2(bilerp(b[r], f00[r].x, f10[r].x, f01[r].x, f11[r].x), bilerp(b[r], f00[r].y, f10[r].y, f01[r].y, f11[r].y));
}

// advect_color (fluid.cpp:72-127) with GLOBAL indices: image rows [irow0, irow0+irows) of the ih-row
// image; velocity rows [row0, row0+rows) (local, cells VS floats apart: the image bands are built so that the
// look-up of fluid.cpp:89-90 never leaves the rank's own velocity band).
template <int VS>
__global__ void __launch_bounds__(256)
    advect_color_slab_kernel(const RowSource S, float4 *__restrict__ out, const float *__restrict__ vp,
                             float dt_over_viw, float dt_over_vih, float viw, float vih, int iw, int ih, int irow0,
                             int irows, int vw, int row0, int rows, int *overflow, float rfiw, float rfih)
{
    constexpr int R = GATHER_ROWS;
    const int i = blockIdx.x * 64 + threadIdx.x, jb = blockIdx.y * (4 * R) + threadIdx.y;
    if (i >= iw || jb >= irows) return;
    const float fiw = (float)iw, fih = (float)ih;
    const int vi = (int)__fmul_rn((float)i, viw);
    bool live[R];
    int jl[R];
    float2 uv[R];
    bool vmiss = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
        live[r] = jb + 4 * r < irows;
        jl[r] = live[r] ? jb + 4 * r : jb;
        int vj = (int)__fmul_rn((float)(irow0 + jl[r]), vih) - row0;
        if (vj < 0 || vj >= rows) {
            vmiss = true;
            vj = 0;
        }
        uv[r] = __ldg(reinterpret_cast<const float2 *>(vp + ((size_t)vj * vw + vi) * VS));
    }
    if (vmiss) {
        atomicExch(overflow, 2);
        return;
    }
    Bilinear b[R];
    const char *c0[R], *c1[R];
    bool miss = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float xp = backtrace_coord((float)i, __fmul_rn(dt_over_viw, uv[r].x), fiw, rfiw);
        const float yp = backtrace_coord((float)(irow0 + jl[r]), __fmul_rn(dt_over_vih, uv[r].y), fih, rfih);
        b[r] = make_bilinear(xp, yp, iw, ih);
        c0[r] = source_row_select(S, b[r].j0, &miss);
        c1[r] = source_row_select(S, b[r].j1, &miss);
    }
    if (miss) {
        atomicExch(overflow, 3);
        return;
    }
    float4 f00[R], f10[R], f01[R], f11[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float4 *q0 = reinterpret_cast<const float4 *>(c0[r]), *q1 = reinterpret_cast<const float4 *>(c1[r]);
        f00[r] = __ldg(q0 + b[r].i0);
        f10[r] = __ldg(q0 + b[r].i1);
        f01[r] = __ldg(q1 + b[r].i0);
        f11[r] = __ldg(q1 + b[r].i1);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        float4 o;
        o.x = bilerp(b[r], f00[r].x, f10[r].x, f01[r].x, f11[r].x);
        o.y = bilerp(b[r], f00[r].y, f10[r].y, f01[r].y, f11[r].y);
        o.z = bilerp(b[r], f00[r].z, f10[r].z, f01[r].z, f11[r].z);
        o.w = bilerp(b[r], f00[r].w, f10[r].w, f01[r].w, f11[r].w);
        if (live[r]) out[(size_t)jl[r] * iw + i] = o;
    }
}

// ---- peer transport: halo rows stored straight into the ring neighbours' memory -------------------------------
// One launch per exchange.  Flags live in each rank's own memory and are written by its neighbours (system-scope
// release stores through the IPC mapping); `seq` counts exchanges and is the same on every rank.
//   1. tell both neighbours "my edge rows for exchange seq are final and my halos may be overwritten"
//   2. wait for the same word from both, then store my top rows into the upper neighbour's bottom halo and my
//      bottom rows into the lower neighbour's top halo (16-byte stores over NVLink, whole grid)
//   3. the last CTA to finish fences, tells both neighbours "pushed", and waits until both have pushed to me,
//      so when the kernel retires this rank's halos are complete.
// A wait that lasts longer than 20 s gives up and raises the slab's error flag (pfs_slab_check) instead of hanging.
struct PushSeg {
    const float4 *src_up, *src_down;
    float4 *dst_up, *dst_down;
    unsigned long long n16;
};
struct PushArgs {
    PushSeg seg[3];
    int nseg;
    int *mine, *up, *down, *error_flag;
    int ordinal;                          // this push is number mine[5] + ordinal
};

__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void wait_flag(const int *p, int seq, int *error_flag)
{
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(p) < seq) {
        if (global_ns() - t0 > 20000000000ull) {
            atomicExch(error_flag, 9);
            return;
        }
    }
}

__global__ void advance_seq_kernel(int *base, int n) { *base += n; }

__global__ void __launch_bounds__(256) halo_push_kernel(const PushArgs A)
{
    int seq = 0;
    if (threadIdx.x == 0) {
        seq = *reinterpret_cast<volatile const int *>(A.mine + 5) + A.ordinal;     // advanced only between pushes, in stream order
        if (blockIdx.x == 0) {
            st_release_sys(A.up + 1, seq);
            st_release_sys(A.down + 0, seq);
        }
        wait_flag(A.mine + 0, seq, A.error_flag);
        wait_flag(A.mine + 1, seq, A.error_flag);
    }
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nth = (size_t)gridDim.x * 256;
    for (int q = 0; q < A.nseg; q++) {
        const PushSeg g = A.seg[q];
        for (size_t i = tid; i < g.n16; i += nth) {
            g.dst_up[i] = g.src_up[i];
            g.dst_down[i] = g.src_down[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(A.mine + 4, 1);
        if (done == (int)gridDim.x - 1) {
            atomicExch(A.mine + 4, 0);
            __threadfence_system();
            st_release_sys(A.up + 3, seq);
            st_release_sys(A.down + 2, seq);
            wait_flag(A.mine + 2, seq, A.error_flag);
            wait_flag(A.mine + 3, seq, A.error_flag);
        }
    }
}

int launch_halo_push(pfs_slab *s, PushArgs &A)
{
    unsigned long long total = 0;
    for (int q = 0; q < A.nseg; q++) total += A.seg[q].n16;
    int blocks = (int)std::min<unsigned long long>(128, (total + 255) / 256);
    if (blocks < 1) blocks = 1;
    A.mine = s->flags;
    A.up = s->link_up.flags;
    A.down = s->link_down.flags;
    A.error_flag = reinterpret_cast<int *>(s->d_scalars + 2);
    A.ordinal = ++s->xord;
    PFS_LAUNCH(halo_push_kernel, blocks, 256, 0, s->stream, A);
    return PFS_OK;
}

// Folds the pushes enqueued so far into the device-side base (in stream order), so that the next push is number base + 1
// again: called before a graph of pushes is captured or replayed, and as the last node of such a graph.
int flush_seq(pfs_slab *s)
{
    if (!s->p2p || s->xord == 0) return PFS_OK;
    PFS_LAUNCH(advance_seq_kernel, 1, 1, 0, s->stream, s->flags + 5, s->xord);
    s->xord = 0;
    return PFS_OK;
}

// ---- transport ----------------------------------------------------------------------------------
// One exchange = for every local slab a list of segments; a segment sends `bytes` from send_up to the
// upper neighbour's recv_from_down and from send_down to the lower neighbour's recv_from_up.
struct Segment {
    const char *send_up, *send_down;
    char *recv_from_up, *recv_from_down;
    size_t bytes;
};

int ring_exchange(const std::vector<pfs_slab *> &local, const std::vector<std::vector<Segment>> &segs)
{
    if (local.empty()) return PFS_OK;
    if (local[0]->comm != nullptr) {
        // NCCL: one process per rank.  Order inside the group: all sends (up, down per segment), then all
        // receives (from_down, from_up per segment) -- with two ranks both neighbours are the same
        // peer, and NCCL pairs sends and receives to one peer in issue order.
        pfs_slab *s = local[0];
        Guard g(s->device);
        const std::vector<Segment> &v = segs[0];
        PFS_NCCL(nccl().GroupStart());
        for (const Segment &sg : v) {
            PFS_NCCL(nccl().Send(sg.send_up, sg.bytes, ncclUint8, s->up(), s->comm, s->stream));
            PFS_NCCL(nccl().Send(sg.send_down, sg.bytes, ncclUint8, s->down(), s->comm, s->stream));
        }
        for (const Segment &sg : v) {
            PFS_NCCL(nccl().Recv(sg.recv_from_down, sg.bytes, ncclUint8, s->down(), s->comm, s->stream));
            PFS_NCCL(nccl().Recv(sg.recv_from_up, sg.bytes, ncclUint8, s->up(), s->comm, s->stream));
        }
        PFS_NCCL(nccl().GroupEnd());
        return PFS_OK;
    }
    // in-process: every rank is local.  Producer streams publish "ready", consumers copy on their own
    // stream, then publish "done" so that a producer never overwrites rows a neighbour is still reading.
    const std::vector<pfs_slab *> &all = local[0]->group;
    for (pfs_slab *s : local) {
        Guard g(s->device);
        PFS_CUDA(cudaEventRecord(s->ev_ready, s->stream));
    }
    for (size_t k = 0; k < local.size(); k++) {
        pfs_slab *s = local[k];
        pfs_slab *up = all[s->up()], *dn = all[s->down()];
        size_t ku = 0, kd = 0;
        for (size_t q = 0; q < local.size(); q++) {
            if (local[q] == up) ku = q;
            if (local[q] == dn) kd = q;
        }
        Guard g(s->device);
        PFS_CUDA(cudaStreamWaitEvent(s->stream, up->ev_ready, 0));
        PFS_CUDA(cudaStreamWaitEvent(s->stream, dn->ev_ready, 0));
        for (size_t i = 0; i < segs[k].size(); i++) {
            const Segment &mine = segs[k][i];
            // my top halo <- the upper neighbour's bottom rows; my bottom halo <- the lower neighbour's top rows
            PFS_CUDA(cudaMemcpyAsync(mine.recv_from_up, segs[ku][i].send_down, mine.bytes, cudaMemcpyDefault, s->stream));
            PFS_CUDA(cudaMemcpyAsync(mine.recv_from_down, segs[kd][i].send_up, mine.bytes, cudaMemcpyDefault, s->stream));
        }
        PFS_CUDA(cudaEventRecord(s->ev_done, s->stream));
    }
    for (pfs_slab *s : local) {
        Guard g(s->device);
        PFS_CUDA(cudaStreamWaitEvent(s->stream, all[s->up()]->ev_done, 0));
        PFS_CUDA(cudaStreamWaitEvent(s->stream, all[s->down()]->ev_done, 0));
    }
    return PFS_OK;
}

// halo exchange of `t` rows of the given planes (same plane units on every slab); `rf` = floats per plane row
// (gw for a scalar plane, 2*gw for a (u,v) plane)
int exchange_planes(const std::vector<pfs_slab *> &local, const std::vector<std::vector<float *>> &planes, int t, size_t rf)
{
    if (local.size() == 1 && local[0]->p2p && local[0]->gw % 4 == 0 && planes[0].size() <= 3) {
        pfs_slab *s = local[0];
        Guard g(s->device);
        PushArgs A;
        A.nseg = 0;
        for (float *p : planes[0]) {
            const size_t k = (size_t)(p - s->planes) / s->plane_floats;           // plane unit: same on every rank
            PushSeg &sg = A.seg[A.nseg++];
            sg.src_up = reinterpret_cast<const float4 *>(p + (size_t)s->halo * rf);
            sg.src_down = reinterpret_cast<const float4 *>(p + (size_t)(s->halo + s->rows - t) * rf);
            sg.dst_up = reinterpret_cast<float4 *>(s->link_up.planes + k * s->link_up.plane_floats +
                                                   (size_t)(s->halo + s->link_up.rows) * rf);
            sg.dst_down = reinterpret_cast<float4 *>(s->link_down.planes + k * s->link_down.plane_floats +
                                                     (size_t)(s->halo - t) * rf);
            sg.n16 = (unsigned long long)t * rf / 4;
        }
        return launch_halo_push(s, A);
    }
    std::vector<std::vector<Segment>> segs(local.size());
    for (size_t k = 0; k < local.size(); k++) {
        pfs_slab *s = local[k];
        const size_t row = rf * sizeof(float);
        for (float *p : planes[k]) {
            char *b = reinterpret_cast<char *>(p);
            Segment sg;
            sg.send_up = b + (size_t)s->halo * row;
            sg.send_down = b + (size_t)(s->halo + s->rows - t) * row;
            sg.recv_from_up = b + (size_t)(s->halo - t) * row;
            sg.recv_from_down = b + (size_t)(s->halo + s->rows) * row;
            sg.bytes = (size_t)t * row;
            segs[k].push_back(sg);
        }
    }
    return ring_exchange(local, segs);
}

// max over all ranks of a non-negative device scalar; returns it on the host (synchronises the streams)
int global_max(const std::vector<pfs_slab *> &local, int slot, float *out)
{
    float m = 0.f;
    bool nan = false;
    for (pfs_slab *s : local) {
        Guard g(s->device);
        if (s->comm != nullptr)
            PFS_NCCL(nccl().AllReduce(s->d_scalars + slot, s->d_scalars + slot, 1, ncclFloat, ncclMax, s->comm, s->stream));
        PFS_CUDA(cudaMemcpyAsync(s->h_scalars + slot, s->d_scalars + slot, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    }
    for (pfs_slab *s : local) {
        Guard g(s->device);
        PFS_CUDA(cudaStreamSynchronize(s->stream));
        const float v = s->h_scalars[slot];
        if (v != v) nan = true;
        m = std::max(m, v);
    }
    *out = nan ? INFINITY : m;
    return PFS_OK;
}

int ensure_bytes(void **ptr, size_t *have_rows, size_t want_rows, size_t row_bytes)
{
    if (*have_rows >= want_rows && *ptr) return PFS_OK;
    if (*ptr) {
        PFS_CUDA(cudaDeviceSynchronize());
        PFS_CUDA(cudaFree(*ptr));
        *ptr = nullptr;
        *have_rows = 0;
    }
    PFS_CUDA(cudaMalloc(ptr, want_rows * row_bytes));
    *have_rows = want_rows;
    return PFS_OK;
}

// Build the gather source of one interleaved field on every local slab: the band itself (read in place
// from the caller's buffer) plus D halo rows from the ring neighbours, or -- when D exceeds a band -- a
// copy of the whole field assembled from every rank's band (always sufficient).
struct FieldBands {
    int width, total;                         // cells per row, rows of the whole field
    size_t cell_bytes;                        // 16: interleaved cells, 8: a (u,v) plane
    std::vector<const char *> band;           // local slabs' bands (first row of the band)
    std::vector<int> first, count;            // global first row / rows of each local band
    bool image;                               // selects the halo / whole buffers of the slab
};

int build_row_sources(const std::vector<pfs_slab *> &local, const FieldBands &F, int D, bool whole,
                      std::vector<RowSource> *out)
{
    const size_t n = local.size();
    const size_t row = (size_t)F.width * F.cell_bytes;
    out->resize(n);
    if (!whole && n == 1 && local[0]->p2p && D <= P2P_GATHER_ROWS && row % 16 == 0 &&
        (F.image ? local[0]->ihalo_p2p : local[0]->vhalo_p2p)) {
        // peer transport: my top D rows become the upper neighbour's "below" rows, my bottom D rows the lower
        // neighbour's "above" rows; both land in the fixed-capacity halos that were IPC-mapped at connect time
        // (capacity: P2P_GATHER_ROWS rows of 16-byte cells per side; rows of 8-byte cells use half of it)
        pfs_slab *s = local[0];
        Guard g(s->device);
        const size_t cap = (size_t)P2P_GATHER_ROWS * F.width;            // float4 units per side
        float4 *mine = F.image ? s->ihalo_p2p : s->vhalo_p2p;
        float4 *up = F.image ? s->link_up.ihalo : s->link_up.vhalo;
        float4 *down = F.image ? s->link_down.ihalo : s->link_down.vhalo;
        PushArgs A;
        A.nseg = 1;
        A.seg[0].src_up = reinterpret_cast<const float4 *>(F.band[0]);
        A.seg[0].src_down = reinterpret_cast<const float4 *>(F.band[0] + (size_t)(F.count[0] - D) * row);
        A.seg[0].dst_up = up + cap;                 // its rows [band_n, band_n + D)
        A.seg[0].dst_down = down;                   // its rows [-D, 0)
        A.seg[0].n16 = (unsigned long long)D * row / 16;
        (*out)[0] = RowSource{F.band[0], reinterpret_cast<const char *>(mine), reinterpret_cast<const char *>(mine + cap),
                              F.first[0], F.count[0], D, F.total, row};
        return launch_halo_push(s, A);
    }
    if (!whole) {
        std::vector<std::vector<Segment>> segs(n);
        for (size_t k = 0; k < n; k++) {
            pfs_slab *s = local[k];
            Guard g(s->device);
            float4 **hb = F.image ? &s->ihalo : &s->vhalo;
            size_t *cap = F.image ? &s->ihalo_rows : &s->vhalo_rows;
            PFS_TRY(ensure_bytes((void **)hb, cap, 2 * (size_t)D, (size_t)F.width * 16));
            char *above = reinterpret_cast<char *>(*hb), *below = above + (size_t)D * row;
            const char *b = F.band[k];
            Segment sg;
            sg.send_up = b;
            sg.send_down = b + (size_t)(F.count[k] - D) * row;
            sg.recv_from_up = above;
            sg.recv_from_down = below;
            sg.bytes = (size_t)D * row;
            segs[k].push_back(sg);
            (*out)[k] = RowSource{F.band[k], above, below, F.first[k], F.count[k], D, F.total, row};
        }
        return ring_exchange(local, segs);
    }
    // whole field: every slab assembles all bands in global row order
    std::vector<char *> buf(n);
    for (size_t k = 0; k < n; k++) {
        pfs_slab *s = local[k];
        Guard g(s->device);
        float4 **wb = F.image ? &s->iwhole : &s->vwhole;
        size_t *cap = F.image ? &s->iwhole_rows : &s->vwhole_rows;
        PFS_TRY(ensure_bytes((void **)wb, cap, (size_t)F.total, (size_t)F.width * 16));
        buf[k] = reinterpret_cast<char *>(*wb);
        if (F.count[k] > 0)
            PFS_CUDA(cudaMemcpyAsync(buf[k] + (size_t)F.first[k] * row, F.band[k], (size_t)F.count[k] * row,
                                     cudaMemcpyDeviceToDevice, s->stream));
        (*out)[k] = RowSource{buf[k], nullptr, nullptr, 0, F.total, 0, F.total, row};
        if (s->comm != nullptr) {
            for (int r = 0; r < s->nranks; r++) {
                int f, c;
                band(s->gh, s->nranks, r, &f, &c);
                if (F.image) {
                    int jf, jc;
                    image_band(s->ih, s->gh, f, c, &jf, &jc);
                    f = jf;
                    c = jc;
                }
                if (c == 0) continue;
                char *p = buf[k] + (size_t)f * row;
                PFS_NCCL(nccl().Broadcast(p, p, (size_t)c * row, ncclUint8, r, s->comm, s->stream));
            }
        } else {
            PFS_CUDA(cudaEventRecord(s->ev_ready, s->stream));
        }
    }
    if (local[0]->comm == nullptr && n > 1) {
        for (size_t k = 0; k < n; k++) {
            pfs_slab *s = local[k];
            Guard g(s->device);
            for (size_t q = 0; q < n; q++) {
                if (q == k || F.count[q] == 0) continue;
                PFS_CUDA(cudaStreamWaitEvent(s->stream, local[q]->ev_ready, 0));
                PFS_CUDA(cudaMemcpyAsync(buf[k] + (size_t)F.first[q] * row, buf[q] + (size_t)F.first[q] * row,
                                         (size_t)F.count[q] * row, cudaMemcpyDefault, s->stream));
            }
            PFS_CUDA(cudaEventRecord(s->ev_done, s->stream));
        }
        for (pfs_slab *s : local) {
            Guard g(s->device);
            for (pfs_slab *o : local)
                if (o != s) PFS_CUDA(cudaStreamWaitEvent(s->stream, o->ev_done, 0));
        }
    }
    return PFS_OK;
}

int check_local(const char *fn, pfs_slab *const *slabs, int n_local, std::vector<pfs_slab *> *out)
{
    if (!slabs || n_local < 1) {
        set_error("%s: no slabs given", fn);
        return PFS_EINVAL;
    }
    for (int k = 0; k < n_local; k++) {
        if (!slabs[k]) {
            set_error("%s: slab %d is null", fn, k);
            return PFS_EINVAL;
        }
        out->push_back(slabs[k]);
    }
    pfs_slab *s0 = slabs[0];
    if (s0->comm == nullptr) {
        if ((int)s0->group.size() != s0->nranks || n_local != s0->nranks) {
            set_error("%s: in-process transport needs every rank of the ring in the call (got %d of %d); "
                      "call pfs_slab_connect_local or pfs_slab_connect_nccl first", fn, n_local, s0->nranks);
            return PFS_ESTATE;
        }
        for (int k = 0; k < n_local; k++) {
            if (slabs[k]->rank != k) {
                set_error("%s: slabs must be passed in rank order", fn);
                return PFS_EINVAL;
            }
        }
    } else if (n_local != 1) {
        set_error("%s: the NCCL transport drives exactly one slab per process", fn);
        return PFS_EINVAL;
    }
    return PFS_OK;
}

}  // namespace
}  // namespace pfs

using namespace pfs;

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" int pfs_slab_partition(int rank, int nranks, int gh, int ih, int *row0, int *rows, int *irow0, int *irows)
{
    if (nranks < 1 || rank < 0 || rank >= nranks || gh < 1) {
        set_error("pfs_slab_partition: bad rank/nranks/height");
        return PFS_EINVAL;
    }
    int f, c;
    band(gh, nranks, rank, &f, &c);
    if (row0) *row0 = f;
    if (rows) *rows = c;
    if (ih > 0) {
        int jf, jc;
        image_band(ih, gh, f, c, &jf, &jc);
        if (irow0) *irow0 = jf;
        if (irows) *irows = jc;
    }
    return PFS_OK;
}

extern "C" int pfs_slab_destroy(pfs_slab *s);

extern "C" int pfs_slab_create(pfs_slab **out, int rank, int nranks, int gw, int gh, int iw, int ih)
{
    const char *fn = "pfs_slab_create";
    if (!out || nranks < 1 || rank < 0 || rank >= nranks || gw < 1 || gh < 1 || iw < 0 || ih < 0) {
        set_error("%s: bad arguments", fn);
        return PFS_EINVAL;
    }
    if ((size_t)gw * gh > ((size_t)1 << 28) || (size_t)iw * ih > ((size_t)1 << 28)) {
        set_error("%s: grid exceeds 2^28 cells (the reference's int32 index limit)", fn);
        return PFS_EINVAL;
    }
    if (gh > PFS_MAX_ROWS || ih > PFS_MAX_ROWS) {
        set_error("%s: height exceeds the supported maximum of %d rows", fn, PFS_MAX_ROWS);
        return PFS_EINVAL;
    }
    if (gh / nranks < MIN_HALO) {
        set_error("%s: every slab needs at least %d rows (grid height %d over %d ranks)", fn, MIN_HALO, gh, nranks);
        return PFS_EINVAL;
    }
    pfs_slab *s = new pfs_slab();
    s->rank = rank;
    s->nranks = nranks;
    s->gw = gw;
    s->gh = gh;
    s->iw = iw;
    s->ih = ih;
    band(gh, nranks, rank, &s->row0, &s->rows);
    if (ih > 0) image_band(ih, gh, s->row0, s->rows, &s->irow0, &s->irows);
    cudaError_t e = cudaGetDevice(&s->device);
    if (e != cudaSuccess) {
        delete s;
        set_error("%s: no usable CUDA device (%s); there is no CPU fallback", fn, cudaGetErrorString(e));
        return PFS_ENODEVICE;
    }
    s->halo = std::min(MAX_HALO, ((gh / nranks) / 8) * 8);      // same on every rank: the thinnest band decides
    if (const char *e = getenv("PFS_SLAB_HALO")) {
        const int v = atoi(e);
        if (v >= MIN_HALO && v <= s->halo) s->halo = (v / 8) * 8;
    }
    s->plane_floats = ((size_t)(s->rows + 2 * s->halo) * gw + 63) & ~(size_t)63;
    int st = PFS_OK;
    auto fail = [&](cudaError_t ce, const char *what) {
        st = cuda_fail(ce, what, __FILE__, __LINE__);
    };
    if ((e = cudaMalloc((void **)&s->planes, pfs::N_SLAB_PLANES * s->plane_floats * sizeof(float))) != cudaSuccess) fail(e, "cudaMalloc planes");
    if (st == PFS_OK && (e = cudaMemset(s->planes, 0, pfs::N_SLAB_PLANES * s->plane_floats * sizeof(float))) != cudaSuccess) fail(e, "cudaMemset");
    if (st == PFS_OK && (e = cudaMalloc((void **)&s->d_scalars, 16 * sizeof(float))) != cudaSuccess) fail(e, "cudaMalloc scalars");
    if (st == PFS_OK && (e = cudaMemset(s->d_scalars, 0, 16 * sizeof(float))) != cudaSuccess) fail(e, "cudaMemset");
    if (st == PFS_OK && (e = cudaMallocHost((void **)&s->h_scalars, 16 * sizeof(float))) != cudaSuccess) fail(e, "cudaMallocHost");
    if (st == PFS_OK && (e = cudaEventCreateWithFlags(&s->ev_color, cudaEventDisableTiming)) != cudaSuccess) fail(e, "event");
    if (st == PFS_OK && (e = cudaEventCreateWithFlags(&s->ev_ready, cudaEventDisableTiming)) != cudaSuccess) fail(e, "event");
    if (st == PFS_OK && (e = cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming)) != cudaSuccess) fail(e, "event");
    if (st == PFS_OK && (e = cudaEventCreateWithFlags(&s->ev_meas, cudaEventDisableTiming)) != cudaSuccess) fail(e, "event");
    if (st == PFS_OK && (e = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming)) != cudaSuccess) fail(e, "event");
    if (st == PFS_OK && (e = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking)) != cudaSuccess) fail(e, "stream");
    if (st != PFS_OK) {
        pfs_slab_destroy(s);
        return st;
    }
    if (nranks == 1) s->group.assign(1, s);   // a single slab is its own ring
    *out = s;
    return PFS_OK;
}

namespace pfs {
namespace {
void p2p_release(pfs_slab *s);
}
}  // namespace pfs

extern "C" int pfs_slab_destroy(pfs_slab *s)
{
    if (!s) return PFS_OK;
    Guard g(s->device);
    cudaDeviceSynchronize();
    p2p_release(s);
    if (s->comm && nccl().loaded) nccl().CommDestroy(s->comm);
    if (s->planes) cudaFree(s->planes);
    if (s->vhalo) cudaFree(s->vhalo);
    if (s->ihalo) cudaFree(s->ihalo);
    if (s->vwhole) cudaFree(s->vwhole);
    if (s->iwhole) cudaFree(s->iwhole);
    if (s->d_scalars) cudaFree(s->d_scalars);
    if (s->h_scalars) cudaFreeHost(s->h_scalars);
    if (s->ev_ready) cudaEventDestroy(s->ev_ready);
    if (s->ev_done) cudaEventDestroy(s->ev_done);
    if (s->ev_meas) cudaEventDestroy(s->ev_meas);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_color) cudaEventDestroy(s->ev_color);
    if (s->img[0]) cudaFree(s->img[0]);
    if (s->img[1]) cudaFree(s->img[1]);
    if (s->side) cudaStreamDestroy(s->side);
    for (auto &gr : s->seg_graphs)
        if (gr.exec) cudaGraphExecDestroy(gr.exec);
    if (s->capture_stream) cudaStreamDestroy(s->capture_stream);
    (void)cudaGetLastError();
    delete s;
    return PFS_OK;
}

extern "C" int pfs_slab_rows(const pfs_slab *s, int *row0, int *rows, int *irow0, int *irows)
{
    if (!s) {
        set_error("pfs_slab_rows: slab is null");
        return PFS_EINVAL;
    }
    if (row0) *row0 = s->row0;
    if (rows) *rows = s->rows;
    if (irow0) *irow0 = s->irow0;
    if (irows) *irows = s->irows;
    return PFS_OK;
}

extern "C" int pfs_slab_connect_local(pfs_slab *const *slabs, int n)
{
    const char *fn = "pfs_slab_connect_local";
    if (!slabs || n < 1) {
        set_error("%s: no slabs", fn);
        return PFS_EINVAL;
    }
    for (int k = 0; k < n; k++) {
        if (!slabs[k] || slabs[k]->nranks != n || slabs[k]->rank != k) {
            set_error("%s: need all %d ranks of one ring, in rank order", fn, n);
            return PFS_EINVAL;
        }
    }
    for (int k = 0; k < n; k++) slabs[k]->group.assign(slabs, slabs + n);
    return PFS_OK;
}

extern "C" int pfs_slab_nccl_unique_id(char id[128])
{
    if (!nccl().loaded) {
        set_error("pfs_slab_nccl_unique_id: libnccl.so.2 could not be loaded (set PFS_NCCL_LIB)");
        return PFS_ESTATE;
    }
    ncclUniqueId u;
    PFS_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return PFS_OK;
}

namespace pfs {
namespace {

// What a rank publishes to the ring when the peer transport is set up (all-gathered through NCCL).
struct PeerInfo {
    cudaIpcMemHandle_t planes, flags, vhalo, ihalo;
    unsigned long long plane_floats;
    int rows, irows, device, has_image;
};

void p2p_release(pfs_slab *s)
{
    for (void *p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    s->ipc_opened.clear();
    if (s->flags) cudaFree(s->flags);
    if (s->vhalo_p2p) cudaFree(s->vhalo_p2p);
    if (s->ihalo_p2p) cudaFree(s->ihalo_p2p);
    s->flags = nullptr;
    s->vhalo_p2p = s->ihalo_p2p = nullptr;
    s->link_up = s->link_down = pfs_slab::PeerLink();
    s->p2p = false;
    (void)cudaGetLastError();
}

// Collective over the ring (called from pfs_slab_connect_nccl).  Every rank shares its planes, its gather halos
// and its flag words through CUDA IPC and maps its two neighbours'.  If any rank fails at any point the whole ring
// stays on NCCL send/recv (the decision is all-reduced), so this never turns a working setup into an error.
int p2p_setup(pfs_slab *s)
{
    if (s->nranks < 2 || !nccl().AllGather) return PFS_OK;
    if (const char *e = getenv("PFS_SLAB_TRANSPORT"))
        if (!strcmp(e, "nccl")) return PFS_OK;
    cudaStream_t st = nullptr;
    int ok = 1;
    PeerInfo mine;
    memset(&mine, 0, sizeof(mine));
    const size_t vbytes = 2 * (size_t)P2P_GATHER_ROWS * s->gw * sizeof(float4);
    const size_t ibytes = 2 * (size_t)P2P_GATHER_ROWS * s->iw * sizeof(float4);
    if (cudaMalloc((void **)&s->flags, P2P_FLAG_INTS * sizeof(int)) != cudaSuccess ||
        cudaMemset(s->flags, 0, P2P_FLAG_INTS * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&s->vhalo_p2p, vbytes) != cudaSuccess ||
        (s->iw > 0 && cudaMalloc((void **)&s->ihalo_p2p, ibytes) != cudaSuccess))
        ok = 0;
    if (ok && (cudaIpcGetMemHandle(&mine.planes, s->planes) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine.flags, s->flags) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine.vhalo, s->vhalo_p2p) != cudaSuccess ||
               (s->ihalo_p2p && cudaIpcGetMemHandle(&mine.ihalo, s->ihalo_p2p) != cudaSuccess)))
        ok = 0;
    (void)cudaGetLastError();
    // stamp the first word of every shared buffer, so that a neighbour can verify that the pointer it gets from
    // cudaIpcOpenMemHandle really is the start of that buffer (removed again once every rank has checked)
    const float fstamp = 1000.f + (float)s->rank;
    const int istamp = 1000 + s->rank;
    if (ok && (cudaMemcpy(s->planes, &fstamp, sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
               cudaMemcpy(s->vhalo_p2p, &fstamp, sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
               (s->ihalo_p2p && cudaMemcpy(s->ihalo_p2p, &fstamp, sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) ||
               cudaMemcpy(s->flags + 8, &istamp, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess))
        ok = 0;
    mine.plane_floats = s->plane_floats;
    mine.rows = s->rows;
    mine.irows = s->irows;
    mine.device = s->device;
    mine.has_image = s->ihalo_p2p ? 1 : 0;

    // all-gather the records (in place), then the "everything worked here" bit
    std::vector<PeerInfo> all(s->nranks);
    char *d_all = nullptr;
    PFS_CUDA(cudaMalloc((void **)&d_all, s->nranks * sizeof(PeerInfo)));
    PFS_CUDA(cudaMemcpy(d_all + (size_t)s->rank * sizeof(PeerInfo), &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice));
    ncclResult_t nr = nccl().AllGather(d_all + (size_t)s->rank * sizeof(PeerInfo), d_all, sizeof(PeerInfo), ncclUint8, s->comm, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    if (nr == ncclSuccess && ce == cudaSuccess)
        ce = cudaMemcpy(all.data(), d_all, s->nranks * sizeof(PeerInfo), cudaMemcpyDeviceToHost);
    cudaFree(d_all);
    if (nr != ncclSuccess || ce != cudaSuccess) {
        p2p_release(s);
        set_error("peer transport: exchanging IPC handles failed");
        return PFS_ECUDA;          // the communicator itself is broken: every rank sees this
    }
    auto open_link = [&](int r, pfs_slab::PeerLink *L) {
        const PeerInfo &pi = all[r];
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, pi.planes, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
        s->ipc_opened.push_back(p);
        L->planes = (float *)p;
        if (cudaIpcOpenMemHandle(&p, pi.flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
        s->ipc_opened.push_back(p);
        L->flags = (int *)p;
        if (cudaIpcOpenMemHandle(&p, pi.vhalo, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
        s->ipc_opened.push_back(p);
        L->vhalo = (float4 *)p;
        if (pi.has_image) {
            if (cudaIpcOpenMemHandle(&p, pi.ihalo, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
            s->ipc_opened.push_back(p);
            L->ihalo = (float4 *)p;
        }
        L->plane_floats = (size_t)pi.plane_floats;
        L->rows = pi.rows;
        L->irows = pi.irows;
        float f[3] = {0.f, 0.f, 1000.f + (float)r};
        int i = 0;
        if (cudaMemcpy(&f[0], L->planes, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(&f[1], L->vhalo, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
            (L->ihalo && cudaMemcpy(&f[2], L->ihalo, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            cudaMemcpy(&i, L->flags + 8, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
            return false;
        const float want = 1000.f + (float)r;
        return f[0] == want && f[1] == want && f[2] == want && i == 1000 + r;
    };
    if (ok) {
        for (int r = 0; r < s->nranks; r++)
            if (all[r].device == s->device && r != s->rank) ok = 0;      // two ranks on one GPU: nothing to gain, keep NCCL
    }
    if (ok && !open_link(s->up(), &s->link_up)) ok = 0;
    if (ok) {
        if (s->down() == s->up())
            s->link_down = s->link_up;                  // two ranks: both neighbours are the same peer, map it once
        else if (!open_link(s->down(), &s->link_down))
            ok = 0;
    }
    (void)cudaGetLastError();
    // unanimous?
    float *d_ok = s->d_scalars + 3;
    const float mine_ok = ok ? 1.f : 0.f;
    float all_ok = 0.f;
    PFS_CUDA(cudaMemcpy(d_ok, &mine_ok, sizeof(float), cudaMemcpyHostToDevice));
    PFS_NCCL(nccl().AllReduce(d_ok, d_ok, 1, ncclFloat, ncclMin, s->comm, st));
    PFS_CUDA(cudaStreamSynchronize(st));
    PFS_CUDA(cudaMemcpy(&all_ok, d_ok, sizeof(float), cudaMemcpyDeviceToHost));
    if (all_ok != 1.f) {
        p2p_release(s);
        PFS_CUDA(cudaMemset(s->planes, 0, sizeof(float)));
        return PFS_OK;
    }
    PFS_CUDA(cudaMemset(s->planes, 0, sizeof(float)));
    PFS_CUDA(cudaMemset(s->flags + 8, 0, sizeof(int)));
    s->p2p = true;
    s->xord = 0;
    return PFS_OK;
}

}  // namespace
}  // namespace pfs

extern "C" int pfs_slab_connect_nccl(pfs_slab *s, const char id[128])
{
    if (!s || !id) {
        set_error("pfs_slab_connect_nccl: null argument");
        return PFS_EINVAL;
    }
    if (!nccl().loaded) {
        set_error("pfs_slab_connect_nccl: libnccl.so.2 could not be loaded (set PFS_NCCL_LIB)");
        return PFS_ESTATE;
    }
    Guard g(s->device);
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    PFS_NCCL(nccl().CommInitRank(&s->comm, s->nranks, u, s->rank));
    s->group.clear();
    return p2p_setup(s);
}

extern "C" const char *pfs_slab_transport(const pfs_slab *s)
{
    if (!s) return "none";
    if (s->comm) return s->p2p ? "p2p" : "nccl";
    return s->group.empty() ? "unconnected" : "local";
}

namespace pfs {
namespace {
int resolve_pending_color(const std::vector<pfs_slab *> &L);      // defined with the step functions below
}
}  // namespace pfs

extern "C" int pfs_slab_check(pfs_slab *const *slabs, int n_local)
{
    std::vector<pfs_slab *> local;
    PFS_TRY(check_local("pfs_slab_check", slabs, n_local, &local));
    if (local[0]->resident) PFS_TRY(resolve_pending_color(local));   // streams: those of the last call
    for (pfs_slab *s : local) {
        Guard g(s->device);
        PFS_CUDA(cudaDeviceSynchronize());
        int flag = 0;
        PFS_CUDA(cudaMemcpy(&flag, s->d_scalars + 2, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag != 0) {
            set_error("slab rank %d: a gather left its halo (code %d) -- this is a bug in the displacement bound", s->rank, flag);
            return PFS_ESTATE;
        }
    }
    return PFS_OK;
}

// ---------------------------------------------------------------------------------------------
// Step diagnostics over the whole ring: local partial sums (kernels_basic.cu), then all-reduce across the
// ranks (NCCL sum / max, or a host fold for slabs living in this process).
// out = {||div||_2, ||p_N - p_{N-1}||_2, ||(u,v)||_2, max(|u|,|v|)} of the WHOLE grid, the same on every rank.
// ---------------------------------------------------------------------------------------------
extern "C" int pfs_slab_step_norms(pfs_slab *const *slabs, int n_local, float *const *vp, float *const *tmp,
                                   double out[4], void *const *streams)
{
    const char *fn = "pfs_slab_step_norms";
    std::vector<pfs_slab *> L;
    PFS_TRY(check_local(fn, slabs, n_local, &L));
    if (!vp || !tmp || !out) {
        set_error("%s: null argument", fn);
        return PFS_EINVAL;
    }
    constexpr int kBlocks = 1184;
    double sums[3] = {0, 0, 0}, mx = 0;
    std::vector<double *> scratch(n_local, nullptr);
    int rc = PFS_OK;
    for (int k = 0; k < n_local && rc == PFS_OK; k++) {
        pfs_slab *s = L[k];
        Guard g(s->device);
        s->stream = streams ? (cudaStream_t)streams[k] : nullptr;
        if (cudaMalloc((void **)&scratch[k], (4 * (size_t)kBlocks + 4) * sizeof(double)) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__);
            break;
        }
        double *res = scratch[k] + 4 * (size_t)kBlocks;
        rc = launch_step_norms(vp[k], tmp[k], (size_t)s->rows * s->gw, scratch[k], kBlocks, res, s->stream);
        if (rc == PFS_OK && s->comm != nullptr) {
            if (nccl().AllReduce(res, res, 3, ncclDouble, ncclSum, s->comm, s->stream) != ncclSuccess ||
                nccl().AllReduce(res + 3, res + 3, 1, ncclDouble, ncclMax, s->comm, s->stream) != ncclSuccess) {
                set_error("%s: NCCL all-reduce failed", fn);
                rc = PFS_ECUDA;
            }
        }
    }
    for (int k = 0; k < n_local; k++) {
        pfs_slab *s = L[k];
        Guard g(s->device);
        if (rc == PFS_OK && scratch[k]) {
            double host[4];
            if (cudaMemcpyAsync(host, scratch[k] + 4 * (size_t)kBlocks, sizeof(host), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
                cudaStreamSynchronize(s->stream) != cudaSuccess) {
                rc = cuda_fail(cudaGetLastError(), "norms readback", __FILE__, __LINE__);
            } else if (s->comm != nullptr) {
                sums[0] = host[0]; sums[1] = host[1]; sums[2] = host[2]; mx = host[3];
            } else {
                sums[0] += host[0]; sums[1] += host[1]; sums[2] += host[2]; mx = std::max(mx, host[3]);
            }
        }
        if (scratch[k]) cudaFree(scratch[k]);
    }
    if (rc != PFS_OK) return rc;
    out[0] = std::sqrt(sums[0]);
    out[1] = std::sqrt(sums[1]);
    out[2] = std::sqrt(sums[2]);
    out[3] = mx;
    return PFS_OK;
}

// ---------------------------------------------------------------------------------------------
// simulate_fluid_step / advect_color_step on slabs.
//
// Two ways to hold the state:
//   stateless  vp[k] / tmp[k] (image[k] / itmp[k]): the k-th local slab's band of the caller's interleaved buffers
//              (rows x gw x 4 floats on that slab's device), read and rewritten every step; pointer exchange as
//              pfs_simulate_fluid_step.
//   resident   pfs_slab_upload / pfs_slab_step / pfs_slab_download: the bands live in the slab's own planes between steps,
//              as in a pfs_ctx (pfs_ctx.cu explains the roles): advect gathers from the projected (u,v) plane, the pressure
//              warm start is a plane of the last step, project writes a (u,v) plane -- no interleaved traffic at all.
//
// Gather depths.  `exact_bound`: take the depth of the velocity advection from a max|v| reduction that the host waits for
// (one synchronisation at the start of the step).  Otherwise the depth is a guess from the previous step's maximum
// (x2, +4 rows), the advect kernel raises a flag if a departure row is missing, and the flag is read back much later --
// just before the only kernel that overwrites the step's inputs (project) -- when it has long been written.  A raised flag
// (never seen outside the tests that provoke it) discards the step's scratch results and reruns it with the exact bound.
// Resident state gives advect_color the same treatment (the stateless call must measure: its result is the caller's at
// once): its flag is looked at at the same point of the NEXT step -- or in pfs_slab_download / pfs_slab_check -- while the
// old image and the velocity it was advected through are both still intact, and a miss reruns it with a measured depth.
// ---------------------------------------------------------------------------------------------
namespace pfs {
namespace {

struct StepIO {
    bool resident;
    float **vp, **tmp;                   // stateless only
    const float *const *forces;          // optional force bands (addForces slot)
};

int min_band_rows(const pfs_slab *s, bool image)
{
    int m = image ? s->ih : s->gh;
    for (int r = 0; r < s->nranks; r++) {
        int f, c;
        band(s->gh, s->nranks, r, &f, &c);
        if (image) {
            int jf, jc;
            image_band(s->ih, s->gh, f, c, &jf, &jc);
            c = jc;
        }
        m = std::min(m, c);
    }
    return m;
}

// advect_color on the bands: image_in[k] -> image_out[k] through the velocity band vel[k] (cells `vs` floats apart).
// speculative: gather depth guessed from `vbound`, "row missing" flag raised in d_scalars[8] and left for resolve_pending_color().
int color_step(const std::vector<pfs_slab *> &L, float *const *image_in, float *const *image_out, const float *const *vel, int vs,
               float dt, bool speculative)
{
    const int n = (int)L.size();
    const int iw = L[0]->iw, ih = L[0]->ih, gw = L[0]->gw, gh = L[0]->gh;
    const float viw = (float)gw / (float)iw, vih = (float)gh / (float)ih;
    const float dt_over_viw = dt / viw, dt_over_vih = dt / vih;
    PhaseScope ph(PFS_PHASE_ADVECT_COLOR, L[0]->stream);
    const int min_rows = min_band_rows(L[0], true);
    bool whole = false;
    int D = 0;
    if (speculative) {
        const double guess = 2.0 * std::fabs((double)dt_over_vih) * (double)L[0]->vbound / (double)ih;
        if (guess * 1.001 + 5.0 < (double)min_rows) D = std::max(8, (int)std::ceil(guess * 1.001) + 4);
        if (D == 0 || D + 3 >= min_rows) speculative = false;
    }
    if (!speculative) {
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            PFS_CUDA(cudaMemsetAsync(s->d_scalars + 1, 0, sizeof(float), s->stream));
            const size_t cells = (size_t)s->rows * gw;
            if (vs == 2)
                PFS_LAUNCH(max_abs_v_kernel<2>, 592, 256, 0, s->stream, vel[k], cells, s->d_scalars + 1);
            else
                PFS_LAUNCH(max_abs_v_kernel<4>, 592, 256, 0, s->stream, vel[k], cells, s->d_scalars + 1);
        }
        float vmax = 0.f;
        PFS_TRY(global_max(L, 1, &vmax));
        const double disp = std::fabs((double)dt_over_vih) * (double)vmax / (double)ih;
        whole = !(disp * 1.001 + 3.0 < (double)min_rows);
        D = whole ? 0 : (int)std::ceil(disp * 1.001) + 2;
    }
    FieldBands F{iw, ih, 16, {}, {}, {}, true};
    for (int k = 0; k < n; k++) {
        F.band.push_back(reinterpret_cast<const char *>(image_in[k]));
        F.first.push_back(L[k]->irow0);
        F.count.push_back(L[k]->irows);
    }
    std::vector<RowSource> src;
    PFS_TRY(build_row_sources(L, F, D, whole, &src));
    for (int k = 0; k < n; k++) {
        pfs_slab *s = L[k];
        Guard g(s->device);
        int *flag = reinterpret_cast<int *>(s->d_scalars + (speculative ? 8 : 2));
        if (speculative) PFS_CUDA(cudaMemsetAsync(s->d_scalars + 8, 0, 2 * sizeof(float), s->stream));
        if (s->irows > 0) {
            dim3 block(64, 4), grid((iw + 63) / 64, (s->irows + 4 * GATHER_ROWS - 1) / (4 * GATHER_ROWS));
            float4 *out = reinterpret_cast<float4 *>(image_out[k]);
            if (vs == 2)
                PFS_LAUNCH(advect_color_slab_kernel<2>, grid, block, 0, s->stream, src[k], out, vel[k], dt_over_viw, dt_over_vih,
                           viw, vih, iw, ih, s->irow0, s->irows, gw, s->row0, s->rows, flag, 1.0f / (float)iw, 1.0f / (float)ih);
            else
                PFS_LAUNCH(advect_color_slab_kernel<4>, grid, block, 0, s->stream, src[k], out, vel[k], dt_over_viw, dt_over_vih,
                           viw, vih, iw, ih, s->irow0, s->irows, gw, s->row0, s->rows, flag, 1.0f / (float)iw, 1.0f / (float)ih);
        }
        if (speculative) {
            // "a row was missing" -> every rank, then the host; nobody waits for it yet
            PFS_LAUNCH(flag_to_float_kernel, 1, 1, 0, s->stream, reinterpret_cast<const int *>(s->d_scalars + 8), s->d_scalars + 8, nullptr);   // [9] = float([8] != 0)
            PFS_CUDA(cudaEventRecord(s->ev_fork, s->stream));
            PFS_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
            if (s->comm != nullptr)
                PFS_NCCL(nccl().AllReduce(s->d_scalars + 9, s->d_scalars + 9, 1, ncclFloat, ncclMax, s->comm, s->side));
            PFS_CUDA(cudaMemcpyAsync(s->h_scalars + 9, s->d_scalars + 9, sizeof(float), cudaMemcpyDeviceToHost, s->side));
            PFS_CUDA(cudaEventRecord(s->ev_color, s->side));
            s->pending_color = true;
            s->pending_dt = dt;
        }
    }
    return PFS_OK;
}

// Resident state: look at the flag of the last speculative advect_color; on a miss run it again with a measured depth
// (old image = img[cur^1], velocity = the projected plane: both untouched since).
int resolve_pending_color(const std::vector<pfs_slab *> &L)
{
    if (!L[0]->pending_color) return PFS_OK;
    bool missed = false;
    for (pfs_slab *s : L) {
        Guard g(s->device);
        PFS_CUDA(cudaEventSynchronize(s->ev_color));
        missed = missed || (s->h_scalars[9] != 0.f);
        s->pending_color = false;
    }
    if (!missed) return PFS_OK;
    std::vector<float *> in, out;
    std::vector<const float *> vel;
    for (pfs_slab *s : L) {
        in.push_back(s->img[s->img_cur ^ 1]);
        out.push_back(s->img[s->img_cur]);
        vel.push_back(s->interior(UV_S, 2 * (size_t)s->gw));
    }
    return color_step(L, in.data(), out.data(), vel.data(), 2, L[0]->pending_dt, false);
}

bool slab_graphs_enabled(pfs_slab *s)
{
    if (s->graph_mode < 0) {
        const char *e = getenv("PFS_SLAB_GRAPH");
        s->graph_mode = (e && e[0] == '0') ? 0 : 1;
    }
    return s->graph_mode == 1;
}

unsigned float_bits(float x)
{
    unsigned u;
    memcpy(&u, &x, sizeof(u));
    return u;
}

    // n sweeps starting from iterate 0 in plane unit pa, all in fused passes; the pass that reaches sweep n also
    // stores iterate n-1 into unit px (as pfs_api.cu's run_diffuse / run_pressure do), so both iterates the reference
    // leaves behind exist afterwards.  `valid` tracks how many halo rows of the current iterate are correct on
    // every slab: an exchange makes it `halo`; a pass of depth t needs t of them and -- by also recomputing the
    // e = valid-t rows just outside the band -- leaves e valid rows on its result.
// `diffusion`: (u,v) planes through the packed kernel; else the pressure planes.
int run_sweeps(const std::vector<pfs_slab *> &L, bool diffusion, int pa, int pb, int px, const SweepParams &proto, int count, int *last,
               int *prev, int *valid_io, const float *const *force_bands)
{
        const int n = (int)L.size();
        const int gw = L[0]->gw, halo = L[0]->halo;
        int cur = pa, oth = pb;
        int left = count;
        int valid = *valid_io;                          // valid halo rows of the iterate in `pa` on entry (0: exchange first)
        const bool vec = (gw % 4 == 0);
        SweepParams p0 = proto;
        const bool packed = diffusion && vec && packed_diffuse_supported(p0);
        const int user_depth = pfs_get_fuse_depth();
        int depth = user_depth > 0 ? std::min(user_depth, MIN_HALO) : (packed ? default_diffuse_depth() : MIN_HALO);
        if (diffusion && !packed) depth = 1;
        if (!vec) depth = 1;
        const size_t rf = diffusion ? 2 * (size_t)gw : (size_t)gw;
        bool prev_in_extra = false;
        auto one_pass = [&](int t, bool final_pass) -> int {
            if (valid < t) {
                std::vector<std::vector<float *>> pl(n);
                for (int k = 0; k < n; k++) pl[k].push_back(L[k]->plane(cur));
                PFS_TRY(exchange_planes(L, pl, halo, rf));
                valid = halo;
            }
            const int e = valid - t;                    // extra rows recomputed on each side of the band
            for (int k = 0; k < n; k++) {
                pfs_slab *s = L[k];
                Guard g(s->device);
                SweepParams p = proto;
                p.w = gw;
                p.h = s->rows + 2 * e;
                p.y_base = halo - e;
                p.wrap = 0;
                int flips = 0, wrote = 0;
                float *a = s->plane(cur), *b = s->plane(oth);
                float *x = (final_pass && t >= 2) ? s->plane(px) : nullptr;
                ForceField ff{(force_bands && final_pass) ? force_bands[k] : nullptr, e, s->rows};
                const ForceField *force = ff.aos ? &ff : nullptr;
                if (diffusion && packed && t >= 2)
                    PFS_TRY(launch_diffuse_packed(a, b, p, t, t, &flips, s->stream, x, &wrote, force));
                else if (diffusion) {
                    PFS_TRY(launch_diffuse_basic(a, b, p, 1, &flips, s->stream));
                    if (force)
                        PFS_TRY(launch_add_forces(b + (size_t)halo * rf, 2, force->aos, gw, s->rows, s->stream));
                } else if (t == 1)
                    PFS_TRY(launch_pressure_basic(a, b, s->plane(DIV), p, 1, &flips, s->stream));
                else
                    PFS_TRY(launch_pressure_fused(a, b, s->plane(DIV), p, t, t, &flips, s->stream, x, &wrote));
                if (flips != 1 || (x != nullptr && !wrote)) {
                    set_error("slab sweeps: a pass of depth %d took %d hops (previous iterate stored: %d)", t, flips, wrote);
                    return PFS_ESTATE;
                }
                if (x != nullptr) prev_in_extra = true;
            }
            valid = (force_bands && final_pass) ? 0 : e;   // rows outside the band got no force: exchange before the next use
            std::swap(cur, oth);
            return PFS_OK;
        };
        // Diffusion: pass depths as even as possible, as launch_diffuse_packed plans them (ceil(count / depth) passes of
        // depth d or d+1).  Pressure: full-depth passes, then the remainder (a depth-7 pass costs more per sweep than 8 or 4).
        const int n_passes = (count + depth - 1) / depth;
        const int base_t = count / n_passes, n_deeper = count % n_passes;
        for (int pass = 0; left > 0; pass++) {
            int t = diffusion ? std::min(left, base_t + (pass < n_deeper ? 1 : 0)) : std::min(left, depth);
            if (t < 1) t = 1;
            if (!diffusion && left - t == 1 && t >= 3) t -= 1;      // never end on a lone single sweep: it could not store iterate n-1
            PFS_TRY(one_pass(t, left - t == 0));
            left -= t;
        }
        *last = cur;
        *prev = prev_in_extra ? px : oth;               // else: the plane the last (single) sweep read
        *valid_io = valid;
        return PFS_OK;
}


int fluid_step(const std::vector<pfs_slab *> &L, const StepIO &io, float dt, float viscosity, int n_diffuse, int n_pressure,
               bool exact_bound)
{
    const int n = (int)L.size();
    const int gw = L[0]->gw, gh = L[0]->gh;
    const int halo = L[0]->halo;
    const int cf = io.resident ? 2 : 4;                  // floats per cell of the field advect gathers from
    std::vector<const float *> src_band(n);
    for (int k = 0; k < n; k++) src_band[k] = io.resident ? L[k]->interior(UV_S, 2 * (size_t)gw) : io.vp[k];

    // phase timing (pfs_phase_times) follows the first local slab's stream
    cudaStream_t ts = L[0]->stream;
    PhaseScope *ph = new PhaseScope(PFS_PHASE_ADVECT, ts);
    struct PhaseHolder {
        PhaseScope **p;
        ~PhaseHolder() { delete *p; }
    } holder{&ph};
    auto next_phase = [&](int phase) {
        delete ph;
        ph = new PhaseScope(phase, ts);
    };

    // ---- advect: displacement bound -> halo depth D of the (u,v) gather source ----
    const int min_rows = min_band_rows(L[0], false);
    static const bool spec_env = !(getenv("PFS_SLAB_SPECULATE") && !strcmp(getenv("PFS_SLAB_SPECULATE"), "0"));
    bool speculate = spec_env && !exact_bound && L[0]->vbound >= 0.f;
    bool whole = false;
    int D = 0;
    if (speculate) {
        // |dt*v/H| in rows from the last known maximum, doubled, +4 (bilinear neighbour, wrap, roundings, growth)
        const double guess = 2.0 * std::fabs((double)dt) * (double)L[0]->vbound / (double)gh;
        if (guess * 1.001 + 5.0 < (double)min_rows)
            D = std::max(8, (int)std::ceil(guess * 1.001) + 4);
        else
            speculate = false;                       // large displacements: measure first
        if (D + 3 >= min_rows) speculate = false;
    }
    if (!speculate) {
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            PFS_CUDA(cudaMemsetAsync(s->d_scalars, 0, sizeof(float), s->stream));
            const size_t cells = (size_t)s->rows * gw;
            if (cf == 2)
                PFS_LAUNCH(max_abs_v_kernel<2>, 592, 256, 0, s->stream, src_band[k], cells, s->d_scalars);
            else
                PFS_LAUNCH(max_abs_v_kernel<4>, 592, 256, 0, s->stream, src_band[k], cells, s->d_scalars);
        }
        float vmax = 0.f;
        PFS_TRY(global_max(L, 0, &vmax));
        for (int k = 0; k < n; k++) L[k]->vbound = vmax;
        // |dt*v/H| in cells, with slack for the float roundings of the kernel's own expression; +2 for the
        // bilinear neighbour and the wrap
        const double disp = std::fabs((double)dt) * (double)vmax / (double)gh;
        whole = !(disp * 1.001 + 3.0 < (double)min_rows);
        D = whole ? 0 : (int)std::ceil(disp * 1.001) + 2;
    }
    {
        FieldBands F{gw, gh, (size_t)cf * sizeof(float), {}, {}, {}, false};
        for (int k = 0; k < n; k++) {
            F.band.push_back(reinterpret_cast<const char *>(src_band[k]));
            F.first.push_back(L[k]->row0);
            F.count.push_back(L[k]->rows);
        }
        std::vector<RowSource> src;
        PFS_TRY(build_row_sources(L, F, D, whole, &src));
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            if (speculate) PFS_CUDA(cudaMemsetAsync(s->d_scalars + 4, 0, 3 * sizeof(float), s->stream));
            dim3 block(64, 4), grid((gw + 63) / 64, (s->rows + 4 * GATHER_ROWS - 1) / (4 * GATHER_ROWS));
            float2 *dst = reinterpret_cast<float2 *>(s->plane(UV_A));
            int *flag = reinterpret_cast<int *>(s->d_scalars + (speculate ? 6 : 2));
            // max|v| of the field being advected: on resident state the project kernel of the step that produced the field
            // left it in [10] (it has spare issue slots; here the reduction costs a sixth of the kernel), else a by-product
            float *vmax_out = (speculate && !io.resident) ? s->d_scalars + 4 : nullptr;
            if (cf == 2)
                PFS_LAUNCH(advect_slab_kernel<2>, grid, block, 0, s->stream, src[k], dst, dt, gw, gh, s->row0, s->rows, s->halo, flag, vmax_out, 1.0f / (float)gw, 1.0f / (float)gh);
            else
                PFS_LAUNCH(advect_slab_kernel<4>, grid, block, 0, s->stream, src[k], dst, dt, gw, gh, s->row0, s->rows, s->halo, flag, vmax_out, 1.0f / (float)gw, 1.0f / (float)gh);
            if (speculate) {
                // (max|v| of this step's input, "a departure row was missing") -> every rank, then the host; nobody waits yet
                // (on a side stream: the sweeps that follow do not depend on it)
                PFS_LAUNCH(flag_to_float_kernel, 1, 1, 0, s->stream, reinterpret_cast<const int *>(s->d_scalars + 6), s->d_scalars + 4,
                           io.resident ? s->d_scalars + 10 : nullptr);
                PFS_CUDA(cudaEventRecord(s->ev_fork, s->stream));
                PFS_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
                if (s->comm != nullptr)
                    PFS_NCCL(nccl().AllReduce(s->d_scalars + 4, s->d_scalars + 4, 2, ncclFloat, ncclMax, s->comm, s->side));
                PFS_CUDA(cudaMemcpyAsync(s->h_scalars + 4, s->d_scalars + 4, 2 * sizeof(float), cudaMemcpyDeviceToHost, s->side));
                PFS_CUDA(cudaEventRecord(s->ev_meas, s->side));
            }
        }
    }

    SweepParams dp;
    dp.w = gw;
    dp.h = L[0]->rows;
    dp.alpha = viscosity * dt;
    dp.beta = (float)(1.0 + 4.0 * (double)dp.alpha);
    int d_valid = 0;
    int dl = UV_A, dpv = UV_B;                                            // iterate n_d, iterate n_d - 1

    // pointer choreography (pfs_simulate_fluid_step): struct `vp` points at buffer Bv after diffuse (the original vp buffer for
    // an odd sweep count, else tmp's), its pressure channel is the warm start; struct `tmp` ends on the buffer holding p_N
    const bool bv_is_x = (n_diffuse & 1) != 0;
    const bool bp_is_bv = (n_pressure & 1) == 0;
    std::vector<float *> Bv(n), Bo(n), Bp(n), Bq(n);
    if (!io.resident) {
        for (int k = 0; k < n; k++) {
            Bv[k] = bv_is_x ? io.vp[k] : io.tmp[k];
            Bo[k] = bv_is_x ? io.tmp[k] : io.vp[k];
            Bp[k] = bp_is_bv ? Bv[k] : Bo[k];
            Bq[k] = bp_is_bv ? Bo[k] : Bv[k];
        }
    }
    // pressure plane units: stateless P_A (filled from the warm-start channel by the divergence kernel), P_B, P_X; resident:
    // the warm-start plane of the last step and the two others (the same on every slab)
    int p_warm = P_A, p_oth = P_B, p_ext = P_X;
    if (io.resident) {
        p_warm = bv_is_x ? L[0]->pX : L[0]->pY;
        const int units[3] = {P_A, P_B, P_X};
        int others[2], q = 0;
        for (int u : units)
            if (u != p_warm) others[q++] = u;
        p_oth = others[0];
        p_ext = others[1];
    }
    SweepParams pp;
    pp.w = gw;
    pp.h = L[0]->rows;
    pp.alpha = 1.0f;
    pp.beta = 4.0f;
    int p_valid = 0;
    int pl_last = p_warm, pl_prev = p_oth;

    // ---- the sweeps: diffusion | divergence (needs one halo row of v) + warm-start pressure, divergence halo | pressure ----
    auto sweeps_segment = [&]() -> int {
        next_phase(PFS_PHASE_DIFFUSE);
        PFS_TRY(run_sweeps(L, true, UV_A, UV_B, UV_X, dp, n_diffuse, &dl, &dpv, &d_valid, io.forces));
        next_phase(PFS_PHASE_DIVERGENCE);
        {
            std::vector<std::vector<float *>> pl(n);
            for (int k = 0; k < n; k++) pl[k].push_back(L[k]->plane(dl));
            if (d_valid < 1) PFS_TRY(exchange_planes(L, pl, 1, 2 * (size_t)gw));
            for (int k = 0; k < n; k++) {
                pfs_slab *s = L[k];
                Guard g(s->device);
                PFS_TRY(launch_divergence(s->plane(dl), s->plane(DIV), io.resident ? nullptr : Bv[k],
                                          io.resident ? nullptr : s->plane(P_A), dt, gw, s->rows, s->stream, halo, 0));
            }
            // the divergence halo and the first pressure halo (the warm start is final by now) in ONE exchange
            for (int k = 0; k < n; k++) {
                pl[k][0] = L[k]->plane(DIV);
                pl[k].push_back(L[k]->plane(p_warm));
            }
            PFS_TRY(exchange_planes(L, pl, halo, (size_t)gw));
            p_valid = halo;
        }
        next_phase(PFS_PHASE_PRESSURE);
        PFS_TRY(run_sweeps(L, false, p_warm, p_oth, p_ext, pp, n_pressure, &pl_last, &pl_prev, &p_valid, nullptr));
        return PFS_OK;
    };
    // One process per GPU over the peer transport, resident state: the segment is the same kernel sequence every step (the
    // plane roles alternate with the sweep-count parities), ~50 launches that otherwise each pay their launch latency on an
    // idle machine.  The second step with the same roles and parameters is captured, later ones replay the graph.  Every
    // rank runs the same pushes in the same order whether it replays or launches them one by one.
    pfs_slab *s0 = L[0];
    bool seg_done = false;
    if (io.resident && n == 1 && s0->p2p && !io.forces && slab_graphs_enabled(s0) && !phase_timing_on()) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        const bool capturing = cudaStreamIsCapturing(s0->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone;
        (void)cudaGetLastError();
        pfs_slab::SegGraph key = {p_warm, n_diffuse, n_pressure, pfs_get_fuse_depth(), halo, float_bits(dt), float_bits(viscosity),
                                  nullptr, 0, 0, 0, 0, 0, 0, 0};
        auto same = [](const pfs_slab::SegGraph &x, const pfs_slab::SegGraph &y) {
            return x.p_warm == y.p_warm && x.nd == y.nd && x.np == y.np && x.fuse == y.fuse && x.halo == y.halo && x.dt_bits == y.dt_bits &&
                   x.visc_bits == y.visc_bits;
        };
        pfs_slab::SegGraph *hit = nullptr;
        if (!capturing)
            for (auto &gr : s0->seg_graphs)
                if (same(gr, key)) hit = &gr;
        Guard g(s0->device);
        if (hit) {
            PFS_TRY(flush_seq(s0));
            PFS_CUDA(cudaGraphLaunch(hit->exec, s0->stream));
            g_launches += hit->launches;
            g_passes += hit->passes;
            dl = hit->dl; dpv = hit->dpv; pl_last = hit->pl_last; pl_prev = hit->pl_prev; p_valid = hit->p_valid;
            seg_done = true;
        } else if (!capturing && std::any_of(s0->seg_seen.begin(), s0->seg_seen.end(),
                                             [&](const pfs_slab::SegGraph &x) { return same(x, key); })) {
            PFS_TRY(flush_seq(s0));
            const unsigned long long l0 = g_launches, q0 = g_passes;
            cudaStream_t user = s0->stream;
            if (!s0->capture_stream) PFS_CUDA(cudaStreamCreateWithFlags(&s0->capture_stream, cudaStreamNonBlocking));
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            bool ok = cudaStreamBeginCapture(s0->capture_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                s0->stream = s0->capture_stream;
                int rc = sweeps_segment();
                if (rc == PFS_OK) rc = flush_seq(s0);
                s0->stream = user;
                ok = (cudaStreamEndCapture(s0->capture_stream, &graph) == cudaSuccess) && rc == PFS_OK && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                if (s0->seg_graphs.size() >= 8) {
                    cudaGraphExecDestroy(s0->seg_graphs.front().exec);
                    s0->seg_graphs.erase(s0->seg_graphs.begin());
                }
                key.exec = exec;
                key.launches = g_launches - l0;
                key.passes = g_passes - q0;
                key.dl = dl; key.dpv = dpv; key.pl_last = pl_last; key.pl_prev = pl_prev; key.p_valid = p_valid;
                s0->seg_graphs.push_back(key);
                PFS_CUDA(cudaGraphLaunch(exec, s0->stream));      // the launches counted during capture are this replay's
                seg_done = true;
            } else {
                (void)cudaGetLastError();                          // capture not possible here: launch one by one from now on
                g_launches = l0;
                g_passes = q0;
                s0->xord = 0;                                      // the pushes of the failed capture never ran
                s0->graph_mode = 0;
                d_valid = 0; dl = UV_A; dpv = UV_B; p_valid = 0; pl_last = p_warm; pl_prev = p_oth;
            }
        }
        if (!seg_done && !capturing) {
            if (s0->seg_seen.size() >= 16) s0->seg_seen.erase(s0->seg_seen.begin());
            s0->seg_seen.push_back(key);
        }
    }
    if (!seg_done) PFS_TRY(sweeps_segment());
    next_phase(PFS_PHASE_PROJECT);

    // ---- late checks: everything so far only wrote scratch planes ----
    if (io.resident) PFS_TRY(resolve_pending_color(L));          // the last advect_color's flag, before project overwrites its velocity
    if (speculate) {
        float vmax = 0.f;
        bool missed = false;
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            PFS_CUDA(cudaEventSynchronize(s->ev_meas));          // recorded right after advect: long done by now
            const float v = s->h_scalars[4];
            vmax = (v != v) ? INFINITY : std::max(vmax, v);
            missed = missed || (s->h_scalars[5] != 0.f);
        }
        for (int k = 0; k < n; k++) L[k]->vbound = vmax;
        if (missed) {
            // the miss was recorded in its own word ([6]); the sticky error word ([2]: gather overflow of an exact bound,
            // peer-transport timeout) is never touched here, so pfs_slab_check still sees whatever it holds
            delete ph;
            ph = nullptr;
            return fluid_step(L, io, dt, viscosity, n_diffuse, n_pressure, true);
        }
    }

    // ---- gradient subtraction (needs one halo row of p_N) + write-back ----
    {
        std::vector<std::vector<float *>> pl(n);
        for (int k = 0; k < n; k++) pl[k].push_back(L[k]->plane(pl_last));
        if (p_valid < 1) PFS_TRY(exchange_planes(L, pl, 1, (size_t)gw));
        const int uv_p = bp_is_bv ? dl : dpv;      // the (u,v) of the buffer holding p_N: diffusion iterate n if it is Bv, else n-1
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            if (io.resident) {
                PFS_CUDA(cudaMemsetAsync(s->d_scalars + 10, 0, sizeof(float), s->stream));
                PFS_TRY(launch_project_uv(s->plane(uv_p), s->plane(pl_last), s->plane(UV_S), dt, gw, s->rows, s->stream, halo, 0,
                                          s->d_scalars + 10));     // max|v| of the new field, for the next step's advect
            }
            else
                PFS_TRY(launch_project_pack(s->plane(uv_p), s->plane(pl_last), s->plane(pl_prev), s->plane(DIV), Bq[k], Bp[k], dt, gw,
                                            s->rows, s->stream, halo, 0));
        }
        if (io.resident) {
            for (pfs_slab *s : L) {
                s->uvY = uv_p;
                s->pX = pl_prev;
                s->pY = pl_last;
                s->dX = s->dY = DIV;
            }
        }
    }
    if (!io.resident) {
        for (int k = 0; k < n; k++) {
            io.vp[k] = Bq[k];
            io.tmp[k] = Bp[k];
        }
    }
    return PFS_OK;
}

int begin_call(const char *fn, pfs_slab *const *slabs, int n_local, void *const *streams, std::vector<pfs_slab *> *L)
{
    PFS_TRY(check_local(fn, slabs, n_local, L));
    for (int k = 0; k < n_local; k++) (*L)[k]->stream = streams ? (cudaStream_t)streams[k] : nullptr;
    return PFS_OK;
}

int check_counts(const char *fn, int n_diffuse, int n_pressure)
{
    if (n_diffuse < 1 || n_pressure < 1) {
        set_error("%s: sweep counts must be >= 1", fn);
        return PFS_EINVAL;
    }
    return PFS_OK;
}

int stateless_fluid_step(const char *fn, pfs_slab *const *slabs, int n_local, float **vp, float **tmp, float dt, float viscosity,
                         int n_diffuse, int n_pressure, const float *const *forces, void *const *streams)
{
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_call(fn, slabs, n_local, streams, &L));
    PFS_TRY(check_counts(fn, n_diffuse, n_pressure));
    if (!vp || !tmp) {
        set_error("%s: vp / tmp arrays are null", fn);
        return PFS_EINVAL;
    }
    for (int k = 0; k < n_local; k++) {
        if (!vp[k] || !tmp[k] || vp[k] == tmp[k] || ((uintptr_t)vp[k] & 15) || ((uintptr_t)tmp[k] & 15)) {
            set_error("%s: slab %d: vp/tmp must be distinct, non-null, 16-byte aligned device buffers", fn, k);
            return PFS_EINVAL;
        }
        if (forces && (!forces[k] || ((uintptr_t)forces[k] & 15))) {
            set_error("%s: slab %d: the force band must be a non-null, 16-byte aligned device buffer", fn, k);
            return PFS_EINVAL;
        }
    }
    return fluid_step(L, StepIO{false, vp, tmp, forces}, dt, viscosity, n_diffuse, n_pressure, false);
}

}  // namespace
}  // namespace pfs

extern "C" int pfs_slab_simulate_fluid_step(pfs_slab *const *slabs, int n_local, float **vp, float **tmp, float dt,
                                            float viscosity, int n_diffuse, int n_pressure, void *const *streams)
{
    return stateless_fluid_step("pfs_slab_simulate_fluid_step", slabs, n_local, vp, tmp, dt, viscosity, n_diffuse, n_pressure, nullptr,
                                streams);
}

// The same step with an external force at the addForces slot (pfs_simulate_fluid_step_forced): forces[k] is the k-th
// local slab's band of the interleaved force field (rows x gw x 4 floats).
extern "C" int pfs_slab_simulate_fluid_step_forced(pfs_slab *const *slabs, int n_local, float **vp, float **tmp, float dt,
                                                   float viscosity, int n_diffuse, int n_pressure,
                                                   const float *const *forces, void *const *streams)
{
    return stateless_fluid_step("pfs_slab_simulate_fluid_step_forced", slabs, n_local, vp, tmp, dt, viscosity, n_diffuse, n_pressure,
                                forces, streams);
}

// advect_color_step on slabs.  image[k] / itmp[k]: the k-th slab's band of image rows (irows x iw x 4);
// vp[k]: its band of the (already stepped) velocity field.  image[k] and itmp[k] are exchanged.
extern "C" int pfs_slab_advect_color_step(pfs_slab *const *slabs, int n_local, float **image, float **itmp,
                                          float *const *vp, float dt, void *const *streams)
{
    const char *fn = "pfs_slab_advect_color_step";
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_call(fn, slabs, n_local, streams, &L));
    if (!image || !itmp || !vp) {
        set_error("%s: null array argument", fn);
        return PFS_EINVAL;
    }
    if (L[0]->iw < 1 || L[0]->ih < 1) {
        set_error("%s: the slabs were created without an image", fn);
        return PFS_EINVAL;
    }
    for (int k = 0; k < n_local; k++) {
        if (!vp[k] || (L[k]->irows > 0 && (!image[k] || !itmp[k]))) {
            set_error("%s: slab %d: null buffer", fn, k);
            return PFS_EINVAL;
        }
    }
    PFS_TRY(color_step(L, image, itmp, vp, 4, dt, false));
    for (int k = 0; k < n_local; k++) std::swap(image[k], itmp[k]);   // fluid.cpp:317-319
    return PFS_OK;
}

// ---------------------------------------------------------------------------------------------
// Resident state: pfs_slab_upload / pfs_slab_step / pfs_slab_download
// ---------------------------------------------------------------------------------------------
extern "C" int pfs_slab_upload(pfs_slab *const *slabs, int n_local, const float *const *vp, const float *const *tmp,
                               const float *const *image, void *const *streams)
{
    const char *fn = "pfs_slab_upload";
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_call(fn, slabs, n_local, streams, &L));
    // The first upload needs everything; later ones may replace the velocity pair (vp AND tmp) or the image alone.
    const bool first = !L[0]->resident;
    if ((vp == nullptr) != (tmp == nullptr)) {
        set_error("%s: vp and tmp are replaced together", fn);
        return PFS_EINVAL;
    }
    if (first && (!vp || (L[0]->iw > 0 && !image))) {
        set_error("%s: the first upload needs vp, tmp%s", fn, L[0]->iw > 0 ? " and image" : "");
        return PFS_EINVAL;
    }
    if (!first && L[0]->pending_color) PFS_TRY(resolve_pending_color(L));
    for (int k = 0; k < n_local; k++) {
        pfs_slab *s = L[k];
        if ((vp && (!vp[k] || !tmp[k] || ((uintptr_t)vp[k] & 15) || ((uintptr_t)tmp[k] & 15))) ||
            (image && s->irows > 0 && (!image[k] || ((uintptr_t)image[k] & 15)))) {
            set_error("%s: slab %d: bands must be non-null, 16-byte aligned device buffers", fn, k);
            return PFS_EINVAL;
        }
        Guard g(s->device);
        const size_t ibytes = (size_t)s->irows * s->iw * 4 * sizeof(float);
        for (int q = 0; q < 2 && ibytes > 0; q++)
            if (!s->img[q]) PFS_CUDA(cudaMalloc((void **)&s->img[q], ibytes));
        const size_t gw = (size_t)s->gw;
        if (vp) {
            s->uvY = UV_A; s->pX = P_A; s->pY = P_B; s->dX = DIV; s->dY = DIV2;
            PFS_TRY(launch_unpack(vp[k], s->interior(UV_S, 2 * gw), s->interior(s->pX, gw), s->interior(s->dX, gw), s->gw, s->rows, s->stream));
            PFS_TRY(launch_unpack(tmp[k], s->interior(s->uvY, 2 * gw), s->interior(s->pY, gw), s->interior(s->dY, gw), s->gw, s->rows, s->stream));
            s->vbound = -1.f;             // the velocity is new: the first gather measures its bound
        }
        if (image && ibytes > 0) {
            if (first) s->img_cur = 0;
            PFS_CUDA(cudaMemcpyAsync(s->img[s->img_cur], image[k], ibytes, cudaMemcpyDefault, s->stream));
        }
        s->resident = true;
        s->pending_color = false;
    }
    return PFS_OK;
}

extern "C" int pfs_slab_download(pfs_slab *const *slabs, int n_local, float *const *vp, float *const *tmp, float *const *image,
                                 void *const *streams)
{
    const char *fn = "pfs_slab_download";
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_call(fn, slabs, n_local, streams, &L));
    for (pfs_slab *s : L) {
        if (!s->resident) {
            set_error("%s: the slabs hold no state (call pfs_slab_upload first)", fn);
            return PFS_ESTATE;
        }
    }
    PFS_TRY(resolve_pending_color(L));
    for (int k = 0; k < n_local; k++) {
        pfs_slab *s = L[k];
        Guard g(s->device);
        const size_t gw = (size_t)s->gw;
        if (vp && vp[k])
            PFS_TRY(launch_pack(vp[k], s->interior(UV_S, 2 * gw), s->interior(s->pX, gw), s->interior(s->dX, gw), s->gw, s->rows, s->stream));
        if (tmp && tmp[k])
            PFS_TRY(launch_pack(tmp[k], s->interior(s->uvY, 2 * gw), s->interior(s->pY, gw), s->interior(s->dY, gw), s->gw, s->rows, s->stream));
        if (image && image[k] && s->irows > 0)
            PFS_CUDA(cudaMemcpyAsync(image[k], s->img[s->img_cur], (size_t)s->irows * s->iw * 4 * sizeof(float), cudaMemcpyDefault, s->stream));
    }
    return PFS_OK;
}

namespace pfs {
namespace {

int begin_resident(const char *fn, pfs_slab *const *slabs, int n_local, void *const *streams, std::vector<pfs_slab *> *L)
{
    PFS_TRY(begin_call(fn, slabs, n_local, streams, L));
    for (pfs_slab *s : *L) {
        if (!s->resident) {
            set_error("%s: the slabs hold no state (call pfs_slab_upload first)", fn);
            return PFS_ESTATE;
        }
    }
    return PFS_OK;
}

int resident_color_step(const std::vector<pfs_slab *> &L, float dt)
{
    static const bool spec_env = !(getenv("PFS_SLAB_SPECULATE") && !strcmp(getenv("PFS_SLAB_SPECULATE"), "0"));
    PFS_TRY(resolve_pending_color(L));                         // two colour steps in a row: the flag words are about to be reused
    std::vector<float *> in, out;
    std::vector<const float *> vel;
    for (pfs_slab *s : L) {
        in.push_back(s->img[s->img_cur]);
        out.push_back(s->img[s->img_cur ^ 1]);
        vel.push_back(s->interior(UV_S, 2 * (size_t)s->gw));
    }
    PFS_TRY(color_step(L, in.data(), out.data(), vel.data(), 2, dt, spec_env && L[0]->vbound >= 0.f));
    for (pfs_slab *s : L) s->img_cur ^= 1;                     // fluid.cpp:317-319
    return PFS_OK;
}

}  // namespace
}  // namespace pfs

// n_steps iterations of the driver loop (main.cpp:236-239) on the resident state; and its two halves on their own
// (simulate_fluid_step / advect_color_step, fluid.hpp:107,116), e.g. to overlap transfers with one of them.
extern "C" int pfs_slab_step(pfs_slab *const *slabs, int n_local, int n_steps, float dt, float viscosity, int n_diffuse,
                             int n_pressure, void *const *streams)
{
    const char *fn = "pfs_slab_step";
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_resident(fn, slabs, n_local, streams, &L));
    PFS_TRY(check_counts(fn, n_diffuse, n_pressure));
    for (int it = 0; it < n_steps; it++) {
        PFS_TRY(fluid_step(L, StepIO{true, nullptr, nullptr, nullptr}, dt, viscosity, n_diffuse, n_pressure, false));
        if (L[0]->iw > 0) PFS_TRY(resident_color_step(L, dt));
    }
    return PFS_OK;
}

extern "C" int pfs_slab_step_fluid(pfs_slab *const *slabs, int n_local, float dt, float viscosity, int n_diffuse, int n_pressure,
                                   void *const *streams)
{
    const char *fn = "pfs_slab_step_fluid";
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_resident(fn, slabs, n_local, streams, &L));
    PFS_TRY(check_counts(fn, n_diffuse, n_pressure));
    return fluid_step(L, StepIO{true, nullptr, nullptr, nullptr}, dt, viscosity, n_diffuse, n_pressure, false);
}

extern "C" int pfs_slab_step_color(pfs_slab *const *slabs, int n_local, float dt, void *const *streams)
{
    const char *fn = "pfs_slab_step_color";
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_resident(fn, slabs, n_local, streams, &L));
    if (L[0]->iw < 1) {
        set_error("%s: the slabs were created without an image", fn);
        return PFS_EINVAL;
    }
    return resident_color_step(L, dt);
}

// ---------------------------------------------------------------------------------------------
// computePressure with a run-time sweep count on slabs (pfs_compute_pressure_adaptive over the ring).
// Batches of `check_every` fused sweeps with the usual halo exchanges; after each batch the ranks' partial sums of
// (p_N - p_{N-1})^2 (warp-shuffle + fixed-order block reduction, kernels_basic.cu) are all-reduced (NCCL, double) and the
// global rms decides -- the same decision on every rank, one host synchronisation per batch, none inside a batch.  The
// bands end up bit-identical to pfs_compute_pressure / the reference's computePressure with n_sweeps = the count returned.
// ---------------------------------------------------------------------------------------------
extern "C" int pfs_slab_compute_pressure_adaptive(pfs_slab *const *slabs, int n_local, float **vp, float **vp_out, float dt, float tol,
                                                  int max_sweeps, int check_every, int *sweeps_out, double *update_rms_out,
                                                  void *const *streams)
{
    const char *fn = "pfs_slab_compute_pressure_adaptive";
    std::vector<pfs_slab *> L;
    PFS_TRY(begin_call(fn, slabs, n_local, streams, &L));
    if (!vp || !vp_out || max_sweeps < 1 || check_every < 2 || !(tol >= 0.0f)) {
        set_error("%s: need vp, vp_out, max_sweeps >= 1, check_every >= 2 and tol >= 0", fn);
        return PFS_EINVAL;
    }
    const int n = n_local;
    for (int k = 0; k < n; k++) {
        if (!vp[k] || !vp_out[k] || vp[k] == vp_out[k] || ((uintptr_t)vp[k] & 15) || ((uintptr_t)vp_out[k] & 15)) {
            set_error("%s: slab %d: vp/vp_out must be distinct, non-null, 16-byte aligned device buffers", fn, k);
            return PFS_EINVAL;
        }
    }
    const int gw = L[0]->gw, halo = L[0]->halo;
    const size_t gws = (size_t)gw;
    constexpr int kBlocks = 1184;
    std::vector<double *> scratch(n, nullptr);
    struct Cleanup {
        std::vector<double *> &v;
        std::vector<pfs_slab *> &L;
        ~Cleanup()
        {
            for (size_t k = 0; k < v.size(); k++)
                if (v[k]) {
                    Guard g(L[k]->device);
                    cudaFree(v[k]);
                }
        }
    } cleanup{scratch, L};
    // (u,v) and the warm-start pressure of the bands -> planes; one halo row of v for the divergence; divergence + its halo
    for (int k = 0; k < n; k++) {
        pfs_slab *s = L[k];
        Guard g(s->device);
        PFS_CUDA(cudaMalloc((void **)&scratch[k], (4 * (size_t)kBlocks + 4) * sizeof(double)));
        PFS_TRY(launch_unpack(vp[k], s->interior(UV_A, 2 * gws), s->interior(P_A, gws), nullptr, gw, s->rows, s->stream));
    }
    {
        std::vector<std::vector<float *>> pl(n);
        for (int k = 0; k < n; k++) pl[k].push_back(L[k]->plane(UV_A));
        PFS_TRY(exchange_planes(L, pl, 1, 2 * gws));
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            PFS_TRY(launch_divergence(s->plane(UV_A), s->plane(DIV), nullptr, nullptr, dt, gw, s->rows, s->stream, halo, 0));
        }
        for (int k = 0; k < n; k++) pl[k][0] = L[k]->plane(DIV);
        PFS_TRY(exchange_planes(L, pl, halo, gws));
    }
    SweepParams pp;
    pp.w = gw;
    pp.h = L[0]->rows;
    pp.alpha = 1.0f;
    pp.beta = 4.0f;
    int cur = P_A, oth = P_B, last = P_A, prev = P_B, valid = 0, done = 0;
    double rms = 0.0;
    const double cells_global = (double)gw * (double)L[0]->gh;
    while (done < max_sweeps) {
        int nb = std::min(check_every, max_sweeps - done);
        if (max_sweeps - done - nb == 1) nb += 1;                 // never leave a batch of one sweep (it keeps no p_{N-1})
        PFS_TRY(run_sweeps(L, false, cur, oth, P_X, pp, nb, &last, &prev, &valid, nullptr));
        done += nb;
        double sum = 0.0;
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            double *res = scratch[k] + 4 * (size_t)kBlocks;
            PFS_TRY(launch_plane_diff_norms(s->interior(last, gws), s->interior(prev, gws), (size_t)s->rows * gws, scratch[k], kBlocks, res,
                                            s->stream));
            if (s->comm != nullptr) PFS_NCCL(nccl().AllReduce(res, res, 1, ncclDouble, ncclSum, s->comm, s->stream));
        }
        for (int k = 0; k < n; k++) {
            pfs_slab *s = L[k];
            Guard g(s->device);
            double host = 0.0;
            PFS_CUDA(cudaMemcpyAsync(&host, scratch[k] + 4 * (size_t)kBlocks, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
            PFS_CUDA(cudaStreamSynchronize(s->stream));
            sum = (s->comm != nullptr) ? host : sum + host;       // NCCL: already the global sum; in-process ring: fold the slabs
        }
        rms = std::sqrt(sum / cells_global);
        if (rms <= (double)tol) break;
        if (done < max_sweeps) {
            cur = last;
            oth = (last == P_A) ? P_B : P_A;                      // the ping-pong plane that does not hold iterate `done`
        }
    }
    // write-back as pfs_compute_pressure: divergence into channel 3 of both buffers, channel 2 <- the iterate each buffer was
    // last written with, data pointers as the reference's swaps leave them
    for (int k = 0; k < n; k++) {
        pfs_slab *s = L[k];
        Guard g(s->device);
        float *in0 = vp[k], *out0 = vp_out[k];
        float *buf_last = (done & 1) ? out0 : in0, *buf_prev = (done & 1) ? in0 : out0;
        PFS_TRY(launch_pack(buf_last, nullptr, s->interior(last, gws), s->interior(DIV, gws), gw, s->rows, s->stream));
        PFS_TRY(launch_pack(buf_prev, nullptr, done >= 2 ? s->interior(prev, gws) : nullptr, s->interior(DIV, gws), gw, s->rows, s->stream));
        vp_out[k] = buf_last;
        vp[k] = buf_prev;
    }
    if (sweeps_out) *sweeps_out = done;
    if (update_rms_out) *update_rms_out = rms;
    return PFS_OK;
}
