// sweeps_fused.cu -- temporally blocked Jacobi sweeps of the pressure (fluid.cpp:239-258): T sweeps of the
// 5-point update per launch.  (The smoothing sweeps of the velocity have their own packed-FP32 kernel, sweeps_packed.cu.)
//
// Why: one sweep is 12 B of compulsory HBM traffic per cell for 5 flops, so a sweep-per-launch kernel sits on the
// HBM roof.  Fusing T sweeps divides the traffic by ~T; the kernel then runs on the FP32 issue rate instead
// (DESIGN.md "fused sweeps").
//
// How (warp-streaming, register-resident time levels -- no __syncthreads, no shared-memory tile):
//   * one warp owns a strip of 128 columns (32 lanes x float4) and streams down the rows of its
//     chunk; each lane keeps, for every time level 0..T-1, the two most recent rows of its four
//     columns in registers (8*T registers), plus a T-row window of the divergence (4*T registers);
//   * at stream step s the fresh level-0 row s arrives (cp.async ring in shared memory, private
//     16-byte slots per lane, so no barrier is ever needed) and level l = 1..T produces row s-l from
//     rows s-l-1, s-l (registers) and s-l+1 (the row level l-1 produced a moment ago); level T's row
//     s-T is the output;
//   * left/right neighbours come from the adjacent lanes by warp shuffle; the outermost cells of
//     the strip have no valid neighbour, so validity shrinks by one cell per level from both ends:
//     the strip loads HL = 4*ceil(T/4) halo columns on each side and stores the inner 128-2*HL;
//   * vertically a chunk of L output rows streams L+2T input rows (T-row halo above and below);
//     rows and columns wrap periodically by index, so no halo copies exist on a single GPU;
//   * register windows rotate by naming: the step loop is unrolled by the rotation period.
// Per-cell arithmetic is the same device function as the one-sweep kernel (pfs_internal.cuh), so
// the result is bit-identical for every T -- tests/test_gpu_operators.py::test_fuse_depth_is_invisible.
#include <stdlib.h>

#include <utility>

#include "pfs_internal.cuh"

namespace pfs {

namespace {

constexpr int WARPS_PER_CTA = 4;
// cp.async ring depth per warp (rows, a power of two).  8 slots = 6 rows (6 KB per warp with the divergence row)
// in flight ahead of the consumer: with 4 slots the twelve warps of an SM kept ~25 KB in flight, less than HBM latency x
// bandwidth asks for, and the 100-sweep pressure solve took 0.665 ms instead of 0.628 ms at 4096^2.
#ifndef PFS_FUSED_RING_SLOTS
#define PFS_FUSED_RING_SLOTS 8
#endif
constexpr int RING_SLOTS = PFS_FUSED_RING_SLOTS;
constexpr int PREFETCH = RING_SLOTS - 2;  // rows in flight ahead of the consumer

struct FusedParams {
    const float *in;            // plane of the current iterate
    float *out;                 // plane receiving iterate +T
    float *prev;                // optional: plane receiving iterate +T-1 as well (null = not wanted)
    const float *rhs;           // divergence plane
    int w, h;
    int strip_out;              // columns stored per strip = 128 - 2*HL
    int halo_cols;              // HL
    int n_strips, n_chunks, chunk_rows;
    int y_base, wrap;           // row map of the planes (SweepParams)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}


__device__ __forceinline__ float4 update_row(const float4 &top, const float4 &cen, const float4 &bot, float left,
                                             float right, const float4 &q)
{
    float4 o;
    o.x = pressure_update(left, cen.y, top.x, bot.x, q.x);
    o.y = pressure_update(cen.x, cen.z, top.y, bot.y, q.y);
    o.z = pressure_update(cen.y, cen.w, top.z, bot.z, q.z);
    o.w = pressure_update(cen.z, right, top.w, bot.w, q.w);
    return o;
}

// Rotation period of the register windows: the 2-row level windows have period 2, the T-row
// divergence window has period T.
template <int T>
struct Unroll {
    static constexpr int value = (T % 2 == 0) ? T : 2 * T;
};

// LAST: the pass that reaches sweep n and also stores iterate n-1; the other passes carry neither the pointer nor the branch.
template <int T, int MINB, bool LAST>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB) fused_sweeps_kernel(const FusedParams P)
{
    constexpr int U = Unroll<T>::value;
    constexpr int SLOT = 2 * 32;                     // float4 per ring slot: the pressure row, then the divergence row
    // Ring addressing: with U == RING_SLOTS every slot a trip touches is an immediate offset; with 2U == RING_SLOTS the
    // trips alternate between the two halves of the ring (one pointer toggle per trip); otherwise the slot is computed.
    constexpr int MODE = (U == RING_SLOTS) ? 0 : ((2 * U == RING_SLOTS) ? 1 : 2);
    __shared__ float4 ring[WARPS_PER_CTA][RING_SLOTS * SLOT];

    pdl_launch_dependents();
    pdl_wait();                                      // the pass before this one has finished; nothing global was touched yet
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS_PER_CTA + warp;
    if (item >= P.n_strips * P.n_chunks) return;     // whole warp leaves together
    const int strip = item % P.n_strips;
    const int chunk = item / P.n_strips;
    const int w = P.w, h = P.h;

    // columns: lane owns unwrapped columns [xc, xc+4); loads wrap periodically, stores do not
    const int x0 = strip * P.strip_out;
    const int xc = x0 - P.halo_cols + 4 * lane;
    int xw = xc % w;
    if (xw < 0) xw += w;
    const bool store_lane = (xc >= x0) && (xc < x0 + P.strip_out) && (xc < w);

    // rows: chunk outputs rows [y0, y0+L); the stream starts T rows above
    const int y0 = chunk * P.chunk_rows;
    const int L = min(P.chunk_rows, h - y0);
    // plane row of the next prefetch: wraps by index on a single GPU, walks into the halo rows of a slab
    int ld_row = y0 - T;
    if (P.wrap) {
        ld_row %= h;
        if (ld_row < 0) ld_row += h;
    }
    const int wrap_at = P.wrap ? h : 0x7fffffff;
    const int n_steps = L + 2 * T;
    // running pointers, one plane row per stream step
    const long long row0 = (long long)P.y_base * w + xw;
    const float *ld_p = P.in + (row0 + (long long)ld_row * w);
    const float *ld_q = P.rhs + (row0 + (long long)ld_row * w);
    const long long cell0 = (long long)(P.y_base + y0) * w + xc;
    float *op = P.out + (cell0 - (long long)(2 * T) * w);                     // output row s - 2T (dereferenced for 0 <= row < L)
    float *pp = (LAST && P.prev) ? P.prev + (cell0 - (long long)(2 * T - 1) * w) : nullptr;   // iterate n-1: row s - 2T + 1

    float4 *const my = &ring[warp][lane];
    float4 *base = my;                               // MODE 1: the half of the ring the current trip consumes

    auto prefetch = [&](int s, float4 *dst) {
        // issue the loads of stream row s (if the stream still needs it) and commit a group either way
        if (s < n_steps) {
            cp_async16(dst, ld_p);
            cp_async16(dst + 32, ld_q);
            ld_p += w;
            ld_q += w;
            if (++ld_row == wrap_at) {
                ld_row = 0;
                ld_p = P.in + row0;
                ld_q = P.rhs + row0;
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < PREFETCH; s++) prefetch(s, my + s * SLOT);

    float4 S[T][2];     // S[l][k]: the two most recent rows of level l (k alternates with the step parity)
    float4 Q[T];        // divergence rows s-T .. s-1; row r lives in Q[r mod T]
#pragma unroll
    for (int l = 0; l < T; l++) {
        S[l][0] = S[l][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        Q[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (int sb = 0; sb < n_steps; sb += U) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s = sb + u;
            const float4 *slot;
            if constexpr (MODE == 0) {
                prefetch(s + PREFETCH, my + ((u + PREFETCH) % RING_SLOTS) * SLOT);
                slot = my + u * SLOT;
            } else if constexpr (MODE == 1) {
                // row s + PREFETCH lands (u + PREFETCH) slots after the trip's first slot: in this half, the other one, or this one again
                const int rel = u + PREFETCH;
                float4 *const other = my + ((base == my) ? U * SLOT : 0);
                prefetch(s + PREFETCH, ((rel / U) % 2 == 0 ? base : other) + (rel % U) * SLOT);
                slot = base + u * SLOT;
            } else {
                prefetch(s + PREFETCH, my + ((s + PREFETCH) & (RING_SLOTS - 1)) * SLOT);
                slot = my + (s & (RING_SLOTS - 1)) * SLOT;
            }
            cp_async_wait<PREFETCH>();                // the group of stream row s has landed
            float4 fresh = slot[0];                   // level-0 row s
            const float4 qnew = slot[32];
            const int older = u & 1;                  // which of S[l][*] holds the older row at this step
#pragma unroll
            for (int l = 1; l <= T; l++) {
                // level l, row s-l, from level l-1 rows s-l-1 (top), s-l (centre), s-l+1 (fresh)
                if (LAST && l == T && pp != nullptr) {
                    // `fresh` is row s-(T-1) of level T-1 = output row y0 + (s - 2T + 1) of the previous iterate;
                    // its valid columns include every column this lane stores
                    if (store_lane && (unsigned)(s - 2 * T + 1) < (unsigned)L) *reinterpret_cast<float4 *>(pp) = fresh;
                }
                const float4 top = S[l - 1][older];
                const float4 cen = S[l - 1][older ^ 1];
                const float left = __shfl_up_sync(0xffffffffu, cen.w, 1);
                const float right = __shfl_down_sync(0xffffffffu, cen.x, 1);
                const float4 q = Q[(u - l + 2 * U * T) % T];
                const float4 o = update_row(top, cen, fresh, left, right, q);
                S[l - 1][older] = fresh;              // level l-1 now holds rows s-l, s-l+1
                fresh = o;
            }
            Q[u % T] = qnew;                          // row s replaces row s-T
            // fresh = level T, row s-T of the stream = output row y0 + (s - 2T)
            if (store_lane && (unsigned)(s - 2 * T) < (unsigned)L) *reinterpret_cast<float4 *>(op) = fresh;
            op += w;
            if constexpr (LAST) {
                if (pp != nullptr) pp += w;
            }
        }
        if constexpr (MODE == 1) base = my + ((base == my) ? U * SLOT : 0);
    }
    cp_async_wait<0>();
}

template <int T>
int launch_one(const FusedParams &P, cudaStream_t s)
{
    const int total = P.n_strips * P.n_chunks;
    const unsigned blocks = (unsigned)((total + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    constexpr int MINB = (T >= 6) ? 3 : 4;
    if (P.prev != nullptr)
        PFS_LAUNCH_PDL((fused_sweeps_kernel<T, MINB, true>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    else
        PFS_LAUNCH_PDL((fused_sweeps_kernel<T, MINB, false>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    return PFS_OK;
}

int launch_depth(int depth, const FusedParams &P, cudaStream_t s)
{
    switch (depth) {
    case 2: return launch_one<2>(P, s);
    case 3: return launch_one<3>(P, s);
    case 4: return launch_one<4>(P, s);
    case 5: return launch_one<5>(P, s);
    case 6: return launch_one<6>(P, s);
    case 7: return launch_one<7>(P, s);
    case 8: return launch_one<8>(P, s);
    default: set_error("fused sweeps: unsupported depth %d", depth); return PFS_EINVAL;
    }
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

}  // namespace

bool fused_sweeps_supported(int w, int h) { return (w % 4 == 0) && w >= 4 && h >= 1; }

constexpr int MAX_FUSE_DEPTH = 8;

int launch_pressure_fused(float *a, float *b, const float *rhs, const SweepParams &p, int n, int depth, int *flips,
                          cudaStream_t s, float *prev, int *prev_written)
{
    if (prev_written) *prev_written = 0;
    if (!fused_sweeps_supported(p.w, p.h)) {
        set_error("fused sweeps need a width that is a multiple of 4 (got %d)", p.w);
        return PFS_EINVAL;
    }
    static const int env_depth = env_int("PFS_FUSE_DEPTH", 0);
    static const int env_rows = env_int("PFS_CHUNK_ROWS", 0);
    if (depth <= 0) depth = env_depth > 0 ? env_depth : MAX_FUSE_DEPTH;
    if (depth > MAX_FUSE_DEPTH) depth = MAX_FUSE_DEPTH;
    int hops = 0;
    float *cur = a, *oth = b;
    int left = n;
    while (left > 0) {
        int t = (left >= depth) ? depth : left;
        if (left - t == 1 && t >= 3) t -= 1;       // never leave a lone last sweep: it could not store iterate n-1
        if (depth < 2 || t < 2) {                  // a single remaining sweep (or depth 1): one plain sweep
            int one = 0;
            PFS_TRY(launch_pressure_basic(cur, oth, rhs, p, 1, &one, s));
            t = 1;
        } else {
            FusedParams P;
            P.in = cur; P.out = oth; P.rhs = rhs;
            const bool last_pass = (left - t == 0) && prev != nullptr;
            P.prev = last_pass ? prev : nullptr;
            if (last_pass && prev_written) *prev_written = 1;
            P.w = p.w; P.h = p.h; P.y_base = p.y_base; P.wrap = p.wrap;
            P.halo_cols = 4 * ((t + 3) / 4);
            P.strip_out = 128 - 2 * P.halo_cols;
            P.n_strips = (p.w + P.strip_out - 1) / P.strip_out;
            // chunk height: enough chunks to fill the machine, tall enough to amortise the 2T halo rows
            const long long slots = (long long)sm_count() * 12;    // resident warps at 12 warps per SM (168 registers)
            const int rows = pick_chunk_rows(p.h, P.n_strips, slots, env_rows);
            P.chunk_rows = rows;
            P.n_chunks = (p.h + rows - 1) / rows;
            PFS_TRY(launch_depth(t, P, s));
        }
        std::swap(cur, oth);
        hops++;
        left -= t;
    }
    *flips = hops;
    return PFS_OK;
}

}  // namespace pfs
