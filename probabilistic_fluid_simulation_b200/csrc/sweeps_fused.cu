// sweeps_fused.cu -- temporally blocked Jacobi sweeps: T sweeps of the 5-point update per launch.
//
// Why: one sweep is 12 B (pressure) / 16 B (diffusion, u and v) of compulsory HBM traffic per cell
// for 5 / ~17 flops, so a sweep-per-launch kernel sits on the HBM roof.  Fusing T sweeps divides the
// traffic by ~T; the kernel then runs on the FP32 issue rate instead (DESIGN.md "fused sweeps").
//
// How (warp-streaming, register-resident time levels -- no __syncthreads, no shared-memory tile):
//   * one warp owns a strip of 128 columns (32 lanes x float4) and streams down the rows of its
//     chunk; each lane keeps, for every time level 0..T-1, the two most recent rows of its four
//     columns in registers (8*T registers), plus -- pressure only -- a T-row window of the
//     divergence (4*T registers);
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
    const float *in0, *in1;     // plane(s) of the current iterate (in1: diffusion's second plane)
    float *out0, *out1;         // plane(s) receiving iterate +T
    float *prev0, *prev1;       // optional: plane(s) receiving iterate +T-1 as well (null = not wanted)
    const float *rhs;           // divergence plane (pressure) or null
    int w, h;
    int strip_out;              // columns stored per strip = 128 - 2*HL
    int halo_cols;              // HL
    int n_strips, n_chunks, chunk_rows, n_planes;
    int y_base, wrap;           // row map of the planes (SweepParams)
    float alpha, beta;
    float rbeta;                // RN(1/beta), binary32
    float div_lo, div_hi;       // |numerator| range in which the FMA division is exact (lo = +inf disables it)
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


// ---------------------------------------------------------------------------------------------
// Correctly rounded division by a loop-invariant constant, 3 instructions instead of ~25.
//
//   y  = RN(1/b)            (host, binary32)
//   q0 = RN(a*y);  e = a - b*q0 (exact, one FMA);  q1 = RN(q0 + e*y)  ==  RN(a/b)
//
// Why q1 is the IEEE quotient: q0 + e*y = a/b + eps*(a/b - q0) with |eps| <= b*2^-25 (b in [1,2),
// y correctly rounded) and |a/b - q0| < 1.5 ulp, so the FMA rounds a value within 0.75*b^2*2^-47
// (< 3 units of 2^-47, relative to a quotient in [1,2)) of the true quotient.  That can only differ
// from RN(a/b) if a rounding midpoint m lies in between, i.e. |A - B*M| <= 2 for the integer
// significands A, B of a, b and the odd 25-bit M of m.  tests/exact_div_check.c enumerates EVERY
// such (a, b) pair with |A - B*M| <= 4 over all 2^23 significands B (23.3 M quotients) and finds no
// mismatch; tests/test_exact_division.py runs it, plus 10^8 random and near-midpoint quotients.
// Preconditions: no underflow in e (|a| >= 2^-96 keeps every bit of e above 2^-149), no overflow
// (|a| <= 2^96, 2^-20 <= b <= 2^20), a != +-0 (the FMA would turn -0 into +0).  Anything outside
// the guard range [div_lo, div_hi] takes __fdiv_rn.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float div_const_fast(float a, float b, float y)
{
    const float q0 = __fmul_rn(a, y);
    const float e = __fmaf_rn(-b, q0, a);
    return __fmaf_rn(e, y, q0);
}

// Out-of-line slow path (kept out of the unrolled loop body: it is essentially never executed).
__device__ __noinline__ float4 div_row_ieee(float sx, float sy, float sz, float sw, float beta)
{
    return make_float4(__fdiv_rn(sx, beta), __fdiv_rn(sy, beta), __fdiv_rn(sz, beta), __fdiv_rn(sw, beta));
}

__device__ __forceinline__ float diffuse_numerator(float l, float r, float t, float b, float c, float alpha)
{
    // fluid.cpp:175-182 numerator: (((alpha*L + alpha*R) + alpha*T) + alpha*B) + 1.0f*u_n
    float s = __fadd_rn(__fmul_rn(alpha, l), __fmul_rn(alpha, r));
    s = __fadd_rn(s, __fmul_rn(alpha, t));
    s = __fadd_rn(s, __fmul_rn(alpha, b));
    return __fadd_rn(s, c);
}

template <int OP>
__device__ __forceinline__ float4 update_row(const float4 &top, const float4 &cen, const float4 &bot, float left,
                                             float right, const float4 &q, const FusedParams &P)
{
    float4 o;
    if constexpr (OP == SWEEP_PRESSURE) {
        o.x = pressure_update(left, cen.y, top.x, bot.x, q.x);
        o.y = pressure_update(cen.x, cen.z, top.y, bot.y, q.y);
        o.z = pressure_update(cen.y, cen.w, top.z, bot.z, q.z);
        o.w = pressure_update(cen.z, right, top.w, bot.w, q.w);
    } else {
        const float alpha = P.alpha, beta = P.beta, y = P.rbeta, lo = P.div_lo, hi = P.div_hi;
        const float sx = diffuse_numerator(left, cen.y, top.x, bot.x, cen.x, alpha);
        const float sy = diffuse_numerator(cen.x, cen.z, top.y, bot.y, cen.y, alpha);
        const float sz = diffuse_numerator(cen.y, cen.w, top.z, bot.z, cen.z, alpha);
        const float sw = diffuse_numerator(cen.z, right, top.w, bot.w, cen.w, alpha);
        const bool fast = (fabsf(sx) >= lo) && (fabsf(sx) <= hi) && (fabsf(sy) >= lo) && (fabsf(sy) <= hi) &&
                          (fabsf(sz) >= lo) && (fabsf(sz) <= hi) && (fabsf(sw) >= lo) && (fabsf(sw) <= hi);
        if (fast) {
            o.x = div_const_fast(sx, beta, y);
            o.y = div_const_fast(sy, beta, y);
            o.z = div_const_fast(sz, beta, y);
            o.w = div_const_fast(sw, beta, y);
        } else {                                  // zeros, denormal-range, huge or non-finite numerators
            o = div_row_ieee(sx, sy, sz, sw, beta);
        }
    }
    return o;
}

// Rotation period of the register windows: the 2-row level windows have period 2, the T-row
// divergence window has period T.
template <int OP, int T>
struct Unroll {
    static constexpr int value = (OP == SWEEP_PRESSURE) ? ((T % 2 == 0) ? T : 2 * T) : 2;
};

template <int OP, int T, int MINB>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB) fused_sweeps_kernel(const FusedParams P)
{
    constexpr int U = Unroll<OP, T>::value;
    constexpr int ROWS_PER_SLOT = (OP == SWEEP_PRESSURE) ? 2 : 1;   // p (+ divergence) per ring slot
    __shared__ float4 ring[WARPS_PER_CTA][RING_SLOTS][ROWS_PER_SLOT][32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int item = blockIdx.x * WARPS_PER_CTA + warp;
    const int total = P.n_strips * P.n_chunks * P.n_planes;
    if (item >= total) return;                       // whole warp leaves together
    const int strip = item % P.n_strips;
    item /= P.n_strips;
    const int chunk = item % P.n_chunks;
    const int plane = item / P.n_chunks;

    const float *__restrict__ in = plane ? P.in1 : P.in0;
    float *__restrict__ out = plane ? P.out1 : P.out0;
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
    ld_row += P.y_base;
    const int wrap_at = P.wrap ? h : 0x7fffffff;
    const int n_steps = L + 2 * T;

    float4 *my = &ring[warp][0][0][lane];
    constexpr int SLOT_STRIDE = ROWS_PER_SLOT * 32;   // in float4 units

    auto prefetch = [&](int s) {
        // issue the loads of stream row s (if the stream still needs it) and commit a group either way
        if (s < n_steps) {
            float4 *dst = my + (s & (RING_SLOTS - 1)) * SLOT_STRIDE;
            const size_t off = (size_t)ld_row * w + xw;
            cp_async16(dst, in + off);
            if constexpr (OP == SWEEP_PRESSURE) cp_async16(dst + 32, P.rhs + off);
            ld_row = (ld_row + 1 == wrap_at) ? 0 : ld_row + 1;
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < PREFETCH; s++) prefetch(s);

    float4 S[T][2];     // S[l][k]: the two most recent rows of level l (k alternates with the step parity)
    float4 Q[T];        // divergence rows s-T .. s-1 (pressure only); row r lives in Q[r mod T]
#pragma unroll
    for (int l = 0; l < T; l++) {
        S[l][0] = S[l][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        Q[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    float *out_ptr = out + (size_t)(P.y_base + y0) * w + xc;   // row y0 of this lane's columns (store lanes only)
    float *prev = plane ? P.prev1 : P.prev0;                   // the reference keeps iterate n-1 in its other buffer
    float *prev_ptr = prev ? prev + (size_t)(P.y_base + y0) * w + xc : nullptr;

    for (int sb = 0; sb < n_steps; sb += U) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s = sb + u;
            prefetch(s + PREFETCH);
            cp_async_wait<PREFETCH>();                // the group of stream row s has landed
            const float4 *slot = my + (s & (RING_SLOTS - 1)) * SLOT_STRIDE;
            float4 fresh = slot[0];                   // level-0 row s
            float4 qnew = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (OP == SWEEP_PRESSURE) qnew = slot[32];
            const int older = u & 1;                  // which of S[l][*] holds the older row at this step
#pragma unroll
            for (int l = 1; l <= T; l++) {
                // level l, row s-l, from level l-1 rows s-l-1 (top), s-l (centre), s-l+1 (fresh)
                if (l == T && prev_ptr != nullptr) {
                    // `fresh` is row s-(T-1) of level T-1 = output row y0 + (s - 2T + 1) of the previous iterate;
                    // its valid columns include every column this lane stores
                    const int prow = s - 2 * T + 1;
                    if (store_lane && prow >= 0 && prow < L)
                        *reinterpret_cast<float4 *>(prev_ptr + (size_t)prow * w) = fresh;
                }
                const float4 top = S[l - 1][older];
                const float4 cen = S[l - 1][older ^ 1];
                const float left = __shfl_up_sync(0xffffffffu, cen.w, 1);
                const float right = __shfl_down_sync(0xffffffffu, cen.x, 1);
                const float4 q = Q[(u - l + 2 * U * T) % T];
                const float4 o = update_row<OP>(top, cen, fresh, left, right, q, P);
                S[l - 1][older] = fresh;              // level l-1 now holds rows s-l, s-l+1
                fresh = o;
            }
            if constexpr (OP == SWEEP_PRESSURE) Q[u % T] = qnew;   // row s replaces row s-T
            // fresh = level T, row s-T of the stream = output row y0 + (s - 2T)
            const int orow = s - 2 * T;
            if (store_lane && orow >= 0 && orow < L)
                *reinterpret_cast<float4 *>(out_ptr + (size_t)orow * w) = fresh;
        }
    }
    cp_async_wait<0>();
}

template <int OP, int T>
int launch_one(const FusedParams &P, cudaStream_t s)
{
    const int total = P.n_strips * P.n_chunks * P.n_planes;
    const unsigned blocks = (unsigned)((total + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    constexpr int MINB = (T >= 6) ? 3 : 4;
    PFS_LAUNCH((fused_sweeps_kernel<OP, T, MINB>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    return PFS_OK;
}

template <int OP>
int launch_depth(int depth, const FusedParams &P, cudaStream_t s)
{
    switch (depth) {
    case 2: return launch_one<OP, 2>(P, s);
    case 3: return launch_one<OP, 3>(P, s);
    case 4: return launch_one<OP, 4>(P, s);
    case 5: return launch_one<OP, 5>(P, s);
    case 6: return launch_one<OP, 6>(P, s);
    case 7: return launch_one<OP, 7>(P, s);
    case 8: return launch_one<OP, 8>(P, s);
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

int launch_sweeps_fused(SweepOp op, float *a0, float *a1, float *b0, float *b1, const float *rhs,
                        const SweepParams &p, int n, int depth, int *flips, cudaStream_t s, float *prev0, float *prev1,
                        int *prev_written)
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
    float *cur0 = a0, *cur1 = a1, *oth0 = b0, *oth1 = b1;
    int left = n;
    while (left > 0) {
        int t = (left >= depth) ? depth : left;
        if (depth < 2 || t < 2) {                  // a single remaining sweep (or depth 1): one plain sweep
            int one = 0;
            PFS_TRY(launch_sweeps_basic(op, cur0, cur1, oth0, oth1, rhs, p, 1, &one, s));
            t = 1;
        } else {
            FusedParams P;
            P.in0 = cur0; P.in1 = cur1; P.out0 = oth0; P.out1 = oth1; P.rhs = rhs;
            const bool last_pass = (left - t == 0) && prev0 != nullptr;
            P.prev0 = last_pass ? prev0 : nullptr;
            P.prev1 = last_pass ? prev1 : nullptr;
            if (last_pass && prev_written) *prev_written = 1;
            P.w = p.w; P.h = p.h; P.y_base = p.y_base; P.wrap = p.wrap;
            P.halo_cols = 4 * ((t + 3) / 4);
            P.strip_out = 128 - 2 * P.halo_cols;
            P.n_strips = (p.w + P.strip_out - 1) / P.strip_out;
            P.n_planes = (op == SWEEP_DIFFUSE) ? 2 : 1;
            // chunk height: enough chunks to fill the machine, tall enough to amortise the 2T halo rows
            const long long slots = (long long)sm_count() * 12;    // resident warps at 12 warps per SM (168 registers)
            const int rows = pick_chunk_rows(p.h, P.n_strips * P.n_planes, slots, env_rows);
            P.chunk_rows = rows;
            P.n_chunks = (p.h + rows - 1) / rows;
            P.alpha = p.alpha; P.beta = p.beta;
            P.rbeta = 1.0f / p.beta;
            const bool fast_div = (p.beta >= 0x1p-20f) && (p.beta <= 0x1p20f);   // false for NaN too
            P.div_lo = fast_div ? 0x1p-96f : __builtin_inff();
            P.div_hi = 0x1p96f;
            if (op == SWEEP_PRESSURE)
                PFS_TRY(launch_depth<SWEEP_PRESSURE>(t, P, s));
            else
                PFS_TRY(launch_depth<SWEEP_DIFFUSE>(t, P, s));
        }
        float *t0 = cur0, *t1 = cur1;
        cur0 = oth0; cur1 = oth1; oth0 = t0; oth1 = t1;
        hops++;
        left -= t;
    }
    *flips = hops;
    return PFS_OK;
}

}  // namespace pfs
