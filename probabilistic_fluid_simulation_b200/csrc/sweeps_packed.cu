// sweeps_packed.cu -- temporally blocked smoothing ("diffusion") sweeps on Blackwell's packed-FP32
// pipe: T sweeps of fluid.cpp:154-186 per launch, u and v advanced together.
//
// The diffusion operator applies the same 5-point update with the same coefficients to channel 0
// (u) and channel 1 (v).  sm_100 has two-wide FP32 instructions (add/mul/fma.rn.f32x2 -> FADD2 /
// FMUL2 / FFMA2) that round each half exactly like the scalar instruction, so a lane keeps (u, v)
// of a cell in one 64-bit register pair and issues ONE instruction per pair of updates.  The
// structure is the warp-streaming scheme of sweeps_fused.cu:
//   * a warp owns a strip of 128 columns x a chunk of L rows of BOTH planes; each lane holds, for
//     every time level 0..T-1, the two most recent rows of its 4 cells x (u,v) (16*T registers);
//   * per stream step one new row of u and of v arrives through a private cp.async ring in shared
//     memory (no barriers anywhere), level l = 1..T produces row s-l, level T is stored;
//   * x neighbours across lanes come by warp shuffle of the already-multiplied alpha*value pairs;
//   * the division by beta = 1 + 4*alpha is the 3-instruction correctly rounded FMA division of
//     div_const_fast() below, applied to pairs.  Its preconditions (numerator magnitude in
//     [2^-96, 2^96], not -0) are not branched on in the hot loop: the loop keeps a running FMNMX3
//     minimum of |numerator| and maximum of |input|; a warp whose extremes leave the safe range
//     raises a flag for its work item, and a second ("repair") launch recomputes exactly those
//     items with __fdiv_rn.  Real velocity fields never raise the flag; exact-zero regions do, and
//     stay correct.
// Results are bit-identical to the one-sweep kernel and to fluid.cpp for every T.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <algorithm>
#include <cmath>
#include <utility>

#include "pfs_internal.cuh"

namespace pfs {

int pick_chunk_rows(int h, int columns_of_items, long long slots, int forced_rows);

namespace {

int env_int(const char *name, int dflt);

constexpr int WARPS_PER_CTA = 4;
#ifndef PFS_CARRY_MAX_DEPTH
#define PFS_CARRY_MAX_DEPTH 6
#endif
constexpr int CARRY_MAX_DEPTH = PFS_CARRY_MAX_DEPTH;
#ifndef PFS_RING_SLOTS
#define PFS_RING_SLOTS 4
#endif
constexpr int RING_SLOTS = PFS_RING_SLOTS;   // cp.async ring depth per warp (rows), a power of two
constexpr int PREFETCH = RING_SLOTS - 2;      // rows in flight ahead of the consumer
constexpr int PRING_SLOTS = 4;                // the (opt-in) packed pressure kernel keeps its 4-slot ring (64 B x 32 lanes per slot)
constexpr int PPREFETCH = PRING_SLOTS - 2;

struct PackedParams {
    const float *in_u, *in_v;
    float *out_u, *out_v;
    float *prev_u, *prev_v;     // optional: planes receiving iterate +T-1 as well (null = not wanted)
    int *flags;                 // one int per work item (main kernel raises, repair kernel consumes)
    int w, h;
    int strip_out, halo_cols;
    int n_strips, n_chunks, chunk_rows;
    int y_base, wrap;           // row map of the planes (SweepParams)
    float alpha, beta, rbeta;
    float zh, zl;               // 2-instruction division (div_const_two2): RN(1/beta) and RN(1/beta - zh); used iff div2
    int div2;
    float guard_lo, guard_hi_in;   // |numerator| >= guard_lo and |input| <= guard_hi_in keep the FMA division exact
    float neg_zero;             // -0.0f, deliberately a RUN-TIME value: see mulc2()
};

// ---- packed binary32 x2 arithmetic (each half rounded to nearest-even like the scalar op) ------
__device__ __forceinline__ unsigned long long pk(float2 a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 upk(unsigned long long r)
{
    float2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
    return upk(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
    return upk(d);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
    return upk(d);
}
// alpha * x as RN(x*alpha + (-0)): identical to mul.rn for every input (adding -0 never changes a
// value, and (+0)+(-0) = +0, (-0)+(-0) = -0 keep the product's zero sign).  Why not mul.rn.f32x2:
// ptxas 12.9 contracts an explicit mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 -- even with
// --fmad=false, and unlike the scalar mul.rn/add.rn pair -- which would round once instead of twice
// and break parity with fluid.cpp.  An FMA whose addend is a kernel parameter cannot be simplified
// back into a multiply, and an FMA result cannot be contracted into a following add.
__device__ __forceinline__ float2 mulc2(float2 x, float2 c, float2 neg_zero) { return fma2(x, c, neg_zero); }

__device__ __forceinline__ float min3abs(float m, float a, float b)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(m), "f"(fabsf(a)), "f"(fabsf(b)));
    return d;
}
__device__ __forceinline__ float max3abs(float m, float a, float b)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(m), "f"(fabsf(a)), "f"(fabsf(b)));
    return d;
}
// Lane shuffles with immediate operands (inline PTX): the intrinsic form made ptxas keep the lane delta and clamp in
// registers and funnel every shuffle through one fixed source/destination register pair, i.e. two MOVs per SHFL.
__device__ __forceinline__ float shfl_up1(float v)
{
    float d;
    asm volatile("shfl.sync.up.b32 %0, %1, 1, 0, 0xffffffff;" : "=f"(d) : "f"(v));
    return d;
}
__device__ __forceinline__ float shfl_down1(float v)
{
    float d;
    asm volatile("shfl.sync.down.b32 %0, %1, 1, 31, 0xffffffff;" : "=f"(d) : "f"(v));
    return d;
}
__device__ __forceinline__ float2 shfl_up2(float2 v) { return make_float2(shfl_up1(v.x), shfl_up1(v.y)); }
__device__ __forceinline__ float2 shfl_down2(float2 v) { return make_float2(shfl_down1(v.x), shfl_down1(v.y)); }

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
// Correctly rounded division by a loop-invariant constant, 3 instructions (per PAIR here):
//   y = RN(1/b) (host, binary32);  q0 = RN(a*y);  e = a - b*q0 (exact in one FMA);  q1 = RN(q0 + e*y)
// q1 == RN(a/b): q0 + e*y = a/b + eps*(a/b - q0) with |eps| <= b*2^-25 (b scaled to [1,2)) and
// |a/b - q0| < 1.5 ulp, i.e. the FMA rounds a value within 0.75*b^2 (< 3) units of 2^-47 of the true
// quotient (quotient scaled to [1,2)).  A different rounding needs a rounding midpoint m in that
// gap, i.e. |A - B*M| <= 2 for the integer significands A, B and the odd 25-bit M of m.
// tests/exact_div_check.c enumerates EVERY (a, b) with |A - B*M| <= 4 over all 2^23 significands B
// (23.3 M quotients): no mismatch.  Preconditions: e must not lose bits to underflow
// (|a| >= 2^-96), nothing overflows (|a| <= 2^96, 2^-20 <= b <= 2^20) and a is not -0 (the final FMA
// would return +0).  +0 is fine.  The guard below keeps the kernel inside them.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 div_const_fast2(float2 a, float2 negb, float2 y)
{
    const float2 q0 = mul2(a, y);
    const float2 e = fma2(negb, q0, a);
    return fma2(e, y, q0);
}

// Two-instruction variant (Brisebarre & Muller, "Correctly rounded multiplication by arbitrary precision constants"):
// with C = 1/b split as zh = RN(C), zl = RN(C - zh),  q = RN(a*zh + RN(a*zl)).  For most divisors this is RN(a/b)
// for EVERY a, for some it is wrong for a few a -- so it is only used for a divisor after div2_constants() (below, host)
// has tried all 2^23 significands of a against the true quotient; scaling a by a power of two scales every
// intermediate exactly as long as a*zl stays normal, which the guard on |numerator| ensures.
__device__ __forceinline__ float2 div_const_two2(float2 a, float2 zh, float2 zl)
{
    return fma2(a, zh, mul2(a, zl));
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}

// NC = cells per lane (4: 128-column strips, 16*T registers of state; 2: 64-column strips, 8*T registers,
// i.e. twice the resident warps at a 7 % wider relative halo -- which one is faster is measured, not assumed:
// profiles/r01_tuning.md).
// UNR = stream steps per trip of the main loop (even: the row slots alternate with the step parity).  Steps past the
// last one are harmless -- their prefetch is skipped and every store is masked by its row range -- so the trip count is
// simply rounded up.  2 is the shipped value; 4 is an opt-in (PFS_DIFFUSE_UNROLL=4) that lets ptxas drop a third of the
// register moves at the loop back-edge (13.4 instead of 14.4 instructions per update, scripts/sass_loop_stats.py) at
// twice the loop body (20 KB); not measured on a GPU yet.
template <int T, bool EXACT, int MINB, int NC, bool DIV2 = false, int UNR = 2>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB) diffuse_packed_kernel(const PackedParams P)
{
    static_assert(NC == 4 || NC == 2, "4 or 2 cells per lane");
    static_assert(UNR >= 2 && UNR % 2 == 0, "the unroll factor must be even");
    __shared__ __align__(16) float ring[WARPS_PER_CTA][RING_SLOTS][2][32 * NC];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS_PER_CTA + warp;
    if (item >= P.n_strips * P.n_chunks) return;          // whole warp leaves together
    if constexpr (EXACT) {
        // repair launch: only flagged items.  It also clears the flag, so the buffer is all zero again when
        // the next pass starts (no memset launch per pass).
        int f = 0;
        if (lane == 0) {
            f = P.flags[item];
            if (f) P.flags[item] = 0;
        }
        if (__shfl_sync(0xffffffffu, f, 0) == 0) return;
    }
    const int strip = item % P.n_strips;
    const int chunk = item / P.n_strips;
    const int w = P.w, h = P.h;

    const int x0 = strip * P.strip_out;
    const int xc = x0 - P.halo_cols + NC * lane;          // unwrapped first column of this lane
    int xw = xc % w;
    if (xw < 0) xw += w;
    const bool store_lane = (xc >= x0) && (xc < x0 + P.strip_out) && (xc < w);

    const int y0 = chunk * P.chunk_rows;
    const int L = min(P.chunk_rows, h - y0);
    int ld_row = y0 - T;
    if (P.wrap) {
        ld_row %= h;
        if (ld_row < 0) ld_row += h;
    }
    ld_row += P.y_base;
    const int wrap_at = P.wrap ? h : 0x7fffffff;
    const int n_steps = L + 2 * T;

    float *my = &ring[warp][0][0][lane * NC];
    constexpr int PLANE_STRIDE = 32 * NC;                 // floats between the u and the v row of a slot
    constexpr int SLOT_STRIDE = 2 * PLANE_STRIDE;

    auto prefetch = [&](int s) {
        if (s < n_steps) {
            float *dst = my + (s & (RING_SLOTS - 1)) * SLOT_STRIDE;
            const size_t off = (size_t)ld_row * w + xw;
            if constexpr (NC == 4) {
                cp_async16(dst, P.in_u + off);
                cp_async16(dst + PLANE_STRIDE, P.in_v + off);
            } else {
                cp_async8(dst, P.in_u + off);
                cp_async8(dst + PLANE_STRIDE, P.in_v + off);
            }
            ld_row = (ld_row + 1 == wrap_at) ? 0 : ld_row + 1;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < PREFETCH; s++) prefetch(s);

    // S[l][k][c]: level l, k alternates with the step parity, c = cell; .x = u, .y = v.
    // Initialised to 1 (not 0) so that warm-up garbage never looks like a zero numerator.
    float2 S[T][2][NC];
#pragma unroll
    for (int l = 0; l < T; l++)
#pragma unroll
        for (int c = 0; c < NC; c++) S[l][0][c] = S[l][1][c] = make_float2(1.f, 1.f);
    // Passes up to CARRY_MAX_DEPTH have registers to spare for a third row per level, so that the alpha product of a
    // row, formed when the row arrives as the bottom row, is kept while the row is the centre and then the top row:
    // every value is multiplied by alpha exactly once per level (the reference multiplies it four times, once per
    // neighbour that reads it -- same value each time).  See the state layout at the prefix lambda below.
    constexpr bool CARRY = (NC == 4 && T <= CARRY_MAX_DEPTH);
    float2 A[CARRY ? T : 1][NC];
#pragma unroll
    for (int l = 0; l < (CARRY ? T : 1); l++)
#pragma unroll
        for (int c = 0; c < NC; c++) A[l][c] = make_float2(1.f, 1.f);

    const float2 alpha2 = make_float2(P.alpha, P.alpha);
    const float2 nz2 = make_float2(P.neg_zero, P.neg_zero);
    const float2 negb2 = make_float2(-P.beta, -P.beta);
    const float2 y2 = make_float2(P.rbeta, P.rbeta);
    const float2 zh2 = make_float2(P.zh, P.zh), zl2 = make_float2(P.zl, P.zl);
    const float beta = P.beta;
    float num_min = __int_as_float(0x7f800000);           // +inf
    float in_max = 0.f;

    float *out_u = P.out_u + (size_t)(P.y_base + y0) * w + xc;
    float *out_v = P.out_v + (size_t)(P.y_base + y0) * w + xc;
    float *prev_u = P.prev_u ? P.prev_u + (size_t)(P.y_base + y0) * w + xc : nullptr;
    float *prev_v = P.prev_v ? P.prev_v + (size_t)(P.y_base + y0) * w + xc : nullptr;

    for (int sb = 0; sb < n_steps; sb += UNR) {
#pragma unroll
        for (int uu = 0; uu < UNR; uu++) {
            const int u = uu & 1;
            const int s = sb + uu;
            prefetch(s + PREFETCH);
            cp_async_wait<PREFETCH>();
            const float *slot = my + (s & (RING_SLOTS - 1)) * SLOT_STRIDE;
            float2 fresh[NC];
            if constexpr (NC == 4) {
                const float4 ru = *reinterpret_cast<const float4 *>(slot);
                const float4 rv = *reinterpret_cast<const float4 *>(slot + PLANE_STRIDE);
                fresh[0] = make_float2(ru.x, rv.x); fresh[1] = make_float2(ru.y, rv.y);
                fresh[2] = make_float2(ru.z, rv.z); fresh[3] = make_float2(ru.w, rv.w);
            } else {
                const float2 ru = *reinterpret_cast<const float2 *>(slot);
                const float2 rv = *reinterpret_cast<const float2 *>(slot + PLANE_STRIDE);
                fresh[0] = make_float2(ru.x, rv.x); fresh[1] = make_float2(ru.y, rv.y);
            }
            if constexpr (!EXACT) {
#pragma unroll
                for (int c = 0; c < NC; c++) in_max = max3abs(in_max, fresh[c].x, fresh[c].y);
            }
            const int older = u;
            // Software pipeline over the levels: the part of level l+1 that only needs rows stored in
            // earlier steps -- the alpha products of its centre and top rows, the lane shuffles and
            // (aL + aR) + aT -- is issued before the tail of level l, which depends on the row level l-1
            // has just produced.  Two independent instruction streams per warp instead of one.
            // State per level.  Without CARRY: slot [older^1] holds the centre row (s-l) as VALUES, slot [older] the top
            // row (s-l-1) already MULTIPLIED by alpha (the product formed when that row was the centre row one step
            // ago).  With CARRY both slots hold alpha products -- [older^1] the centre row's, [older] the top row's --
            // and A[] holds the centre row's values; the top slot is dead after the prefix, so the product of the
            // incoming bottom row is written straight into it and the slots just swap roles with the step parity.
            float2 part[NC], part_next[NC], aCen[NC], aCen_next[NC];
            auto prefix = [&](int l, float2(&dst)[NC], float2(&aC)[NC]) {
#pragma unroll
                for (int c = 0; c < NC; c++)        // alpha * centre row: carried from the step that produced it, or formed now
                    aC[c] = CARRY ? S[l - 1][older ^ 1][c] : mulc2(S[l - 1][older ^ 1][c], alpha2, nz2);
                const float2 aLft = shfl_up2(aC[NC - 1]);                 // alpha * (x-1) of cell 0
                const float2 aRgt = shfl_down2(aC[0]);                    // alpha * (x+1) of the last cell
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const float2 lft = (c == 0) ? aLft : aC[c - 1];
                    const float2 rgt = (c == NC - 1) ? aRgt : aC[c + 1];
                    dst[c] = add2(add2(lft, rgt), S[l - 1][older][c]);    // (aL + aR) + aT, fluid.cpp:175-182
                }
            };
            prefix(1, part, aCen);
#pragma unroll
            for (int l = 1; l <= T; l++) {
                float2 o[NC];
                if (l == T && prev_u != nullptr) {
                    // `fresh` is row s-(T-1) of level T-1: the previous iterate, which the reference keeps in its
                    // other buffer (fluid.cpp:188-194); every column this lane stores is valid at that level too
                    const int prow = s - 2 * T + 1;
                    if (store_lane && prow >= 0 && prow < L) {
                        if constexpr (NC == 4) {
                            *reinterpret_cast<float4 *>(prev_u + (size_t)prow * w) =
                                make_float4(fresh[0].x, fresh[1].x, fresh[2].x, fresh[3].x);
                            *reinterpret_cast<float4 *>(prev_v + (size_t)prow * w) =
                                make_float4(fresh[0].y, fresh[1].y, fresh[2].y, fresh[3].y);
                        } else {
                            *reinterpret_cast<float2 *>(prev_u + (size_t)prow * w) = make_float2(fresh[0].x, fresh[1].x);
                            *reinterpret_cast<float2 *>(prev_v + (size_t)prow * w) = make_float2(fresh[0].y, fresh[1].y);
                        }
                    }
                }
                if (l < T) prefix(l + 1, part_next, aCen_next);
                float2 aBot[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const float2 aB = mulc2(fresh[c], alpha2, nz2);       // alpha * bottom row (s-l+1)
                    aBot[c] = aB;
                    // ... + aB) + 1.0f*u_n
                    const float2 num = add2(add2(part[c], aB), CARRY ? A[CARRY ? l - 1 : 0][c] : S[l - 1][older ^ 1][c]);
                    if constexpr (EXACT) {
                        o[c] = make_float2(__fdiv_rn(num.x, beta), __fdiv_rn(num.y, beta));
                    } else {
                        o[c] = DIV2 ? div_const_two2(num, zh2, zl2) : div_const_fast2(num, negb2, y2);
                        num_min = min3abs(num_min, num.x, num.y);
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    if constexpr (CARRY) {
                        S[l - 1][older][c] = aBot[c];         // the dead top slot takes the new centre row's alpha product
                        A[l - 1][c] = fresh[c];               // ... and A its values
                    } else {
                        S[l - 1][older ^ 1][c] = aCen[c];     // the centre row becomes the top row: keep its alpha product
                        S[l - 1][older][c] = fresh[c];        // the bottom row becomes the centre row: keep its values
                        aCen[c] = aCen_next[c];
                    }
                    fresh[c] = o[c];
                    part[c] = part_next[c];
                }
            }
            const int orow = s - 2 * T;
            if (store_lane && orow >= 0 && orow < L) {
                if constexpr (NC == 4) {
                    *reinterpret_cast<float4 *>(out_u + (size_t)orow * w) =
                        make_float4(fresh[0].x, fresh[1].x, fresh[2].x, fresh[3].x);
                    *reinterpret_cast<float4 *>(out_v + (size_t)orow * w) =
                        make_float4(fresh[0].y, fresh[1].y, fresh[2].y, fresh[3].y);
                } else {
                    *reinterpret_cast<float2 *>(out_u + (size_t)orow * w) = make_float2(fresh[0].x, fresh[1].x);
                    *reinterpret_cast<float2 *>(out_v + (size_t)orow * w) = make_float2(fresh[0].y, fresh[1].y);
                }
            }
        }
    }
    cp_async_wait<0>();
    if constexpr (!EXACT) {
        // !(x >= lo) also catches a NaN minimum
        const bool bad = !(num_min >= P.guard_lo) || !(in_max <= P.guard_hi_in);
        if (__any_sync(0xffffffffu, bad) && lane == 0) P.flags[item] = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// Packed pressure sweeps (fluid.cpp:239-258): the pressure field is one plane, so the two halves of a
// register pair are two DIFFERENT strips (A = strip 2k, B = strip 2k+1) of the same rows: identical
// instruction stream, independent data.  State per lane: 16*T registers of pressure rows plus 8*T of
// divergence rows (window of T rows, rotating with period T; the step loop is unrolled by lcm(2,T)).
// (((pL + pR) + pT) + pB) + b, then *0.25 -- the scaling is written as FFMA2(x, 0.25, -0) with a
// run-time -0 so that ptxas cannot contract it into the next level's first add (see mulc2()).
// ---------------------------------------------------------------------------------------------
struct PressurePackedParams {
    const float *in;
    float *out;
    const float *rhs;
    int w, h;
    int strip_out, halo_cols;
    int n_strips, n_pairs, n_chunks, chunk_rows;
    int y_base, wrap;
    float neg_zero;
};

template <int T>
struct PressureUnroll {
    static constexpr int value = (T % 2 == 0) ? T : 2 * T;
};

template <int T, int MINB>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB) pressure_packed_kernel(const PressurePackedParams P)
{
    constexpr int U = PressureUnroll<T>::value;
    __shared__ float4 ring[WARPS_PER_CTA][PRING_SLOTS][4][32];   // per slot: p(A), p(B), div(A), div(B)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS_PER_CTA + warp;
    if (item >= P.n_pairs * P.n_chunks) return;
    const int pair = item % P.n_pairs;
    const int chunk = item / P.n_pairs;
    const int w = P.w, h = P.h;

    const int stripA = 2 * pair;
    const bool haveB = (stripA + 1 < P.n_strips);
    const int stripB = haveB ? stripA + 1 : stripA;          // an odd last strip is computed twice, stored once
    const int x0A = stripA * P.strip_out, x0B = stripB * P.strip_out;
    const int xcA = x0A - P.halo_cols + 4 * lane, xcB = x0B - P.halo_cols + 4 * lane;
    int xwA = xcA % w, xwB = xcB % w;
    if (xwA < 0) xwA += w;
    if (xwB < 0) xwB += w;
    const bool storeA = (xcA >= x0A) && (xcA < x0A + P.strip_out) && (xcA < w);
    const bool storeB = haveB && (xcB >= x0B) && (xcB < x0B + P.strip_out) && (xcB < w);

    const int y0 = chunk * P.chunk_rows;
    const int L = min(P.chunk_rows, h - y0);
    int ld_row = y0 - T;
    if (P.wrap) {
        ld_row %= h;
        if (ld_row < 0) ld_row += h;
    }
    ld_row += P.y_base;
    const int wrap_at = P.wrap ? h : 0x7fffffff;
    const int n_steps = L + 2 * T;

    float4 *my = &ring[warp][0][0][lane];
    constexpr int SLOT_STRIDE = 4 * 32;

    auto prefetch = [&](int s) {
        if (s < n_steps) {
            float4 *dst = my + (s & (PRING_SLOTS - 1)) * SLOT_STRIDE;
            const size_t row = (size_t)ld_row * w;
            cp_async16(dst, P.in + row + xwA);
            cp_async16(dst + 32, P.in + row + xwB);
            cp_async16(dst + 64, P.rhs + row + xwA);
            cp_async16(dst + 96, P.rhs + row + xwB);
            ld_row = (ld_row + 1 == wrap_at) ? 0 : ld_row + 1;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < PPREFETCH; s++) prefetch(s);

    float2 S[T][2][4];      // pressure rows: level, parity slot, cell; .x = strip A, .y = strip B
    float2 Q[T][4];         // divergence rows s-T .. s-1; row r lives in Q[r mod T]
#pragma unroll
    for (int l = 0; l < T; l++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            S[l][0][c] = S[l][1][c] = make_float2(0.f, 0.f);
            Q[l][c] = make_float2(0.f, 0.f);
        }
    const float2 quarter2 = make_float2(0.25f, 0.25f);
    const float2 nz2 = make_float2(P.neg_zero, P.neg_zero);

    float *outA = P.out + (size_t)(P.y_base + y0) * w + xcA;
    float *outB = P.out + (size_t)(P.y_base + y0) * w + xcB;

    for (int sb = 0; sb < n_steps; sb += U) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s = sb + u;
            prefetch(s + PPREFETCH);
            cp_async_wait<PPREFETCH>();
            const float4 *slot = my + (s & (PRING_SLOTS - 1)) * SLOT_STRIDE;
            const float4 pa = slot[0], pb = slot[32], qa = slot[64], qb = slot[96];
            float2 fresh[4] = {make_float2(pa.x, pb.x), make_float2(pa.y, pb.y), make_float2(pa.z, pb.z),
                               make_float2(pa.w, pb.w)};
            const float2 qnew[4] = {make_float2(qa.x, qb.x), make_float2(qa.y, qb.y), make_float2(qa.z, qb.z),
                                    make_float2(qa.w, qb.w)};
            const int older = u & 1;
#pragma unroll
            for (int l = 1; l <= T; l++) {
                float2 o[4];
                const float2 lft = shfl_up2(S[l - 1][older ^ 1][3]);
                const float2 rgt = shfl_down2(S[l - 1][older ^ 1][0]);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float2 pl = (c == 0) ? lft : S[l - 1][older ^ 1][c - 1];
                    const float2 pr = (c == 3) ? rgt : S[l - 1][older ^ 1][c + 1];
                    // fluid.cpp:249-255: ((((pL + pR) + pT) + pB) + 1.0f*b) / 4.0f
                    float2 sum = add2(add2(add2(pl, pr), S[l - 1][older][c]), fresh[c]);
                    sum = add2(sum, Q[(u - l + 2 * U * T) % T][c]);
                    o[c] = mulc2(sum, quarter2, nz2);
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    S[l - 1][older][c] = fresh[c];
                    fresh[c] = o[c];
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) Q[u % T][c] = qnew[c];
            const int orow = s - 2 * T;
            if (orow >= 0 && orow < L) {
                if (storeA)
                    *reinterpret_cast<float4 *>(outA + (size_t)orow * w) =
                        make_float4(fresh[0].x, fresh[1].x, fresh[2].x, fresh[3].x);
                if (storeB)
                    *reinterpret_cast<float4 *>(outB + (size_t)orow * w) =
                        make_float4(fresh[0].y, fresh[1].y, fresh[2].y, fresh[3].y);
            }
        }
    }
    cp_async_wait<0>();
}

template <int T>
int launch_pressure_packed_t(const PressurePackedParams &P, cudaStream_t s)
{
    const int total = P.n_pairs * P.n_chunks;
    const unsigned blocks = (unsigned)((total + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    constexpr int MINB = (T > 2) ? 2 : 3;
    PFS_LAUNCH((pressure_packed_kernel<T, MINB>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    return PFS_OK;
}

template <int T, int MINB, int NC>
int launch_packed_mb(const PackedParams &P, cudaStream_t s)
{
    const int total = P.n_strips * P.n_chunks;
    const unsigned blocks = (unsigned)((total + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    static const int unroll = env_int("PFS_DIFFUSE_UNROLL", 2);
    if (NC == 4 && P.div2 && unroll == 4)
        PFS_LAUNCH((diffuse_packed_kernel<T, false, MINB, NC, (NC == 4), (NC == 4 ? 4 : 2)>), blocks, WARPS_PER_CTA * 32, 0,
                   s, P);
    else if (NC == 4 && P.div2)
        PFS_LAUNCH((diffuse_packed_kernel<T, false, MINB, NC, (NC == 4)>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    else
        PFS_LAUNCH((diffuse_packed_kernel<T, false, MINB, NC>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    PFS_LAUNCH((diffuse_packed_kernel<T, true, MINB, NC>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    --g_passes;     // the repair launch belongs to the same pass
    return PFS_OK;
}

// resident CTAs per SM the kernel is compiled for (register cap 65536 / (128 * MINB)).  4 cells per lane:
// depth <= 4 fits three CTAs (12 warps/SM) without spilling, deeper passes get two (8 warps/SM, up to 255
// registers).  2 cells per lane: four CTAs (16 warps/SM, 128 registers).
constexpr int packed_minb(int t, int nc) { return nc == 2 ? 4 : (t > 4 ? 2 : 3); }

template <int T>
int launch_packed(const PackedParams &P, int cells, cudaStream_t s)
{
    if (cells == 2) return launch_packed_mb<T, packed_minb(T, 2), 2>(P, s);
    return launch_packed_mb<T, packed_minb(T, 4), 4>(P, s);
}

int launch_packed_depth(int t, const PackedParams &P, int cells, cudaStream_t s)
{
    switch (t) {
    case 1: return launch_packed<1>(P, cells, s);
    case 2: return launch_packed<2>(P, cells, s);
    case 3: return launch_packed<3>(P, cells, s);
    case 4: return launch_packed<4>(P, cells, s);
    case 5: return launch_packed<5>(P, cells, s);
    case 6: return launch_packed<6>(P, cells, s);
    case 7: return launch_packed<7>(P, cells, s);
    case 8: return launch_packed<8>(P, cells, s);
    default: set_error("packed diffusion: unsupported depth %d", t); return PFS_EINVAL;
    }
}

// per-device flag buffers (grown on demand)
struct FlagBuf {
    int *ptr = nullptr;
    size_t n = 0;
};
std::mutex g_flag_mutex;
std::map<int, FlagBuf> g_flags;

int get_flags(size_t n, int **out)
{
    int dev = 0;
    PFS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_flag_mutex);
    FlagBuf &fb = g_flags[dev];
    if (fb.n < n) {
        if (fb.ptr) {
            PFS_CUDA(cudaDeviceSynchronize());
            PFS_CUDA(cudaFree(fb.ptr));
            fb.ptr = nullptr;
            fb.n = 0;
        }
        size_t want = n < 16384 ? 16384 : 2 * n;
        PFS_CUDA(cudaMalloc((void **)&fb.ptr, want * sizeof(int)));
        PFS_CUDA(cudaMemset(fb.ptr, 0, want * sizeof(int)));   // flags stay zero between passes: the repair kernel clears what it consumes
        fb.n = want;
    }
    *out = fb.ptr;
    return PFS_OK;
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Is the two-instruction division (div_const_two2) correctly rounded for EVERY numerator with this divisor?
// Tried, not assumed: all 2^23 significands of a in [1,2) against RN(a/b) (formed in double and rounded once more:
// innocuous for a quotient of two 24-bit numbers since 53 >= 2*24+2).  ~10 ms on 8 host threads, once per divisor value.
// ---------------------------------------------------------------------------------------------
struct Div2Entry {
    bool ok;
    float zh, zl;
};
std::mutex g_div2_mutex;
std::map<uint32_t, Div2Entry> g_div2;

Div2Entry div2_constants(float beta)
{
    uint32_t key;
    memcpy(&key, &beta, sizeof(key));
    std::lock_guard<std::mutex> lock(g_div2_mutex);
    auto it = g_div2.find(key);
    if (it != g_div2.end()) return it->second;
    Div2Entry e;
    const double C = 1.0 / (double)beta;
    e.zh = (float)C;
    e.zl = (float)(C - (double)e.zh);
    constexpr int NT = 8;
    bool good[NT];
    std::vector<std::thread> th;
    for (int t = 0; t < NT; t++) {
        th.emplace_back([&, t] {
            bool ok = true;
            const uint32_t lo = (uint32_t)t << 20, hi = lo + (1u << 20);
            for (uint32_t m = lo; m < hi && ok; m++) {
                const uint32_t bits = 0x3f800000u | m;
                float a;
                memcpy(&a, &bits, sizeof(a));
                volatile float prod = a * e.zl;                       // rounded to binary32 on its own
                const float q = fmaf(a, e.zh, prod);
                ok = (q == (float)((double)a / (double)beta));
            }
            good[t] = ok;
        });
    }
    for (auto &x : th) x.join();
    e.ok = true;
    for (int t = 0; t < NT; t++) e.ok = e.ok && good[t];
    g_div2[key] = e;
    return e;
}

// Rows per work item: every warp streams rows + 2T input rows for `rows` output rows, so chunks should be
// tall; but there should also be about one resident wave of warps (`slots`), and all chunks should be the
// same height (a short last chunk costs a whole extra wave).  -> as few chunks as fill the machine once,
// equal heights, 8 <= rows <= 512 (small grids cannot fill the machine with tall chunks: there parallelism
// beats halo overhead, 1024^2 runs 1.6x faster with 9-row chunks than with 32-row ones).
int pick_chunk_rows(int h, int columns_of_items, long long slots, int forced_rows)
{
    int rows = forced_rows;
    if (rows <= 0) {
        long long chunks = slots / (columns_of_items > 0 ? columns_of_items : 1);
        if (chunks < 1) chunks = 1;
        rows = (int)((h + chunks - 1) / chunks);
        if (rows < 8) rows = 8;
        if (rows > 512) rows = 512;
    }
    if (rows > h) rows = h;
    const int n = (h + rows - 1) / rows;
    return (h + n - 1) / n;            // equalise
}

void packed_release_device_buffers()
{
    std::lock_guard<std::mutex> lock(g_flag_mutex);
    for (auto &kv : g_flags) {
        if (kv.second.ptr && cudaSetDevice(kv.first) == cudaSuccess) cudaFree(kv.second.ptr);
    }
    g_flags.clear();
}

bool packed_diffuse_supported(const SweepParams &p)
{
    // alpha >= 0 makes every sweep a convex combination (|values| never exceed the input maximum),
    // which is what lets the guard bound numerators by checking inputs only.
    return (p.w % 4 == 0) && p.w >= 4 && p.h >= 1 && p.alpha >= 0.f && p.beta >= 1.f && p.beta <= 0x1p20f;
}

// 2 or 3: instructions of the constant division the fused passes use for this divisor (diagnostic; host arithmetic only)
int packed_division_ops(float beta)
{
    static const bool div2_env = !(getenv("PFS_DIFFUSE_DIV2") && getenv("PFS_DIFFUSE_DIV2")[0] == '0');
    return (div2_env && div2_constants(beta).ok) ? 2 : 3;
}

int default_diffuse_depth()
{
    static const int d = env_int("PFS_DIFFUSE_DEPTH", 0);
    return (d > 0 && d <= 8) ? d : 6;      // measured best at 4096^2 (profiles/r01_tuning.md)
}

// n diffusion sweeps, up to `depth` per launch, ping-ponging (a0,a1) <-> (b0,b1).
int launch_diffuse_packed(float *a0, float *a1, float *b0, float *b1, const SweepParams &p, int n, int depth,
                          int *flips, cudaStream_t s, float *prev0, float *prev1, int *prev_written)
{
    if (prev_written) *prev_written = 0;
    static const int env_rows = env_int("PFS_DIFFUSE_ROWS", 0);
    static const int env_warps = env_int("PFS_DIFFUSE_WARPS_PER_SM", 0);
    if (depth <= 0) depth = default_diffuse_depth();
    if (depth > 8) depth = 8;
    int hops = 0;
    float *cur0 = a0, *cur1 = a1, *oth0 = b0, *oth1 = b1;
    int left = n;
    while (left > 0) {
        const int t = left < depth ? left : depth;
        PackedParams P;
        P.in_u = cur0; P.in_v = cur1; P.out_u = oth0; P.out_v = oth1;
        const bool last_pass = (left - t == 0) && prev0 != nullptr && t >= 2;
        P.prev_u = last_pass ? prev0 : nullptr;
        P.prev_v = last_pass ? prev1 : nullptr;
        if (last_pass && prev_written) *prev_written = 1;
        P.w = p.w; P.h = p.h; P.y_base = p.y_base; P.wrap = p.wrap;
        static const int env_cells = env_int("PFS_DIFFUSE_CELLS", 0);
        const int cells = (env_cells == 2) ? 2 : 4;                         // cells per lane
        P.halo_cols = cells * ((t + cells - 1) / cells);
        P.strip_out = 32 * cells - 2 * P.halo_cols;
        P.n_strips = (p.w + P.strip_out - 1) / P.strip_out;
        // chunk height: one resident wave of warps if the grid allows it (pick_chunk_rows)
        const int minb = packed_minb(t, cells);
        const int warps_per_sm = env_warps > 0 ? env_warps : 4 * minb;
        const long long slots = (long long)sm_count() * warps_per_sm;
        P.chunk_rows = pick_chunk_rows(p.h, P.n_strips, slots, env_rows);
        P.n_chunks = (p.h + P.chunk_rows - 1) / P.chunk_rows;
        P.alpha = p.alpha; P.beta = p.beta; P.rbeta = 1.0f / p.beta;
        P.guard_lo = 0x1p-96f;
        static const bool div2_env = !(getenv("PFS_DIFFUSE_DIV2") && getenv("PFS_DIFFUSE_DIV2")[0] == '0');
        const Div2Entry d2 = (div2_env && cells == 4) ? div2_constants(p.beta) : Div2Entry{false, 0.f, 0.f};
        P.div2 = d2.ok ? 1 : 0;
        P.zh = d2.zh;
        P.zl = d2.zl;
        if (d2.ok && d2.zl != 0.f)       // numerator * zl must stay a normal number (scale invariance of the proof)
            P.guard_lo = std::max(P.guard_lo, 0x1p-124f / std::fabs(d2.zl));
        P.guard_hi_in = 0x1p60f;
        P.neg_zero = -0.0f;
        PFS_TRY(get_flags((size_t)P.n_strips * P.n_chunks, &P.flags));
        PFS_TRY(launch_packed_depth(t, P, cells, s));
        float *t0 = cur0, *t1 = cur1;
        cur0 = oth0; cur1 = oth1; oth0 = t0; oth1 = t1;
        hops++;
        left -= t;
    }
    *flips = hops;
    return PFS_OK;
}

bool packed_pressure_supported(const SweepParams &p) { return (p.w % 4 == 0) && p.w >= 4 && p.h >= 1; }

// n pressure sweeps, up to `depth` (even, <= 6) per launch, ping-ponging a <-> b; an odd remainder is one
// plain sweep.
int launch_pressure_packed(float *a, float *b, const float *rhs, const SweepParams &p, int n, int depth, int *flips,
                           cudaStream_t s)
{
    static const int env_depth = env_int("PFS_PRESSURE_DEPTH", 0);
    static const int env_rows = env_int("PFS_PRESSURE_ROWS", 0);
    static const int env_warps = env_int("PFS_PRESSURE_WARPS_PER_SM", 0);
    if (depth <= 0) depth = env_depth > 0 ? env_depth : 6;
    if (depth > 8) depth = 8;
    depth &= ~1;
    int hops = 0;
    float *cur = a, *oth = b;
    int left = n;
    while (left > 0) {
        int t = (left >= depth) ? depth : (left & ~1);
        if (depth < 2 || t < 2) {
            int one = 0;
            PFS_TRY(launch_sweeps_basic(SWEEP_PRESSURE, cur, cur, oth, oth, rhs, p, 1, &one, s));
            t = 1;
        } else {
            PressurePackedParams P;
            P.in = cur; P.out = oth; P.rhs = rhs;
            P.w = p.w; P.h = p.h; P.y_base = p.y_base; P.wrap = p.wrap;
            P.halo_cols = 4 * ((t + 3) / 4);
            P.strip_out = 128 - 2 * P.halo_cols;
            P.n_strips = (p.w + P.strip_out - 1) / P.strip_out;
            P.n_pairs = (P.n_strips + 1) / 2;
            const int warps_per_sm = env_warps > 0 ? env_warps : 8;
            const long long slots = (long long)sm_count() * warps_per_sm;
            P.chunk_rows = pick_chunk_rows(p.h, P.n_pairs, slots, env_rows);
            P.n_chunks = (p.h + P.chunk_rows - 1) / P.chunk_rows;
            P.neg_zero = -0.0f;
            switch (t) {
            case 2: PFS_TRY(launch_pressure_packed_t<2>(P, s)); break;
            case 4: PFS_TRY(launch_pressure_packed_t<4>(P, s)); break;
            case 6: PFS_TRY(launch_pressure_packed_t<6>(P, s)); break;
            case 8: PFS_TRY(launch_pressure_packed_t<8>(P, s)); break;
            default: set_error("packed pressure: unsupported depth %d", t); return PFS_EINVAL;
            }
        }
        std::swap(cur, oth);
        hops++;
        left -= t;
    }
    *flips = hops;
    return PFS_OK;
}

}  // namespace pfs
