// sweeps_packed.cu -- temporally blocked smoothing ("diffusion") sweeps on Blackwell's packed-FP32
// pipe: T sweeps of fluid.cpp:154-186 per launch, u and v advanced together.
//
// The diffusion operator applies the same 5-point update with the same coefficients to channel 0
// (u) and channel 1 (v).  sm_100 has two-wide FP32 instructions (add/mul/fma.rn.f32x2 -> FADD2 /
// FMUL2 / FFMA2) that round each half exactly like the scalar instruction, so a lane keeps (u, v)
// of a cell in one 64-bit register pair and issues ONE instruction per pair of updates.  The velocity
// planes are stored the same way -- (u, v) interleaved, 8 bytes per cell -- so a 16-byte load or store
// moves two cells straight into / out of two register pairs, with no repacking.
// The structure is the warp-streaming scheme of sweeps_fused.cu:
//   * a warp owns a strip of 128 columns x a chunk of L rows; each lane holds, for every time level
//     0..T-1, three rows of its 4 cells (24*T registers): the alpha products of the centre and the
//     top row and the values of the centre row, so every value is multiplied by alpha exactly once
//     per level (the reference multiplies it once per neighbour that reads it -- same value);
//   * per stream step one new row arrives through a private cp.async ring in shared memory (no
//     barriers anywhere), level l = 1..T produces row s-l, level T is stored;
//   * x neighbours across lanes come by warp shuffle of the already-multiplied alpha*value pairs;
//   * the division by beta = 1 + 4*alpha is a correctly rounded 2- or 3-instruction division by a
//     constant (below), applied to pairs.  Its preconditions (numerator magnitude in [2^-96, 2^96],
//     not -0) are not branched on in the hot loop: the loop keeps a running FMNMX3 minimum of
//     |numerator| and maximum of |input|; a warp whose extremes leave the safe range recomputes its
//     own work item with __fdiv_rn before it retires (same kernel, same warp, out of line).  Real
//     velocity fields never take that path; exact-zero regions do, and stay correct.
// Results are bit-identical to the one-sweep kernel and to fluid.cpp for every T.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "pfs_internal.cuh"

namespace pfs {

namespace {

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

constexpr int WARPS_PER_CTA = 4;
constexpr int MAX_DEPTH = 6;                  // deeper passes would not fit three rows per level in 255 registers
// cp.async ring depth per warp (rows): eight slots (six rows in flight ahead of the consumer) where registers allow, four
// (two rows in flight) at depth 6, where the second ring pointer pushed ptxas over the 255-register cap into spills.
__host__ __device__ constexpr int ring_slots(int t) { return t >= 6 ? 4 : 8; }
constexpr int UNROLL = 4;                     // stream steps per trip of the main loop; ring slots are immediates within a trip
// Ring slot = one row of the strip (128 cells x 8 bytes).  Global loads are issued so that every cp.async instruction of
// a warp covers 512 CONTIGUOUS bytes (lane i fetches 16-byte piece 32k+i of the row, k = 0,1) -- a lane fetching its own
// 32 bytes as two pieces would touch half of every 32-byte sector per instruction, and the same pattern as a pure copy
// runs at 3.8 instead of 5.7 TB/s (scripts/ubench/stream_copy*.cu).  In shared memory piece p (cells 2p, 2p+1) goes where
// its OWNER lane o = p/2 reads it back with two conflict-free 16-byte loads: the lower pieces of all lanes in floats
// [0,128), the upper pieces in [144,272) (the 64-byte skew keeps the asynchronous writes conflict-free as well).
constexpr int SLOT_HI = 144;                  // float offset of the upper pieces within a slot
constexpr int SLOT_FLOATS = 288;

struct PackedParams {
    const float *in;            // (u,v) plane of the current iterate
    float *out;                 // (u,v) plane receiving iterate +T
    float *prev;                // optional: (u,v) plane receiving iterate +T-1 as well (null = not wanted)
    const float *force;         // optional: interleaved [rows][w][4] field whose channels 0,1 are added to iterate +T as
                                // it is stored (the addForces slot, fluid.cpp:302); row 0 = output row force_skip
    int force_skip, force_rows;
    int w, h;
    int strip_out, halo_cols;
    int n_strips, n_chunks, chunk_rows;
    int y_base, wrap;           // row map of the planes (SweepParams)
    float alpha, beta, rbeta;
    float zh, zl;               // 2-instruction division (div_const_two2): RN(1/beta) and RN(1/beta - zh)
    float guard_lo, guard_hi_in;   // |numerator| >= guard_lo and |input| <= guard_hi_in keep the constant division exact
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

// Four cells = 32 bytes in one 256-bit store (STG.E.256, sm_100): a warp's row goes out as 1 KB without gaps.
// WIDE = false (the out-of-line exact path) uses two 16-byte stores: ptxas 12.9 was seen to emit a single 32-bit STG for
// the 256-bit inline-asm store inside code that also calls the IEEE-division slow path (only the first float of the row
// reached memory).  tests/test_sass_stores.py checks the store forms of every instantiation in the built library.
template <bool WIDE>
__device__ __forceinline__ void store_row(float *p, const float2 (&c)[4])
{
    if constexpr (WIDE) {
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(c[0].x), "f"(c[0].y), "f"(c[1].x),
                     "f"(c[1].y), "f"(c[2].x), "f"(c[2].y), "f"(c[3].x), "f"(c[3].y)
                     : "memory");
    } else {
        *reinterpret_cast<float4 *>(p) = make_float4(c[0].x, c[0].y, c[1].x, c[1].y);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(c[2].x, c[2].y, c[3].x, c[3].y);
    }
}

// ---------------------------------------------------------------------------------------------
// One work item (strip x chunk) streamed by one warp.  EXACT: IEEE division, no guard (the out-of-line repair
// path); otherwise the constant division with the guard, returning "this lane saw a value outside the safe range".
// Lane geometry in the (u,v) plane: the lane owns the unwrapped cells [xc, xc+4) -- 32 contiguous, 32-byte aligned bytes
// (the halo and the width are multiples of four), stored with ONE 256-bit instruction per row; a warp's store covers
// 1 KB without gaps.  Loads: see SLOT_FLOATS.
// ---------------------------------------------------------------------------------------------
// LAST: the pass that reaches sweep n -- the only one that may also store iterate +T-1 and add the external force; the
// other passes (16 of 17 at 100 sweeps) carry neither the pointers nor the branches.
template <int T, bool EXACT, bool DIV2, bool LAST>
__device__ __forceinline__ bool stream_item(const PackedParams &P, const int item, float *my, const int lane)
{
    constexpr int RING_SLOTS = ring_slots(T);
    constexpr int PREFETCH = RING_SLOTS - 2;              // rows in flight ahead of the consumer
    const int strip = item % P.n_strips;
    const int chunk = item / P.n_strips;
    const int w = P.w, h = P.h;
    const long long rs = 2 * (long long)w;                // floats per plane row

    const int x0 = strip * P.strip_out;
    const int xc = x0 - P.halo_cols + 4 * lane;           // unwrapped first cell of this lane
    const bool st = (xc >= x0) && (xc < min(x0 + P.strip_out, w));
    // the two 16-byte pieces this lane FETCHES (not the ones it owns): cells xs + 2*lane and xs + 64 + 2*lane (+1), wrapped
    int xlo = (x0 - P.halo_cols + 2 * lane) % w;
    if (xlo < 0) xlo += w;
    int xhi = (xlo + 64) % w;
    // ... and where they go in a ring slot: owner lane o = 16k + lane/2, lower or upper piece by lane parity
    float *const fetch = my - lane * 4 + (lane & 1) * SLOT_HI + 4 * (lane >> 1);

    const int y0 = chunk * P.chunk_rows;
    const int L = min(P.chunk_rows, h - y0);
    int ld_row = y0 - T;
    if (P.wrap) {
        ld_row %= h;
        if (ld_row < 0) ld_row += h;
    }
    const int wrap_at = P.wrap ? h : 0x7fffffff;
    const int n_steps = L + 2 * T;
    const float *const in0 = P.in + (long long)P.y_base * rs;     // plane row of interior row 0
    const float *ld_ptr = in0 + (long long)ld_row * rs;           // ld_row may be negative on a slab (halo rows)
    const int off_lo = 2 * xlo, off_hi = 2 * xhi;

    // S[l][k][c]: alpha products of level l's centre row (k = older^1) and top row (k = older); k alternates with
    // the step parity.  A[l][c]: the centre row's values.  .x = u, .y = v.
    // Initialised to 1 (not 0) so that warm-up garbage never looks like a zero numerator.
    float2 S[T][2][4], A[T][4];
#pragma unroll
    for (int l = 0; l < T; l++)
#pragma unroll
        for (int c = 0; c < 4; c++) S[l][0][c] = S[l][1][c] = A[l][c] = make_float2(1.f, 1.f);

    const float2 alpha2 = make_float2(P.alpha, P.alpha);
    const float2 nz2 = make_float2(P.neg_zero, P.neg_zero);
    const float2 negb2 = make_float2(-P.beta, -P.beta);
    const float2 y2 = make_float2(P.rbeta, P.rbeta);
    const float2 zh2 = make_float2(P.zh, P.zh), zl2 = make_float2(P.zl, P.zl);
    const float beta = P.beta;
    float num_min = __int_as_float(0x7f800000);           // +inf
    float in_max = 0.f;

    // row pointers of this lane's cells, advanced by one plane row per stream step: `op` is where the row produced at
    // step s goes (output row s - 2T; dereferenced only for 0 <= row < L), `pp` the same for iterate +T-1 (row s-2T+1)
    const long long cell0 = (long long)(P.y_base + y0) * rs + 2 * (long long)xc;
    float *op = P.out + (cell0 - (long long)(2 * T) * rs);
    float *pp = (LAST && P.prev) ? P.prev + (cell0 - (long long)(2 * T - 1) * rs) : nullptr;
    const float *fp = nullptr;                            // force row of the output row, interleaved cells (4 floats each)
    if (LAST && P.force) fp = P.force + 4 * ((long long)w * (long long)(y0 - P.force_skip - 2 * T) + (long long)xc);

    // `slot` is a compile-time constant at every call site (the step loop is unrolled by UNROLL and starts at a multiple
    // of it; with eight slots `ring` alternates between the two halves of the ring from trip to trip), so ring addresses
    // are a base register plus an immediate
    float *ring = my;                                     // owner view of the half being consumed (eight slots: it alternates)
    const long long fetch_delta = fetch - my;             // the same slot seen through this lane's fetch position
    auto prefetch = [&](int s, int slot, float *owner_base) {
        float *base = owner_base + fetch_delta;
        if (s < n_steps) {
            float *dst = base + slot * SLOT_FLOATS;
            cp_async16(dst, ld_ptr + off_lo);
            cp_async16(dst + 64, ld_ptr + off_hi);
            ld_ptr += rs;
            if (++ld_row == wrap_at) {
                ld_row = 0;
                ld_ptr = in0;
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < PREFETCH; s++) prefetch(s, s, my);

    for (int sb = 0; sb < n_steps; sb += UNROLL) {
        // the half of the ring this trip consumes, and the one the rows prefetched PREFETCH steps ahead land in
        float *const other = (RING_SLOTS == 8) ? my + ((ring == my) ? UNROLL * SLOT_FLOATS : 0) : my;
#pragma unroll
        for (int uu = 0; uu < UNROLL; uu++) {
            const int older = uu & 1;
            const int s = sb + uu;
            {
                const int ahead = uu + PREFETCH;              // 2..5 (four slots) or 6..9 (eight slots) steps ahead
                if constexpr (RING_SLOTS == 4)
                    prefetch(s + PREFETCH, ahead % 4, my);
                else
                    prefetch(s + PREFETCH, ahead % 4, (ahead < 8) ? other : ring);
            }
            cp_async_wait<PREFETCH>();
            __syncwarp();     // cp.async completion is per thread, and the pieces a lane reads were fetched by two other lanes
            const float *slot = ring + uu * SLOT_FLOATS;
            float2 fresh[4];
            {
                const float4 lo = *reinterpret_cast<const float4 *>(slot);
                const float4 hi = *reinterpret_cast<const float4 *>(slot + SLOT_HI);
                fresh[0] = make_float2(lo.x, lo.y); fresh[1] = make_float2(lo.z, lo.w);
                fresh[2] = make_float2(hi.x, hi.y); fresh[3] = make_float2(hi.z, hi.w);
            }
            // Software pipeline over the levels: the part of level l+1 that only needs rows stored in
            // earlier steps -- the lane shuffles of its centre row's alpha products and (aL + aR) + aT -- is issued
            // before the tail of level l, which depends on the row level l-1 has just produced.  Two independent
            // instruction streams per warp instead of one.  The top slot is dead after the prefix, so the product of
            // the incoming bottom row is written straight into it and the slots just swap roles with the step parity.
            float2 part[4], part_next[4];
            auto prefix = [&](int l, float2(&dst)[4]) {
                const float2(&aC)[4] = S[l - 1][older ^ 1];
                const float2 aLft = shfl_up2(aC[3]);                      // alpha * (x-1) of cell 0
                const float2 aRgt = shfl_down2(aC[0]);                    // alpha * (x+1) of cell 3
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float2 lft = (c == 0) ? aLft : aC[c - 1];
                    const float2 rgt = (c == 3) ? aRgt : aC[c + 1];
                    dst[c] = add2(add2(lft, rgt), S[l - 1][older][c]);    // (aL + aR) + aT, fluid.cpp:175-182
                }
            };
            prefix(1, part);
#pragma unroll
            for (int l = 1; l <= T; l++) {
                if (LAST && l == T && pp != nullptr) {
                    // `fresh` is row s-(T-1) of level T-1: the previous iterate, which the reference keeps in its
                    // other buffer (fluid.cpp:188-194); every column this lane stores is valid at that level too
                    if (st && (unsigned)(s - 2 * T + 1) < (unsigned)L) store_row<!EXACT>(pp, fresh);
                }
                if (l < T) prefix(l + 1, part_next);
                float2 o[4], aBot[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    aBot[c] = mulc2(fresh[c], alpha2, nz2);               // alpha * bottom row (s-l+1)
                    const float2 num = add2(add2(part[c], aBot[c]), A[l - 1][c]);   // ... + aB) + 1.0f*u_n
                    if constexpr (EXACT) {
                        o[c] = make_float2(__fdiv_rn(num.x, beta), __fdiv_rn(num.y, beta));
                    } else {
                        o[c] = DIV2 ? div_const_two2(num, zh2, zl2) : div_const_fast2(num, negb2, y2);
                        num_min = min3abs(num_min, num.x, num.y);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    S[l - 1][older][c] = aBot[c];         // the dead top slot takes the new centre row's alpha product
                    A[l - 1][c] = fresh[c];               // ... and A its values
                    fresh[c] = o[c];
                    part[c] = part_next[c];
                }
            }
            if ((unsigned)(s - 2 * T) < (unsigned)L) {
                if (LAST && fp != nullptr) {
                    // addForces slot (fluid.cpp:302): channels 0,1 of the force field are added to the row being stored
                    const int frow = y0 + (s - 2 * T) - P.force_skip;
                    if ((unsigned)frow < (unsigned)P.force_rows) {
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            if (st) {
                                const float2 f = __ldg(reinterpret_cast<const float2 *>(fp + 4 * c));
                                fresh[c] = make_float2(__fadd_rn(fresh[c].x, f.x), __fadd_rn(fresh[c].y, f.y));
                            }
                        }
                    }
                }
                if (st) store_row<!EXACT>(op, fresh);
            }
            if constexpr (!EXACT) {
                // the guard on the inputs, read from A[0] (= the level-0 row that arrived in this step) only now: right
                // after the shared-memory load it made the warp sit out the load's latency instead of starting on the
                // work that does not depend on the new row
#pragma unroll
                for (int c = 0; c < 4; c++) in_max = max3abs(in_max, A[0][c].x, A[0][c].y);
            }
            op += rs;
            if constexpr (LAST) {
                if (pp != nullptr) pp += rs;
                if (fp != nullptr) fp += 4 * (long long)w;
            }
        }
        ring = other;
    }
    cp_async_wait<0>();
    if constexpr (EXACT) return false;
    // !(x >= lo) also catches a NaN minimum
    return !(num_min >= P.guard_lo) || !(in_max <= P.guard_hi_in);
}

// The repair path, out of line so that it costs the hot loop neither registers nor instruction-cache footprint.
template <int T>
__device__ __noinline__ void repair_item(const PackedParams &P, const int item, float *my, const int lane)
{
    (void)stream_item<T, true, false, true>(P, item, my, lane);
}

template <int T, int MINB, bool DIV2, bool LAST>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB) diffuse_packed_kernel(const PackedParams P)
{
    __shared__ __align__(16) float ring[WARPS_PER_CTA][ring_slots(T) * SLOT_FLOATS];   // 18 or 36 KB per CTA
    pdl_launch_dependents();
    pdl_wait();                                           // the pass before this one has finished; nothing global was touched yet
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS_PER_CTA + warp;
    if (item >= P.n_strips * P.n_chunks) return;          // whole warp leaves together
    float *my = &ring[warp][lane * 4];                    // owner view: lower piece at +0, upper at +SLOT_HI of a slot
    if (P.guard_lo > 1e30f) {                 // test hook (PFS_DIFFUSE_FORCE_REPAIR=1): the exact path only
        repair_item<T>(P, item, my, lane);
        return;
    }
    const bool bad = stream_item<T, false, DIV2, LAST>(P, item, my, lane);
    if (__any_sync(0xffffffffu, bad)) repair_item<T>(P, item, my, lane);
}

// resident CTAs per SM the kernel is compiled for (register cap 65536 / (128 * MINB)): depth 2 fits four, depth <= 4 three CTAs
// (12 warps/SM), deeper passes get two (8 warps/SM, up to 255 registers)
constexpr int packed_minb(int t) { return t > 4 ? 2 : (t > 2 ? 3 : 4); }

template <int T>
int launch_packed(const PackedParams &P, bool div2, cudaStream_t s)
{
    const int total = P.n_strips * P.n_chunks;
    const unsigned blocks = (unsigned)((total + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    const bool last = (P.prev != nullptr) || (P.force != nullptr);
    if (div2 && last)
        PFS_LAUNCH_PDL((diffuse_packed_kernel<T, packed_minb(T), true, true>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    else if (div2)
        PFS_LAUNCH_PDL((diffuse_packed_kernel<T, packed_minb(T), true, false>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    else if (last)
        PFS_LAUNCH_PDL((diffuse_packed_kernel<T, packed_minb(T), false, true>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    else
        PFS_LAUNCH_PDL((diffuse_packed_kernel<T, packed_minb(T), false, false>), blocks, WARPS_PER_CTA * 32, 0, s, P);
    return PFS_OK;
}

int launch_packed_depth(int t, const PackedParams &P, bool div2, cudaStream_t s)
{
    switch (t) {
    case 2: return launch_packed<2>(P, div2, s);
    case 3: return launch_packed<3>(P, div2, s);
    case 4: return launch_packed<4>(P, div2, s);
    case 5: return launch_packed<5>(P, div2, s);
    case 6: return launch_packed<6>(P, div2, s);
    default: set_error("packed diffusion: unsupported depth %d", t); return PFS_EINVAL;
    }
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
    return (d > 0 && d <= MAX_DEPTH) ? d : 5;      // measured best at 4096^2: 1.45 ms per 100 sweeps against 1.49 at depth 6 (profiles/r02_tuning.md)
}

// n diffusion sweeps on (u,v) planes, up to `depth` per launch, ping-ponging a <-> b.
int launch_diffuse_packed(float *a, float *b, const SweepParams &p, int n, int depth, int *flips, cudaStream_t s,
                          float *prev, int *prev_written, const ForceField *force)
{
    if (prev_written) *prev_written = 0;
    static const int env_rows = env_int("PFS_DIFFUSE_ROWS", 0);
    static const int env_warps = env_int("PFS_DIFFUSE_WARPS_PER_SM", 0);
    static const bool div2_env = !(getenv("PFS_DIFFUSE_DIV2") && getenv("PFS_DIFFUSE_DIV2")[0] == '0');
    if (depth <= 0) depth = default_diffuse_depth();
    if (depth > MAX_DEPTH) depth = MAX_DEPTH;
    int hops = 0;
    float *cur = a, *oth = b;
    int left = n;
    // Pass depths as even as possible: ceil(n / depth) passes of depth d or d+1 (100 sweeps at depth 6: 15 x 6 + 2 x 5, not
    // 16 x 6 + 4 -- shallow passes pay the same memory traffic for fewer sweeps).  A single sweep is left only for n == 1.
    const int n_passes = (n + depth - 1) / depth;
    const int base = n / n_passes, n_deeper = n % n_passes;
    int pass = 0;
    while (left > 0) {
        const int t = std::min(left, base + (pass < n_deeper ? 1 : 0));
        pass++;
        if (t < 2) {
            // a single sweep (n == 1, or depth 1): the plain kernel; iterate n-1 is then simply what it read
            int one = 0;
            PFS_TRY(launch_diffuse_basic(cur, oth, p, 1, &one, s));
            if (left - 1 == 0 && force)
                PFS_TRY(launch_add_forces(oth + (size_t)(p.y_base + force->skip_rows) * 2 * p.w, 2, force->aos, p.w,
                                          force->rows, s));
        } else {
            PackedParams P;
            P.in = cur; P.out = oth;
            const bool last_pass = (left - t == 0);
            P.prev = (last_pass && prev != nullptr) ? prev : nullptr;
            if (P.prev && prev_written) *prev_written = 1;
            P.force = (last_pass && force) ? force->aos : nullptr;
            P.force_skip = force ? force->skip_rows : 0;
            P.force_rows = force ? force->rows : 0;
            P.w = p.w; P.h = p.h; P.y_base = p.y_base; P.wrap = p.wrap;
            P.halo_cols = 4 * ((t + 3) / 4);
            P.strip_out = 128 - 2 * P.halo_cols;
            P.n_strips = (p.w + P.strip_out - 1) / P.strip_out;
            // chunk height: one resident wave of warps if the grid allows it (pick_chunk_rows)
            const int warps_per_sm = env_warps > 0 ? env_warps : WARPS_PER_CTA * packed_minb(t);
            const long long slots = (long long)sm_count() * warps_per_sm;
            P.chunk_rows = pick_chunk_rows(p.h, P.n_strips, slots, env_rows);
            P.n_chunks = (p.h + P.chunk_rows - 1) / P.chunk_rows;
            P.alpha = p.alpha; P.beta = p.beta; P.rbeta = 1.0f / p.beta;
            P.guard_lo = 0x1p-96f;
            const Div2Entry d2 = div2_env ? div2_constants(p.beta) : Div2Entry{false, 0.f, 0.f};
            P.zh = d2.zh;
            P.zl = d2.zl;
            if (d2.ok && d2.zl != 0.f)       // numerator * zl must stay a normal number (scale invariance of the proof)
                P.guard_lo = std::max(P.guard_lo, 0x1p-124f / std::fabs(d2.zl));
            P.guard_hi_in = 0x1p60f;
            static const bool force_exact = getenv("PFS_DIFFUSE_FORCE_REPAIR") && getenv("PFS_DIFFUSE_FORCE_REPAIR")[0] == '1';
            if (force_exact) P.guard_lo = __builtin_inff();     // test hook: every work item takes the out-of-line exact path
            P.neg_zero = -0.0f;
            PFS_TRY(launch_packed_depth(t, P, d2.ok, s));
        }
        std::swap(cur, oth);
        hops++;
        left -= t;
    }
    *flips = hops;
    return PFS_OK;
}

}  // namespace pfs
