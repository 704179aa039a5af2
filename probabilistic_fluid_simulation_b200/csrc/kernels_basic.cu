// kernels_basic.cu -- the non-iterated kernels of the fluid step (advect, divergence, project+pack,
// advect_color, interleaved<->planar movers) and the one-sweep-per-launch Jacobi kernel that is the
// reference point / remainder path for the temporally blocked sweeps in sweeps_fused.cu.
//
// All kernels are HBM-streaming: scalar planes (pressure, divergence) and (u,v) planes (velocity, 8 bytes per cell),
// float4 per thread along x when W % 4 == 0 (V = 4), scalar otherwise (V = 1).  Periodic wrap is resolved by index
// (no halo copies on one GPU).
#include <math.h>
#include <stdlib.h>

#include "pfs_internal.cuh"

namespace pfs {

namespace {

constexpr int BX = 32;   // threads along x
constexpr int BY = 8;    // threads along y

template <int V>
__device__ __forceinline__ void load_vec(const float *p, float (&o)[V])
{
    if constexpr (V == 4) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
        o[0] = __ldg(p);
    }
}

// Four (u,v) cells = 32 contiguous, 32-byte aligned bytes in ONE 256-bit access (LDG/STG.E.256, sm_100): a warp covers
// 1 KB without gaps.  Two 16-byte accesses per lane would touch half of every 32-byte sector per instruction.
__device__ __forceinline__ void load_uv4(const float2 *p, float2 (&c)[4])
{
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(c[0].x), "=f"(c[0].y), "=f"(c[1].x), "=f"(c[1].y), "=f"(c[2].x), "=f"(c[2].y), "=f"(c[3].x), "=f"(c[3].y)
                 : "l"(p));
}
__device__ __forceinline__ void store_uv4(float2 *p, const float2 (&c)[4])
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(c[0].x), "f"(c[0].y), "f"(c[1].x),
                 "f"(c[1].y), "f"(c[2].x), "f"(c[2].y), "f"(c[3].x), "f"(c[3].y)
                 : "memory");
}

template <int V>
__device__ __forceinline__ void store_vec(float *p, const float (&o)[V])
{
    if constexpr (V == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        *p = o[0];
    }
}

// Common per-thread stencil addressing: this thread owns cells [x, x+V) of row j.
// j = interior row (0..h-1, what the interleaved caller buffers are indexed with); rj, jm, jp = PLANE
// rows of the cell and of its vertical neighbours (see SweepParams::y_base / wrap).
struct StencilPos {
    int x, j, rj, xm, xp, jm, jp;
    bool valid;
};

template <int V>
__device__ __forceinline__ StencilPos stencil_pos(int w, int h, int y_base, int wrap)
{
    StencilPos s;
    int xv = blockIdx.x * BX + threadIdx.x;
    s.j = blockIdx.y * BY + threadIdx.y;
    s.x = xv * V;
    s.valid = (s.x < w) && (s.j < h);
    s.xm = (s.x == 0) ? w - 1 : s.x - 1;          // ((i-1) % w + w) % w, fluid.cpp:159
    s.xp = (s.x + V >= w) ? 0 : s.x + V;          // (i+1) % w, fluid.cpp:160
    int jm = s.j - 1, jp = s.j + 1;
    if (wrap) {
        if (jm < 0) jm = h - 1;                   // fluid.cpp:161
        if (jp >= h) jp = 0;                      // fluid.cpp:162
    }
    s.rj = y_base + s.j;
    s.jm = y_base + jm;
    s.jp = y_base + jp;
    return s;
}

inline dim3 stencil_grid(int w, int h, int v, int planes)
{
    int wv = (w + v - 1) / v;
    return dim3((wv + BX - 1) / BX, (h + BY - 1) / BY, planes);
}

// ---------------------------------------------------------------------------------------------
// one Jacobi sweep of the pressure (fluid.cpp:239-258) on a scalar plane
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(BX *BY)
    pressure_sweep_kernel(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ rhs, int w, int h,
                          int y_base, int wrap)
{
    StencilPos s = stencil_pos<V>(w, h, y_base, wrap);
    if (!s.valid) return;
    const float *rc = in + (size_t)s.rj * w;
    float c[V], t[V], b[V], o[V], q[V];
    load_vec<V>(rc + s.x, c);
    load_vec<V>(in + (size_t)s.jm * w + s.x, t);
    load_vec<V>(in + (size_t)s.jp * w + s.x, b);
    float l = __ldg(rc + s.xm), r = __ldg(rc + s.xp);
    load_vec<V>(rhs + (size_t)s.rj * w + s.x, q);
#pragma unroll
    for (int k = 0; k < V; k++) {
        float left = (k == 0) ? l : c[k - 1];
        float right = (k == V - 1) ? r : c[k + 1];
        o[k] = pressure_update(left, right, t[k], b[k], q[k]);
    }
    store_vec<V>(out + (size_t)s.rj * w + s.x, o);
}

// ---------------------------------------------------------------------------------------------
// one smoothing sweep of the velocity (fluid.cpp:154-186) on a (u,v) plane: one cell (8 bytes) per thread.
// Reference point of the packed fused passes, and the path of widths that are not a multiple of four and of
// coefficients outside the packed kernel's range (negative viscosity, huge alpha).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    diffuse_sweep_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int w, int h, float alpha, float beta,
                         int y_base, int wrap)
{
    const int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y;
    if (i >= w || j >= h) return;
    const int im = (i == 0) ? w - 1 : i - 1, ip = (i + 1 >= w) ? 0 : i + 1;     // fluid.cpp:159-160
    int jm = j - 1, jp = j + 1;
    if (wrap) {
        if (jm < 0) jm = h - 1;                                                  // fluid.cpp:161
        if (jp >= h) jp = 0;                                                     // fluid.cpp:162
    }
    const float2 *rc = in + (size_t)(y_base + j) * w;
    const float2 c = __ldg(rc + i), l = __ldg(rc + im), r = __ldg(rc + ip);
    const float2 t = __ldg(in + (size_t)(y_base + jm) * w + i), b = __ldg(in + (size_t)(y_base + jp) * w + i);
    float2 o;
    o.x = diffuse_update(l.x, r.x, t.x, b.x, c.x, alpha, beta);
    o.y = diffuse_update(l.y, r.y, t.y, b.y, c.y, alpha, beta);
    out[(size_t)(y_base + j) * w + i] = o;
}

// ---------------------------------------------------------------------------------------------
// divergence (fluid.cpp:221-237) + optional extraction of the warm-start pressure (channel 2)
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(BX *BY)
    divergence_kernel(const float2 *__restrict__ uv, float *__restrict__ div, const float *__restrict__ p0_src,
                      float *__restrict__ p0, float gamma, int w, int h, int y_base, int wrap)
{
    StencilPos s = stencil_pos<V>(w, h, y_base, wrap);
    if (!s.valid) return;
    const float2 *rc = uv + (size_t)s.rj * w, *rt = uv + (size_t)s.jm * w, *rb = uv + (size_t)s.jp * w;
    float uc[V], vt[V], vb[V], o[V];
    if constexpr (V == 4) {
        // four cells = one 256-bit load per row; u comes from the centre row, v from the rows above and below
        float2 c[4], t[4], b[4];
        load_uv4(rc + s.x, c);
        load_uv4(rt + s.x, t);
        load_uv4(rb + s.x, b);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uc[k] = c[k].x;
            vt[k] = t[k].y;
            vb[k] = b[k].y;
        }
    } else {
        uc[0] = __ldg(rc + s.x).x;
        vt[0] = __ldg(rt + s.x).y;
        vb[0] = __ldg(rb + s.x).y;
    }
    const float ul = __ldg(rc + s.xm).x, urr = __ldg(rc + s.xp).x;
#pragma unroll
    for (int k = 0; k < V; k++) {
        float left = (k == 0) ? ul : uc[k - 1];
        float right = (k == V - 1) ? urr : uc[k + 1];
        o[k] = divergence_value(right, left, vb[k], vt[k], gamma);
    }
    store_vec<V>(div + (size_t)s.rj * w + s.x, o);
    if (p0 != nullptr) {
        const float *src = p0_src + ((size_t)s.j * w + s.x) * 4 + 2;   // interleaved buffers have no halo rows
        float pv[V];
#pragma unroll
        for (int k = 0; k < V; k++) pv[k] = __ldg(src + 4 * k);
        store_vec<V>(p0 + (size_t)s.rj * w + s.x, pv);
    }
}

// ---------------------------------------------------------------------------------------------
// subtract pressure gradient (fluid.cpp:273-294) fused with the write-back of both interleaved
// post-state buffers (full 16-byte cells, so no read-modify-write of the caller's buffers).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BX *BY)
    project_pack_kernel(const float2 *__restrict__ uv, const float *__restrict__ pn, const float *__restrict__ pprev,
                        const float *__restrict__ div, float4 *__restrict__ out_q, float4 *__restrict__ out_p, float dt,
                        int w, int h, int y_base, int wrap)
{
    // One cell per thread: a warp's two 16-byte stores per cell then cover 512 contiguous bytes of each interleaved
    // buffer (full sectors); four cells per thread would leave every store instruction scattered over 32 half-written sectors.
    StencilPos s = stencil_pos<1>(w, h, y_base, wrap);
    if (!s.valid) return;
    const float *pr = pn + (size_t)s.rj * w;
    const size_t off = (size_t)s.rj * w + s.x;                 // planes (with halo rows)
    const size_t cell = (size_t)s.j * w + s.x;                 // interleaved buffers (no halo rows)
    const float pc = __ldg(pr + s.x), pl = __ldg(pr + s.xm), prr = __ldg(pr + s.xp);
    const float pt = __ldg(pn + (size_t)s.jm * w + s.x), pb = __ldg(pn + (size_t)s.jp * w + s.x);
    const float2 v = __ldg(uv + off);
    const float pp = __ldg(pprev + off), dd = __ldg(div + off);
    const float un = project_component(v.x, prr, pl, dt);
    const float vn = project_component(v.y, pb, pt, dt);
    out_q[cell] = make_float4(un, vn, pp, dd);
    out_p[cell] = make_float4(v.x, v.y, pc, dd);
}

// The same subtraction with the result kept in a (u,v) plane (persistent-state contexts): two cells per thread, one
// 16-byte load and store of the velocity.  vmax (optional): running maximum of |v| of the projected field -- what bounds
// the row displacement of the gathers that follow (advect_color of this step, advect of the next); NaN counts as +inf.
template <int V>
__global__ void __launch_bounds__(BX *BY)
    project_uv_kernel(const float2 *__restrict__ uv, const float *__restrict__ pn, float2 *__restrict__ uv_out, float dt,
                      int w, int h, int y_base, int wrap, float *vmax)
{
    StencilPos s = stencil_pos<V>(w, h, y_base, wrap);
    float m = 0.f;
    if (s.valid) {
        const float *pr = pn + (size_t)s.rj * w;
        float pc[V], pt[V], pb[V];
        load_vec<V>(pr + s.x, pc);
        load_vec<V>(pn + (size_t)s.jm * w + s.x, pt);
        load_vec<V>(pn + (size_t)s.jp * w + s.x, pb);
        const float pl = __ldg(pr + s.xm), prr = __ldg(pr + s.xp);
        const float2 *src = uv + (size_t)s.rj * w + s.x;
        float2 *dst = uv_out + (size_t)s.rj * w + s.x;
        float2 v[V], o[V];
        if constexpr (V == 4)
            load_uv4(src, v);
        else
            v[0] = __ldg(src);
#pragma unroll
        for (int k = 0; k < V; k++) {
            const float left = (k == 0) ? pl : pc[k - 1];
            const float right = (k == V - 1) ? prr : pc[k + 1];
            o[k].x = project_component(v[k].x, right, left, dt);
            o[k].y = project_component(v[k].y, pb[k], pt[k], dt);
            const float av = fabsf(o[k].y);
            m = (av > m || av != av) ? av : m;
        }
        if constexpr (V == 4)
            store_uv4(dst, o);
        else
            *dst = o[0];
    }
    if (vmax != nullptr) {                      // every thread of the block gets here
        __shared__ float warp_max[BX * BY / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float t = __shfl_xor_sync(0xffffffffu, m, o);
            m = (t > m || t != t) ? t : m;
        }
        const int tid = threadIdx.y * BX + threadIdx.x;
        if ((tid & 31) == 0) warp_max[tid >> 5] = m;
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < BX * BY / 32; q++) {
                const float t = warp_max[q];
                m = (t > m || t != t) ? t : m;
            }
            if (m != m) m = __int_as_float(0x7f800000);
            if (m > *reinterpret_cast<volatile float *>(vmax))
                atomicMax(reinterpret_cast<int *>(vmax), __float_as_int(m));    // non-negative floats order like their bits
        }
    }
}

// stand-alone subtractPressureGradient on interleaved buffers (operator API only)
__global__ void __launch_bounds__(256)
    subtract_gradient_aos_kernel(const float4 *__restrict__ in, float *__restrict__ out, float dt, int w, int h)
{
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y;
    if (i >= w || j >= h) return;
    int im = (i == 0) ? w - 1 : i - 1, ip = (i + 1 >= w) ? 0 : i + 1;
    int jm = (j == 0) ? h - 1 : j - 1, jp = (j + 1 >= h) ? 0 : j + 1;
    const float *f = reinterpret_cast<const float *>(in);
    float pl = __ldg(f + ((size_t)j * w + im) * 4 + 2), pr = __ldg(f + ((size_t)j * w + ip) * 4 + 2);
    float pt = __ldg(f + ((size_t)jm * w + i) * 4 + 2), pb = __ldg(f + ((size_t)jp * w + i) * 4 + 2);
    float2 uv = __ldg(reinterpret_cast<const float2 *>(f + ((size_t)j * w + i) * 4));
    float2 o;
    o.x = project_component(uv.x, pr, pl, dt);
    o.y = project_component(uv.y, pb, pt, dt);
    *reinterpret_cast<float2 *>(out + ((size_t)j * w + i) * 4) = o;
}

// ---------------------------------------------------------------------------------------------
// advect (fluid.cpp:24-70): back-trace, periodic wrap, bilinear gather of (u, v) from the interleaved field or from
// a (u,v) plane.  One 8-byte load fetches both components of a corner; neighbouring threads'
// corners share 32-byte sectors, so the gather runs out of L1 for coherent flows.
// ---------------------------------------------------------------------------------------------
// R rows per thread (rows j and j + 4 of an 8-row tile for R = 2): the kernel is two dependent round trips to memory per
// cell (own velocity, then the four corners) with ~120 instructions between and after them; at one cell per thread it
// spent most of its time waiting on the scoreboard (ncu: issue 66 %, long-scoreboard stalls dominant).  Two independent
// cells per thread, staged (all own loads, all index arithmetic, all gathers, all blends), overlap those waits.
template <int SS, int DS, int R>     // floats per source / destination cell: 4 = interleaved [u,v,p,div], 2 = (u,v) plane
__global__ void __launch_bounds__(256)
    advect_kernel(const float *__restrict__ src, float *__restrict__ dst, float dt, int w, int h, float rfw, float rfh)
{
    const int i = blockIdx.x * 64 + threadIdx.x, j0 = blockIdx.y * (4 * R) + threadIdx.y;
    if (i >= w || j0 >= h) return;
    const float fw = (float)w, fh = (float)h;
    // a field has at most 2^28 cells (check_dims): 32-bit cell indices, one widening multiply-add per address
    const unsigned uw = (unsigned)w;
    bool live[R];
    unsigned cell[R];
    float2 uv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int j = j0 + 4 * r;
        live[r] = j < h;
        cell[r] = (unsigned)(live[r] ? j : j0) * uw + (unsigned)i;
        uv[r] = __ldg(reinterpret_cast<const float2 *>(src + (size_t)cell[r] * SS));
    }
    Bilinear b[R];
    unsigned r0[R], r1[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        // fluid.cpp:39,41: (float)i - dt*u/fwidth  ==  i - ((dt*u)/fwidth)
        // rfw, rfh: the correctly rounded reciprocals of the extents, formed once on the host (1.0f / fw)
        const float xp = backtrace_coord((float)i, __fmul_rn(dt, uv[r].x), fw, rfw);
        const float yp = backtrace_coord((float)(live[r] ? j0 + 4 * r : j0), __fmul_rn(dt, uv[r].y), fh, rfh);
        b[r] = make_bilinear(xp, yp, w, h);
        r0[r] = (unsigned)b[r].j0 * uw;
        r1[r] = (unsigned)b[r].j1 * uw;
    }
    float2 f00[R], f10[R], f01[R], f11[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        f00[r] = __ldg(reinterpret_cast<const float2 *>(src + (size_t)(r0[r] + (unsigned)b[r].i0) * SS));
        f10[r] = __ldg(reinterpret_cast<const float2 *>(src + (size_t)(r0[r] + (unsigned)b[r].i1) * SS));
        f01[r] = __ldg(reinterpret_cast<const float2 *>(src + (size_t)(r1[r] + (unsigned)b[r].i0) * SS));
        f11[r] = __ldg(reinterpret_cast<const float2 *>(src + (size_t)(r1[r] + (unsigned)b[r].i1) * SS));
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float un = bilerp(b[r], f00[r].x, f10[r].x, f01[r].x, f11[r].x);
        const float vn = bilerp(b[r], f00[r].y, f10[r].y, f01[r].y, f11[r].y);
        if (live[r]) *reinterpret_cast<float2 *>(dst + (size_t)cell[r] * DS) = make_float2(un, vn);
    }
}

// ---------------------------------------------------------------------------------------------
// advect_color (fluid.cpp:72-127): one thread per pixel, velocity point-sampled at
// ((int)(i*viw), (int)(j*vih)), four float4 texel gathers, one coalesced float4 store.
// ---------------------------------------------------------------------------------------------
// (png_byte)(x * 255.0) of includes/utils.hpp:129-131 without the double: the product of a binary32 and 255 rounded TOWARD
// ZERO lies on the same side of every integer as the exact product (integers up to 2^24 are representable), so its
// truncation is the truncation of the exact product -- which is what the reference's double multiply computes.
__device__ __forceinline__ unsigned frame_byte(float x)
{
    const int t = __float2int_rz(__fmul_rz(x, 255.0f));
    return (unsigned)min(max(t, 0), 255);           // outside [0, 256/255) the reference's cast is undefined; saturate
}

template <int VS, bool BYTES, int R>     // floats per velocity cell: 4 = interleaved buffer, 2 = (u,v) plane; BYTES: also the frame;
__global__ void __launch_bounds__(256)   // R rows per thread, staged as in advect_kernel
    advect_color_kernel(const float4 *__restrict__ image, float4 *__restrict__ out, const float *__restrict__ vp,
                        float dt_over_viw, float dt_over_vih, float viw, float vih, int iw, int ih, int vw, float rfiw, float rfih,
                        unsigned *__restrict__ rgba8)
{
    const int i = blockIdx.x * 64 + threadIdx.x, jb = blockIdx.y * (4 * R) + threadIdx.y;
    if (i >= iw || jb >= ih) return;
    const float fiw = (float)iw, fih = (float)ih;
    const unsigned uw = (unsigned)iw;    // <= 2^28 pixels (check_dims)
    const int vi = (int)__fmul_rn((float)i, viw);   // fluid.cpp:89
    bool live[R];
    int jj[R];
    float2 uv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        live[r] = jb + 4 * r < ih;
        jj[r] = live[r] ? jb + 4 * r : jb;
        const int vj = (int)__fmul_rn((float)jj[r], vih);   // fluid.cpp:90
        uv[r] = __ldg(reinterpret_cast<const float2 *>(vp + (size_t)((unsigned)vj * (unsigned)vw + (unsigned)vi) * VS));
    }
    Bilinear b[R];
    unsigned r0[R], r1[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        // fluid.cpp:97-98: (float)i - (dt/viw) * u / fiwidth
        const float xp = backtrace_coord((float)i, __fmul_rn(dt_over_viw, uv[r].x), fiw, rfiw);
        const float yp = backtrace_coord((float)jj[r], __fmul_rn(dt_over_vih, uv[r].y), fih, rfih);
        b[r] = make_bilinear(xp, yp, iw, ih);
        r0[r] = (unsigned)b[r].j0 * uw;
        r1[r] = (unsigned)b[r].j1 * uw;
    }
    float4 f00[R], f10[R], f01[R], f11[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        f00[r] = __ldg(image + (r0[r] + (unsigned)b[r].i0));
        f10[r] = __ldg(image + (r0[r] + (unsigned)b[r].i1));
        f01[r] = __ldg(image + (r1[r] + (unsigned)b[r].i0));
        f11[r] = __ldg(image + (r1[r] + (unsigned)b[r].i1));
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        float4 o;
        o.x = bilerp(b[r], f00[r].x, f10[r].x, f01[r].x, f11[r].x);
        o.y = bilerp(b[r], f00[r].y, f10[r].y, f01[r].y, f11[r].y);
        o.z = bilerp(b[r], f00[r].z, f10[r].z, f01[r].z, f11[r].z);
        o.w = bilerp(b[r], f00[r].w, f10[r].w, f01[r].w, f11[r].w);
        if (live[r]) {
            const unsigned pix = (unsigned)jj[r] * uw + (unsigned)i;
            out[pix] = o;
            if (BYTES)      // the frame the reference's writer would form from this pixel: R, G, B, A in memory order
                rgba8[pix] = frame_byte(o.x) | (frame_byte(o.y) << 8) | (frame_byte(o.z) << 16) | (frame_byte(o.w) << 24);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// interleaved <-> planar movers (operator API; the fused step never needs a separate pass)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    unpack_kernel(const float4 *__restrict__ aos, float2 *uv, float *p, float *div, size_t n)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float4 v = __ldg(aos + i);
    if (uv) uv[i] = make_float2(v.x, v.y);
    if (p) p[i] = v.z;
    if (div) div[i] = v.w;
}

__global__ void __launch_bounds__(256)
    pack_kernel(float *__restrict__ aos, const float2 *uv, const float *p, const float *div, size_t n)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (uv && p && div) {
        const float2 v = uv[i];
        reinterpret_cast<float4 *>(aos)[i] = make_float4(v.x, v.y, p[i], div[i]);
        return;
    }
    if (uv) *reinterpret_cast<float2 *>(aos + i * 4) = uv[i];
    if (p && div) {
        *reinterpret_cast<float2 *>(aos + i * 4 + 2) = make_float2(p[i], div[i]);
    } else {
        if (p) aos[i * 4 + 2] = p[i];
        if (div) aos[i * 4 + 3] = div[i];
    }
}

// addForces slot (fluid.cpp:198-208; the reference body is an empty loop over every cell and channel): the velocity
// channels take the force, dst[cell] (u,v) += force[cell] channels 0,1, one rounded addition each.
template <int DS>
__global__ void __launch_bounds__(256) add_forces_kernel(float *__restrict__ dst, const float4 *__restrict__ force, size_t n)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 f = __ldg(force + i);
    float2 *d = reinterpret_cast<float2 *>(dst + i * DS);
    const float2 v = *d;
    *d = make_float2(__fadd_rn(v.x, f.x), __fadd_rn(v.y, f.y));
}

// ---------------------------------------------------------------------------------------------
// Opt-in stochastic forcing at the reference's (empty) addForces slot, fluid.cpp:198-208 / :302.
// The reference has no stochastic term; oracle/fluid_oracle.c restates THIS definition (Philox-4x32-10,
// counter = (cell_lo, cell_hi, step, draw), key = seed; 16 uniform 16-bit integers per cell, 8 summed per
// component) so that the two agree bit for bit.  Integer arithmetic up to the last three float operations.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// u, v: either two planes (stride 1, row pitch w, first row y_base) or channels 0,1 of an interleaved
// buffer (stride 4).  row0 = global row of local row 0 (cell index = (row0 + j) * w + i).
__global__ void __launch_bounds__(256)
    stochastic_force_kernel(float *__restrict__ u, float *__restrict__ v, int stride, float sigma, float norm,
                            uint32_t seed_lo, uint32_t seed_hi, uint32_t step, int w, int h, int row0, int y_base)
{
    const int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y;
    if (i >= w || j >= h) return;
    const unsigned long long cell = (unsigned long long)(row0 + j) * (unsigned long long)w + (unsigned long long)i;
    uint32_t a[4], b[4];
    philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), step, 0u, seed_lo, seed_hi, a);
    philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), step, 1u, seed_lo, seed_hi, b);
    int su = 0, sv = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        su += (int)(a[q] & 0xffffu) + (int)(a[q] >> 16);
        sv += (int)(b[q] & 0xffffu) + (int)(b[q] >> 16);
    }
    const float gu = __fmul_rn((float)(2 * su - 8 * 65535), norm);
    const float gv = __fmul_rn((float)(2 * sv - 8 * 65535), norm);
    const size_t o = ((size_t)(y_base + j) * w + i) * (size_t)stride;
    u[o] = __fadd_rn(u[o], __fmul_rn(sigma, gu));
    v[o] = __fadd_rn(v[o], __fmul_rn(sigma, gv));
}

// ---------------------------------------------------------------------------------------------
// Step diagnostics (not in the reference, which runs a fixed sweep count with no convergence test,
// fluid.cpp:239): from the two post-state buffers of simulate_fluid_step,
//   [0] sum div^2          (vp ch3)                 [1] sum (p_N - p_{N-1})^2  (tmp ch2 - vp ch2)
//   [2] sum (u^2 + v^2)    (vp ch0,1, projected)    [3] max(|u|, |v|)
// Warp-shuffle tree + one partial per block (double), then a single block folds the partials in a fixed
// order, so the result does not depend on scheduling.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    norms_partial_kernel(const float4 *__restrict__ vp, const float4 *__restrict__ tmp, size_t n, double *partials)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float4 a = __ldg(vp + i), b = __ldg(tmp + i);
        s0 += (double)a.w * (double)a.w;
        const double r = (double)b.z - (double)a.z;
        s1 += r * r;
        s2 += (double)a.x * (double)a.x + (double)a.y * (double)a.y;
        m = fmaxf(m, fmaxf(fabsf(a.x), fabsf(a.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    __shared__ double sh[8][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[warp][0] = s0; sh[warp][1] = s1; sh[warp][2] = s2; sh[warp][3] = (double)m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        for (int q = 0; q < 8; q++) {
            t0 += sh[q][0]; t1 += sh[q][1]; t2 += sh[q][2]; t3 = fmax(t3, sh[q][3]);
        }
        double *o = partials + 4 * (size_t)blockIdx.x;
        o[0] = t0; o[1] = t1; o[2] = t2; o[3] = t3;
    }
}

// Same reduction for two SoA planes: [0] sum (a-b)^2, [1] 0, [2] sum a^2, [3] max |a-b|
// (a = iterate N, b = iterate N-1 of the pressure solve; used by pfs_compute_pressure_adaptive).
__global__ void __launch_bounds__(256)
    plane_diff_partial_kernel(const float *__restrict__ a, const float *__restrict__ b, size_t n, double *partials)
{
    double s0 = 0.0, s2 = 0.0;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float x = __ldg(a + i), y = __ldg(b + i);
        const double r = (double)x - (double)y;
        s0 += r * r;
        s2 += (double)x * (double)x;
        m = fmaxf(m, fabsf(__fsub_rn(x, y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    __shared__ double sh[8][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[warp][0] = s0; sh[warp][1] = s2; sh[warp][2] = (double)m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t2 = 0, t3 = 0;
        for (int q = 0; q < 8; q++) {
            t0 += sh[q][0]; t2 += sh[q][1]; t3 = fmax(t3, sh[q][2]);
        }
        double *o = partials + 4 * (size_t)blockIdx.x;
        o[0] = t0; o[1] = 0.0; o[2] = t2; o[3] = t3;
    }
}

// One warp folds the per-block partials: lane l takes partials l, l+32, ... in that order, then a shuffle tree.
// The order depends only on nblocks, so the result does not depend on scheduling.
__global__ void norms_final_kernel(const double *partials, int nblocks, double *out)
{
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int q = threadIdx.x; q < nblocks; q += 32) {
        t0 += partials[4 * q + 0]; t1 += partials[4 * q + 1]; t2 += partials[4 * q + 2];
        t3 = fmax(t3, partials[4 * q + 3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        t3 = fmax(t3, __shfl_xor_sync(0xffffffffu, t3, o));
    }
    if (threadIdx.x == 0) {
        out[0] = t0; out[1] = t1; out[2] = t2; out[3] = t3;
    }
}

// ---------------------------------------------------------------------------------------------
// Frame packing (includes/utils.hpp:129-131): byte = (png_byte)(x * 255.0) -- a DOUBLE multiply and a
// truncation, exactly as the reference's write_png_from_array does on the host.  One pixel (4 channels) per
// thread, one 4-byte store.  Values outside [0, 256/255) are undefined behaviour in the reference's cast; here
// they saturate.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_rgba8_kernel(const float4 *__restrict__ image, uchar4 *__restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 c = __ldg(image + i);
    auto q = [](float x) -> unsigned char {
        const double v = (double)x * 255.0;
        const int t = (int)v;                       // truncation toward zero
        return (unsigned char)(t < 0 ? 0 : (t > 255 ? 255 : t));
    };
    out[i] = make_uchar4(q(c.x), q(c.y), q(c.z), q(c.w));
}

// ---------------------------------------------------------------------------------------------
// Red-black successive over-relaxation of the pressure equation -- NOT the reference's solver (fluid.cpp:239-266 is a
// fixed number of Jacobi sweeps) and not a parity path: an opt-in alternative reported beside it.
// One half-sweep per launch, in place: cells with (i + j) % 2 == colour take p <- p + omega * (jacobi(p) - p), where
// jacobi(p) is the reference's update with the four neighbours of the other colour.  The sum of squared updates goes to
// per-block partials (double), folded by norms_final_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    sor_half_sweep_kernel(float *__restrict__ p, const float *__restrict__ rhs, int w, int h, float omega, int colour,
                          double *partials, int accumulate)
{
    const int j = blockIdx.y * 4 + threadIdx.y;
    const int i = 2 * (blockIdx.x * 64 + threadIdx.x) + ((j + colour) & 1);
    double d2 = 0.0;
    if (i < w && j < h) {
        const int im = (i == 0) ? w - 1 : i - 1, ip = (i + 1 >= w) ? 0 : i + 1;
        const int jm = (j == 0) ? h - 1 : j - 1, jp = (j + 1 >= h) ? 0 : j + 1;
        const size_t c = (size_t)j * w + i;
        const float old = p[c];
        const float gs = pressure_update(p[(size_t)j * w + im], p[(size_t)j * w + ip], p[(size_t)jm * w + i], p[(size_t)jp * w + i], rhs[c]);
        const float upd = __fmul_rn(omega, __fsub_rn(gs, old));
        p[c] = __fadd_rn(old, upd);
        d2 = (double)upd * (double)upd;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    __shared__ double sh[8];
    const int tid = threadIdx.y * 64 + threadIdx.x;
    if ((tid & 31) == 0) sh[tid >> 5] = d2;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int q = 0; q < 8; q++) t += sh[q];
        double *o = partials + 4 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
        o[0] = accumulate ? o[0] + t : t;
        o[1] = o[2] = o[3] = 0.0;
    }
}

inline bool vec4_ok(int w, const void *a, const void *b = nullptr, const void *c = nullptr, const void *d = nullptr,
                    const void *e = nullptr, const void *f = nullptr)
{
    auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return (w % 4 == 0) && al(a) && al(b) && al(c) && al(d) && al(e) && al(f);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------------
// rows per thread of the two gather kernels (PFS_GATHER_ROWS=1|2, default 2; see advect_kernel)
static int env_rows_per_thread()
{
    const char *e = getenv("PFS_GATHER_ROWS");
    return (e && e[0] == '1') ? 1 : 2;
}

int launch_unpack(const float *aos, float *uv, float *p, float *div, int w, int h, cudaStream_t s)
{
    size_t n = (size_t)w * h;
    unsigned blocks = (unsigned)((n + 255) / 256);
    PFS_LAUNCH(unpack_kernel, blocks, 256, 0, s, reinterpret_cast<const float4 *>(aos), reinterpret_cast<float2 *>(uv), p,
               div, n);
    return PFS_OK;
}

int launch_pack(float *aos, const float *uv, const float *p, const float *div, int w, int h, cudaStream_t s)
{
    size_t n = (size_t)w * h;
    unsigned blocks = (unsigned)((n + 255) / 256);
    PFS_LAUNCH(pack_kernel, blocks, 256, 0, s, aos, reinterpret_cast<const float2 *>(uv), p, div, n);
    return PFS_OK;
}

int launch_add_forces(float *dst, int dst_stride, const float *force_aos, int w, int rows, cudaStream_t s)
{
    const size_t n = (size_t)w * rows;
    if (n == 0) return PFS_OK;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    const float4 *f = reinterpret_cast<const float4 *>(force_aos);
    if (dst_stride == 2)
        PFS_LAUNCH(add_forces_kernel<2>, blocks, 256, 0, s, dst, f, n);
    else
        PFS_LAUNCH(add_forces_kernel<4>, blocks, 256, 0, s, dst, f, n);
    return PFS_OK;
}

int launch_advect(const float *src, int src_stride, float *dst, int dst_stride, float dt, int w, int h, cudaStream_t s)
{
    static const int rows_env = env_rows_per_thread();
    const int R = rows_env;
    dim3 block(64, 4), grid((w + 63) / 64, (h + 4 * R - 1) / (4 * R));
    const float rfw = 1.0f / (float)w, rfh = 1.0f / (float)h;     // binary32 division on the host: correctly rounded, as __frcp_rn
    if (src_stride == 4 && dst_stride == 2 && R == 2)
        PFS_LAUNCH((advect_kernel<4, 2, 2>), grid, block, 0, s, src, dst, dt, w, h, rfw, rfh);
    else if (src_stride == 4 && dst_stride == 2)
        PFS_LAUNCH((advect_kernel<4, 2, 1>), grid, block, 0, s, src, dst, dt, w, h, rfw, rfh);
    else if (src_stride == 2 && dst_stride == 2 && R == 2)
        PFS_LAUNCH((advect_kernel<2, 2, 2>), grid, block, 0, s, src, dst, dt, w, h, rfw, rfh);
    else if (src_stride == 2 && dst_stride == 2)
        PFS_LAUNCH((advect_kernel<2, 2, 1>), grid, block, 0, s, src, dst, dt, w, h, rfw, rfh);
    else if (src_stride == 4 && dst_stride == 4 && R == 2)
        PFS_LAUNCH((advect_kernel<4, 4, 2>), grid, block, 0, s, src, dst, dt, w, h, rfw, rfh);
    else if (src_stride == 4 && dst_stride == 4)
        PFS_LAUNCH((advect_kernel<4, 4, 1>), grid, block, 0, s, src, dst, dt, w, h, rfw, rfh);
    else {
        set_error("advect: unsupported cell strides %d -> %d", src_stride, dst_stride);
        return PFS_EINVAL;
    }
    return PFS_OK;
}

int launch_pressure_basic(float *a, float *b, const float *rhs, const SweepParams &p, int n, int *flips, cudaStream_t s)
{
    const bool v4 = vec4_ok(p.w, a, b, rhs);
    dim3 block(BX, BY), grid = stencil_grid(p.w, p.h, v4 ? 4 : 1, 1);
    for (int it = 0; it < n; it++) {
        const float *in = (it & 1) ? b : a;
        float *out = (it & 1) ? a : b;
        if (v4)
            PFS_LAUNCH(pressure_sweep_kernel<4>, grid, block, 0, s, in, out, rhs, p.w, p.h, p.y_base, p.wrap);
        else
            PFS_LAUNCH(pressure_sweep_kernel<1>, grid, block, 0, s, in, out, rhs, p.w, p.h, p.y_base, p.wrap);
    }
    *flips = n;
    return PFS_OK;
}

int launch_diffuse_basic(float *a_uv, float *b_uv, const SweepParams &p, int n, int *flips, cudaStream_t s)
{
    dim3 block(64, 4), grid((p.w + 63) / 64, (p.h + 3) / 4);
    for (int it = 0; it < n; it++) {
        const float2 *in = reinterpret_cast<const float2 *>((it & 1) ? b_uv : a_uv);
        float2 *out = reinterpret_cast<float2 *>((it & 1) ? a_uv : b_uv);
        PFS_LAUNCH(diffuse_sweep_kernel, grid, block, 0, s, in, out, p.w, p.h, p.alpha, p.beta, p.y_base, p.wrap);
    }
    *flips = n;
    return PFS_OK;
}

int launch_divergence(const float *uv, float *div, const float *p0_src_aos, float *p0, float dt, int w, int h,
                      cudaStream_t s, int y_base, int wrap)
{
    const float gamma = (float)(-1.0 / (double)dt);   // fluid.cpp:218
    const bool v4 = vec4_ok(w, uv, div, p0);
    dim3 block(BX, BY), grid = stencil_grid(w, h, v4 ? 4 : 1, 1);
    const float2 *uv2 = reinterpret_cast<const float2 *>(uv);
    if (v4)
        PFS_LAUNCH(divergence_kernel<4>, grid, block, 0, s, uv2, div, p0_src_aos, p0, gamma, w, h, y_base, wrap);
    else
        PFS_LAUNCH(divergence_kernel<1>, grid, block, 0, s, uv2, div, p0_src_aos, p0, gamma, w, h, y_base, wrap);
    return PFS_OK;
}

int launch_project_pack(const float *uv, const float *p_n, const float *p_prev, const float *div, float *out_q,
                        float *out_p, float dt, int w, int h, cudaStream_t s, int y_base, int wrap)
{
    dim3 block(BX, BY), grid = stencil_grid(w, h, 1, 1);
    PFS_LAUNCH(project_pack_kernel, grid, block, 0, s, reinterpret_cast<const float2 *>(uv), p_n, p_prev, div,
               reinterpret_cast<float4 *>(out_q), reinterpret_cast<float4 *>(out_p), dt, w, h, y_base, wrap);
    return PFS_OK;
}

int launch_project_uv(const float *uv, const float *p_n, float *uv_out, float dt, int w, int h, cudaStream_t s, int y_base,
                      int wrap, float *vmax_out)
{
    const bool v4 = vec4_ok(w, uv, p_n, uv_out);
    dim3 block(BX, BY), grid = stencil_grid(w, h, v4 ? 4 : 1, 1);
    const float2 *in = reinterpret_cast<const float2 *>(uv);
    float2 *out = reinterpret_cast<float2 *>(uv_out);
    if (v4)
        PFS_LAUNCH(project_uv_kernel<4>, grid, block, 0, s, in, p_n, out, dt, w, h, y_base, wrap, vmax_out);
    else
        PFS_LAUNCH(project_uv_kernel<1>, grid, block, 0, s, in, p_n, out, dt, w, h, y_base, wrap, vmax_out);
    return PFS_OK;
}

int launch_subtract_gradient_aos(const float *vp_aos, float *out_aos, float dt, int w, int h, cudaStream_t s)
{
    dim3 block(64, 4), grid((w + 63) / 64, (h + 3) / 4);
    PFS_LAUNCH(subtract_gradient_aos_kernel, grid, block, 0, s, reinterpret_cast<const float4 *>(vp_aos), out_aos, dt,
               w, h);
    return PFS_OK;
}

int launch_pack_rgba8(const float *image_aos, unsigned char *out, size_t pixels, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((pixels + 255) / 256);
    PFS_LAUNCH(pack_rgba8_kernel, blocks, 256, 0, s, reinterpret_cast<const float4 *>(image_aos),
               reinterpret_cast<uchar4 *>(out), pixels);
    return PFS_OK;
}

int launch_plane_diff_norms(const float *a, const float *b, size_t cells, double *partials, int max_blocks, double *out4,
                            cudaStream_t s)
{
    int blocks = (int)std::min<size_t>((size_t)max_blocks, (cells + 255) / 256);
    if (blocks < 1) blocks = 1;
    PFS_LAUNCH(plane_diff_partial_kernel, blocks, 256, 0, s, a, b, cells, partials);
    PFS_LAUNCH(norms_final_kernel, 1, 32, 0, s, partials, blocks, out4);
    return PFS_OK;
}

int launch_step_norms(const float *vp_aos, const float *tmp_aos, size_t cells, double *partials, int max_blocks,
                      double *out4, cudaStream_t s)
{
    int blocks = (int)((cells + 255) / 256);
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    PFS_LAUNCH(norms_partial_kernel, blocks, 256, 0, s, reinterpret_cast<const float4 *>(vp_aos),
               reinterpret_cast<const float4 *>(tmp_aos), cells, partials);
    PFS_LAUNCH(norms_final_kernel, 1, 32, 0, s, partials, blocks, out4);
    return PFS_OK;
}

int sor_partial_blocks(int w, int h) { return ((w / 2 + 1 + 63) / 64) * ((h + 3) / 4); }

// One full red-black sweep (two launches) of p in place; the sum of squared updates lands in out4[0] (device).
int launch_sor_sweep(float *p, const float *rhs, int w, int h, float omega, double *partials, double *out4, cudaStream_t s)
{
    dim3 block(64, 4), grid((w / 2 + 1 + 63) / 64, (h + 3) / 4);
    PFS_LAUNCH(sor_half_sweep_kernel, grid, block, 0, s, p, rhs, w, h, omega, 0, partials, 0);
    PFS_LAUNCH(sor_half_sweep_kernel, grid, block, 0, s, p, rhs, w, h, omega, 1, partials, 1);
    if (out4) PFS_LAUNCH(norms_final_kernel, 1, 32, 0, s, partials, (int)(grid.x * grid.y), out4);
    return PFS_OK;
}

int launch_stochastic_force(float *u, float *v, int stride, float sigma, unsigned long long seed, unsigned step, int w,
                            int h, int row0, int y_base, cudaStream_t s)
{
    // same expression as oracle_stochastic_norm()
    const float norm = (float)(1.0 / sqrt(4.0 * 8.0 * (65536.0 * 65536.0 - 1.0) / 12.0));
    dim3 block(64, 4), grid((w + 63) / 64, (h + 3) / 4);
    PFS_LAUNCH(stochastic_force_kernel, grid, block, 0, s, u, v, stride, sigma, norm, (uint32_t)seed,
               (uint32_t)(seed >> 32), (uint32_t)step, w, h, row0, y_base);
    return PFS_OK;
}

int launch_advect_color(const float *image, float *out, const float *vel, int vel_stride, float dt, int iw, int ih, int vw,
                        int vh, cudaStream_t s, unsigned char *rgba8)
{
    // fluid.cpp:82-83 and the (dt/viw), (dt/vih) factors of :97-98, all binary32
    const float viw = (float)vw / (float)iw;
    const float vih = (float)vh / (float)ih;
    const float dt_over_viw = dt / viw;
    const float dt_over_vih = dt / vih;
    static const int R = env_rows_per_thread();
    dim3 block(64, 4), grid((iw + 63) / 64, (ih + 4 * R - 1) / (4 * R));
    const float4 *img = reinterpret_cast<const float4 *>(image);
    float4 *o4 = reinterpret_cast<float4 *>(out);
    const float rfiw = 1.0f / (float)iw, rfih = 1.0f / (float)ih;
    unsigned *b4 = reinterpret_cast<unsigned *>(rgba8);
#define PFS_COLOR(VS_, BY_, R_)                                                                                                   \
    PFS_LAUNCH((advect_color_kernel<VS_, BY_, R_>), grid, block, 0, s, img, o4, vel, dt_over_viw, dt_over_vih, viw, vih, iw, ih, vw, \
               rfiw, rfih, b4)
    const int variant = (vel_stride == 2 ? 0 : 4) + (rgba8 ? 2 : 0) + (R == 2 ? 1 : 0);
    switch (variant) {
    case 0: PFS_COLOR(2, false, 1); break;
    case 1: PFS_COLOR(2, false, 2); break;
    case 2: PFS_COLOR(2, true, 1); break;
    case 3: PFS_COLOR(2, true, 2); break;
    case 4: PFS_COLOR(4, false, 1); break;
    case 5: PFS_COLOR(4, false, 2); break;
    case 6: PFS_COLOR(4, true, 1); break;
    default: PFS_COLOR(4, true, 2); break;
    }
#undef PFS_COLOR
    return PFS_OK;
}

}  // namespace pfs
