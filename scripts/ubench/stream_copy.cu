// stream_copy.cu -- how fast can the access pattern of the fused sweep kernels move data, with no arithmetic at all?
// One warp streams a strip of 128 cells x 8 bytes (1 KB per row, two 16-byte chunks per lane) down a chunk of rows:
// cp.async ring in shared memory -> registers -> 16-byte stores of the inner `out_cols` cells, exactly like
// sweeps_packed.cu.  Knobs: ring depth, resident warps per SM, chunk height, row pitch padding, plain loads instead of
// cp.async, stores on/off.  Prints GB/s of (bytes read + bytes written) per configuration.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_copy stream_copy.cu && ./stream_copy
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct P {
    const float *in;
    float *out;
    int w, h;              // cells per row, rows
    long long pitch;       // floats per row (>= 2*w)
    int strip_out, halo, n_strips, n_chunks, chunk_rows;
    int do_store;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

template <int SLOTS, bool ASYNC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) copy_kernel(const P p)
{
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS + warp;
    if (item >= p.n_strips * p.n_chunks) return;
    float *my = smem + (size_t)warp * SLOTS * 256 + lane * 4;
    const int strip = item % p.n_strips, chunk = item / p.n_strips;
    const int x0 = strip * p.strip_out, xc = x0 - p.halo + 4 * lane;
    int xlo = xc % p.w; if (xlo < 0) xlo += p.w;
    int xhi = xlo + 2; if (xhi >= p.w) xhi -= p.w;
    const int xend = min(x0 + p.strip_out, p.w);
    const bool st_lo = xc >= x0 && xc < xend, st_hi = xc + 2 >= x0 && xc + 2 < xend;
    const int y0 = chunk * p.chunk_rows, L = min(p.chunk_rows, p.h - y0);
    const float *ld = p.in + (long long)y0 * p.pitch;
    float *op = p.out + (long long)y0 * p.pitch + 2 * (long long)xc;
    float4 acc = make_float4(0, 0, 0, 0);
    constexpr int PRE = SLOTS - 2;
    if constexpr (ASYNC) {
        for (int s = 0; s < PRE; s++) {
            if (s < L) { cp_async16(my + (s % SLOTS) * 256, ld + 2 * xlo); cp_async16(my + (s % SLOTS) * 256 + 128, ld + 2 * xhi); ld += p.pitch; }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        for (int s = 0; s < L; s++) {
            const int n = s + PRE;
            if (n < L) { cp_async16(my + (n % SLOTS) * 256, ld + 2 * xlo); cp_async16(my + (n % SLOTS) * 256 + 128, ld + 2 * xhi); ld += p.pitch; }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group %0;\n" ::"n"(PRE) : "memory");
            const float4 a = *reinterpret_cast<const float4 *>(my + (s % SLOTS) * 256);
            const float4 b = *reinterpret_cast<const float4 *>(my + (s % SLOTS) * 256 + 128);
            if (p.do_store) {
                if (st_lo) *reinterpret_cast<float4 *>(op) = a;
                if (st_hi) *reinterpret_cast<float4 *>(op + 4) = b;
            } else { acc.x += a.x + b.x; acc.y += a.y + b.y; }
            op += p.pitch;
        }
    } else {
        // register ring: PRE rows of plain 16-byte loads in flight
        float4 ra[PRE], rb[PRE];
#pragma unroll
        for (int s = 0; s < PRE; s++) {
            if (s < L) { ra[s] = __ldcg(reinterpret_cast<const float4 *>(ld + 2 * xlo)); rb[s] = __ldcg(reinterpret_cast<const float4 *>(ld + 2 * xhi)); ld += p.pitch; }
        }
        for (int sb = 0; sb < L; sb += PRE) {
#pragma unroll
            for (int u = 0; u < PRE; u++) {
                const int s = sb + u;
                const float4 a = ra[u], b = rb[u];
                if (s + PRE < L) { ra[u] = __ldcg(reinterpret_cast<const float4 *>(ld + 2 * xlo)); rb[u] = __ldcg(reinterpret_cast<const float4 *>(ld + 2 * xhi)); ld += p.pitch; }
                if (s < L) {
                    if (p.do_store) {
                        if (st_lo) *reinterpret_cast<float4 *>(op) = a;
                        if (st_hi) *reinterpret_cast<float4 *>(op + 4) = b;
                    } else { acc.x += a.x + b.x; acc.y += a.y + b.y; }
                    op += p.pitch;
                }
            }
        }
    }
    if (!p.do_store && acc.x == 12345.678f) p.out[0] = acc.x + acc.y;
}

__global__ void __launch_bounds__(256) linear_copy(const float4 *__restrict__ a, float4 *__restrict__ b, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) b[i] = a[i];
}

template <int SLOTS, bool ASYNC, int WARPS>
void run(const char *name, P p, int warps_per_sm, int sms)
{
    const int total = p.n_strips * p.n_chunks;
    const int blocks = (total + WARPS - 1) / WARPS;
    // occupancy control through dynamic shared memory: each CTA claims 1/ctas_per_sm of the SM's shared memory
    const int ctas_per_sm = warps_per_sm / WARPS;
    size_t smem = (size_t)WARPS * SLOTS * 1024;
    const size_t want = (size_t)(200 * 1024) / ctas_per_sm / 1024 * 1024;
    if (want > smem) smem = want;
    CK(cudaFuncSetAttribute(copy_kernel<SLOTS, ASYNC, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) copy_kernel<SLOTS, ASYNC, WARPS><<<blocks, WARPS * 32, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; i++) copy_kernel<SLOTS, ASYNC, WARPS><<<blocks, WARPS * 32, smem>>>(p);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double bytes = (double)p.w * p.h * 8 * (p.do_store ? 2 : 1);
    printf("  %-46s warps/SM %2d chunk_rows %4d items %5d pitch %6lld B : %7.1f us  %6.0f GB/s\n", name, warps_per_sm, p.chunk_rows, total,
           p.pitch * 4, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int w = 4096, h = 4096;
    float *in, *out;
    const long long max_pitch = 2 * w + 8192 + 64;
    CK(cudaMalloc(&in, (size_t)max_pitch * h * 4 + 4096));
    CK(cudaMalloc(&out, (size_t)max_pitch * h * 4 + 4096));
    CK(cudaMemset(in, 0, (size_t)max_pitch * h * 4));
    auto mk = [&](int warps_per_sm, long long pitch, int halo, int do_store, int forced_rows) {
        P p;
        p.in = in; p.out = out; p.w = w; p.h = h; p.pitch = pitch; p.halo = halo; p.strip_out = 128 - 2 * halo;
        p.n_strips = (w + p.strip_out - 1) / p.strip_out;
        int chunks = sms * warps_per_sm / p.n_strips;
        int rows = forced_rows > 0 ? forced_rows : (h + chunks - 1) / chunks;
        int n = (h + rows - 1) / rows;
        p.chunk_rows = (h + n - 1) / n; p.n_chunks = (h + p.chunk_rows - 1) / p.chunk_rows; p.do_store = do_store;
        return p;
    };
    printf("stream copy of a 4096 x 4096 (u,v) plane (128 MiB in, 128 MiB out), %d SMs\n", sms);
    // (a) chunk height scan (exact heights, 32 warps per SM so that every case is one resident wave)
    for (int rows : {512, 384, 256, 192, 160, 144, 136, 132, 130, 129, 128, 127, 126, 124, 120, 112, 96, 80, 64, 48, 32}) {
        P p = mk(32, 2 * w, 6, 1, rows);
        p.chunk_rows = rows; p.n_chunks = (h + rows - 1) / rows;
        run<8, true, 4>("ring 8, load+store, exact chunk height", p, 32, sms);
    }
    // (a2) row pitch scan at 128-row chunks
    for (long long pad : {0LL, 64LL, 256LL, 1024LL, 2048LL, 4096LL, 8192LL + 64}) {
        run<8, true, 4>("ring 8, load+store, padded pitch", mk(8, 2 * w + pad, 6, 1, 0), 8, sms);
    }
    // (b) a plain grid-stride copy in the same harness (what the memory system does for a linear stream)
    {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const size_t n4 = (size_t)w * h * 2 / 4;
        for (int i = 0; i < 3; i++) linear_copy<<<sms * 16, 256>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<float4 *>(out), n4);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 10; i++) linear_copy<<<sms * 16, 256>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<float4 *>(out), n4);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  linear grid-stride float4 copy: %7.1f us  %6.0f GB/s\n", ms / 10 * 1e3, (double)w * h * 16 / (ms / 10 * 1e-3) / 1e9);
    }
    return 0;
}
