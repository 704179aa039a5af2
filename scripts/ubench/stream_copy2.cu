// stream_copy2.cu -- the column-streaming pattern of the fused sweep kernels as a pure copy, second look: how the number
// of concurrent row fronts (chunks), the bytes a warp moves per row (strip width) and the prefetch depth trade off.
// Every work item = one warp streaming a strip of CPL*64 cells (8 B each) down `rows` rows; all items are resident at once.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct P { const float *in; float *out; int w, h; long long pitch; int n_strips, n_chunks, chunk_rows, do_store; };

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

template <int SLOTS, int CPL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) copy_kernel(const P p)
{
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS + warp;
    if (item >= p.n_strips * p.n_chunks) return;
    constexpr int SLOT = CPL * 128;                         // floats per ring slot
    float *my = smem + (size_t)warp * SLOTS * SLOT + lane * 4;
    const int strip = item % p.n_strips, chunk = item / p.n_strips;
    const long long x = (long long)strip * CPL * 128 + lane * 4;          // float offset within a row of the lane's first chunk
    const int y0 = chunk * p.chunk_rows, L = min(p.chunk_rows, p.h - y0);
    const float *ld = p.in + (long long)y0 * p.pitch + x;
    float *op = p.out + (long long)y0 * p.pitch + x;
    constexpr int PRE = SLOTS - 2;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int s = 0; s < PRE; s++) {
        if (s < L) {
#pragma unroll
            for (int c = 0; c < CPL; c++) cp_async16(my + (s % SLOTS) * SLOT + c * 128, ld + c * 128);
            ld += p.pitch;
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    for (int s = 0; s < L; s++) {
        const int n = s + PRE;
        if (n < L) {
#pragma unroll
            for (int c = 0; c < CPL; c++) cp_async16(my + (n % SLOTS) * SLOT + c * 128, ld + c * 128);
            ld += p.pitch;
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group %0;\n" ::"n"(PRE) : "memory");
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const float4 a = *reinterpret_cast<const float4 *>(my + (s % SLOTS) * SLOT + c * 128);
            if (p.do_store) *reinterpret_cast<float4 *>(op + c * 128) = a; else { acc.x += a.x; acc.y += a.y; }
        }
        op += p.pitch;
    }
    if (!p.do_store && acc.x == 12345.678f) p.out[0] = acc.x + acc.y;
}

template <int SLOTS, int CPL, int WARPS>
void run(P p, int fronts, int sms)
{
    p.n_strips = p.w * 2 / (CPL * 128);
    p.n_chunks = fronts;
    p.chunk_rows = (p.h + fronts - 1) / fronts;
    const int total = p.n_strips * p.n_chunks;
    const int blocks = (total + WARPS - 1) / WARPS;
    const size_t smem = (size_t)WARPS * SLOTS * CPL * 512;
    CK(cudaFuncSetAttribute(copy_kernel<SLOTS, CPL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, copy_kernel<SLOTS, CPL, WARPS>, WARPS * 32, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) copy_kernel<SLOTS, CPL, WARPS><<<blocks, WARPS * 32, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; i++) copy_kernel<SLOTS, CPL, WARPS><<<blocks, WARPS * 32, smem>>>(p);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double bytes = (double)p.w * p.h * 8 * (p.do_store ? 2 : 1);
    printf("  ring %d  bytes/row/warp %4d  strips %3d  fronts %3d (rows %4d)  warps %5d (%s one wave: %d CTAs/SM fit)  %s : %7.1f us %6.0f GB/s\n",
           SLOTS, CPL * 512, p.n_strips, fronts, p.chunk_rows, total, blocks <= per_sm * sms ? "" : "NOT", per_sm, p.do_store ? "copy" : "read", ms * 1e3,
           bytes / (ms * 1e-3) / 1e9);
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int w = 4096, h = 4096;
    float *in, *out;
    CK(cudaMalloc(&in, (size_t)2 * w * h * 4 + 4096));
    CK(cudaMalloc(&out, (size_t)2 * w * h * 4 + 4096));
    CK(cudaMemset(in, 0, (size_t)2 * w * h * 4));
    P p; p.in = in; p.out = out; p.w = w; p.h = h; p.pitch = 2 * w;
    printf("4096 x 4096 cells of 8 bytes (32 KiB rows); %d SMs\n", sms);
    for (int store : {1, 0}) {
        p.do_store = store;
        for (int fronts : {4, 8, 16, 32, 64}) {
            run<8, 1, 4>(p, fronts, sms);      // 512 B per row per warp, 64 strips
            run<8, 2, 4>(p, fronts, sms);      // 1 KiB, 32 strips
            run<8, 4, 4>(p, fronts, sms);      // 2 KiB, 16 strips
            run<4, 2, 4>(p, fronts, sms);      // 1 KiB, shallower ring
            run<16, 2, 4>(p, fronts, sms);     // 1 KiB, deeper ring
        }
    }
    return 0;
}
