// dispatch_cost.cu -- what the instructions that accompany the packed FP32 arithmetic of the sweep kernels cost when they
// are issued between FADD2s: N independent FADD2 chains per thread plus ONE candidate instruction per FADD2, 4 warps per
// scheduler.  Reported: cycles per (FADD2 + candidate) pair per scheduler; FADD2 alone is 2.0.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dispatch_cost dispatch_cost.cu && ./dispatch_cost
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int CH = 8, ITERS = 2048, UNROLL = 4;

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

enum Kind { NONE, FMNMX2, FMNMX3, FMNMX3_4CH, VIMNMX3, FSETP_OR, LOP3, IADD3, MOV32, SHFL, LDS128, FADD1, FMUL1, PRMT, FMNMX3_ABS_4CH, N_KINDS };
const char *NAMES[] = {"(FADD2 alone)", "FMNMX (2-input, 1 chain)", "FMNMX3 (1 chain)", "FMNMX3 (4 chains)", "VIMNMX.U32 2-input (4 chains)", "FSETP.LT.OR (pred chain)",
                       "LOP3 (4 chains)", "IADD3 (4 chains)", "MOV", "SHFL.UP", "LDS.128", "FADD scalar", "FMUL scalar", "PRMT", "FMNMX3 |a|,|b| (4 chains)"};

template <int K>
__global__ void __launch_bounds__(512) cost_kernel(float *out, float a_in, int iters, long long *cycles)
{
    __shared__ float4 sm[512];
    const float a = a_in + (float)threadIdx.x * 1e-7f;
    unsigned long long p[CH];
    float x[CH], m[4] = {1e30f, 1e30f, 1e30f, 1e30f};
    unsigned um[4] = {~0u, ~0u, ~0u, ~0u};
    int acc[4] = {0, 1, 2, 3};
    int pred = 0;
    float4 ld = make_float4(0, 0, 0, 0);
    sm[threadIdx.x] = make_float4(a, a, a, a);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CH; c++) {
        x[c] = (float)(threadIdx.x + c) * 1e-3f + 1.f;
        p[c] = pk(x[c], x[c] + 1.f);
    }
    const unsigned long long pa = pk(a, a + 1.f);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(&sm[threadIdx.x]);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(pa));
                if constexpr (K == FMNMX2) asm volatile("min.f32 %0, %0, %1;" : "+f"(m[0]) : "f"(x[c]));
                if constexpr (K == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[0]) : "f"(x[c]), "f"(x[(c + 1) % CH]));
                if constexpr (K == FMNMX3_4CH) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[c & 3]) : "f"(x[c]), "f"(x[(c + 1) % CH]));
                if constexpr (K == FMNMX3_ABS_4CH) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[c & 3]) : "f"(fabsf(x[c])), "f"(fabsf(x[(c + 1) % CH])));
                if constexpr (K == VIMNMX3) asm volatile("min.u32 %0, %0, %1;" : "+r"(um[c & 3]) : "r"(__float_as_uint(x[c])));
                if constexpr (K == FSETP_OR) asm volatile("{ .reg .pred q; setp.lt.f32 q, %1, %2; selp.s32 %0, 1, %0, q; }" : "+r"(pred) : "f"(x[c]), "f"(a));
                if constexpr (K == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x80;" : "+r"(acc[c & 3]) : "r"(__float_as_int(x[c])), "r"(__float_as_int(x[(c + 1) % CH])));
                if constexpr (K == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(acc[c & 3]) : "r"(__float_as_int(x[c])));
                if constexpr (K == MOV32) asm volatile("mov.b32 %0, %1;" : "=f"(x[c]) : "f"(m[c & 3]));
                if constexpr (K == SHFL) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+f"(x[c]));
                if constexpr (K == LDS128) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ld.x), "=f"(ld.y), "=f"(ld.z), "=f"(ld.w) : "r"(sbase + (c & 1) * 16));
                if constexpr (K == FADD1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(a));
                if constexpr (K == FMUL1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(a));
                if constexpr (K == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(acc[c & 3]) : "r"(__float_as_int(x[c])));
            }
        }
    }
    const long long t1 = clock64();
    float s = m[0] + m[1] + m[2] + m[3] + (float)(acc[0] + acc[1] + acc[2] + acc[3] + pred) + (float)(um[0] ^ um[1] ^ um[2] ^ um[3]) + ld.x + ld.y + ld.z + ld.w;
#pragma unroll
    for (int c = 0; c < CH; c++) s += x[c] + (float)(p[c] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int K>
int run(int sms, float *out, long long *d_cycles)
{
    const int w = 4, threads = w * 4 * 32;
    cost_kernel<K><<<sms, threads>>>(out, 1.0001f, 16, d_cycles);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    cost_kernel<K><<<sms, threads>>>(out, 1.0001f, ITERS, d_cycles);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost));
    const double pairs = (double)ITERS * UNROLL * CH * w;
    printf("  FADD2 + %-28s: %.3f cycles per pair per scheduler  (candidate costs %.2f)\n", NAMES[K], (double)cyc / pairs, (double)cyc / pairs - 2.0);
    return 0;
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *out;
    long long *d_cycles;
    CK(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(float)));
    CK(cudaMalloc(&d_cycles, sizeof(long long)));
    if (run<NONE>(sms, out, d_cycles) || run<FMNMX2>(sms, out, d_cycles) || run<FMNMX3>(sms, out, d_cycles) || run<FMNMX3_4CH>(sms, out, d_cycles) ||
        run<FMNMX3_ABS_4CH>(sms, out, d_cycles) || run<VIMNMX3>(sms, out, d_cycles) || run<FSETP_OR>(sms, out, d_cycles) || run<LOP3>(sms, out, d_cycles) ||
        run<IADD3>(sms, out, d_cycles) || run<MOV32>(sms, out, d_cycles) || run<SHFL>(sms, out, d_cycles) || run<LDS128>(sms, out, d_cycles) ||
        run<FADD1>(sms, out, d_cycles) || run<FMUL1>(sms, out, d_cycles) || run<PRMT>(sms, out, d_cycles))
        return 1;
    return 0;
}
