// fp32_rates.cu -- what one SM sub-partition of a B200 can issue per cycle for the FP32 instruction forms the two sweep
// kernels are made of (scalar FADD/FMUL/FFMA with register and immediate operands, packed FADD2/FMUL2/FFMA2, and mixes
// with the ALU-pipe / shuffle instructions that accompany them).  Every test is N_CHAINS independent dependency chains per
// thread, so that with W warps per scheduler the result is throughput- and not latency-bound once W*N_CHAINS is large.
// Output: cycles per warp-instruction per scheduler (SM sub-partition) = elapsed SM cycles * 4 schedulers / (warp-instructions per SM).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_rates fp32_rates.cu && ./fp32_rates
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int CH = 8;        // independent chains per thread
constexpr int ITERS = 4096;  // loop trips; each trip issues CH * UNROLL instructions of the kind under test
constexpr int UNROLL = 4;

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

enum Kind { FADD_RR, FMUL_RR, FFMA_RRR, FFMA_RIR, FADD_RI, FADD2_RR, FMUL2_RR, FFMA2_RRR, FFMA2_BCAST, MIX_FADD2_FMNMX, MIX_FADD2_SHFL,
            MIX_FADD_FMNMX, MIX_FFMA2_IADD, N_KINDS };
const char *NAMES[] = {"FADD  r,r", "FMUL  r,r", "FFMA  r,r,r", "FFMA  r,imm,r", "FADD  r,imm", "FADD2 rr,rr", "FMUL2 rr,rr", "FFMA2 rr,rr,rr",
                       "FFMA2 rr,bcast,bcast", "FADD2 + FMNMX3 (1:1)", "FADD2 + SHFL (2:1)", "FADD + FMNMX3 (1:1)", "FFMA2 + IADD3 (1:1)"};

template <int K>
__global__ void __launch_bounds__(1024) rate_kernel(float *out, float a_in, float b_in, int iters, long long *cycles)
{
    // per-thread operands: uniform-register forms would not be what the sweep kernels issue
    const float a = a_in + (float)threadIdx.x * 1e-7f, b = b_in + (float)threadIdx.x * 1e-7f;
    float x[CH], y[CH];
    unsigned long long p[CH];
    float m = 0.f;
    int ia = threadIdx.x;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        x[c] = (float)(threadIdx.x + c) * 1e-3f;
        y[c] = (float)(threadIdx.x * 3 + c) * 1e-3f;
        p[c] = pk(x[c], y[c]);
    }
    const unsigned long long pa = pk(a, a), pb = pk(b, b);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                if constexpr (K == FADD_RR) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(a));
                if constexpr (K == FMUL_RR) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(a));
                if constexpr (K == FFMA_RRR) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(a), "f"(b));
                if constexpr (K == FFMA_RIR) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, %1;" : "+f"(x[c]) : "f"(b));
                if constexpr (K == FADD_RI) asm volatile("add.rn.f32 %0, %0, 0f3A83126F;" : "+f"(x[c]));
                if constexpr (K == FADD2_RR) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(pa));
                if constexpr (K == FMUL2_RR) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(pa));
                if constexpr (K == FFMA2_RRR) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(pa), "l"(p[(c + 1) % CH]));
                if constexpr (K == FFMA2_BCAST) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(pa), "l"(pb));
                if constexpr (K == MIX_FADD2_FMNMX) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(pa));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(x[c]), "f"(y[c]));
                }
                if constexpr (K == MIX_FADD2_SHFL) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(pa));
                    if (c & 1) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+f"(y[c]));
                }
                if constexpr (K == MIX_FADD_FMNMX) {
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(a));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(y[c]), "f"(b));
                }
                if constexpr (K == MIX_FFMA2_IADD) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(pa), "l"(pb));
                    asm volatile("add.s32 %0, %0, %1;" : "+r"(ia) : "r"(it));
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = m + (float)ia;
#pragma unroll
    for (int c = 0; c < CH; c++) s += x[c] + y[c] + (float)(p[c] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int K>
int run(int warps_per_sched, int sms, float *out, long long *d_cycles)
{
    const int threads = warps_per_sched * 4 * 32;          // one CTA per SM
    rate_kernel<K><<<sms, threads>>>(out, 1.0001f, 0.5f, 16, d_cycles);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    rate_kernel<K><<<sms, threads>>>(out, 1.0001f, 0.5f, ITERS, d_cycles);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost));
    const int per_item = (K == MIX_FADD2_FMNMX || K == MIX_FADD_FMNMX || K == MIX_FFMA2_IADD) ? 2 : 1;
    const double instr = (double)ITERS * UNROLL * CH * per_item * warps_per_sched + (K == MIX_FADD2_SHFL ? (double)ITERS * UNROLL * CH / 2 * warps_per_sched : 0.0);
    printf("  %-22s warps/scheduler %d : %.3f cycles per warp-instruction per scheduler\n", NAMES[K], warps_per_sched, (double)cyc / instr);
    return 0;
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *out;
    long long *d_cycles;
    CK(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(float) * 2));
    CK(cudaMalloc(&d_cycles, sizeof(long long)));
    printf("SMs %d; %d chains per thread, clock64 of warp 0 of CTA 0\n", sms, CH);
    for (int w : {1, 2, 3, 4, 8}) {
        if (run<FADD_RR>(w, sms, out, d_cycles)) return 1;
        if (run<FMUL_RR>(w, sms, out, d_cycles)) return 1;
        if (run<FFMA_RRR>(w, sms, out, d_cycles)) return 1;
        if (run<FFMA_RIR>(w, sms, out, d_cycles)) return 1;
        if (run<FADD_RI>(w, sms, out, d_cycles)) return 1;
        if (run<FADD2_RR>(w, sms, out, d_cycles)) return 1;
        if (run<FMUL2_RR>(w, sms, out, d_cycles)) return 1;
        if (run<FFMA2_RRR>(w, sms, out, d_cycles)) return 1;
        if (run<FFMA2_BCAST>(w, sms, out, d_cycles)) return 1;
        if (run<MIX_FADD2_FMNMX>(w, sms, out, d_cycles)) return 1;
        if (run<MIX_FADD2_SHFL>(w, sms, out, d_cycles)) return 1;
        if (run<MIX_FADD_FMNMX>(w, sms, out, d_cycles)) return 1;
        if (run<MIX_FFMA2_IADD>(w, sms, out, d_cycles)) return 1;
        printf("\n");
    }
    return 0;
}
