// Do the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2) round subnormal results like the scalar ones?
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(unsigned long long r) { float2 a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a; }
__global__ void k(const float *x, const float *y, float *out, int n, float alpha, float nz)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long d;
    // packed fma(x, alpha, -0), packed add(x, y), packed mul(x, alpha)
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(x[i], y[i])), "l"(pk(alpha, alpha)), "l"(pk(nz, nz)));
    float2 f = upk(d);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(x[i], y[i])), "l"(pk(y[i], x[i])));
    float2 a = upk(d);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(x[i], y[i])), "l"(pk(alpha, alpha)));
    float2 m = upk(d);
    out[8 * i + 0] = f.x; out[8 * i + 1] = __fmaf_rn(x[i], alpha, nz);
    out[8 * i + 2] = a.x; out[8 * i + 3] = __fadd_rn(x[i], y[i]);
    out[8 * i + 4] = m.x; out[8 * i + 5] = __fmul_rn(x[i], alpha);
    out[8 * i + 6] = f.y; out[8 * i + 7] = __fmaf_rn(y[i], alpha, nz);
}
int main()
{
    const int n = 1 << 16;
    float *hx = (float *)malloc(n * 4), *hy = (float *)malloc(n * 4), *ho = (float *)malloc(n * 32);
    srand(1);
    for (int i = 0; i < n; i++) {
        unsigned a = (rand() & 0x7fffff) | ((unsigned)(rand() % 12) << 23) | ((unsigned)(rand() & 1) << 31);   // exponents 0..11: subnormal and tiny
        unsigned b = (rand() & 0x7fffff) | ((unsigned)(rand() % 12) << 23) | ((unsigned)(rand() & 1) << 31);
        memcpy(&hx[i], &a, 4); memcpy(&hy[i], &b, 4);
    }
    float *x, *y, *o;
    cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&o, n * 32);
    cudaMemcpy(x, hx, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(y, hy, n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(x, y, o, n, 0.03f, -0.0f);
    cudaMemcpy(ho, o, n * 32, cudaMemcpyDeviceToHost);
    int bad[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int q = 0; q < 4; q++)
            if (memcmp(&ho[8 * i + 2 * q], &ho[8 * i + 2 * q + 1], 4)) {
                if (bad[q]++ < 3) printf("  op %d: packed %a scalar %a (x %a y %a)\n", q, ho[8 * i + 2 * q], ho[8 * i + 2 * q + 1], hx[i], hy[i]);
            }
    printf("mismatches packed vs scalar: fma %d add %d mul %d fma.y %d of %d\n", bad[0], bad[1], bad[2], bad[3], n);
    return 0;
}
