/* The fast periodic wrap of pfs_internal.cuh (wrap_coord) against the reference expression
 * fmod(fmod(x, ext) + ext, ext) of fluid.cpp:48-49, bit for bit, on the CPU (fmodf is exact on both sides). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline uint32_t f2u(float f){uint32_t u; memcpy(&u,&f,4); return u;}
static inline float u2f(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static float ref_wrap(float x, float ext){ return fmodf(fmodf(x, ext) + ext, ext); }
static float fast_wrap(float x, float ext){
    float r = (x >= 0.0f && x < ext) ? x : fmodf(x, ext);
    float t = r + ext;
    if (t >= ext && t < ext + ext) return t - ext;
    if (t >= 0.0f && t < ext) return t;
    return fmodf(t, ext);
}
int main(void){
    const float exts[] = {1.f, 2.f, 3.f, 29.f, 37.f, 64.f, 100.f, 512.f, 768.f, 1024.f, 4096.f, 16384.f, 268435456.f};
    uint64_t st = 0x9E3779B97F4A7C15ULL; long tested = 0, bad = 0;
    for (unsigned e = 0; e < sizeof(exts)/sizeof(exts[0]); e++) {
        const float W = exts[e];
        const float edge[] = {0.f, -0.f, W, -W, W*2, -W*2, nextafterf(W, 0.f), nextafterf(W, 2*W), nextafterf(0.f, 1.f), nextafterf(0.f, -1.f),
                              -1e-30f, 1e-30f, -W + nextafterf(0.f,1.f), W*0.5f, -W*0.5f, 1e30f, -1e30f, 3.4e38f, -3.4e38f};
        for (unsigned k = 0; k < sizeof(edge)/sizeof(edge[0]); k++) { tested++; if (f2u(ref_wrap(edge[k],W)) != f2u(fast_wrap(edge[k],W))) { bad++; printf("edge mismatch x=%a W=%a\n", edge[k], W);} }
        for (long i = 0; i < 4000000; i++) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            float x;
            switch (i & 3) {
            case 0: x = (float)((double)(st >> 11) / 9007199254740992.0) * W; break;                    /* inside the domain */
            case 1: x = ((float)((double)(st >> 11) / 9007199254740992.0) * 6.0f - 3.0f) * W; break;      /* a few periods around it */
            case 2: x = u2f((uint32_t)st); if (x != x) x = 0.5f; break;                                   /* any finite or infinite bit pattern */
            default: x = W - (float)((double)(st >> 40) * 1e-7); break;                                   /* just below the upper edge */
            }
            float a = ref_wrap(x, W), b = fast_wrap(x, W);
            tested++;
            if (f2u(a) != f2u(b) && !(a != a && b != b)) { bad++; if (bad < 10) printf("mismatch x=%a W=%a ref=%a fast=%a\n", x, W, a, b); }
        }
    }
    printf("wrap values tested %ld  mismatches %ld\n", tested, bad);
    return bad != 0;
}
