/* Independent check of the two-instruction constant division of sweeps_packed.cu (div_const_two2):
 *   zh = RN(1/b), zl = RN(1/b - zh),  q = RN(a*zh + RN(a*zl))
 * For a divisor b it prints whether q == RN(a/b) for ALL 2^23 significands of a in [1,2) (every other a is one of
 * these times a power of two, and the scaling is exact while a*zl stays normal).  The library takes the same
 * decision per divisor at run time (div2_constants); tests/test_exact_division.py compares the two on random
 * divisors, including divisors for which the two-instruction form is WRONG for some a (about 1 in 30).
 *   usage: div2_check <count> <seed>   -> lines "<hex bits of b> <1|0> <number of wrong quotients, capped at 9>"
 *          div2_check bits <hex> ...   -> the same lines for the given divisors
 * Exact reference: (float)((double)a/(double)b) -- rounding the double quotient once more is innocuous for a
 * quotient of two 24-bit numbers (53 >= 2*24+2).  Compile with -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int wrong_quotients(float b)
{
    const double C = 1.0 / (double)b;
    const float zh = (float)C;
    const float zl = (float)(C - (double)zh);
    int bad = 0;
    for (uint32_t m = 0; m < (1u << 23) && bad < 9; m++) {
        const uint32_t bits = 0x3f800000u | m;
        float a;
        memcpy(&a, &bits, 4);
        volatile float t = a * zl;
        const float q = fmaf(a, zh, t);
        if (q != (float)((double)a / (double)b)) bad++;
    }
    return bad;
}

int main(int argc, char **argv)
{
    if (argc > 2 && strcmp(argv[1], "bits") == 0) {
        for (int k = 2; k < argc; k++) {
            const uint32_t bits = (uint32_t)strtoul(argv[k], NULL, 16);
            float b;
            memcpy(&b, &bits, 4);
            const int bad = wrong_quotients(b);
            printf("%08x %d %d\n", bits, bad == 0, bad);
        }
        return 0;
    }
    const int count = argc > 1 ? atoi(argv[1]) : 30;
    uint64_t x = argc > 2 ? strtoull(argv[2], NULL, 10) : 1;
    for (int k = 0; k < count; k++) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;            /* LCG: reproducible everywhere */
        const uint32_t mant = (uint32_t)(x >> 41);                          /* 23 bits */
        const uint32_t expo = 127u + (uint32_t)((x >> 20) % 3);             /* binades [1,2), [2,4), [4,8) */
        const uint32_t bits = (expo << 23) | mant;
        float b;
        memcpy(&b, &bits, 4);
        const int bad = wrong_quotients(b);
        printf("%08x %d %d\n", bits, bad == 0, bad);
    }
    return 0;
}
