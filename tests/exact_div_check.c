// Exhaustive worst-case search for q1 = fma(a - b*q0, y, q0), q0 = RN(a*y), y = RN(1/b):
// all 2^23 significands B of b, all quotients whose distance to a rounding midpoint is
// |A' - B*M| <= 4 units of 2^-47 (the analysis shows only < 3 can possibly fail).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float u2f(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static inline uint32_t f2u(float f){uint32_t u; memcpy(&u,&f,4); return u;}
static uint32_t inv24(uint32_t b){ // inverse of odd b mod 2^32 (Newton)
    uint32_t x=b; for(int i=0;i<5;i++) x*=2-b*x; return x; }
int main(){
    long tested=0,bad1=0,bad2=0;
    for(uint32_t B=1u<<23; B<(1u<<24); B++){
        float b = ldexpf((float)B,-23);
        float y = 1.0f/b;
        int t = __builtin_ctz(B); if (t>3) t=3;   // only j up to 4 => t<=2 matters; cap
        for(int j=-4;j<=4;j++){
            if(!j) continue;
            int tt = __builtin_ctz(B);
            if (tt>24) tt=24;
            if (tt>2) { if (j % (1<<3)) { /* need 2^tt | j, impossible for |j|<=4 when tt>=3 */ continue; } }
            if (j % (1<<tt)) continue;
            uint32_t Bo = B>>tt; int jo = j/(1<<tt);
            uint32_t modbits = 24-tt; uint32_t mask = (modbits==32)?0xffffffffu:((1u<<modbits)-1);
            uint32_t M0 = ((uint32_t)((int64_t)jo * (int64_t)inv24(Bo))) & mask;
            // candidates M = M0 + i*2^modbits in [2^24, 2^25), odd
            for(uint64_t M=M0; M<(1ull<<25); M+= (1ull<<modbits)){
                if (M < (1ull<<24) || !(M&1)) continue;
                uint64_t P = (uint64_t)B*M; int64_t Ai = (int64_t)P - j;
                if (Ai & ((1ll<<24)-1)) continue;
                uint64_t A = (uint64_t)Ai>>24; float a;
                if (A < (1ull<<24)) { if (A < (1ull<<23)) continue; a = ldexpf((float)A,-23); }
                else { if (A&1) continue; if ((A>>1) >= (1ull<<24)) continue; a = ldexpf((float)(A>>1),-22); }
                float want = a/b;
                float q0=a*y; float e=fmaf(-b,q0,a); float q1=fmaf(e,y,q0);
                float e1=fmaf(-b,q1,a); float q2=fmaf(e1,y,q1);
                tested++;
                if (f2u(q1)!=f2u(want)) { bad1++; if(bad1<10) printf("1-iter FAIL a=%a b=%a q1=%a want=%a j=%d\n",a,b,q1,want,j);}                
                if (f2u(q2)!=f2u(want)) { bad2++; if(bad2<10) printf("2-iter FAIL a=%a b=%a\n",a,b);}                
            }
        }
    }
    printf("worst-case quotients tested %ld  one-iteration mismatches %ld  two-iteration mismatches %ld\n",tested,bad1,bad2);
    /* random numerators over the whole guarded exponent range x random betas of the form 1+4*alpha */
    uint64_t st = 88172645463325252ULL; long rtested = 0, rbad = 0;
    for (int bi = 0; bi < 2000; bi++) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        float alpha = u2f(0x30000000u + (uint32_t)(st % 0x10800000u));
        float b = (float)(1.0 + 4.0 * (double)alpha);
        if (!(b <= 1048576.0f)) continue;
        float y = 1.0f / b;
        for (int i = 0; i < 10000; i++) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            uint32_t m = (uint32_t)(st & 0x7fffffu); int e = 127 - 96 + (int)((st >> 23) % 193); uint32_t sg = (uint32_t)((st >> 40) & 1) << 31;
            float a = u2f(sg | ((uint32_t)e << 23) | m);
            float q0 = a * y; float r = fmaf(-b, q0, a); float q1 = fmaf(r, y, q0);
            rtested++;
            if (f2u(q1) != f2u(a / b)) rbad++;
        }
    }
    printf("random quotients tested %ld  mismatches %ld\n", rtested, rbad);
    return (bad1 != 0 || rbad != 0);
}
