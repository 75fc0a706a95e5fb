/* Exhaustive check of the constant-division fast path of csrc/nig_math.cuh (DivFast):
 *   q = RN(x*rc); r = fma(-q, c, x); q' = fma(r, rc, q),  rc = RN(1/c)
 * against IEEE x / c for ALL 2^32 binary32 inputs and every constant the kernels divide by.
 * Result (gcc 13.3, x86-64 FMA): for c in {5, 20, 50, 100, 1000, 418000} the only mismatches are -0, +-inf and
 * finite |x| < 2^-122, i.e. the fast path is exact on the guarded domain 2^-120 <= |x| <= FLT_MAX.
 *   gcc -O2 -mfma -fopenmp -ffp-contract=off tools/verify_cdiv.c -o /tmp/verify_cdiv -lm && /tmp/verify_cdiv
 * (about 5 minutes on 8 cores) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(void)
{
    const float cs[] = {5.f, 20.f, 50.f, 100.f, 1000.f, 418000.f};
    int fail = 0;
    for (int ci = 0; ci < 6; ci++) {
        const float c = cs[ci], rc = 1.0f / c;
        uint64_t bad_guarded = 0, bad_any = 0;
#pragma omp parallel for reduction(+ : bad_guarded, bad_any) schedule(static)
        for (int64_t i = 0; i < (1LL << 32); i++) {
            const float x = asf((uint32_t)i), t = x / c;
            const float q = x * rc, r = fmaf(-q, c, x), q2 = fmaf(r, rc, q);
            if (asu(q2) != asu(t) && !(t != t && q2 != q2)) {
                bad_any++;
                const float ax = fabsf(x);
                if (ax >= 0x1.0p-120f && ax <= 0x1.fffffep+127f) bad_guarded++;
            }
        }
        printf("c = %-8g rc = %a : mismatches on the guarded domain %llu, anywhere %llu\n", c, rc,
               (unsigned long long)bad_guarded, (unsigned long long)bad_any);
        fail |= bad_guarded != 0;
    }
    return fail;
}
