/* CPU restatement of the arithmetic of csrc/solver_kernels.cu: div_tiny (the exact division of numerators below 2^-100,
 * denormals included, that the sweeps' rare path uses instead of the compiler's IEEE slow path) checked against the
 * host's IEEE single-precision division.  TEST INFRASTRUCTURE (tests/test_div_tiny_cpu.py); the GPU self-test
 * rtdd_selftest_division mode 4 checks the device code itself, with div_fast's MUFU-based quotient.
 * Here Q = RN(S / b) is taken from the host's correctly rounded division, which is what div_fast's sequence delivers
 * for operands in its range.  Build: gcc -O2 -ffp-contract=off -fno-fast-math. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint64_t state = 88172645463325252ULL;
static uint32_t rnd(void) { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return (uint32_t)(state >> 16); }

static float div_tiny(float a, float b, int fix_ties)
{
    const float S = a * 0x1p100f;                 /* exact */
    const float Q = S / b;                        /* correctly rounded quotient of operands in div_fast's range */
    const float q = Q * 0x1p-100f;                /* second rounding: only inexact for denormal results */
    const float Qs = Q * 0x1p49f;                 /* in units of the smallest denormal (exact, or +-inf for large Q) */
    const float fl = floorf(Qs);
    const float side = fmaf(-b, Q, S);            /* exact sign of S - b * Q */
    const int fix = fix_ties && (Qs - fl) == 0.5f && side != 0.0f;
    const float n = (side > 0.0f) ? fl + 1.0f : fl;
    const float qf = copysignf(n * u2f(1u), Q);
    return fix ? qf : q;
}

int main(int argc, char **argv)
{
    const long count = argc > 1 ? atol(argv[1]) : 20000000L;
    const int fix_ties = argc > 2 ? atoi(argv[2]) : 1;
    long bad = 0;
    for (long it = 0; it < count; it++) {
        const uint32_t r0 = rnd(), r1 = rnd(), r2 = rnd();
        uint32_t mag;
        switch (r2 & 3u) {                                        /* numerators in (0, 2^-100) */
        case 0: mag = 1u + r0 % ((27u << 23) - 1u); break;
        case 1: mag = 1u + r0 % 4096u; break;                      /* few-bit denormals: midpoint ties are frequent */
        case 2: mag = 1u + r0 % (1u << 23); break;                 /* denormals */
        default: mag = (1u << 23) + r0 % (26u << 23); break;       /* small normals */
        }
        const float a = u2f(mag | ((r2 >> 31) << 31));
        uint32_t mb = r1 & 0x7FFFFFu;
        if (((r2 >> 16) & 7u) == 0u) mb &= 0x700000u;              /* simple mantissas: exact ties and near-ties */
        if (((r2 >> 16) & 7u) == 1u) mb &= 0x7FF000u;
        const float b = u2f(((127u - 100u + (r2 >> 8) % 103u) << 23) | mb);   /* denominators in [2^-100, 8) */
        if (f2u(a / b) != f2u(div_tiny(a, b, fix_ties))) bad++;
    }
    printf("%ld %ld\n", count, bad);
    return 0;
}
