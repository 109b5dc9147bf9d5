/* noisecheck.c -- the glottal-noise sample of flowgen_shimmer.c:387,394 in two FP64 operations.
 *
 * reference:  w = (short)ceil((1.0*random()/RAND_MAX)*NDW - NDW/2.0)           (divide, multiply, subtract, ceil)
 * kernel:     t = fma((double)min(r, M-1), fl(1/M), -0.5);   w = ceil(t * (double)NDW)
 *
 * Why they agree: the exact value is NDW*(2r-M)/(2M), M = 2^31-1 prime.  It is an integer only for r = 0 or r = M
 * (M divides neither NDW < M nor 2r-M otherwise): r = 0 is exact in both forms; at r = M the kernel evaluates r = M-1
 * instead, NDW/2 - NDW/(2M), which has the ceiling of NDW/2 because NDW/(2M) < 1/2.  Everywhere else it is at least
 * 1/(2M) ~ 2^-32 away from the nearest integer, while both floating-point forms err by less than 2^-34 for
 * NDW < 2^19 -- so neither can cross an integer and both ceilings equal the exact ceiling.
 *
 * This program checks the claim: every NDW up to NDW_FULL against r at and around every breakpoint of the
 * ceiling, plus random (r, NDW) pairs up to 2^19.  Build: gcc -O2 -ffp-contract=off noisecheck.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define M 2147483647.0
static const double INV_M = 1.0 / 2147483647.0;

static int w_ref(int32_t r, int32_t ndw)
{
    volatile double u = (1.0 * r) / M;
    volatile double a = u * ndw;
    volatile double b = ndw / 2.0;
    volatile double c = a - b;
    return (int)(short)ceil(c);
}
static int w_fast(int32_t r, int32_t ndw)
{
    double t = fma((double)(r < 2147483646 ? r : 2147483646), INV_M, -0.5);
    volatile double p = t * (double)ndw;
    return (int)(short)(int32_t)ceil(p);
}

int main(int argc, char **argv)
{
    const int ndw_full = argc > 1 ? atoi(argv[1]) : 2048;
    unsigned long long checked = 0, bad = 0;
    for (int32_t ndw = 0; ndw <= ndw_full; ndw++) {
        /* breakpoints: w changes where ndw*(2r-M)/(2M) crosses an integer k, r ~ (2k*M/ndw + M)/2 */
        for (int k = -(ndw / 2) - 1; k <= ndw / 2 + 1; k++) {
            const double rc = ndw ? (2.0 * k * M / ndw + M) / 2.0 : 0.0;
            for (int d = -2; d <= 2; d++) {
                double rr = floor(rc) + d;
                if (rr < 0 || rr > M) continue;
                const int32_t r = (int32_t)rr;
                checked++;
                if (w_ref(r, ndw) != w_fast(r, ndw)) { if (bad++ < 10) printf("MISMATCH r=%d ndw=%d ref=%d fast=%d\n", r, ndw, w_ref(r, ndw), w_fast(r, ndw)); }
            }
        }
        const int32_t edge[] = {0, 1, 2, 1073741823, 1073741824, 2147483645, 2147483646, 2147483647};
        for (unsigned e = 0; e < sizeof edge / sizeof *edge; e++) { checked++; if (w_ref(edge[e], ndw) != w_fast(edge[e], ndw)) { bad++; printf("MISMATCH edge r=%d ndw=%d\n", edge[e], ndw); } }
    }
    uint64_t s = 88172645463325252ull;
    for (long it = 0; it < 200000000L; it++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const int32_t r = (int32_t)(s & 0x7fffffff);
        const int32_t ndw = (int32_t)((s >> 31) & 0x7ffff);
        checked++;
        if (w_ref(r, ndw) != w_fast(r, ndw)) { if (bad++ < 10) printf("MISMATCH r=%d ndw=%d\n", r, ndw); }
    }
    printf("%llu checked, %llu mismatches\n", checked, bad);
    return bad != 0;
}
