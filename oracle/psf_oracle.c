/*
 * psf_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle; never shipped,
 * never on the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it).
 *
 * A plain-C, single-threaded restatement of the algorithm of the reference's
 * native electron thrower  PSF()  (reference: wayne/pyparallel_menu.c:10-113),
 * split into its three logical steps so that each can be pinned separately:
 *
 *   wo_rand_r            glibc rand_r restated (reference calls rand_r at
 *                        pyparallel_menu.c:57-58; glibc stdlib/rand_r.c is the
 *                        third-party arithmetic, container glibc 2.39)
 *   wo_fill_normals      the Box-Muller table  A[2*ssum]  with the reference's
 *                        per-OpenMP-thread chunking and seeding
 *                        (pyparallel_menu.c:40-64), emulated serially
 *   wo_bin_electrons     the serial scatter / histogram loop
 *                        (pyparallel_menu.c:87-108)
 *   wo_psf               the three glued together == PSF() for a given
 *                        (test, threads)
 *
 * Pinning: tests/test_oracle_psf.py checks wo_psf bit-for-bit against the
 * UNMODIFIED reference compiled from /root/reference (oracle/_ref/, built by
 * oracle/Makefile) for several (test, threads) pairs, and wo_rand_r against
 * the C library's rand_r; committed golden histograms in tests/golden/ carry
 * that pin to machines where /root/reference does not exist.
 *
 * Build: see oracle/Makefile (gcc -O2 -fPIC -shared, -ffp-contract=off so the
 * a*b+c expressions are never fused -- the reference is built without FMA on
 * x86-64, setup.py:72-73).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WO_PI 3.14159265358979323846      /* pyparallel_menu.c:8 */
#define WO_RAND_MAX 2147483647            /* glibc RAND_MAX */

/* glibc rand_r: three steps of  s <- s*1103515245 + 12345 (mod 2^32), taking
 * 11, 10 and 10 bits from bit 16 upwards of the successive states. */
int wo_rand_r(uint32_t *state)
{
    uint32_t s = *state;
    uint32_t out;
    s = s * 1103515245u + 12345u;
    out = (s >> 16) & 0x7ffu;            /* (s/65536) % 2048 */
    s = s * 1103515245u + 12345u;
    out = (out << 10) ^ ((s >> 16) & 0x3ffu);
    s = s * 1103515245u + 12345u;
    out = (out << 10) ^ ((s >> 16) & 0x3ffu);
    *state = s;
    return (int)out;
}

/* First electron index handled by OpenMP thread t of T (int arithmetic as in
 * pyparallel_menu.c:47-49; the last thread runs to ssum). */
static int chunk_begin(int t, int T, int ssum)
{
    return (t >= T) ? ssum : (int)(t * ssum / T);
}

/* A[i] = x-normal of electron i, A[i+ssum] = y-normal (pyparallel_menu.c:55-62).
 * Thread t seeds with 25234 + 17*t + test (:52) and walks its chunk in order. */
void wo_fill_normals(double *A, int ssum, int test, int threads)
{
    for (int t = 0; t < threads; ++t) {
        int lo = chunk_begin(t, threads, ssum);
        int hi = (t == threads - 1) ? ssum : chunk_begin(t + 1, threads, ssum);
        uint32_t state = (uint32_t)(25234 + 17 * t + test);
        for (int i = lo; i < hi; ++i) {
            double theta = 2. * WO_PI * wo_rand_r(&state) / ((double)WO_RAND_MAX);
            double R = sqrt(-2. * log(wo_rand_r(&state) / ((double)WO_RAND_MAX)));
            A[i] = R * cos(theta);
            A[i + ssum] = R * sin(theta);
        }
    }
}

/* Truncating double->int conversion with the x86 out-of-range behaviour the
 * reference gets from cvttsd2si (INT_MIN for NaN / overflow), written out so
 * the oracle does not depend on undefined behaviour. */
static int trunc_to_int(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return INT32_MIN;
    return (int)v;
}

/* Scatter loop (pyparallel_menu.c:87-108).  frame[nr*nc] must be zeroed by the
 * caller (:68-83 does that in the reference).  Returns electrons consumed. */
long wo_bin_electrons(const int *counts, int size, const double *x_pos,
                      const double *y_pos, const double *ratio,
                      const double *sigl, const double *sigh, int nr, int nc,
                      const double *A, long ssum, int *frame)
{
    long e = 0;
    for (int b = 0; b < size; ++b) {
        int n_wide = trunc_to_int(counts[b] * ratio[b]);     /* :89 */
        for (int j = 0; j < counts[b]; ++j, ++e) {
            /* the first n_wide electrons of the bin take the wide Gaussian
             * (:90-98), the rest the narrow one (:99-107) */
            double s = (j < n_wide) ? sigh[b] : sigl[b];
            int xp = trunc_to_int(A[e] * s + x_pos[b]);
            int yp = trunc_to_int(A[e + ssum] * s + y_pos[b]);
            if (xp > 0 && xp < nr && yp > 0 && yp < nc)      /* :93, :102 */
                frame[(long)yp * nc + xp] += 1;
        }
    }
    return e;
}

/* == PSF(): returns 0 on success; frame is caller-owned int[nr*nc]. */
int wo_psf(const int *counts, int size, const double *x_pos,
           const double *y_pos, const double *ratio, const double *sigl,
           const double *sigh, int nr, int nc, int test, int threads,
           int *frame)
{
    long ssum = 0;
    for (int b = 0; b < size; ++b)
        ssum += counts[b];                                   /* :19-34 */
    double *A = (double *)malloc((size_t)(2 * ssum + 1) * sizeof(double));
    if (!A)
        return -1;
    wo_fill_normals(A, (int)ssum, test, threads);
    memset(frame, 0, (size_t)nr * nc * sizeof(int));
    wo_bin_electrons(counts, size, x_pos, y_pos, ratio, sigl, sigh, nr, nc, A,
                     ssum, frame);
    free(A);
    return 0;
}
