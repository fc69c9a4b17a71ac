// counts_native.cuh -- Poisson photon counts of the (sub-sample, bin) cells for the
// native mode (exposure_generator.py:602-628 with np.random.poisson's distribution),
// written for throughput.  k_counts (stage1.cuh) draws every cell with the PTRS
// rejection sampler: ~540 instructions per cell at the pace of the slowest lane of
// the warp.  Here a thread owns ONE bin for a block of consecutive sub-samples.  The
// cell means of one bin differ between neighbouring sub-samples only through the
// planet signal and the scan-speed variation, so the thread tabulates the Poisson
// CDF once -- a window of CW_T partial sums around the mode, in shared memory, for a
// mean a few per cent below the bin's (scan-speed variations modulate the cell means
// by that much from one sub-sample to the next) -- and then every draw is
//     one binary search of the window            (Poisson of the tabulated mean lam_t)
//   + one small Poisson draw of the remainder    (mean lam - lam_t, a few per cent of lam),
// exact because independent Poissons add.  Draws that fall outside the window are
// finished by the same recurrence (upwards from the window's last term, downwards
// from its first), so the sampler is inversion of ONE uniform throughout.  Means
// above CW_LAM_MAX (where a window of CW_T terms would be left too often) take the
// PTRS sampler of k_counts with that kernel's counters.
//
// Uniforms: Philox4x32-10 call (2^31 + 0, bin, sub-sample >> 1, WB_STREAM_COUNTS);
// words (x, y) serve the even sub-sample of the pair, (z, w) the odd one.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace wb {

constexpr int CW_T = 64;             // window entries per thread
constexpr int CW_THREADS = 128;      // 32 KB of shared memory per CTA, 6 CTAs per SM
#ifndef CW_BLOCK_N
#define CW_BLOCK_N 32
#endif
#ifndef CW_MIN_BLOCKS
#define CW_MIN_BLOCKS 6
#endif
constexpr int CW_BLOCK = CW_BLOCK_N; // sub-samples per thread
constexpr float CW_LAM_MAX = 100.0f; // window: [lam - 3.7 sigma, lam + 2.7 sigma] at this mean
constexpr float CW_MARGIN = 0.045f;  // the window is tabulated this far (relative) below the current mean
constexpr float CW_REUSE = 0.25f;    // ... and serves means up to lam_t (1 + 2 CW_MARGIN) + CW_REUSE

struct CountWindow {
    float lam;   // tabulated mean (-1: none)
    float below; // P(X < k0)
    float top;   // P(X <= k0 + CW_T - 1)
    float p0;    // P(X = k0)
    float ptop;  // P(X = k0 + CW_T - 1)
    int k0;
};

// tbl: this thread's column, consecutive entries CW_THREADS floats apart (bank = lane)
__device__ __forceinline__ void count_window_build(float *tbl, CountWindow &t, float lam)
{
    float sl;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sl) : "f"(lam));
    const int k0 = max(0, (int)(lam - 3.7f * sl) - 1);
    float p, below = 0.0f;
    if (k0 == 0) {
        p = expf(-lam); // lam < ~16 here
    } else {
        // P(X = k0) from its logarithm in fp64 (e^-lam itself underflows fp32 beyond lam = 87)
        const double lp = -(double)lam + (double)k0 * log((double)lam) - log_factorial((double)k0);
        p = (float)exp(lp);
        // lower tail by the downward recurrence p(k-1) = p(k) k / lam
        float q = p;
        const float il = rcp_ftz(lam);
        for (int k = k0; k > 0 && q > 1e-9f * below + 1e-30f; --k) {
            q *= (float)k * il;
            below += q;
        }
    }
    float s = below + p;
    tbl[0] = s;
    float kf = (float)k0, pk = p;
#pragma unroll 8
    for (int j = 1; j < CW_T; ++j) {
        kf += 1.0f;
        pk *= lam * rcp_ftz(kf);
        s += pk;
        tbl[j * CW_THREADS] = s;
    }
    t.lam = lam;
    t.below = below;
    t.top = s;
    t.p0 = p;
    t.ptop = pk;
    t.k0 = k0;
}

// Poisson(lam) for the SMALL remainder mean (a few per cent of the cell's mean: ~0..6) by inversion of
// one uniform, like poisson_inversion_u, but with the first eight terms of the chop-down search
// unrolled and predicated: a warp's walk runs at the pace of its slowest lane (~lam + 3 sigma
// dependent trips of ~10 instructions, 23 % of this kernel's instructions on the configs[3]
// shape); straight-line, the eight steps are five instructions each with the reciprocals as
// immediates, and only a lane beyond k = 8 (2e-3 of the draws at lam = 2.7) enters the loop.
__device__ __forceinline__ int poisson_small_u(float u, float lam)
{
    u = fminf(u, 0.99999994f);
    float p, s;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(-1.4426950408889634f * lam)); // e^-lam (lam < ~90)
    s = p;
    int k = 0;
#pragma unroll
    for (int j = 1; j <= 8; ++j) {
        k += (u > s) ? 1 : 0; // s = P(X <= j - 1)
        p *= lam * (1.0f / (float)j);
        s += p;
    }
    if (u > s) { // beyond k = 8: the rest of the search as in poisson_inversion_u
        k = 8;
        do {
            ++k;
            p *= lam * rcp_ftz((float)k);
            s += p;
        } while (u > s && !(p < 1e-10f && (float)k > lam));
    }
    return k;
}

__device__ __forceinline__ int count_window_draw(const float *tbl, const CountWindow &t, float u)
{
    u = fminf(u, 0.99999994f);
    if (u > t.top) { // above the window: continue the recurrence
        int k = t.k0 + CW_T - 1;
        float p = t.ptop, s = t.top;
        while (u > s) {
            ++k;
            p *= t.lam * rcp_ftz((float)k);
            s += p;
            if (p < 1e-10f && (float)k > t.lam)
                break;
        }
        return k;
    }
    if (!(u > t.below)) { // below the window: walk down from k0 - 1
        int k = t.k0 - 1;
        float c = t.below;
        float q = t.p0 * (float)t.k0 * rcp_ftz(t.lam); // P(X = k0 - 1)
        while (k > 0 && u <= c - q) {
            c -= q;
            q *= (float)k * rcp_ftz(t.lam);
            --k;
        }
        return k;
    }
    int j = 0; // entries < u; tbl[CW_T-1] >= u
#pragma unroll
    for (int step = CW_T / 2; step >= 1; step >>= 1) {
        WB_DEV_ASSERT(j + step - 1 >= 0 && j + step - 1 < CW_T);
        if (tbl[(j + step - 1) * CW_THREADS] < u)
            j += step;
    }
    WB_DEV_ASSERT(j < CW_T && tbl[j * CW_THREADS] >= u);
    return t.k0 + j;
}

// grid = (ceil(W / CW_THREADS), ceil(N / CW_BLOCK)); POISSON mode only
__global__ void __launch_bounds__(CW_THREADS, CW_MIN_BLOCKS)
k_counts_window(int N, int W, const double *__restrict__ flux, const double *__restrict__ depth,
                long long depth_ld, const double *__restrict__ cheb_coef, int cheb_order,
                const double *__restrict__ cheb_x, const double *__restrict__ sens,
                const double *__restrict__ dwl, const double *__restrict__ dur_ms, double scale,
                uint32_t k0, uint32_t k1, double *expected, int *counts, unsigned long long *totals,
                const double *__restrict__ sep_row)
{
    extern __shared__ float s_cdf[]; // [CW_T][CW_THREADS]
    const int w = blockIdx.x * CW_THREADS + threadIdx.x;
    const int s0 = blockIdx.y * CW_BLOCK;
    const int s1 = min(N, s0 + CW_BLOCK);
    const bool live = w < W;
    float *tbl = s_cdf + threadIdx.x;
    double f0 = 0.0, sn = 0.0, dw = 0.0, x = 0.0;
    // separable planet signal depth[s][w] = sep_row[s] * depth[w] (depth = the per-bin factor)
    const bool use_sep = live && !cheb_coef && depth && sep_row;
    const bool use_depth = live && !cheb_coef && depth && !sep_row;
    double dnext = 0.0, sep_col = 0.0;
    if (live) {
        f0 = flux[w];
        sn = sens[w];
        dw = dwl[w];
        if (cheb_coef)
            x = cheb_x[w];
        if (use_depth)
            dnext = depth[(size_t)s0 * depth_ld + w];
        if (use_sep)
            sep_col = depth[w];
    }
    CountWindow win;
    win.lam = -1.0f;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int s = s0; s < s1; ++s) {
        long long c = 0;
        const double dcur = dnext;
        if (use_depth && s + 1 < s1) // next sub-sample's planet signal in flight
            dnext = depth[(size_t)(s + 1) * depth_ld + w];
        if (live) {
            double f = f0;
            if (cheb_coef) {
                // Clenshaw recurrence in numpy.polynomial.chebyshev.chebval's order
                const double *cf = cheb_coef + (size_t)s * cheb_order;
                double d;
                if (cheb_order == 1) {
                    d = cf[0];
                } else {
                    const double x2 = 2 * x;
                    double c0 = cf[cheb_order - 2], c1 = cf[cheb_order - 1];
                    for (int j = 3; j <= cheb_order; ++j) {
                        const double tmp = c0;
                        c0 = cf[cheb_order - j] - c1;
                        c1 = tmp + c1 * x2;
                    }
                    d = c0 + c1 * x;
                }
                f = f * (1. - d);
            } else if (use_depth)
                f = f * (1. - dcur);
            else if (use_sep)
                f = f * (1. - sep_col * sep_row[s]);
            // the unit algebra of k_counts (exposure_generator.py:602-623), same order
            double e = f * sn;
            e = e * dw;
            e = e * 1e4;
            e = e * dur_ms[s];
            e = e * 1e-3;
            e = e * scale;
            if (expected)
                expected[(size_t)s * W + w] = e;
            if ((s & 1) == 0) // CW_BLOCK is even, so every block starts on an even sub-sample
                rnd = philox4x32_10(make_uint4(0x80000000u, (uint32_t)w, (uint32_t)(s >> 1), WB_STREAM_COUNTS),
                                    k0, k1);
            const uint32_t wu = (s & 1) ? rnd.z : rnd.x, wu2 = (s & 1) ? rnd.w : rnd.y;
            if (!(e > 0.0)) {
                c = 0;
            } else if (e < (double)CW_LAM_MAX) {
                const float lam = (float)e;
                float rest = lam - win.lam;
                if (!(rest >= 0.0f && rest <= fmaf(2.0f * CW_MARGIN, win.lam, CW_REUSE))) {
                    // tabulate a little below this mean: scan-speed variations move the
                    // neighbours' means by a few per cent either way, a transit by parts in 1e5
                    count_window_build(tbl, win, lam * (1.0f - CW_MARGIN));
                    rest = lam - win.lam;
                }
                c = count_window_draw(tbl, win, u01f(wu));
                if (rest > 0.0f)
                    c += poisson_small_u(u01f(wu2), rest);
            } else {
                PhiloxStream g(k0, k1, (uint32_t)w, (uint32_t)s, WB_STREAM_COUNTS);
                c = poisson_draw_fast(g, e);
                if (c > 2147483647LL)
                    c = 2147483647LL; // the reference's counters are 32-bit (pyparallel_menu.c:12)
            }
            WB_DEV_ASSERT(c >= 0 && s < N && w < W);
            if (counts)
                counts[(size_t)s * W + w] = (int)c;
        }
        // photons of this sub-sample: one hardware warp reduction when every lane's count fits
        // 26 bits (always, short of a saturated detector), the shuffle tree otherwise
        unsigned long long v;
        if (__all_sync(0xffffffffu, c < (1LL << 26)))
            v = __reduce_add_sync(0xffffffffu, (unsigned)c);
        else
            v = warp_sum_u64((unsigned long long)c);
        if (lane_id() == 0 && v)
            atomicAdd(&totals[s], v);
    }
}

} // namespace wb
