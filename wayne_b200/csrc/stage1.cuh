// stage1.cuh -- trace / dispersion / sensitivity / expected counts (fp64).
//
// Everything here is tiny next to the photon stage, so the kernels are written
// for EXACTNESS: each expression keeps the operation order of the numpy
// expression it replaces and the library is built with -fmad=false, so no
// a*b+c is fused; the parity tests then see differences of at most a few ulp
// (gate: 1e-6 relative).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace wb {

// np.poly1d / np.polyval Horner form: y = y*x + p[i], starting from 0.
__device__ __forceinline__ double polyval4(const double *p, double x)
{
    double y = p[0]; // 0*x + p0
    y = y * x + p[1];
    y = y * x + p[2];
    y = y * x + p[3];
    return y;
}

// np.interp (end-clamped linear interpolation, numpy/_core/src/multiarray/
// compiled_base.c arr_interp): binary search for xp[j] <= x < xp[j+1].
__device__ inline double interp1(double x, const double *xp, const double *fp, int n)
{
    if (x > xp[n - 1])
        return fp[n - 1];
    if (x < xp[0])
        return fp[0];
    int lo = 0, hi = n; // invariant xp[lo] <= x
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= xp[mid])
            lo = mid;
        else
            hi = mid;
    }
    const int j = lo;
    if (j == n - 1 || xp[j] == x)
        return fp[j];
    const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double r = slope * (x - xp[j]) + fp[j];
    if (isnan(r)) {
        r = slope * (x - xp[j + 1]) + fp[j + 1];
        if (isnan(r) && fp[j] == fp[j + 1])
            r = fp[j];
    }
    return r;
}

// grism.py:111-118 (+ tools.py:106-128): one thread per wavelength bin.
struct Poly12 {
    double c[12];
};

__global__ void k_bin_tables(int W, const double *__restrict__ wl, const Poly12 poly, int ns,
                             const double *__restrict__ sens_wl,
                             const double *__restrict__ sens_val, double *ratio, double *sigl,
                             double *sigh, double *sens, double *dwl)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W)
        return;
    const double x = wl[w];
    ratio[w] = polyval4(poly.c + 0, x);
    sigl[w] = polyval4(poly.c + 4, x);
    sigh[w] = polyval4(poly.c + 8, x);
    sens[w] = interp1(x, sens_wl, sens_val, ns);
    // bin_centers_to_widths: half-gap to the previous centre (first bin reuses
    // the second's) plus half-gap to the next centre (last bin reuses its own).
    auto half_gap = [&](int i) {
        const int k = (i == 0) ? 1 : i;
        return (wl[k] - wl[k - 1]) / 2.;
    };
    const double g0 = half_gap(w);
    const double g1 = (w == W - 1) ? half_gap(W - 1) : half_gap(w + 1);
    dwl[w] = g0 + g1;
}

// Per chunk of `chunk_bins` bins (one CTA of the native thrower): the first and the last bin that
// can hold electrons at all (flux * sensitivity * width > 0), or lo = -1 when the wavelength grid is
// not monotonic inside the chunk.  Bin positions are linear in the wavelength (wl_to_x, wl_to_y),
// so on a monotonic grid the footprint of a chunk in ANY sub-sample is spanned by those two bins:
// the thrower places its tile from them instead of scanning the chunk's counts and positions
// (a 2048-bin pass and two barriers at the head of every CTA).  span[2c] = lo, span[2c+1] = hi.
__global__ void k_chunk_spans(int W, int chunk_bins, const double *__restrict__ wl,
                              const double *__restrict__ flux, const double *__restrict__ sens,
                              const double *__restrict__ dwl, int *span)
{
    __shared__ int s_lo, s_hi, s_bad;
    const int w0 = blockIdx.x * chunk_bins, w1 = min(W, w0 + chunk_bins);
    if (threadIdx.x == 0) {
        s_lo = 0x7fffffff;
        s_hi = -1;
        s_bad = 0;
    }
    __syncthreads();
    int lo = 0x7fffffff, hi = -1, bad = 0;
    for (int w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
        const double e = flux[w] * sens[w] * dwl[w];
        if (e > 0.0) { // (false for NaN)
            lo = min(lo, w);
            hi = max(hi, w);
        }
        const double a = wl[w];
        if (!(fabs(a) < 1e300) || (w + 1 < w1 && !(wl[w + 1] >= a) && !(wl[w + 1] <= a)))
            bad = 1; // NaN / inf
    }
    // monotonic either way: every step has the sign of the chunk's end-to-end difference
    if (w1 - w0 > 1) {
        const double d = wl[w1 - 1] - wl[w0];
        for (int w = w0 + threadIdx.x; w + 1 < w1; w += blockDim.x) {
            const double step = wl[w + 1] - wl[w];
            if ((d >= 0 && step < 0) || (d <= 0 && step > 0))
                bad = 1;
        }
    }
    if (lo <= hi) {
        atomicMin(&s_lo, lo);
        atomicMax(&s_hi, hi);
    }
    if (bad)
        atomicOr(&s_bad, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        span[2 * blockIdx.x] = s_bad ? -1 : (s_lo <= s_hi ? s_lo : w0);
        span[2 * blockIdx.x + 1] = s_lo <= s_hi ? s_hi : w0;
    }
}

struct TraceCoef {
    double x_ref, y_ref, m_t, c_t, m_w, c_w, m_wl, c_wl;
};

// grism.py:779-803 then :591-602 for one reference position.
__device__ __forceinline__ TraceCoef make_trace(double x, double y, const double *a,
                                                const double *b)
{
    TraceCoef t;
    t.x_ref = x;
    t.y_ref = y;
    t.m_t = a[3] + a[4] * x + a[5] * y + a[6] * (x * x) + a[7] * x * y + a[8] * (y * y);
    t.c_t = a[0] + a[1] * x + a[2] * y;
    t.m_w = b[3] + b[4] * x + b[5] * y + b[6] * (x * x) + b[7] * x * y + b[8] * (y * y);
    t.c_w = (b[0] + b[1] * x) + b[2] * y;
    // two probe points at x_ref+10 / +20 on the trace -> linear lambda(x)
    const double X0 = x + 10, X1 = x + 20;
    const double Y0 = t.m_t * (X0 - x) + t.c_t + y;
    const double Y1 = t.m_t * (X1 - x) + t.c_t + y;
    const double d0 = sqrt((Y0 - y) * (Y0 - y) + (X0 - x) * (X0 - x));
    const double d1 = sqrt((Y1 - y) * (Y1 - y) + (X1 - x) * (X1 - x));
    const double l0 = (t.m_w * d0 + t.c_w) * 1e-4; // angstrom -> micron
    const double l1 = (t.m_w * d1 + t.c_w) * 1e-4;
    t.m_wl = (l1 - l0) / (X1 - X0);
    t.c_wl = l0 - t.m_wl * X0;
    return t;
}

// _SpectrumTrace.wl_to_x / wl_to_y (grism.py:635-669) minus the sub-array
// shift (exposure_generator.py:630-632).
__device__ __forceinline__ void trace_xy(const TraceCoef &t, double wl, double sub_scale,
                                         double &x, double &y)
{
    const double xp = (wl - t.c_wl) / t.m_wl;
    const double yp = t.m_t * (xp - t.x_ref) + t.c_t + t.y_ref;
    x = xp - sub_scale;
    y = yp - sub_scale;
}

__device__ __forceinline__ TraceCoef load_trace(const double *tr)
{
    TraceCoef t;
    t.x_ref = tr[0];
    t.y_ref = tr[1];
    t.m_t = tr[2];
    t.c_t = tr[3];
    t.m_w = tr[4];
    t.c_w = tr[5];
    t.m_wl = tr[6];
    t.c_wl = tr[7];
    return t;
}

struct Coef18 {
    double a[9], b[9];
};

__global__ void k_trace_table(int N, const double *__restrict__ xr, const double *__restrict__ yr,
                              Coef18 c, double *trace)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N)
        return;
    const TraceCoef t = make_trace(xr[s], yr[s], c.a, c.b);
    double *o = trace + (size_t)s * WB200_TRACE_STRIDE;
    o[0] = t.x_ref;
    o[1] = t.y_ref;
    o[2] = t.m_t;
    o[3] = t.c_t;
    o[4] = t.m_w;
    o[5] = t.c_w;
    o[6] = t.m_wl;
    o[7] = t.c_wl;
}

__global__ void k_trace_positions(int N, int W, const double *__restrict__ trace,
                                  const double *__restrict__ wl, double sub_scale, double *xpos,
                                  double *ypos)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (w >= W || s >= N)
        return;
    const TraceCoef t = load_trace(trace + (size_t)s * WB200_TRACE_STRIDE);
    double x, y;
    trace_xy(t, wl[w], sub_scale, x, y);
    xpos[(size_t)s * W + w] = x;
    ypos[(size_t)s * W + w] = y;
}

// exposure_generator.py:344-348 and :602-628.  grid = (ceil(W/128), ceil(N/COUNTS_SPT)).
// A thread owns one bin for COUNTS_SPT consecutive sub-samples: the bin's flux,
// sensitivity and width are loaded once, the next sub-sample's planet-signal load
// is in flight during the current draw, and the per-sub-sample photon totals leave
// through one warp reduction + atomic each (no CTA barrier: a warp that is held up in
// the Poisson rejection loop does not stall the others).
constexpr int COUNTS_SPT = 4;
constexpr int COUNTS_THREADS = 128;
#ifndef COUNTS_MIN_BLOCKS
#define COUNTS_MIN_BLOCKS 8
#endif

__global__ void __launch_bounds__(COUNTS_THREADS, COUNTS_MIN_BLOCKS)
k_counts(int N, int W, const double *__restrict__ flux, const double *__restrict__ depth,
         long long depth_ld, const double *__restrict__ cheb_coef, int cheb_order,
         const double *__restrict__ cheb_x, const double *__restrict__ sens,
         const double *__restrict__ dwl, const double *__restrict__ dur_ms, double scale, int mode,
         uint32_t k0, uint32_t k1, double *expected, int *counts, unsigned long long *totals,
         const double *__restrict__ sep_row)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int s0 = blockIdx.y * COUNTS_SPT;
    const bool live = w < W;
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
        if (use_depth && s0 < N)
            dnext = depth[(size_t)s0 * depth_ld + w];
        if (use_sep)
            sep_col = depth[w];
    }
#pragma unroll 1
    for (int i = 0; i < COUNTS_SPT; ++i) {
        const int s = s0 + i;
        if (s >= N)
            break; // CTA-uniform
        long long c = 0;
        const double dcur = dnext;
        if (use_depth && i + 1 < COUNTS_SPT && s + 1 < N) // next sub-sample's planet signal in flight
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
            double e = f * sn;       // ph / s / angstrom
            e = e * dw;              // (ph/s/A) * micron
            e = e * 1e4;             // micron -> angstrom : ph / s
            e = e * dur_ms[s];       // ph/s * ms
            e = e * 1e-3;            // ms -> s : photons
            e = e * scale;           // visit trend (exposure_generator.py:620-621)
            if (expected)
                expected[(size_t)s * W + w] = e;
            if (mode == WB200_COUNT_ROUND) {
                c = (long long)rint(e);
            } else if (mode == WB200_COUNT_POISSON) {
                PhiloxStream g(k0, k1, (uint32_t)w, (uint32_t)s, WB_STREAM_COUNTS);
                c = poisson_draw_fast(g, e);
            } else if (counts) {
                c = counts[(size_t)s * W + w];
            }
            if (c < 0)
                c = 0;
            if (c > 2147483647LL)
                c = 2147483647LL; // the reference's counters are 32-bit (pyparallel_menu.c:12)
            if (mode != WB200_COUNT_NONE && counts)
                counts[(size_t)s * W + w] = (int)c;
        }
        const unsigned long long v = warp_sum_u64((unsigned long long)c);
        if (lane_id() == 0 && v)
            atomicAdd(&totals[s], v);
    }
}

// ---------------------------------------------------------------------------
// Transit light curves (wayne_b200/lightcurve.py::planet_signal on the device).
// One CTA per sub-sample, one thread per Chebyshev node of the radius ratio:
// 1 - flux by Gauss-Legendre quadrature over the occulted annuli, then the
// discrete Chebyshev transform of the CTA's `order` node values.
// ---------------------------------------------------------------------------
struct Claret4 {
    double c[4];
};

__device__ __forceinline__ double claret_intensity(double r, const Claret4 &k)
{
    const double mu = sqrt(fmax(1.0 - r * r, 0.0));
    const double sq = sqrt(mu);
    return 1.0 - k.c[0] * (1.0 - sq) - k.c[1] * (1.0 - mu) - k.c[2] * (1.0 - mu * sq) -
           k.c[3] * (1.0 - mu * mu);
}

__global__ void __launch_bounds__(32)
k_transit_cheb(int N, int order, const double *__restrict__ zs, double pmin, double pmax, Claret4 ld,
               double total, int n_gl, const double *__restrict__ glx, const double *__restrict__ glw,
               double *coef)
{
    __shared__ double f[32];
    const int s = blockIdx.x, k = threadIdx.x;
    const double PI = 3.141592653589793;
    if (k < order) {
        const double xk = cos(PI * (k + 0.5) / order);
        const double p = 0.5 * (pmax - pmin) * xk + 0.5 * (pmax + pmin);
        const double z = zs[s];
        double blocked = 0.0;
        if (z < 1.0 + p && p > 0.0) {
            const double lo = fmin(fmax(z - p, 0.0), 1.0), hi = fmin(fmax(z + p, 0.0), 1.0);
            const double span = hi - lo;
            for (int i = 0; i < n_gl; ++i) {
                const double sv = 0.5 * (glx[i] + 1.0), w = 0.5 * glw[i];
                const double r = hi - span * sv * sv; // limb-side square-root behaviour at sv -> 0
                const double jac = 2.0 * span * sv;
                double ang;
                if (z == 0.0)
                    ang = (r <= p) ? PI : 0.0;
                else if (r <= p - z)
                    ang = PI;
                else
                    ang = acos(fmin(fmax((r * r + z * z - p * p) / (2.0 * r * z), -1.0), 1.0));
                blocked += w * jac * claret_intensity(r, ld) * 2.0 * r * ang;
            }
        }
        f[k] = blocked / total; // = 1 - flux
    }
    __syncthreads();
    if (k < order) {
        double c = 0.0;
        for (int j = 0; j < order; ++j)
            c += f[j] * cos(PI * k * (j + 0.5) / order);
        c *= 2.0 / order;
        if (k == 0)
            c *= 0.5;
        coef[(size_t)s * order + k] = c;
    }
}

// Exclusive prefix of counts along bins; one CTA (256 threads) per sub-sample.
__global__ void __launch_bounds__(256)
k_count_offsets(int W, const int *__restrict__ counts, int *offsets)
{
    const int s = blockIdx.x;
    const int *c = counts + (size_t)s * W;
    int *o = offsets + (size_t)s * W;
    __shared__ int wsum[8];
    __shared__ int carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int base = 0; base < W; base += 256) {
        const int w = base + threadIdx.x;
        const int v = (w < W) ? c[w] : 0;
        const int inc = warp_incl_scan(v);
        if (lane_id() == 31)
            wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        int pre = carry;
        for (int i = 0; i < (int)(threadIdx.x >> 5); ++i)
            pre += wsum[i];
        if (w < W)
            o[w] = pre + inc - v;
        __syncthreads();
        if (threadIdx.x == 255)
            carry = pre + inc;
        __syncthreads();
    }
}

} // namespace wb
