// gather.cuh -- stage 3b: per-sub-sample flat field and ordered accumulation.
//
// Replaces, per sub-sample, G141.get_flat_field's `indices` branch
// (wayne/grism.py:359-385), `new_pixel_array *= flat_field`
// (wayne/exposure_generator.py:641-645) and `pixel_array += sample_frame`
// (:359).  The reference builds a full L x L flat per sub-sample; here a pixel
// thread walks the sub-samples of its read interval IN ORDER, reads its count
// from each sub-sample's HBM window, evaluates that sub-sample's flat at the
// pixel only where the count is non-zero, and adds count*flat to a register
// accumulator -- the same sequence of fp64 operations per pixel as the
// reference, so the interval planes are bit-reproducible (no float atomics).
//
// A CTA owns a 32 x 8 pixel rectangle of one read interval.  It first filters
// the interval's sub-samples down to those whose window overlaps the rectangle
// (one candidate per thread, ORDER-PRESERVING ballot compaction into shared
// memory) and then loops over the survivors only: CTAs outside the scanned band
// finish after the filter, CTAs inside it do no rejection work in the hot loop.
//
// EXACT = true   the reference's expression, operation for operation
//                (fp64 divide + sqrt per evaluation): compat / parity mode.
// EXACT = false  native mode: per-sample 1/sqrt(den) and 1/(wmax-wmin) hoisted,
//                Horner FMAs -- same value to ~1e-15 relative before the float32
//                rounding the reference applies to the flat value.
//
// Bound: HBM reads of the int32 windows (one coalesced row segment per
// (CTA row, sub-sample) overlap) + one RMW of the fp64 interval plane.
#pragma once
#include "common.cuh"
#include "stage1.cuh"

namespace wb {

constexpr int GX = 32, GY = 8;       // pixel rectangle of a CTA
constexpr int G_THREADS = GX * GY;   // candidates filtered per pass

struct GatherSample {
    int ox, oy, s, pad;
    double x_ref, y_ref, a_t_i, den, m_w, c_w; // den = a_t_i^2+1 (EXACT) or 1/sqrt(a_t_i^2+1)
};

// grism.py:359-385 for one pixel; Xd, Yd are flat-plane indices (python-style
// negative wrap is applied by the caller before the plane lookups).
__device__ __forceinline__ double flat_value_exact(const GatherSample &g, double Xd, double Yd,
                                                   double f0, double f1, double f2, double f3,
                                                   double wmin, double wmax)
{
    const double arr = g.y_ref - Yd + g.a_t_i * g.x_ref - g.a_t_i * Xd;
    const double d = sqrt((arr * arr) / g.den);
    const double wl = g.m_w * d + g.c_w;
    const double n = (wl - wmin) / (wmax - wmin);
    const double n2 = n * n;
    const double n3 = n2 * n;
    return f0 + (f1 * n) + (f2 * n2) + (f3 * n3);
}

__device__ __forceinline__ double flat_value_fast(const GatherSample &g, double Xd, double Yd,
                                                  double f0, double f1, double f2, double f3,
                                                  double wmin, double inv_range)
{
    const double arr = fma(g.a_t_i, g.x_ref - Xd, g.y_ref - Yd);
    const double wl = fma(g.m_w, fabs(arr) * g.den, g.c_w);
    const double n = (wl - wmin) * inv_range;
    return fma(n, fma(n, fma(n, f3, f2), f1), f0);
}

template <bool EXACT>
__global__ void __launch_bounds__(G_THREADS) k_gather(const wb200_gather_args a)
{
    __shared__ GatherSample sm[G_THREADS];
    __shared__ int s_wcount[G_THREADS / 32];
    const int r = blockIdx.z;
    const int c_l = blockIdx.x * GX + threadIdx.x; // light-sensitive coordinates
    const int r_l = blockIdx.y * GY + threadIdx.y;
    const int tid = threadIdx.y * GX + threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const bool live = (c_l < a.L) && (r_l < a.L);

    // sub-samples of read interval r present in this window batch
    const int first = (r == 0) ? 0 : a.d_read_end[r - 1] + 1;
    const int last = a.d_read_end[r];
    const int lo = max(first, a.sample0) - a.sample0;
    const int hi = min(last + 1, a.sample0 + a.n_samples) - a.sample0;
    if (lo >= hi)
        return;

    // CTA pixel rectangle (light-sensitive coords == the frame coords of PSF())
    const int tx0 = blockIdx.x * GX, ty0 = blockIdx.y * GY;

    const size_t pix = ((size_t)r * a.F + (r_l + a.border)) * a.F + (c_l + a.border);
    double acc = 0.0;
    bool loaded = false, touched = false;

    // flat-plane indices of this pixel (grism.py:361-363), python negative wrap
    // (flat_xs/flat_ys are meshgrids looked up at the wrapped index, so the
    // wrapped value is also what enters the wavelength formula)
    int Xi = c_l + a.flat_off, Yi = r_l + a.flat_off;
    if (Xi < 0)
        Xi += a.flat_n;
    if (Yi < 0)
        Yi += a.flat_n;
    const double Xd = (double)Xi, Yd = (double)Yi;
    const bool flat_ok = (unsigned)Xi < (unsigned)a.flat_n && (unsigned)Yi < (unsigned)a.flat_n;
    const size_t fidx = (size_t)Yi * a.flat_n + Xi;
    double f0 = 1, f1 = 0, f2 = 0, f3 = 0;
    bool have_f = false;
    const double inv_range = 1.0 / (a.flat_wmax - a.flat_wmin);

    for (int base = lo; base < hi; base += G_THREADS) {
        // ---- filter: one candidate per thread, order-preserving compaction ----
        const int s = base + tid;
        int ox = 0, oy = 0;
        bool sel = false;
        if (s < hi) {
            ox = a.d_win_ox[s];
            oy = a.d_win_oy[s];
            sel = !(ox >= tx0 + GX || ox + a.win_w <= tx0 || oy >= ty0 + GY || oy + a.win_h <= ty0);
        }
        const unsigned ballot = __ballot_sync(FULL, sel);
        __syncthreads(); // previous pass's sm[] fully consumed
        if (lane == 0)
            s_wcount[warp] = __popc(ballot);
        __syncthreads();
        int pos = __popc(ballot & ((1u << lane) - 1u)), n = 0;
#pragma unroll
        for (int w = 0; w < G_THREADS / 32; ++w) {
            const int c = s_wcount[w];
            if (w < warp)
                pos += c;
            n += c;
        }
        if (n == 0)
            continue; // uniform over the CTA
        if (sel) {
            GatherSample g;
            g.ox = ox;
            g.oy = oy;
            g.s = s;
            g.pad = 0;
            const double *t = a.d_trace + (size_t)s * WB200_TRACE_STRIDE;
            g.x_ref = t[0];
            g.y_ref = t[1];
            g.a_t_i = 1 / t[2];
            const double den = g.a_t_i * g.a_t_i + 1;
            g.den = EXACT ? den : 1.0 / sqrt(den);
            g.m_w = t[4];
            g.c_w = t[5];
            sm[pos] = g;
        }
        __syncthreads();
        if (!live)
            continue;
        // ---- hot loop: only windows that overlap this CTA's rectangle ----------
        for (int i = 0; i < n; ++i) {
            const GatherSample &g = sm[i];
            const int wx = c_l - g.ox, wy = r_l - g.oy;
            if ((unsigned)wx >= (unsigned)a.win_w || (unsigned)wy >= (unsigned)a.win_h)
                continue;
            const int h = a.d_win[((size_t)g.s * a.win_h + wy) * a.win_w + wx];
            if (h == 0)
                continue;
            if (!loaded) {
                acc = a.d_acc[pix];
                loaded = true;
            }
            double v = (double)h;
            if (a.add_flat) {
                if (!have_f && flat_ok) {
                    f0 = a.d_flat[0][fidx];
                    f1 = a.d_flat[1][fidx];
                    f2 = a.d_flat[2][fidx];
                    f3 = a.d_flat[3][fidx];
                    have_f = true;
                }
                double fv = EXACT ? flat_value_exact(g, Xd, Yd, f0, f1, f2, f3, a.flat_wmin, a.flat_wmax)
                                  : flat_value_fast(g, Xd, Yd, f0, f1, f2, f3, a.flat_wmin, inv_range);
                if (a.flat_f32)
                    fv = (double)__double2float_rn(fv); // stored in ones_like(flat_f0)
                v = v * fv;
            }
            acc = acc + v;
            touched = true;
        }
    }
    if (live && touched)
        a.d_acc[pix] = acc;
}

} // namespace wb
