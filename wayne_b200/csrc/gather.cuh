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
// Bound: HBM reads of the int32 windows (one coalesced row segment per
// (CTA row, sub-sample) overlap) + one RMW of the fp64 interval plane.
#pragma once
#include "common.cuh"
#include "stage1.cuh"

namespace wb {

constexpr int GX = 32, GY = 8;       // pixel tile of a CTA
constexpr int G_SAMPLES = 128;       // sub-sample descriptors staged per pass

struct GatherSample {
    int ox, oy;
    double x_ref, y_ref, a_t_i, den, m_w, c_w;
};

// grism.py:359-385 for one pixel; X, Y are flat-plane indices (python-style
// negative wrap is applied by the caller before the plane lookups).
__device__ __forceinline__ double flat_value(const GatherSample &g, double Xd, double Yd,
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

__global__ void __launch_bounds__(GX * GY) k_gather(const wb200_gather_args a)
{
    __shared__ GatherSample sm[G_SAMPLES];
    const int r = blockIdx.z;
    const int c_l = blockIdx.x * GX + threadIdx.x; // light-sensitive coordinates
    const int r_l = blockIdx.y * GY + threadIdx.y;
    const int tid = threadIdx.y * GX + threadIdx.x;
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

    size_t pix = ((size_t)r * a.F + (r_l + a.border)) * a.F + (c_l + a.border);
    double acc = live ? a.d_acc[pix] : 0.0;
    bool touched = false;

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

    for (int base = lo; base < hi; base += G_SAMPLES) {
        const int n = min(G_SAMPLES, hi - base);
        __syncthreads();
        for (int i = tid; i < n; i += GX * GY) {
            const int s = base + i;
            GatherSample g;
            g.ox = a.d_win_ox[s];
            g.oy = a.d_win_oy[s];
            const double *t = a.d_trace + (size_t)s * WB200_TRACE_STRIDE;
            g.x_ref = t[0];
            g.y_ref = t[1];
            g.a_t_i = 1 / t[2];
            g.den = g.a_t_i * g.a_t_i + 1;
            g.m_w = t[4];
            g.c_w = t[5];
            sm[i] = g;
        }
        __syncthreads();
        for (int i = 0; i < n; ++i) {
            const GatherSample &g = sm[i];
            // CTA-uniform rejection of windows that miss the pixel tile
            if (g.ox >= tx0 + GX || g.ox + a.win_w <= tx0 || g.oy >= ty0 + GY ||
                g.oy + a.win_h <= ty0)
                continue;
            const int wx = c_l - g.ox, wy = r_l - g.oy;
            if (!live || (unsigned)wx >= (unsigned)a.win_w || (unsigned)wy >= (unsigned)a.win_h)
                continue;
            const int h = a.d_win[((size_t)(base + i) * a.win_h + wy) * a.win_w + wx];
            if (h == 0)
                continue;
            double v = (double)h;
            if (a.add_flat) {
                if (!have_f && flat_ok) {
                    f0 = a.d_flat[0][fidx];
                    f1 = a.d_flat[1][fidx];
                    f2 = a.d_flat[2][fidx];
                    f3 = a.d_flat[3][fidx];
                    have_f = true;
                }
                double fv = flat_value(g, Xd, Yd, f0, f1, f2, f3, a.flat_wmin, a.flat_wmax);
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
