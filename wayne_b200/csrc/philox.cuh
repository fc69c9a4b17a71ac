// philox.cuh -- counter-based random streams of the native (WB200_RNG_PHILOX)
// mode: Philox4x32-10 (Salmon et al., SC'11; Random123 constants), uniform /
// normal / Poisson samplers built on it.  Every draw is addressed by
// (key, counter) only, so results do not depend on launch geometry, batching
// or the number of GPUs.
//
// Counter layout used across the library: {draw index, minor id, major id,
// stream id}; see the WB_STREAM_* ids.
#pragma once
#include "common.cuh"

namespace wb {

enum : uint32_t {
    WB_STREAM_COUNTS = 1,  // Poisson draw of a (sub-sample, bin) cell
    WB_STREAM_PHOTONS = 2, // Box-Muller pairs of a cell's electrons
    WB_STREAM_NOISE = 3,   // per-read background noise normal
    WB_STREAM_SKY = 4,     // per-read sky Poisson
    WB_STREAM_DARK = 5,    // dark-current normal
    WB_STREAM_READ = 6,    // read-noise normal
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// (0,1] style uniforms: (x + 0.5) * 2^-32, never 0.
__device__ __forceinline__ float u01f(uint32_t x)
{
    return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}
__device__ __forceinline__ double u01d(uint32_t x)
{
    return ((double)x + 0.5) * 2.3283064365386963e-10;
}

// fp32 Box-Muller on the SFU: lg2 + sqrt + sin + cos (4 MUFU ops per pair).
// theta is taken in (-pi, pi] where sin.approx/cos.approx are at their stated
// accuracy (abs err 2^-21.4).  |z| <= sqrt(2*33*ln2) = 6.77.
constexpr float WB_ZMAX_PHILOX = 6.8f;
__device__ __forceinline__ void box_muller_f(uint32_t a, uint32_t b, float &zx, float &zy)
{
    const float u1 = u01f(a);
    const float th = fmaf((float)b, 1.4629180792671596e-09f, -3.14159265358979f); // 2pi*2^-32*b - pi
    // -2 ln u = -2 ln2 * lg2(u)
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * __log2f(u1)));
    float s, c;
    __sincosf(th, &s, &c);
    zx = r * c;
    zy = r * s;
}

// the same pair promoted to fp64: the native per-pixel noise terms (read noise,
// dark, background) need the DISTRIBUTION, not 53-bit normals
__device__ __forceinline__ void box_muller_fd(uint32_t a, uint32_t b, double &z0, double &z1)
{
    float x, y;
    box_muller_f(a, b, x, y);
    z0 = (double)x;
    z1 = (double)y;
}

// fp64 normal pair from two 32-bit words (per-pixel noise terms; accuracy over speed)
__device__ __forceinline__ void box_muller_d(uint32_t a, uint32_t b, double &z0, double &z1)
{
    const double u1 = u01d(a);
    const double u2 = u01d(b);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// A small stateful view of one Philox stream: counter.x walks, the other three
// words address the stream.  Used by the Poisson samplers which need an
// unbounded number of uniforms.
struct PhiloxStream {
    uint32_t k0, k1, cy, cz, cw, cx;
    uint4 buf;
    int have;
    // `first` = first block counter of the stream (blocks below it belong to other users of the
    // same (minor, major, stream) triple: the per-pair call of the read kernels is block 1)
    __device__ __forceinline__ PhiloxStream(uint32_t k0_, uint32_t k1_, uint32_t minor,
                                            uint32_t major, uint32_t stream, uint32_t first = 0)
        : k0(k0_), k1(k1_), cy(minor), cz(major), cw(stream), cx(first), have(0)
    {
    }
    __device__ __forceinline__ uint32_t next()
    {
        if (have == 0) {
            buf = philox4x32_10(make_uint4(cx++, cy, cz, cw), k0, k1);
            have = 4;
        }
        uint32_t v = buf.x;
        buf.x = buf.y;
        buf.y = buf.z;
        buf.z = buf.w;
        --have;
        return v;
    }
    __device__ __forceinline__ double uniform() { return u01d(next()); }
};

__constant__ float c_logfact_f[10] = {0.f, 0.f, 0.6931471806f, 1.791759469f, 3.178053830f, 4.787491743f,
                                      6.579251212f, 8.525161361f, 10.60460290f, 12.80182748f};

// lgamma(k+1) for integer-valued k >= 0: table below 10, Stirling series above
// (truncation error < 1e-12 for k >= 10) -- one fp64 log instead of lgamma().
__device__ __forceinline__ double log_factorial(double k)
{
    if (k < 10.0) {
        const double t[10] = {0.0, 0.0, 0.6931471805599453, 1.791759469228055, 3.1780538303479458,
                              4.787491742782046, 6.579251212010101, 8.525161361065415,
                              10.604602902745251, 12.801827480081469};
        return t[(int)k];
    }
    const double x = k + 1.0;
    const double xi = 1.0 / x, xi2 = xi * xi;
    return (k + 0.5) * log(x) - x + 0.9189385332046727 +
           xi * (8.333333333333333e-2 -
                 xi2 * (2.777777777777778e-3 - xi2 * (7.936507936507937e-4 - xi2 * 5.952380952380952e-4)));
}

// Exact Poisson sampler (distribution-exact, not a normal approximation):
//  lam < 10 : multiplication method (Knuth)
//  lam >= 10: PTRS transformed rejection (Hoermann 1993, "The transformed
//             rejection method for generating Poisson random variables") --
//             the same two algorithms numpy's legacy poisson uses, so the
//             distribution matches np.random.poisson (exposure_generator.py:626,495).
// Written for the GPU: ~88% of the proposals leave through the squeeze test,
// which needs no logarithm; log(lam) and the normalisation constant are only
// formed when a lane first reaches the full acceptance test, which itself is
// one log of a ratio plus log_factorial().
__device__ inline long long poisson_draw(PhiloxStream &g, double lam)
{
    if (!(lam > 0.0))
        return 0;
    if (lam < 10.0) {
        const double enlam = exp(-lam);
        long long k = 0;
        double prod = g.uniform();
        while (prod > enlam) {
            prod *= g.uniform();
            ++k;
        }
        return k;
    }
    const double slam = sqrt(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    double invalpha = 0.0, loglam = 0.0;
    bool have = false;
    for (;;) {
        const double U = g.uniform() - 0.5;
        const double V = g.uniform();
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr)
            return (long long)kf;
        if (kf < 0.0 || (us < 0.013 && V > us))
            continue;
        if (!have) {
            invalpha = 1.1239 + 1.1328 / (b - 3.4);
            loglam = log(lam);
            have = true;
        }
        // log(V) + log(invalpha) - log(a/us^2 + b) <= -lam + k log(lam) - lgamma(k+1)
        if (log(V * invalpha / (a / (us * us) + b)) <= (-lam + kf * loglam - log_factorial(kf)))
            return (long long)kf;
    }
}

// The same PTRS sampler with the expensive pieces in fp32 on the SFU -- used by
// the native kernels, where only the DISTRIBUTION matters (tests/test_rng_gpu.py
// checks it against the exact pmf by chi-square):
//   * the proposal constants (sqrt, reciprocals) and 2a/us are fp32; the same
//     rounded values enter both the proposal and the acceptance test;
//   * U, V, us and k stay fp64 (exact uniforms, integer-exact k for any lam);
//   * the acceptance test's right-hand side is rewritten around m = k+1 as
//       (m - lam) - k*log1p((m-lam)/lam) - 0.5*log(m) - C - Stirling(1/m),
//     whose fp32 rounding error scales with |m - lam| ~ sqrt(lam) rather than
//     with lam*log(lam), so lg2.approx is accurate enough (error <~ 1e-5).
// lam >= 4e6 (never reached by photon counts per cell in practice) and lam < 10
// use the fp64 sampler above.
__device__ __forceinline__ float log1p_f(float d)
{
    const float u = 1.0f + d;
    const float lu = __logf(u);
    // Kahan: log1p(d) = log(u) * d / (u - 1) restores the bits lost in 1 + d
    return (u == 1.0f) ? d : lu * __fdividef(d, u - 1.0f);
}

// Small means (lam < 40: the sky per read interval, faint spectral bins): CDF
// inversion by chop-down search from 0 in fp32 -- ONE uniform, lam + 1 multiply-
// add steps on average, no logarithm, so a warp is not dragged through PTRS's
// acceptance test (which for lam ~ 16 rejects the squeeze 55% of the time).
// Probabilities carry ~1e-6 relative error (fp32 running sum), far below what
// any test of the distribution can resolve.
__device__ __forceinline__ float rcp_ftz(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ long long poisson_inversion_u(float u, float lam)
{
    // u is clamped below 1 (a double uniform can round to 1.0f) and the search
    // stops once the terms no longer move the fp32 sum: the quantile is then
    // ~lam + 5.5 sigma, i.e. the tail is truncated at the 1 - 6e-8 level.
    // (reads_native.cuh tabulates exactly these partial sums: keep the operations
    // of the recurrence -- p *= lam * rcp(k); s += p -- identical there.)
    u = fminf(u, 0.99999994f);
    float p = expf(-lam), s = p;
    int k = 0;
    while (u > s) {
        ++k;
        p *= lam * rcp_ftz((float)k);
        s += p;
        if (p < 1e-10f && (float)k > lam)
            break;
    }
    return k;
}
__device__ __forceinline__ long long poisson_inversion_f(PhiloxStream &g, float lam)
{
    return poisson_inversion_u((float)g.uniform(), lam);
}

__device__ inline long long poisson_draw_fast(PhiloxStream &g, double lam)
{
    if (!(lam > 0.0))
        return 0;
    if (lam < 40.0)
        return poisson_inversion_f(g, (float)lam);
    if (lam >= 4.0e6)
        return poisson_draw(g, lam);
    const float lamf = (float)lam;
    float slam;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(slam) : "f"(lamf));
    const float b = fmaf(2.53f, slam, 0.931f);
    const float a = fmaf(0.02483f, b, -0.059f);
    const float vr = 0.9277f - __fdividef(3.6224f, b - 2.0f);
    const float two_a = 2.0f * a;
    float invalpha = 0.f, inv_lam = 0.f;
    bool have = false;
    for (;;) {
        const double U = g.uniform() - 0.5;
        const double V = g.uniform();
        const double us = 0.5 - fabs(U);
        const float usf = (float)us;
        const float t = fmaf(__fdividef(two_a, usf), (float)U, b * (float)U);
        const double kf = floor((double)t + lam + 0.43);
        if (usf >= 0.07f && (float)V <= vr)
            return (long long)kf;
        if (kf < 0.0 || (usf < 0.013f && V > us))
            continue;
        if (!have) {
            invalpha = 1.1239f + __fdividef(1.1328f, b - 3.4f);
            inv_lam = __fdividef(1.0f, lamf);
            have = true;
        }
        const float lhs = __logf(__fdividef((float)V * invalpha, fmaf(a, __fdividef(1.0f, usf * usf), b)));
        const double m = kf + 1.0;
        const float dm = (float)(m - lam);          // exact difference, then rounded
        const float mf = (float)m;
        const float xi = __fdividef(1.0f, mf), xi2 = xi * xi;
        float rhs;
        if (kf < 10.0) {
            // only reachable for lam < ~30, where every term is O(10): fp32 is
            // accurate to ~1e-6 here, and no lane drags its warp through an fp64 log
            rhs = fmaf((float)kf, __logf(lamf), -lamf) - c_logfact_f[(int)kf];
        } else {
            const float series = xi * (8.333333333e-2f - xi2 * (2.777777778e-3f - xi2 * 7.936507937e-4f));
            rhs = dm - (float)kf * log1p_f(dm * inv_lam) - 0.5f * __logf(mf) - 0.9189385332f - series;
        }
        if (lhs <= rhs)
            return (long long)kf;
    }
}

} // namespace wb
