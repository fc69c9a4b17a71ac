// reads_native.cuh -- stage 4 for the native (Philox) mode: the same fused
// per-pixel ramp pass as reads.cuh::k_reads<0, FAST> (same Philox counters, same
// order of terms: wayne/exposure_generator.py:468-515, 361-389, 407-444;
// wayne/exposure.py:49-131; wayne/detector.py:151-198, 318-350), written for
// throughput.  It takes no host-drawn planes (those belong to the parity mode,
// which stays on k_reads).  What differs from the generic kernel:
//
//  * sky Poisson: a pixel draws one Poisson of the SAME mean per read interval,
//    so the CDF window of the chop-down recurrence is built once per (pixel,
//    mean) in shared memory and every draw is a 6-probe binary search.  The
//    serial walk it replaces cost ~lam + 3 sigma dependent steps per draw, at the
//    pace of the slowest lane of the warp -- 57 % of the generic kernel's
//    instructions on the C4 workload (profiles/README.md).  The table holds the
//    very partial sums the walk forms, so each draw equals
//    poisson_inversion_u(u, lam) exactly; draws outside the window take the walk.
//  * int64 fixed point -> double by the 2^52 magic add (no I2F.F64.S64 on the
//    XU pipe), Newton steps with FMAs and an fp32-seeded reciprocal, the ftz SFU
//    forms for the normals, derivative coefficients formed from c2..c4 (three
//    planes fewer to read);
//  * next read's interval / dark / dark-error loads are issued before this
//    read's arithmetic (software pipeline), 128 threads per CTA, <= 96 registers.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "photons.cuh"
#include "reads.cuh"

namespace wb {

constexpr int RN_THREADS = 128;
#ifndef RN_UNROLL
#define RN_UNROLL 1 // reads per trip of the ramp loop (2 measured slower: 0.228 against 0.212 ms; so did 4 CTAs per SM)
#endif
constexpr int RN_UNROLL_N = RN_UNROLL;
#ifndef RN_MIN_BLOCKS
#define RN_MIN_BLOCKS 5
#endif
constexpr int SKY_T = 32;            // CDF window entries per pixel (32 KB per CTA)
constexpr float SKY_LAM_MAX = 24.0f; // means below this are one draw from the window
constexpr size_t RN_SMEM = sizeof(float) * SKY_T * 2 * RN_THREADS + sizeof(double) * 10 * RN_THREADS;

// exact for |q| < 2^51: q + bits(1.5 * 2^52), then subtract 1.5 * 2^52
__device__ __forceinline__ double ll2d_fast(long long q)
{
    // |q| < 2^51 <=> the high word lies in [-2^19, 2^19): one add and one unsigned compare
    if ((unsigned)((int)(q >> 32) + (1 << 19)) < (1u << 20))
        return __longlong_as_double(q + 0x4338000000000000LL) - 6755399441055744.0;
    return (double)q;
}

// fp32 Box-Muller, ftz SFU forms (u1 >= 2^-33 is never denormal)
__device__ __forceinline__ void box_muller_ftz(uint32_t a, uint32_t b, float &zx, float &zy)
{
    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float th = fmaf((float)b, 1.4629180792671596e-09f, -3.14159265358979f);
    const float r = sqrt_approx(-1.3862943611198906f * lg2_approx(u1));
    zx = r * cos_approx(th);
    zy = r * sin_approx(th);
}

struct SkyWindow {
    float lam;   // mean the window was built for (-1: none)
    float below; // partial sum at k0 - 1 (-1 when k0 == 0)
    float top;   // partial sum at k0 + SKY_T - 1
    int k0;
};

// tbl points at this thread's column; consecutive k are 2*RN_THREADS floats apart
// ([k][pixel h][thread]: bank = thread % 32 whatever k a lane probes).
// The recurrence is poisson_inversion_u's, operation for operation.
__device__ __forceinline__ void sky_window_build(float *tbl, SkyWindow &t, float lam)
{
    const int k0 = max(0, (int)(lam - 2.5f * sqrt_approx(lam)) - 1);
    float p = expf(-lam), s = 0.0f, below = -1.0f, mult = 1.0f;
    if (k0 > 0) {
        s = p;
        for (int k = 1; k < k0; ++k) {
            p *= lam * rcp_ftz((float)k);
            s += p;
        }
        below = s;
        mult = lam * rcp_ftz((float)k0);
    }
    p *= mult; // k = k0 (for k0 == 0: p0 itself, s0 = 0 + p0)
    s += p;
    tbl[0] = s;
    float kf = (float)k0;
#pragma unroll
    for (int j = 1; j < SKY_T; ++j) {
        kf += 1.0f;
        p *= lam * rcp_ftz(kf);
        s += p;
        tbl[j * (2 * RN_THREADS)] = s;
    }
    t.lam = lam;
    t.below = below;
    t.top = s;
    t.k0 = k0;
}

__device__ __noinline__ int sky_serial_walk(float u, float lam)
{
    return (int)poisson_inversion_u(u, lam);
}

// One Poisson(t.lam) draw: identical to poisson_inversion_u(u, t.lam) for every u.
__device__ __forceinline__ int sky_window_draw(const float *tbl, const SkyWindow &t, float u)
{
    u = fminf(u, 0.99999994f);
    if (!(u > t.below) || u > t.top)
        return sky_serial_walk(u, t.lam); // outside the window (rare)
    // j = number of entries < u (<= SKY_T-1 because tbl[SKY_T-1] >= u)
    int j = 0;
#pragma unroll
    for (int step = SKY_T / 2; step >= 1; step >>= 1) {
        WB_DEV_ASSERT(j + step - 1 >= 0 && j + step - 1 < SKY_T);
        if (tbl[(j + step - 1) * (2 * RN_THREADS)] < u)
            j += step;
    }
    WB_DEV_ASSERT(j < SKY_T && tbl[j * (2 * RN_THREADS)] >= u);
    return t.k0 + j;
}

// A window built for mean t.lam also serves every mean in [t.lam, t.lam + SKY_REUSE]:
// Poisson(lam) = Poisson(t.lam) + Poisson(lam - t.lam), independent terms, and the
// second is almost always 0 (read intervals of one ramp differ by ~1 ms: the window
// is built once or twice per pixel instead of once per read).
constexpr float SKY_REUSE = 0.25f;

// Poisson(lam) for any lam > 0 from uniform words: u (the window draw) and u2 (the
// small remainder); means >= SKY_LAM_MAX are the sum of m window draws of mean ~lam/m
// (Poisson is closed under sums), their uniforms from extra Philox calls.
__device__ __forceinline__ int sky_draw(float *tbl, SkyWindow &t, float lam, uint32_t w, uint32_t w2,
                                        uint32_t pid, uint32_t r, uint32_t key0, uint32_t key1)
{
    int tot;
    float rest;
    if (lam < SKY_LAM_MAX) {
        rest = lam - t.lam;
        if (!(rest >= 0.0f && rest <= SKY_REUSE)) {
            sky_window_build(tbl, t, lam);
            rest = 0.0f;
        }
        tot = sky_window_draw(tbl, t, (float)u01d(w));
    } else {
        const int m = (int)(lam * 0.0625f) + 1;
        rest = fmaf(-(float)m, t.lam, lam);
        if (!(rest >= 0.0f && rest <= SKY_REUSE)) {
            sky_window_build(tbl, t, lam * rcp_ftz((float)m));
            rest = fmaxf(fmaf(-(float)m, t.lam, lam), 0.0f);
        }
        tot = sky_window_draw(tbl, t, (float)u01d(w));
        uint4 q = make_uint4(0, 0, 0, 0);
        for (int j = 1; j < m; ++j) {
            if ((j & 3) == 1)
                q = philox4x32_10(make_uint4(2u + (uint32_t)(j >> 2), pid, r, WB_STREAM_SKY), key0, key1);
            tot += sky_window_draw(tbl, t, (float)u01d(q.x));
            q = make_uint4(q.y, q.z, q.w, q.x);
        }
    }
    if (rest > 0.0f)
        tot += (int)poisson_inversion_u((float)u01d(w2), rest);
    return tot;
}

__device__ __noinline__ double cosmic_energy(const int32_t *next, const int32_t *read,
                                             const double *energy, int c, int r)
{
    double e = 0.0;
    for (; c >= 0; c = next[c])
        if (read[c] == r)
            e += energy[c];
    return e;
}

struct NLCoef4 {
    double b0, c2, c3, c4;
};

// detector.py:318-350 per pixel: Newton on u*(b0 + c2 u + c3 u^2 + c4 u^3) = p from u = p
// until the step is < 1e-3 (the reference's threshold), with FMAs and a reciprocal
// seeded in fp32 (relative error 6e-8 of a step that ends below 1e-3).
__device__ __forceinline__ double newton_native(double p, const NLCoef4 &k)
{
    double u = p;
    for (int it = 0; it < 64; ++it) {
        // value and derivative in one Horner sweep (no 2*c2, 3*c3, 4*c4 constants)
        const double A = fma(k.c4, u, k.c3);
        const double Bv = fma(u, A, k.c2);
        const double C = fma(u, Bv, k.b0);
        const double f = fma(u, C, -p);
        const double d = fma(u, fma(u, fma(k.c4, u, A), Bv), C);
        float ri;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ri) : "f"((float)d));
        const double step = f * (double)ri;
        u = u - step;
        if (!(fabs(step) >= 1e-3)) // also leaves on NaN
            break;
    }
    return u;
}

// a pair of adjacent plane values promoted to double: PLANES32 = the planes are held in the
// calibration files' own float32 (every product the reference forms with them -- float32 sky x
// rate, float32 2.35 / pfl, float32 1 + c1, the float32 dark-error floor -- is float32-valued, so
// the promotion is exact and half the bytes move)
template <bool PLANES32>
__device__ __forceinline__ double2 ld_plane2(const void *base, size_t idx)
{
    if (PLANES32) {
        float2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];"
                     : "=f"(r.x), "=f"(r.y)
                     : "l"(reinterpret_cast<const float *>(base) + idx));
        return make_double2((double)r.x, (double)r.y);
    }
    return ld_stream2(reinterpret_cast<const double *>(base) + idx);
}

__device__ __forceinline__ float2 ld_plane2_f32(const void *base, size_t idx)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];"
                 : "=f"(r.x), "=f"(r.y)
                 : "l"(reinterpret_cast<const float *>(base) + idx));
    return r;
}

// np.clip: a NaN stays a NaN (fmin / fmax would return the bound), one compare and one select per bound
__device__ __forceinline__ double clip_np(double v, double lo, double hi)
{
    v = (v < lo) ? lo : v;
    return (v > hi) ? hi : v;
}

template <bool OUT32, bool PLANES32>
__global__ void __launch_bounds__(RN_THREADS, RN_MIN_BLOCKS) k_reads_native(const wb200_reads_args a)
{
    extern __shared__ float s_sky[]; // [SKY_T][2][RN_THREADS] floats, then [10][RN_THREADS] doubles
    const int F = a.F, B = a.border, R = a.n_reads;
    const int half = F >> 1;
    const uint32_t plane = (uint32_t)F * (uint32_t)F;
    const uint32_t idx = blockIdx.x * RN_THREADS + threadIdx.x;
    if (idx >= plane / 2)
        return;
    const int Y = (int)(idx / (uint32_t)half);
    const int X = (int)(idx - (uint32_t)Y * (uint32_t)half) * 2;
    const uint32_t p = (uint32_t)Y * (uint32_t)F + (uint32_t)X;
    WB_DEV_ASSERT(p + 1 < plane && (p & 1) == 0);
    const bool rowin = Y >= B && Y < F - B;
    const bool in0 = rowin && X >= B && X < F - B;
    const bool in1 = rowin && (X + 1) >= B && (X + 1) < F - B;

    const bool sky_on = a.add_sky != 0, dark_on = a.add_dark != 0;
    const bool acc_fixed = a.acc_fixed != 0;

    // ---- per-pixel constants -------------------------------------------------
    double ginv[2] = {1.0 / a.const_gain, 1.0 / a.const_gain};
    if (a.d_gain) {
        const double2 g = ld_plane2<PLANES32>(a.d_gain, p);
        ginv[0] = 1.0 / g.x;
        ginv[1] = 1.0 / g.y;
    }
    float skyf[2] = {0.f, 0.f};
    double skyd[2] = {0., 0.};
    if (sky_on) {
        const double2 s = ld_plane2<PLANES32>(a.d_sky, p);
        skyd[0] = s.x;
        skyd[1] = s.y;
        skyf[0] = __double2float_rn(s.x);
        skyf[1] = __double2float_rn(s.y);
    }
    int chead[2] = {-1, -1};
    if (a.d_cos_head) {
        const int2 t = *reinterpret_cast<const int2 *>(a.d_cos_head + p);
        chead[0] = t.x;
        chead[1] = t.y;
    }
    double zc[2] = {0., 0.};
    if (a.d_zero) {
        const double2 z = ld_stream2(a.d_zero + p);
        zc[0] = z.x;
        zc[1] = z.y;
    }
    if (a.clip) {
        zc[0] = clip_np(zc[0], a.clip_lo, a.clip_hi);
        zc[1] = clip_np(zc[1], a.clip_lo, a.clip_hi);
    }
    if (!in0)
        zc[0] = 0.0;
    if (!in1)
        zc[1] = 0.0;
    // per-pixel constants used once per read live in shared memory, not in registers
    double *sc = reinterpret_cast<double *>(s_sky + SKY_T * 2 * RN_THREADS) + threadIdx.x;
    if (a.add_nonlinear) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 t = ld_plane2<PLANES32>(a.d_nl[i], p);
            sc[(2 * i) * RN_THREADS] = t.x;
            sc[(2 * i + 1) * RN_THREADS] = t.y;
        }
    }
    sc[8 * RN_THREADS] = zc[0];
    sc[9 * RN_THREADS] = zc[1];

    SkyWindow win[2];
    win[0].lam = win[1].lam = -1.0f;
    float *tbl = s_sky + threadIdx.x;

    // ---- software pipeline: loads of read r+1 are in flight during read r -------
    longlong2 nq = make_longlong2(0, 0);
    double2 nacc = make_double2(0., 0.), ndk = make_double2(0., 0.), nde = make_double2(0., 0.);
    float2 ndk32 = make_float2(0.f, 0.f), nde32 = make_float2(0.f, 0.f); // PLANES32: promoted at use
    double ndt = 0.0;
    auto issue = [&](int r) {
        const size_t off = (size_t)r * plane + p;
        ndt = a.d_dt[r];
        if (acc_fixed) {
            longlong2 *q = reinterpret_cast<longlong2 *>(
                reinterpret_cast<long long *>(const_cast<void *>(a.d_acc)) + off);
            nq = *q;
        } else
            nacc = ld_stream2(reinterpret_cast<const double *>(a.d_acc) + off);
        if (dark_on) {
            if (PLANES32) {
                ndk32 = ld_plane2_f32(a.d_dark, off);
                nde32 = ld_plane2_f32(a.d_dark_err, off);
            } else {
                ndk = ld_plane2<false>(a.d_dark, off);
                nde = ld_plane2<false>(a.d_dark_err, off);
            }
        }
    };
    issue(0);

    double cum[2] = {0., 0.};
#pragma unroll RN_UNROLL_N
    for (int r = 0; r < R; ++r) {
        double acc[2];
        if (acc_fixed) {
            acc[0] = ll2d_fast(nq.x) * (1.0 / 16777216.0);
            acc[1] = ll2d_fast(nq.y) * (1.0 / 16777216.0);
            // the interval planes are left zeroed for the next exposure (no memset pass).  The store
            // comes HERE, once this read's load has been consumed: right behind the load it would
            // have to wait for the load's data (same address) and stall the software pipeline.
            if (a.zero_acc)
                *reinterpret_cast<longlong2 *>(reinterpret_cast<long long *>(const_cast<void *>(a.d_acc)) +
                                               (size_t)r * plane + p) = make_longlong2(0, 0);
        } else {
            acc[0] = nacc.x;
            acc[1] = nacc.y;
        }
        const double2 dk = PLANES32 ? make_double2((double)ndk32.x, (double)ndk32.y) : ndk;
        const double2 de = PLANES32 ? make_double2((double)nde32.x, (double)nde32.y) : nde;
        const double dt = ndt;
        if (r + 1 < R)
            issue(r + 1);

        // one Philox call per (pixel pair, read): x, y = the two sky uniforms, (z, w) =
        // the pair's dark normals -- the counters of k_reads<0, FAST>
        uint4 qs = make_uint4(0, 0, 0, 0);
        if (sky_on || dark_on)
            qs = philox4x32_10(make_uint4(1u, p, (uint32_t)r, WB_STREAM_SKY), a.key0, a.key1);
        // the read-noise call of this read: (x, y) its normals, z / w the two pixels' sky
        // remainder uniforms
        uint4 qr = make_uint4(0, 0, 0, 0);
        if (a.add_read_noise || sky_on)
            qr = philox4x32_10(make_uint4(0, p, (uint32_t)(r + 1), WB_STREAM_READ), a.key0, a.key1);
        float lamf[2] = {0.f, 0.f};
        double lamd[2] = {0., 0.};
        if (sky_on) {
            const double bg = a.sky_rate * dt;
            if (a.sky_f32) {
                const float bgf = __double2float_rn(bg);
                lamd[0] = (double)__fmul_rn(skyf[0], bgf);
                lamd[1] = (double)__fmul_rn(skyf[1], bgf);
            } else {
                lamd[0] = skyd[0] * bg;
                lamd[1] = skyd[1] * bg;
            }
            lamf[0] = (float)lamd[0];
            lamf[1] = (float)lamd[1];
        }
        double v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double px = 0.0;
            if (h ? in1 : in0) {
                px = acc[h];
                if (a.add_noise) {
                    const uint4 q = philox4x32_10(make_uint4(0, p + h, (uint32_t)r, WB_STREAM_NOISE),
                                                  a.key0, a.key1);
                    float z0, z1;
                    box_muller_ftz(q.x, q.y, z0, z1);
                    px = px + (a.noise_mean * dt + (a.noise_std * dt) * (double)z0);
                }
                if (sky_on) {
                    if (!(lamd[h] > 0.0)) {
                        // no sky photons
                    } else if (!(lamf[h] < 1.0e6f)) {
                        px = px + lamd[h]; // beyond any physical sky level: the mean itself
                    } else {
                        px = px + (double)sky_draw(tbl + h * RN_THREADS, win[h], lamf[h], h ? qs.y : qs.x,
                                                   h ? qr.w : qr.z, p + h, (uint32_t)r, a.key0, a.key1);
                    }
                }
                if (chead[h] >= 0)
                    px = px + cosmic_energy(a.d_cos_next, a.d_cos_read, a.d_cos_energy, chead[h], r);
                px = px * ginv[h];
            }
            cum[h] = cum[h] + px;
            v[h] = cum[h];
        }
        if (dark_on) {
            float z0, z1;
            box_muller_ftz(qs.z, qs.w, z0, z1);
            const double sd0 = (de.x > 0) ? de.x : 0.00001; // detector.py:189-190
            const double sd1 = (de.y > 0) ? de.y : 0.00001;
            v[0] = v[0] + (dk.x + sd0 * (double)z0);
            v[1] = v[1] + (dk.y + sd1 * (double)z1);
        }
        if (a.add_nonlinear) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const NLCoef4 nl = {sc[h * RN_THREADS], sc[(2 + h) * RN_THREADS], sc[(4 + h) * RN_THREADS],
                                    sc[(6 + h) * RN_THREADS]};
                v[h] = newton_native(v[h], nl);
            }
        }
        if (a.clip) {
            v[0] = clip_np(v[0], a.clip_lo, a.clip_hi);
            v[1] = clip_np(v[1], a.clip_lo, a.clip_hi);
        }
        if (!in0)
            v[0] = 0.0; // reset_reference_pixels (exposure.py:122-131)
        if (!in1)
            v[1] = 0.0;
        v[0] = v[0] + sc[8 * RN_THREADS]; // add_zero_read (exposure.py:94-104)
        v[1] = v[1] + sc[9 * RN_THREADS];
        if (a.add_read_noise) {
            float z0, z1;
            box_muller_ftz(qr.x, qr.y, z0, z1);
            v[0] = v[0] + a.read_noise * (double)z0; // detector.py:198
            v[1] = v[1] + a.read_noise * (double)z1;
        }
        const size_t o = (size_t)(r + 1) * plane + p;
        if (OUT32)
            *reinterpret_cast<float2 *>(reinterpret_cast<float *>(a.d_out) + o) =
                make_float2((float)v[0], (float)v[1]);
        else
            st_stream2(reinterpret_cast<double *>(a.d_out) + o, make_double2(v[0], v[1]));
    }

    // the zero read itself: clipped, border-zeroed, then read noise
    double v0 = sc[8 * RN_THREADS], v1 = sc[9 * RN_THREADS];
    if (a.add_read_noise) {
        const uint4 q = philox4x32_10(make_uint4(0, p, 0u, WB_STREAM_READ), a.key0, a.key1);
        float z0, z1;
        box_muller_ftz(q.x, q.y, z0, z1);
        v0 = v0 + a.read_noise * (double)z0;
        v1 = v1 + a.read_noise * (double)z1;
    }
    if (OUT32)
        *reinterpret_cast<float2 *>(reinterpret_cast<float *>(a.d_out) + p) = make_float2((float)v0, (float)v1);
    else
        st_stream2(reinterpret_cast<double *>(a.d_out) + p, make_double2(v0, v1));
}

} // namespace wb
