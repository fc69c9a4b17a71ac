// reads.cuh -- stage 4: the fused per-pixel pass over the whole ramp.
//
// Replaces _add_read_reductions (wayne/exposure_generator.py:468-515), the
// cumulative read stack (:361-389), _post_exposure_reductions (:407-444) and
// the Exposure / WFC3_IR methods they call (wayne/exposure.py:49-131,
// wayne/detector.py:151-198, 318-350).  The reference makes ~12 full-frame
// numpy passes per read; here one thread owns two adjacent pixels for all
// NSAMP reads, keeps the cumulative signal in registers, reads every
// calibration plane exactly once (128-bit loads) and writes every read exactly
// once (128-bit stores).  Pure HBM streaming: the roofline is
//   bytes = R*F^2*8 (interval planes) + planes read once + (R+1)*F^2*8 written.
//
// Operation order per pixel and read (fp64, no fused multiply-add):
//   px = acc [+ noise] [+ sky] [+ cosmics]; px /= gain; (border px = 0)
//   cum += px; v = cum [+ dark]; [v = nonlinear(v)]; [clip]; border -> 0;
//   v += zero_read'; [v += rn * z]            (zero_read' = clipped, border-zeroed)
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace wb {

struct NLCoef {
    double b0, c2, c3, c4, d2, d3, d4; // 1+c1, c2, c3, c4, 2*c2, 3*c3, 4*c4
};

// One Newton step of detector.py:339-343, same operation order.
__device__ __forceinline__ double newton_step(double p, double u0, const NLCoef &k)
{
    const double num = -p + u0 * (k.b0 + u0 * (k.c2 + u0 * (k.c3 + k.c4 * u0)));
    const double den = k.b0 + k.d2 * u0 + k.d3 * u0 * u0 + k.d4 * u0 * u0 * u0;
    return u0 - (num / den);
}

constexpr int NEWTON_LIMIT = 10000; // detector.py:338

// iterate until this pixel moves < 1e-3; returns the value, *iters = steps taken
__device__ __forceinline__ double newton_converge(double p, const NLCoef &k, int *iters)
{
    double u0 = p, u1 = p;
    int it = 0;
    while (it < NEWTON_LIMIT) {
        u1 = newton_step(p, u0, k);
        ++it;
        if (fabs(u1 - u0) < 1e-3)
            break;
        if (isnan(u1)) {
            it = NEWTON_LIMIT;
            break;
        }
        u0 = u1;
    }
    *iters = it;
    return u1;
}

__device__ __forceinline__ double newton_fixed(double p, const NLCoef &k, int n)
{
    double u0 = p, u1 = p;
    for (int it = 0; it < n; ++it) {
        u1 = newton_step(p, u0, k);
        u0 = u1;
    }
    return u1;
}

__device__ __forceinline__ void load_pair(const double *plane, size_t i, double v[2])
{
    const double2 t = ld_stream2(plane + i);
    v[0] = t.x;
    v[1] = t.y;
}

// PASS 0: everything fused, per-pixel Newton convergence (native mode).
// PASS 1: exact mode, first half -- stores the pre-non-linearity ramp and
//         records, per read, how many Newton steps the slowest pixel needs
//         (the reference stops all pixels together, detector.py:344).
// PASS 2: exact mode, second half -- applies exactly that many steps to every
//         pixel, then clip / border / zero read / read noise.
// FAST (native mode only): fp32-SFU Box-Muller for the noise normals and a
// reciprocal gain -- distributions unchanged, values not bit-tied to numpy.
template <int PASS, bool FAST = false>
__global__ void __launch_bounds__(256) k_reads(const wb200_reads_args a)
{
    const int F = a.F, B = a.border, R = a.n_reads;
    const int half = F >> 1;
    const size_t plane = (size_t)F * F;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < plane / 2;
    const int Y = valid ? (int)(idx / half) : 0;
    const int X = valid ? (int)(idx - (size_t)Y * half) * 2 : 0;
    const size_t p = (size_t)Y * F + X;
    __shared__ int s_iter[16];
    if (PASS == 1) {
        if (threadIdx.x < 16)
            s_iter[threadIdx.x] = 0;
        __syncthreads();
    }

    bool inside[2];
    for (int h = 0; h < 2; ++h)
        inside[h] = valid && Y >= B && Y < F - B && (X + h) >= B && (X + h) < F - B;

    double sky[2] = {0, 0}, gain[2] = {a.const_gain, a.const_gain}, zero[2] = {0, 0};
    NLCoef nl[2];
    int chead[2] = {-1, -1};
    if (valid) {
        if (PASS != 2) {
            if (a.add_sky && !a.d_draw_sky)
                load_pair(reinterpret_cast<const double *>(a.d_sky), p, sky);
            if (a.d_gain)
                load_pair(reinterpret_cast<const double *>(a.d_gain), p, gain);
            if (a.d_cos_head) {
                const int2 t = *reinterpret_cast<const int2 *>(a.d_cos_head + p);
                chead[0] = t.x;
                chead[1] = t.y;
            }
        }
        if (a.d_zero && PASS != 1)
            load_pair(a.d_zero, p, zero);
        if (a.add_nonlinear) {
            double t[7][2];
            for (int i = 0; i < 7; ++i)
                load_pair(reinterpret_cast<const double *>(a.d_nl[i]), p, t[i]);
            for (int h = 0; h < 2; ++h)
                nl[h] = NLCoef{t[0][h], t[1][h], t[2][h], t[3][h], t[4][h], t[5][h], t[6][h]};
        }
    }

    if (FAST) {
        gain[0] = 1.0 / gain[0];
        gain[1] = 1.0 / gain[1];
    }
    // zero read as it is after clip + reference-pixel reset (exposure.py:82-131)
    double zc[2];
    for (int h = 0; h < 2; ++h) {
        double z = zero[h];
        if (a.clip)
            z = fmin(fmax(z, a.clip_lo), a.clip_hi);
        if (!inside[h])
            z = 0.0;
        zc[h] = z;
    }

    double cum[2] = {0, 0};
    for (int r = 0; r < R; ++r) {
        double v[2] = {0, 0};
        int iters = 0;
        if (valid) {
            if (PASS != 2) {
                const double dt = a.d_dt[r];
                double acc[2], dn[2] = {0, 0}, ds[2] = {0, 0};
                if (a.acc_fixed) {
                    const longlong2 q = *reinterpret_cast<const longlong2 *>(
                        reinterpret_cast<const long long *>(a.d_acc) + (size_t)r * plane + p);
                    acc[0] = (double)q.x * (1.0 / 16777216.0);
                    acc[1] = (double)q.y * (1.0 / 16777216.0);
                } else {
                    load_pair(reinterpret_cast<const double *>(a.d_acc) + (size_t)r * plane, p, acc);
                }
                if (a.add_noise && a.d_draw_noise)
                    load_pair(a.d_draw_noise + (size_t)r * plane, p, dn);
                if (a.add_sky && a.d_draw_sky)
                    load_pair(a.d_draw_sky + (size_t)r * plane, p, ds);
                // FAST: one Philox call per (pixel pair, read) feeds both pixels' sky
                // uniforms (x, y) and the pair's dark normals (z, w)
                uint4 qs = make_uint4(0, 0, 0, 0);
                if (FAST && ((a.add_sky && !a.d_draw_sky) || (a.add_dark && !a.d_draw_dark)))
                    qs = philox4x32_10(make_uint4(1u, (uint32_t)p, (uint32_t)r, WB_STREAM_SKY), a.key0, a.key1);
                for (int h = 0; h < 2; ++h) {
                    double px = 0.0;
                    if (inside[h]) {
                        const uint32_t pid = (uint32_t)(p + h);
                        px = acc[h];
                        if (a.add_noise) {
                            if (a.d_draw_noise) {
                                px = px + dn[h];
                            } else {
                                const uint4 q = philox4x32_10(
                                    make_uint4(0, pid, (uint32_t)r, WB_STREAM_NOISE), a.key0, a.key1);
                                double z0, z1;
                                if (FAST)
                                    box_muller_fd(q.x, q.y, z0, z1);
                                else
                                    box_muller_d(q.x, q.y, z0, z1);
                                px = px + (a.noise_mean * dt + (a.noise_std * dt) * z0);
                            }
                        }
                        if (a.add_sky) {
                            if (a.d_draw_sky) {
                                px = px + ds[h];
                            } else {
                                // FAST: block 1 of pixel p is the pair's shared call `qs` above (its z, w feed
                                // the dark normals): the large-mean sampler's own blocks start at 2, as in
                                // reads_native.cuh, so no word serves two noise terms
                                PhiloxStream g(a.key0, a.key1, pid, (uint32_t)r, WB_STREAM_SKY, FAST ? 2u : 0u);
                                const double bg = a.sky_rate * dt;
                                const double lam = a.sky_f32
                                    ? (double)__fmul_rn(__double2float_rn(sky[h]), __double2float_rn(bg))
                                    : sky[h] * bg;
                                if (FAST && lam < 40.0)
                                    px = px + (double)poisson_inversion_u((float)u01d(h ? qs.y : qs.x), (float)lam);
                                else
                                    px = px + (double)(FAST ? poisson_draw_fast(g, lam) : poisson_draw(g, lam));
                            }
                        }
                        if (chead[h] >= 0) {
                            double e = 0.0;
                            for (int c = chead[h]; c >= 0; c = a.d_cos_next[c])
                                if (a.d_cos_read[c] == r)
                                    e += a.d_cos_energy[c];
                            px = px + e;
                        }
                        px = FAST ? px * gain[h] : px / gain[h];
                    }
                    cum[h] = cum[h] + px;
                    v[h] = cum[h];
                }
                if (a.add_dark) {
                    if (a.d_draw_dark) {
                        double dd[2];
                        load_pair(a.d_draw_dark + (size_t)r * plane, p, dd);
                        v[0] = v[0] + dd[0];
                        v[1] = v[1] + dd[1];
                    } else {
                        double dk[2], de[2];
                        load_pair(reinterpret_cast<const double *>(a.d_dark) + (size_t)r * plane, p, dk);
                        load_pair(reinterpret_cast<const double *>(a.d_dark_err) + (size_t)r * plane, p, de);
                        double z[2];
                        if (FAST) {
                            box_muller_fd(qs.z, qs.w, z[0], z[1]);
                        } else {
                            const uint4 q = philox4x32_10(
                                make_uint4(0, (uint32_t)p, (uint32_t)r, WB_STREAM_DARK), a.key0, a.key1);
                            box_muller_d(q.x, q.y, z[0], z[1]);
                        }
                        for (int h = 0; h < 2; ++h) {
                            const double sd = (de[h] > 0) ? de[h] : 0.00001; // detector.py:189-190
                            v[h] = v[h] + (dk[h] + sd * z[h]);
                        }
                    }
                }
            } else {
                load_pair(reinterpret_cast<const double *>(a.d_out) + (size_t)(r + 1) * plane, p, v);
            }

            if (a.add_nonlinear) {
                for (int h = 0; h < 2; ++h) {
                    if (PASS == 0) {
                        int it;
                        v[h] = newton_converge(v[h], nl[h], &it);
                    } else if (PASS == 1) {
                        int it;
                        (void)newton_converge(v[h], nl[h], &it);
                        iters = max(iters, it);
                    } else {
                        v[h] = newton_fixed(v[h], nl[h], a.d_newton_iters[r]);
                    }
                }
            }
        }
        if (PASS == 1) {
            if (a.add_nonlinear) {
                iters = warp_max_i(iters);
                if (lane_id() == 0 && iters)
                    atomicMax(&s_iter[r & 15], iters);
            }
            if (valid)
                st_stream2(reinterpret_cast<double *>(a.d_out) + (size_t)(r + 1) * plane + p,
                           make_double2(v[0], v[1]));
            continue;
        }
        if (!valid)
            continue;
        for (int h = 0; h < 2; ++h) {
            if (a.clip)
                v[h] = fmin(fmax(v[h], a.clip_lo), a.clip_hi);
            if (!inside[h])
                v[h] = 0.0; // reset_reference_pixels (exposure.py:122-131)
            v[h] = v[h] + zc[h]; // add_zero_read (exposure.py:94-104)
        }
        if (a.add_read_noise) {
            double z[2];
            if (a.d_draw_rn) {
                load_pair(a.d_draw_rn + (size_t)(r + 1) * plane, p, z);
            } else {
                const uint4 q = philox4x32_10(
                    make_uint4(0, (uint32_t)p, (uint32_t)(r + 1), WB_STREAM_READ), a.key0, a.key1);
                if (FAST)
                    box_muller_fd(q.x, q.y, z[0], z[1]);
                else
                    box_muller_d(q.x, q.y, z[0], z[1]);
            }
            v[0] = v[0] + a.read_noise * z[0]; // detector.py:198
            v[1] = v[1] + a.read_noise * z[1];
        }
        if (a.out_f32)
            *reinterpret_cast<float2 *>(reinterpret_cast<float *>(a.d_out) + (size_t)(r + 1) * plane + p) =
                make_float2((float)v[0], (float)v[1]);
        else
            st_stream2(reinterpret_cast<double *>(a.d_out) + (size_t)(r + 1) * plane + p,
                       make_double2(v[0], v[1]));
    }

    if (PASS == 1) {
        __syncthreads();
        if (a.add_nonlinear && threadIdx.x < 16 && threadIdx.x < R && s_iter[threadIdx.x])
            atomicMax(&a.d_newton_iters[threadIdx.x], s_iter[threadIdx.x]);
        return;
    }
    if (!valid)
        return;
    // the zero read itself: clipped, border-zeroed, then read noise
    double v[2] = {zc[0], zc[1]};
    if (a.add_read_noise) {
        double z[2];
        if (a.d_draw_rn) {
            load_pair(a.d_draw_rn, p, z);
        } else {
            const uint4 q =
                philox4x32_10(make_uint4(0, (uint32_t)p, 0u, WB_STREAM_READ), a.key0, a.key1);
            if (FAST)
                box_muller_fd(q.x, q.y, z[0], z[1]);
            else
                box_muller_d(q.x, q.y, z[0], z[1]);
        }
        v[0] = v[0] + a.read_noise * z[0];
        v[1] = v[1] + a.read_noise * z[1];
    }
    if (a.out_f32)
        *reinterpret_cast<float2 *>(reinterpret_cast<float *>(a.d_out) + p) =
            make_float2((float)v[0], (float)v[1]);
    else
        st_stream2(reinterpret_cast<double *>(a.d_out) + p, make_double2(v[0], v[1]));
}

// cosmic_rays.py:70-86: hits -> per-pixel chains (order inside a chain is
// irrelevant: energies are integers and sum exactly in fp64).
__global__ void k_cosmic_chains(int n, const int *__restrict__ pixel, int n_pixels, int *head,
                                int *next)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int px = pixel[i];
    if ((unsigned)px >= (unsigned)n_pixels) {
        next[i] = -1;
        return;
    }
    next[i] = atomicExch(&head[px], i);
}

} // namespace wb
