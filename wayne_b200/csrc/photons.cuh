// photons.cuh -- stages 2+3: throw every electron of an exposure and bin it.
//
// Replaces the reference's PSF() (wayne/pyparallel_menu.c:10-113): its
// 16 B/electron normal table, and its SERIAL scatter loop, for all sub-samples
// of an exposure in one launch.
//
// Two kernels
//   k_throw<MODE>        generic / parity kernel: RANDR (the reference's rand_r
//                        streams reproduced per electron, fp64) and HOST (caller's
//                        normal table) -- bit-exact against pyparallel_menu.c --
//                        plus a baseline PHILOX path kept as an A/B reference
//                        (WB200_GENERIC_THROW=1).
//   k_throw_philox<DIRECT>  the native kernel (further down): instruction-tuned
//                        Philox thrower; with DIRECT the flat field and the
//                        accumulation into the read-interval planes are fused
//                        into the tile flush.
//
// Mapping to B200 (both kernels)
//   grid  = (bin chunks, sub-samples): thousands of CTAs over 148 SMs.
//   CTA   = 256 threads, a TH x TW int32 histogram tile in shared memory placed
//           on the footprint of its bin chunk (trace segment +- ~4 sigma_h).
//   warp  = takes 32 consecutive bins; a warp-level prefix sum of their work
//           units turns the ragged "counts[bin] electrons per bin" loop into a
//           dense stream of units (load balance is exact whatever the counts).
//   unit  = PHILOX: one Philox4x32-10 call = up to 4 electrons of one width of one bin
//           RANDR / HOST: one electron (fp64, bit-compatible with the reference)
//   bin   = shared-memory atomic add into the tile; electrons that leave the
//           tile but not the frame go straight to HBM (rare: |z| > ~4).  At the
//           end the tile is flushed with integer atomics -- integer adds
//           commute, so the result is deterministic.
//   HBM traffic per electron is ~0 by construction; the stage is bound by issue
//   slots for RNG + Box-Muller -- the quarter-rate IMAD.WIDE of the Philox rounds first
//   of all (measured: profiles/r01_pipe_rates_s3.txt) -- not by the atomics.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "stage1.cuh"
#include "gather.cuh"

namespace wb {

// ---- glibc rand_r restated for the compat stream ---------------------------
// state <- state*1103515245 + 12345 (mod 2^32); one rand_r() = 3 steps.
// LCG jump tables: after 2^k steps, s -> A[k]*s + C[k].
__constant__ uint32_t c_lcg_A[32];
__constant__ uint32_t c_lcg_C[32];

__device__ __forceinline__ uint32_t lcg_skip(uint32_t s, uint32_t n)
{
    while (n) {
        const int k = __ffs(n) - 1;
        s = c_lcg_A[k] * s + c_lcg_C[k];
        n &= n - 1;
    }
    return s;
}

__device__ __forceinline__ int rand_r_dev(uint32_t &s)
{
    s = s * 1103515245u + 12345u;
    uint32_t out = (s >> 16) & 0x7ffu;
    s = s * 1103515245u + 12345u;
    out = (out << 10) ^ ((s >> 16) & 0x3ffu);
    s = s * 1103515245u + 12345u;
    out = (out << 10) ^ ((s >> 16) & 0x3ffu);
    return (int)out;
}

// first electron of OpenMP thread t of T over ssum electrons (pyparallel_menu.c:47-49)
__device__ __forceinline__ long long chunk_begin(int t, int T, long long ssum)
{
    return (t >= T) ? ssum : (t * ssum) / T;
}

// The reference's normal pair of electron e of a sub-sample (pyparallel_menu.c:52-62).
__device__ inline void randr_normals(long long e, long long ssum, int T, int test, double &zx,
                                     double &zy)
{
    int t = (int)((e * T) / ssum);
    if (t >= T)
        t = T - 1;
    while (t + 1 < T && chunk_begin(t + 1, T, ssum) <= e)
        ++t;
    while (t > 0 && chunk_begin(t, T, ssum) > e)
        --t;
    const long long first = chunk_begin(t, T, ssum);
    uint32_t st = (uint32_t)(25234 + 17 * t + test);
    st = lcg_skip(st, (uint32_t)(6ull * (unsigned long long)(e - first)));
    const int r1 = rand_r_dev(st);
    const int r2 = rand_r_dev(st);
    const double theta = (2. * 3.14159265358979323846) * (double)r1 / 2147483647.0;
    const double R = sqrt(-2. * log((double)r2 / 2147483647.0));
    zx = R * cos(theta);
    zy = R * sin(theta);
}

struct PhotonParams {
    wb200_photon_args a;
    int sample0; // global index of the launch's first sub-sample (Philox counters)
    int n_split; // generic kernel: CTAs sharing one chunk's electrons (grid.z)
    // native kernel: 1-D grid.  Blocks [0, fine_b0) take chunk_bins bins each, `chunks` per
    // sub-sample; blocks from fine_b0 on belong to the LAST sub-samples (from fine_s0) and take
    // fine_chunk bins each, fine_per per sub-sample, so that the grid drains in shorter quanta.
    // (An experiment kept as an A/B switch, WB200_THROW_FINE: it measured slower, see throw_photons;
    // by default fine_s0 = n_samples and every block is a full chunk.)
    int chunks, fine_b0, fine_s0, fine_chunk, fine_per;
    // native kernel, optional: [chunks][2] first / last bin of each full chunk that can hold
    // electrons (k_chunk_spans; lo = -1: scan the chunk instead)
    const int *d_chunk_span;
};

template <int TW, int TH>
struct Tile {
    int x0, y0;         // frame coordinates of tile cell (0,0)
    int lox, hix;       // accepted tile-relative x range [lox, hix)  (frame: 0 < x < nr)
    int loy, hiy;
};

// Frame-accepted electron at absolute pixel (xa, ya) that missed the tile.
__device__ __forceinline__ void to_window(const wb200_photon_args &a, int s_local, int ox, int oy,
                                          int xa, int ya)
{
    const int wx = xa - ox, wy = ya - oy;
    if ((unsigned)wx < (unsigned)a.win_w && (unsigned)wy < (unsigned)a.win_h)
        atomicAdd(&a.d_win[((size_t)s_local * a.win_h + wy) * a.win_w + wx], 1);
    else
        atomicAdd((unsigned long long *)a.d_lost, 1ull);
}

// Random stream of the native thrower: Philox4x32-10 with a FIXED key, so the ten
// round keys are instruction immediates (a run-time key costs ~9 constant loads per
// call: ptxas re-loads kernel parameters inside the loop whatever the register
// budget), over the counter
//     (4 j = first electron of unit j of the bin, hy + sub-sample, bin, hw ^ WB_STREAM_PHOTONS)
// where (hy, hw) is a 64-bit hash of the exposure's (key0, key1) made on the host
// (wb::throw_keys).  Philox is a bijection of the counter for any key, so distinct
// (exposure, sub-sample, bin, unit) give distinct blocks; two exposures could only
// share blocks if their hashed words collided to within the sample / stream range
// (probability ~2^-49 per pair of exposures).
constexpr uint32_t WB_TK0 = 0xA4093822u, WB_TK1 = 0x299F31D0u; // key: hex digits of pi
constexpr uint32_t WB_PHILOX_M0 = 0xD2511F53u, WB_PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t WB_PHILOX_W0 = 0x9E3779B9u, WB_PHILOX_W1 = 0xBB67AE85u;

struct ThrowKeys {
    uint32_t hy, hw;
};

__host__ __device__ inline ThrowKeys throw_keys(uint32_t key0, uint32_t key1)
{
    unsigned long long z = ((unsigned long long)key1 << 32 | key0) + 0x9E3779B97F4A7C15ull; // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ThrowKeys{(uint32_t)z, (uint32_t)(z >> 32)};
}

// rounds 2..10 of Philox4x32-10 under the fixed key (round 1 is done by the caller,
// which caches its bin-dependent half)
// 32 x 32 -> (lo, hi) as two 32-bit registers (one IMAD.WIDE; written in PTX so that
// NVVM does not widen the xors that follow to 64 bits)
__device__ __forceinline__ void mul_wide(uint32_t a, uint32_t m, uint32_t &lo, uint32_t &hi)
{
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(lo), "=r"(hi)
        : "r"(a), "r"(m));
}
template <int R>
__device__ __forceinline__ uint4 philox_round_fixed(uint4 c)
{
    constexpr uint32_t k0 = WB_TK0 + (uint32_t)R * WB_PHILOX_W0, k1 = WB_TK1 + (uint32_t)R * WB_PHILOX_W1;
    uint32_t lo0, hi0, lo1, hi1;
    mul_wide(c.x, WB_PHILOX_M0, lo0, hi0);
    mul_wide(c.z, WB_PHILOX_M1, lo1, hi1);
    return make_uint4((hi1 ^ c.y) ^ k0, lo1, (hi0 ^ c.w) ^ k1, lo0);
}
__device__ __forceinline__ uint4 philox4x32_rounds2to10_fixed(uint4 c)
{
    c = philox_round_fixed<1>(c);
    c = philox_round_fixed<2>(c);
    c = philox_round_fixed<3>(c);
    c = philox_round_fixed<4>(c);
    c = philox_round_fixed<5>(c);
    c = philox_round_fixed<6>(c);
    c = philox_round_fixed<7>(c);
    c = philox_round_fixed<8>(c);
    c = philox_round_fixed<9>(c);
    return c;
}

// the whole call: counter (e0, hy + sub-sample, bin, hw ^ stream), e0 = 4 * unit = the index of the
// unit's first electron in the bin's list (wide electrons first, each width padded to a multiple of 4)
__device__ __forceinline__ uint4 philox4x32_10_throw(uint32_t e0, uint32_t sample, uint32_t bin, ThrowKeys k,
                                                     uint32_t stream = WB_STREAM_PHOTONS)
{
    uint32_t lo0, hi0, lo1, hi1;
    mul_wide(e0, WB_PHILOX_M0, lo0, hi0);
    mul_wide(bin, WB_PHILOX_M1, lo1, hi1);
    return philox4x32_rounds2to10_fixed(make_uint4((hi1 ^ (k.hy + sample)) ^ WB_TK0, lo1,
                                                   (hi0 ^ (k.hw ^ stream)) ^ WB_TK1, lo0));
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// u1 >= 2^-33 is never denormal, so the flush-to-zero forms are exact here and
// save the denormal pre-scaling the non-ftz forms expand to
__device__ __forceinline__ float lg2_approx(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sin_approx(float x)
{
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float cos_approx(float x)
{
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void red_shared_inc(uint32_t addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// floor(v) for |v| < 2^22 without a conversion instruction: adding 1.5*2^23
// with round-toward-minus-infinity leaves floor(v) in the low mantissa bits.
__device__ __forceinline__ int floor_magic(float v)
{
#ifdef WB_FLOOR_F2I
    return __float2int_rd(v); // A/B variant: conversion on the XU pipe
#else
    return __float_as_int(__fadd_rd(v, 12582912.0f)) - 0x4B400000;
#endif
}

// ---- one Philox call = FOUR electrons --------------------------------------------
// IMAD.WIDE.U32 issues once per 4 cycles per SM sub-partition on B200 (measured:
// wb200_microbench(8), profiles/), so a Philox4x32-10 call costs ~80 cycles per warp
// whatever else is in flight: at two electrons per call the multiplier pipe, not
// the SFU or the shared-memory atomics, bounds the thrower.  Each 32-bit word
// therefore feeds one electron's Box-Muller pair:
//     high 16 bits -> radius uniform  u1 = (k + 1/2) 2^-16
//     low  16 bits -> angle           theta = (t - 32768 + 1/2) 2 pi 2^-16
// Radius fields k < 16 (2.4e-4 of the electrons: everything beyond 4.08 sigma) are
// refined with a 32-bit uniform v from a second call (stream WB_STREAM_PHOTON_TAIL,
// same counter): u1 = (k + v) 2^-16, so the tail is continuous out to 8.2 sigma as
// with a 49-bit uniform.  In the bulk the radius lattice is finer than 3e-3 sigma
// (2.6e-5 sigma at the median) and the angle lattice 9.6e-5 rad -- far below a pixel
// for any PSF width, and tests/test_rng_gpu.py checks the binned distributions
// against the exact double-Gaussian cell probabilities.
constexpr uint32_t WB_STREAM_PHOTON_TAIL = 7;
constexpr uint32_t WB_TAIL_WORD = 16u << 16; // words below this have radius field < 16
constexpr float WB_ZMAX_THROW = 8.3f;        // sqrt(2 * 49 ln 2) = 8.24

// (float)(2^23 + 16-bit field): one PRMT, no conversion instruction
__device__ __forceinline__ float hi16_biased(uint32_t w)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632));
}
__device__ __forceinline__ float lo16_biased(uint32_t w)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610));
}
__device__ __forceinline__ float throw_u1(uint32_t w) // (k + 1/2) 2^-16, exact
{
    return fmaf(hi16_biased(w), 1.52587890625e-05f, -127.99999237060547f);
}
__device__ __forceinline__ float throw_u1_tail(uint32_t w, uint32_t t) // (k + v) 2^-16
{
    const float v = fmaf((float)t, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    return ((hi16_biased(w) - 8388608.0f) + v) * 1.52587890625e-05f;
}
// angle in (-pi, pi) up to a constant rotation of ~3e-5 rad (the rounding of the two
// constants; one fused multiply-add, exact product): (t - 32768 + 1/2) 2 pi 2^-16
__device__ __forceinline__ float throw_theta(uint32_t w)
{
    // c = 2 pi / 65536 as float; d = -c * (2^23 + 32768) + c / 2, rounded from double
    return fmaf(lo16_biased(w), 9.58738019107841e-05f, -807.3892822265625f);
}
// sqrt(2 ln 2): the PSF widths are staged pre-multiplied, so the radius is
// sigma' * sqrt(-lg2 u1) with no multiply in between
constexpr float WB_SQRT_2LN2 = 1.1774100225154747f;
// position of one electron thrown from (cx, cy): centre + sigma sqrt(-2 ln u1) (cos, sin)(theta),
// sigma_s = sigma * WB_SQRT_2LN2
__device__ __forceinline__ void throw_position(float u1, uint32_t w, float sigma_s, float cx, float cy,
                                               float &x, float &y)
{
    const float rs = sqrt_approx(-lg2_approx(u1)) * sigma_s;
    const float th = throw_theta(w);
    x = fmaf(cos_approx(th), rs, cx);
    y = fmaf(sin_approx(th), rs, cy);
}
__device__ __forceinline__ uint32_t word_of(const uint4 &r, int i)
{
    return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w));
}

template <int MODE, int TW, int TH>
__global__ void __launch_bounds__(256) k_throw(const PhotonParams p)
{
    const wb200_photon_args &a = p.a;
    extern __shared__ int tile[]; // TH*TW
    __shared__ float s_red[4][8];
    __shared__ int s_org[2];

    const int s_local = blockIdx.y;
    const int s_glob = p.sample0 + s_local;
    const int W = a.n_bins;
    const int w0 = blockIdx.x * a.chunk_bins;
    const int w1 = min(W, w0 + a.chunk_bins);
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    const long long ssum = a.d_totals ? (long long)a.d_totals[s_local] : 0;
    if (ssum == 0 && a.d_totals)
        return; // nothing to throw in this sub-sample (uniform over the CTA)

    TraceCoef tc = {}; // (read only when the positions come from the trace, but never indeterminate)
    if (!a.d_xpos)
        tc = load_trace(a.d_trace + (size_t)s_local * WB200_TRACE_STRIDE);
    const size_t row = (size_t)s_local * W;

    auto bin_xy = [&](int w, double &x, double &y) {
        if (a.d_xpos) {
            x = a.d_xpos[row + w];
            y = a.d_ypos[row + w];
        } else {
            trace_xy(tc, a.d_wl[w], a.sub_scale, x, y);
        }
    };

    // ---- place the tile on the chunk's footprint -------------------------
    float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
    for (int w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
        if (a.d_counts[row + w] <= 0)
            continue;
        double x, y;
        bin_xy(w, x, y);
        // NaN / inf positions never bin; keep them out of the bounding box
        if (!(fabs(x) < 1e9) || !(fabs(y) < 1e9))
            continue;
        xmin = fminf(xmin, (float)x);
        xmax = fmaxf(xmax, (float)x);
        ymin = fminf(ymin, (float)y);
        ymax = fmaxf(ymax, (float)y);
    }
    xmin = warp_min(xmin);
    xmax = warp_max(xmax);
    ymin = warp_min(ymin);
    ymax = warp_max(ymax);
    if (lane == 0) {
        s_red[0][warp] = xmin;
        s_red[1][warp] = xmax;
        s_red[2][warp] = ymin;
        s_red[3][warp] = ymax;
    }
    for (int i = threadIdx.x; i < TW * TH; i += blockDim.x)
        tile[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < nwarps; ++i) {
            xmin = fminf(xmin, s_red[0][i]);
            xmax = fmaxf(xmax, s_red[1][i]);
            ymin = fminf(ymin, s_red[2][i]);
            ymax = fmaxf(ymax, s_red[3][i]);
        }
        int ox = 0, oy = 0;
        if (xmin <= xmax) {
            ox = (int)floorf(0.5f * (xmin + xmax)) - TW / 2;
            oy = (int)floorf(0.5f * (ymin + ymax)) - TH / 2;
        }
        s_org[0] = ox;
        s_org[1] = oy;
    }
    __syncthreads();
    Tile<TW, TH> T;
    T.x0 = s_org[0];
    T.y0 = s_org[1];
    T.lox = max(0, 1 - T.x0);
    T.hix = min(TW, a.nr - T.x0);
    T.loy = max(0, 1 - T.y0);
    T.hiy = min(TH, a.nc - T.y0);
    const int wox = a.d_win_ox[s_local], woy = a.d_win_oy[s_local];

    // ---- the electrons ------------------------------------------------------
    const int ngroups = (w1 - w0 + 31) >> 5;
    for (int g = warp; g < ngroups; g += nwarps) {
        const int wb = w0 + (g << 5);
        const int w = wb + lane;
        int cnt = 0, nh = 0, off = 0;
        double bx = 0, by = 0, bsl = 0, bsh = 0;
        if (w < w1) {
            cnt = a.d_counts[row + w];
            if (cnt > 0) {
                bin_xy(w, bx, by);
                bsl = a.d_sigl[w];
                bsh = a.d_sigh[w];
                // N = counts*psf_ratio truncated: the first N electrons of the
                // bin take the wide Gaussian (pyparallel_menu.c:89-98)
                nh = __double2int_rz((double)cnt * a.d_ratio[w]);
                if (MODE != WB200_RNG_PHILOX)
                    off = a.d_offsets[row + w];
            } else {
                cnt = 0;
            }
        }
        // PHILOX: a unit is one Philox call = up to four electrons of one width: the
        // ceil(nh/4) wide units of the bin come first, then the narrow ones
        const int nhc = max(0, min(nh, cnt)), uh = (nhc + 3) >> 2;
        const int units = (MODE == WB200_RNG_PHILOX) ? (uh + ((cnt - nhc + 3) >> 2)) : cnt;
        const int incl = warp_incl_scan(units);
        const int total = __shfl_sync(FULL, incl, 31);
        const int excl = incl - units;

        if (MODE == WB200_RNG_PHILOX) {
            // tile-relative fp32 copies of the bin parameters
            const float fx = (float)(bx - (double)T.x0);
            const float fy = (float)(by - (double)T.y0);
            const float fsl = (float)bsl * WB_SQRT_2LN2, fsh = (float)bsh * WB_SQRT_2LN2;
            for (int base = 32 * (int)blockIdx.z; base < total; base += 32 * (int)gridDim.z) {
                const int q = base + lane;
                int b = 0;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int v = __shfl_sync(FULL, incl, b + step - 1);
                    if (v <= q)
                        b += step;
                }
                const int ucnt = __shfl_sync(FULL, cnt, b);
                const int unh = __shfl_sync(FULL, nhc, b);
                const int uuh = __shfl_sync(FULL, uh, b);
                const int uex = __shfl_sync(FULL, excl, b);
                const float ux = __shfl_sync(FULL, fx, b);
                const float uy = __shfl_sync(FULL, fy, b);
                const float usl = __shfl_sync(FULL, fsl, b);
                const float ush = __shfl_sync(FULL, fsh, b);
                if (q >= total)
                    continue;
                const int j = q - uex;
                const ThrowKeys tk = throw_keys(a.key0, a.key1);
                const uint4 r = philox4x32_10_throw((uint32_t)(4 * j), (uint32_t)s_glob, (uint32_t)(wb + b), tk);
                const bool wide = j < uuh;
                const float sg = wide ? ush : usl;
                const int rem = wide ? (unh - 4 * j) : (ucnt - unh - 4 * (j - uuh)); // electrons left
                uint4 t = make_uint4(0, 0, 0, 0);
                if (min(min(r.x, r.y), min(r.z, r.w)) < WB_TAIL_WORD)
                    t = philox4x32_10_throw((uint32_t)(4 * j), (uint32_t)s_glob, (uint32_t)(wb + b), tk,
                                            WB_STREAM_PHOTON_TAIL);
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (h >= rem)
                        break;
                    const uint32_t wd = word_of(r, h);
                    const float u1 = (wd < WB_TAIL_WORD) ? throw_u1_tail(wd, word_of(t, h)) : throw_u1(wd);
                    float px, py;
                    throw_position(u1, wd, sg, ux, uy, px, py);
                    // floor on tile-relative coordinates == the reference's (int)
                    // truncation on frame coordinates for every accepted electron
                    // (x in (-1,1) is rejected either way by the strict 0 < x test)
                    const int ix = __float2int_rd(px);
                    const int iy = __float2int_rd(py);
                    if (ix >= T.lox && ix < T.hix && iy >= T.loy && iy < T.hiy) {
                        atomicAdd(&tile[iy * TW + ix], 1);
                    } else {
                        const int xa = ix + T.x0, ya = iy + T.y0;
                        if (xa > 0 && xa < a.nr && ya > 0 && ya < a.nc)
                            to_window(a, s_local, wox, woy, xa, ya);
                    }
                }
            }
        } else {
            const double *An = nullptr;
            if (MODE == WB200_RNG_HOST)
                An = a.d_normals + a.d_normals_base[s_local];
            const int test = (MODE == WB200_RNG_RANDR) ? a.d_seeds[s_local] : 0;
            // (gridDim.z CTAs share a chunk's electron stream in rows of 32: few bins and many
            // electrons -- one PSF() call -- still fill the machine; integer adds commute)
            for (int base = 32 * (int)blockIdx.z; base < total; base += 32 * (int)gridDim.z) {
                const int q = base + lane;
                int b = 0;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int v = __shfl_sync(FULL, incl, b + step - 1);
                    if (v <= q)
                        b += step;
                }
                const int unh = __shfl_sync(FULL, nh, b);
                const int uex = __shfl_sync(FULL, excl, b);
                const int uoff = __shfl_sync(FULL, off, b);
                const double ux = shfl_d(bx, b);
                const double uy = shfl_d(by, b);
                const double usl = shfl_d(bsl, b);
                const double ush = shfl_d(bsh, b);
                if (q >= total)
                    continue;
                const int j = q - uex;
                const long long e = (long long)uoff + j;
                double zx, zy;
                if (MODE == WB200_RNG_HOST) {
                    zx = An[e];
                    zy = An[e + ssum];
                } else {
                    randr_normals(e, ssum, a.threads, test, zx, zy);
                }
                const double sg = (j < unh) ? ush : usl;
                // xpos = (int)(A*sigma + x_pos): separate multiply and add, C
                // truncation (pyparallel_menu.c:91-92); out-of-range values
                // saturate and fail the bounds test like the reference's INT_MIN
                const int xa = __double2int_rz(__dadd_rn(__dmul_rn(zx, sg), ux));
                const int ya = __double2int_rz(__dadd_rn(__dmul_rn(zy, sg), uy));
                if (xa > 0 && xa < a.nr && ya > 0 && ya < a.nc) {
                    const int ix = xa - T.x0, iy = ya - T.y0;
                    if ((unsigned)ix < (unsigned)TW && (unsigned)iy < (unsigned)TH)
                        atomicAdd(&tile[iy * TW + ix], 1);
                    else
                        to_window(a, s_local, wox, woy, xa, ya);
                }
            }
        }
    }
    __syncthreads();

    // ---- flush the tile into the sub-sample's HBM window ----------------------
    for (int i = threadIdx.x; i < TW * TH; i += blockDim.x) {
        const int v = tile[i];
        if (v) {
            const int iy = i / TW, ix = i - iy * TW;
            const int wx = ix + T.x0 - wox, wy = iy + T.y0 - woy;
            if ((unsigned)wx < (unsigned)a.win_w && (unsigned)wy < (unsigned)a.win_h)
                atomicAdd(&a.d_win[((size_t)s_local * a.win_h + wy) * a.win_w + wx], v);
            else
                atomicAdd((unsigned long long *)a.d_lost, (unsigned long long)v);
        }
    }
}


// ---------------------------------------------------------------------------
// Native (Philox) electron thrower, written for instruction count: the
// baseline generic kernel above spent 111 warp-instructions per 32 electrons
// (profiles/r01_k_throw_opcode_histogram_baseline.txt), this one 64
// (profiles/r01_k_throw_opcodes_s3.txt).  Here
//   * a warp takes its 32-bin groups from a shared counter; a lane owns a contiguous
//     RUN of units of the group: one binary search per lane and group, then a
//     sequential walk that only steps to the next bin when its units are exhausted
//     (bin parameters live in shared memory, 32 B per bin) -- no warp collectives
//     in the hot loop;
//   * the Philox key is fixed (round keys are immediates), the bin's half of round 1
//     is cached, and one call feeds FOUR electrons (see "one Philox call" above);
//   * sqrt.approx / lg2.approx / sin.approx / cos.approx (4 MUFU per electron),
//     PRMT field extraction (no I2F), floor by the round-down magic-number add
//     (FADD.RM, no F2I on the XU pipe);
//   * tile-relative unsigned bounds tests and branch-free shared increments.
// Same counters as the generic kernel's PHILOX path -- (first electron of the unit, sub-sample, bin,
// stream) -- so a given key throws the same electrons whatever the launch geometry.
// ---------------------------------------------------------------------------
// keeps a CTA-uniform value in a vector register instead of letting ptxas rebuild it
// from the kernel parameters every iteration
__device__ __forceinline__ uint32_t pin_reg(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// cell address if (ix, iy) lies in the accepted range [0, nx) x [0, ny), else `other`: two chained
// compares and ONE select (the compiler's own lowering selects once per axis)
__device__ __forceinline__ uint32_t select_in_range(uint32_t ix, uint32_t nx, uint32_t iy, uint32_t ny,
                                                    uint32_t cell, uint32_t other)
{
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\tsetp.lt.and.u32 p, %3, %4, p;\n\tselp.b32 %0, %5, %6, p;\n\t}"
        : "=r"(r)
        : "r"(ix), "r"(nx), "r"(iy), "r"(ny), "r"(cell), "r"(other));
    return r;
}

struct BinPar {        // 32 bytes, two 128-bit shared loads per trip of the electron loop
    uint32_t r1x, r1y; // the bin's half of Philox round 1: hi(M1*bin) ^ c.y ^ k0, lo(M1*bin)
    int nh;            // wide electrons: units j with 4j < nh are wide and hold min(4, nh - 4j)
    int nlx;           // nl + 4*ceil(nh/4): the other units hold min(4, nlx - 4j); units = ceil(nlx/4)
    float fx, fy, sl, sh;
};

// DIRECT = false: tiles are flushed (integer red) into the per-sub-sample HBM
//                  windows and k_gather applies the flat afterwards.
// DIRECT = true : no windows and no gather pass.  At flush each non-zero tile
//                  cell evaluates ITS sub-sample's flat once (grism.py:359-385,
//                  fast form) and adds count*flat to the read interval's plane
//                  as a 2^-24 fixed-point 64-bit integer: integer atomics
//                  commute, so the planes stay bit-reproducible whatever the
//                  launch geometry, and HBM never sees a per-sub-sample buffer.
constexpr double WB_ACC_SCALE = 16777216.0; // 2^24 per electron

struct DirectSample {
    GatherSample g;
    long long *acc; // plane of this sub-sample's read interval
    double inv_range;
};

__device__ __forceinline__ void deposit(const wb200_gather_args &ga, const DirectSample &d, int xa,
                                        int ya, int count)
{
    double v = (double)count;
    if (ga.add_flat) {
        int Xi = xa + ga.flat_off, Yi = ya + ga.flat_off;
        if (Xi < 0)
            Xi += ga.flat_n;
        if (Yi < 0)
            Yi += ga.flat_n;
        if ((unsigned)Xi < (unsigned)ga.flat_n && (unsigned)Yi < (unsigned)ga.flat_n) {
            const size_t fi = (size_t)Yi * ga.flat_n + Xi;
            double f[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                f[i] = ga.flat_planes_f32 ? (double)reinterpret_cast<const float *>(ga.d_flat[i])[fi] : ga.d_flat[i][fi];
            double fv = flat_value_fast(d.g, (double)Xi, (double)Yi, f[0], f[1], f[2], f[3], ga.flat_wmin, d.inv_range);
            if (ga.flat_f32)
                fv = (double)__double2float_rn(fv);
            v *= fv;
        }
    }
    WB_DEV_ASSERT((unsigned)(ya + ga.border) < (unsigned)ga.F && (unsigned)(xa + ga.border) < (unsigned)ga.F);
    const long long q = __double2ll_rn(v * WB_ACC_SCALE);
    // (the plane pointer comes out of shared memory: say that it is a global address, or the
    // compiler emits a generic atomic with a shared-memory CAS path)
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(d.acc + (size_t)(ya + ga.border) * ga.F + (xa + ga.border)),
                 "l"(q)
                 : "memory");
}

#ifndef WB_THROW_MIN_BLOCKS
#define WB_THROW_MIN_BLOCKS 5
#endif

// (float)(2^23 + 16-bit field) with the bias word in a REGISTER and the selector an immediate: as
// written by __byte_perm the selector travels in a register, and ptxas made a fresh copy of it for
// each of the eight extractions of a trip
__device__ __forceinline__ float hi16_biased_r(uint32_t w, uint32_t bias)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(w), "r"(bias));
    return __uint_as_float(r);
}
__device__ __forceinline__ float lo16_biased_r(uint32_t w, uint32_t bias)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7610;" : "=r"(r) : "r"(w), "r"(bias));
    return __uint_as_float(r);
}

// `cell` if |px| < hx and |py| < hy, else `other`: two chained compares on the coordinates
// themselves and ONE select (the compiler's own lowering of the && selects once per axis)
__device__ __forceinline__ uint32_t select_in_box(float px, float hx, float py, float hy, uint32_t cell,
                                                  uint32_t other)
{
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\t.reg .f32 ax, ay;\n\tabs.ftz.f32 ax, %1;\n\tabs.ftz.f32 ay, %3;\n\t"
        "setp.lt.ftz.f32 p, ax, %2;\n\tsetp.lt.and.ftz.f32 p, ay, %4, p;\n\tselp.b32 %0, %5, %6, p;\n\t}"
        : "=r"(r)
        : "f"(px), "f"(hx), "f"(py), "f"(hy), "r"(cell), "r"(other));
    return r;
}

// the same with a third chained compare: `cell` only if also a > b (the electron exists)
__device__ __forceinline__ uint32_t select_in_box_if_gt(float px, float hx, float py, float hy, int a, int b,
                                                        uint32_t cell, uint32_t other)
{
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\t.reg .f32 ax, ay;\n\tabs.ftz.f32 ax, %1;\n\tabs.ftz.f32 ay, %3;\n\t"
        "setp.lt.ftz.f32 p, ax, %2;\n\tsetp.lt.and.ftz.f32 p, ay, %4, p;\n\tsetp.gt.and.s32 p, %7, %8, p;\n\t"
        "selp.b32 %0, %5, %6, p;\n\t}"
        : "=r"(r)
        : "f"(px), "f"(hx), "f"(py), "f"(hy), "r"(cell), "r"(other), "r"(a), "r"(b));
    return r;
}

// CTA-uniform values of the native thrower that only the rare paths read: shared memory
struct ThrowShared {
    DirectSample ds;
    int hx, hy, ax0, ay0, wox, woy, s_local, pad;
    unsigned tally[2]; // slow-path electrons binned inside the frame / dropped outside it
    uint32_t bias;     // 0x4B000000, read back from here so that it lives in a REGISTER (see hi16_biased_r)
    uint32_t pad2;
    wb200_gather_args ga;  // copies of the kernel parameters for the replay (a separate function:
    wb200_photon_args a;   // parameters taken by reference would be spilled to the stack)
};

// The electrons of one lane's run of units.
//   REPLAY = false  the hot loop: every electron of every unit is positioned; one inside the
//                   accepted part of the tile (fast float test) increments its cell, one outside it
//                   the warp's spare word `dump`, and so does one that does not exist (the warp
//                   knows how many of those its group holds).
//                   No branch depends on where an electron went.
//   REPLAY = true   run again by a warp that found its spare word non-zero (an electron in ~4e6
//                   leaves a tile that is clear of the frame's edges): the same units, the same
//                   arithmetic, and now ONLY the electrons that failed the fast test are looked
//                   at -- frame test of the reference (pyparallel_menu.c:93), then straight to HBM.
template <int TW, int TH, bool DIRECT, bool REPLAY>
__device__ __forceinline__ void throw_run(const BinPar *pb, int j, int n, uint32_t cwk, uint32_t cellk,
                                          uint32_t dump, float hxf, float hyf, ThrowShared &sh)
{
    constexpr int ROW_SHIFT = (TW == 64 ? 8 : TW == 128 ? 9 : TW == 256 ? 10 : 11);
    static_assert((4 * TW) == (1 << ROW_SHIFT), "ROW_SHIFT = log2(4 TW)");
    const uint32_t bias = sh.bias; // (a value ptxas cannot see: it would make it the immediate again)
    // The walk counts ELECTRONS, not units: j4 = 4 * (unit index) is the index of the unit's first
    // electron within its bin's padded list, and it is also the first counter word of the unit's
    // Philox call -- "this unit is used up" is then j4 >= nlx, the electrons left are n - j4, and
    // no shift, no ceil(nlx / 4) and no separate unit counter are formed per trip.
    int j4 = 4 * j;
    static_assert(sizeof(BinPar) == 32, "the walk steps by 32 bytes");
    uint32_t pba = (uint32_t)__cvta_generic_to_shared(pb);
    int nlx_cur = 0x7fffffff; // padded electron count of the bin pb points at (known after the first load)
    do {
        // step to the next staged bin when this one is used up, then (re)load the bin:
        // branch-free -- at four electrons per unit some lane steps on most trips
        // (one compare and two predicated instructions; as C++ selects ptxas spent five on it)
        asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, %2;\n\t@p add.u32 %0, %0, 32;\n\t@p mov.b32 %1, 0;\n\t}"
            : "+r"(pba), "+r"(j4)
            : "r"(nlx_cur));
        BinPar cur;
        asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
            : "=r"(cur.r1x), "=r"(cur.r1y), "=r"(cur.nh), "=r"(cur.nlx)
            : "r"(pba));
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];"
            : "=f"(cur.fx), "=f"(cur.fy), "=f"(cur.sl), "=f"(cur.sh)
            : "r"(pba));
        nlx_cur = cur.nlx;
        const uint32_t r1x = cur.r1x, r1y = cur.r1y;
        uint32_t pjl, pjh;
        mul_wide((uint32_t)j4, WB_PHILOX_M0, pjl, pjh);
        const uint4 r = philox4x32_rounds2to10_fixed(make_uint4(r1x, r1y, pjh ^ cwk, pjl));
        // the unit's width and how many of its four electrons exist
        const bool wide = j4 < cur.nh;
        const float sg = wide ? cur.sh : cur.sl;
        const int rem = (wide ? cur.nh : cur.nlx) - j4;
        // radius uniforms (k + 1/2) 2^-16; fields k < 16 (everything beyond 4.08 sigma, 1e-3 of the
        // units) are refined with a 32-bit uniform from a second call
        float u1[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
            u1[h] = fmaf(hi16_biased_r(word_of(r, h), bias), 1.52587890625e-05f, -127.99999237060547f);
        if (min(min(r.x, r.y), min(r.z, r.w)) < WB_TAIL_WORD) {
            const uint4 t = philox4x32_rounds2to10_fixed(
                make_uint4(r1x, r1y, pjh ^ cwk ^ (WB_STREAM_PHOTONS ^ WB_STREAM_PHOTON_TAIL), pjl));
#pragma unroll
            for (int h = 0; h < 4; ++h)
                if (word_of(r, h) < WB_TAIL_WORD)
                    u1[h] = throw_u1_tail(word_of(r, h), word_of(t, h));
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t w = word_of(r, h);
            const float rs = sqrt_approx(-lg2_approx(u1[h])) * sg;
            const float th = fmaf(lo16_biased_r(w, bias), 9.58738019107841e-05f, -807.3892822265625f);
            const float px = fmaf(cos_approx(th), rs, cur.fx);
            const float py = fmaf(sin_approx(th), rs, cur.fy);
            // bits(p + 1.5 2^23, rounded down) = 0x4B400000 + floor(p)
            const uint32_t bx = __float_as_uint(__fadd_rd(px, 12582912.0f));
            const uint32_t by = __float_as_uint(__fadd_rd(py, 12582912.0f));
            if (!REPLAY) {
                // one select: the cell if the electron exists AND lies in the accepted box, else the
                // warp's spare word.  Electrons that do not exist (the last unit of a width holds 1..4)
                // are counted there too; the warp knows how many its group has (`expect`), so only a
                // count ABOVE that means that some electron left the box.
                // (a predicated increment was tried: ATOMS.POPC.INC cannot be predicated -- ptxas
                // branches around it -- and an ATOMS.ADD of 0 / 1 loses the POPC form)
                const uint32_t cell = cellk + (bx << 2) + (by << ROW_SHIFT);
                const uint32_t ad = (h == 0) ? select_in_box(px, hxf, py, hyf, cell, dump)
                                             : select_in_box_if_gt(px, hxf, py, hyf, rem, h, cell, dump);
                WB_DEV_ASSERT(ad >= dump - (uint32_t)(8 * 8 + TW * TH * 4) && ad <= dump && (ad & 3) == 0);
                red_shared_inc(ad);
            } else {
                if ((h > 0 && rem <= h) || (fabsf(px) < hxf && fabsf(py) < hyf))
                    continue; // does not exist / was binned by the hot loop
                const int xa = (int)(bx - 0x4B400000u) + sh.ax0, ya = (int)(by - 0x4B400000u) + sh.ay0;
                if (xa > 0 && xa < sh.a.nr && ya > 0 && ya < sh.a.nc) {
                    if (DIRECT) {
                        deposit(sh.ga, sh.ds, xa, ya, 1);
                        atomicAdd(&sh.tally[0], 1u);
                    } else
                        to_window(sh.a, sh.s_local, sh.wox, sh.woy, xa, ya);
                } else if (DIRECT)
                    atomicAdd(&sh.tally[1], 1u); // dropped like the reference's (pyparallel_menu.c:93)
            }
        }
        j4 += 4;
    } while (--n > 0);
}

template <int TW, int TH, bool DIRECT>
__device__ __noinline__ void throw_replay(const BinPar *pb, int j, int n, uint32_t cwk, float hxf, float hyf,
                                          ThrowShared &sh)
{
    throw_run<TW, TH, DIRECT, true>(pb, j, n, cwk, 0u, 0u, hxf, hyf, sh);
}

template <int TW, int TH, bool DIRECT>
__global__ void __launch_bounds__(256, WB_THROW_MIN_BLOCKS)
k_throw_philox(const PhotonParams p, const ThrowKeys keys, const wb200_gather_args ga)
{
    const wb200_photon_args &a = p.a;
    // TH*TW tile cells, then per warp two spare words: [0] electrons outside the accepted part of
    // the tile (a non-zero count makes the warp replay its group), [1] electrons that do not exist
    extern __shared__ int tile[];
    __shared__ float s_red[4][8];
    __shared__ int s_org[2];
    __shared__ __align__(16) BinPar s_bin[8][32];
    __shared__ ThrowShared sh;
    __shared__ double s_pos[4]; // x = wl * [0] + [1], y = x * [2] + [3] (frame coordinates of a bin)

    const int W = a.n_bins;
    int s_local, w0, w1;
    if ((int)blockIdx.x < p.fine_b0) {
        s_local = (int)blockIdx.x / p.chunks;
        w0 = ((int)blockIdx.x - s_local * p.chunks) * a.chunk_bins;
        w1 = min(W, w0 + a.chunk_bins);
    } else {
        const int b = (int)blockIdx.x - p.fine_b0, ds = b / p.fine_per;
        s_local = p.fine_s0 + ds;
        w0 = (b - ds * p.fine_per) * p.fine_chunk;
        w1 = min(W, w0 + p.fine_chunk);
    }
    const uint32_t s_glob = (uint32_t)(p.sample0 + s_local);
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    if (a.d_totals && a.d_totals[s_local] == 0)
        return;
    if (DIRECT && ga.d_read_end[ga.n_reads - 1] < (int)s_glob)
        return; // a sub-sample after the last read belongs to no read (exposure_generator.py:361)

    // Bin positions (_SpectrumTrace.wl_to_x / wl_to_y, grism.py:635-669, minus the sub-array shift,
    // exposure_generator.py:630-632) with the division hoisted out of the per-bin work:
    //     x = (wl - c_wl) / m_wl - sub = wl (1 / m_wl) - (c_wl / m_wl + sub)
    //     y = m_t (x + sub - x_ref) + c_t + y_ref - sub = m_t x + (m_t (sub - x_ref) + c_t + y_ref - sub)
    // two fp64 FMAs per bin (1e-13 px from trace_xy's operation order).
    if (threadIdx.x == 0 && !a.d_xpos) {
        const TraceCoef tc = load_trace(a.d_trace + (size_t)s_local * WB200_TRACE_STRIDE);
        s_pos[0] = 1.0 / tc.m_wl;
        s_pos[1] = -(tc.c_wl / tc.m_wl + a.sub_scale);
        s_pos[2] = tc.m_t;
        s_pos[3] = tc.m_t * (a.sub_scale - tc.x_ref) + tc.c_t + tc.y_ref - a.sub_scale;
    }
    const size_t row = (size_t)s_local * W;
    auto bin_xy = [&](int w, double &x, double &y) {
        if (a.d_xpos) {
            x = a.d_xpos[row + w];
            y = a.d_ypos[row + w];
        } else {
            x = fma(a.d_wl[w], s_pos[0], s_pos[1]);
            y = fma(x, s_pos[2], s_pos[3]);
        }
    };

    // ---- place the tile on the chunk's footprint ----------------------------------------
    // Bin positions are linear in the wavelength, so on a monotonic grid the footprint is spanned
    // by the chunk's first and last populated bins, which stage 1 has found once per exposure
    // (k_chunk_spans): one thread places the tile from those two positions while the others zero
    // it.  Otherwise (explicit positions, a non-monotonic grid, the fine-grained tail) the chunk's
    // counts and positions are scanned as in the generic kernel.
    int span_lo = -1, span_hi = -1;
    if (p.d_chunk_span && !a.d_xpos && (int)blockIdx.x < p.fine_b0) {
        const int c = (int)blockIdx.x - s_local * p.chunks;
        span_lo = p.d_chunk_span[2 * c];
        span_hi = p.d_chunk_span[2 * c + 1];
    }
    if (span_lo >= 0) {
        // (thread 0 reads back the position constants it has just written itself: no barrier yet)
        if (threadIdx.x == 0) {
            double x0, y0, x1, y1;
            bin_xy(span_lo, x0, y0);
            bin_xy(span_hi, x1, y1);
            int ox = 0, oy = 0;
            if (fabs(x0) < 1e9 && fabs(y0) < 1e9 && fabs(x1) < 1e9 && fabs(y1) < 1e9) {
                ox = (int)floorf(0.5f * ((float)x0 + (float)x1)) - TW / 2;
                oy = (int)floorf(0.5f * ((float)y0 + (float)y1)) - TH / 2;
            }
            s_org[0] = ox;
            s_org[1] = oy;
        }
        for (int i = threadIdx.x; i < TW * TH + 16; i += blockDim.x)
            tile[i] = 0;
    } else {
        __syncthreads(); // the position constants
        float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
        for (int w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
            if (a.d_counts[row + w] <= 0)
                continue;
            double x, y;
            bin_xy(w, x, y);
            if (!(fabs(x) < 1e9) || !(fabs(y) < 1e9))
                continue;
            xmin = fminf(xmin, (float)x);
            xmax = fmaxf(xmax, (float)x);
            ymin = fminf(ymin, (float)y);
            ymax = fmaxf(ymax, (float)y);
        }
        xmin = warp_min(xmin);
        xmax = warp_max(xmax);
        ymin = warp_min(ymin);
        ymax = warp_max(ymax);
        if (lane == 0) {
            s_red[0][warp] = xmin;
            s_red[1][warp] = xmax;
            s_red[2][warp] = ymin;
            s_red[3][warp] = ymax;
        }
        for (int i = threadIdx.x; i < TW * TH + 16; i += blockDim.x)
            tile[i] = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < nwarps; ++i) {
                xmin = fminf(xmin, s_red[0][i]);
                xmax = fmaxf(xmax, s_red[1][i]);
                ymin = fminf(ymin, s_red[2][i]);
                ymax = fmaxf(ymax, s_red[3][i]);
            }
            int ox = 0, oy = 0;
            if (xmin <= xmax) {
                ox = (int)floorf(0.5f * (xmin + xmax)) - TW / 2;
                oy = (int)floorf(0.5f * (ymin + ymax)) - TH / 2;
            }
            s_org[0] = ox;
            s_org[1] = oy;
        }
        __syncthreads();
    }
    // accepted tile-relative range [lo, lo+n): frame test 0 < x < nr, 0 < y < nc, trimmed to an even
    // size and CENTRED: with hx = n/2 the fast test of an electron at centred coordinate p is
    // |p| < hx -- one float compare per axis on the coordinate itself -- and everything that fails
    // it (the trimmed last column included) takes the exact slow path
    auto accepted = [&](int tx0, int ty0, int &lox, int &loy, int &hx, int &hy) {
        lox = max(0, 1 - tx0);
        loy = max(0, 1 - ty0);
        hx = max(0, min(TW, a.nr - tx0) - lox) >> 1;
        hy = max(0, min(TH, a.nc - ty0) - loy) >> 1;
    };
    // the sub-sample's flat-field / accumulation constants are CTA-uniform and only used by the flush
    // and by electrons that leave the tile: shared memory, not registers.  Thread 0 fills them in
    // right behind the tile origin (which it wrote itself), so that ONE barrier publishes the
    // position constants, the origin, these and the zeroed tile.
    __shared__ int s_next; // next 32-bin group to hand out (dynamic: counts are ragged)
    if (threadIdx.x == 0) {
        int lox, loy, hx, hy;
        accepted(s_org[0], s_org[1], lox, loy, hx, hy);
        s_next = 0;
        sh.hx = hx;
        sh.hy = hy;
        sh.ax0 = s_org[0] + lox + hx;
        sh.ay0 = s_org[1] + loy + hy;
        sh.wox = sh.woy = 0;
        sh.s_local = s_local;
        sh.tally[0] = sh.tally[1] = 0u;
        sh.bias = 0x4B000000u;
        sh.ga = ga;
        sh.a = a;
        if (!DIRECT) {
            sh.wox = a.d_win_ox[s_local];
            sh.woy = a.d_win_oy[s_local];
        } else {
            const double *t = a.d_trace + (size_t)s_local * WB200_TRACE_STRIDE;
            DirectSample &ds = sh.ds;
            ds.g.ox = ds.g.oy = ds.g.s = ds.g.pad = 0;
            ds.g.x_ref = t[0];
            ds.g.y_ref = t[1];
            ds.g.a_t_i = 1 / t[2];
            ds.g.den = 1.0 / sqrt(ds.g.a_t_i * ds.g.a_t_i + 1);
            ds.g.m_w = t[4];
            ds.g.c_w = t[5];
            ds.inv_range = 1.0 / (ga.flat_wmax - ga.flat_wmin);
            int r = 0; // read interval of this sub-sample: first r with read_end[r] >= s
            while (r + 1 < ga.n_reads && ga.d_read_end[r] < (int)s_glob)
                ++r;
            ds.acc = reinterpret_cast<long long *>(ga.d_acc) + (size_t)r * ga.F * ga.F;
        }
    }
    __syncthreads();
    const int tx0 = s_org[0], ty0 = s_org[1];
    int lox, loy, hx, hy;
    accepted(tx0, ty0, lox, loy, hx, hy);
    // Bin positions are staged relative to the CENTRE of the accepted range, frame pixel
    // (ax0, ay0).  floor() of a centred coordinate p comes from the round-down magic add:
    // bits(p + 1.5 2^23) = 0x4B400000 + floor(p), and the tile cell of (floor px, floor py) is
    //     tile + ((fy + loy + hy) TW + fx + lox + hx) 4  =  cellk + (bits_x << 2) + (bits_y << log2(4 TW))
    // with every constant (the two magic offsets too: shifts and adds wrap mod 2^32) folded into
    // cellk -- no subtraction, no separate address arithmetic per electron.
    static_assert((TW & (TW - 1)) == 0, "TW must be a power of two");
    constexpr int ROW_SHIFT = (TW == 64 ? 8 : TW == 128 ? 9 : TW == 256 ? 10 : 11);
    const int ax0 = tx0 + lox + hx, ay0 = ty0 + loy + hy;
    const float hxf = (float)hx, hyf = (float)hy;
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    const uint32_t cellk = pin_reg(tile_s + (uint32_t)(((loy + hy) * TW + lox + hx) * 4) - (0x4B400000u << 2) -
                                   (0x4B400000u << ROW_SHIFT));
    const uint32_t dump = pin_reg(tile_s + (uint32_t)(TW * TH * 4 + warp * 8));
    int *const mydump = tile + TW * TH + warp * 2;
    const int wox = sh.wox, woy = sh.woy;
    BinPar *mybins = s_bin[warp];
    // round-1 constants of the thrower's Philox stream (see ThrowKeys)
    const uint32_t cyk = (keys.hy + s_glob) ^ WB_TK0;
    const uint32_t cwk = pin_reg((keys.hw ^ WB_STREAM_PHOTONS) ^ WB_TK1);

    const int ngroups = (w1 - w0 + 31) >> 5;
    int expect = 0; // electrons this warp has thrown at its spare word because they do not exist
    for (;;) {
        int g = 0;
        if (lane == 0)
            g = atomicAdd(&s_next, 1);
        g = __shfl_sync(FULL, g, 0);
        if (g >= ngroups)
            break;
        const int wb = w0 + (g << 5);
        const int w = wb + lane;
        // (prefetching the counts of the group this warp is likely to take next into L1 was tried:
        // 1.841 against 1.828 ms)
        BinPar bp;
        int cnt = 0, nh = 0;
        bp.fx = bp.fy = bp.sl = bp.sh = 0.f;
        if (w < w1) {
            cnt = a.d_counts[row + w];
            if (cnt > 0) {
                double bx, by;
                bin_xy(w, bx, by);
                // first N = (int)(counts*ratio) electrons take the wide Gaussian
                // (pyparallel_menu.c:89-98)
                nh = __double2int_rz((double)cnt * a.d_ratio[w]);
                bp.fx = (float)(bx - (double)ax0);
                bp.fy = (float)(by - (double)ay0);
                bp.sl = (float)a.d_sigl[w];
                bp.sh = (float)a.d_sigh[w];
                // bins that cannot reach the frame throw nothing: NaN / far-away
                // positions, NaN widths (the reference's INT_MIN path), and widths
                // beyond 1e5 px (keeps |coordinate| < 2^22 for the magic-add floor)
                bool bad = !(fabsf(bp.fx) < 3.0e6f) || !(fabsf(bp.fy) < 3.0e6f);
                if (nh > 0 && !(fabsf(bp.sh) <= 1.0e5f))
                    bad = true;
                if (cnt - nh > 0 && !(fabsf(bp.sl) <= 1.0e5f))
                    bad = true;
                if (bad) {
                    if (DIRECT && a.d_tally) // not thrown, but accounted for: dropped outside the frame
                        atomicAdd((unsigned long long *)a.d_tally + 1, (unsigned long long)cnt);
                    cnt = 0;
                }
                bp.sl *= WB_SQRT_2LN2; // the radius is sigma' sqrt(-lg2 u1): no multiply in between
                bp.sh *= WB_SQRT_2LN2;
            } else {
                cnt = 0;
            }
        }
        nh = max(0, min(nh, cnt));
        bp.nh = nh;
        bp.nlx = cnt - nh + ((nh + 3) & ~3);
        const int units = (bp.nlx + 3) >> 2; // ceil(nh/4) wide ones, then ceil(nl/4) narrow ones
        // electrons of the group's padded lists that do not exist: they are thrown at the spare word
        // (a running total: the spare word is only cleared when a group is replayed)
        expect += (int)__reduce_add_sync(FULL, (unsigned)(((-nh) & 3) + ((-(cnt - nh)) & 3)));
        mul_wide((uint32_t)w, WB_PHILOX_M1, bp.r1y, bp.r1x);
        bp.r1x ^= cyk;
        const int incl = warp_incl_scan(units);
        const int total = __shfl_sync(FULL, incl, 31);
        // bins that have units are staged back to back, so the walk steps with "+1"
        const uint32_t have = __ballot_sync(FULL, units > 0);
        __syncwarp();
        if (units > 0)
            mybins[__popc(have & ((1u << lane) - 1u))] = bp;
        __syncwarp();
        if (total == 0)
            continue;
        // this lane's run of units
        const int K = (total + 31) >> 5;
        const int q = lane * K;
        const int qend = min(q + K, total);
        // first bin of the run: largest b with excl[b] <= q (all lanes search)
        int b = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int v = __shfl_sync(FULL, incl, b + step - 1);
            if (v <= q)
                b += step;
        }
        const int excl_b = __shfl_sync(FULL, incl - units, b);
        const BinPar *pb = mybins + __popc(have & ((1u << b) - 1u));
        // flat walk over the run (all lanes execute the same body every trip; a lane steps to
        // its next bin when the current one has no units left)
        if (q < qend)
            throw_run<TW, TH, DIRECT, false>(pb, q - excl_b, qend - q, cwk, cellk, dump, hxf, hyf, sh);
        __syncwarp();
        if (mydump[0] != expect) { // some electron of the group left the accepted part of the tile
            __syncwarp();
            if (lane == 0)
                mydump[0] = 0;
            expect = 0;
            if (q < qend)
                throw_replay<TW, TH, DIRECT>(pb, q - excl_b, qend - q, cwk, hxf, hyf, sh);
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- flush the tile ---------------------------------------------------------
    // (a queue that compacts the non-zero cells before the fp64 part -- every lane busy whatever the
    // fill -- was measured SLOWER: 1.950 against 1.913 ms; the halo rows are not that sparse)
    unsigned long long binned = 0;
    for (int i = threadIdx.x; i < TW * TH; i += blockDim.x) {
        const int v = tile[i];
        if (v) {
            const int iy = i / TW, ix = i - iy * TW;
            if (DIRECT) {
                deposit(ga, sh.ds, ix + tx0, iy + ty0, v);
                binned += (unsigned long long)(unsigned)v;
            } else {
                const int wx = ix + tx0 - wox, wy = iy + ty0 - woy;
                if ((unsigned)wx < (unsigned)a.win_w && (unsigned)wy < (unsigned)a.win_h)
                    atomicAdd(&a.d_win[((size_t)s_local * a.win_h + wy) * a.win_w + wx], v);
                else
                    atomicAdd((unsigned long long *)a.d_lost, (unsigned long long)v);
            }
        }
    }
    if (DIRECT && a.d_tally) {
        // electron bookkeeping of the exposure: binned + dropped == thrown, exactly
        binned = warp_sum_u64(binned);
        if (lane == 0 && binned)
            atomicAdd((unsigned long long *)a.d_tally, binned);
        if (threadIdx.x < 2 && sh.tally[threadIdx.x]) // (complete: every warp passed the barrier before the flush)
            atomicAdd((unsigned long long *)a.d_tally + threadIdx.x, (unsigned long long)sh.tally[threadIdx.x]);
    }
}

} // namespace wb
