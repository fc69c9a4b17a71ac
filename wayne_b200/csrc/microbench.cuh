// microbench.cuh -- measured denominators for the photon stage's roofline:
// shared-memory atomic rate (conflict-free / same-address / PSF-like),
// global red rate, Philox and Box-Muller issue rates.  Run through
// wb200_microbench(); numbers are recorded in profiles/.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "photons.cuh"

namespace wb {

constexpr int MB_TILE = 8192; // ints of shared memory per CTA

// which: 0 same address per warp, 1 conflict-free (lane -> own bank),
//        2 PSF-like (3x3 neighbourhood per warp, pseudo-random), 3 random over the tile
__global__ void __launch_bounds__(256) k_mb_smem_atomic(int which, int iters, int *sink)
{
    __shared__ int tile[MB_TILE];
    for (int i = threadIdx.x; i < MB_TILE; i += blockDim.x)
        tile[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t h = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        int idx;
        if (which == 0)
            idx = (warp * 131 + it) & (MB_TILE - 1);
        else if (which == 1)
            idx = ((warp * 64 + (it & 63)) * 32 + lane) & (MB_TILE - 1);
        else if (which == 2) {
            const int dx = (h >> 28) % 3, dy = (h >> 20) % 3;
            idx = ((warp * 5 + dy) * 128 + ((it >> 4) & 63) + dx) & (MB_TILE - 1);
        } else
            idx = (h >> 12) & (MB_TILE - 1);
        atomicAdd(&tile[idx], 1);
    }
    __syncthreads();
    int acc = 0;
    for (int i = threadIdx.x; i < MB_TILE; i += blockDim.x)
        acc += tile[i];
    if (acc == -1)
        sink[0] = acc;
}

// global red.add.s32 over a 64 MB region (spread) or a 256 KB window (hot)
__global__ void __launch_bounds__(256) k_mb_global_red(int which, int iters, int *buf, int n)
{
    uint32_t h = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 99u;
    const int mask = (which == 4) ? (n - 1) : (65536 - 1);
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        atomicAdd(&buf[(h >> 6) & mask], 1);
    }
}

// Issue-rate probes for single instruction kinds (which = 8..12): 64 dependent-free
// instructions per iteration per thread, 8 independent chains
template <int which>
__global__ void __launch_bounds__(256) k_mb_pipe(int iters, int *sink)
{
    uint32_t x[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
        f[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (which == 8) {        // IMAD.WIDE.U32: hi ^ lo feeds the next
                    uint32_t lo, hi;
                    asm volatile("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
                                 : "=r"(lo), "=r"(hi) : "r"(x[i]), "r"(0xD2511F53u));
                    x[i] = hi + lo;
                } else if (which == 9) { // IMAD (32-bit)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(0xD2511F53u), "r"(0x9E3779B9u));
                } else if (which == 10) { // LOP3
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(0x9E3779B9u));
                } else if (which == 11) { // MUFU.LG2
                    asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                } else if (which == 12) { // FFMA (3-register form)
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7]));
                } else {                 // 13: I2FP.F32.U32 + back (F2I excluded: xor with bits)
                    float t;
                    asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(t) : "r"(x[i]));
                    x[i] ^= __float_as_uint(t);
                }
            }
        }
    }
    uint32_t acc = 0;
    float accf = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc ^= x[i];
        accf += f[i];
    }
    if (acc == 0x7fffffffu || accf == 1.2345f)
        sink[0] = 1;
}

// The thrower's random recipe and nothing else (no bins, no tile, no atomics): the ceiling
// of GENERATING electron positions on this GPU.
//   which 6: Philox4x32-10 calls only (fixed key, as in the thrower)
//   which 7: one call + four electrons (16-bit radius / angle fields, ftz SFU Box-Muller)
__global__ void __launch_bounds__(256) k_mb_rng(int which, int iters, int *sink)
{
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    float accf = 0.f;
    uint32_t acci = 0;
    const ThrowKeys k = throw_keys(0x1234u, 0x5678u);
    for (int it = 0; it < iters; ++it) {
        const uint4 r = philox4x32_10_throw((uint32_t)it, 7u, t, k);
        if (which == 6) {
            acci ^= r.x ^ r.y ^ r.z ^ r.w;
        } else {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const uint32_t w = word_of(r, h);
                float x, y;
                throw_position(throw_u1(w), w, 1.25f * WB_SQRT_2LN2, 3.0f, 4.0f, x, y);
                accf += x + y;
            }
        }
    }
    if (accf == 1.2345f || acci == 0x7fffffffu)
        sink[0] = 1;
}

// test hook behind wb200_philox_words
__global__ void k_philox_words(int which, uint4 c, uint32_t k0, uint32_t k1, uint32_t *out)
{
    uint4 r;
    ThrowKeys tk = throw_keys(k0, k1);
    if (which == 0)
        r = philox4x32_10(c, k0, k1);
    else
        r = philox4x32_10_throw(c.x, c.y, c.z, tk, c.w);
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
    out[3] = r.w;
    out[4] = tk.hy;
    out[5] = tk.hw;
}

} // namespace wb
