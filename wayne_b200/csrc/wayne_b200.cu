// wayne_b200.cu -- the C ABI of libwayne_b200.so (see include/wayne_b200.h).
// One translation unit: kernels live in the .cuh files next to this one.
//
// Build (see wayne_b200/build.py):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//        -fmad=false -Xcompiler -fPIC -shared -o libwayne_b200.so wayne_b200.cu
// -fmad=false: the fp64 paths must reproduce numpy / C expressions that are
// never fused on the reference's x86-64 build; the fp32 fast path asks for its
// FMAs explicitly (fmaf).
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "gather.cuh"
#include "philox.cuh"
#include "photons.cuh"
#include "microbench.cuh"
#include "reads.cuh"
#include "reads_native.cuh"
#include "stage1.cuh"
#include "counts_native.cuh"

namespace wb {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

static bool g_lcg_ready[64] = {};

// jump tables of the rand_r LCG, uploaded once per device
static int ensure_lcg_tables()
{
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && g_lcg_ready[dev])
        return WB200_OK;
    uint32_t A[32], C[32];
    uint32_t a = 1103515245u, c = 12345u;
    for (int k = 0; k < 32; ++k) {
        A[k] = a;
        C[k] = c;
        c = a * c + c; // two applications: a*(a*s+c)+c
        a = a * a;
    }
    WB_CUDA(cudaMemcpyToSymbol(c_lcg_A, A, sizeof(A)));
    WB_CUDA(cudaMemcpyToSymbol(c_lcg_C, C, sizeof(C)));
    if (dev >= 0 && dev < 64)
        g_lcg_ready[dev] = true;
    return WB200_OK;
}

constexpr int TILE_W = 128, TILE_H = 64;

template <int MODE>
static int launch_throw(const PhotonParams &p, cudaStream_t st)
{
    const wb200_photon_args &a = p.a;
    const int chunks = (a.n_bins + a.chunk_bins - 1) / a.chunk_bins;
    dim3 grid(chunks, a.n_samples, p.n_split > 1 ? p.n_split : 1);
    const size_t smem = (size_t)TILE_W * TILE_H * sizeof(int);
    k_throw<MODE, TILE_W, TILE_H><<<grid, 256, smem, st>>>(p);
    WB_LAUNCHED("k_throw");
    return WB200_OK;
}

int throw_photons(const wb200_photon_args *a, int sample0, cudaStream_t st, const wb200_gather_args *direct,
                  int n_split = 1, const int *d_chunk_span = nullptr)
{
    WB_REQUIRE(a != nullptr, "null args");
    WB_REQUIRE(a->n_samples >= 0 && a->n_bins > 0, "bad sizes");
    WB_REQUIRE(a->chunk_bins > 0 && (a->chunk_bins % 32) == 0, "chunk_bins must be a multiple of 32");
    WB_REQUIRE(a->d_counts, "null counts");
    WB_REQUIRE(direct || (a->d_win && a->d_win_ox && a->d_win_oy && a->d_lost), "null window buffers");
    WB_REQUIRE(a->d_ratio && a->d_sigl && a->d_sigh, "null psf tables");
    WB_REQUIRE((a->d_xpos && a->d_ypos) || (a->d_trace && a->d_wl), "no positions");
    WB_REQUIRE(a->n_samples <= 65535, "at most 65535 sub-samples per launch");
    if (a->n_samples == 0)
        return WB200_OK;
    PhotonParams p;
    p.a = *a;
    p.sample0 = sample0;
    p.n_split = n_split;
    p.chunks = p.fine_b0 = p.fine_s0 = p.fine_chunk = p.fine_per = 0;
    p.d_chunk_span = nullptr;
    WB_REQUIRE(!direct || a->rng_mode == WB200_RNG_PHILOX, "direct accumulation is a native-mode path");
    switch (a->rng_mode) {
    case WB200_RNG_PHILOX: {
        if (getenv("WB200_GENERIC_THROW") && !direct)      // A/B switch: the baseline generic kernel
            return launch_throw<WB200_RNG_PHILOX>(p, st);
        const ThrowKeys keys = throw_keys(a->key0, a->key1);
        const int chunks = (a->n_bins + a->chunk_bins - 1) / a->chunk_bins;
        // Optionally the last ~1.25 waves' worth of sub-samples in chunks 1/fine_div as long (see
        // PhotonParams).  OFF by default: measured on the configs[3] shape it LOSES (1.891 ms with
        // fine_div = 4 against 1.868 ms with equal CTAs) -- the grid does not drain in lock step, so
        // there is no idle last wave to recover, and short CTAs pay tile set-up and flush more often.
        int fine_div = 1;
        if (const char *env = getenv("WB200_THROW_FINE")) // A/B switch
            fine_div = atoi(env);
        p.chunks = chunks;
        p.fine_s0 = a->n_samples;
        p.fine_chunk = a->chunk_bins;
        p.fine_per = chunks;
        if (fine_div > 1 && a->chunk_bins / fine_div >= 32) {
            const int slots = 148 * WB_THROW_MIN_BLOCKS;
            int n_fine = (slots + slots / 4 + chunks - 1) / chunks;
            n_fine = n_fine < a->n_samples / 4 ? n_fine : a->n_samples / 4;
            p.fine_s0 = a->n_samples - n_fine;
            p.fine_chunk = (a->chunk_bins / fine_div + 31) / 32 * 32;
            p.fine_per = (a->n_bins + p.fine_chunk - 1) / p.fine_chunk;
        }
        p.fine_b0 = p.fine_s0 * chunks;
        p.d_chunk_span = d_chunk_span;
        dim3 grid(p.fine_b0 + (a->n_samples - p.fine_s0) * p.fine_per);
        const size_t smem = (size_t)TILE_W * TILE_H * sizeof(int) + 64; // + two spare words per warp
        if (direct) {
            k_throw_philox<TILE_W, TILE_H, true><<<grid, 256, smem, st>>>(p, keys, *direct);
        } else {
            wb200_gather_args none;
            memset(&none, 0, sizeof(none));
            k_throw_philox<TILE_W, TILE_H, false><<<grid, 256, smem, st>>>(p, keys, none);
        }
        WB_LAUNCHED("k_throw_philox");
        return WB200_OK;
    }
    case WB200_RNG_RANDR: {
        WB_REQUIRE(a->d_offsets && a->d_totals && a->d_seeds && a->threads >= 1, "RANDR inputs");
        int rc = ensure_lcg_tables();
        if (rc)
            return rc;
        return launch_throw<WB200_RNG_RANDR>(p, st);
    }
    case WB200_RNG_HOST:
        WB_REQUIRE(a->d_offsets && a->d_totals && a->d_normals && a->d_normals_base, "HOST inputs");
        return launch_throw<WB200_RNG_HOST>(p, st);
    default:
        return fail(WB200_ERR_ARG, "unknown rng_mode%s%s");
    }
}
int launch_counts(const wb200_counts_args *a, cudaStream_t st)
{
    WB_REQUIRE(a != nullptr, "null args");
    WB_REQUIRE(a->n_samples > 0 && a->n_samples <= 65535 && a->n_bins > 0, "bad sizes");
    WB_REQUIRE(a->d_flux && a->d_sens && a->d_dwl && a->d_dur_ms && a->d_totals, "null buffer");
    WB_REQUIRE(a->count_mode >= 0 && a->count_mode <= 2, "bad count_mode");
    WB_REQUIRE(a->count_mode == WB200_COUNT_NONE || a->d_counts, "counts output needed");
    WB_REQUIRE(!a->d_cheb_coef || (a->d_cheb_x && a->cheb_order >= 1 && a->cheb_order <= 32),
               "Chebyshev planet signal: need x and 1 <= order <= 32");
    WB_CUDA(cudaMemsetAsync(a->d_totals, 0, sizeof(uint64_t) * a->n_samples, st));
    if (a->count_mode == WB200_COUNT_POISSON && !getenv("WB200_GENERIC_COUNTS")) {
        // native mode: CDF-window sampler, a thread owns a bin for CW_BLOCK sub-samples
        static_assert(CW_BLOCK % 2 == 0, "Philox words are shared by sub-sample pairs");
        dim3 wgrid((a->n_bins + CW_THREADS - 1) / CW_THREADS, (a->n_samples + CW_BLOCK - 1) / CW_BLOCK);
        k_counts_window<<<wgrid, CW_THREADS, sizeof(float) * CW_T * CW_THREADS, st>>>(
            a->n_samples, a->n_bins, a->d_flux, a->d_depth, (long long)a->depth_ld, a->d_cheb_coef,
            a->cheb_order, a->d_cheb_x, a->d_sens, a->d_dwl, a->d_dur_ms, a->scale, a->key0, a->key1,
            a->d_expected, a->d_counts, (unsigned long long *)a->d_totals, a->d_sep_row);
        WB_LAUNCHED("k_counts_window");
        return WB200_OK;
    }
    dim3 grid((a->n_bins + COUNTS_THREADS - 1) / COUNTS_THREADS, (a->n_samples + COUNTS_SPT - 1) / COUNTS_SPT);
    k_counts<<<grid, COUNTS_THREADS, 0, st>>>(a->n_samples, a->n_bins, a->d_flux, a->d_depth, (long long)a->depth_ld,
                                   a->d_cheb_coef, a->cheb_order, a->d_cheb_x, a->d_sens, a->d_dwl,
                                   a->d_dur_ms, a->scale, a->count_mode, a->key0, a->key1, a->d_expected,
                                   a->d_counts, (unsigned long long *)a->d_totals, a->d_sep_row);
    WB_LAUNCHED("k_counts");
    return WB200_OK;
}


int launch_reads(const wb200_reads_args *a, cudaStream_t st)
{
    WB_REQUIRE(a != nullptr, "null args");
    WB_REQUIRE(a->n_reads >= 1 && a->n_reads <= 15, "1 <= n_reads <= 15");
    WB_REQUIRE(a->F > 2 * a->border && (a->F % 2) == 0, "F must be even and > 2*border");
    WB_REQUIRE(a->d_dt && a->d_acc && a->d_out, "null buffer");
    WB_REQUIRE(!a->add_sky || a->d_sky || a->d_draw_sky, "sky plane missing");
    WB_REQUIRE(!a->add_dark || a->d_draw_dark || (a->d_dark && a->d_dark_err), "dark planes missing");
    if (a->add_nonlinear)
        for (int i = 0; i < (a->planes_f32 ? 4 : 7); ++i)
            WB_REQUIRE(a->d_nl[i], "non-linearity planes missing");
    WB_REQUIRE(!a->d_cos_head || a->n_cosmics == 0 || (a->d_cos_next && a->d_cos_read && a->d_cos_energy),
               "cosmic lists missing");
    WB_REQUIRE(!a->zero_acc || a->acc_fixed, "zero_acc applies to the fixed-point interval planes");
    const size_t pairs = (size_t)a->F * a->F / 2;
    const int blocks = (int)((pairs + 255) / 256);
    const bool native = a->fast_math && !a->d_draw_noise && !a->d_draw_sky && !a->d_draw_dark && !a->d_draw_rn &&
                        !(a->exact_newton && a->add_nonlinear) &&
                        (a->planes_f32 || (!getenv("WB200_EXACT_READS") && !getenv("WB200_GENERIC_READS")));
    WB_REQUIRE(native || !a->planes_f32, "float32 planes are read by the native kernel only");
    if (native) {
        // native mode without host-drawn planes: the throughput kernel (reads_native.cuh)
        static const size_t smem = RN_SMEM;
        static const cudaError_t attr[4] = {
            cudaFuncSetAttribute(k_reads_native<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            cudaFuncSetAttribute(k_reads_native<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            cudaFuncSetAttribute(k_reads_native<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            cudaFuncSetAttribute(k_reads_native<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)};
        for (cudaError_t e : attr)
            WB_CUDA(e);
        const int nblocks = (int)((pairs + RN_THREADS - 1) / RN_THREADS);
        if (a->planes_f32) {
            if (a->out_f32)
                k_reads_native<true, true><<<nblocks, RN_THREADS, smem, st>>>(*a);
            else
                k_reads_native<false, true><<<nblocks, RN_THREADS, smem, st>>>(*a);
        } else {
            if (a->out_f32)
                k_reads_native<true, false><<<nblocks, RN_THREADS, smem, st>>>(*a);
            else
                k_reads_native<false, false><<<nblocks, RN_THREADS, smem, st>>>(*a);
        }
        WB_LAUNCHED("k_reads_native");
        return WB200_OK;
    }
    if (a->exact_newton && a->add_nonlinear) {
        WB_REQUIRE(!a->out_f32, "exact_newton needs the float64 output");
        WB_REQUIRE(a->d_newton_iters, "newton scratch missing");
        WB_CUDA(cudaMemsetAsync(a->d_newton_iters, 0, sizeof(int32_t) * 16, st));
        k_reads<1><<<blocks, 256, 0, st>>>(*a);
        WB_LAUNCHED("k_reads<1>");
        k_reads<2><<<blocks, 256, 0, st>>>(*a);
        WB_LAUNCHED("k_reads<2>");
    } else {
        if (a->fast_math && !getenv("WB200_EXACT_READS"))
            k_reads<0, true><<<blocks, 256, 0, st>>>(*a);
        else
            k_reads<0, false><<<blocks, 256, 0, st>>>(*a);
        WB_LAUNCHED("k_reads<0>");
    }
    if (a->zero_acc) // the generic kernels do not write the planes back
        WB_CUDA(cudaMemsetAsync(const_cast<void *>(a->d_acc), 0, sizeof(long long) * (size_t)a->n_reads * a->F * a->F, st));
    return WB200_OK;
}
} // namespace wb

using namespace wb;


extern "C" {

const char *wb200_last_error(void) { return g_err; }
int wb200_version(void) { return 100; }
uint64_t wb200_launch_count(void) { return g_launches.load(); }

int wb200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail(WB200_ERR_CUDA, "cudaGetDeviceCount: %s%s", cudaGetErrorString(e));
        return WB200_ERR_CUDA;
    }
    return n;
}

int wb200_bin_tables(int n_bins, const double *d_wl, const double *psf_poly12, int n_sens,
                     const double *d_sens_wl, const double *d_sens_val, double *d_ratio,
                     double *d_sigl, double *d_sigh, double *d_sens, double *d_dwl, void *stream)
{
    WB_REQUIRE(n_bins >= 2 && n_sens >= 1, "need >= 2 bins and a sensitivity table");
    WB_REQUIRE(d_wl && psf_poly12 && d_sens_wl && d_sens_val, "null input");
    WB_REQUIRE(d_ratio && d_sigl && d_sigh && d_sens && d_dwl, "null output");
    // the 12 polynomial coefficients are a HOST array: they travel by value as a
    // kernel parameter (no staging allocation, no copy)
    Poly12 poly;
    memcpy(poly.c, psf_poly12, sizeof(poly.c));
    k_bin_tables<<<(n_bins + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        n_bins, d_wl, poly, n_sens, d_sens_wl, d_sens_val, d_ratio, d_sigl, d_sigh, d_sens, d_dwl);
    WB_LAUNCHED("k_bin_tables");
    return WB200_OK;
}

int wb200_trace_table(int n_samples, const double *d_xref, const double *d_yref,
                      const double *trace_coeff9, const double *wl_sol9, double *d_trace,
                      void *stream)
{
    WB_REQUIRE(n_samples >= 0 && d_xref && d_yref && trace_coeff9 && wl_sol9 && d_trace, "bad args");
    if (n_samples == 0)
        return WB200_OK;
    Coef18 c;
    memcpy(c.a, trace_coeff9, sizeof(c.a));
    memcpy(c.b, wl_sol9, sizeof(c.b));
    k_trace_table<<<(n_samples + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_samples, d_xref,
                                                                             d_yref, c, d_trace);
    WB_LAUNCHED("k_trace_table");
    return WB200_OK;
}

int wb200_trace_positions(int n_samples, int n_bins, const double *d_trace, const double *d_wl,
                          double sub_scale, double *d_xpos, double *d_ypos, void *stream)
{
    WB_REQUIRE(n_samples > 0 && n_samples <= 65535 && n_bins > 0, "bad sizes");
    WB_REQUIRE(d_trace && d_wl && d_xpos && d_ypos, "null buffer");
    dim3 grid((n_bins + 127) / 128, n_samples);
    k_trace_positions<<<grid, 128, 0, (cudaStream_t)stream>>>(n_samples, n_bins, d_trace, d_wl,
                                                              sub_scale, d_xpos, d_ypos);
    WB_LAUNCHED("k_trace_positions");
    return WB200_OK;
}

int wb200_counts_ex(const wb200_counts_args *a, void *stream) { return launch_counts(a, (cudaStream_t)stream); }

int wb200_counts(int n_samples, int n_bins, const double *d_flux, const double *d_depth,
                 int64_t depth_ld, const double *d_sens, const double *d_dwl,
                 const double *d_dur_ms, double scale, int count_mode, uint32_t key0,
                 uint32_t key1, double *d_expected, int32_t *d_counts, uint64_t *d_totals,
                 void *stream)
{
    wb200_counts_args a;
    memset(&a, 0, sizeof(a));
    a.n_samples = n_samples;
    a.n_bins = n_bins;
    a.count_mode = count_mode;
    a.key0 = key0;
    a.key1 = key1;
    a.scale = scale;
    a.depth_ld = depth_ld;
    a.d_flux = d_flux;
    a.d_depth = d_depth;
    a.d_sens = d_sens;
    a.d_dwl = d_dwl;
    a.d_dur_ms = d_dur_ms;
    a.d_expected = d_expected;
    a.d_counts = d_counts;
    a.d_totals = d_totals;
    return wb200_counts_ex(&a, stream);
}

int wb200_transit_cheb(int n_samples, int order, const double *d_z, double p_min, double p_max,
                       const double *ld4, int n_gl, const double *d_gl_x, const double *d_gl_w,
                       double *d_coef, void *stream)
{
    WB_REQUIRE(n_samples > 0 && order >= 1 && order <= 32, "1 <= order <= 32");
    WB_REQUIRE(d_z && ld4 && d_gl_x && d_gl_w && d_coef && n_gl >= 2, "null buffer");
    WB_REQUIRE(p_max >= p_min && p_min >= 0.0, "bad radius-ratio interval");
    Claret4 ld;
    memcpy(ld.c, ld4, sizeof(ld.c));
    double total = 1.0; // int_0^1 I(r) 2 pi r dr, with int mu^(n/2) 2 r dr = 4/(n+4)
    for (int n = 1; n <= 4; ++n)
        total -= ld.c[n - 1] * (1.0 - 4.0 / (n + 4));
    total *= 3.141592653589793;
    k_transit_cheb<<<n_samples, 32, 0, (cudaStream_t)stream>>>(n_samples, order, d_z, p_min, p_max, ld,
                                                               total, n_gl, d_gl_x, d_gl_w, d_coef);
    WB_LAUNCHED("k_transit_cheb");
    return WB200_OK;
}

int wb200_count_offsets(int n_samples, int n_bins, const int32_t *d_counts, int32_t *d_offsets,
                        void *stream)
{
    WB_REQUIRE(n_samples > 0 && n_bins > 0 && d_counts && d_offsets, "bad args");
    k_count_offsets<<<n_samples, 256, 0, (cudaStream_t)stream>>>(n_bins, d_counts, d_offsets);
    WB_LAUNCHED("k_count_offsets");
    return WB200_OK;
}

int wb200_throw_photons(const wb200_photon_args *args, void *stream)
{
    return throw_photons(args, 0, (cudaStream_t)stream, nullptr);
}

// Batches of an exposure carry the global index of their first sub-sample so
// Philox counters do not depend on the batching.
int wb200_throw_photons_at(const wb200_photon_args *args, int sample0, void *stream)
{
    return throw_photons(args, sample0, (cudaStream_t)stream, nullptr);
}

int wb200_throw_photons_direct(const wb200_photon_args *args, const wb200_gather_args *g,
                               int sample0, void *stream)
{
    WB_REQUIRE(args && g, "null args");
    WB_REQUIRE(g->n_reads > 0 && g->L > 0 && g->F >= g->L && g->d_read_end && g->d_acc, "bad geometry");
    WB_REQUIRE(args->d_trace, "direct accumulation needs the trace table");
    WB_REQUIRE(!g->add_flat || (g->d_flat[0] && g->d_flat[1] && g->d_flat[2] && g->d_flat[3]),
               "flat planes missing");
    return throw_photons(args, sample0, (cudaStream_t)stream, g);
}

int wb200_gather_flat(const wb200_gather_args *a, void *stream)
{
    WB_REQUIRE(a != nullptr, "null args");
    WB_REQUIRE(a->n_reads > 0 && a->n_reads <= 65535 && a->L > 0 && a->F >= a->L, "bad geometry");
    WB_REQUIRE(a->d_read_end && a->d_win && a->d_win_ox && a->d_win_oy && a->d_trace && a->d_acc,
               "null buffer");
    WB_REQUIRE(!a->add_flat || (a->d_flat[0] && a->d_flat[1] && a->d_flat[2] && a->d_flat[3]),
               "flat planes missing");
    if (a->n_samples == 0)
        return WB200_OK;
    dim3 grid((a->L + GX - 1) / GX, (a->L + GY - 1) / GY, a->n_reads);
    dim3 block(GX, GY);
    if (a->exact)
        k_gather<true><<<grid, block, 0, (cudaStream_t)stream>>>(*a);
    else
        k_gather<false><<<grid, block, 0, (cudaStream_t)stream>>>(*a);
    WB_LAUNCHED("k_gather");
    return WB200_OK;
}

int wb200_cosmic_chains(int n_hits, const int32_t *d_pixel, int32_t n_pixels, int32_t *d_head,
                        int32_t *d_next, void *stream)
{
    WB_REQUIRE(n_hits >= 0 && n_pixels > 0 && d_head, "bad args");
    cudaStream_t st = (cudaStream_t)stream;
    WB_CUDA(cudaMemsetAsync(d_head, 0xff, sizeof(int32_t) * (size_t)n_pixels, st));
    if (n_hits == 0)
        return WB200_OK;
    WB_REQUIRE(d_pixel && d_next, "null hit list");
    k_cosmic_chains<<<(n_hits + 127) / 128, 128, 0, st>>>(n_hits, d_pixel, n_pixels, d_head, d_next);
    WB_LAUNCHED("k_cosmic_chains");
    return WB200_OK;
}

int wb200_reads(const wb200_reads_args *args, void *stream) { return launch_reads(args, (cudaStream_t)stream); }

// ---------------------------------------------------------------------------
// PSF drop-in (wayne/pyparallel_menu.c:10-113): host buffers in, host frame out.
// ---------------------------------------------------------------------------

namespace {
// Per-thread scratch of the PSF() drop-in: device buffers, pinned staging and a stream that
// live as long as the calling thread, so a call costs one host->device copy, the kernels and
// one device->host copy -- not seven cudaMalloc / cudaFree pairs and eight pageable copies.
struct PsfScratch {
    int device = -1;
    cudaStream_t st = nullptr;
    char *h_in = nullptr, *d_in = nullptr;
    size_t in_cap = 0;
    int *h_out = nullptr, *d_win = nullptr;
    size_t out_cap = 0;
    int32_t *d_off = nullptr;
    size_t off_cap = 0;
    double *d_norm = nullptr;
    size_t norm_cap = 0;
    void release()
    {
        if (device < 0)
            return;
        if (h_in)
            cudaFreeHost(h_in);
        if (h_out)
            cudaFreeHost(h_out);
        for (void *p : {(void *)d_in, (void *)d_win, (void *)d_off, (void *)d_norm})
            if (p)
                cudaFree(p);
        if (st)
            cudaStreamDestroy(st);
        *this = PsfScratch();
    }
    ~PsfScratch() { release(); }
};
thread_local PsfScratch g_psf;

struct PsfHeader { // travels in front of the packed inputs
    uint64_t lost, total;
    int32_t ox, oy, seed, pad;
    int64_t nbase;
};
} // namespace

int wb200_psf_host(const int *counts, int size, const double *x_pos, const double *y_pos,
                   const double *psf_ratio, const double *psf_sigmal, const double *psf_sigmah,
                   int nr, int nc, int test, int threads, int rng_mode, const double *normals,
                   int *frame_out)
{
    WB_REQUIRE(size >= 0 && nr > 0 && nc > 0 && frame_out, "bad sizes");
    WB_REQUIRE(size == 0 || (counts && x_pos && y_pos && psf_ratio && psf_sigmal && psf_sigmah),
               "null input");
    WB_REQUIRE(rng_mode != WB200_RNG_HOST || normals, "normals missing");
    WB_REQUIRE(rng_mode != WB200_RNG_RANDR || threads >= 1, "threads >= 1");
    const size_t npix = (size_t)nr * nc;
    if (size == 0) {
        memset(frame_out, 0, npix * sizeof(int));
        return WB200_OK;
    }
    long long ssum = 0;
    for (int i = 0; i < size; ++i)
        ssum += counts[i] > 0 ? counts[i] : 0;
    WB_REQUIRE(ssum < 2147483647LL, "more than 2^31-1 electrons (the reference's int overflows too)");

    PsfScratch &S = g_psf;
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (S.device != dev) {
        S.release();
        WB_CUDA(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
        S.device = dev;
    }
    const size_t nb = (size_t)size;
    // packed inputs: header | counts int32 [nb] (padded to 8) | x, y, ratio, sigl, sigh double [nb] each
    const size_t off_counts = sizeof(PsfHeader);
    const size_t off_dbl = off_counts + (nb * 4 + 7) / 8 * 8;
    const size_t in_bytes = off_dbl + nb * 8 * 5;
    if (S.in_cap < in_bytes) {
        if (S.h_in)
            cudaFreeHost(S.h_in);
        if (S.d_in)
            cudaFree(S.d_in);
        S.h_in = S.d_in = nullptr;
        S.in_cap = 0;
        const size_t cap = in_bytes + in_bytes / 2 + 4096;
        WB_CUDA(cudaHostAlloc((void **)&S.h_in, cap, cudaHostAllocDefault));
        WB_CUDA(cudaMalloc((void **)&S.d_in, cap));
        S.in_cap = cap;
    }
    if (S.out_cap < npix) {
        if (S.h_out)
            cudaFreeHost(S.h_out);
        if (S.d_win)
            cudaFree(S.d_win);
        S.h_out = S.d_win = nullptr;
        S.out_cap = 0;
        WB_CUDA(cudaHostAlloc((void **)&S.h_out, npix * sizeof(int), cudaHostAllocDefault));
        WB_CUDA(cudaMalloc((void **)&S.d_win, npix * sizeof(int)));
        S.out_cap = npix;
    }
    if (S.off_cap < nb) {
        if (S.d_off)
            cudaFree(S.d_off);
        S.d_off = nullptr;
        S.off_cap = 0;
        WB_CUDA(cudaMalloc((void **)&S.d_off, (nb + nb / 2 + 256) * 4));
        S.off_cap = nb + nb / 2 + 256;
    }
    PsfHeader hd = {0, (uint64_t)ssum, 0, 0, test, 0, 0};
    memcpy(S.h_in, &hd, sizeof(hd));
    int32_t *hc = (int32_t *)(S.h_in + off_counts);
    for (size_t i = 0; i < nb; ++i) // negative counts never throw (the reference's loops simply do not run)
        hc[i] = counts[i] > 0 ? counts[i] : 0;
    const double *src[5] = {x_pos, y_pos, psf_ratio, psf_sigmal, psf_sigmah};
    for (int i = 0; i < 5; ++i)
        memcpy(S.h_in + off_dbl + (size_t)i * nb * 8, src[i], nb * 8);
    cudaStream_t st = S.st;
    WB_CUDA(cudaMemcpyAsync(S.d_in, S.h_in, in_bytes, cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemsetAsync(S.d_win, 0, npix * 4, st));
    if (rng_mode == WB200_RNG_HOST) {
        const size_t n = (size_t)ssum * 2;
        if (S.norm_cap < n) {
            if (S.d_norm)
                cudaFree(S.d_norm);
            S.d_norm = nullptr;
            S.norm_cap = 0;
            WB_CUDA(cudaMalloc((void **)&S.d_norm, (n ? n : 1) * 8));
            S.norm_cap = n ? n : 1;
        }
        WB_CUDA(cudaMemcpyAsync(S.d_norm, normals, n * 8, cudaMemcpyHostToDevice, st));
    }
    int rc = wb200_count_offsets(1, size, (const int32_t *)(S.d_in + off_counts), S.d_off, st);
    if (rc)
        return rc;

    wb200_photon_args a;
    memset(&a, 0, sizeof(a));
    a.n_samples = 1;
    a.n_bins = size;
    // One sub-sample must still fill 148 SMs: chunks of at most 256 bins (a CTA's eight warps
    // take a 32-bin group each), and the electron stream of a chunk shared by n_split CTAs so
    // that ~4 warps per scheduler are busy whatever the number of bins.
    int chunk = ((size + 147) / 148 + 31) / 32 * 32;
    if (chunk > 256)
        chunk = 256;
    a.chunk_bins = chunk < 32 ? 32 : chunk;
    const int chunks = (size + a.chunk_bins - 1) / a.chunk_bins;
    const long long groups = (size + 31) / 32;
    long long n_split = (148LL * 16 + groups - 1) / groups;                       // target ~2400 busy warps
    const long long by_work = (ssum / 32 + groups * 8 - 1) / (groups * 8 > 0 ? groups * 8 : 1); // >= 8 rows of 32 electrons per warp
    if (n_split > by_work)
        n_split = by_work;
    if (n_split < 1)
        n_split = 1;
    if (n_split > 64)
        n_split = 64;
    (void)chunks;
    a.nr = nr;
    a.nc = nc;
    a.rng_mode = rng_mode;
    a.threads = threads;
    a.win_w = nc;
    a.win_h = nr;
    a.sub_scale = 0.0;
    a.key0 = (uint32_t)test;
    a.key1 = 0x57415945u; // "WAYE"
    a.d_counts = (const int32_t *)(S.d_in + off_counts);
    a.d_offsets = S.d_off;
    PsfHeader *dh = (PsfHeader *)S.d_in;
    a.d_totals = &dh->total;
    const double *dd = (const double *)(S.d_in + off_dbl);
    a.d_xpos = dd;
    a.d_ypos = dd + nb;
    a.d_ratio = dd + 2 * nb;
    a.d_sigl = dd + 3 * nb;
    a.d_sigh = dd + 4 * nb;
    a.d_lost = &dh->lost;
    a.d_win_ox = &dh->ox;
    a.d_win_oy = &dh->oy;
    a.d_seeds = &dh->seed;
    a.d_normals_base = &dh->nbase;
    a.d_normals = S.d_norm;
    a.d_win = S.d_win;
    rc = throw_photons(&a, 0, st, nullptr, rng_mode == WB200_RNG_PHILOX ? 1 : (int)n_split);
    if (rc)
        return rc;
    WB_CUDA(cudaMemcpyAsync(S.h_out, S.d_win, npix * 4, cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaMemcpyAsync(S.h_in, S.d_in, sizeof(PsfHeader), cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaStreamSynchronize(st));
    if (((PsfHeader *)S.h_in)->lost)
        return fail(WB200_ERR_LOST, "electrons fell outside the frame window%s%s");
    memcpy(frame_out, S.h_out, npix * sizeof(int));
    return WB200_OK;
}

int *PSF(int *counts, int size, double *x_pos, double *y_pos, double *psf_ratio,
         double *psf_sigmal, double *psf_sigmah, int nr, int nc, int test, int threads)
{
    if (nr <= 0 || nc <= 0) {
        fail(WB200_ERR_ARG, "PSF: bad frame size%s%s");
        return nullptr;
    }
    int *frame = (int *)malloc((size_t)nr * nc * sizeof(int)); // caller frees (pyparallel.pyx:36)
    if (!frame) {
        fail(WB200_ERR_NOMEM, "PSF: malloc failed%s%s");
        return nullptr;
    }
    const int rc = wb200_psf_host(counts, size, x_pos, y_pos, psf_ratio, psf_sigmal, psf_sigmah, nr,
                                  nc, test, threads, WB200_RNG_RANDR, nullptr, frame);
    if (rc != WB200_OK) {
        free(frame);
        return nullptr;
    }
    return frame;
}

int wb200_philox_words(int which, const uint32_t *c, const uint32_t *k, uint32_t *out)
{
    WB_REQUIRE((which == 0 || which == 1) && c && k && out, "bad args");
    uint32_t *d = nullptr;
    WB_CUDA(cudaMalloc(&d, 6 * sizeof(uint32_t)));
    k_philox_words<<<1, 1>>>(which, make_uint4(c[0], c[1], c[2], c[3]), k[0], k[1], d);
    WB_LAUNCHED("k_philox_words");
    const cudaError_t e = cudaMemcpy(out, d, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    WB_CUDA(e);
    return WB200_OK;
}

int wb200_microbench(int which, int iters, double *ms_out, double *ops_out)
{
    WB_REQUIRE(which >= 0 && which <= 13 && iters > 0 && ms_out && ops_out, "bad args");
    cudaDeviceProp prop;
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    WB_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8;
    int *buf = nullptr;
    const int n = 1 << 24; // 64 MB of ints
    WB_CUDA(cudaMalloc(&buf, (size_t)n * 4));
    WB_CUDA(cudaMemset(buf, 0, (size_t)n * 4));
    cudaEvent_t e0, e1;
    WB_CUDA(cudaEventCreate(&e0));
    WB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        WB_CUDA(cudaEventRecord(e0));
        if (which <= 3)
            k_mb_smem_atomic<<<blocks, 256>>>(which, iters, buf);
        else if (which <= 5)
            k_mb_global_red<<<blocks, 256>>>(which, iters, buf, n);
        else if (which <= 7)
            k_mb_rng<<<blocks, 256>>>(which, iters, buf);
        else if (which == 8)
            k_mb_pipe<8><<<blocks, 256>>>(iters, buf);
        else if (which == 9)
            k_mb_pipe<9><<<blocks, 256>>>(iters, buf);
        else if (which == 10)
            k_mb_pipe<10><<<blocks, 256>>>(iters, buf);
        else if (which == 11)
            k_mb_pipe<11><<<blocks, 256>>>(iters, buf);
        else if (which == 12)
            k_mb_pipe<12><<<blocks, 256>>>(iters, buf);
        else
            k_mb_pipe<13><<<blocks, 256>>>(iters, buf);
        WB_LAUNCHED("microbench");
        WB_CUDA(cudaEventRecord(e1));
        WB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        WB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best)
            best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *ms_out = best;
    double ops = (double)blocks * 256.0 * iters;
    if (which == 7)
        ops *= 4.0; // four electrons per Philox call
    if (which >= 8)
        ops *= 64.0; // probed instructions per iteration and thread
    *ops_out = ops;
    return WB200_OK;
}

} // extern "C"

#include "context.cuh"
