// context.cuh -- the exposure-level interface (include/wayne_b200.h, "Exposure-level
// interface"): one call per exposure on an opaque per-GPU context.
//
// The reference crosses from Python into native code once per SUB-SAMPLE
// (wayne/pyparallel.pyx:27-30 from wayne/exposure_generator.py:636-639, inside the loop
// :336-394) and does the per-read and post-exposure passes in numpy (:361-389, :407-444).
// Here the host hands over the exposure's small inputs and the library queues the whole
// native-mode pipeline on one stream:
//
//   staging copy (upload stream) -> k_bin_tables, k_trace_table -> k_counts_window / k_counts
//   -> [k_cosmic_chains] -> k_throw_philox<DIRECT> -> k_reads_native<PLANES32> -> k_exposure_stats
//
// What the context owns
//   planes   resident calibration planes, float32 (the files' dtype), uploaded once
//   scratch  tables [5][W], trace [N][8], counts [N][W], totals [N], the int64 fixed-point
//            interval planes acc [R][F][F] (kept ZERO between exposures: k_reads_native
//            writes the zeros back, so no memset pass), cosmic chains
//   staging  a ring of pinned host + device buffers for the small per-exposure arrays
//            (one host->device copy per exposure, on the context's upload stream so it
//            never queues behind the previous exposure's kernels), each slot guarded by
//            two events: "copied" (the compute stream waits for it) and "done" (recorded
//            after the exposure's last kernel; the slot is reused only after it)
#pragma once
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "counts_native.cuh"
#include "photons.cuh"
#include "reads_native.cuh"
#include "stage1.cuh"

namespace wb {
int launch_reads(const wb200_reads_args *a, cudaStream_t st);   // wayne_b200.cu
int launch_counts(const wb200_counts_args *a, cudaStream_t st); // wayne_b200.cu
int throw_photons(const wb200_photon_args *a, int sample0, cudaStream_t st, const wb200_gather_args *direct,
                  int n_split, const int *d_chunk_span);

// electrons thrown / binned / dropped of one exposure, for the caller's bookkeeping
__global__ void __launch_bounds__(256)
k_exposure_stats(int N, const unsigned long long *__restrict__ totals, const unsigned long long *__restrict__ tally,
                 unsigned long long *stats)
{
    __shared__ unsigned long long part[8];
    unsigned long long v = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        v += totals[i];
    v = warp_sum_u64(v);
    if (lane_id() == 0)
        part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i)
            t += part[i];
        stats[0] = t;
        stats[1] = tally[0];
        stats[2] = tally[1];
        stats[3] = 0;
    }
}
} // namespace wb

struct wb200_ctx {
    int device = 0;
    char err[512] = "";
    bool have_inst = false;
    wb200_instrument inst;
    cudaStream_t upload = nullptr;

    struct Buf {
        void *d = nullptr;
        size_t cap = 0;
    };
    Buf planes[WB200_PLANE_COUNT];
    int64_t plane_count[WB200_PLANE_COUNT] = {};
    Buf acc, tally;
    // Two LANES of the per-exposure scratch that stage 1 and the count sampler write: the tables,
    // traces and counts of exposure i+1 are made on the context's own `pre` stream while exposure
    // i's electrons are still being thrown on the caller's stream (the thrower's grid drains over
    // the duration of one CTA, the ramp pass runs at a third of the SM's warps: the count
    // sampler's CTAs fill those holes).  `pre_done`: the lane's stage-1 work (recorded on `pre`);
    // `freed`: the last kernel READING the lane has been queued on the caller's stream.
    static constexpr int LANES = 2;
    struct Lane {
        Buf tables, trace, counts, totals, cos_head, cos_next, span;
        cudaEvent_t pre_done = nullptr, freed = nullptr;
        bool used = false;
    } lane[LANES];
    int next_lane = 0, last_lane = 0;
    cudaStream_t pre = nullptr;
    cudaEvent_t inputs = nullptr; // "device-resident inputs are ready" (recorded on the caller's stream)
    bool overlap = true;          // WB200_CTX_SERIAL=1: everything on the caller's stream

    static constexpr int SLOTS = 6;
    struct Slot {
        char *h = nullptr;
        char *d = nullptr;
        size_t cap = 0;
        char *big = nullptr; // dense planet signal of the exposure that owns the slot
        size_t big_cap = 0;
        cudaEvent_t copied = nullptr, done = nullptr;
        bool busy = false;
    } slot[SLOTS];
    int next_slot = 0;

    // geometry cache of the thrower (bins per CTA): depends on the set-up, not on the pointing
    int chunk_bins = 0, chunk_W = 0, chunk_N = 0;
    double chunk_wl0 = 0, chunk_wl1 = 0;
    int64_t exposures = 0, last_launches = 0;
    int last_N = 0, last_W = 0;

    // optional per-stage timing (bench / profiles): CUDA events on the launching stream
    bool profile = false;
    struct Mark {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Mark> marks;
    cudaEvent_t open_ev = nullptr;
    int open_stage = -1;
};

namespace wb {

inline int ctx_fail(wb200_ctx *c, int code, const char *fmt, const char *a = "", const char *b = "")
{
    fail(code, fmt, a, b);
    if (c)
        snprintf(c->err, sizeof(c->err), "%s", g_err);
    return code;
}

#define CTX_CUDA(c, call)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return wb::ctx_fail(c, WB200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));     \
    } while (0)
#define CTX_REQUIRE(c, cond, msg)                                                                 \
    do {                                                                                          \
        if (!(cond))                                                                              \
            return wb::ctx_fail(c, WB200_ERR_ARG, "%s (%s)", msg, #cond);                         \
    } while (0)
// a stage-level call failed: its message is in g_err already
#define CTX_STAGE(c, call)                                                                        \
    do {                                                                                          \
        int rc__ = (call);                                                                        \
        if (rc__ != WB200_OK) {                                                                   \
            snprintf((c)->err, sizeof((c)->err), "%s", wb::g_err);                                \
            return rc__;                                                                          \
        }                                                                                         \
    } while (0)

// stage brackets: begin(stage) ... end() around a group of launches on `st`
inline void mark_begin(wb200_ctx *c, int stage, cudaStream_t st)
{
    if (!c->profile)
        return;
    if (cudaEventCreate(&c->open_ev) != cudaSuccess) {
        c->open_ev = nullptr;
        return;
    }
    cudaEventRecord(c->open_ev, st);
    c->open_stage = stage;
}
inline void mark_end(wb200_ctx *c, cudaStream_t st)
{
    if (!c->profile || !c->open_ev)
        return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) == cudaSuccess) {
        cudaEventRecord(b, st);
        c->marks.push_back({c->open_stage, c->open_ev, b});
    } else {
        cudaEventDestroy(c->open_ev);
    }
    c->open_ev = nullptr;
}

inline int ctx_reserve(wb200_ctx *c, wb200_ctx::Buf &b, size_t bytes, bool zero = false)
{
    if (bytes <= b.cap)
        return WB200_OK;
    if (b.d) {
        CTX_CUDA(c, cudaDeviceSynchronize()); // an earlier exposure may still be reading it
        cudaFree(b.d);
        b.d = nullptr;
        b.cap = 0;
    }
    if (cudaMalloc(&b.d, bytes) != cudaSuccess) {
        cudaGetLastError();
        return ctx_fail(c, WB200_ERR_NOMEM, "cudaMalloc of a scratch buffer failed%s%s");
    }
    b.cap = bytes;
    if (zero)
        CTX_CUDA(c, cudaMemset(b.d, 0, bytes));
    return WB200_OK;
}

// Bins per CTA of the thrower: the chunk's trace segment plus 3 sigma of the wide Gaussian either
// side must fit the 128-pixel shared tile (the few electrons of the end bins that leave it are
// replayed; fewer, longer CTAs pay the tile zeroing / placement / flush less often: 2048 bins per
// CTA measured 1.91 ms, 1376 1.97, 1024 2.01, 512 2.24 on the configs[3] shape); equal chunks; at
// least ~16 CTAs per SM in the grid.  The trace
// length depends on the grism and the wavelength grid, hardly on the pointing: evaluated on every
// (N/32)-th sub-sample and cached per (W, N, wavelength range).
inline int choose_chunk_bins(const wb200_instrument &I, const wb200_exposure_args &a)
{
    const int N = a.n_samples, W = a.n_bins;
    double wl_lo = a.wl[0], wl_hi = a.wl[0], sig = 0.0;
    for (int w = 0; w < W; ++w) {
        const double x = a.wl[w];
        wl_lo = fmin(wl_lo, x);
        wl_hi = fmax(wl_hi, x);
        double sh = 0, sl = 0;
        for (int k = 0; k < 4; ++k) {
            sl = sl * x + I.psf_poly12[4 + k];
            sh = sh * x + I.psf_poly12[8 + k];
        }
        sig = fmax(sig, fmax(fabs(sl), fabs(sh)));
    }
    const double *ta = I.trace_coeff9, *tb = I.wl_sol9;
    double extent = 0.0;
    const int stride = N / 32 > 1 ? N / 32 : 1;
    for (int s = 0; s < N; s += stride) {
        const double x = a.xref[s], y = a.yref[s];
        const double m_t = ta[3] + ta[4] * x + ta[5] * y + ta[6] * x * x + ta[7] * x * y + ta[8] * y * y;
        const double c_t = ta[0] + ta[1] * x + ta[2] * y;
        const double m_w = tb[3] + tb[4] * x + tb[5] * y + tb[6] * x * x + tb[7] * x * y + tb[8] * y * y;
        const double c_w = tb[0] + tb[1] * x + tb[2] * y;
        const double X0 = x + 10, X1 = x + 20;
        const double Y0 = m_t * (X0 - x) + c_t + y, Y1 = m_t * (X1 - x) + c_t + y;
        const double l0 = (m_w * sqrt((Y0 - y) * (Y0 - y) + (X0 - x) * (X0 - x)) + c_w) * 1e-4;
        const double l1 = (m_w * sqrt((Y1 - y) * (Y1 - y) + (X1 - x) * (X1 - x)) + c_w) * 1e-4;
        const double m_wl = (l1 - l0) / (X1 - X0), c_wl = l0 - m_wl * X0;
        const double xa = (wl_lo - c_wl) / m_wl, xb = (wl_hi - c_wl) / m_wl;
        if (isfinite(xa) && isfinite(xb))
            extent = fmax(extent, fabs(xb - xa));
    }
    extent += 1.0;
    const double core = fmax(8.0, 128.0 - 2.0 * 3.0 * sig);
    long long chunk = extent > core ? (long long)(W * core / extent) : W;
    chunk = chunk / 32 * 32;
    if (chunk < 32)
        chunk = 32;
    const long long n_chunks = (W + chunk - 1) / chunk;
    chunk = (long long)ceil(W / (double)n_chunks / 32.0) * 32;
    if (chunk < 32)
        chunk = 32;
    const long long want_ctas = 148 * 16;
    if ((long long)N * ((W + chunk - 1) / chunk) < want_ctas) {
        const long long per = want_ctas / N > 1 ? want_ctas / N : 1;
        long long c2 = (long long)ceil(W / (double)per / 32.0) * 32;
        if (c2 < 32)
            c2 = 32;
        if (c2 < chunk)
            chunk = c2;
    }
    return (int)chunk;
}

} // namespace wb

extern "C" {

int wb200_ctx_create(int device, wb200_ctx **ctx_out)
{
    WB_REQUIRE(ctx_out != nullptr, "null output");
    *ctx_out = nullptr;
    int n = 0;
    WB_CUDA(cudaGetDeviceCount(&n));
    WB_REQUIRE(device >= 0 && device < n, "no such CUDA device");
    int prev = 0;
    WB_CUDA(cudaGetDevice(&prev));
    WB_CUDA(cudaSetDevice(device));
    wb200_ctx *c = new (std::nothrow) wb200_ctx();
    if (!c)
        return wb::fail(WB200_ERR_NOMEM, "out of host memory%s%s");
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->upload, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&c->pre, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaEventCreateWithFlags(&c->inputs, cudaEventDisableTiming);
    for (int i = 0; i < wb200_ctx::LANES && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->lane[i].pre_done, cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(&c->lane[i].freed, cudaEventDisableTiming);
    }
    if (const char *env = getenv("WB200_CTX_SERIAL"))
        c->overlap = atoi(env) == 0;
    for (int i = 0; i < wb200_ctx::SLOTS && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->slot[i].copied, cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(&c->slot[i].done, cudaEventDisableTiming);
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        wb200_ctx_destroy(c);
        return wb::fail(WB200_ERR_CUDA, "context set-up: %s%s", cudaGetErrorString(e));
    }
    *ctx_out = c;
    return WB200_OK;
}

int wb200_ctx_destroy(wb200_ctx *c)
{
    if (!c)
        return WB200_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &p : c->planes)
        if (p.d)
            cudaFree(p.d);
    for (wb200_ctx::Buf *b : {&c->acc, &c->tally})
        if (b->d)
            cudaFree(b->d);
    for (auto &l : c->lane) {
        for (wb200_ctx::Buf *b : {&l.tables, &l.trace, &l.counts, &l.totals, &l.cos_head, &l.cos_next, &l.span})
            if (b->d)
                cudaFree(b->d);
        if (l.pre_done)
            cudaEventDestroy(l.pre_done);
        if (l.freed)
            cudaEventDestroy(l.freed);
    }
    if (c->inputs)
        cudaEventDestroy(c->inputs);
    if (c->pre)
        cudaStreamDestroy(c->pre);
    for (auto &m : c->marks) {
        cudaEventDestroy(m.a);
        cudaEventDestroy(m.b);
    }
    for (auto &s : c->slot) {
        if (s.h)
            cudaFreeHost(s.h);
        if (s.d)
            cudaFree(s.d);
        if (s.big)
            cudaFree(s.big);
        if (s.copied)
            cudaEventDestroy(s.copied);
        if (s.done)
            cudaEventDestroy(s.done);
    }
    if (c->upload)
        cudaStreamDestroy(c->upload);
    cudaSetDevice(prev);
    delete c;
    return WB200_OK;
}

const char *wb200_ctx_last_error(const wb200_ctx *c) { return c ? c->err : wb::g_err; }

int wb200_ctx_set_instrument(wb200_ctx *c, const wb200_instrument *inst)
{
    WB_REQUIRE(c != nullptr, "null context");
    CTX_REQUIRE(c, inst != nullptr, "null instrument");
    CTX_REQUIRE(c, inst->L > 0 && inst->border >= 0 && inst->F == inst->L + 2 * inst->border, "bad geometry");
    CTX_REQUIRE(c, inst->F % 2 == 0 && inst->F > 2 * inst->border, "F must be even and > 2*border");
    CTX_REQUIRE(c, inst->flat_n > 0 && inst->n_sens >= 1, "flat / sensitivity sizes");
    c->inst = *inst;
    c->have_inst = true;
    c->chunk_bins = 0;
    return WB200_OK;
}

int wb200_ctx_upload_plane(wb200_ctx *c, int which, const void *host, int dtype, int64_t count)
{
    WB_REQUIRE(c != nullptr, "null context");
    CTX_REQUIRE(c, which >= 0 && which < WB200_PLANE_COUNT, "unknown plane");
    CTX_REQUIRE(c, dtype == WB200_F32 || dtype == WB200_F64, "dtype must be WB200_F32 or WB200_F64");
    int prev = 0;
    CTX_CUDA(c, cudaGetDevice(&prev));
    CTX_CUDA(c, cudaSetDevice(c->device));
    struct Restore {
        int d;
        ~Restore() { cudaSetDevice(d); }
    } restore{prev};
    wb200_ctx::Buf &b = c->planes[which];
    if (!host) {
        if (b.d) {
            CTX_CUDA(c, cudaDeviceSynchronize());
            cudaFree(b.d);
        }
        b = wb200_ctx::Buf();
        c->plane_count[which] = 0;
        return WB200_OK;
    }
    CTX_REQUIRE(c, count > 0, "empty plane");
    const bool keep64 = which == WB200_PLANE_ZERO || which == WB200_PLANE_SENS_WL || which == WB200_PLANE_SENS_VAL;
    const size_t bytes = (size_t)count * (keep64 ? 8 : 4);
    std::vector<char> conv;
    const void *src = host;
    if (keep64 && dtype == WB200_F32) {
        conv.resize(bytes);
        for (int64_t i = 0; i < count; ++i)
            ((double *)conv.data())[i] = (double)((const float *)host)[i];
        src = conv.data();
    } else if (!keep64 && dtype == WB200_F64) {
        conv.resize(bytes);
        for (int64_t i = 0; i < count; ++i) {
            const double v = ((const double *)host)[i];
            const float f = (float)v;
            if ((double)f != v && v == v)
                return wb::ctx_fail(c, WB200_ERR_ARG,
                                    "plane value is not float32-representable: the resident planes are held "
                                    "in the calibration files' float32%s%s");
            ((float *)conv.data())[i] = f;
        }
        src = conv.data();
    }
    CTX_STAGE(c, wb::ctx_reserve(c, b, bytes));
    CTX_CUDA(c, cudaMemcpy(b.d, src, bytes, cudaMemcpyHostToDevice));
    c->plane_count[which] = count;
    return WB200_OK;
}

int wb200_exposure_run(wb200_ctx *c, const wb200_exposure_args *a, void *d_out, void *stream)
{
    using namespace wb;
    WB_REQUIRE(c != nullptr, "null context");
    CTX_REQUIRE(c, c->have_inst, "wb200_ctx_set_instrument has not been called");
    CTX_REQUIRE(c, a != nullptr && d_out != nullptr, "null args / output");
    const wb200_instrument &I = c->inst;
    const int N = a->n_samples, W = a->n_bins, R = a->n_reads, F = I.F;
    CTX_REQUIRE(c, N >= 1 && N <= 65535, "1..65535 sub-samples per exposure");
    CTX_REQUIRE(c, W >= 2, "need at least two wavelength bins");
    CTX_REQUIRE(c, R >= 1 && R <= 15, "1 <= n_reads <= 15");
    CTX_REQUIRE(c, a->count_mode == WB200_COUNT_ROUND || a->count_mode == WB200_COUNT_POISSON,
                "count_mode must be ROUND or POISSON (host-supplied counts belong to the parity path)");
    CTX_REQUIRE(c, a->wl && a->xref && a->yref && a->dur_ms && a->dt_s && a->read_end, "null host array");
    CTX_REQUIRE(c, a->flux || a->d_flux, "stellar flux missing");
    CTX_REQUIRE(c, a->cheb_order >= 0 && a->cheb_order <= 32, "0 <= cheb_order <= 32");
    CTX_REQUIRE(c, !a->cheb_order || (a->cheb_x && (a->cheb_coef || a->d_cheb_coef)), "Chebyshev planet signal incomplete");
    CTX_REQUIRE(c, !a->sep_row == !a->sep_col, "separable planet signal needs both factors");
    CTX_REQUIRE(c, !(a->sep_row && (a->cheb_order || a->d_depth || a->depth)) && !(a->depth && (a->d_depth || a->cheb_order)),
                "one form of the planet signal at a time");
    CTX_REQUIRE(c, !(a->depth || a->d_depth) || a->depth_ld >= W, "depth_ld must be at least n_bins");
    CTX_REQUIRE(c, a->n_cosmics >= 0 && (a->n_cosmics == 0 || (a->cos_pixel && a->cos_read && a->cos_energy)),
                "cosmic hit list incomplete");
    const size_t plane = (size_t)F * F;
    auto have = [&](int which, size_t n) { return c->planes[which].d && (size_t)c->plane_count[which] >= n; };
    CTX_REQUIRE(c, have(WB200_PLANE_SENS_WL, I.n_sens) && have(WB200_PLANE_SENS_VAL, I.n_sens), "sensitivity table not uploaded");
    if (a->add_flat)
        for (int i = 0; i < 4; ++i)
            CTX_REQUIRE(c, have(WB200_PLANE_FLAT0 + i, (size_t)I.flat_n * I.flat_n), "flat planes not uploaded");
    CTX_REQUIRE(c, !a->add_sky || have(WB200_PLANE_SKY, plane), "sky plane not uploaded");
    CTX_REQUIRE(c, !a->add_gain || have(WB200_PLANE_GAIN, plane), "gain plane not uploaded");
    CTX_REQUIRE(c, !a->add_zero || have(WB200_PLANE_ZERO, plane), "zero-read plane not uploaded");
    CTX_REQUIRE(c, !a->add_dark || (have(WB200_PLANE_DARK, R * plane) && have(WB200_PLANE_DARK_ERR, R * plane)),
                "dark planes not uploaded (or fewer reads than n_reads)");
    if (a->add_nonlinear)
        for (int i = 0; i < 4; ++i)
            CTX_REQUIRE(c, have(WB200_PLANE_NL0 + i, plane), "non-linearity planes not uploaded");
    for (int r = 0; r < R; ++r)
        CTX_REQUIRE(c, a->read_end[r] >= 0 && a->read_end[r] < N && (r == 0 || a->read_end[r] >= a->read_end[r - 1]),
                    "read_end must be non-decreasing indexes of sub-samples");

    cudaStream_t st = (cudaStream_t)stream;
    int prev = 0;
    CTX_CUDA(c, cudaGetDevice(&prev));
    if (prev != c->device)
        CTX_CUDA(c, cudaSetDevice(c->device));
    struct Restore {
        int d, cur;
        ~Restore()
        {
            if (d != cur)
                cudaSetDevice(d);
        }
    } restore{prev, c->device};
    const uint64_t launches0 = g_launches.load();

    // ---- scratch -------------------------------------------------------------------
    wb200_ctx::Lane &Ln = c->lane[c->next_lane];
    c->last_lane = c->next_lane;
    c->next_lane = (c->next_lane + 1) % wb200_ctx::LANES;
    CTX_STAGE(c, ctx_reserve(c, Ln.tables, sizeof(double) * 5 * W));
    CTX_STAGE(c, ctx_reserve(c, Ln.trace, sizeof(double) * WB200_TRACE_STRIDE * N));
    CTX_STAGE(c, ctx_reserve(c, Ln.counts, sizeof(int32_t) * (size_t)N * W));
    CTX_STAGE(c, ctx_reserve(c, Ln.totals, sizeof(uint64_t) * N));
    CTX_STAGE(c, ctx_reserve(c, c->tally, sizeof(uint64_t) * 4));
    CTX_STAGE(c, ctx_reserve(c, c->acc, sizeof(long long) * 15 * plane, /*zero=*/true));
    if (a->n_cosmics) {
        CTX_STAGE(c, ctx_reserve(c, Ln.cos_head, sizeof(int32_t) * plane));
        CTX_STAGE(c, ctx_reserve(c, Ln.cos_next, sizeof(int32_t) * a->n_cosmics));
    }

    // ---- stage the small inputs: ONE host->device copy on the upload stream ------------
    struct Piece {
        const void *src;
        size_t bytes, off;
    };
    std::vector<Piece> pieces;
    size_t total = 0;
    auto add = [&](const void *src, size_t bytes) {
        Piece p{src, bytes, total};
        total += (bytes + 15) / 16 * 16;
        pieces.push_back(p);
        return pieces.size() - 1;
    };
    const size_t i_wl = add(a->wl, 8ull * W), i_x = add(a->xref, 8ull * N), i_y = add(a->yref, 8ull * N);
    const size_t i_dur = add(a->dur_ms, 8ull * N), i_dt = add(a->dt_s, 8ull * R), i_re = add(a->read_end, 4ull * R);
    const size_t i_flux = a->d_flux ? 0 : add(a->flux, 8ull * W);
    size_t i_cx = 0, i_cc = 0, i_cp = 0, i_cr = 0, i_ce = 0;
    if (a->cheb_order) {
        i_cx = add(a->cheb_x, 8ull * W);
        if (!a->d_cheb_coef)
            i_cc = add(a->cheb_coef, 8ull * (size_t)N * a->cheb_order);
    }
    size_t i_sr = 0, i_sc = 0;
    if (a->sep_row) {
        i_sr = add(a->sep_row, 8ull * N);
        i_sc = add(a->sep_col, 8ull * W);
    }
    if (a->n_cosmics) {
        i_cp = add(a->cos_pixel, 4ull * a->n_cosmics);
        i_cr = add(a->cos_read, 4ull * a->n_cosmics);
        i_ce = add(a->cos_energy, 8ull * a->n_cosmics);
    }
    wb200_ctx::Slot &S = c->slot[c->next_slot];
    c->next_slot = (c->next_slot + 1) % wb200_ctx::SLOTS;
    if (S.busy) {
        // the pinned staging may be rewritten once the slot's previous copies have run; its device
        // buffers once the exposure that read them has finished -- the upload stream waits for
        // that, not the host
        CTX_CUDA(c, cudaEventSynchronize(S.copied));
        CTX_CUDA(c, cudaStreamWaitEvent(c->upload, S.done, 0));
    }
    const size_t big_bytes = a->depth ? sizeof(double) * ((size_t)(N - 1) * a->depth_ld + W) : 0;
    if (S.big_cap < big_bytes) {
        if (S.big) {
            CTX_CUDA(c, cudaEventSynchronize(S.done));
            cudaFree(S.big);
        }
        S.big = nullptr;
        S.big_cap = 0;
        CTX_CUDA(c, cudaMalloc((void **)&S.big, big_bytes));
        S.big_cap = big_bytes;
    }
    if (S.cap < total) {
        if (S.d)
            CTX_CUDA(c, cudaEventSynchronize(S.done));
        if (S.h)
            cudaFreeHost(S.h);
        if (S.d)
            cudaFree(S.d);
        S.h = S.d = nullptr;
        S.cap = 0;
        const size_t cap = total + total / 4 + 4096;
        CTX_CUDA(c, cudaHostAlloc((void **)&S.h, cap, cudaHostAllocDefault));
        CTX_CUDA(c, cudaMalloc((void **)&S.d, cap));
        S.cap = cap;
    }
    for (const Piece &p : pieces)
        memcpy(S.h + p.off, p.src, p.bytes);
    CTX_CUDA(c, cudaMemcpyAsync(S.d, S.h, total, cudaMemcpyHostToDevice, c->upload));
    if (a->depth) // same stream, behind the small arrays: one FIFO, nothing overtakes anything
        CTX_CUDA(c, cudaMemcpyAsync(S.big, a->depth, big_bytes, cudaMemcpyHostToDevice, c->upload));
    CTX_CUDA(c, cudaEventRecord(S.copied, c->upload));
    // Stage 1 and the count sampler go on the context's `pre` stream (see wb200_ctx::Lane): ordered
    // after the staging copy, after the last reader of this lane's buffers (the exposure before
    // the previous one) and -- only when the caller hands over device-resident inputs it does not
    // declare ready -- after what the caller's stream holds now.  With per-stage timing on, or
    // WB200_CTX_SERIAL=1, everything stays on the caller's stream.
    const bool side = c->overlap && !c->profile;
    cudaStream_t ps = side ? c->pre : st;
    if (side) {
        if (Ln.used)
            CTX_CUDA(c, cudaStreamWaitEvent(ps, Ln.freed, 0));
        if ((a->d_flux || a->d_depth || a->d_cheb_coef) && !a->device_inputs_ready) {
            CTX_CUDA(c, cudaEventRecord(c->inputs, st));
            CTX_CUDA(c, cudaStreamWaitEvent(ps, c->inputs, 0));
        }
    }
    CTX_CUDA(c, cudaStreamWaitEvent(ps, S.copied, 0));
    S.busy = true;
    auto dev = [&](size_t i) { return (void *)(S.d + pieces[i].off); };
    const double *d_wl = (const double *)dev(i_wl), *d_dur = (const double *)dev(i_dur);
    const double *d_flux = a->d_flux ? a->d_flux : (const double *)dev(i_flux);

    // ---- bins per CTA of the thrower (cached per set-up) ------------------------------------
    if (c->chunk_bins == 0 || c->chunk_W != W || c->chunk_N != N || c->chunk_wl0 != a->wl[0] ||
        c->chunk_wl1 != a->wl[W - 1]) {
        c->chunk_bins = choose_chunk_bins(I, *a);
        if (const char *env = getenv("WB200_CHUNK_BINS")) { // A/B switch (a multiple of 32)
            const int v = atoi(env) / 32 * 32;
            if (v >= 32)
                c->chunk_bins = v;
        }
        c->chunk_W = W;
        c->chunk_N = N;
        c->chunk_wl0 = a->wl[0];
        c->chunk_wl1 = a->wl[W - 1];
    }
    const int n_chunks = (W + c->chunk_bins - 1) / c->chunk_bins;
    CTX_STAGE(c, ctx_reserve(c, Ln.span, sizeof(int) * 2 * (size_t)n_chunks));

    // ---- stage 1: tables and traces ---------------------------------------------------
    double *tab = (double *)Ln.tables.d;
    double *d_ratio = tab, *d_sigl = tab + W, *d_sigh = tab + 2 * W, *d_sens = tab + 3 * W, *d_dwl = tab + 4 * W;
    mark_begin(c, 0, ps);
    CTX_STAGE(c, wb200_bin_tables(W, d_wl, I.psf_poly12, I.n_sens, (const double *)c->planes[WB200_PLANE_SENS_WL].d,
                                  (const double *)c->planes[WB200_PLANE_SENS_VAL].d, d_ratio, d_sigl, d_sigh,
                                  d_sens, d_dwl, ps));
    // first / last populated bin of every chunk: the thrower places its tiles from them
    k_chunk_spans<<<n_chunks, 256, 0, ps>>>(W, c->chunk_bins, d_wl, d_flux, d_sens, d_dwl, (int *)Ln.span.d);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CTX_CUDA(c, cudaGetLastError());
    CTX_STAGE(c, wb200_trace_table(N, (const double *)dev(i_x), (const double *)dev(i_y), I.trace_coeff9,
                                   I.wl_sol9, (double *)Ln.trace.d, ps));
    mark_end(c, ps);

    // ---- counts -------------------------------------------------------------------------
    wb200_counts_args ca;
    memset(&ca, 0, sizeof(ca));
    ca.n_samples = N;
    ca.n_bins = W;
    ca.count_mode = a->count_mode;
    ca.cheb_order = a->cheb_order;
    ca.key0 = a->key0;
    ca.key1 = a->key1;
    ca.scale = a->scale;
    ca.depth_ld = a->depth_ld;
    ca.d_flux = d_flux;
    ca.d_depth = a->cheb_order ? nullptr : (a->depth ? (const double *)S.big : a->d_depth);
    if (a->sep_row) {
        ca.d_depth = (const double *)dev(i_sc);
        ca.d_sep_row = (const double *)dev(i_sr);
    }
    if (a->cheb_order) {
        ca.d_cheb_coef = a->d_cheb_coef ? a->d_cheb_coef : (const double *)dev(i_cc);
        ca.d_cheb_x = (const double *)dev(i_cx);
    }
    ca.d_sens = d_sens;
    ca.d_dwl = d_dwl;
    ca.d_dur_ms = d_dur;
    ca.d_counts = (int32_t *)Ln.counts.d;
    ca.d_totals = (uint64_t *)Ln.totals.d;
    mark_begin(c, 1, ps);
    CTX_STAGE(c, launch_counts(&ca, ps));
    mark_end(c, ps);

    // ---- cosmic-ray chains (read by the per-pixel pass) ---------------------------------------
    if (a->n_cosmics) {
        mark_begin(c, 2, ps);
        CTX_STAGE(c, wb200_cosmic_chains(a->n_cosmics, (const int32_t *)dev(i_cp), (int32_t)plane,
                                         (int32_t *)Ln.cos_head.d, (int32_t *)Ln.cos_next.d, ps));
        mark_end(c, ps);
    }

    if (side) { // the caller's stream picks up when the lane's stage-1 work is complete
        CTX_CUDA(c, cudaEventRecord(Ln.pre_done, ps));
        CTX_CUDA(c, cudaStreamWaitEvent(st, Ln.pre_done, 0));
    }

    // ---- electrons: throw, bin, flat, accumulate ------------------------------------------
    CTX_CUDA(c, cudaMemsetAsync(c->tally.d, 0, sizeof(uint64_t) * 4, st));
    wb200_photon_args pa;
    memset(&pa, 0, sizeof(pa));
    pa.n_samples = N;
    pa.n_bins = W;
    pa.chunk_bins = c->chunk_bins;
    pa.nr = pa.nc = I.L;
    pa.rng_mode = WB200_RNG_PHILOX;
    pa.sub_scale = I.sub_scale;
    pa.key0 = a->key0;
    pa.key1 = a->key1;
    pa.d_counts = (const int32_t *)Ln.counts.d;
    pa.d_totals = (const uint64_t *)Ln.totals.d;
    pa.d_trace = (const double *)Ln.trace.d;
    pa.d_wl = d_wl;
    pa.d_ratio = d_ratio;
    pa.d_sigl = d_sigl;
    pa.d_sigh = d_sigh;
    pa.d_tally = (uint64_t *)c->tally.d;
    wb200_gather_args ga;
    memset(&ga, 0, sizeof(ga));
    ga.n_samples = N;
    ga.n_reads = R;
    ga.L = I.L;
    ga.F = F;
    ga.border = I.border;
    ga.add_flat = a->add_flat ? 1 : 0;
    ga.flat_off = I.flat_off;
    ga.flat_n = I.flat_n;
    ga.flat_f32 = I.flat_f32;
    ga.flat_planes_f32 = 1;
    ga.flat_wmin = I.flat_wmin;
    ga.flat_wmax = I.flat_wmax;
    ga.d_read_end = (const int32_t *)dev(i_re);
    ga.d_trace = (const double *)Ln.trace.d;
    for (int i = 0; i < 4; ++i)
        ga.d_flat[i] = (const double *)c->planes[WB200_PLANE_FLAT0 + i].d;
    ga.d_acc = (double *)c->acc.d;
    mark_begin(c, 3, st);
    CTX_STAGE(c, throw_photons(&pa, 0, st, &ga, 1, getenv("WB200_THROW_SCAN") ? nullptr : (const int *)Ln.span.d));
    mark_end(c, st);

    // ---- the per-pixel ramp pass ---------------------------------------------------------------
    wb200_reads_args ra;
    memset(&ra, 0, sizeof(ra));
    ra.n_reads = R;
    ra.F = F;
    ra.border = I.border;
    ra.out_f32 = a->out_f32 ? 1 : 0;
    ra.add_noise = (a->add_noise && a->noise_mean != 0.0 && a->noise_std != 0.0) ? 1 : 0;
    ra.add_sky = (a->add_sky && a->sky_rate != 0.0) ? 1 : 0;
    ra.add_dark = a->add_dark ? 1 : 0;
    ra.add_nonlinear = a->add_nonlinear ? 1 : 0;
    ra.clip = a->clip ? 1 : 0;
    ra.add_read_noise = (a->add_read_noise && I.read_noise != 0.0) ? 1 : 0;
    ra.n_cosmics = a->n_cosmics;
    ra.key0 = a->key0;
    ra.key1 = a->key1;
    ra.noise_mean = a->noise_mean;
    ra.noise_std = a->noise_std;
    ra.sky_rate = a->sky_rate;
    ra.sky_f32 = 1;
    ra.fast_math = 1;
    ra.acc_fixed = 1;
    ra.zero_acc = getenv("WB200_CTX_MEMSET_ACC") ? 0 : 1; // A/B switch: memset pass instead of the write-back
    ra.planes_f32 = 1;
    ra.const_gain = I.const_gain;
    ra.clip_lo = I.clip_lo;
    ra.clip_hi = I.clip_hi;
    ra.read_noise = I.read_noise;
    ra.d_dt = (const double *)dev(i_dt);
    ra.d_acc = c->acc.d;
    ra.d_sky = ra.add_sky ? c->planes[WB200_PLANE_SKY].d : nullptr;
    ra.d_gain = a->add_gain ? c->planes[WB200_PLANE_GAIN].d : nullptr;
    ra.d_zero = a->add_zero ? (const double *)c->planes[WB200_PLANE_ZERO].d : nullptr;
    if (a->add_dark) {
        ra.d_dark = c->planes[WB200_PLANE_DARK].d;
        ra.d_dark_err = c->planes[WB200_PLANE_DARK_ERR].d;
    }
    if (a->add_nonlinear)
        for (int i = 0; i < 4; ++i)
            ra.d_nl[i] = c->planes[WB200_PLANE_NL0 + i].d;
    if (a->n_cosmics) {
        ra.d_cos_head = (const int32_t *)Ln.cos_head.d;
        ra.d_cos_next = (const int32_t *)Ln.cos_next.d;
        ra.d_cos_read = (const int32_t *)dev(i_cr);
        ra.d_cos_energy = (const double *)dev(i_ce);
    }
    ra.d_out = d_out;
    mark_begin(c, 4, st);
    const int rc = launch_reads(&ra, st);
    mark_end(c, st);
    if (rc == WB200_OK && !ra.zero_acc)
        cudaMemsetAsync(c->acc.d, 0, sizeof(long long) * (size_t)R * plane, st);
    if (rc != WB200_OK) {
        // the interval planes may hold this exposure's electrons: restore the "zero between exposures" invariant
        cudaMemsetAsync(c->acc.d, 0, sizeof(long long) * (size_t)R * plane, st);
        snprintf(c->err, sizeof(c->err), "%s", g_err);
        return rc;
    }
    if (a->d_stats) {
        k_exposure_stats<<<1, 256, 0, st>>>(N, (const unsigned long long *)Ln.totals.d,
                                            (const unsigned long long *)c->tally.d, (unsigned long long *)a->d_stats);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CTX_CUDA(c, cudaGetLastError());
    }
    CTX_CUDA(c, cudaEventRecord(S.done, st));
    CTX_CUDA(c, cudaEventRecord(Ln.freed, st));
    Ln.used = true;
    c->exposures += 1;
    c->last_launches = (int64_t)(g_launches.load() - launches0);
    c->last_N = N;
    c->last_W = W;
    return WB200_OK;
}

int wb200_ctx_profile(wb200_ctx *c, int enable)
{
    WB_REQUIRE(c != nullptr, "null context");
    c->profile = enable != 0;
    return WB200_OK;
}

int wb200_ctx_stage_times(wb200_ctx *c, double ms_out[8], int64_t n_out[8])
{
    WB_REQUIRE(c != nullptr, "null context");
    CTX_REQUIRE(c, ms_out != nullptr && n_out != nullptr, "null output");
    for (int i = 0; i < 8; ++i) {
        ms_out[i] = 0.0;
        n_out[i] = 0;
    }
    int prev = 0;
    CTX_CUDA(c, cudaGetDevice(&prev));
    CTX_CUDA(c, cudaSetDevice(c->device));
    cudaError_t e = cudaDeviceSynchronize();
    for (auto &m : c->marks) {
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, m.a, m.b) == cudaSuccess && m.stage >= 0 && m.stage < 8) {
            ms_out[m.stage] += ms;
            n_out[m.stage] += 1;
        }
        cudaEventDestroy(m.a);
        cudaEventDestroy(m.b);
    }
    c->marks.clear();
    cudaSetDevice(prev);
    CTX_CUDA(c, e);
    return WB200_OK;
}

int wb200_ctx_info(const wb200_ctx *c, int64_t info[8])
{
    WB_REQUIRE(c != nullptr && info != nullptr, "null argument");
    int used = 0;
    for (const auto &s : c->slot)
        used += s.cap ? 1 : 0;
    const int64_t v[8] = {c->chunk_bins, used, c->exposures, c->last_launches, c->last_N, c->last_W, 0, 0};
    memcpy(info, v, sizeof(v));
    return WB200_OK;
}

int wb200_ctx_read_scratch(wb200_ctx *c, int which, void *host_out, int64_t bytes)
{
    WB_REQUIRE(c != nullptr, "null context");
    CTX_REQUIRE(c, host_out != nullptr && which >= 0 && which <= 3, "bad argument");
    const size_t N = (size_t)c->last_N, W = (size_t)c->last_W;
    const wb200_ctx::Lane &Ll = c->lane[c->last_lane]; // the lane of the last exposure
    const wb200_ctx::Buf *b[4] = {&Ll.counts, &Ll.totals, &Ll.trace, &Ll.tables};
    const size_t want[4] = {4 * N * W, 8 * N, 8 * WB200_TRACE_STRIDE * N, 8 * 5 * W};
    CTX_REQUIRE(c, b[which]->d && (size_t)bytes == want[which] && want[which] <= b[which]->cap,
                "size does not match the last exposure");
    int prev = 0;
    CTX_CUDA(c, cudaGetDevice(&prev));
    CTX_CUDA(c, cudaSetDevice(c->device));
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess)
        e = cudaMemcpy(host_out, b[which]->d, (size_t)bytes, cudaMemcpyDeviceToHost);
    cudaSetDevice(prev);
    CTX_CUDA(c, e);
    return WB200_OK;
}

} // extern "C"
