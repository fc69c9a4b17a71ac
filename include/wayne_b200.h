/*
 * wayne_b200.h -- C ABI of libwayne_b200.so: the B200 (sm_100a) replacement for
 * the hot path of ucl-exoplanets/wayne (per-exposure detector-image synthesis).
 *
 * Boundary rules: extern "C", plain pointers and sizes, no torch / C++ types.
 * Every `d_` / "device" pointer is a CUDA device pointer owned by the caller
 * (the Python host layer allocates them as torch tensors); every `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  All entry
 * points return WB200_OK (0) or a negative status; wb200_last_error() returns
 * the message of the last failure on the calling thread.  Nothing throws
 * across the boundary.  Launches are asynchronous on `stream` unless stated.
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * the reference tree).
 */
#ifndef WAYNE_B200_H
#define WAYNE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB200_OK 0
#define WB200_ERR_ARG -1
#define WB200_ERR_CUDA -2
#define WB200_ERR_NOMEM -3
#define WB200_ERR_LOST -4      /* an electron fell outside its HBM window */

/* How the electron normals are produced (a12 of the scope table). */
#define WB200_RNG_PHILOX 0     /* native: Philox4x32-10 counter streams, fp32 Box-Muller   */
#define WB200_RNG_RANDR 1      /* compat: glibc rand_r stream of the reference, fp64,       */
                               /*         chunked/seeded per OpenMP thread like PSF()       */
#define WB200_RNG_HOST 2       /* deterministic: normals supplied by the caller (A table)   */

/* How expected counts become integer counts (exposure_generator.py:625-628). */
#define WB200_COUNT_NONE 0     /* counts supplied by the caller                              */
#define WB200_COUNT_ROUND 1    /* np.round (half to even)  -- add_stellar_noise=False         */
#define WB200_COUNT_POISSON 2  /* Philox Poisson draw      -- add_stellar_noise=True, native */

const char *wb200_last_error(void);
int wb200_version(void);
/* Number of CUDA devices visible, or a negative status. */
int wb200_device_count(void);
/* Kernels launched by this library since load (all threads); the bench's
 * "gpu_launches" claim is read from here. */
uint64_t wb200_launch_count(void);

/* --------------------------------------------------------------------------
 * Drop-in for the reference's only native symbol
 *   int *PSF(int *counts,int size,double *x_pos,double *y_pos,double *psf_ratio,
 *            double *psf_sigmal,double *psf_sigmah,int nr,int nc,int test,int threads)
 * (wayne/pyparallel_menu.h:1-3, wayne/pyparallel_menu.c:10-113; bound by
 * wayne/pyparallel.pyx:10-12).  Same contract: HOST pointers in (borrowed,
 * never written), a malloc'd int[nr*nc] out that the CALLER frees, result
 * identical to the reference for the same (test, threads) -- the rand_r stream
 * and the per-thread chunking are reproduced on the GPU.  Returns NULL on a
 * CUDA failure (the reference cannot fail; see wb200_last_error()).
 * Synchronous.  Runs on the current CUDA device.
 * -------------------------------------------------------------------------- */
int *PSF(int *counts, int size, double *x_pos, double *y_pos, double *psf_ratio,
         double *psf_sigmal, double *psf_sigmah, int nr, int nc, int test,
         int threads);

/* Same computation, caller-owned output, explicit status, optional
 * caller-supplied normals (rng_mode WB200_RNG_HOST: normals = A[2*ssum] laid
 * out as in pyparallel_menu.c:61-62) or Philox (key = test).  HOST pointers. */
int wb200_psf_host(const int *counts, int size, const double *x_pos,
                   const double *y_pos, const double *psf_ratio,
                   const double *psf_sigmal, const double *psf_sigmah, int nr,
                   int nc, int test, int threads, int rng_mode,
                   const double *normals, int *frame_out);

/* --------------------------------------------------------------------------
 * Stage 1a: wavelength-only tables, one thread per bin.
 * Replaces G141.set_current_wavelength_only_dependent_array (wayne/grism.py:
 * 111-118: three np.poly1d evaluations + np.interp of the sensitivity table)
 * and tools.bin_centers_to_widths (wayne/tools.py:106-128).
 * d_wl [n_bins] microns.  psf_poly = ratio[4], sigmal[4], sigmah[4], highest
 * power first (grism.py:85-90).  d_sens_wl/d_sens_val [n_sens] (microns).
 * Outputs [n_bins] each: ratio, sigl, sigh, sens, dwl (microns).
 * -------------------------------------------------------------------------- */
int wb200_bin_tables(int n_bins, const double *d_wl, const double *psf_poly12,
                     int n_sens, const double *d_sens_wl,
                     const double *d_sens_val, double *d_ratio, double *d_sigl,
                     double *d_sigh, double *d_sens, double *d_dwl,
                     void *stream);

/* --------------------------------------------------------------------------
 * Stage 1b: field-dependent trace / dispersion per sub-sample, one thread per
 * sub-sample.  Replaces wavelength_calibration_coeffs (wayne/grism.py:779-803)
 * and _SpectrumTrace.__init__/_get_x_to_wl_poly_coeffs (grism.py:491-506,
 * 553-602).  d_xref/d_yref [n_samples] = the sub-sample reference position
 * (incl. jitter and scan offset).  trace_coeff9 / wl_sol9 = grism.py:756-776.
 * d_trace [n_samples][8] = {x_ref, y_ref, m_t, c_t, m_w, c_w, m_wl, c_wl}.
 * -------------------------------------------------------------------------- */
#define WB200_TRACE_STRIDE 8
int wb200_trace_table(int n_samples, const double *d_xref, const double *d_yref,
                      const double *trace_coeff9, const double *wl_sol9,
                      double *d_trace, void *stream);

/* Parity/debug helper: x_pos,y_pos [n_samples][n_bins] on the detector for
 * every (sub-sample, bin) = _SpectrumTrace.wl_to_x / wl_to_y (grism.py:635-669)
 * minus sub_scale (exposure_generator.py:630-632).  The photon kernel evaluates
 * the same device function inline. */
int wb200_trace_positions(int n_samples, int n_bins, const double *d_trace,
                          const double *d_wl, double sub_scale, double *d_xpos,
                          double *d_ypos, void *stream);

/* --------------------------------------------------------------------------
 * Stage 1c + Poisson: expected electrons per (sub-sample, bin) and their
 * integer draw.  Replaces _gen_subsample's flux->counts algebra
 * (wayne/exposure_generator.py:344-348, 602-628; _flux_to_counts :649-687).
 *   expected = flux[w]*(1-depth[s][w]) * sens[w] * dwl[w] * 1e4 * dur_ms[s]
 *              * 1e-3 * scale
 * d_depth may be NULL (no planet); depth_ld = row stride of d_depth.
 * d_expected (nullable) [n_samples][n_bins] float64 out.
 * d_counts  (nullable) [n_samples][n_bins] int32 out, per count_mode.
 * d_totals  [n_samples] uint64 out: electrons per sub-sample (zeroed here).
 * Philox key = (key0,key1); counter = (attempt, bin, sample, stream id).
 * -------------------------------------------------------------------------- */
int wb200_counts(int n_samples, int n_bins, const double *d_flux,
                 const double *d_depth, int64_t depth_ld, const double *d_sens,
                 const double *d_dwl, const double *d_dur_ms, double scale,
                 int count_mode, uint32_t key0, uint32_t key1,
                 double *d_expected, int32_t *d_counts, uint64_t *d_totals,
                 void *stream);

/* Same stage with the planet signal given as a per-sub-sample Chebyshev
 * expansion in the bin's radius ratio instead of an [n_samples][n_bins] array:
 *   depth[s][w] = sum_k coef[s][k] T_k(x[w])     (Clenshaw, numpy chebval order)
 * -- the light curves of Observation.generate_lightcurves (wayne/observation.py:
 * 293-357, 441-443) evaluated inside the counts kernel, so the 8 bytes per
 * (sub-sample, bin) planet-signal array is never built or copied.
 * d_cheb_coef [n_samples][cheb_order], d_cheb_x [n_bins]; cheb_order 1..32.
 * When d_cheb_coef is NULL, d_depth / depth_ld are used as in wb200_counts. */
typedef struct wb200_counts_args {
    int32_t n_samples, n_bins, count_mode, cheb_order;
    uint32_t key0, key1;
    double scale;
    int64_t depth_ld;
    const double *d_flux;
    const double *d_depth;
    const double *d_cheb_coef;
    const double *d_cheb_x;
    const double *d_sens;
    const double *d_dwl;
    const double *d_dur_ms;
    double *d_expected;
    int32_t *d_counts;
    uint64_t *d_totals;
    const double *d_sep_row;   /* nullable [n_samples]: SEPARABLE planet signal          */
                               /*   depth[s][w] = d_sep_row[s] * d_depth[w]             */
                               /* (d_depth then holds the per-bin factor, depth_ld      */
                               /* unused): a transit of one light-curve shape times a   */
                               /* depth spectrum, bit-identical to the dense product    */
} wb200_counts_args;

int wb200_counts_ex(const wb200_counts_args *args, void *stream);

/* Light curves on the device: the Chebyshev planet-signal
 * coefficients wb200_counts_ex consumes, from the projected star-planet
 * separations of the sub-samples.  Replaces the per-wavelength pylightcurve calls
 * of Observation.generate_lightcurves (wayne/observation.py:293-357).
 *   depth(s, p) = 1 - F(z[s], p): Claret four-coefficient limb darkening, occulted
 *   intensity by Gauss-Legendre quadrature over stellar annuli (n_gl nodes,
 *   d_gl_x / d_gl_w on [-1, 1]), evaluated at the `order` Chebyshev nodes of the
 *   radius-ratio interval [p_min, p_max], then transformed to coefficients.
 * d_z [n_samples] (use a value >= 1 + p_max, e.g. 1e30, when the planet is behind
 * the star or out of transit); ld4 HOST array; d_coef [n_samples][order] out. */
int wb200_transit_cheb(int n_samples, int order, const double *d_z, double p_min, double p_max,
                       const double *ld4, int n_gl, const double *d_gl_x, const double *d_gl_w,
                       double *d_coef, void *stream);

/* Exclusive prefix of counts along bins, per sub-sample (electron offsets of
 * the compat / deterministic modes = the reference's running electron_counter,
 * pyparallel_menu.c:86-107).  d_offsets [n_samples][n_bins] int32. */
int wb200_count_offsets(int n_samples, int n_bins, const int32_t *d_counts,
                        int32_t *d_offsets, void *stream);

/* --------------------------------------------------------------------------
 * Stages 2+3: throw every electron and bin it.  Replaces PSF()'s normal table
 * and scatter loop (wayne/pyparallel_menu.c:40-64, 87-108) for ALL sub-samples
 * of an exposure in one launch: grid = (bin chunks, sub-samples); each CTA
 * histograms into a shared-memory tile and flushes it with integer atomics
 * into its sub-sample's HBM window win[s][WH][WW] (origin win_ox/oy[s], frame
 * coordinates).  Windows must be zeroed by the caller.
 * -------------------------------------------------------------------------- */
typedef struct wb200_photon_args {
    int32_t n_samples;        /* sub-samples in this launch                       */
    int32_t n_bins;           /* W                                                */
    int32_t chunk_bins;       /* bins per CTA (multiple of 32)                    */
    int32_t nr, nc;           /* bounds test 0<x<nr, 0<y<nc (pyparallel_menu.c:93) */
    int32_t rng_mode;         /* WB200_RNG_*                                      */
    int32_t threads;          /* RANDR: the reference's OpenMP thread count       */
    int32_t win_w, win_h;     /* WW, WH                                           */
    double sub_scale;         /* 507 - SUBARRAY/2 (exposure_generator.py:630)     */
    uint32_t key0, key1;      /* PHILOX key                                       */
    const int32_t *d_counts;  /* [n_samples][n_bins]                              */
    const int32_t *d_offsets; /* [n_samples][n_bins] (RANDR / HOST), else NULL    */
    const uint64_t *d_totals; /* [n_samples] electrons per sub-sample             */
    const double *d_xpos;     /* explicit positions [n_samples][n_bins] or NULL   */
    const double *d_ypos;
    const double *d_trace;    /* [n_samples][8] when positions are NULL           */
    const double *d_wl;       /* [n_bins] microns                                 */
    const double *d_ratio;    /* [n_bins]                                         */
    const double *d_sigl;
    const double *d_sigh;
    const int32_t *d_seeds;   /* RANDR: `test` per sub-sample                     */
    const double *d_normals;  /* HOST: concatenated A tables                      */
    const int64_t *d_normals_base; /* HOST: offset of sub-sample s's A table      */
    int32_t *d_win;           /* [n_samples][win_h][win_w] int32, pre-zeroed      */
    const int32_t *d_win_ox;  /* [n_samples] window origin x (frame coords)       */
    const int32_t *d_win_oy;
    uint64_t *d_lost;         /* [1] electrons inside the frame but outside their */
                              /* window (pre-zeroed; must stay 0)                 */
    uint64_t *d_tally;        /* direct accumulation only, nullable, ADDED to:    */
                              /* [0] electrons binned inside the frame, [1] the   */
                              /* ones dropped outside it (pyparallel_menu.c:93),  */
                              /* so [0] + [1] == sum(counts) of the sub-samples   */
                              /* that belong to a read                            */
} wb200_photon_args;

int wb200_throw_photons(const wb200_photon_args *args, void *stream);
/* Same, for a batch of an exposure: sample0 = exposure-wide index of the batch's
 * first sub-sample (Philox counters use exposure-wide indices, so the result
 * does not depend on how an exposure is batched). */
int wb200_throw_photons_at(const wb200_photon_args *args, int sample0, void *stream);

/* --------------------------------------------------------------------------
 * Stage 3b: flat field + accumulation into the per-read-interval planes, in
 * sub-sample order per pixel (deterministic).  Replaces G141.get_flat_field's
 * indices branch (wayne/grism.py:359-385), `new_pixel_array *= flat_field`
 * (exposure_generator.py:641-645) and `pixel_array += sample_frame` (:359).
 * d_acc [n_reads][F][F] float64, bordered layout (light-sensitive pixel (r,c)
 * lives at (r+border, c+border)); accumulated INTO (zero it first).
 * -------------------------------------------------------------------------- */
typedef struct wb200_gather_args {
    int32_t n_samples;          /* sub-samples in the window buffer            */
    int32_t sample0;            /* global index of the first of them           */
    int32_t n_reads;            /* R                                           */
    int32_t L, F, border;       /* geometry                                    */
    int32_t win_w, win_h;
    int32_t add_flat;
    int32_t flat_off;           /* (1014 - SUBARRAY) floor-div 2 (grism.py:361) */
    int32_t flat_n;             /* side of the flat planes (1014)              */
    int32_t flat_f32;           /* 1: round the flat value to float32 before it */
                                /* multiplies -- the reference stores it in    */
                                /* np.ones_like(flat_f0), the FITS float32     */
                                /* dtype (grism.py:380-385)                    */
    int32_t exact;              /* 1: the reference's flat expression operation */
                                /* for operation (fp64 divide + sqrt; parity   */
                                /* mode); 0: hoisted reciprocals + FMAs (native)*/
    int32_t flat_planes_f32;    /* direct accumulation only: d_flat[] point at float32 */
                                /* planes (the file's dtype) instead of float64     */
    double flat_wmin, flat_wmax;
    const int32_t *d_read_end;  /* [R] global index of each read's last sample */
    const int32_t *d_win;
    const int32_t *d_win_ox;    /* indexed by local sample                     */
    const int32_t *d_win_oy;
    const double *d_trace;      /* indexed by local sample                     */
    const double *d_flat[4];    /* f0..f3 [flat_n][flat_n]                     */
    double *d_acc;
} wb200_gather_args;

int wb200_gather_flat(const wb200_gather_args *args, void *stream);

/* Native mode, stages 2+3+3b in ONE launch: throw (Philox), bin in shared-memory
 * tiles, and at tile flush apply the sub-sample's flat and add count*flat into
 * the read interval's plane -- no per-sub-sample HBM windows, no gather pass.
 * `g` supplies geometry, flat planes, d_read_end, d_trace (win_* fields unused);
 * g->d_acc is [n_reads][F][F] INT64 fixed point, 2^24 per electron (integer
 * atomics commute: bit-reproducible), zeroed by the caller; wb200_reads reads it
 * with acc_fixed = 1.  args->d_win* / d_lost are unused. */
int wb200_throw_photons_direct(const wb200_photon_args *args, const wb200_gather_args *g,
                               int sample0, void *stream);

/* --------------------------------------------------------------------------
 * Stage 4: the fused per-pixel pass over all reads.  Replaces
 * _add_read_reductions (wayne/exposure_generator.py:468-515), the cumulative
 * read stack (:378-382), _post_exposure_reductions (:407-444) and the
 * Exposure/WFC3_IR methods it calls (wayne/exposure.py:49-131,
 * wayne/detector.py:151-198, 318-350).  One thread owns a pixel for the whole
 * ramp, so every calibration plane is read once.  All planes are [F][F]
 * float64 in the bordered layout; nullable planes switch their term off.
 * Draw planes (compat mode) carry host-generated numpy draws; when NULL and
 * the term is enabled the kernel draws from Philox.
 * d_out [n_reads+1][F][F]: read 0 is the zero read.
 * -------------------------------------------------------------------------- */
typedef struct wb200_reads_args {
    int32_t n_reads;            /* R = NSAMP-1                                  */
    int32_t F, border;
    int32_t out_f32;            /* 0: float64 output, 1: float32 output         */
    int32_t add_noise, add_sky, add_dark, add_nonlinear, clip, add_read_noise;
    int32_t exact_newton;       /* 1: reference's global stopping rule          */
    int32_t n_cosmics;          /* length of the hit list                       */
    uint32_t key0, key1;
    double noise_mean, noise_std;   /* per second (exposure_generator.py:477-479) */
    double sky_rate;            /* counts/s                                     */
    int32_t sky_f32;            /* 1: Poisson mean = float32(sky)*float32(rate*dt), */
                                /* the in-place float32 product of :489-493     */
    int32_t fast_math;          /* 1 (native mode): fp32-SFU normals, reciprocal */
                                /* gain; 0: fp64 expressions of the reference    */
    int32_t acc_fixed;          /* 1: d_acc is int64 fixed point (2^24 / electron) */
    int32_t zero_acc;           /* 1 (native kernel, acc_fixed): every interval-plane  */
                                /* element is written back as 0 once it has been read, */
                                /* so the next exposure needs no memset pass           */
    double const_gain;          /* 2.35, used when d_gain == NULL               */
    double clip_lo, clip_hi;    /* -20, 78000                                   */
    double read_noise;          /* 14.1/2.35                                    */
    const double *d_dt;         /* [R] read interval lengths, seconds           */
    const void *d_acc;          /* [R][F][F] electrons per interval: float64, or */
                                /* int64 fixed point when acc_fixed             */
    const void *d_sky;          /* master sky plane (float64; float32 when planes_f32) */
    const void *d_gain;         /* 2.35/pfl plane or NULL            (same)     */
    const double *d_zero;       /* zero read (initial bias) or NULL (= zeros)   */
    const void *d_dark;         /* [R][F][F] or NULL                 (same)     */
    const void *d_dark_err;     /* [R][F][F]                         (same)     */
    const void *d_nl[7];        /* 1+c1, c2, c3, c4, 2*c2, 3*c3, 4*c4 (same)    */
    const double *d_draw_noise; /* [R][F][F] compat: N(mu*dt, sd*dt) draws      */
    const double *d_draw_sky;   /* [R][F][F] compat: Poisson draws              */
    const double *d_draw_dark;  /* [R][F][F] compat: N(dark, err) draws         */
    const double *d_draw_rn;    /* [R+1][F][F] compat: standard normals         */
    const int32_t *d_cos_head;  /* [F][F] first hit of the pixel or -1 (NULL=off) */
    const int32_t *d_cos_next;  /* [n_cosmics]                                  */
    const int32_t *d_cos_read;  /* [n_cosmics] read interval of the hit         */
    const double *d_cos_energy; /* [n_cosmics]                                  */
    int32_t *d_newton_iters;    /* [R] scratch for exact_newton (zeroed here)   */
    void *d_out;
    int32_t planes_f32;         /* 1 (native kernel only): sky, gain, nl, dark and dark  */
                                /* error planes are float32 -- the dtype of the reference's */
                                /* calibration FITS files (grism.py:66-76, detector.py:56-57, */
                                /* 183-190) -- promoted in registers                     */
    int32_t pad2;
} wb200_reads_args;

int wb200_reads(const wb200_reads_args *args, void *stream);

/* Cosmic-ray hit list -> per-pixel chains (cosmic_rays.py:70-86 scatter).
 * d_head [F*F] is set to -1 and filled; d_next [n] out. */
int wb200_cosmic_chains(int n_hits, const int32_t *d_pixel, int32_t n_pixels,
                        int32_t *d_head, int32_t *d_next, void *stream);

/* --------------------------------------------------------------------------
 * Exposure-level interface: ONE call per exposure.
 *
 * The reference's boundary between Python and native code is one call per
 * sub-sample (wayne/pyparallel.pyx:27-30, called from wayne/exposure_generator.py:
 * 636-639 inside the loop :336-394).  Here the whole native-mode exposure --
 * stage 1 tables and traces, Philox counts, the electron throw with the flat
 * fused into the tile flush, the per-pixel ramp pass (:178-405 and :407-444) --
 * is one call on an opaque per-GPU context that owns the resident calibration
 * planes, the scratch buffers and the staging of the small per-exposure inputs.
 * A C / C++ host needs nothing but this section.
 *
 * Rules: one context per (GPU, instrument configuration); calls on ONE context
 * must be serialised by the caller (distinct contexts may be used from distinct
 * threads; ctypes releases the GIL).  Every call returns WB200_OK or a negative
 * status; wb200_ctx_last_error(ctx) (or wb200_last_error() on the calling
 * thread) has the message.  Nothing throws across the boundary.
 * wb200_exposure_run is asynchronous on `stream`: it returns when the work is
 * queued.  Host arrays passed to it are copied before it returns and may be
 * reused immediately.  The output (and d_stats) is complete in `stream` order.
 * Consecutive exposures overlap on the device: the tables, traces and counts of
 * an exposure (stage 1) are made on a stream of the context's own, into the
 * other of two scratch lanes, while the previous exposure's electrons are still
 * being thrown on `stream` -- its electron throw and ramp pass then follow on
 * `stream` as before.  Results do not depend on it (same counters, same
 * kernels); WB200_CTX_SERIAL=1 in the environment at context creation, or
 * per-stage timing (wb200_ctx_profile), keeps everything on `stream`.
 * -------------------------------------------------------------------------- */
typedef struct wb200_ctx wb200_ctx;

typedef struct wb200_instrument {
    int32_t subarray;           /* SUBARRAY                                          */
    int32_t L, F, border;       /* light-sensitive side, full side (detector.py:102-124), 5 */
    int32_t flat_off;           /* (1014 - SUBARRAY) floor-div 2 (grism.py:361-363)  */
    int32_t flat_n;             /* side of the flat-field planes                     */
    int32_t flat_f32;           /* 1: flat value rounded to float32 (grism.py:380-385) */
    int32_t n_sens;             /* length of the sensitivity table                   */
    double sub_scale;           /* 507 - SUBARRAY//2 (exposure_generator.py:630)     */
    double flat_wmin, flat_wmax;
    double psf_poly12[12];      /* grism.py:85-90: ratio, sigma_l, sigma_h (highest power first) */
    double trace_coeff9[9];     /* grism.py:756-776                                  */
    double wl_sol9[9];
    double const_gain;          /* 2.35 (detector.py:29)                             */
    double clip_lo, clip_hi;    /* -20, 78000 (detector.py:26-27)                    */
    double read_noise;          /* 14.1 / 2.35 (detector.py:33)                      */
} wb200_instrument;

/* Resident calibration planes.  Image planes are held in float32 -- the dtype of
 * the reference's calibration FITS files (grism.py:66-76, 411-423; detector.py:
 * 56-57, 183-190, 200-209) -- in the bordered F x F layout (light-sensitive pixel
 * (r, c) at (r + border, c + border)); float64 input is accepted when every value
 * is float32-representable (anything else would change results and is refused).
 * ZERO (the initial bias, a float64 file) and the sensitivity table stay float64. */
#define WB200_PLANE_FLAT0 0     /* f0..f3 [flat_n][flat_n]                           */
#define WB200_PLANE_FLAT1 1
#define WB200_PLANE_FLAT2 2
#define WB200_PLANE_FLAT3 3
#define WB200_PLANE_SKY 4       /* master sky [F][F]                                 */
#define WB200_PLANE_GAIN 5      /* 2.35 / pfl [F][F], 1 in the border                */
#define WB200_PLANE_NL0 6       /* 1 + c1, c2, c3, c4 [F][F] (detector.py:328-343)   */
#define WB200_PLANE_NL1 7
#define WB200_PLANE_NL2 8
#define WB200_PLANE_NL3 9
#define WB200_PLANE_DARK 10     /* [n_reads][F][F]: read r uses super-dark NSAMP r+2 */
#define WB200_PLANE_DARK_ERR 11 /* [n_reads][F][F], non-positive entries already 1e-5 */
#define WB200_PLANE_ZERO 12     /* zero read / initial bias [F][F] float64           */
#define WB200_PLANE_SENS_WL 13  /* [n_sens] microns, float64                         */
#define WB200_PLANE_SENS_VAL 14 /* [n_sens] float64                                  */
#define WB200_PLANE_COUNT 15
#define WB200_F32 0
#define WB200_F64 1

int wb200_ctx_create(int device, wb200_ctx **ctx_out);
int wb200_ctx_destroy(wb200_ctx *ctx);
const char *wb200_ctx_last_error(const wb200_ctx *ctx);
int wb200_ctx_set_instrument(wb200_ctx *ctx, const wb200_instrument *inst);
/* host -> device, synchronous; replaces a plane uploaded before.  host = NULL drops it. */
int wb200_ctx_upload_plane(wb200_ctx *ctx, int which, const void *host, int dtype, int64_t count);

typedef struct wb200_exposure_args {
    int32_t n_samples, n_bins, n_reads;   /* N, W, R = NSAMP - 1                     */
    int32_t count_mode;         /* WB200_COUNT_ROUND | WB200_COUNT_POISSON (exposure_generator.py:625-628) */
    int32_t cheb_order;         /* > 0: planet signal as Chebyshev coefficients (see wb200_counts_ex) */
    int32_t n_cosmics;
    int32_t add_flat, add_sky, add_gain, add_dark, add_nonlinear, clip, add_read_noise,
        add_zero, add_noise;    /* the switches of scanning_frame (exposure_generator.py:178-192) */
    int32_t out_f32;            /* 0: float64 reads like the reference, 1: float32   */
    uint32_t key0, key1;        /* Philox key of the exposure                        */
    int32_t device_inputs_ready; /* 1: d_flux / d_depth / d_cheb_coef (if any) are complete already --   */
                                /* nothing queued on `stream` produces them -- so the library may     */
                                /* start this exposure's tables and counts on its own stream while the */
                                /* previous exposure is still running; 0: ordered after `stream`       */
    double scale;               /* visit-trend scale factor (:620-621)               */
    double sky_rate;            /* counts/s                                          */
    double noise_mean, noise_std;
    int64_t depth_ld;           /* row stride of d_depth                             */
    /* small HOST arrays (copied before the call returns) */
    const double *wl;           /* [W] microns, cropped to the grism limits          */
    const double *flux;         /* [W] or NULL when d_flux is given                  */
    const double *xref, *yref;  /* [N] sub-sample reference positions (jitter, scan) */
    const double *dur_ms;       /* [N] sub-sample durations                          */
    const double *dt_s;         /* [R] read interval lengths                         */
    const int32_t *read_end;    /* [R] index of the last sub-sample of each read     */
    const double *cheb_x;       /* [W] or NULL                                       */
    const double *cheb_coef;    /* HOST [N][cheb_order], or NULL when d_cheb_coef    */
    const double *sep_row;      /* separable planet signal (see wb200_counts_args):  */
    const double *sep_col;      /* HOST [N] and [W]; both NULL otherwise             */
    const double *depth;        /* HOST dense planet signal [N][depth_ld] (first used */
                                /* column), or NULL: uploaded by the library on its   */
                                /* upload stream, right behind the small arrays, into */
                                /* a ring of device buffers (pinned memory makes the  */
                                /* copy asynchronous; pageable memory works, slower)  */
    const int32_t *cos_pixel;   /* [n_cosmics] bordered flat pixel index             */
    const int32_t *cos_read;    /* [n_cosmics] read interval                         */
    const double *cos_energy;   /* [n_cosmics] electrons                             */
    /* large inputs already on the device (nullable) */
    const double *d_depth;      /* planet signal [N][depth_ld] (first used column)   */
    const double *d_cheb_coef;  /* [N][cheb_order]                                   */
    const double *d_flux;       /* [W]                                               */
    uint64_t *d_stats;          /* nullable device [4], written: electrons thrown,   */
                                /* binned inside the frame, dropped outside it, 0    */
} wb200_exposure_args;

/* d_out: device [n_reads+1][F][F] float64 (float32 when out_f32), read 0 = zero read. */
int wb200_exposure_run(wb200_ctx *ctx, const wb200_exposure_args *args, void *d_out, void *stream);

/* Diagnostics / parity tests: geometry chosen for the last exposure and synchronous
 * copies of the context's scratch buffers to the host.
 *   info[0] = bins per CTA of the thrower, [1] = staging slots in use, [2] = exposures run,
 *   [3] = kernels launched by the last exposure.
 *   which: 0 counts int32 [N][W]; 1 totals uint64 [N]; 2 trace float64 [N][8];
 *          3 tables float64 [5][W] (ratio, sigma_l, sigma_h, sens, dwl) */
int wb200_ctx_info(const wb200_ctx *ctx, int64_t info[8]);
/* Per-stage device timing for the bench: with profiling on, wb200_exposure_run brackets its
 * stages with CUDA events on the launching stream (no synchronisation).  wb200_ctx_stage_times
 * synchronises the device, returns the accumulated milliseconds and launch-group counts since
 * the last call and resets them.  Stages: 0 tables + traces, 1 counts, 2 cosmic chains,
 * 3 electron throw (+ flat + accumulation), 4 per-pixel ramp pass. */
int wb200_ctx_profile(wb200_ctx *ctx, int enable);
int wb200_ctx_stage_times(wb200_ctx *ctx, double ms_out[8], int64_t n_out[8]);
int wb200_ctx_read_scratch(wb200_ctx *ctx, int which, void *host_out, int64_t bytes);

/* Microbenchmarks used for the roofline denominators (profiles/).  which:
 *   0..3  shared-memory atomics (same address / conflict-free / PSF-like 3x3 / random)
 *   4, 5  global red (spread over 64 MB / hot 256 KB window)
 *   6     Philox4x32-10 calls of the thrower's fixed-key stream alone
 *   7     the thrower's whole random recipe (one call + four fp32 SFU Box-Muller pairs)
 *   8..13 single-instruction issue-rate probes: IMAD.WIDE.U32, IMAD, LOP3, MUFU.LG2,
 *         FFMA, I2FP
 * Returns the best elapsed ms of 3 timed launches in *ms_out and the operations per
 * launch in *ops_out.  Synchronous. */
int wb200_microbench(int which, int iters, double *ms_out, double *ops_out);

/* Test hook: raw words of the library's counter-based random streams, computed on the
 * device (host pointers in and out; synchronous).
 *   which = 0: Philox4x32-10 (Salmon et al., SC'11) of counter c[4] under key k[2], the
 *              function behind the count / per-pixel samplers: the Random123 known-answer
 *              vectors apply (tests/test_rng_gpu.py);
 *   which = 1: the native thrower's call with first counter word c[0] (the thrower passes 4 x the
 *              unit index = the index of the unit's first electron) for bin c[2] in sub-sample c[1]
 *              of the exposure keyed k[2], stream id c[3] (2 electrons, 7 their tail refinement):
 *              fixed-key Philox4x32-10 over (c[0], hy + sub-sample, bin, hw ^ stream) with
 *              (hy, hw) = splitmix64 of the key; out[4], out[5] return hy, hw. */
int wb200_philox_words(int which, const uint32_t *c, const uint32_t *k, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* WAYNE_B200_H */
