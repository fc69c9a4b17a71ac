/*
 * exposure_host.c -- a plain-C host of libwayne_b200.so: one exposure through the
 * exposure-level interface of include/wayne_b200.h, no Python, no torch.
 *
 *     gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_host/exposure_host.c \
 *         -Lwayne_b200 -lwayne_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/wayne_b200 \
 *         -o exposure_host
 *     ./exposure_host bundle.bin reads.bin
 *
 * What the reference does in ExposureGenerator.scanning_frame (wayne/exposure_generator.py:
 * 178-405) plus _post_exposure_reductions (:407-444) -- thousands of Python -> C crossings
 * (wayne/pyparallel.pyx:27-30) and a dozen numpy passes per read -- is here:
 * wb200_ctx_create, wb200_ctx_set_instrument, wb200_ctx_upload_plane x n, wb200_exposure_run.
 *
 * The input bundle is a flat file of records written by tests/test_c_host_gpu.py:
 *     int32 tag, int32 dtype (0 f32, 1 f64, 2 i32, 3 raw bytes), int64 count, payload
 * tags 0..14 = the WB200_PLANE_* ids; 100 = wb200_instrument (raw); 101 = wb200_exposure_args
 * scalars (raw, pointers ignored); 110.. = the host arrays of the exposure (see below).
 * Output: the NSAMP reads, float64 [R+1][F][F], raw.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "wayne_b200.h"

enum { T_INST = 100, T_ARGS = 101, T_WL = 110, T_FLUX, T_XREF, T_YREF, T_DUR, T_DT, T_READ_END,
       T_SEP_ROW, T_SEP_COL, T_COS_PIXEL, T_COS_READ, T_COS_ENERGY };

typedef struct { int32_t tag, dtype; int64_t count; void *data; } record;

static size_t elem_size(int dtype) { return dtype == 0 ? 4 : dtype == 1 ? 8 : dtype == 2 ? 4 : 1; }

static const record *find(const record *r, int n, int tag)
{
    for (int i = 0; i < n; ++i)
        if (r[i].tag == tag)
            return &r[i];
    return NULL;
}

int main(int argc, char **argv)
{
    if (argc != 3) {
        fprintf(stderr, "usage: %s bundle.bin reads.bin\n", argv[0]);
        return 2;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) {
        perror(argv[1]);
        return 2;
    }
    record rec[64];
    int n = 0;
    while (n < 64) {
        int32_t head[2];
        int64_t count;
        if (fread(head, 4, 2, f) != 2 || fread(&count, 8, 1, f) != 1)
            break;
        rec[n].tag = head[0];
        rec[n].dtype = head[1];
        rec[n].count = count;
        rec[n].data = malloc((size_t)count * elem_size(head[1]) + 8);
        if (fread(rec[n].data, elem_size(head[1]), (size_t)count, f) != (size_t)count) {
            fprintf(stderr, "truncated bundle\n");
            return 2;
        }
        ++n;
    }
    fclose(f);

    wb200_ctx *ctx = NULL;
    if (wb200_ctx_create(0, &ctx) != WB200_OK) {
        fprintf(stderr, "wb200_ctx_create: %s\n", wb200_last_error());
        return 1;
    }
    const record *ri = find(rec, n, T_INST), *ra = find(rec, n, T_ARGS);
    if (!ri || !ra || (size_t)ri->count != sizeof(wb200_instrument) || (size_t)ra->count != sizeof(wb200_exposure_args)) {
        fprintf(stderr, "bundle lacks the instrument / argument records (or was written for another header)\n");
        return 2;
    }
    wb200_instrument inst;
    memcpy(&inst, ri->data, sizeof(inst));
    if (wb200_ctx_set_instrument(ctx, &inst) != WB200_OK) {
        fprintf(stderr, "set_instrument: %s\n", wb200_ctx_last_error(ctx));
        return 1;
    }
    for (int i = 0; i < n; ++i)
        if (rec[i].tag >= 0 && rec[i].tag < WB200_PLANE_COUNT)
            if (wb200_ctx_upload_plane(ctx, rec[i].tag, rec[i].data, rec[i].dtype == 1 ? WB200_F64 : WB200_F32,
                                       rec[i].count) != WB200_OK) {
                fprintf(stderr, "upload_plane %d: %s\n", rec[i].tag, wb200_ctx_last_error(ctx));
                return 1;
            }

    wb200_exposure_args a;
    memcpy(&a, ra->data, sizeof(a)); /* scalars; every pointer is set below */
#define PTR(field, tag, type)                                        \
    do {                                                             \
        const record *r_ = find(rec, n, tag);                        \
        a.field = r_ ? (const type *)r_->data : NULL;                \
    } while (0)
    PTR(wl, T_WL, double);
    PTR(flux, T_FLUX, double);
    PTR(xref, T_XREF, double);
    PTR(yref, T_YREF, double);
    PTR(dur_ms, T_DUR, double);
    PTR(dt_s, T_DT, double);
    PTR(read_end, T_READ_END, int32_t);
    PTR(sep_row, T_SEP_ROW, double);
    PTR(sep_col, T_SEP_COL, double);
    PTR(cos_pixel, T_COS_PIXEL, int32_t);
    PTR(cos_read, T_COS_READ, int32_t);
    PTR(cos_energy, T_COS_ENERGY, double);
    a.cheb_x = a.cheb_coef = a.depth = NULL;
    a.d_depth = a.d_cheb_coef = a.d_flux = NULL;
    a.cheb_order = 0;

    const size_t n_out = (size_t)(a.n_reads + 1) * inst.F * inst.F;
    double *d_out = NULL, *h_out = NULL;
    uint64_t *d_stats = NULL, stats[4];
    cudaStream_t st;
    if (cudaMalloc((void **)&d_out, n_out * 8) != cudaSuccess || cudaMalloc((void **)&d_stats, 32) != cudaSuccess ||
        cudaMallocHost((void **)&h_out, n_out * 8) != cudaSuccess || cudaStreamCreate(&st) != cudaSuccess) {
        fprintf(stderr, "cuda allocation failed\n");
        return 1;
    }
    a.d_stats = d_stats;
    /* twice: the second exposure finds the read-interval planes zeroed by the first one's ramp pass */
    for (int rep = 0; rep < 2; ++rep)
        if (wb200_exposure_run(ctx, &a, d_out, st) != WB200_OK) {
            fprintf(stderr, "wb200_exposure_run: %s\n", wb200_ctx_last_error(ctx));
            return 1;
        }
    cudaMemcpyAsync(h_out, d_out, n_out * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(stats, d_stats, 32, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        fprintf(stderr, "stream failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    FILE *o = fopen(argv[2], "wb");
    if (!o || fwrite(h_out, 8, n_out, o) != n_out) {
        perror(argv[2]);
        return 2;
    }
    fclose(o);
    printf("electrons thrown %llu, binned %llu, dropped %llu; %d reads of %d x %d written; %llu kernels launched\n",
           (unsigned long long)stats[0], (unsigned long long)stats[1], (unsigned long long)stats[2], a.n_reads + 1,
           inst.F, inst.F, (unsigned long long)wb200_launch_count());
    wb200_ctx_destroy(ctx);
    return 0;
}
