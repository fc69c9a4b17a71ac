// common.cuh -- error plumbing, launch accounting and warp helpers shared by
// every kernel file of libwayne_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/wayne_b200.h"

namespace wb {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define WB_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess)                                                \
            return wb::fail(WB200_ERR_CUDA, "%s: %s", #call,                   \
                            cudaGetErrorString(e__));                          \
    } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define WB_LAUNCHED(name)                                                      \
    do {                                                                       \
        wb::g_launches.fetch_add(1, std::memory_order_relaxed);                \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess)                                                \
            return wb::fail(WB200_ERR_CUDA, "launch %s: %s", name,             \
                            cudaGetErrorString(e__));                          \
    } while (0)

// Device-side index checks of the bounds-checked build (-DWB_BOUNDS_CHECK, see
// tools/bounds_check_build.py: compute-sanitizer is not available on the GPU pool).  A failed
// check traps, which the next CUDA call reports as an error, so a test run under that build
// fails loudly.  Compiled out of the product build.
#ifdef WB_BOUNDS_CHECK
#define WB_DEV_ASSERT(cond)                                                    \
    do {                                                                       \
        if (!(cond))                                                           \
            __trap();                                                          \
    } while (0)
#else
#define WB_DEV_ASSERT(cond) ((void)0)
#endif

#define WB_REQUIRE(cond, msg)                                                  \
    do {                                                                       \
        if (!(cond))                                                           \
            return wb::fail(WB200_ERR_ARG, "%s (%s)", msg, #cond);             \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(FULL, v, d);
        if (lane >= d)
            v += n;
    }
    return v;
}

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1)
        v = fminf(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1)
        v = fmaxf(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1)
        v += __shfl_xor_sync(FULL, v, d);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1)
        v = max(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
__device__ __forceinline__ double shfl_d(double v, int src)
{
    return __shfl_sync(FULL, v, src);
}

// 128-bit streaming accesses for planes that are touched once per exposure
__device__ __forceinline__ double2 ld_stream2(const double *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream2(double *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p),
                 "d"(v.x), "d"(v.y)
                 : "memory");
}
__device__ __forceinline__ void st_stream4f(float *p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

} // namespace wb
