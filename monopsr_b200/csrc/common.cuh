// monopsr_b200/csrc/common.cuh -- shared host/device helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MPB_API extern "C" __attribute__((visibility("default")))

namespace mpb {

// Launch accounting: every kernel launch made by this library bumps this counter
// (read through mpb_launch_count(); bench.py reports it as "gpu_launches").
extern unsigned long long g_launch_count;
inline void count_launch(int k = 1) { __atomic_fetch_add(&g_launch_count, (unsigned long long)k, __ATOMIC_RELAXED); }

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? 0 : (int)e; }

#define MPB_CUDA_TRY(expr)                         \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

// Check the launch that was just issued.
#define MPB_LAUNCH_CHECK()                         \
    do {                                           \
        mpb::count_launch();                       \
        cudaError_t _e = cudaGetLastError();       \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// Stream-ordered scratch from a library-private pool (no effect on the process's default
// pool; memory is cached across calls, so an allocation costs ~1 us after the first use).
inline cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool;
        e = cudaMemPoolCreate(&pool, &props);
        if (e != cudaSuccess) return e;
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        pools[dev] = pool;
    }
    return cudaMallocFromPoolAsync(p, bytes, pools[dev], s);
}
inline cudaError_t scratch_free(void* p, cudaStream_t s) { return cudaFreeAsync(p, s); }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mpb
