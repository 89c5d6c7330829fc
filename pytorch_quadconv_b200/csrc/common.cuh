// common.cuh -- shared helpers for the sm_100a QuadConv kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include "../../include/quadconv_b200.h"

#define QC_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

#define QC_CUDA(x)                                         \
    do {                                                   \
        cudaError_t e__ = (x);                             \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

static inline int qc_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Batch lanes of the packed layout: Bp = min(32, pow2ceil(B)), nblk = ceil(B / 32).
static inline int qc_bp(int B)
{
    int bp = 1;
    while (bp < B && bp < 32) bp <<= 1;
    return bp;
}
static inline int qc_nblk(int B) { return (B + 31) / 32; }

// Packed feature layout  Xp[nblk][N][CP4/4][Bp][4]  (fp32): per point, channels in groups of four,
// batch lane next, the four channels of a group innermost -- so the thread that owns batch lane b
// reads four channels with one 128-bit load and a warp's load is one contiguous 512 B segment.
// CP4 = C rounded up to 4; padded channels and padded batch lanes hold zeros.
__host__ __device__ inline int qc_cp4(int C) { return (C + 3) & ~3; }
__host__ __device__ inline int qc_off(int c, int b, int Bp) { return ((((c >> 2) * Bp) + b) << 2) + (c & 3); }

// Per-process caches of launch-invariant host state (hoisted out of the per-launch path: a small-mesh layer
// step is ~25 launches of a few microseconds each).  All keyed by the CURRENT device, so one process may
// drive several GPUs.
constexpr int QC_MAX_DEVICES = 64;
static inline int qc_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= QC_MAX_DEVICES) ? 0 : dev;
}

static inline int qc_num_sms()
{
    static int sms[QC_MAX_DEVICES] = {0};
    const int dev = qc_device();
    if (sms[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = v > 0 ? v : 148;
    }
    return sms[dev];
}

// Environment switches are read once per process.
#define QC_ENV_SET(name) ([]() -> bool { static const bool v = getenv(name) != nullptr; return v; }())

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) and size increase.
template <auto Kernel>
static inline cudaError_t qc_set_max_smem(int bytes)
{
    static int cur[QC_MAX_DEVICES] = {0};
    const int dev = qc_device();
    if (bytes <= cur[dev]) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) cur[dev] = bytes;
    return e;
}

__device__ __forceinline__ int qc_warp_max_i(int v)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
