// common.cuh -- shared helpers for the sm_100a QuadConv kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
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

static inline int qc_num_sms()
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

__device__ __forceinline__ int qc_warp_max_i(int v)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
