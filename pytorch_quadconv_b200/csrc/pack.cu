// pack.cu -- layout conversion between the reference's [B,C,N] activations
// (torch_quadconv/quadconv.py:238 docstring, point index innermost) and the packed
// [nblk][N][C][Bp] layout (batch innermost) the fused kernels gather from, plus the small weight
// repacks of the last filter-MLP layer.  Pure HBM-bound copies: every global access is a full
// 32-float (128 B) segment on one side and a Bp-float segment on the other, staged through a
// padded shared-memory tile.
#include <cstdint>
#include "common.cuh"

namespace {

// grid: (ceil(N/64), CP4/4, nblk); block 256.  Tile: 64 points x 4 channels x Bp batch lanes, held as
// 4*Bp rows of 64 points with an odd row stride (65): the [B,C,N] side moves whole 256-byte row pieces
// as float4 per lane (VEC: every row start is 16-byte aligned, i.e. N % 4 == 0), the packed side moves
// 16 bytes per lane = the 4 channels of one (point, batch lane), 512 contiguous bytes per warp.
constexpr int PT = 64, RS = 65;

template <bool VEC>
__global__ void __launch_bounds__(256)
k_pack(const float* __restrict__ x, int B, int C, int N, int Bp, float* __restrict__ xp)
{
    __shared__ float tile[4 * 32 * RS];   // [(channel in group) * Bp + batch lane][point]
    const int n0 = blockIdx.x * PT, cg = blockIdx.y, blk = blockIdx.z;
    const int CG = gridDim.y;
    const int rows = 4 * Bp;
    if (VEC) {
        const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;   // half-warp per row, float4 per lane
        const int n = n0 + hl * 4;
        for (int r = hw; r < rows; r += 16) {
            const int e = r / Bp, bl = r % Bp;
            const int c = cg * 4 + e, b = blk * 32 + bl;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C && b < B && n < N) v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)b * C + c) * N + n));
            float* t = tile + r * RS + hl * 4;
            t[0] = v.x, t[1] = v.y, t[2] = v.z, t[3] = v.w;
        }
    } else {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = w; r < rows; r += 8) {
            const int e = r / Bp, bl = r % Bp;
            const int c = cg * 4 + e, b = blk * 32 + bl;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int n = n0 + half * 32 + lane;
                float v = 0.f;
                if (c < C && b < B && n < N) v = x[((size_t)b * C + c) * N + n];
                tile[r * RS + half * 32 + lane] = v;
            }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < PT * Bp; idx += 256) {
        const int bl = idx % Bp, nl = idx / Bp;
        const int n = n0 + nl;
        if (n < N) {
            const float4 v = make_float4(tile[(0 * Bp + bl) * RS + nl], tile[(1 * Bp + bl) * RS + nl],
                                         tile[(2 * Bp + bl) * RS + nl], tile[(3 * Bp + bl) * RS + nl]);
            *reinterpret_cast<float4*>(xp + ((((size_t)blk * N + n) * CG + cg) * Bp + bl) * 4) = v;
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
k_unpack(const float* __restrict__ xp, int B, int C, int N, int Bp, const float* __restrict__ bias,
         float* __restrict__ x)
{
    __shared__ float tile[4 * 32 * RS];
    const int n0 = blockIdx.x * PT, cg = blockIdx.y, blk = blockIdx.z;
    const int CG = gridDim.y;
    const int rows = 4 * Bp;
    for (int idx = threadIdx.x; idx < PT * Bp; idx += 256) {
        const int bl = idx % Bp, nl = idx / Bp;
        const int n = n0 + nl;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N) v = __ldg(reinterpret_cast<const float4*>(xp + ((((size_t)blk * N + n) * CG + cg) * Bp + bl) * 4));
        tile[(0 * Bp + bl) * RS + nl] = v.x;
        tile[(1 * Bp + bl) * RS + nl] = v.y;
        tile[(2 * Bp + bl) * RS + nl] = v.z;
        tile[(3 * Bp + bl) * RS + nl] = v.w;
    }
    __syncthreads();
    if (VEC) {
        const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
        const int n = n0 + hl * 4;
        for (int r = hw; r < rows; r += 16) {
            const int e = r / Bp, bl = r % Bp;
            const int c = cg * 4 + e, b = blk * 32 + bl;
            if (c < C && b < B && n < N) {
                const float* t = tile + r * RS + hl * 4;
                float4 v = make_float4(t[0], t[1], t[2], t[3]);
                if (bias != nullptr) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + (size_t)c * N + n));
                    v.x += bv.x, v.y += bv.y, v.z += bv.z, v.w += bv.w;
                }
                *reinterpret_cast<float4*>(x + ((size_t)b * C + c) * N + n) = v;
            }
        }
    } else {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = w; r < rows; r += 8) {
            const int e = r / Bp, bl = r % Bp;
            const int c = cg * 4 + e, b = blk * 32 + bl;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int n = n0 + half * 32 + lane;
                if (c < C && b < B && n < N) {
                    const float bv = bias != nullptr ? bias[(size_t)c * N + n] : 0.f;
                    x[((size_t)b * C + c) * N + n] = tile[r * RS + half * 32 + lane] + bv;
                }
            }
        }
    }
}

__global__ void k_batch_sum(const float* __restrict__ x, int B, int CN, float* __restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CN; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += x[(size_t)b * CN + i];
        out[i] = s;
    }
}

__global__ void k_repack_wlast(const float* __restrict__ w, int Ci, int Co, int H, int HP, int mode,
                               float* __restrict__ out, int total)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (mode == 0) {  // out[(i*H+h)*CoP + j]
            int CoP = (Co + 3) & ~3;
            int j = t % CoP, r = t / CoP, h = r % H, i = r / H;
            if (j < Co) v = w[((size_t)i * Co + j) * H + h];
        } else if (mode == 1) {  // out[(j*H+h)*CiP + i]
            int CiP = (Ci + 3) & ~3;
            int i = t % CiP, r = t / CiP, h = r % H, j = r / H;
            if (i < Ci) v = w[((size_t)i * Co + j) * H + h];
        } else {  // out[(j*Ci+i)*HP + h]
            int h = t % HP, r = t / HP, i = r % Ci, j = r / Ci;
            if (h < H) v = w[((size_t)i * Co + j) * H + h];
        }
        out[t] = v;
    }
}

// dW2 in mode-1 layout [(j*H+h)*CiP + i]  ->  d_w_last[(i*Co+j)*H + h] +=
__global__ void k_unpack_dwlast(const float* __restrict__ dW2, int Ci, int Co, int H, float* __restrict__ dw)
{
    int CiP = (Ci + 3) & ~3;
    int total = Ci * Co * H;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int h = t % H, r = t / H, j = r % Co, i = r / Co;
        dw[t] += dW2[((size_t)j * H + h) * CiP + i];
    }
}

}  // namespace

extern "C" {

int qc_pack_features(const float* x, int B, int C, int N, float* xp, void* stream)
{
    if (B <= 0 || C <= 0 || N <= 0 || C > 4 * 65535) return QC_EINVAL;
    dim3 grid(qc_ceil_div(N, PT), qc_cp4(C) / 4, qc_nblk(B));
    if (N % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
        k_pack<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, C, N, qc_bp(B), xp);
    else
        k_pack<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, C, N, qc_bp(B), xp);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_unpack_features(const float* xp, int B, int C, int N, const float* bias, float* x, void* stream)
{
    if (B <= 0 || C <= 0 || N <= 0 || C > 4 * 65535) return QC_EINVAL;
    dim3 grid(qc_ceil_div(N, PT), qc_cp4(C) / 4, qc_nblk(B));
    const bool vec = N % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    if (vec)
        k_unpack<true><<<grid, 256, 0, (cudaStream_t)stream>>>(xp, B, C, N, qc_bp(B), bias, x);
    else
        k_unpack<false><<<grid, 256, 0, (cudaStream_t)stream>>>(xp, B, C, N, qc_bp(B), bias, x);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_batch_sum(const float* x, int B, int C, int N, float* out, void* stream)
{
    int CN = C * N;
    k_batch_sum<<<min(qc_ceil_div(CN, 256), qc_num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(x, B, CN, out);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_repack_wlast(const float* w_last, int Ci, int Co, int H, int HP, int mode, float* out, void* stream)
{
    int total;
    if (mode == 0) total = Ci * H * ((Co + 3) & ~3);
    else if (mode == 1) total = Co * H * ((Ci + 3) & ~3);
    else if (mode == 2) total = Co * Ci * HP;
    else return QC_EINVAL;
    k_repack_wlast<<<min(qc_ceil_div(total, 256), qc_num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(w_last, Ci, Co, H, HP, mode, out, total);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_unpack_dwlast(const float* dW2, int Ci, int Co, int H, float* d_w_last, void* stream)
{
    int total = Ci * Co * H;
    k_unpack_dwlast<<<min(qc_ceil_div(total, 256), qc_num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(dW2, Ci, Co, H, d_w_last);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_version(void) { return 1; }

}  // extern "C"
