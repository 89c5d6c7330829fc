// pack.cu -- layout conversion between the reference's [B,C,N] activations
// (torch_quadconv/quadconv.py:238 docstring, point index innermost) and the packed
// [nblk][N][C][Bp] layout (batch innermost) the fused kernels gather from, plus the small weight
// repacks of the last filter-MLP layer.  Pure HBM-bound copies: every global access is a full
// 32-float (128 B) segment on one side and a Bp-float segment on the other, staged through a
// padded shared-memory tile.
#include "common.cuh"

namespace {

// grid: (ceil(N/32), CP4/4, nblk); block 256 = 8 warps.  Tile: 32 points x 4 channels x Bp lanes.
__global__ void __launch_bounds__(256)
k_pack(const float* __restrict__ x, int B, int C, int N, int Bp, float* __restrict__ xp)
{
    __shared__ float tile[4][32 * 33 + 8];  // [channel in group][batch lane * 33 + point]; +8: the 4 planes hit distinct banks
    const int n0 = blockIdx.x * 32, cg = blockIdx.y, blk = blockIdx.z;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int CG = gridDim.y;
    for (int r = w; r < 4 * Bp; r += 8) {
        const int e = r / Bp, bl = r % Bp;
        const int c = cg * 4 + e, b = blk * 32 + bl, n = n0 + lane;
        float v = 0.f;
        if (c < C && b < B && n < N) v = x[((size_t)b * C + c) * N + n];
        tile[e][bl * 33 + lane] = v;
    }
    __syncthreads();
    // write: per point Bp*4 contiguous floats, (lane, channel-in-group) innermost
    for (int idx = threadIdx.x; idx < 32 * Bp * 4; idx += 256) {
        const int e = idx & 3, bl = (idx >> 2) % Bp, nl = (idx >> 2) / Bp;
        const int n = n0 + nl;
        if (n < N) xp[((((size_t)blk * N + n) * CG + cg) * Bp + bl) * 4 + e] = tile[e][bl * 33 + nl];
    }
}

__global__ void __launch_bounds__(256)
k_unpack(const float* __restrict__ xp, int B, int C, int N, int Bp, const float* __restrict__ bias,
         float* __restrict__ x)
{
    __shared__ float tile[4][32 * 33 + 8];
    const int n0 = blockIdx.x * 32, cg = blockIdx.y, blk = blockIdx.z;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int CG = gridDim.y;
    for (int idx = threadIdx.x; idx < 32 * Bp * 4; idx += 256) {
        const int e = idx & 3, bl = (idx >> 2) % Bp, nl = (idx >> 2) / Bp;
        const int n = n0 + nl;
        tile[e][bl * 33 + nl] = n < N ? xp[((((size_t)blk * N + n) * CG + cg) * Bp + bl) * 4 + e] : 0.f;
    }
    __syncthreads();
    const int n = n0 + lane;
    for (int r = w; r < 4 * Bp; r += 8) {
        const int e = r / Bp, bl = r % Bp;
        const int c = cg * 4 + e, b = blk * 32 + bl;
        if (c < C && b < B && n < N) {
            const float bv = bias != nullptr ? bias[(size_t)c * N + n] : 0.f;
            x[((size_t)b * C + c) * N + n] = tile[e][bl * 33 + lane] + bv;
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
    dim3 grid(qc_ceil_div(N, 32), qc_cp4(C) / 4, qc_nblk(B));
    k_pack<<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, C, N, qc_bp(B), xp);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_unpack_features(const float* xp, int B, int C, int N, const float* bias, float* x, void* stream)
{
    if (B <= 0 || C <= 0 || N <= 0 || C > 4 * 65535) return QC_EINVAL;
    dim3 grid(qc_ceil_div(N, 32), qc_cp4(C) / 4, qc_nblk(B));
    k_unpack<<<grid, 256, 0, (cudaStream_t)stream>>>(xp, B, C, N, qc_bp(B), bias, x);
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
