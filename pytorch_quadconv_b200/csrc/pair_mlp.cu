// pair_mlp.cu -- K2a: hidden layers of the filter MLP, one thread per (out,in) pair.
//
// Reference: per pair the reference gathers the quadrature weight (quadconv.py:247), forms the
// offset eval_locs = P_out[idx0] - P_in[idx1] (quadconv.py:250-253) and pushes it through
// self.filter = Sequential(Linear(bias=False), Sin, ..., Linear(bias=False)) (quadconv.py:130-145,
// utils/misc.py:18-19).  The LAST Linear (h_L -> Ci*Co) is linear in its input, so it is folded
// into the contraction kernel (contract.cu); this file evaluates everything before it:
//     phi[p][h] = w[in(p)] * x_L[h],  x_0 = z_p,  x_{l+1} = sin(W_l x_l)
// and its backward.  The hidden widths are tiny (4..32), so this is CUDA-core work: the weights sit
// zero-padded in shared memory and are read as warp-uniform broadcasts; activations stay in
// registers.  sinf/cosf are the accurate versions (no fast-math) to hold the 1e-5 fp32 bound.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

struct MlpDev {
    int bf16;   // operands of every Linear rounded to bf16 (qc_mlp_t::operand_bf16)
    int n_hidden;
    int dims[QC_MAX_MLP_LAYERS + 1];
    const float* w[QC_MAX_MLP_LAYERS];
    float* dw[QC_MAX_MLP_LAYERS];
};

__device__ __forceinline__ float rnd_operand(float v, int bf16)
{
    return bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}

template <int WMAX>
__device__ __forceinline__ void load_weights(const MlpDev& m, float* Wsm)
{
    // Wsm[l][o*WMAX + k], zero padded
    for (int l = 0; l < m.n_hidden; ++l) {
        const int nin = m.dims[l], nout = m.dims[l + 1];
        for (int t = threadIdx.x; t < WMAX * WMAX; t += blockDim.x) {
            int o = t / WMAX, k = t % WMAX;
            Wsm[l * WMAX * WMAX + t] = (o < nout && k < nin) ? rnd_operand(m.w[l][o * nin + k], m.bf16) : 0.f;
        }
    }
}

template <int WMAX>
__global__ void __launch_bounds__(128)
k_pair_phi_fwd(const float* __restrict__ out_pts, const float* __restrict__ in_pts,
               const float* __restrict__ quad_w, const int* __restrict__ pair_row,
               const int* __restrict__ pair_col, int64_t P, int d, MlpDev m, int HP,
               float* __restrict__ phi)
{
    extern __shared__ float smem[];
    float* Wsm = smem;
    load_weights<WMAX>(m, Wsm);
    __syncthreads();
    const int H = m.dims[m.n_hidden];
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const int r = pair_row[p], c = pair_col[p];
        float a[WMAX];
#pragma unroll
        for (int k = 0; k < WMAX; ++k) a[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < d && k < WMAX) a[k] = __fsub_rn(out_pts[(size_t)r * d + k], in_pts[(size_t)c * d + k]);
#pragma unroll
        for (int k = 0; k < WMAX; ++k) a[k] = rnd_operand(a[k], m.bf16);
        for (int l = 0; l < m.n_hidden; ++l) {
            const float* W = Wsm + l * WMAX * WMAX;
            float t[WMAX];
#pragma unroll
            for (int o = 0; o < WMAX; ++o) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < WMAX; ++k) acc = fmaf(W[o * WMAX + k], a[k], acc);
                t[o] = rnd_operand(sinf(acc), m.bf16);
            }
#pragma unroll
            for (int o = 0; o < WMAX; ++o) a[o] = t[o];
        }
        const float wq = quad_w ? quad_w[c] : 1.f;
        float* dst = phi + (size_t)p * HP;
#pragma unroll
        for (int h = 0; h < WMAX; ++h)
            if (h < HP) dst[h] = h < H ? wq * a[h] : 0.f;
        // HP may exceed WMAX only when H <= WMAX < HP, which the host never configures
    }
}

// Backward.  Block = 128 pairs per sweep.  Per layer the block stages (da, x) of its pairs in
// shared memory and every thread owns a fixed set of dW[o][k] entries, so the weight gradient is
// reduced without atomics inside the block; one atomicAdd per entry per block at the end.
// LT > 0: the number of hidden layers is a compile-time constant, so the per-layer activations and
// cosines live in registers (with a runtime layer count they are dynamically indexed and ptxas puts
// them in local memory); LT == 0 is the general version.
template <int WMAX, int LT>
__global__ void __launch_bounds__(128)
k_pair_phi_bwd(const float* __restrict__ out_pts, const float* __restrict__ in_pts,
               const float* __restrict__ quad_w, const int* __restrict__ pair_row,
               const int* __restrict__ pair_col, int64_t P, int d, MlpDev m, int HP,
               const float* __restrict__ dphi, int nslices, float* __restrict__ d_quad_w)
{
    constexpr int NT = 128;
    extern __shared__ float smem[];
    float* Wsm = smem;                                  // [n_hidden][WMAX*WMAX]
    float* dWs = Wsm + m.n_hidden * WMAX * WMAX;        // [n_hidden][WMAX*WMAX]
    float* da_s = dWs + m.n_hidden * WMAX * WMAX;       // [NT][WMAX]
    float* x_s = da_s + NT * WMAX;                      // [NT][WMAX]
    load_weights<WMAX>(m, Wsm);
    for (int t = threadIdx.x; t < m.n_hidden * WMAX * WMAX; t += NT) dWs[t] = 0.f;
    __syncthreads();
    const int H = m.dims[m.n_hidden];
    constexpr int LMAX = LT > 0 ? LT : QC_MAX_MLP_LAYERS;
    const int L = LT > 0 ? LT : m.n_hidden;

    for (int64_t base = (int64_t)blockIdx.x * NT; base < P; base += (int64_t)gridDim.x * NT) {
        const int64_t p = base + threadIdx.x;
        const bool valid = p < P;
        // forward recompute, keeping every layer input and the cosine of every pre-activation
        float xs[LMAX + 1][WMAX];
        float cs[LMAX][WMAX];
        int r = 0, c = 0;
        if (valid) {
            r = pair_row[p];
            c = pair_col[p];
        }
#pragma unroll
        for (int k = 0; k < WMAX; ++k) xs[0][k] = 0.f;
        if (valid) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < d && k < WMAX)
                    xs[0][k] = rnd_operand(__fsub_rn(out_pts[(size_t)r * d + k], in_pts[(size_t)c * d + k]), m.bf16);
        }
#pragma unroll(LT > 0 ? LMAX : 1)
        for (int l = 0; l < LMAX; ++l) {
            if (l >= L) break;
            const float* W = Wsm + l * WMAX * WMAX;
#pragma unroll
            for (int o = 0; o < WMAX; ++o) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < WMAX; ++k) acc = fmaf(W[o * WMAX + k], xs[l][k], acc);
                float sv, cv;
                sincosf(acc, &sv, &cv);
                xs[l + 1][o] = rnd_operand(sv, m.bf16);   // straight-through: the derivative ignores the rounding
                cs[l][o] = cv;
            }
        }
        const float wq = (valid && quad_w) ? quad_w[c] : 1.f;
        float dx[WMAX];
        float dwq = 0.f;
        float gs[WMAX];   // d loss / d phi of this pair: the slices are summed in a fixed order, 128 bits at a time
#pragma unroll
        for (int h = 0; h < WMAX; ++h) gs[h] = 0.f;
        if (valid) {
            for (int sl = 0; sl < nslices; ++sl) {
                const float4* src = reinterpret_cast<const float4*>(dphi + ((size_t)sl * P + p) * HP);
#pragma unroll
                for (int q = 0; q < WMAX / 4; ++q)
                    if (4 * q < HP) {
                        const float4 v = __ldg(src + q);
                        gs[4 * q] += v.x, gs[4 * q + 1] += v.y, gs[4 * q + 2] += v.z, gs[4 * q + 3] += v.w;
                    }
            }
        }
#pragma unroll
        for (int h = 0; h < WMAX; ++h) {
            const float g = h < H ? gs[h] : 0.f;
            dwq = fmaf(g, LT > 0 ? xs[LT][h] : xs[L][h], dwq);
            dx[h] = wq * g;
        }
        if (valid && d_quad_w) atomicAdd(&d_quad_w[c], dwq);

#pragma unroll(LT > 0 ? LMAX : 1)
        for (int l = LMAX - 1; l >= 0; --l) {
            if (l >= L) continue;
            const float* W = Wsm + l * WMAX * WMAX;
            float da[WMAX];
#pragma unroll
            for (int o = 0; o < WMAX; ++o) {
                da[o] = dx[o] * cs[l][o];
                da_s[threadIdx.x * WMAX + o] = da[o];
                x_s[threadIdx.x * WMAX + o] = xs[l][o];
            }
            __syncthreads();
            // dW_l[o][k] += sum_q da[q][o] * x[q][k]: a register-tiled mini-GEMM over the block's 128 pairs.
            // Every thread owns a TL x TL tile of (o,k) and a range of q; vector loads of TL values feed
            // TL*TL FMAs; the q ranges are combined with shared-memory atomics.
            {
                constexpr int TL = WMAX >= 16 ? 4 : 2;
                constexpr int TPR = WMAX / TL, NTILE = TPR * TPR;
                constexpr int NQR = NT / NTILE, QL = NT / NQR;
                const int tile = threadIdx.x % NTILE, qr = threadIdx.x / NTILE;
                const int o0 = (tile / TPR) * TL, k0 = (tile % TPR) * TL;
                float acc[TL][TL];
#pragma unroll
                for (int i = 0; i < TL; ++i)
#pragma unroll
                    for (int j = 0; j < TL; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                for (int q = qr * QL; q < (qr + 1) * QL; ++q) {
                    float dv[TL], xv[TL];
                    if (TL == 4) {
                        const float4 a4 = *reinterpret_cast<const float4*>(da_s + q * WMAX + o0);
                        const float4 b4 = *reinterpret_cast<const float4*>(x_s + q * WMAX + k0);
                        dv[0] = a4.x, dv[1] = a4.y, dv[TL - 2] = a4.z, dv[TL - 1] = a4.w;
                        xv[0] = b4.x, xv[1] = b4.y, xv[TL - 2] = b4.z, xv[TL - 1] = b4.w;
                    } else {
                        const float2 a2 = *reinterpret_cast<const float2*>(da_s + q * WMAX + o0);
                        const float2 b2 = *reinterpret_cast<const float2*>(x_s + q * WMAX + k0);
                        dv[0] = a2.x, dv[1] = a2.y;
                        xv[0] = b2.x, xv[1] = b2.y;
                    }
#pragma unroll
                    for (int i = 0; i < TL; ++i)
#pragma unroll
                        for (int j = 0; j < TL; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < TL; ++i)
#pragma unroll
                    for (int j = 0; j < TL; ++j)
                        atomicAdd(&dWs[l * WMAX * WMAX + (o0 + i) * WMAX + k0 + j], acc[i][j]);
            }
            if (l > 0) {
#pragma unroll
                for (int k = 0; k < WMAX; ++k) {
                    float acc = 0.f;
#pragma unroll
                    for (int o = 0; o < WMAX; ++o) acc = fmaf(W[o * WMAX + k], da[o], acc);
                    dx[k] = acc;
                }
            }
            __syncthreads();
        }
    }
    // flush
    for (int l = 0; l < L; ++l) {
        const int nin = m.dims[l], nout = m.dims[l + 1];
        for (int t = threadIdx.x; t < WMAX * WMAX; t += NT) {
            int o = t / WMAX, k = t % WMAX;
            if (o < nout && k < nin) atomicAdd(&m.dw[l][o * nin + k], dWs[l * WMAX * WMAX + t]);
        }
    }
}

int pick_wmax(const qc_mlp_t* mlp, int d)
{
    int mx = d;
    for (int l = 1; l <= mlp->n_hidden; ++l) mx = mlp->dims[l] > mx ? mlp->dims[l] : mx;
    if (mx <= 4) return 4;
    if (mx <= 8) return 8;
    if (mx <= 16) return 16;
    if (mx <= 32) return 32;
    return -1;
}

int fill_dev(const qc_mlp_t* mlp, int d, MlpDev& m)
{
    if (mlp->n_hidden < 0 || mlp->n_hidden > QC_MAX_MLP_LAYERS) return QC_EUNSUP;
    if (mlp->dims[0] != d) return QC_EINVAL;
    m.n_hidden = mlp->n_hidden;
    m.bf16 = mlp->operand_bf16 != 0;
    for (int l = 0; l <= QC_MAX_MLP_LAYERS; ++l) m.dims[l] = l <= mlp->n_hidden ? mlp->dims[l] : 0;
    for (int l = 0; l < QC_MAX_MLP_LAYERS; ++l) {
        m.w[l] = l < mlp->n_hidden ? mlp->w[l] : nullptr;
        m.dw[l] = nullptr;
    }
    return 0;
}

template <int WMAX>
int launch_fwd(const float* out_pts, const float* in_pts, const float* quad_w, const int* pair_row,
               const int* pair_col, int64_t P, int d, const MlpDev& m, int HP, float* phi, cudaStream_t st)
{
    size_t smem = (size_t)(m.n_hidden > 0 ? m.n_hidden : 1) * WMAX * WMAX * 4;
    QC_CUDA((qc_set_max_smem<k_pair_phi_fwd<WMAX>>((int)smem)));
    int blocks = min(qc_ceil_div(P, 128), qc_num_sms() * 16);
    k_pair_phi_fwd<WMAX><<<blocks, 128, smem, st>>>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi);
    QC_CHECK_LAUNCH();
    return 0;
}

template <int WMAX, int LT>
int launch_bwd(const float* out_pts, const float* in_pts, const float* quad_w, const int* pair_row,
               const int* pair_col, int64_t P, int d, const MlpDev& m, int HP, const float* dphi,
               int nslices, float* d_quad_w, cudaStream_t st)
{
    size_t smem = ((size_t)2 * m.n_hidden * WMAX * WMAX + 2 * 128 * WMAX) * 4;
    QC_CUDA((qc_set_max_smem<k_pair_phi_bwd<WMAX, LT>>((int)smem)));
    int blocks = min(qc_ceil_div(P, 128), qc_num_sms() * 16);
    k_pair_phi_bwd<WMAX, LT><<<blocks, 128, smem, st>>>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, dphi, nslices, d_quad_w);
    QC_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" {

int qc_pair_phi_fwd(const float* out_pts, const float* in_pts, const float* quad_w,
                    const int* pair_row, const int* pair_col, int64_t P, int d,
                    const qc_mlp_t* mlp, int HP, float* phi, void* stream)
{
    if (P <= 0) return 0;
    if (d < 1 || d > 3) return QC_EUNSUP;
    MlpDev m;
    int rc = fill_dev(mlp, d, m);
    if (rc) return rc;
    const int H = m.dims[m.n_hidden];
    const int wmax = pick_wmax(mlp, d);
    if (wmax < 0 || HP < H || HP > 32 || (HP & 3)) return QC_EUNSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const int wsel = wmax > HP ? wmax : HP;  // the phi row (HP wide) is written from registers
    switch (wsel) {
        case 4: return launch_fwd<4>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
        case 8: return launch_fwd<8>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
        case 16: return launch_fwd<16>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
        case 32: return launch_fwd<32>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
    }
    return QC_EUNSUP;
}

int qc_pair_phi_bwd(const float* out_pts, const float* in_pts, const float* quad_w,
                    const int* pair_row, const int* pair_col, int64_t P, int d,
                    const qc_mlp_t* mlp, int HP, const float* dphi, int nslices,
                    float* const* d_mlp_w, float* d_quad_w, void* stream)
{
    if (P <= 0) return 0;
    if (d < 1 || d > 3 || nslices < 1) return QC_EUNSUP;
    MlpDev m;
    int rc = fill_dev(mlp, d, m);
    if (rc) return rc;
    for (int l = 0; l < m.n_hidden; ++l) m.dw[l] = d_mlp_w[l];
    const int H = m.dims[m.n_hidden];
    const int wmax = pick_wmax(mlp, d);
    if (wmax < 0 || HP < H || HP > 32 || (HP & 3)) return QC_EUNSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const int wsel = wmax > HP ? wmax : HP;
#define QC_BWD(W, LTV) return launch_bwd<W, LTV>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, dphi, nslices, d_quad_w, st)
    const int L = m.n_hidden;
    switch (wsel) {   // register-resident activations when (2L+1)*W floats fit comfortably, else the general kernel
        case 4: if (L == 1) QC_BWD(4, 1); if (L == 2) QC_BWD(4, 2); if (L == 3) QC_BWD(4, 3); if (L == 4) QC_BWD(4, 4); QC_BWD(4, 0);
        case 8: if (L == 1) QC_BWD(8, 1); if (L == 2) QC_BWD(8, 2); if (L == 3) QC_BWD(8, 3); if (L == 4) QC_BWD(8, 4); QC_BWD(8, 0);
        case 16: if (L == 1) QC_BWD(16, 1); if (L == 2) QC_BWD(16, 2); QC_BWD(16, 0);
        case 32: QC_BWD(32, 0);   // 160+ registers per thread: the occupancy loss costs more than the local-memory traffic
    }
#undef QC_BWD
    return QC_EUNSUP;
}

}  // extern "C"
