// pool.cu -- K5/K6: Mesh_MaxPool as a segmented max (and its adjoint gather) with analytic
// backward.  Reference: torch_quadconv/mesh_pool.py:88-89 (zeros + scatter_reduce_ amax,
// include_self=False over an int64 [B,C,N] index) and :93 (gather), backward by autograd.
// Here the group structure is one int32 CSR shared by every (batch, channel) row, so the
// reduction is atomic-free and deterministic; ties split the gradient evenly like torch's amax.
// Rows are [BC][N] with the point index innermost; threads walk the point index so stores are
// coalesced and the member reads of a group hit nearby cache lines.
//
// ACT = 1 fuses the user-side activation that precedes the pool in the reference's autoencoders, pool(sin(x))
// (SURVEY 8 f-4): the sine is evaluated on the fly (accurate sinf, the function torch.sin launches), the backward
// recomputes it for the arg-max test and multiplies the routed gradient by cosf(x) -- one launch and one [B,C,N] pass
// per direction instead of three launches and three passes.  ACT = 0 is the reference's pool.
#include "common.cuh"

namespace {

template <int ACT> __device__ __forceinline__ float act_f(float v) { return ACT == 1 ? sinf(v) : v; }
template <int ACT> __device__ __forceinline__ float act_d(float v) { return ACT == 1 ? cosf(v) : 1.f; }

template <int ACT>
__global__ void k_segmax_fwd(const float* __restrict__ x, const int* __restrict__ grp_ptr,
                             const int* __restrict__ grp_mem, int BC, int Nin, int Nout, float* __restrict__ out)
{
    const int64_t total = (int64_t)BC * Nout;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(t % Nout);
        const int64_t r = t / Nout;
        const int beg = grp_ptr[g], end = grp_ptr[g + 1];
        float m = 0.f;  // empty group keeps the reference's zero initialisation
        if (end > beg) {
            // NaN is sticky, like torch's amax (fmaxf would drop it and hide a diverged activation)
            m = -INFINITY;
            for (int e = beg; e < end; ++e) {
                const float v = act_f<ACT>(__ldg(x + r * Nin + grp_mem[e]));
                m = (v != v || v > m) ? v : m;
                if (m != m) break;
            }
        }
        out[t] = m;
    }
}

template <int ACT>
__global__ void k_segmax_bwd(const float* __restrict__ x, const float* __restrict__ out,
                             const float* __restrict__ grad_out, const int* __restrict__ group,
                             const int* __restrict__ grp_ptr, const int* __restrict__ grp_mem, int BC, int Nin,
                             int Nout, float* __restrict__ grad_x)
{
    const int64_t total = (int64_t)BC * Nin;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % Nin);
        const int64_t r = t / Nin;
        const int g = group[k];
        const float m = out[r * Nout + g];
        float gx = 0.f;
        if (m != m) {
            // torch: (src == result) is false for every member, grad / 0 ties, false * inf -> NaN for the whole group
            gx = m;
        } else if (act_f<ACT>(x[t]) == m) {
            // torch's amax backward counts `self == result` as one more tie even with
            // include_self=False, and self is the reference's zeros() tensor (mesh_pool.py:88):
            // a group whose maximum is exactly 0 spreads its gradient over ties+1.  Reproduced
            // for parity with the reference's autograd.
            int ties = (m == 0.f) ? 1 : 0;
            for (int e = grp_ptr[g]; e < grp_ptr[g + 1]; ++e) ties += (act_f<ACT>(__ldg(x + r * Nin + grp_mem[e])) == m);
            gx = grad_out[r * Nout + g] / (float)ties * act_d<ACT>(x[t]);
        }
        grad_x[t] = gx;
    }
}

template <int ACT>
__global__ void k_gather_fwd(const float* __restrict__ x, const int* __restrict__ index, int BC, int Nin,
                             int Nout, float* __restrict__ out)
{
    const int64_t total = (int64_t)BC * Nout;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % Nout);
        const int64_t r = t / Nout;
        out[t] = act_f<ACT>(__ldg(x + r * Nin + index[k]));
    }
}

template <int ACT>
__global__ void k_gather_bwd(const float* __restrict__ grad_out, const float* __restrict__ x, const int* __restrict__ grp_ptr,
                             const int* __restrict__ grp_mem, int BC, int Nfine, int Ncoarse, float* __restrict__ grad_x)
{
    const int64_t total = (int64_t)BC * Ncoarse;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(t % Ncoarse);
        const int64_t r = t / Ncoarse;
        float s = 0.f;
        for (int e = grp_ptr[g]; e < grp_ptr[g + 1]; ++e) s += __ldg(grad_out + r * Nfine + grp_mem[e]);
        grad_x[t] = ACT ? s * act_d<ACT>(x[t]) : s;
    }
}

inline int grid_for(int64_t total) { return min(qc_ceil_div(total > 0 ? total : 1, 256), qc_num_sms() * 16); }

}  // namespace

extern "C" {

// act: 0 = none (the reference's pool), 1 = sin applied to the input first (see the header comment)
int qc_segmax_act_fwd(const float* x, const int* grp_ptr, const int* grp_mem, int BC, int Nin, int Nout, int act,
                      float* out, void* stream)
{
    if (BC <= 0 || Nin <= 0 || Nout <= 0 || act < 0 || act > 1) return QC_EINVAL;
    const int grid = grid_for((int64_t)BC * Nout);
    if (act) k_segmax_fwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, grp_ptr, grp_mem, BC, Nin, Nout, out);
    else k_segmax_fwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, grp_ptr, grp_mem, BC, Nin, Nout, out);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_segmax_act_bwd(const float* x, const float* out, const float* grad_out, const int* group, const int* grp_ptr,
                      const int* grp_mem, int BC, int Nin, int Nout, int act, float* grad_x, void* stream)
{
    if (BC <= 0 || Nin <= 0 || Nout <= 0 || act < 0 || act > 1) return QC_EINVAL;
    const int grid = grid_for((int64_t)BC * Nin);
    if (act) k_segmax_bwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, grad_out, group, grp_ptr, grp_mem, BC, Nin, Nout, grad_x);
    else k_segmax_bwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, grad_out, group, grp_ptr, grp_mem, BC, Nin, Nout, grad_x);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_gather_act_fwd(const float* x, const int* index, int BC, int Nin, int Nout, int act, float* out, void* stream)
{
    if (BC <= 0 || Nin <= 0 || Nout <= 0 || act < 0 || act > 1) return QC_EINVAL;
    const int grid = grid_for((int64_t)BC * Nout);
    if (act) k_gather_fwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, index, BC, Nin, Nout, out);
    else k_gather_fwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, index, BC, Nin, Nout, out);
    QC_CHECK_LAUNCH();
    return 0;
}

// x (the coarse input of the forward call) is read only when act != 0
int qc_gather_act_bwd(const float* grad_out, const float* x, const int* grp_ptr, const int* grp_mem, int BC, int Nfine,
                      int Ncoarse, int act, float* grad_x, void* stream)
{
    if (BC <= 0 || Nfine <= 0 || Ncoarse <= 0 || act < 0 || act > 1 || (act && x == nullptr)) return QC_EINVAL;
    const int grid = grid_for((int64_t)BC * Ncoarse);
    if (act) k_gather_bwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, x, grp_ptr, grp_mem, BC, Nfine, Ncoarse, grad_x);
    else k_gather_bwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, x, grp_ptr, grp_mem, BC, Nfine, Ncoarse, grad_x);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_segmax_fwd(const float* x, const int* grp_ptr, const int* grp_mem, int BC, int Nin, int Nout,
                  float* out, void* stream)
{
    return qc_segmax_act_fwd(x, grp_ptr, grp_mem, BC, Nin, Nout, 0, out, stream);
}

int qc_segmax_bwd(const float* x, const float* out, const float* grad_out, const int* group,
                  const int* grp_ptr, const int* grp_mem, int BC, int Nin, int Nout, float* grad_x,
                  void* stream)
{
    return qc_segmax_act_bwd(x, out, grad_out, group, grp_ptr, grp_mem, BC, Nin, Nout, 0, grad_x, stream);
}

int qc_gather_fwd(const float* x, const int* index, int BC, int Nin, int Nout, float* out, void* stream)
{
    return qc_gather_act_fwd(x, index, BC, Nin, Nout, 0, out, stream);
}

int qc_gather_bwd(const float* grad_out, const int* grp_ptr, const int* grp_mem, int BC, int Nfine,
                  int Ncoarse, float* grad_x, void* stream)
{
    return qc_gather_act_bwd(grad_out, nullptr, grp_ptr, grp_mem, BC, Nfine, Ncoarse, 0, grad_x, stream);
}

}  // extern "C"
