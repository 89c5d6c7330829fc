// weight_map.cu -- fused learned quadrature-weight map.
//
// Reference: MeshHandler.build_weight_map / weight_map, torch_quadconv/mesh_handler.py:85-116: for every
// simplex the coordinates of its E vertices go through Linear(E*d,8)-act-Linear(8,8)-act-Linear(8,8)-act-
// Linear(8,E)-act (act = Sigmoid by default) and the E outputs are scatter-added to the vertices.  The
// reference (and every QuadConv.forward, quadconv.py:247) runs this as ~15 ATen kernels forward and as
// many backward; here it is one kernel per direction plus a deterministic per-vertex gather-sum
// (incidence CSR built once per mesh level) instead of the atomic scatter_add_.
#include "common.cuh"

namespace {

constexpr int WM_HID = 8;    // hidden width (fixed by the reference)
constexpr int WM_MAXIN = 12; // E*d <= 12 (tetrahedra in 3-D)
constexpr int WM_MAXE = 4;
constexpr int WM_NW = WM_HID * WM_MAXIN + 2 * WM_HID * WM_HID + WM_MAXE * WM_HID;   // per-block weight-gradient sums
constexpr int WM_NB = 3 * WM_HID + WM_MAXE;                                        // ... and bias-gradient sums
constexpr int WM_PART = WM_NW + WM_NB;

struct WmParams {
    const float* W[4];
    const float* b[4];
    float* dW[4];
    float* db[4];
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct WmSmem {
    float W0[WM_HID * WM_MAXIN], b0[WM_HID];
    float W1[WM_HID * WM_HID], b1[WM_HID];
    float W2[WM_HID * WM_HID], b2[WM_HID];
    float W3[WM_MAXE * WM_HID], b3[WM_MAXE];
};

__device__ __forceinline__ void wm_load(const WmParams& p, WmSmem& s, int F0, int E)
{
    for (int t = threadIdx.x; t < WM_HID * WM_MAXIN; t += blockDim.x) {
        int o = t / WM_MAXIN, k = t % WM_MAXIN;
        s.W0[t] = k < F0 ? p.W[0][o * F0 + k] : 0.f;
    }
    for (int t = threadIdx.x; t < WM_HID * WM_HID; t += blockDim.x) {
        s.W1[t] = p.W[1][t];
        s.W2[t] = p.W[2][t];
    }
    for (int t = threadIdx.x; t < WM_MAXE * WM_HID; t += blockDim.x) s.W3[t] = (t / WM_HID) < E ? p.W[3][t] : 0.f;
    for (int t = threadIdx.x; t < WM_HID; t += blockDim.x) {
        s.b0[t] = p.b[0][t];
        s.b1[t] = p.b[1][t];
        s.b2[t] = p.b[2][t];
    }
    for (int t = threadIdx.x; t < WM_MAXE; t += blockDim.x) s.b3[t] = t < E ? p.b[3][t] : 0.f;
}

// forward of one simplex; keeps every activation for the backward
__device__ __forceinline__ void wm_forward(const WmSmem& s, const float (&x)[WM_MAXIN], float (&a1)[WM_HID],
                                           float (&a2)[WM_HID], float (&a3)[WM_HID], float (&o)[WM_MAXE])
{
#pragma unroll
    for (int r = 0; r < WM_HID; ++r) {
        float acc = s.b0[r];
#pragma unroll
        for (int k = 0; k < WM_MAXIN; ++k) acc = fmaf(s.W0[r * WM_MAXIN + k], x[k], acc);
        a1[r] = sigmoidf_(acc);
    }
#pragma unroll
    for (int r = 0; r < WM_HID; ++r) {
        float acc = s.b1[r];
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) acc = fmaf(s.W1[r * WM_HID + k], a1[k], acc);
        a2[r] = sigmoidf_(acc);
    }
#pragma unroll
    for (int r = 0; r < WM_HID; ++r) {
        float acc = s.b2[r];
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) acc = fmaf(s.W2[r * WM_HID + k], a2[k], acc);
        a3[r] = sigmoidf_(acc);
    }
#pragma unroll
    for (int r = 0; r < WM_MAXE; ++r) {
        float acc = s.b3[r];
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) acc = fmaf(s.W3[r * WM_HID + k], a3[k], acc);
        o[r] = sigmoidf_(acc);
    }
}

__device__ __forceinline__ void wm_gather(const float* __restrict__ pts, const int* __restrict__ adj, int m, int E,
                                          int dd, float (&x)[WM_MAXIN], int (&v)[WM_MAXE])
{
#pragma unroll
    for (int k = 0; k < WM_MAXIN; ++k) x[k] = 0.f;
#pragma unroll
    for (int e = 0; e < WM_MAXE; ++e) {
        v[e] = 0;
        if (e < E) {
            v[e] = adj[(size_t)m * E + e];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < dd) x[e * dd + c] = pts[(size_t)v[e] * dd + c];   // el_points row layout (mesh_handler.py:104)
        }
    }
}

__global__ void __launch_bounds__(128)
k_wm_fwd(const float* __restrict__ pts, const int* __restrict__ adj, int M, int E, int dd, WmParams p, float* __restrict__ el_w)
{
    __shared__ WmSmem s;
    wm_load(p, s, E * dd, E);
    __syncthreads();
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        float x[WM_MAXIN], a1[WM_HID], a2[WM_HID], a3[WM_HID], o[WM_MAXE];
        int v[WM_MAXE];
        wm_gather(pts, adj, m, E, dd, x, v);
        wm_forward(s, x, a1, a2, a3, o);
#pragma unroll
        for (int e = 0; e < WM_MAXE; ++e)
            if (e < E) el_w[(size_t)m * E + e] = o[e];
    }
}

// weights[v] = sum of el_w over the (simplex, slot) incidences of v, ascending flat index (the order
// of the reference's CPU scatter_add_, mesh_handler.py:110) -- deterministic, no atomics.
__global__ void k_vertex_sum(const float* __restrict__ el_w, const int* __restrict__ inc_ptr,
                             const int* __restrict__ inc_idx, int N, float* __restrict__ w)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int t = inc_ptr[v]; t < inc_ptr[v + 1]; ++t) acc += el_w[inc_idx[t]];
        w[v] = acc;
    }
}

// Backward.  Block = 128 simplices per sweep.  Per layer the block stages (dz, input activation) in
// shared memory and every thread owns fixed entries of dW/db, so the reduction inside the block is
// atomic-free; one atomicAdd per parameter entry per block at the end.
__global__ void __launch_bounds__(128)
k_wm_bwd(const float* __restrict__ pts, const int* __restrict__ adj, int M, int E, int dd, WmParams p,
         const float* __restrict__ gw, float* __restrict__ part)
{
    constexpr int NT = 128;
    __shared__ WmSmem s;
    __shared__ float dz_s[NT * WM_HID];
    __shared__ float in_s[NT * WM_MAXIN];
    __shared__ float accW[WM_HID * WM_MAXIN + 2 * WM_HID * WM_HID + WM_MAXE * WM_HID];
    __shared__ float accb[3 * WM_HID + WM_MAXE];
    const int F0 = E * dd;
    wm_load(p, s, F0, E);
    for (int t = threadIdx.x; t < WM_HID * WM_MAXIN + 2 * WM_HID * WM_HID + WM_MAXE * WM_HID; t += NT) accW[t] = 0.f;
    for (int t = threadIdx.x; t < 3 * WM_HID + WM_MAXE; t += NT) accb[t] = 0.f;
    __syncthreads();
    float* aW0 = accW;
    float* aW1 = aW0 + WM_HID * WM_MAXIN;
    float* aW2 = aW1 + WM_HID * WM_HID;
    float* aW3 = aW2 + WM_HID * WM_HID;

    // accumulate dW[o][k] += sum_q dz[q][o] * in[q][k], db[o] += sum_q dz[q][o] for one layer
    auto reduce_layer = [&](float* aW, float* ab, int nout, int nin, int ldw) {
        __syncthreads();
        for (int t = threadIdx.x; t < nout * nin; t += NT) {
            const int o = t / nin, k = t % nin;
            float acc = 0.f;
#pragma unroll 8
            for (int q = 0; q < NT; ++q) acc = fmaf(dz_s[q * WM_HID + o], in_s[q * WM_MAXIN + k], acc);
            aW[o * ldw + k] += acc;
        }
        for (int o = threadIdx.x; o < nout; o += NT) {
            float acc = 0.f;
            for (int q = 0; q < NT; ++q) acc += dz_s[q * WM_HID + o];
            ab[o] += acc;
        }
        __syncthreads();
    };

    for (int base = blockIdx.x * NT; base < M; base += gridDim.x * NT) {
        const int m = base + threadIdx.x;
        const bool valid = m < M;
        float x[WM_MAXIN], a1[WM_HID], a2[WM_HID], a3[WM_HID], o[WM_MAXE];
        int v[WM_MAXE];
        wm_gather(pts, adj, valid ? m : 0, E, dd, x, v);
        wm_forward(s, x, a1, a2, a3, o);
        float dz[WM_HID];
        // output layer: dz3[e] = gw[v_e] * o (1 - o)
#pragma unroll
        for (int e = 0; e < WM_HID; ++e) {
            dz[e] = 0.f;
            if (e < WM_MAXE && e < E && valid) dz[e] = gw[v[e]] * o[e] * (1.f - o[e]);
        }
        float da[WM_HID];
        // ---- layer 3 (8 -> E)
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) {
            dz_s[threadIdx.x * WM_HID + k] = dz[k];
            in_s[threadIdx.x * WM_MAXIN + k] = valid ? a3[k] : 0.f;
        }
        reduce_layer(aW3, accb + 3 * WM_HID, E, WM_HID, WM_HID);
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int e = 0; e < WM_MAXE; ++e) acc = fmaf(s.W3[e * WM_HID + k], dz[e], acc);
            da[k] = acc;
        }
        // ---- layer 2 (8 -> 8)
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) {
            dz[k] = da[k] * a3[k] * (1.f - a3[k]);
            dz_s[threadIdx.x * WM_HID + k] = valid ? dz[k] : 0.f;
            in_s[threadIdx.x * WM_MAXIN + k] = valid ? a2[k] : 0.f;
        }
        reduce_layer(aW2, accb + 2 * WM_HID, WM_HID, WM_HID, WM_HID);
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < WM_HID; ++r) acc = fmaf(s.W2[r * WM_HID + k], dz[r], acc);
            da[k] = acc;
        }
        // ---- layer 1 (8 -> 8)
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) {
            dz[k] = da[k] * a2[k] * (1.f - a2[k]);
            dz_s[threadIdx.x * WM_HID + k] = valid ? dz[k] : 0.f;
            in_s[threadIdx.x * WM_MAXIN + k] = valid ? a1[k] : 0.f;
        }
        reduce_layer(aW1, accb + WM_HID, WM_HID, WM_HID, WM_HID);
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < WM_HID; ++r) acc = fmaf(s.W1[r * WM_HID + k], dz[r], acc);
            da[k] = acc;
        }
        // ---- layer 0 (E*d -> 8); points carry no gradient (mesh_handler.py:35)
#pragma unroll
        for (int k = 0; k < WM_HID; ++k) {
            dz[k] = da[k] * a1[k] * (1.f - a1[k]);
            dz_s[threadIdx.x * WM_HID + k] = valid ? dz[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < WM_MAXIN; ++k) in_s[threadIdx.x * WM_MAXIN + k] = valid ? x[k] : 0.f;
        reduce_layer(aW0, accb, WM_HID, F0, WM_MAXIN);
    }
    __syncthreads();
    if (part != nullptr) {
        // deterministic mode: this block's sums go to its own row of the workspace; k_wm_finish adds the rows in block order
        float* row = part + (size_t)blockIdx.x * WM_PART;
        for (int t = threadIdx.x; t < WM_NW; t += NT) row[t] = accW[t];
        for (int t = threadIdx.x; t < WM_NB; t += NT) row[WM_NW + t] = accb[t];
        return;
    }
    for (int t = threadIdx.x; t < WM_HID * F0; t += NT) atomicAdd(&p.dW[0][t], aW0[(t / F0) * WM_MAXIN + t % F0]);
    for (int t = threadIdx.x; t < WM_HID * WM_HID; t += NT) {
        atomicAdd(&p.dW[1][t], aW1[t]);
        atomicAdd(&p.dW[2][t], aW2[t]);
    }
    for (int t = threadIdx.x; t < E * WM_HID; t += NT) atomicAdd(&p.dW[3][t], aW3[t]);
    for (int t = threadIdx.x; t < WM_HID; t += NT) {
        atomicAdd(&p.db[0][t], accb[t]);
        atomicAdd(&p.db[1][t], accb[WM_HID + t]);
        atomicAdd(&p.db[2][t], accb[2 * WM_HID + t]);
    }
    for (int t = threadIdx.x; t < E; t += NT) atomicAdd(&p.db[3][t], accb[3 * WM_HID + t]);
}

// dW / db += the blocks' partial sums in a FIXED order (bit-reproducible): one warp per entry, lane j adds the rows
// j, j + 32, ... in sequence, then a fixed xor-shuffle tree combines the 32 lanes
__global__ void __launch_bounds__(128) k_wm_finish(const float* __restrict__ part, int nblocks, int E, int dd, WmParams p)
{
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= WM_PART) return;
    float v = 0.f;
    for (int b = lane; b < nblocks; b += 32) v += part[(size_t)b * WM_PART + t];
#pragma unroll
    for (int mk = 16; mk >= 1; mk >>= 1) v += __shfl_xor_sync(0xffffffffu, v, mk);
    if (lane != 0) return;
    const int F0 = E * dd;
    constexpr int O1 = WM_HID * WM_MAXIN, O2 = O1 + WM_HID * WM_HID, O3 = O2 + WM_HID * WM_HID;
    if (t < O1) {
        const int o = t / WM_MAXIN, k = t % WM_MAXIN;
        if (k < F0) p.dW[0][o * F0 + k] += v;
    } else if (t < O2) {
        p.dW[1][t - O1] += v;
    } else if (t < O3) {
        p.dW[2][t - O2] += v;
    } else if (t < WM_NW) {
        if (t - O3 < E * WM_HID) p.dW[3][t - O3] += v;
    } else {
        const int u = t - WM_NW;
        if (u < WM_HID) p.db[0][u] += v;
        else if (u < 2 * WM_HID) p.db[1][u - WM_HID] += v;
        else if (u < 3 * WM_HID) p.db[2][u - 2 * WM_HID] += v;
        else if (u - 3 * WM_HID < E) p.db[3][u - 3 * WM_HID] += v;
    }
}

int fill(WmParams& p, const float* const* wb, float* const* dwb)
{
    for (int l = 0; l < 4; ++l) {
        p.W[l] = wb[2 * l];
        p.b[l] = wb[2 * l + 1];
        p.dW[l] = dwb ? dwb[2 * l] : nullptr;
        p.db[l] = dwb ? dwb[2 * l + 1] : nullptr;
    }
    return 0;
}

}  // namespace

extern "C" {

int qc_weight_map_fwd(const float* pts, const int* adj, int M, int E, int dd, const float* const* wb,
                      float* el_w, void* stream)
{
    if (M <= 0 || E < 2 || E > WM_MAXE || dd < 1 || dd > 3 || E * dd > WM_MAXIN) return QC_EUNSUP;
    WmParams p;
    fill(p, wb, nullptr);
    k_wm_fwd<<<min(qc_ceil_div(M, 128), qc_num_sms() * 16), 128, 0, (cudaStream_t)stream>>>(pts, adj, M, E, dd, p, el_w);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_vertex_sum(const float* el_w, const int* inc_ptr, const int* inc_idx, int N, float* w, void* stream)
{
    if (N <= 0) return QC_EINVAL;
    k_vertex_sum<<<min(qc_ceil_div(N, 256), qc_num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(el_w, inc_ptr, inc_idx, N, w);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_weight_map_bwd(const float* pts, const int* adj, int M, int E, int dd, const float* const* wb,
                      const float* gw, float* const* dwb, void* stream)
{
    if (M <= 0 || E < 2 || E > WM_MAXE || dd < 1 || dd > 3 || E * dd > WM_MAXIN) return QC_EUNSUP;
    WmParams p;
    fill(p, wb, dwb);
    k_wm_bwd<<<min(qc_ceil_div(M, 128), qc_num_sms() * 8), 128, 0, (cudaStream_t)stream>>>(pts, adj, M, E, dd, p, gw, nullptr);
    QC_CHECK_LAUNCH();
    return 0;
}

size_t qc_weight_map_bwd_workspace_bytes(int M)
{
    if (M <= 0) return 0;
    return (size_t)min(qc_ceil_div(M, 128), qc_num_sms() * 8) * WM_PART * sizeof(float);
}

int qc_weight_map_bwd_ws(const float* pts, const int* adj, int M, int E, int dd, const float* const* wb,
                         const float* gw, float* const* dwb, void* ws, size_t ws_bytes, void* stream)
{
    if (M <= 0 || E < 2 || E > WM_MAXE || dd < 1 || dd > 3 || E * dd > WM_MAXIN) return QC_EUNSUP;
    const int blocks = min(qc_ceil_div(M, 128), qc_num_sms() * 8);
    if (ws == nullptr || ws_bytes < (size_t)blocks * WM_PART * sizeof(float)) return QC_EWORKSPACE;
    WmParams p;
    fill(p, wb, dwb);
    k_wm_bwd<<<blocks, 128, 0, (cudaStream_t)stream>>>(pts, adj, M, E, dd, p, gw, reinterpret_cast<float*>(ws));
    QC_CHECK_LAUNCH();
    k_wm_finish<<<qc_ceil_div(WM_PART, 4), 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(ws), blocks, E, dd, p);
    QC_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
