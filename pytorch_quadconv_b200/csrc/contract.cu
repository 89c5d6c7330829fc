// contract.cu -- K2b/K3/K4: fused last-MLP-layer + gather-contract-segmented-reduce (fp32 CUDA-core
// path), and the d(phi) kernel of the backward pass.
//
// Reference (torch_quadconv/quadconv.py): filters = self.G(eval_locs) materialises [P,Ci,Co]
// (:257,:91); _integrate gathers features[:,:,idx1] into [B,Ci,P], runs einsum('n,nij,bin->bjn')
// as P tiny bmm's and scatter_add_'s [B,Co,P] into the output (:213-227).  None of these
// P-sized tensors exist here.  The last Linear has no bias/activation (:143), so with
// phi[p][h] = w * (last hidden activation) the layer is
//     T[o; b,c,h] = sum_{p in N(o)} phi[p][h] * X[in(p)][c][b]          (stage B, segmented sum)
//     Y[o][j][b]  = sum_{c,h} T[o; b,c,h] * W2[(c,h)][j]                (stage C, dense)
// (hard part 3b).
//
// Thread mapping (tcgen05-ready: one thread = one (point, batch) row of the dense stage):
//   lane   -> (point slot s, batch lane b) with Bp = 2^LB batch lanes per point, 32/Bp points per warp
//   warp   -> (wp, wk): wk selects a chunk of CXC = 64/HP gathered channels (K split of stage C)
//   thread accumulates T[CXC][HP] (64 registers) for its row over the point's pair segment; the
//   gathers X[src][c][*] are contiguous Bp-float segments (128 B for Bp = 32), phi and the pair
//   indices are sub-group-uniform broadcast loads.  Segments are contiguous in the CSR, so the
//   per-point reduction needs no atomics and is bit-reproducible.
#include "common.cuh"
#include "contract_args.cuh"

namespace {



constexpr int OUT_STRIDE = 33;  // conflict-free row stride of the output staging tile
constexpr int Z_STRIDE = 36;    // float4-aligned row stride of the Z tile (aliases the output tile)

template <int HP, bool DW>
__global__ void __launch_bounds__(256, DW ? 1 : 2) k_contract(const ContractArgs a)
{
    constexpr int CXC = 64 / HP;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int WK = 1 << a.LWK, WP = 8 >> a.LWK;
    const int wk = warp & (WK - 1), wp = warp >> a.LWK;
    const int Bp = 1 << a.LB, PPW = 32 >> a.LB;
    const int s = lane >> a.LB, b = lane & (Bp - 1);
    const int blk = blockIdx.z;
    const int cx0 = (blockIdx.y * WK + wk) * CXC;
    const bool kvalid = cx0 < a.Cx;
    const int rows = WP * 32;
    const int KS = WK * 64 + 4;

    int* row_dst = reinterpret_cast<int*>(smem);  // [256]
    float* out_s = smem + 256;                    // [256][Z_STRIDE] (output tile / Z tile)
    float* T_s = out_s + 256 * Z_STRIDE;          // DW: [rows][KS]
    float* dW_s = T_s + (size_t)rows * KS;        // DW: [WK*64][CyS]

    if (DW) {
        for (int i = threadIdx.x; i < WK * 64 * a.CyS; i += 256) dW_s[i] = 0.f;
    }

    const int ppt = WP * PPW;
    const int ntiles = (a.Ndst + ppt - 1) / ppt;
    const size_t xrow = (size_t)qc_cp4(a.Cx) * Bp, yrow = (size_t)qc_cp4(a.Cy) * Bp;  // floats per point
    const float* xb = a.X + (size_t)blk * a.Nsrc * xrow;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * ppt + wp * PPW + s;
        const bool pvalid = t < a.Ndst;
        const int dstp = pvalid ? (a.order ? a.order[t] : t) : -1;
        int beg = 0, len = 0;
        if (pvalid && kvalid) {
            beg = a.seg_ptr[dstp];
            len = a.seg_ptr[dstp + 1] - beg;
        }
        const int maxlen = qc_warp_max_i(len);

        float T[CXC][HP];
#pragma unroll
        for (int i = 0; i < CXC; ++i)
#pragma unroll
            for (int h = 0; h < HP; ++h) T[i][h] = 0.f;

        // ---- stage B: segmented gather-accumulate -------------------------------------------
#pragma unroll 2
        for (int it = 0; it < maxlen; ++it) {
            if (it < len) {
                const int q = beg + it;
                const int u = a.src_idx[q];
                const int pp = a.phi_idx ? a.phi_idx[q] : q;
                const float4* php = reinterpret_cast<const float4*>(a.phi + (size_t)pp * HP);
                float ph[HP];
#pragma unroll
                for (int hq = 0; hq < HP / 4; ++hq) {
                    float4 v = __ldg(php + hq);
                    ph[4 * hq + 0] = v.x;
                    ph[4 * hq + 1] = v.y;
                    ph[4 * hq + 2] = v.z;
                    ph[4 * hq + 3] = v.w;
                }
                const float* xr = xb + (size_t)u * xrow;
                float f[CXC];
#pragma unroll
                for (int i = 0; i < CXC; ++i) f[i] = (cx0 + i < a.Cx) ? __ldg(xr + qc_off(cx0 + i, b, Bp)) : 0.f;
#pragma unroll
                for (int i = 0; i < CXC; ++i)
#pragma unroll
                    for (int h = 0; h < HP; ++h) T[i][h] = fmaf(f[i], ph[h], T[i][h]);
            }
        }

        // ---- stage C: dense contraction with the repacked last MLP layer --------------------
        for (int jt = 0; jt < a.Cy; jt += 32) {
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
            if (kvalid) {
#pragma unroll
                for (int i = 0; i < CXC; ++i) {
                    if (cx0 + i < a.Cx) {
#pragma unroll
                        for (int h = 0; h < HP; ++h) {
                            if (h < a.H) {
                                const float4* wrow = reinterpret_cast<const float4*>(
                                    a.W2 + ((size_t)(cx0 + i) * a.H + h) * a.CyP + jt);
                                const float tv = T[i][h];
#pragma unroll
                                for (int jq = 0; jq < 8; ++jq) {
                                    if (jt + jq * 4 < a.CyP) {
                                        const float4 w = __ldg(wrow + jq);
                                        acc[4 * jq + 0] = fmaf(tv, w.x, acc[4 * jq + 0]);
                                        acc[4 * jq + 1] = fmaf(tv, w.y, acc[4 * jq + 1]);
                                        acc[4 * jq + 2] = fmaf(tv, w.z, acc[4 * jq + 2]);
                                        acc[4 * jq + 3] = fmaf(tv, w.w, acc[4 * jq + 3]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            // combine the WK partial sums of each row through shared memory: every warp owns a
            // [32][OUT_STRIDE] slot, the store phase adds the WK slots of a row in fixed order
            // (deterministic, no atomics)
            __syncthreads();
            {
                float* orow = out_s + (size_t)(warp * 32 + lane) * OUT_STRIDE;
#pragma unroll
                for (int j = 0; j < 32; ++j) orow[j] = acc[j];
            }
            if (wk == 0) row_dst[wp * 32 + lane] = dstp;
            __syncthreads();
            for (int idx = threadIdx.x; idx < rows * 32; idx += 256) {
                const int row = idx & (rows - 1), jj = idx / rows;
                const int j = jt + jj;
                const int dp = row_dst[row];
                if (dp >= 0 && j < a.Cy) {
                    const float* src = out_s + (size_t)(((row >> 5) * WK) * 32 + (row & 31)) * OUT_STRIDE + jj;
                    float v = 0.f;
                    for (int k = 0; k < WK; ++k) v += src[(size_t)k * 32 * OUT_STRIDE];
                    float* y = a.Y + ((size_t)blk * a.Ndst + dp) * yrow + qc_off(j, row & (Bp - 1), Bp);
                    if (a.KSPLIT > 1) atomicAdd(y, v);
                    else *y = v;
                }
            }
        }

        // ---- optional: dW2[(c,h)][j] += sum_rows T[row][(c,h)] * Z[row][j] -------------------
        if (DW) {
            __syncthreads();
            float* trow = T_s + (size_t)(wp * 32 + lane) * KS + wk * 64;
#pragma unroll
            for (int i = 0; i < CXC; ++i)
#pragma unroll
                for (int hq = 0; hq < HP / 4; ++hq)
                    *reinterpret_cast<float4*>(trow + i * HP + 4 * hq) =
                        make_float4(T[i][4 * hq], T[i][4 * hq + 1], T[i][4 * hq + 2], T[i][4 * hq + 3]);
            float* Z_s = out_s;
            const int KC = WK * 64;
            const int ntq = (KC / 4) * 8;
            for (int jt = 0; jt < a.Cy; jt += 32) {
                __syncthreads();
                for (int idx = threadIdx.x; idx < rows * 32; idx += 256) {
                    const int row = idx & (rows - 1), jj = idx / rows;
                    const int j = jt + jj;
                    const int dp = row_dst[row];
                    float v = 0.f;
                    if (dp >= 0 && j < a.Cy)
                        v = __ldg(a.Z + ((size_t)blk * a.Ndst + dp) * yrow + qc_off(j, row & (Bp - 1), Bp));
                    Z_s[row * Z_STRIDE + jj] = v;
                }
                __syncthreads();
                for (int tt = threadIdx.x; tt < ntq; tt += 256) {
                    const int kq = tt >> 3, jq = tt & 7;
                    float c[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) c[r][cc] = 0.f;
                    for (int row = 0; row < rows; ++row) {
                        const float4 tv = *reinterpret_cast<const float4*>(T_s + (size_t)row * KS + kq * 4);
                        const float4 zv = *reinterpret_cast<const float4*>(Z_s + row * Z_STRIDE + jq * 4);
                        const float tva[4] = {tv.x, tv.y, tv.z, tv.w};
                        const float zva[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) c[r][cc] = fmaf(tva[r], zva[cc], c[r][cc]);
                    }
                    float* dst = dW_s + (size_t)(kq * 4) * a.CyS + jt + jq * 4;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) dst[r * a.CyS + cc] += c[r][cc];
                }
            }
        }
    }

    if (DW) {
        __syncthreads();
        const int KC = WK * 64;
        for (int idx = threadIdx.x; idx < KC * a.CyS; idx += 256) {
            const int k = idx / a.CyS, j = idx % a.CyS;
            const int wkk = k >> 6, i = (k & 63) / HP, h = k % HP;
            const int cx = (blockIdx.y * WK + wkk) * CXC + i;
            if (cx < a.Cx && h < a.H && j < a.Cy) atomicAdd(&a.dW2[((size_t)cx * a.H + h) * a.CyP + j], dW_s[idx]);
        }
    }
}

// ---------------------------------------------------------------------------------------------


__host__ __device__ constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }

// Reduce v[HP] over the 2^LB lanes of a sub-group.  Afterwards lane holds HP>>NH sums for
// h = hsel*(HP>>NH) + k, replicated over the low (LB-NH) lane bits.
template <int HP, int LB>
__device__ __forceinline__ int subgroup_reduce(float (&v)[HP], int lane)
{
    constexpr int LH = ilog2c(HP);
    constexpr int NH = LB < LH ? LB : LH;
    int hsel = 0;
#pragma unroll
    for (int st = 0; st < LB; ++st) {
        const int mask = 1 << (LB - 1 - st);
        if (st < NH) {
            const int half = HP >> (st + 1);
            const bool upper = (lane & mask) != 0;
#pragma unroll
            for (int k = 0; k < half; ++k) {
                const float send = upper ? v[k] : v[k + half];
                const float keep = upper ? v[k + half] : v[k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
            hsel = (hsel << 1) | (upper ? 1 : 0);
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
        }
    }
    return hsel;
}

template <int HP, int LB>
__global__ void __launch_bounds__(256, 2) k_dphi(const DphiArgs a)
{
    constexpr int CXC = 64 / HP;
    constexpr int Bp = 1 << LB, PPW = 32 >> LB;
    constexpr int LH = ilog2c(HP);
    constexpr int NH = LB < LH ? LB : LH;
    constexpr int NF = HP >> NH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int WK = 1 << a.LWK, WP = 8 >> a.LWK;
    const int wk = warp & (WK - 1), wp = warp >> a.LWK;
    const int s = lane >> LB, b = lane & (Bp - 1);
    const int blk = blockIdx.z;
    const int cx0 = (blockIdx.y * WK + wk) * CXC;
    const bool kvalid = cx0 < a.Cx;
    const int ppt = WP * PPW;
    const int ntiles = (a.Ndst + ppt - 1) / ppt;
    const size_t xrow = (size_t)qc_cp4(a.Cx) * Bp, grow = (size_t)qc_cp4(a.Cg) * Bp;
    const float* xb = a.X + (size_t)blk * a.Nsrc * xrow;
    const bool writer = (lane & ((1 << (LB - NH)) - 1)) == 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * ppt + wp * PPW + s;
        const bool pvalid = t < a.Ndst;
        const int dstp = pvalid ? (a.order ? a.order[t] : t) : -1;
        int beg = 0, len = 0;
        if (pvalid && kvalid) {
            beg = a.seg_ptr[dstp];
            len = a.seg_ptr[dstp + 1] - beg;
        }
        const int maxlen = qc_warp_max_i(len);

        // dT[c][h] = sum_j G[dst][j][b] * W3[j][c][h]
        float dT[CXC][HP];
#pragma unroll
        for (int i = 0; i < CXC; ++i)
#pragma unroll
            for (int h = 0; h < HP; ++h) dT[i][h] = 0.f;
        if (pvalid && kvalid) {
            const float* gp = a.G + ((size_t)blk * a.Ndst + dstp) * grow;
            for (int j = 0; j < a.Cg; ++j) {
                const float g = __ldg(gp + qc_off(j, b, Bp));
                const float4* w3 = reinterpret_cast<const float4*>(a.W3 + ((size_t)j * a.Cx + cx0) * HP);
#pragma unroll
                for (int i = 0; i < CXC; ++i) {
                    if (cx0 + i < a.Cx) {
#pragma unroll
                        for (int hq = 0; hq < HP / 4; ++hq) {
                            const float4 w = __ldg(w3 + i * (HP / 4) + hq);
                            dT[i][4 * hq + 0] = fmaf(g, w.x, dT[i][4 * hq + 0]);
                            dT[i][4 * hq + 1] = fmaf(g, w.y, dT[i][4 * hq + 1]);
                            dT[i][4 * hq + 2] = fmaf(g, w.z, dT[i][4 * hq + 2]);
                            dT[i][4 * hq + 3] = fmaf(g, w.w, dT[i][4 * hq + 3]);
                        }
                    }
                }
            }
        }

        for (int it = 0; it < maxlen; ++it) {
            const bool act = it < len;
            float part[HP];
#pragma unroll
            for (int h = 0; h < HP; ++h) part[h] = 0.f;
            int q = 0;
            if (act) {
                q = beg + it;
                const int u = a.src_idx[q];
                const float* xr = xb + (size_t)u * xrow;
                float f[CXC];
#pragma unroll
                for (int i = 0; i < CXC; ++i) f[i] = (cx0 + i < a.Cx) ? __ldg(xr + qc_off(cx0 + i, b, Bp)) : 0.f;
#pragma unroll
                for (int i = 0; i < CXC; ++i)
#pragma unroll
                    for (int h = 0; h < HP; ++h) part[h] = fmaf(f[i], dT[i][h], part[h]);
            }
            const int hsel = subgroup_reduce<HP, LB>(part, lane);
            if (act && writer) {
                float* dst = a.dphi + (size_t)q * HP + hsel * NF;
#pragma unroll
                for (int k = 0; k < NF; ++k)
                    if (hsel * NF + k < a.H) atomicAdd(dst + k, part[k]);
            }
        }
    }
}

inline int ilog2(int x)
{
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

template <int HP, bool DW>
int launch_contract(const ContractArgs& a, int nblk, size_t smem, cudaStream_t st)
{
    QC_CUDA((qc_set_max_smem<k_contract<HP, DW>>((int)smem)));
    const int WP = 8 >> a.LWK, PPW = 32 >> a.LB;
    const int ntiles = qc_ceil_div(a.Ndst, WP * PPW);
    int gx = qc_num_sms() * (DW ? 1 : 2);
    int per = a.KSPLIT * nblk;
    gx = (gx + per - 1) / per;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, a.KSPLIT, nblk);
    k_contract<HP, DW><<<grid, 256, smem, st>>>(a);
    QC_CHECK_LAUNCH();
    return 0;
}

template <int HP, int LB>
int launch_dphi(const DphiArgs& a, int nksplit, int nblk, cudaStream_t st)
{
    const int WP = 8 >> a.LWK, PPW = 32 >> LB;
    const int ntiles = qc_ceil_div(a.Ndst, WP * PPW);
    int gx = qc_num_sms() * 2;
    int per = nksplit * nblk;
    gx = (gx + per - 1) / per;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, nksplit, nblk);
    k_dphi<HP, LB><<<grid, 256, 0, st>>>(a);
    QC_CHECK_LAUNCH();
    return 0;
}

template <int HP>
int dispatch_dphi(const DphiArgs& a, int LB, int nksplit, int nblk, cudaStream_t st)
{
    switch (LB) {
        case 0: return launch_dphi<HP, 0>(a, nksplit, nblk, st);
        case 1: return launch_dphi<HP, 1>(a, nksplit, nblk, st);
        case 2: return launch_dphi<HP, 2>(a, nksplit, nblk, st);
        case 3: return launch_dphi<HP, 3>(a, nksplit, nblk, st);
        case 4: return launch_dphi<HP, 4>(a, nksplit, nblk, st);
        case 5: return launch_dphi<HP, 5>(a, nksplit, nblk, st);
    }
    return QC_EINVAL;
}

}  // namespace

extern "C" {

size_t qc_contract_workspace_bytes(int Cx, int H, int HP, int Cy, int Ndst, int B)
{
    (void)H;
    return qc_contract_tc_ws_bytes(Ndst, Cx, Cy, HP, B);
}

int qc_contract(const float* Xp, int Nsrc, int Cx, const int* seg_ptr, const int* src_idx,
                const int* phi_idx, const float* phi, int H, int HP, const float* W2, int Cy,
                const int* order, int Ndst, int B, float* Yp, const float* Zp, float* dW2,
                void* workspace, size_t workspace_bytes, void* stream)
{
    if (Nsrc <= 0 || Ndst <= 0 || Cx <= 0 || Cy <= 0 || B <= 0 || H <= 0 || H > HP) return QC_EINVAL;
    if (!(HP == 4 || HP == 8 || HP == 16 || HP == 32)) return QC_EUNSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const bool dw = (Zp != nullptr && dW2 != nullptr);
    const int Bp = qc_bp(B), nblk = qc_nblk(B);
    const int CXC = 64 / HP;
    const int NKC = qc_ceil_div(Cx, CXC);
    ContractArgs a;
    a.X = Xp; a.Y = Yp; a.Z = Zp; a.dW2 = dW2; a.dW_scratch = nullptr; a.dW_counts = nullptr; a.dw_ordered = 0;
    a.ws = workspace; a.ws_bytes = workspace_bytes;
    a.seg_ptr = seg_ptr; a.src_idx = src_idx; a.phi_idx = phi_idx; a.phi = phi; a.W2 = W2; a.order = order;
    a.Nsrc = Nsrc; a.Ndst = Ndst; a.Cx = Cx; a.Cy = Cy; a.CyP = (Cy + 3) & ~3; a.CyS = (Cy + 31) & ~31;
    a.H = H; a.LB = ilog2(Bp);
    int WK = 1;
    while (WK * 2 <= 8 && WK * 2 <= NKC) WK *= 2;
    if (dw) {
        if (a.CyS > 256) return QC_EUNSUP;
        while (WK > 1 && WK * a.CyS > 256) WK /= 2;
    }
    a.LWK = ilog2(WK);
    a.KSPLIT = qc_ceil_div(NKC, WK);
    if (a.KSPLIT > 65535 || nblk > 65535) return QC_EUNSUP;
    {
        int rc = qc_contract_tc(a, HP, B, dw, st);       // tcgen05 dense stage (contract_tc.cu)
        if (rc != 0) return rc == 1 ? 0 : rc;
        rc = qc_contract_fast(a, HP, B, dw, st);         // FP32-pipe fast path (contract_fast.cu)
        if (rc != 0) return rc == 1 ? 0 : rc;
    }
    const int WP = 8 / WK;
    size_t smem = (256 + 256 * Z_STRIDE) * 4;
    if (dw) smem += ((size_t)WP * 32 * (WK * 64 + 4) + (size_t)WK * 64 * a.CyS) * 4;
    if (a.KSPLIT > 1)
        QC_CUDA(cudaMemsetAsync(Yp, 0, (size_t)nblk * Ndst * qc_cp4(Cy) * Bp * 4, st));
#define QC_LC(HPV)                                                            \
    case HPV:                                                                 \
        return dw ? launch_contract<HPV, true>(a, nblk, smem, st)             \
                  : launch_contract<HPV, false>(a, nblk, smem, st);
    switch (HP) {
        QC_LC(4)
        QC_LC(8)
        QC_LC(16)
        QC_LC(32)
    }
#undef QC_LC
    return QC_EUNSUP;
}

int qc_dphi_num_slices(int Nsrc, int Cx, int Cg, int HP, int B)
{
    if (HP <= 0 || B <= 0) return 1;
    if (qc_dphi_fast_ok(Nsrc, Cx, Cg, HP, B)) return qc_ceil_div(Cx, 64 / HP) * qc_nblk(B);
    return 1;
}

int qc_dphi(const float* Xp, int Nsrc, int Cx, const float* Gp, int Cg, const int* seg_ptr,
            const int* src_idx, const float* W3, int H, int HP, const int* order, int Ndst, int B,
            int64_t P, float* dphi, void* stream)
{
    if (Nsrc <= 0 || Ndst <= 0 || Cx <= 0 || Cg <= 0 || B <= 0 || H <= 0 || H > HP) return QC_EINVAL;
    if (!(HP == 4 || HP == 8 || HP == 16 || HP == 32)) return QC_EUNSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const int Bp = qc_bp(B), nblk = qc_nblk(B);
    const int CXC = 64 / HP;
    const int NKC = qc_ceil_div(Cx, CXC);
    DphiArgs a;
    a.X = Xp; a.G = Gp; a.W3 = W3; a.seg_ptr = seg_ptr; a.src_idx = src_idx; a.order = order; a.dphi = dphi;
    a.Nsrc = Nsrc; a.Ndst = Ndst; a.Cx = Cx; a.Cg = Cg; a.H = H; a.P = P;
    int WK = 1;
    while (WK * 2 <= 8 && WK * 2 <= NKC) WK *= 2;
    a.LWK = ilog2(WK);
    const int nksplit = qc_ceil_div(NKC, WK);
    if (nksplit > 65535 || nblk > 65535) return QC_EUNSUP;
    const int LB = ilog2(Bp);
    {
        int rc = qc_dphi_tc(a, HP, B, st);
        if (rc != 0) return rc == 1 ? 0 : rc;
        rc = qc_dphi_fast(a, HP, B, st);
        if (rc != 0) return rc == 1 ? 0 : rc;
    }
    QC_CUDA(cudaMemsetAsync(dphi, 0, (size_t)P * HP * 4, st));   // the generic kernel accumulates with atomics
    switch (HP) {
        case 4: return dispatch_dphi<4>(a, LB, nksplit, nblk, st);
        case 8: return dispatch_dphi<8>(a, LB, nksplit, nblk, st);
        case 16: return dispatch_dphi<16>(a, LB, nksplit, nblk, st);
        case 32: return dispatch_dphi<32>(a, LB, nksplit, nblk, st);
    }
    return QC_EUNSUP;
}

}  // extern "C"
