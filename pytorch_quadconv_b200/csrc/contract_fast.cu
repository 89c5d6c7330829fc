// contract_fast.cu -- fast fp32 path of the fused gather-contract kernels for the shapes the
// headline configurations use (last hidden width H <= 8, <= 32 output channels, batch lanes 8 or 32).
//
// Same factorised math and thread mapping as contract.cu (one thread = one (point, batch) row), but
//   * the K split of the dense stage is a loop IN TIME over chunks of CXC = 64/HP gathered
//     channels instead of a split across warps: the 8 warps of a CTA work on 8 spatially adjacent
//     destination points and the SAME channel chunk, so the rows they gather overlap and hit in L1;
//     the dense-stage partial sums stay in registers (no cross-warp combine, deterministic);
//   * batch lanes are a template parameter, so the CXC gathers of a pair are one 64-bit address
//     plus immediates; indices are prefetched two pairs ahead and feature rows one pair ahead;
//   * the slice of the repacked last MLP layer used by a chunk is staged in shared memory and
//     read as warp-uniform 128-bit broadcasts.
// Reference lines replaced: quadconv.py:143,:91 (last Linear + reshape), :213-227 (_integrate).
#include <cstdlib>
#include "contract_args.cuh"

namespace {

constexpr int TS = 68;  // row stride of the T tile in shared memory (floats, 16-byte aligned rows)
constexpr int ZS = 36;  // row stride of the Z tile

constexpr int PCH = 64;  // pairs staged per chunk (indices + phi rows in shared memory)

template <int HP, int LB, bool DW>
__global__ void __launch_bounds__(256, 1)
k_contract_fast(const ContractArgs a, const int nkc, const int dw_in_smem)
{
    constexpr int CXC = 64 / HP, Bp = 1 << LB, PPW = 32 >> LB;
    constexpr int NG = CXC / 4, HQ = HP / 4;
    constexpr int D = 8;  // gathered pairs in flight per thread
    constexpr int PHS = PCH * HP + 4;  // per-point phi slab stride (padded: distinct banks per sub-group)
    extern __shared__ __align__(16) float smem[];
    float* W2s = smem;                                   // [64][32]
    float* phi_s = W2s + 64 * 32;                        // [8*PPW][PHS]
    int* idx_s = reinterpret_cast<int*>(phi_s + 8 * PPW * PHS);  // [8*PPW][PCH]
    float* T_s = reinterpret_cast<float*>(idx_s + 8 * PPW * PCH);  // DW: [256][TS]
    float* Z_s = T_s + 256 * TS;                         // DW: [256][ZS]
    float* dW_s = Z_s + 256 * ZS;                        // DW && dw_in_smem: [nkc*64][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = lane >> LB, b = lane & (Bp - 1);
    const int blk = blockIdx.y;
    const unsigned xrow = (unsigned)qc_cp4(a.Cx) * Bp;   // floats per source point
    const size_t yrow = (size_t)qc_cp4(a.Cy) * Bp;
    const float* Xb = a.X + (size_t)blk * a.Nsrc * xrow + b * 4;
    constexpr int ppt = 8 * PPW;
    const int ntiles = (a.Ndst + ppt - 1) / ppt;

    if (DW && dw_in_smem)
        for (int i = tid; i < nkc * 64 * 32; i += 256) dW_s[i] = 0.f;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * ppt + warp * PPW + s;
        const bool pvalid = t < a.Ndst;
        const int dstp = pvalid ? (a.order ? a.order[t] : t) : -1;
        int beg = 0, len = 0;
        if (pvalid) {
            beg = a.seg_ptr[dstp];
            len = a.seg_ptr[dstp + 1] - beg;
        }
        const int maxlen = PPW == 1 ? len : qc_warp_max_i(len);
        const int* sidx = a.src_idx + beg;
        const int* pidx = a.phi_idx ? a.phi_idx + beg : nullptr;

        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;

        if (DW) {
            __syncthreads();  // the previous tile's last mini-GEMM is done with Z_s
            float* zr = Z_s + (warp * 32 + lane) * ZS;
            const float* zg = a.Z + ((size_t)blk * a.Ndst + (pvalid ? dstp : 0)) * yrow + b * 4;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pvalid && 4 * g < a.Cy) v = __ldg(reinterpret_cast<const float4*>(zg + (size_t)g * Bp * 4));
                *reinterpret_cast<float4*>(zr + 4 * g) = v;
            }
        }

        for (int kc = 0; kc < nkc; ++kc) {
            const int cx0 = kc * CXC;
            const int nval = min(CXC, a.Cx - cx0);
            const int ngv = (nval + 3) >> 2;  // channel groups of this chunk that exist
            __syncthreads();  // W2s / T_s of the previous chunk are no longer read
            for (int idx = tid; idx < 64 * 32; idx += 256) {
                const int k = idx >> 5, j = idx & 31;
                const int i = k / HP, h = k % HP;
                float v = 0.f;
                if (i < nval && h < a.H && j < a.CyP) v = __ldg(a.W2 + ((size_t)(cx0 + i) * a.H + h) * a.CyP + j);
                W2s[idx] = v;
            }

            // ---- stage B: segmented gather-accumulate, D pairs in flight ------------------------
            float T[CXC][HP];
#pragma unroll
            for (int i = 0; i < CXC; ++i)
#pragma unroll
                for (int h = 0; h < HP; ++h) T[i][h] = 0.f;

            const float* Xc = Xb + (size_t)(cx0 >> 2) * (Bp * 4);
            float* my_phi = phi_s + (warp * PPW + s) * PHS;
            int* my_idx = idx_s + (warp * PPW + s) * PCH;
            float4 buf[D][NG];
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int g = 0; g < NG; ++g) buf[d][g] = make_float4(0.f, 0.f, 0.f, 0.f);
            auto issue = [&](float4 (&xb)[NG], int e) {
                const unsigned u = (unsigned)my_idx[e];
                const float4* xr = reinterpret_cast<const float4*>(Xc + (size_t)(u * xrow));
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    if (g < ngv) xb[g] = __ldg(xr + g * Bp);
            };
            auto consume = [&](const float4 (&xb)[NG], int e) {
                float ph[HP];
                const float4* pr = reinterpret_cast<const float4*>(my_phi + e * HP);
#pragma unroll
                for (int q = 0; q < HQ; ++q) {
                    const float4 v = pr[q];
                    ph[4 * q] = v.x, ph[4 * q + 1] = v.y, ph[4 * q + 2] = v.z, ph[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float f[4] = {xb[g].x, xb[g].y, xb[g].z, xb[g].w};
#pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2)
#pragma unroll
                        for (int h = 0; h < HP; ++h) T[4 * g + e2][h] = fmaf(f[e2], ph[h], T[4 * g + e2][h]);
                }
            };
            for (int c0 = 0; c0 < maxlen; c0 += PCH) {
                const int clen = min(PCH, len - c0);        // this lane's point (may be <= 0)
                const int cmax = min(PCH, maxlen - c0);     // warp-uniform
                __syncwarp();
                // stage the pair indices and phi rows of this chunk (coalesced, once per chunk)
                for (int e = b; e < clen; e += Bp) {
                    const int p = c0 + e;
                    my_idx[e] = sidx[p];
                    const int pp = pidx ? pidx[p] : beg + p;
                    const float4* php = reinterpret_cast<const float4*>(a.phi + (size_t)pp * HP);
#pragma unroll
                    for (int q = 0; q < HQ; ++q) reinterpret_cast<float4*>(my_phi + e * HP)[q] = __ldg(php + q);
                }
                __syncwarp();
#pragma unroll
                for (int d = 0; d < D; ++d)
                    if (d < clen) issue(buf[d], d);
                for (int it = 0; it < cmax; it += D) {
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        if (it + d < clen) {
                            consume(buf[d], it + d);
                            if (it + d + D < clen) issue(buf[d], it + d + D);
                        }
                    }
                }
            }

            __syncthreads();  // W2s staged
            // ---- stage C (partial over this channel chunk): acc[j] += sum_k T[k] * W2s[k][j] ----
#pragma unroll
            for (int i = 0; i < CXC; ++i)
#pragma unroll
                for (int h = 0; h < HP; ++h) {
                    const float tv = T[i][h];
                    const float4* wr = reinterpret_cast<const float4*>(W2s + (i * HP + h) * 32);
#pragma unroll
                    for (int jq = 0; jq < 8; ++jq) {
                        const float4 w = wr[jq];
                        acc[4 * jq + 0] = fmaf(tv, w.x, acc[4 * jq + 0]);
                        acc[4 * jq + 1] = fmaf(tv, w.y, acc[4 * jq + 1]);
                        acc[4 * jq + 2] = fmaf(tv, w.z, acc[4 * jq + 2]);
                        acc[4 * jq + 3] = fmaf(tv, w.w, acc[4 * jq + 3]);
                    }
                }

            if (DW) {
                // dW2[(c,h)][j] += sum_rows T[row][(c,h)] * Z[row][j] for this chunk
                float* tr = T_s + (warp * 32 + lane) * TS;
#pragma unroll
                for (int i = 0; i < CXC; ++i)
#pragma unroll
                    for (int hq = 0; hq < HP / 4; ++hq)
                        *reinterpret_cast<float4*>(tr + i * HP + 4 * hq) =
                            make_float4(T[i][4 * hq], T[i][4 * hq + 1], T[i][4 * hq + 2], T[i][4 * hq + 3]);
                __syncthreads();
                const int half = tid >> 7, tq = tid & 127;
                const int kq = tq >> 3, jq = tq & 7;
                float c[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) c[r][cc] = 0.f;
                const float* tp = T_s + (half * 128) * TS + kq * 4;
                const float* zp = Z_s + (half * 128) * ZS + jq * 4;
#pragma unroll 4
                for (int row = 0; row < 128; ++row) {
                    const float4 tv = *reinterpret_cast<const float4*>(tp + row * TS);
                    const float4 zv = *reinterpret_cast<const float4*>(zp + row * ZS);
                    const float tva[4] = {tv.x, tv.y, tv.z, tv.w};
                    const float zva[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) c[r][cc] = fmaf(tva[r], zva[cc], c[r][cc]);
                }
                if (dw_in_smem) {
                    float* dst = dW_s + ((size_t)kc * 64 + kq * 4) * 32 + jq * 4;
                    if (half == 0) {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) dst[r * 32 + cc] += c[r][cc];
                    }
                    __syncthreads();
                    if (half == 1) {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) dst[r * 32 + cc] += c[r][cc];
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int k = kq * 4 + r, i = k / HP, h = k % HP;
                        if (i < nval && h < a.H) {
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                const int j = jq * 4 + cc;
                                if (j < a.Cy) atomicAdd(a.dW2 + ((size_t)(cx0 + i) * a.H + h) * a.CyP + j, c[r][cc]);
                            }
                        }
                    }
                }
            }
        }

        if (pvalid) {
            float* y = a.Y + ((size_t)blk * a.Ndst + dstp) * yrow + b * 4;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                if (4 * g < a.Cy)
                    *reinterpret_cast<float4*>(y + (size_t)g * Bp * 4) =
                        make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
        }
    }

    if (DW && dw_in_smem) {
        __syncthreads();
        for (int idx = tid; idx < nkc * 64 * 32; idx += 256) {
            const int kc = idx >> 11, k = (idx >> 5) & 63, j = idx & 31;
            const int i = k / HP, h = k % HP;
            const int cx = kc * CXC + i;
            if (cx < a.Cx && h < a.H && j < a.Cy) atomicAdd(a.dW2 + ((size_t)cx * a.H + h) * a.CyP + j, dW_s[idx]);
        }
    }
}

__host__ __device__ constexpr int ilog2f(int x) { return x <= 1 ? 0 : 1 + ilog2f(x >> 1); }

template <int HP, int LB>
__device__ __forceinline__ int subgroup_reduce_f(float (&v)[HP], int lane)
{
    constexpr int LH = ilog2f(HP);
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
__global__ void __launch_bounds__(256, 1)
k_dphi_fast(const DphiArgs a, const int nkc, const size_t slice_stride)
{
    constexpr int CXC = 64 / HP, Bp = 1 << LB, PPW = 32 >> LB;
    constexpr int NG = CXC / 4;
    constexpr int D = 8;
    constexpr int LH = ilog2f(HP);
    constexpr int NH = LB < LH ? LB : LH;
    constexpr int NF = HP >> NH;
    extern __shared__ __align__(16) float smem[];
    float* W3s = smem;  // [32 j][64 k]
    int* idx_s = reinterpret_cast<int*>(W3s + 32 * 64);  // [8*PPW][PCH]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = lane >> LB, b = lane & (Bp - 1);
    const int blk = blockIdx.y;
    const unsigned xrow = (unsigned)qc_cp4(a.Cx) * Bp;
    const size_t grow = (size_t)qc_cp4(a.Cg) * Bp;
    const float* Xb = a.X + (size_t)blk * a.Nsrc * xrow + b * 4;
    constexpr int ppt = 8 * PPW;
    const int ntiles = (a.Ndst + ppt - 1) / ppt;
    const bool writer = (lane & ((1 << (LB - NH)) - 1)) == 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * ppt + warp * PPW + s;
        const bool pvalid = t < a.Ndst;
        const int dstp = pvalid ? (a.order ? a.order[t] : t) : -1;
        int beg = 0, len = 0;
        if (pvalid) {
            beg = a.seg_ptr[dstp];
            len = a.seg_ptr[dstp + 1] - beg;
        }
        const int maxlen = qc_warp_max_i(len);
        const int* sidx = a.src_idx + beg;

        float g[32];
        {
            const float* gp = a.G + ((size_t)blk * a.Ndst + (pvalid ? dstp : 0)) * grow + b * 4;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pvalid && 4 * q < a.Cg) v = __ldg(reinterpret_cast<const float4*>(gp + (size_t)q * Bp * 4));
                g[4 * q] = v.x, g[4 * q + 1] = v.y, g[4 * q + 2] = v.z, g[4 * q + 3] = v.w;
            }
        }

        for (int kc = 0; kc < nkc; ++kc) {
            const int cx0 = kc * CXC;
            const int nval = min(CXC, a.Cx - cx0);
            const int ngv = (nval + 3) >> 2;
            __syncthreads();
            for (int idx = tid; idx < 32 * 64; idx += 256) {
                const int j = idx >> 6, k = idx & 63;
                const int i = k / HP, h = k % HP;
                float v = 0.f;
                if (j < a.Cg && i < nval) v = __ldg(a.W3 + ((size_t)j * a.Cx + cx0 + i) * HP + h);
                W3s[idx] = v;
            }
            __syncthreads();
            // dT[k] = sum_j g[j] * W3s[j][k]
            float dT[CXC][HP];
#pragma unroll
            for (int i = 0; i < CXC; ++i)
#pragma unroll
                for (int h = 0; h < HP; ++h) dT[i][h] = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float gv = g[j];
                const float4* wr = reinterpret_cast<const float4*>(W3s + j * 64);
#pragma unroll
                for (int i = 0; i < CXC; ++i)
#pragma unroll
                    for (int hq = 0; hq < HP / 4; ++hq) {
                        const float4 w = wr[i * (HP / 4) + hq];
                        dT[i][4 * hq + 0] = fmaf(gv, w.x, dT[i][4 * hq + 0]);
                        dT[i][4 * hq + 1] = fmaf(gv, w.y, dT[i][4 * hq + 1]);
                        dT[i][4 * hq + 2] = fmaf(gv, w.z, dT[i][4 * hq + 2]);
                        dT[i][4 * hq + 3] = fmaf(gv, w.w, dT[i][4 * hq + 3]);
                    }
            }

            const float* Xc = Xb + (size_t)(cx0 >> 2) * (Bp * 4);
            int* my_idx = idx_s + (warp * PPW + s) * PCH;
            float4 xbuf[D][NG];
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int q = 0; q < NG; ++q) xbuf[d][q] = make_float4(0.f, 0.f, 0.f, 0.f);
            auto issue = [&](float4 (&xb)[NG], int e) {
                const unsigned u = (unsigned)my_idx[e];
                const float4* xr = reinterpret_cast<const float4*>(Xc + (size_t)(u * xrow));
#pragma unroll
                for (int q = 0; q < NG; ++q)
                    if (q < ngv) xb[q] = __ldg(xr + q * Bp);
            };
            for (int c0 = 0; c0 < maxlen; c0 += PCH) {
                const int clen = min(PCH, len - c0);
                const int cmax = min(PCH, maxlen - c0);
                __syncwarp();
                for (int e = b; e < clen; e += Bp) my_idx[e] = sidx[c0 + e];
                __syncwarp();
#pragma unroll
                for (int d = 0; d < D; ++d)
                    if (d < clen) issue(xbuf[d], d);
                for (int it = 0; it < cmax; it += D) {
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        if (it + d < cmax) {   // warp-uniform: the shuffles below need every lane
                            const bool act = it + d < clen;
                            float part[HP];
#pragma unroll
                            for (int h = 0; h < HP; ++h) part[h] = 0.f;
#pragma unroll
                            for (int q = 0; q < NG; ++q) {
                                const float f[4] = {xbuf[d][q].x, xbuf[d][q].y, xbuf[d][q].z, xbuf[d][q].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e)
#pragma unroll
                                    for (int h = 0; h < HP; ++h) part[h] = fmaf(f[e], dT[4 * q + e][h], part[h]);
                            }
                            if (!act) {
#pragma unroll
                                for (int h = 0; h < HP; ++h) part[h] = 0.f;
                            }
                            const int hsel = subgroup_reduce_f<HP, LB>(part, lane);
                            if (act && writer) {
                                // one slice per (batch block, channel chunk): plain stores, summed by
                                // qc_pair_phi_bwd in fixed order (deterministic, no read-modify-write)
                                float* dst = a.dphi + (size_t)(blk * nkc + kc) * slice_stride +
                                             (size_t)(beg + c0 + it + d) * HP + hsel * NF;
#pragma unroll
                                for (int k = 0; k < NF; ++k) dst[k] = part[k];
                            }
                            if (it + d + D < clen) issue(xbuf[d], it + d + D);
                        }
                    }
                }
            }
        }
    }
}

template <int HP, int LB, bool DW>
int launch_cf(const ContractArgs& a, int B, cudaStream_t st)
{
    constexpr int CXC = 64 / HP, PPW = 32 >> LB;
    const int nkc = qc_ceil_div(a.Cx, CXC);
    const int nblk = qc_nblk(B);
    size_t smem = (64 * 32 + 8 * PPW * (PCH * HP + 4) + 8 * PPW * PCH) * 4;
    int dw_in_smem = 0;
    if (DW) {
        smem += (256 * TS + 256 * ZS) * 4;
        if ((size_t)nkc * 64 * 32 * 4 <= 64 * 1024) {
            dw_in_smem = 1;
            smem += (size_t)nkc * 64 * 32 * 4;
        }
    }
    QC_CUDA((qc_set_max_smem<k_contract_fast<HP, LB, DW>>((int)smem)));
    const int ntiles = qc_ceil_div(a.Ndst, 8 * PPW);
    int gx = qc_num_sms();
    gx = (gx + nblk - 1) / nblk;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, nblk);
    k_contract_fast<HP, LB, DW><<<grid, 256, smem, st>>>(a, nkc, dw_in_smem);
    QC_CHECK_LAUNCH();
    return 1;
}

template <int HP, int LB>
int launch_df(const DphiArgs& a, int B, cudaStream_t st)
{
    constexpr int CXC = 64 / HP, PPW = 32 >> LB;
    const int nkc = qc_ceil_div(a.Cx, CXC);
    const int nblk = qc_nblk(B);
    const size_t smem = (32 * 64 + 8 * PPW * PCH) * 4;
    const int ntiles = qc_ceil_div(a.Ndst, 8 * PPW);
    int gx = qc_num_sms();
    gx = (gx + nblk - 1) / nblk;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, nblk);
    QC_CUDA((qc_set_max_smem<k_dphi_fast<HP, LB>>((int)smem)));
    k_dphi_fast<HP, LB><<<grid, 256, smem, st>>>(a, nkc, (size_t)a.P * HP);
    QC_CHECK_LAUNCH();
    return 1;
}

}  // namespace

int qc_contract_fast(const ContractArgs& a, int HP, int B, bool dw, cudaStream_t st)
{
    const int Bp = qc_bp(B);
    if (a.Cy > 32 || !(HP == 4 || HP == 8) || !(Bp == 8 || Bp == 32)) return 0;
    if ((size_t)a.Nsrc * qc_cp4(a.Cx) * Bp >= ((size_t)1 << 32)) return 0;   // 32-bit row offsets
    if (QC_ENV_SET("QC_NO_FAST")) return 0;
#define QC_F(HPV, LBV) (dw ? launch_cf<HPV, LBV, true>(a, B, st) : launch_cf<HPV, LBV, false>(a, B, st))
    if (HP == 4) return Bp == 8 ? QC_F(4, 3) : QC_F(4, 5);
    return Bp == 8 ? QC_F(8, 3) : QC_F(8, 5);
#undef QC_F
}

bool qc_dphi_fast_ok(int Nsrc, int Cx, int Cg, int HP, int B)
{
    const int Bp = qc_bp(B);
    if (Cg > 32 || !(HP == 4 || HP == 8) || !(Bp == 8 || Bp == 32)) return false;
    if ((size_t)Nsrc * qc_cp4(Cx) * Bp >= ((size_t)1 << 32)) return false;
    if (QC_ENV_SET("QC_NO_FAST")) return false;
    return true;
}

int qc_dphi_fast(const DphiArgs& a, int HP, int B, cudaStream_t st)
{
    const int Bp = qc_bp(B);
    if (a.Cg > 32 || !(HP == 4 || HP == 8) || !(Bp == 8 || Bp == 32)) return 0;
    if ((size_t)a.Nsrc * qc_cp4(a.Cx) * Bp >= ((size_t)1 << 32)) return 0;
    if (QC_ENV_SET("QC_NO_FAST")) return 0;
    if (HP == 4) return Bp == 8 ? launch_df<4, 3>(a, B, st) : launch_df<4, 5>(a, B, st);
    return Bp == 8 ? launch_df<8, 3>(a, B, st) : launch_df<8, 5>(a, B, st);
}
