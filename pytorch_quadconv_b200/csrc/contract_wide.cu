// contract_wide.cu -- the gather stages of the QuadConv contraction for WIDE layers (last hidden
// width H > 8 or more than 32 channels on the dense side), where the weight slices of the fused
// tcgen05 kernels (contract_tc.cu) no longer fit in shared memory / TMEM.
//
// For those shapes the layer is evaluated in the same factorised form, but the intermediate
//     T[b][o][(i,h)] = sum_{p in seg(o)} phi[p][h] * X[src(p)][b][i]            (qc_gather_accumulate)
// is written to HBM once and the dense stage  Y = T * W2  (and, backward, dT = G * W2^T,
// dW2 = T^T * Z) is a plain row-major GEMM done by the caller; the backward per-pair term
//     dphi[p][h] = sum_{b,i} X[src(p)][b][i] * dT[b][o(p)][(i,h)]               (qc_gather_dot)
// reads dT rows back the same way.  T is Cx*H floats per (point, batch) row: 8 KB at 64 channels x
// H = 32, i.e. 0.65 GB at 10k points x batch 8 -- small against the 61 GFMA the stage costs.
//
// Both kernels walk the pair list like the tcgen05 kernels do: one thread = one (point, batch lane)
// row, 64 accumulators per thread (CXC channels x HC <= 16 filter components per pass), gathers
// staged global -> shared by cp.async into a per-warp ring, pair n+1 moved shared -> registers
// under the packed FMAs (fma.rn.f32x2) of pair n.
// Reference lines replaced: quadconv.py:143,:91 (last Linear + reshape), :213-227 (_integrate).
#include <cstdint>
#include "contract_args.cuh"
#include "tc_common.cuh"

namespace {

struct WideArgs {
    const float* X;        // packed features  [nblk][Nsrc][CP4/4][Bp][4]
    const int* seg_ptr;    // [Ndst+1]
    const int* src_idx;    // [P] source point of every pair, grouped by destination
    const int* phi_idx;    // [P] row of phi for every pair (nullptr: identity)
    const float* phi;      // [P][HP]
    const int* order;      // processing order of destination points (nullptr: identity)
    float* T;              // gather_accumulate: out [B][Ndst][Cx*H]
    const float* dT;       // gather_dot: in  [B][Ndst][Cx*H]
    float* dphi;           // gather_dot: out [slices][P][HP]
    int Nsrc, Ndst, Cx, H, HP, B;
    int64_t P;
};

constexpr int PCH = 64;   // pairs staged per chunk of a segment

// Lane mapping shared by both kernels.  A warp works on ONE destination point.  With 32 batch lanes
// (LB = 5) lane = batch entry and a pass covers CXC = 64/HC channels.  With 8 batch lanes (LB = 3)
// the 32 lanes are (channel group g = lane >> 3, batch entry b = lane & 7): a point's packed row
// [c4][b][4] read 512 bytes at a time is exactly that order, so the gathers are the same contiguous
// 512-byte pieces, a pass covers 4x the channels, and the per-pair reduction of qc_gather_dot runs
// over all 32 lanes in both cases.
template <int HC, int LB>
struct Lanes {
    static constexpr int Bp = 1 << LB, G = 32 >> LB;
    static constexpr int CXC = 64 / HC;        // channels per lane and pass
    static constexpr int NG = CXC / 4;         // 4-channel groups per lane and pass
    static constexpr int CW = CXC * G;         // channels per warp and pass
    static constexpr int D = 8 / NG < 2 ? 2 : 8 / NG;   // ring slots (pairs in flight per warp)
    static constexpr int RING = D * NG * 512;
};

// ---------------------------------------------------------------------------------------------
// T[b][o][(i,h)] = sum_p phi[p][h] X[src(p)][b][i]
// ---------------------------------------------------------------------------------------------
template <int HC, int LB, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_gather_acc(const WideArgs a)
{
    using LM = Lanes<HC, LB>;
    constexpr int CXC = LM::CXC, NG = LM::NG, HQ = HC / 4, Bp = LM::Bp, G = LM::G, CW = LM::CW;
    constexpr int D = LM::D, RING = LM::RING;
    constexpr int PHS = PCH * HC + 4, IDS = PCH + 16;
    extern __shared__ __align__(128) float smem[];
    float* phi_s = smem;                                         // [NW][PHS]
    int* idx_s = reinterpret_cast<int*>(phi_s + NW * PHS);       // [NW][IDS]
    float* ring_s = reinterpret_cast<float*>(idx_s + NW * IDS);  // [NW][RING bytes]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // broadcast: the compiler then knows it is warp-uniform
    const int g = lane >> LB, b = lane & (Bp - 1);
    const int blk = blockIdx.y;
    const int bglob = blk * 32 + b;
    const unsigned xrow = (unsigned)qc_cp4(a.Cx) * Bp;
    const float* Xb = a.X + (size_t)blk * a.Nsrc * xrow + lane * 4;
    const int ntiles = (a.Ndst + NW - 1) / NW;
    const int ncg = (a.Cx + CW - 1) / CW, nhh = a.HP / HC;
    const size_t trow = (size_t)a.Cx * a.H;

    for (int i = tid; i < NW * RING / 4; i += NW * 32) ring_s[i] = 0.f;   // stale ring slots are read (and must be finite)
    for (int i = tid; i < NW * IDS; i += NW * 32) idx_s[i] = 0;
    __syncthreads();

    float* my_phi = phi_s + warp * PHS;
    int* my_idx = idx_s + warp * IDS;
    const uint32_t my_idx_sa = tc::smem_u32(my_idx);
    const uint32_t my_ring = tc::smem_u32(ring_s) + warp * RING + lane * 16;
    const float* my_ring_p = ring_s + (warp * RING + lane * 16) / 4;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * NW + warp;
        const bool pvalid = t < a.Ndst;
        const int dstp = pvalid ? (a.order ? a.order[t] : t) : -1;
        int beg = 0, len = 0;
        if (pvalid) {
            beg = a.seg_ptr[dstp];
            len = a.seg_ptr[dstp + 1] - beg;
        }
        const int* sidx = a.src_idx + beg;
        const int* pidx = a.phi_idx ? a.phi_idx + beg : nullptr;
        float* trow_p = (pvalid && bglob < a.B) ? a.T + ((size_t)bglob * a.Ndst + dstp) * trow : nullptr;

        for (int cg = 0; cg < ncg; ++cg) {
            const int cx0 = cg * CW;
            const float* Xc = Xb + (size_t)(cx0 >> 2) * (Bp * 4);
            bool gv[NG];   // this lane's q-th channel group exists
#pragma unroll
            for (int q = 0; q < NG; ++q) gv[q] = cx0 + 4 * (q * G + g) < a.Cx;
            for (int hh = 0; hh < nhh; ++hh) {
                float T[CXC][HC];
#pragma unroll
                for (int i = 0; i < CXC; ++i)
#pragma unroll
                    for (int h = 0; h < HC; ++h) T[i][h] = 0.f;

                auto issue_u = [&](int slot, unsigned u) {
                    const float4* xr = reinterpret_cast<const float4*>(Xc + (size_t)u * xrow);
#pragma unroll
                    for (int q = 0; q < NG; ++q)
                        if (gv[q]) tc::cp_async16(my_ring + (slot * NG + q) * 512, xr + q * 32);
                };
                struct Staged { float4 x[NG]; float4 ph[HQ]; };
                auto load = [&](Staged& r, int slot, int e) {
                    const float4* pr = reinterpret_cast<const float4*>(my_phi + e * HC);
#pragma unroll
                    for (int q = 0; q < HQ; ++q) r.ph[q] = pr[q];
#pragma unroll
                    for (int q = 0; q < NG; ++q)
                        r.x[q] = *reinterpret_cast<const float4*>(my_ring_p + (slot * NG + q) * 128);
                };
                auto fma = [&](const Staged& r) {
                    float ph[HC];
#pragma unroll
                    for (int q = 0; q < HQ; ++q)
                        ph[4 * q] = r.ph[q].x, ph[4 * q + 1] = r.ph[q].y, ph[4 * q + 2] = r.ph[q].z, ph[4 * q + 3] = r.ph[q].w;
#pragma unroll
                    for (int q = 0; q < NG; ++q) {
                        const float f[4] = {r.x[q].x, r.x[q].y, r.x[q].z, r.x[q].w};
#pragma unroll
                        for (int e2 = 0; e2 < 4; e2 += 2)
#pragma unroll
                            for (int h = 0; h < HC; ++h)
                                tc::fma2(T[4 * q + e2][h], T[4 * q + e2 + 1][h], f[e2], f[e2 + 1], ph[h], ph[h]);
                    }
                };
                for (int c0 = 0; c0 < len; c0 += PCH) {
                    const int clen = min(PCH, len - c0);
                    __syncwarp();
                    for (int e = lane; e < clen; e += 32) {
                        const int p = c0 + e;
                        my_idx[e] = sidx[p];
                        const int pp = pidx ? pidx[p] : beg + p;
                        const float4* php = reinterpret_cast<const float4*>(a.phi + (size_t)pp * a.HP + hh * HC);
#pragma unroll
                        for (int q = 0; q < HQ; ++q) reinterpret_cast<float4*>(my_phi + e * HC)[q] = __ldg(php + q);
                    }
                    __syncwarp();
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        if (d < clen) issue_u(d, (unsigned)my_idx[d]);
                        tc::cp_async_commit();
                    }
                    Staged cur, nxt;
                    tc::cp_async_wait<D - 1>();
                    load(cur, 0, 0);
                    for (int it = 0; it < clen; it += D) {
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            tc::cp_async_wait<D - 2>();
                            load(nxt, (d + 1) % D, min(it + d + 1, PCH - 1));
                            const unsigned un = tc::lds_u32_early(my_idx_sa + (it + d + D) * 4);
                            if (it + d < clen) {
                                fma(cur);
                                if (it + d + D < clen) issue_u(d, un);
                            }
                            tc::cp_async_commit();
                            cur = nxt;
                        }
                    }
                }
                // this lane's piece of the row: CXC channels x HC components, HC contiguous floats per channel
                if (trow_p != nullptr) {
                    const bool vec = (a.H & 3) == 0 && (reinterpret_cast<uintptr_t>(a.T) & 15) == 0;
#pragma unroll
                    for (int i = 0; i < CXC; ++i) {
                        const int c = cx0 + 4 * ((i >> 2) * G + g) + (i & 3);
                        if (c < a.Cx) {
                            float* dst = trow_p + (size_t)c * a.H + hh * HC;
                            if (vec) {
#pragma unroll
                                for (int q = 0; q < HQ; ++q)
                                    if (hh * HC + 4 * q < a.H)
                                        *reinterpret_cast<float4*>(dst + 4 * q) =
                                            make_float4(T[i][4 * q], T[i][4 * q + 1], T[i][4 * q + 2], T[i][4 * q + 3]);
                            } else {
#pragma unroll
                                for (int h = 0; h < HC; ++h)
                                    if (hh * HC + h < a.H) dst[h] = T[i][h];
                            }
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dphi[slice = blk][p][h] = sum_{b in blk, i} X[src(p)][b][i] dT[b][o(p)][(i,h)]
// Two pairs per step: lane bit 4 picks the pair a half-warp ends up with, the low four lane bits
// reduce it (k_dphi_tc's butterfly: the h slots of dT are permuted per lane once per pass so that the
// halving exchanges always keep the low slots).  A pair's value is owned by the same lane in every
// channel pass, so the passes after the first accumulate with a plain read-modify-write whose read
// is issued before the FMAs.
// ---------------------------------------------------------------------------------------------
template <int HC, int LB, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_gather_dot(const WideArgs a, const size_t slice_stride)
{
    using LM = Lanes<HC, LB>;
    constexpr int CXC = LM::CXC, NG = LM::NG, Bp = LM::Bp, G = LM::G, CW = LM::CW;
    constexpr int D = LM::D, RING = LM::RING;
    constexpr int IDS = PCH + 16;
    constexpr int LH = HC == 16 ? 4 : (HC == 8 ? 3 : 2);
    constexpr int LBr = 4;
    constexpr int NH = LBr < LH ? LBr : LH;
    constexpr int NF = HC >> NH;
    static_assert(NF == 1, "the butterfly ends with one value per lane");
    static_assert(D % 2 == 0, "two-pair steps walk the ring two slots at a time");
    extern __shared__ __align__(128) float smem[];
    int* idx_s = reinterpret_cast<int*>(smem);                     // [NW][IDS]
    float* ring_s = reinterpret_cast<float*>(idx_s + NW * IDS);    // [NW][RING bytes]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // broadcast: the compiler then knows it is warp-uniform
    const int g = lane >> LB, b = lane & (Bp - 1);
    const int blk = blockIdx.y;
    const int bglob = blk * 32 + b;
    const unsigned xrow = (unsigned)qc_cp4(a.Cx) * Bp;
    const float* Xb = a.X + (size_t)blk * a.Nsrc * xrow + lane * 4;
    const int ntiles = (a.Ndst + NW - 1) / NW;
    const int ncg = (a.Cx + CW - 1) / CW, nhh = a.HP / HC;
    const size_t trow = (size_t)a.Cx * a.H;
    const bool writer = (lane & ((1 << (LBr - NH)) - 1)) == 0;
    const int hsel = (lane >> (LBr - NH)) & ((1 << NH) - 1);
    const int up = (lane >> 4) & 1;
    float* dslice = a.dphi + (size_t)blk * slice_stride;

    for (int i = tid; i < NW * RING / 4; i += NW * 32) ring_s[i] = 0.f;
    for (int i = tid; i < NW * IDS; i += NW * 32) idx_s[i] = 0;
    __syncthreads();

    int* my_idx = idx_s + warp * IDS;
    const uint32_t my_idx_sa = tc::smem_u32(my_idx);
    const uint32_t my_ring = tc::smem_u32(ring_s) + warp * RING + lane * 16;
    const float* my_ring_p = ring_s + (warp * RING + lane * 16) / 4;
    const float* own_p = my_ring_p + up * (NG * 128);
    const float* oth_p = my_ring_p + (1 - up) * (NG * 128);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * NW + warp;
        const bool pvalid = t < a.Ndst;
        const int dstp = pvalid ? (a.order ? a.order[t] : t) : -1;
        int beg = 0, len = 0;
        if (pvalid) {
            beg = a.seg_ptr[dstp];
            len = a.seg_ptr[dstp + 1] - beg;
        }
        const int* sidx = a.src_idx + beg;
        const float* drow = (pvalid && bglob < a.B) ? a.dT + ((size_t)bglob * a.Ndst + dstp) * trow : nullptr;

        for (int cg = 0; cg < ncg; ++cg) {
            const int cx0 = cg * CW;
            const float* Xc = Xb + (size_t)(cx0 >> 2) * (Bp * 4);
            const bool first = cg == 0;
            bool gv[NG];
#pragma unroll
            for (int q = 0; q < NG; ++q) gv[q] = cx0 + 4 * (q * G + g) < a.Cx;
            for (int hh = 0; hh < nhh; ++hh) {
                // dT piece of this lane: CXC channels x HC components
                float dT[CXC][HC];
                {
                    const bool vec = (a.H & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dT) & 15) == 0;
#pragma unroll
                    for (int i = 0; i < CXC; ++i) {
                        const int c = cx0 + 4 * ((i >> 2) * G + g) + (i & 3);
#pragma unroll
                        for (int h = 0; h < HC; ++h) dT[i][h] = 0.f;
                        if (drow != nullptr && c < a.Cx) {
                            const float* src = drow + (size_t)c * a.H + hh * HC;
                            if (vec) {
#pragma unroll
                                for (int q = 0; q < HC / 4; ++q)
                                    if (hh * HC + 4 * q < a.H) {
                                        const float4 v = __ldg(reinterpret_cast<const float4*>(src + 4 * q));
                                        dT[i][4 * q] = v.x, dT[i][4 * q + 1] = v.y, dT[i][4 * q + 2] = v.z, dT[i][4 * q + 3] = v.w;
                                    }
                            } else {
#pragma unroll
                                for (int h = 0; h < HC; ++h)
                                    if (hh * HC + h < a.H) dT[i][h] = __ldg(src + h);
                            }
                        }
                    }
                    // slot k <-> component k ^ hsel: the butterfly keeps low slots, sends high ones
#pragma unroll
                    for (int st = 0; st < NH; ++st) {
                        const int bit = HC >> (st + 1);
                        const bool upper = (lane & (1 << (LBr - 1 - st))) != 0;
#pragma unroll
                        for (int c = 0; c < CXC; ++c)
#pragma unroll
                            for (int k = 0; k < HC; ++k)
                                if (!(k & bit)) {
                                    const float x = dT[c][k], y = dT[c][k | bit];
                                    dT[c][k] = upper ? y : x;
                                    dT[c][k | bit] = upper ? x : y;
                                }
                    }
                }
                auto issue_u = [&](int slot, unsigned u) {
                    const float4* xr = reinterpret_cast<const float4*>(Xc + (size_t)u * xrow);
#pragma unroll
                    for (int q = 0; q < NG; ++q)
                        if (gv[q]) tc::cp_async16(my_ring + (slot * NG + q) * 512, xr + q * 32);
                };
                for (int c0 = 0; c0 < len; c0 += PCH) {
                    const int clen = min(PCH, len - c0);
                    __syncwarp();
                    for (int e = lane; e < clen; e += 32) my_idx[e] = sidx[c0 + e];
                    __syncwarp();
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        if (d < clen) issue_u(d, (unsigned)my_idx[d]);
                        tc::cp_async_commit();
                    }
                    for (int it = 0; it < clen; it += D) {
#pragma unroll
                        for (int d = 0; d < D; d += 2) {
                            tc::cp_async_wait<D - 2>();       // slots d and d+1 have landed
                            if (it + d < clen) {
                                float4 xa[NG], xb[NG];
#pragma unroll
                                for (int q = 0; q < NG; ++q) {
                                    xa[q] = *reinterpret_cast<const float4*>(own_p + (d * NG + q) * 128);
                                    xb[q] = *reinterpret_cast<const float4*>(oth_p + (d * NG + q) * 128);
                                }
                                const unsigned u0 = tc::lds_u32_early(my_idx_sa + (it + d + D) * 4), u1 = tc::lds_u32_early(my_idx_sa + (it + d + 1 + D) * 4);
                                const bool wr = writer && it + d + up < clen;
                                float* dst = dslice + (size_t)(beg + c0 + it + d + up) * a.HP + hh * HC + hsel;
                                const float old = (wr && !first) ? *dst : 0.f;
                                float acc0[HC], acc1[HC];
#pragma unroll
                                for (int h = 0; h < HC; ++h) acc0[h] = 0.f, acc1[h] = 0.f;
#pragma unroll
                                for (int q = 0; q < NG; ++q) {
                                    const float fa[4] = {xa[q].x, xa[q].y, xa[q].z, xa[q].w};
                                    const float fb[4] = {xb[q].x, xb[q].y, xb[q].z, xb[q].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
#pragma unroll
                                        for (int h = 0; h < HC; h += 2) {
                                            tc::fma2(acc0[h], acc0[h + 1], dT[4 * q + e][h], dT[4 * q + e][h + 1], fa[e], fa[e]);
                                            tc::fma2(acc1[h], acc1[h + 1], dT[4 * q + e][h], dT[4 * q + e][h + 1], fb[e], fb[e]);
                                        }
                                }
#pragma unroll
                                for (int k = 0; k < HC; ++k) acc0[k] += __shfl_xor_sync(0xffffffffu, acc1[k], 16);
#pragma unroll
                                for (int st = 0; st < LBr; ++st) {
                                    const int mask = 1 << (LBr - 1 - st);
                                    if (st < NH) {
                                        const int half = HC >> (st + 1);
#pragma unroll
                                        for (int k = 0; k < half; ++k) acc0[k] += __shfl_xor_sync(0xffffffffu, acc0[k + half], mask);
                                    } else {
                                        acc0[0] += __shfl_xor_sync(0xffffffffu, acc0[0], mask);
                                    }
                                }
                                if (wr) *dst = old + acc0[0];
                                if (it + d + D < clen) issue_u(d, u0);
                                tc::cp_async_commit();
                                if (it + d + 1 + D < clen) issue_u(d + 1, u1);
                                tc::cp_async_commit();
                            } else {
                                tc::cp_async_commit();
                                tc::cp_async_commit();
                            }
                        }
                    }
                }
            }
        }
    }
}

template <int HC, int LB> constexpr size_t acc_smem(int NW)
{
    return (size_t)NW * (PCH * HC + 4) * 4 + (size_t)NW * (PCH + 16) * 4 + (size_t)NW * Lanes<HC, LB>::RING;
}
template <int HC, int LB> constexpr size_t dot_smem(int NW)
{
    return (size_t)NW * (PCH + 16) * 4 + (size_t)NW * Lanes<HC, LB>::RING;
}

template <int HC, int LB, int NW>
int launch_acc(const WideArgs& a, cudaStream_t st)
{
    const size_t smem = acc_smem<HC, LB>(NW);
    QC_CUDA((qc_set_max_smem<k_gather_acc<HC, LB, NW>>((int)smem)));
    const int nblk = qc_nblk(a.B);
    const int ntiles = qc_ceil_div(a.Ndst, NW);
    int gx = qc_ceil_div(qc_num_sms(), nblk);
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    k_gather_acc<HC, LB, NW><<<dim3(gx, nblk), NW * 32, smem, st>>>(a);
    QC_CHECK_LAUNCH();
    return 0;
}

template <int HC, int LB, int NW>
int launch_dot(const WideArgs& a, cudaStream_t st)
{
    const size_t smem = dot_smem<HC, LB>(NW);
    QC_CUDA((qc_set_max_smem<k_gather_dot<HC, LB, NW>>((int)smem)));
    const int nblk = qc_nblk(a.B);
    const int ntiles = qc_ceil_div(a.Ndst, NW);
    int gx = qc_ceil_div(qc_num_sms(), nblk);
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    k_gather_dot<HC, LB, NW><<<dim3(gx, nblk), NW * 32, smem, st>>>(a, (size_t)a.P * a.HP);
    QC_CHECK_LAUNCH();
    return 0;
}

bool wide_ok(int Cx, int H, int HP, int B)
{
    const int Bp = qc_bp(B);
    return (HP == 4 || HP == 8 || HP == 16 || HP == 32) && H > 0 && H <= HP && (Bp == 8 || Bp == 32) && Cx > 0;
}

}  // namespace

extern "C" {

int qc_gather_supported(int Cx, int H, int HP, int B) { return wide_ok(Cx, H, HP, B) ? 1 : 0; }

int qc_gather_dot_num_slices(int Cx, int HP, int B)
{
    (void)Cx; (void)HP;
    return qc_nblk(B);
}

int qc_gather_accumulate(const float* Xp, int Nsrc, int Cx, const int* seg_ptr, const int* src_idx,
                         const int* phi_idx, const float* phi, int H, int HP, const int* order, int Ndst,
                         int B, float* T, void* stream)
{
    if (Nsrc <= 0 || Ndst <= 0 || B <= 0) return QC_EINVAL;
    if (!wide_ok(Cx, H, HP, B)) return QC_EUNSUP;
    WideArgs a{};
    a.X = Xp; a.seg_ptr = seg_ptr; a.src_idx = src_idx; a.phi_idx = phi_idx; a.phi = phi; a.order = order;
    a.T = T; a.Nsrc = Nsrc; a.Ndst = Ndst; a.Cx = Cx; a.H = H; a.HP = HP; a.B = B; a.P = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool b32 = qc_bp(B) == 32;
    switch (HP) {
        case 4: return b32 ? launch_acc<4, 5, 12>(a, st) : launch_acc<4, 3, 12>(a, st);
        case 8: return b32 ? launch_acc<8, 5, 12>(a, st) : launch_acc<8, 3, 12>(a, st);
        default: return b32 ? launch_acc<16, 5, 12>(a, st) : launch_acc<16, 3, 12>(a, st);
    }
}

int qc_gather_dot(const float* Xp, int Nsrc, int Cx, const float* dT, const int* seg_ptr, const int* src_idx,
                  int H, int HP, const int* order, int Ndst, int B, int64_t P, float* dphi, void* stream)
{
    if (Nsrc <= 0 || Ndst <= 0 || B <= 0 || P < 0) return QC_EINVAL;
    if (!wide_ok(Cx, H, HP, B)) return QC_EUNSUP;
    WideArgs a{};
    a.X = Xp; a.seg_ptr = seg_ptr; a.src_idx = src_idx; a.order = order; a.dT = dT; a.dphi = dphi;
    a.Nsrc = Nsrc; a.Ndst = Ndst; a.Cx = Cx; a.H = H; a.HP = HP; a.B = B; a.P = P;
    cudaStream_t st = (cudaStream_t)stream;
    const bool b32 = qc_bp(B) == 32;
    switch (HP) {
        case 4: return b32 ? launch_dot<4, 5, 12>(a, st) : launch_dot<4, 3, 12>(a, st);
        case 8: return b32 ? launch_dot<8, 5, 12>(a, st) : launch_dot<8, 3, 12>(a, st);
        default: return b32 ? launch_dot<16, 5, 12>(a, st) : launch_dot<16, 3, 12>(a, st);
    }
}

}  // extern "C"
