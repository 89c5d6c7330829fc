// dphi_mma.cu -- reduced-precision tensor-core path of d(phi): the gradient of the per-pair filter activations,
//     dphi[p][h] = sum_{b,i} f[b,i,in(p)] * dT[out(p); b,i,h],      dT[o; b,i,h] = sum_j g[b,j,o] * W_last[i*Co+j][h]
// (autograd of torch_quadconv/quadconv.py:216-219 through the factorised form of DESIGN.md section 3; the fp32 twin
// is k_dphi_tc in contract_tc.cu).  Same tiles as contract_mma.cu -- 16 destination points and the union U of their
// neighbours -- and the same gathered rows, used the other way round (a sampled dense-dense product):
//
//   dense    dT[(o,b)][(h,i)] = G[(o,b)][j] * W3[(h,i)][j]            per channel chunk, G = the tile's own grad rows
//   regroup  fp32 accumulator -> bf16, stored as the K-major operand A2[(o,h)][(b,i)]
//   sddmm    D[(o,h)][u] += A2[(o,h)][(b,i)] * X[u][(b,i)]            for every gathered union row u (K = (b,i), the
//            gathered stage is the K-major B operand: the same bytes contract_mma.cu reads MN-major), accumulated
//            over the channel chunks in TMEM (one column per union slot)
//   extract  dphi[p][h] = D[(out(p),h)][slot(p)] for the tile's pairs (the other entries of D are not needed)
//
// Warp roles (448 threads, 1 CTA per SM, persistent): warps 0-3 and 10-13 regroup / extract, 4-7 producers (G tile +
// gather ring, cp.async tracked by mbarriers), warp 8 issues the dense MMAs, warp 9 the sddmm MMAs.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"
#include <cstring>

extern "C" int qc_mma_get_variant(void);

namespace {

constexpr int TP = 16;                 // destination points per tile
constexpr int KST = 32;                // union rows per ring stage
constexpr int NST = 8;                 // ring stages
constexpr uint32_t A2_K8 = 144;        // bytes between groups of 8 k in the sddmm A operand (16 B pad: bank spread)
constexpr int NTHREADS = 448;
constexpr int DT_COL = 384;            // TMEM: sddmm accumulator in columns [0, max_ku), dense accumulators from here
constexpr int STG_LD = 33;             // floats per row of the extraction staging tile

struct DphiMmaArgs {
    const __nv_bfloat16* X;     // packed source features [nblk][Nsrc][nch][Bp][CC]
    const __nv_bfloat16* G;     // packed grad_out rows of the destination points [nblk][Ndst][nchg][Bp][CCg]
    float* dphi;                // [P][HPphi] fp32 (zero-initialised by the caller when nblk > 1)
    const int* order;           // processing order of the destination points
    const int2* tile_ent;       // {pair, (point in tile << 16) | union slot} grouped by tile
    const int* tile_pptr;
    const int* tile_uptr;
    const int* union_idx;
    const float* w_last;        // [Ci*Co][H]
    int Nsrc, Ndst, Cs, Cg, H, HPphi, nch, nchg, CCg, ntiles, nblk, max_ku;
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__host__ __device__ constexpr size_t dphi_smem_bytes(int CC, int Bp, int nch, int CgP, int max_ku)
{
    const int NC = Bp * CC, rows = TP * Bp;
    return (size_t)rows * CgP * 2                       // G tile
           + (size_t)nch * 8 * CC * CgP * 2             // W3, all chunks
           + 2 * (size_t)TP * (NC / 8) * A2_K8          // A2, two buffers
           + (size_t)NST * KST * NC * 2                 // gather ring (128-byte-swizzle rows)
           + 2 * (size_t)128 * STG_LD * 4               // extraction staging, one per regroup group
           + (size_t)max_ku * 8 + 64 * 8 + 1024;        // union lists, barriers, alignment slack
}

// TMA = true: the union rows arrive by TMA gather4 (tma.cuh; xmap = X as [nblk * Nsrc][nch * NC] bf16), issued by the
// lanes of one producer warp, instead of per-lane cp.async copies
template <int CC, int LB, bool TMA>
__global__ void __launch_bounds__(NTHREADS, 1) k_dphi_mma(const DphiMmaArgs a, const __grid_constant__ CUtensorMap xmap)
{
    constexpr int Bp = 1 << LB, NC = Bp * CC;
    static_assert(NC % 64 == 0, "the gather ring uses 128-byte rows");
    constexpr int rows = TP * Bp, MBC = rows / 128;
    constexpr int ND = 8 * CC;                            // N of the dense stage = (h, channel in chunk)
    constexpr uint32_t A2SB = (NC / 8) * A2_K8;           // bytes between groups of 8 rows (= points) of A2
    constexpr uint32_t A2BUF = TP * A2SB;
    constexpr uint32_t XBLK = KST * 128, STAGE = (NC / 64) * XBLK;
    const int nch = a.nch;
    const int CgP = (a.Cg + 15) & ~15;                    // K of the dense stage
    const uint32_t GSB = (uint32_t)(CgP / 8) * 128;       // bytes between groups of 8 rows of the G tile
    const uint32_t W3CH = (uint32_t)ND * CgP * 2;

    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* x_s = smem;                                           // [NST][NC/64][KST][128 B]
    unsigned char* g_s = x_s + NST * STAGE;                              // [rows/8][CgP/8][8 rows][8 j] bf16 (K-major)
    unsigned char* w3_s = g_s + (size_t)rows * CgP * 2;                  // [nch][ND/8][CgP/8][8 n][8 j]
    unsigned char* a2_s = w3_s + (size_t)nch * W3CH;                     // [2][16][NC/8 k-groups x 144]
    float* stg = reinterpret_cast<float*>(a2_s + 2 * A2BUF);            // [2][128][STG_LD]
    int* uidx_s = reinterpret_cast<int*>(stg + 2 * 128 * STG_LD);        // [2][max_ku]
    uint64_t* bars = reinterpret_cast<uint64_t*>(uidx_s + 2 * a.max_ku);
    uint64_t* x_full = bars;                 // [NST] producers (128 copy completions) -> sddmm issuer
    uint64_t* x_empty = x_full + NST;        // [NST] sddmm commit -> producers
    uint64_t* a2_full = x_empty + NST;       // [2]   regroup (8 warps) -> sddmm issuer
    uint64_t* a2_empty = a2_full + 2;        // [2]   sddmm commit -> regroup
    uint64_t* g_full = a2_empty + 2;         // producers (128 copy completions) -> dense issuer
    uint64_t* g_empty = g_full + 1;          // dense commit -> producers
    uint64_t* dt_full = g_empty + 1;         // dense commit -> regroup
    uint64_t* dt_empty = dt_full + 1;        // regroup (8) -> dense issuer
    uint64_t* dp_full = dt_empty + 1;        // sddmm commit -> extraction
    uint64_t* dp_empty = dp_full + 1;        // extraction (8) -> sddmm issuer
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(dp_empty + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    if (warp == 0) tc::tmem_alloc(tmem_base_s, 512);
    if (tid == 32) {
        for (int i = 0; i < NST; ++i) tc::mbar_init(&x_full[i], TMA ? 1 : 128), tc::mbar_init(&x_empty[i], 1);
        for (int i = 0; i < 2; ++i) tc::mbar_init(&a2_full[i], 8), tc::mbar_init(&a2_empty[i], 1);
        tc::mbar_init(g_full, 128), tc::mbar_init(g_empty, 1);
        tc::mbar_init(dt_full, 1), tc::mbar_init(dt_empty, 8);
        tc::mbar_init(dp_full, 1), tc::mbar_init(dp_empty, 8);
        tc::mbar_fence_init();
    }
    // W3[(h, channel in chunk)][j] = W_last[(i*Co + j)][h], bf16, K-major, per channel chunk
    {
        const int total = nch * ND * CgP;
        for (int idx = tid; idx < total; idx += NTHREADS) {
            const int c = idx / (ND * CgP), r = idx % (ND * CgP);
            const int n = r / CgP, j = r % CgP;
            const int h = n / CC, i = c * CC + n % CC;
            float v = 0.f;
            if (i < a.Cs && j < a.Cg && h < a.H) v = __ldg(a.w_last + ((size_t)i * a.Cg + j) * a.H + h);
            *reinterpret_cast<__nv_bfloat16*>(w3_s + (size_t)c * W3CH + tc::kmajor16_off(n, j, GSB, 128)) = __float2bfloat16_rn(v);
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_base_s;
    const int nwork = a.ntiles * a.nblk;

    if (warp >= 4 && warp < 8) {
        // ===== producers: the tile's own grad rows, then the gather ring ========================================
        const int pt = tid - 128;
        const uint32_t x_sa = tc::smem_u32(x_s), g_sa = tc::smem_u32(g_s);
        const size_t srow = (size_t)nch * NC;
        const size_t grow = (size_t)a.nchg * Bp * a.CCg;    // elements per destination point of G
        uint32_t xit = 0, tt = 0;
        for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
            const int tile = wi % a.ntiles, blk = wi / a.ntiles;
            const int u0 = a.tile_uptr[tile], ku = a.tile_uptr[tile + 1] - u0, nst = ku / KST;
            int* uix = uidx_s + (tt & 1) * a.max_ku;
            const uint32_t rowbytes = (uint32_t)(nch * NC * 2);
            for (int i = pt; i < ku; i += 128) {
                const int u = __ldg(a.union_idx + u0 + i);
                if (TMA) uix[i] = u >= 0 ? blk * a.Nsrc + u : a.nblk * a.Nsrc;   // row of the 2-D view; past the end = zero fill
                else uix[i] = u >= 0 ? (int)((uint32_t)u * rowbytes) : -1;      // byte offset of the row (< 4 GB, host-checked)
            }
            tc::named_bar_sync(2, 128);
            // G tile: row = (point in tile, batch lane), K-major over the grad channels; pieces of min(16, 2*CCg) bytes
            tc::mbar_wait(g_empty, (tt & 1) ^ 1);
            {
                const __nv_bfloat16* Gb = a.G + (size_t)blk * a.Ndst * grow;
                const int pb = a.CCg >= 8 ? 16 : 8;                 // bytes per piece
                const int ppr = CgP * 2 / pb;                        // pieces per row
                for (int idx = pt; idx < rows * ppr; idx += 128) {
                    const int row = idx / ppr, pc = idx % ppr;
                    const int ol = row / Bp, b = row % Bp;
                    const int j0 = pc * (pb / 2);                    // first channel of the piece
                    const int t = tile * TP + ol;
                    const bool valid = t < a.Ndst && j0 < a.nchg * a.CCg;
                    const int dstp = valid ? (a.order ? a.order[t] : t) : 0;
                    const __nv_bfloat16* src = Gb + (size_t)dstp * grow + ((size_t)(j0 / a.CCg) * Bp + b) * a.CCg + j0 % a.CCg;
                    const uint32_t dst = g_sa + tc::kmajor16_off(row, j0, GSB, 128);
                    if (pb == 16) tc::cp_async16_zfill(dst, src, valid);
                    else tc::cp_async8_zfill(dst, src, valid);
                }
                tc::cp_async_mbar_arrive_noinc(g_full);
            }
            const char* Xbytes = reinterpret_cast<const char*>(a.X + (size_t)blk * a.Nsrc * srow);
            constexpr int NJ = KST * (NC / 64) / 16;
            const int cidx = pt & 7, l0 = pt >> 3;
            for (int c = 0; c < nch; ++c)
                for (int st = 0; st < nst; ++st) {
                    const uint32_t s = xit % NST, ph = (xit / NST) & 1;
                    // byte offsets of this stage's rows, loaded together before the wait (see contract_mma.cu)
                    const uint32_t* uo = reinterpret_cast<const uint32_t*>(uix) + st * KST;
                    if constexpr (TMA) {
                        if (pt < 32) {
                            constexpr int NI = (KST / 4) * (NC / 64);          // gather4 instructions per stage
                            const int rg = pt % (KST / 4), b64 = pt / (KST / 4);
                            int4 r = make_int4(0, 0, 0, 0);
                            if (pt < NI) r = *reinterpret_cast<const int4*>(uo + rg * 4);
                            tc::mbar_wait(&x_empty[s], ph ^ 1);
                            if (pt == 0) tma::mbar_arrive_expect_tx(&x_full[s], STAGE);
                            __syncwarp();
                            if (pt < NI)
                                tma::gather4(x_sa + s * STAGE + b64 * XBLK + rg * 512, &xmap, &x_full[s], c * NC + b64 * 64, r.x,
                                             r.y, r.z, r.w);
                        }
                        ++xit;
                        continue;
                    }
                    uint32_t off[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) off[j] = uo[(l0 + 16 * j) % KST];
                    const char* base = Xbytes + (size_t)c * NC * 2 + cidx * 16;
                    const uint32_t dbase = x_sa + s * STAGE + ((l0 % KST) * 128) + ((cidx ^ (l0 & 7)) << 4);
                    tc::mbar_wait(&x_empty[s], ph ^ 1);
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int line = l0 + 16 * j;
                        const bool valid = off[j] != 0xffffffffu;
                        tc::cp_async16_zfill(dbase + (line / KST) * XBLK + ((line % KST) - (l0 % KST)) * 128,
                                             base + (line / KST) * 128 + (valid ? off[j] : 0u), valid);
                    }
                    tc::cp_async_mbar_arrive_noinc(&x_full[s]);
                    ++xit;
                }
        }
    } else if (warp == 8) {
        // ===== dense issuer: dT chunk = G tile x W3 chunk ========================================================
        const uint32_t idD = tc::make_idesc_bf16(128, ND, false, false);
        const uint64_t da0 = tc::make_smem_desc(tc::smem_u32(g_s), 128, GSB);
        const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(w3_s), 128, GSB);
        uint32_t cc = 0, tt = 0;
        for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
            tc::mbar_wait(g_full, tt & 1);
            tc::fence_after_sync();
            for (int c = 0; c < nch; ++c, ++cc) {
                tc::mbar_wait(dt_empty, (cc & 1) ^ 1);           // the previous chunk has left the dense accumulators
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    for (int mb = 0; mb < MBC; ++mb)
                        for (int ks = 0; ks < CgP / 16; ++ks)
                            tc::mma_bf16_ss(tmem + DT_COL + mb * ND, da0 + (uint64_t)((mb * 16 * GSB + ks * 256) >> 4),
                                            db0 + (uint64_t)((c * W3CH + ks * 256) >> 4), idD, ks > 0);
                    tc::mma_commit(dt_full);
                }
                __syncwarp();
            }
            if (tc::elect_one()) tc::mma_commit(g_empty);
            __syncwarp();
        }
    } else if (warp == 9) {
        // ===== sddmm issuer: D[(o,h)][slot] += A2 chunk x gathered rows^T ========================================
        const uint32_t idS = tc::make_idesc_bf16(128, KST, false, false);
        const uint64_t da0 = tc::make_smem_desc(tc::smem_u32(a2_s), A2_K8, A2SB);
        const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(x_s), 16, 1024) | ((uint64_t)2 << 61);   // K-major, 128 B swizzle
        uint32_t xit = 0, cc = 0, tt = 0;
        for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
            const int tile = wi % a.ntiles;
            const int nst = (__ldg(a.tile_uptr + tile + 1) - __ldg(a.tile_uptr + tile)) / KST;
            tc::mbar_wait(dp_empty, (tt & 1) ^ 1);               // the previous tile's values have been extracted
            tc::fence_after_sync();
            for (int c = 0; c < nch; ++c, ++cc) {
                const uint32_t buf = cc & 1;
                tc::mbar_wait(&a2_full[buf], (cc >> 1) & 1);
                tc::fence_after_sync();
                const uint64_t dac = da0 + (uint64_t)((buf * A2BUF) >> 4);
                for (int st = 0; st < nst; ++st, ++xit) {
                    const uint32_t s = xit % NST, ph = (xit / NST) & 1;
                    tc::mbar_wait(&x_full[s], ph);
                    tc::fence_after_sync();
                    if (tc::elect_one()) {
                        const uint64_t dbs = db0 + (uint64_t)((s * STAGE) >> 4);
#pragma unroll
                        for (int ks = 0; ks < NC / 16; ++ks)     // K = the chunk's (batch, channel) columns, 16 per MMA
                            tc::mma_bf16_ss(tmem + st * KST, dac + (uint64_t)((ks * 2 * A2_K8) >> 4),
                                            dbs + (uint64_t)(((ks / 4) * XBLK + (ks % 4) * 32) >> 4), idS, c > 0 || ks > 0);
                        tc::mma_commit(&x_empty[s]);
                    }
                    __syncwarp();
                }
                if (tc::elect_one()) tc::mma_commit(&a2_empty[buf]);
                __syncwarp();
            }
            if (tc::elect_one()) tc::mma_commit(dp_full);
            __syncwarp();
        }
    } else if (warp < 4 || warp >= 10) {
        // ===== regroup (per chunk) + extraction (per tile), two groups of 4 warps ================================
        const int grp = warp >= 10 ? 1 : 0;
        const int m = (warp & 3) * 32 + lane;                    // TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float* my_stg = stg + grp * 128 * STG_LD;
        const int gtid = m;                                      // thread index inside the group
        uint32_t cc = 0, tt = 0;
        for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
            const int tile = wi % a.ntiles;
            for (int c = 0; c < nch; ++c, ++cc) {
                const uint32_t buf = cc & 1;
                tc::mbar_wait(dt_full, cc & 1);
                tc::mbar_wait(&a2_empty[buf], ((cc >> 1) & 1) ^ 1);   // the sddmm of chunk cc-2 has read this buffer
                tc::fence_after_sync();
                // dense accumulator row = (point in tile, batch lane), columns (h, channel in chunk) -> A2[(point, h)][(b, channel)]
                for (int mb = (MBC > 1 ? grp : 0); mb < MBC; mb += (MBC > 1 ? 2 : 1)) {
                    if (MBC == 1 && grp == 1) break;
                    const int row = mb * 128 + m;
                    const int ol = row / Bp, b = row % Bp;
                    unsigned char* dst = a2_s + buf * A2BUF + (size_t)ol * A2SB + (size_t)((b * CC) >> 3) * A2_K8 + ((b * CC) & 7) * 2;
#pragma unroll
                    for (int g = 0; g < ND / 16; ++g) {
                        uint32_t r[16];
                        tc::tmem_ld16(lane_base + tmem + DT_COL + mb * ND + g * 16, r);
                        tc::wait_ld();
                        if (CC == 4) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)          // h = 4g + q, its 4 channels
                                *reinterpret_cast<uint2*>(dst + (4 * g + q) * 16) =
                                    make_uint2(pack_bf16(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])),
                                               pack_bf16(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                        } else if (CC == 8) {
#pragma unroll
                            for (int q = 0; q < 2; ++q)          // h = 2g + q, its 8 channels
                                *reinterpret_cast<uint4*>(dst + (2 * g + q) * 16) =
                                    make_uint4(pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                                               pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                                               pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                                               pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
                        } else {
#pragma unroll
                            for (int q = 0; q < 2; ++q)          // h = g, channels 8q .. 8q+7 (k-groups 2b, 2b+1)
                                *reinterpret_cast<uint4*>(dst + g * 16 + q * A2_K8) =
                                    make_uint4(pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                                               pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                                               pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                                               pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
                        }
                    }
                }
                tc::fence_before_sync();
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(dt_empty);
                    tc::mbar_arrive(&a2_full[buf]);
                }
            }
            // extraction: 32 union slots (columns) at a time through a staging tile; group g takes every other block
            const int p0 = a.tile_pptr[tile], np = a.tile_pptr[tile + 1] - p0;
            const int ku = a.tile_uptr[tile + 1] - a.tile_uptr[tile];
            tc::mbar_wait(dp_full, tt & 1);
            tc::fence_after_sync();
            for (int cb = grp; cb < ku / 32; cb += 2) {
                uint32_t r0[16], r1[16];
                tc::tmem_ld16(lane_base + tmem + cb * 32, r0);
                tc::tmem_ld16(lane_base + tmem + cb * 32 + 16, r1);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) my_stg[m * STG_LD + j] = __uint_as_float(r0[j]), my_stg[m * STG_LD + 16 + j] = __uint_as_float(r1[j]);
                tc::named_bar_sync(3 + grp, 128);
                for (int i = gtid; i < np; i += 128) {
                    const int2 e = __ldg(a.tile_ent + p0 + i);
                    const int slot = e.y & 0xffff, ol = e.y >> 16;
                    if ((slot >> 5) == cb) {
                        const float* sr = my_stg + (ol * 8) * STG_LD + (slot & 31);
                        float* out = a.dphi + (size_t)e.x * a.HPphi;
                        if (a.nblk == 1) {
                            for (int h = 0; h < a.HPphi; ++h) out[h] = h < a.H ? sr[h * STG_LD] : 0.f;
                        } else {
                            for (int h = 0; h < a.H; ++h) atomicAdd(out + h, sr[h * STG_LD]);
                        }
                    }
                }
                tc::named_bar_sync(3 + grp, 128);
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(dp_empty);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int CC, int LB, bool TMA>
int launch_dphi_v(const DphiMmaArgs& a, cudaStream_t st)
{
    const int CgP = (a.Cg + 15) & ~15;
    const size_t smem = dphi_smem_bytes(CC, 1 << LB, a.nch, CgP, a.max_ku);
    if (smem > 227 * 1024) return 0;
    CUtensorMap xmap;
    memset(&xmap, 0, sizeof(xmap));
    if (TMA && !tma::make_map_2d(&xmap, a.X, 2, (uint64_t)a.nblk * a.Nsrc, (uint64_t)a.nch * (CC << LB), 64, 1)) return 0;
    QC_CUDA((qc_set_max_smem<k_dphi_mma<CC, LB, TMA>>((int)smem)));
    int grid = min(qc_num_sms(), a.ntiles * a.nblk);
    if (grid < 1) grid = 1;
    k_dphi_mma<CC, LB, TMA><<<grid, NTHREADS, smem, st>>>(a, xmap);
    QC_CHECK_LAUNCH();
    return 1;
}

// QC_MMA_VAR bit 3 selects the TMA producers (same switch as contract_mma.cu); the cp.async variant is the fallback
template <int CC, int LB>
int launch_dphi(const DphiMmaArgs& a, cudaStream_t st)
{
    if (qc_mma_get_variant() & 8) {
        const int rc = launch_dphi_v<CC, LB, true>(a, st);
        if (rc != 0) return rc;
    }
    return launch_dphi_v<CC, LB, false>(a, st);
}

}  // namespace

extern "C" {

int qc_mma_chunk_channels(int C, int B);

int qc_dphi_mma_smem_bytes(int Cs, int Cg, int B, int max_ku)
{
    const int CC = qc_mma_chunk_channels(Cs, B);
    if (CC == 0) return 0;
    return (int)dphi_smem_bytes(CC, qc_bp(B), qc_ceil_div(Cs, CC), (Cg + 15) & ~15, max_ku);
}

int qc_dphi_mma(const void* Xh, int Nsrc, int Cs, const void* Gh, int Cg, const void* tile_ent, const int* tile_pptr,
                const int* tile_uptr, const int* union_idx, int ntiles, int max_ku, const float* w_last, int H, int HPphi,
                const int* order, int Ndst, int B, float* dphi, void* stream)
{
    const int Bp = qc_bp(B);
    const int CC = qc_mma_chunk_channels(Cs, B), CCg = qc_mma_chunk_channels(Cg, B);
    if (CC == 0 || CCg == 0 || Bp * CC < 64 || H > 8 || H < 1 || (HPphi != 4 && HPphi != 8) || HPphi < H || Cg > 64 ||
        max_ku < KST || (max_ku % KST) || max_ku > DT_COL)
        return QC_EUNSUP;
    if ((size_t)Nsrc * qc_ceil_div(Cs, CC) * Bp * CC * 2 >= 0xffffffffull) return QC_EUNSUP;   // 32-bit row offsets
    DphiMmaArgs a;
    a.X = reinterpret_cast<const __nv_bfloat16*>(Xh);
    a.G = reinterpret_cast<const __nv_bfloat16*>(Gh);
    a.dphi = dphi;
    a.order = order;
    a.tile_ent = reinterpret_cast<const int2*>(tile_ent);
    a.tile_pptr = tile_pptr;
    a.tile_uptr = tile_uptr;
    a.union_idx = union_idx;
    a.w_last = w_last;
    a.Nsrc = Nsrc, a.Ndst = Ndst, a.Cs = Cs, a.Cg = Cg, a.H = H, a.HPphi = HPphi, a.nch = qc_ceil_div(Cs, CC);
    a.nchg = qc_ceil_div(Cg, CCg), a.CCg = CCg, a.ntiles = ntiles, a.nblk = qc_nblk(B), a.max_ku = max_ku;
    const cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (Bp == 32) rc = launch_dphi<4, 5>(a, st);
    else if (Bp == 16) rc = CC == 8 ? launch_dphi<8, 4>(a, st) : launch_dphi<4, 4>(a, st);
    else if (Bp == 8) rc = CC == 16 ? launch_dphi<16, 3>(a, st) : (CC == 8 ? launch_dphi<8, 3>(a, st) : 0);
    if (rc == 0) return QC_EUNSUP;
    return rc == 1 ? 0 : rc;
}

}  // extern "C"
