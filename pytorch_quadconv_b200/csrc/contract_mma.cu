// contract_mma.cu -- reduced-precision tensor-core path of the fused gather-contract kernel: BOTH stages of the
// factorised layer on tcgen05 (bf16 operands, fp32 accumulation in TMEM).
//
// Reference: QuadConv._integrate + the last Linear of the filter MLP (torch_quadconv/quadconv.py:213-227, :143),
// in the factorised form of DESIGN.md section 3:
//     T[o; b,i,h] = sum_{p in N(o)} phi[p][h] * f[b,i,in(p)]          stage B  (gather stage)
//     Y[o; b,j]   = sum_{i,h} T[o; b,i,h] * W_last[i*Co+j][h]          stage C  (dense stage)
// The fp32 kernels (contract_tc.cu) run stage B on the FP32 pipe, one (point, batch) row per thread, and
// re-gather every neighbour row for every destination point (19.7 GB of L2->SM traffic at C3).  Here a CTA
// works on a TILE of 16 spatially adjacent destination points (cell-sorted order) and on the UNION U of their
// neighbours (184 rows on average at C3 instead of 16 x 48 = 770: adjacent points share 3/4 of their support):
//
//   stage B  D[(o,h)][(b,i)] = sum_{u in U} Phi[(o,h)][u] * X[u][(b,i)]       one dense GEMM per tile
//            M = 16 points x 8 filter components = 128, N = a chunk of (batch, channel) columns, K = |U|;
//            Phi is the tile's zero-filled [128 x |U|] block of phi (a block of the block-sparse matrix whose
//            pattern is eval_indices), X the gathered bf16 feature rows: both MN-major operands in shared memory;
//   regroup  the fp32 accumulator [(o,h)][(b,i)] is read from TMEM, rounded to bf16 and stored as the K-major
//            A operand [(o,b)][(h,i)] of stage C (h and b trade places between the two GEMMs);
//   stage C  Y[(o,b)][j] += T[(o,b)][(h,i)] * W2[(h,i)][j], accumulated over the channel chunks in TMEM.
//
// Warp roles (288 threads, 1 CTA per SM, persistent over tiles):
//   warps 0-3  regroup + output: TMEM -> registers -> shared / global (one TMEM lane per thread)
//   warps 4-7  producers: build Phi for the tile, gather the union rows chunk by chunk into a ring of K-step
//              stages with cp.async (16 bytes per lane straight into the operand layout)
//   warp  8    one thread issues every tcgen05.mma and the tcgen05.commit that releases stages / buffers
// All hand-offs are mbarriers; the MMAs of channel chunk c+1 run while chunk c is regrouped.
#include <cuda_bf16.h>
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int TP = 16;                 // destination points per tile
constexpr int KST = 32;                // union rows per ring stage (two K steps of 16)
constexpr int NST = 8;                 // ring stages
constexpr int LAG = 6;                 // a stage is published once the copies of LAG later stages are issued
constexpr int DCOLS = 128;             // TMEM columns per stage-B accumulator buffer
constexpr int PF_CH = 0;               // L2 prefetch distance of the gather in channel chunks (0 = off: measured slower,
                                       // the prefetches compete with the copies for the same LSU request slots)
constexpr int ENT_R = 8;               // pair entries per producer thread held in registers (1024 pairs per tile)
constexpr uint32_t PHI_K8 = TP * 128;  // bytes between groups of 8 union rows in the Phi operand (16 points x 128 B)
constexpr uint32_t AC_K8 = 144;        // bytes between groups of 8 k in the stage-C A operand (16 B pad: bank spread)
constexpr int NTHREADS = 448;         // 4 regroup + 4 producer + 2 issuer + 4 regroup warps

struct MmaArgs {
    const __nv_bfloat16* X;     // packed source features [nblk][Nsrc][nch][Bp][CC]
    float* Y;                   // packed destination rows, fp32 [nblk][Ndst][CP4/4][Bp][4] (common.cuh)
    const int* order;           // processing order of the destination points (tile t = entries 16t .. 16t+15); may be null
    const int2* tile_ent;       // pair entries grouped by tile: {phi row, (point in tile << 16) | union slot}
    const int* tile_pptr;       // [ntiles+1] offsets into tile_ent
    const int* tile_uptr;       // [ntiles+1] offsets into union_idx (multiples of 32, at least 32 per tile)
    const int* union_idx;       // union lists, -1 = padding
    const float* phi;           // [P][HPphi] fp32
    const float* w_last;        // [Ci*Co][H] fp32 (filter.{2L}.weight)
    // DW mode (d W_last = sum over rows of T x grad rows instead of stage C):
    const __nv_bfloat16* G;     // packed rows of the destination-side tensor [nblk][Ndst][nchg][Bp][CCg]
    float* dw_part;             // per-CTA partial results [grid][groups][MDW][CgP]
    int Cg, nchg, CCg;
    int Nsrc, Ndst, Cs, Cd, H, HPphi, nch, transposed, ntiles, nblk, max_ku;   // HPphi = floats per phi row (4 or 8)
    long long* dbg;             // QC_MMA_PROFILE builds: wait-cycle counters of CTA 0 (null otherwise)
};

#ifdef QC_MMA_PROFILE
#define PROF_DECL(n) long long n = 0
#define PROF_T0() const long long t0__ = clock64()
#define PROF_ADD(n) n += clock64() - t0__
#define PROF_WAIT(n, ...) do { const long long t0__ = clock64(); __VA_ARGS__; n += clock64() - t0__; } while (0)
#else
#define PROF_DECL(n)
#define PROF_WAIT(n, ...) do { __VA_ARGS__; } while (0)
#endif

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// byte step between groups of 8 rows of the stage-C A operand (CC k-groups of one channel chunk), padded so that
// the regroup stores of a warp (4 points x 8 filter components at one batch lane) spread over all banks
__host__ __device__ constexpr uint32_t ac_row8(int CC, int Bp)
{
    return (uint32_t)CC * AC_K8 + (Bp >= 32 ? 16u : (Bp == 16 ? 32u : 64u));
}

__host__ __device__ constexpr size_t mma_smem_bytes(int CC, int Bp, int nch, int CdP, int max_ku, bool dw = false, int CgP = 0)
{
    const int cpp = 1;
    const size_t ring = (Bp * CC) % 64 == 0 ? (size_t)NST * KST * Bp * CC * 2 : (size_t)NST * (KST / 8) * (Bp * CC / 8) * 144;
    return (size_t)max_ku * 256 + ring + (size_t)(TP * Bp / 8) * ac_row8(CC * cpp, Bp) +
           (dw ? (size_t)TP * Bp * CgP * 2 : (size_t)nch * CdP * 16 * CC) + (size_t)max_ku * 8 + 64 * 8;
}

// VAR (gather experiment): bit 0 = cp.async.cg (bypass L1), bit 1 = n-chunk stride 144 B so that a quarter warp copies
// ONE row's 128 contiguous bytes (8 chunks) without bank conflicts instead of 16 bytes of 8 rows
// DW = true: the dW_last pass.  Stage B and the regroup are unchanged; instead of stage C the regrouped tile T[(o,b)][(h,i)]
// is the MN-major A operand of   D_dw[(h,i)][j] += sum_{(o,b)} T[(o,b)][(h,i)] * G[(o,b)][j]   (K = the tile's rows), G = the
// tile's own rows of the destination-side tensor (grad_out in the forward view).  M = 64 or 128 k-values = CPP channel
// chunks side by side in the regroup tile; the accumulators stay in TMEM for the whole kernel (one per chunk group) and
// every CTA writes its partial sums once at the end.
template <int CC, int LB, int VAR, bool DW>
__global__ void __launch_bounds__(NTHREADS, 1) k_contract_mma(const MmaArgs a, const __grid_constant__ CUtensorMap xmap)
{
    constexpr int Bp = 1 << LB, NC = Bp * CC, QN = NC / 8;
    // (one channel chunk per accumulator: pairing two 32-value chunks into one M = 64 operand would need both halves of
    //  the regroup tile resident, 37 KB that the gather ring needs more)
    constexpr int MDW = 8 * CC;                           // valid rows of the dW accumulator (the MMA is issued with M = 128)
    constexpr int CPP = 1;                                // channel chunks side by side in the regroup tile
    constexpr bool CG = (VAR & 1) != 0, ROWMAJ = (VAR & 2) != 0;
    // bit 2: the gathered operand in the 128-byte-swizzle MN-major layout (rows of 64 columns = 128 contiguous bytes,
    // 16-byte chunks XOR-ed with the row index inside 1 KB atoms): a 128-byte line from L2 lands as ONE shared-memory
    // wavefront (the no-swizzle layouts scatter it into eight 16-byte pieces, 6x the LSU data-pipe traffic, ncu)
    constexpr bool SW = (VAR & 4) != 0 && (NC % 64 == 0);
    // bit 3: the union rows arrive by TMA (tma.cuh): one gather4 instruction per four rows and 64-column block, issued
    // by the lanes of ONE producer warp, completion counted in bytes on the stage's mbarrier; padding slots name a row
    // past the end of the tensor and are zero-filled by the copy unit.  xmap: X as [nblk * Nsrc][nch * NC] bf16.
    constexpr bool TMA = SW && (VAR & 8) != 0;
    constexpr uint32_t XSB = ROWMAJ ? 144 : 128;          // bytes between 8-column chunks of the gathered operand
    constexpr int rows = TP * Bp, MBC = rows / 128;       // rows of the stage-C GEMM per tile, its 128-row blocks
    constexpr uint32_t ACR = ac_row8(CC * CPP, Bp);
    constexpr uint32_t XK8 = (uint32_t)QN * XSB;          // bytes between k-groups (8 union rows) of a ring stage
    constexpr uint32_t XBLK = KST * 128;                  // SW: bytes between 64-column blocks of a stage
    constexpr uint32_t STAGE = SW ? (NC / 64) * XBLK : (KST / 8) * XK8;
    const int nch = a.nch;
    const int CdP = DW ? 0 : max(16, (a.Cd + 15) & ~15);  // N of stage C
    const uint32_t W2CH = (uint32_t)CdP * 16 * CC;        // bytes of one chunk of the stage-C B operand
    const int CgP = DW ? ((a.Cg + 15) & ~15) : 0;         // N of the dW GEMM
    constexpr uint32_t GMN8 = (rows / 8) * 128;           // DW: bytes between 8-channel groups of the G tile (MN-major B operand)

    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* phi_s = smem;                                        // [max_ku/8][16 points][8 k][8 h] bf16
    unsigned char* x_s = phi_s + (size_t)a.max_ku * 256;                 // [NST][4][QN][8 k][8 n] bf16
    unsigned char* ac_s = x_s + NST * STAGE;                             // [rows/8][ACR]
    unsigned char* w2_s = ac_s + (size_t)(rows / 8) * ACR;               // [nch][CdP/8][CC][8 n][8 k] bf16  (DW: the G tile)
    unsigned char* g_s = w2_s;                                           // DW: [CgP/8][rows/8][8 rows][8 j] bf16
    int* uidx_s = reinterpret_cast<int*>(w2_s + (DW ? (size_t)rows * CgP * 2 : (size_t)nch * W2CH));     // [2][max_ku]
    uint64_t* bars = reinterpret_cast<uint64_t*>(uidx_s + 2 * a.max_ku);
    uint64_t* x_full = bars;                 // [NST]  producers (128 copy-completion arrivals) -> MMA
    uint64_t* x_empty = x_full + NST;        // [NST]  MMA commit -> producers
    uint64_t* d_full = x_empty + NST;        // [2]    MMA commit -> regroup
    uint64_t* d_empty = d_full + 2;          // [2]    regroup (4) -> MMA
    uint64_t* ac_full = d_empty + 2;         // [2] regroup (8 warps) -> MMA; DW: one per chunk slot of the regroup tile, so that
                                             // a barrier is never more than one phase ahead of its waiter
    uint64_t* ac_empty = ac_full + 2;        // MMA commit -> regroup
    uint64_t* phi_full = ac_empty + 1;       // producers (4) -> MMA
    uint64_t* phi_empty = phi_full + 1;      // MMA commit -> producers
    uint64_t* y_full = phi_empty + 1;        // MMA commit -> output
    uint64_t* y_empty = y_full + 1;          // output (4) -> MMA
    uint64_t* g_full = y_empty + 1;          // DW: producers (128 copy completions) -> dW issuer
    uint64_t* g_empty = g_full + 1;          // DW: dW commit -> producers
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(g_empty + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // ---- one-time setup ---------------------------------------------------------------------------------
    if (warp == 0) tc::tmem_alloc(tmem_base_s, 512);
    if (tid == 32) {
        for (int i = 0; i < NST; ++i) tc::mbar_init(&x_full[i], TMA ? 1 : 128), tc::mbar_init(&x_empty[i], 1);
        for (int i = 0; i < 2; ++i) tc::mbar_init(&d_full[i], 1), tc::mbar_init(&d_empty[i], 8);
        tc::mbar_init(&ac_full[0], 8), tc::mbar_init(&ac_full[1], 8), tc::mbar_init(ac_empty, 1);
        tc::mbar_init(phi_full, 4), tc::mbar_init(phi_empty, 1);
        tc::mbar_init(y_full, 1), tc::mbar_init(y_empty, 8);
        tc::mbar_init(g_full, 128), tc::mbar_init(g_empty, 1);
        tc::mbar_fence_init();
    }
    // stage-C weights, bf16, K-major [j][(h, channel in chunk)] per channel chunk
    if (!DW) {
        constexpr int KC = 8 * CC;
        const int total = nch * CdP * KC;
        for (int idx = tid; idx < total; idx += NTHREADS) {
            const int c = idx / (CdP * KC), r = idx % (CdP * KC);
            const int n = r / KC, k = r % KC;
            const int h = k / CC, cs = c * CC + k % CC;
            float v = 0.f;
            if (cs < a.Cs && n < a.Cd && h < a.H)
                v = __ldg(a.w_last + (a.transposed ? ((size_t)n * a.Cs + cs) : ((size_t)cs * a.Cd + n)) * a.H + h);
            const uint32_t off = (uint32_t)c * W2CH + tc::kmajor16_off(n, k, CC * 128, 128);
            *reinterpret_cast<__nv_bfloat16*>(w2_s + off) = __float2bfloat16_rn(v);
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_base_s;
    const uint32_t tmem_y = tmem + 2 * DCOLS;
    const int nwork = a.ntiles * a.nblk;

    if (warp >= 4 && warp < 8) {
        // ===== producers ====================================================================================
        const int pt = tid - 128;
        const uint32_t x_sa = tc::smem_u32(x_s);
        uint32_t xit = 0, tt = 0;
        PROF_DECL(w_xempty); PROF_DECL(w_phiempty); PROF_DECL(w_cpwait); PROF_DECL(w_phibuild); PROF_DECL(w_total);
#ifdef QC_MMA_PROFILE
        const long long tstart = clock64();
#endif
        // pair entries of a tile, ENT_R per thread in registers, fetched one tile ahead
        int2 ent[ENT_R];
        auto fetch_entries = [&](int wi2) {
            if (wi2 < nwork) {
                const int tile2 = wi2 % a.ntiles;
                const int p0 = a.tile_pptr[tile2], np = a.tile_pptr[tile2 + 1] - p0;
#pragma unroll
                for (int r = 0; r < ENT_R; ++r) {
                    const int i = pt + r * 128;
                    ent[r] = i < np ? __ldg(a.tile_ent + p0 + i) : make_int2(-1, 0);
                }
            }
        };
        auto phi_row = [&](int pid) -> uint4 {
            const float4 p0 = __ldg(reinterpret_cast<const float4*>(a.phi + (size_t)pid * a.HPphi));
            float4 p1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.HPphi == 8) p1 = __ldg(reinterpret_cast<const float4*>(a.phi + (size_t)pid * 8 + 4));
            return make_uint4(pack_bf16(p0.x, p0.y), pack_bf16(p0.z, p0.w), pack_bf16(p1.x, p1.y), pack_bf16(p1.z, p1.w));
        };
        auto phi_put = [&](int meta, const uint4& v) {
            const int ol = meta >> 16, slot = meta & 0xffff;
            *reinterpret_cast<uint4*>(phi_s + (size_t)(slot >> 3) * PHI_K8 + ol * 128 + (slot & 7) * 16) = v;
        };
        fetch_entries(blockIdx.x);
        for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
            const int tile = wi % a.ntiles, blk = wi / a.ntiles;
            const int u0 = a.tile_uptr[tile], ku = a.tile_uptr[tile + 1] - u0, nst = ku / KST;
            const int p0 = a.tile_pptr[tile], np = a.tile_pptr[tile + 1] - p0;
            // phi rows of this tile (the loads fly while the MMAs finish the previous tile)
            uint4 pv[ENT_R];
#pragma unroll
            for (int r = 0; r < ENT_R; ++r)
                if (ent[r].x >= 0) pv[r] = phi_row(ent[r].x);
            // union lists: this tile's (loaded one tile ago) and the next tile's (for the L2 prefetch below); the buffer
            // being refilled was last read two tiles ago, every producer is past it after the barrier
            int* uix = uidx_s + (tt & 1) * a.max_ku;
            int* uix_next = uidx_s + ((tt + 1) & 1) * a.max_ku;
            const int wi2 = wi + gridDim.x;
            const bool has_next = wi2 < nwork;
            int nst_next = 0, blk_next = 0;
            tc::named_bar_sync(2, 128);
            const uint32_t rowbytes = (uint32_t)(nch * NC * 2);
            // cp.async: byte offset of the row inside this batch block (< 4 GB: checked by the host), -1 = padding;
            // TMA: row index of the 2-D view (batch block included), the row count = out of bounds = zero fill
            const int oob_row = a.nblk * a.Nsrc;
            auto row_off = [&](int u, int bk) {
                if (TMA) return u >= 0 ? bk * a.Nsrc + u : oob_row;
                return u >= 0 ? (int)((uint32_t)u * rowbytes) : -1;
            };
            if (tt == 0)
                for (int i = pt; i < ku; i += 128) uix[i] = row_off(__ldg(a.union_idx + u0 + i), blk);
            if (has_next) {
                const int tile2 = wi2 % a.ntiles;
                const int v0 = a.tile_uptr[tile2], kv = a.tile_uptr[tile2 + 1] - v0;
                nst_next = kv / KST, blk_next = wi2 / a.ntiles;
                for (int i = pt; i < kv; i += 128) uix_next[i] = row_off(__ldg(a.union_idx + v0 + i), blk_next);
            }
            tc::named_bar_sync(2, 128);

            // gather: per channel chunk, per stage: 32 union rows x NC columns, 16 bytes per lane, straight into the
            // MN-major operand layout.
            const size_t srow = (size_t)nch * NC;               // elements per source point
            const __nv_bfloat16* Xb = a.X + (size_t)blk * a.Nsrc * srow;
            // lane -> (row in k-group, column chunk): default 8 rows x 4 chunks per warp; ROWMAJ: 8 chunks (128 B) of 4 rows
            constexpr int QL = ROWMAJ ? (QN < 8 ? QN : 8) : 1;            // consecutive chunks handled by consecutive lanes
            const int k8 = ROWMAJ ? (pt / QL) & 7 : pt & 7;
            const int qq = ROWMAJ ? (pt % QL) + QL * (pt / (QL * 8)) : (pt >> 3);
            const int q = qq % QN, kg0 = qq / QN;
            constexpr int KGSTEP = 16 / QN;
            const uint32_t dst0 = x_sa + q * XSB + k8 * 16;
            const __nv_bfloat16* Xq = Xb + q * 8;
            // The union list holds BYTE offsets of the rows (0xffffffff = padding): the copy addresses of a stage are
            // formed from 32-bit loads issued together BEFORE the wait on the ring slot, then the copies go out back to
            // back (the index load -> 64-bit address arithmetic -> copy chain of one piece after the other made the
            // producers latency bound at ~700 cycles per stage: the slowest role of the kernel).
            const char* Xbytes = reinterpret_cast<const char*>(Xb);
            auto issue_stage = [&](int c, int st) {
                const uint32_t s = xit % NST, ph = (xit / NST) & 1;
                const uint32_t* uo = reinterpret_cast<const uint32_t*>(uix) + st * KST;
                if constexpr (TMA) {
                    if (pt < 32) {
                        constexpr int NI = (KST / 4) * (NC / 64);          // gather4 instructions per stage
                        const int rg = pt % (KST / 4), b64 = pt / (KST / 4);
                        int4 r = make_int4(0, 0, 0, 0);
                        if (pt < NI) r = *reinterpret_cast<const int4*>(uo + rg * 4);
                        PROF_WAIT(w_xempty, tc::mbar_wait(&x_empty[s], ph ^ 1));
                        if (pt == 0) tma::mbar_arrive_expect_tx(&x_full[s], STAGE);
                        __syncwarp();
                        if (pt < NI)
                            tma::gather4(x_sa + s * STAGE + b64 * XBLK + rg * 512, &xmap, &x_full[s], c * NC + b64 * 64, r.x, r.y,
                                         r.z, r.w);
                    }
                    ++xit;
                    return;
                }
                if constexpr (SW) {
                    // 8 lanes = the 8 chunks of one 128-byte row piece; line = (block of 64 columns, union row)
                    constexpr int NJ = KST * (NC / 64) / 16;
                    const int cidx = pt & 7, l0 = pt >> 3;
                    uint32_t off[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) off[j] = uo[(l0 + 16 * j) % KST];
                    const char* base = Xbytes + (size_t)c * NC * 2 + cidx * 16;
                    const uint32_t dbase = x_sa + s * STAGE + ((l0 % KST) * 128) + ((cidx ^ (l0 & 7)) << 4);
                    PROF_WAIT(w_xempty, tc::mbar_wait(&x_empty[s], ph ^ 1));
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        // line = l0 + 16 j: row (l0 + 16 j) % 32 (same row & 7 for every j), block (l0 + 16 j) / 32
                        const int line = l0 + 16 * j;
                        const bool valid = off[j] != 0xffffffffu;
                        const char* src = base + (line / KST) * 128 + (valid ? off[j] : 0u);
                        const uint32_t dst = dbase + (line / KST) * XBLK + ((line % KST) - (l0 % KST)) * 128;
                        if (CG) tc::cp_async16_cg_zfill(dst, src, valid);
                        else tc::cp_async16_zfill(dst, src, valid);
                    }
                } else {
                    constexpr int NJ = (KST / 8 + KGSTEP - 1) / KGSTEP;
                    uint32_t off[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int kg = kg0 + j * KGSTEP;
                        off[j] = kg < KST / 8 ? uo[kg * 8 + k8] : 0xffffffffu;
                    }
                    const char* base = Xbytes + ((size_t)c * NC + q * 8) * 2;
                    PROF_WAIT(w_xempty, tc::mbar_wait(&x_empty[s], ph ^ 1));
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int kg = kg0 + j * KGSTEP;
                        if (kg < KST / 8) {
                            const bool valid = off[j] != 0xffffffffu;
                            if (CG) tc::cp_async16_cg_zfill(dst0 + s * STAGE + kg * XK8, base + (valid ? off[j] : 0u), valid);
                            else tc::cp_async16_zfill(dst0 + s * STAGE + kg * XK8, base + (valid ? off[j] : 0u), valid);
                        }
                    }
                }
                // the stage's mbarrier counts one arrival per producer thread, delivered by the copy unit when this
                // thread's copies have landed: no wait on the issuing side, nothing held back
                tc::cp_async_mbar_arrive_noinc(&x_full[s]);
                ++xit;
            };
            // the first stages of this tile go out BEFORE its Phi is built: they only need ring slots, which free up
            // as the MMAs finish the previous tile, so the gather stream never drains at a tile boundary
            const int total = nch * nst;
            const int pre = min(total, NST - 1);
            int c = 0, st = 0;
            for (int j = 0; j < pre; ++j) {
                issue_stage(c, st);
                if (++st == nst) st = 0, ++c;
            }
            PROF_WAIT(w_phiempty, tc::mbar_wait(phi_empty, (tt & 1) ^ 1));   // the MMAs are done with the previous Phi
#ifdef QC_MMA_PROFILE
            const long long tb0 = clock64();
#endif
            for (int i = pt; i < ku * 16; i += 128) reinterpret_cast<uint4*>(phi_s)[i] = make_uint4(0, 0, 0, 0);
            tc::named_bar_sync(2, 128);
#pragma unroll
            for (int r = 0; r < ENT_R; ++r)
                if (ent[r].x >= 0) phi_put(ent[r].y, pv[r]);
            for (int i = pt + ENT_R * 128; i < np; i += 128) {   // tiles with more than 1024 pairs
                const int2 e = __ldg(a.tile_ent + p0 + i);
                phi_put(e.y, phi_row(e.x));
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(phi_full);
#ifdef QC_MMA_PROFILE
            w_phibuild += clock64() - tb0;
#endif
            fetch_entries(wi + gridDim.x);
            if (DW) {
                // the tile's own rows of the destination-side tensor: MN-major B operand [j][(point, batch lane)],
                // pieces of min(16, 2*CCg) bytes (one 16-byte line = 8 channels of one row)
                tc::mbar_wait(g_empty, (tt & 1) ^ 1);
                const size_t grow = (size_t)a.nchg * Bp * a.CCg;
                const __nv_bfloat16* Gb = a.G + (size_t)blk * a.Ndst * grow;
                const uint32_t g_sa = tc::smem_u32(g_s);
                const int pb = a.CCg >= 8 ? 16 : 8;
                const int ppr = CgP * 2 / pb;
                for (int idx = pt; idx < rows * ppr; idx += 128) {
                    const int row = idx % rows, pc = idx / rows;
                    const int ol2 = row / Bp, b = row % Bp;
                    const int j0 = pc * (pb / 2);
                    const int t = tile * TP + ol2;
                    const bool valid = t < a.Ndst && j0 < a.nchg * a.CCg;
                    const int dstp = valid ? (a.order ? a.order[t] : t) : 0;
                    const __nv_bfloat16* src = Gb + (size_t)dstp * grow + ((size_t)(j0 / a.CCg) * Bp + b) * a.CCg + j0 % a.CCg;
                    const uint32_t dst = g_sa + tc::mnmajor16_off(j0, row, GMN8, 128);
                    if (pb == 16) tc::cp_async16_zfill(dst, src, valid);
                    else tc::cp_async8_zfill(dst, src, valid);
                }
                tc::cp_async_mbar_arrive_noinc(g_full);
            }
            for (int j = pre; j < total; ++j) {
                issue_stage(c, st);
                if (++st == nst) st = 0, ++c;
            }
        }
#ifdef QC_MMA_PROFILE
        if (a.dbg && blockIdx.x == 0 && pt == 0) {
            a.dbg[0] = clock64() - tstart, a.dbg[1] = w_xempty, a.dbg[2] = w_phiempty, a.dbg[3] = w_cpwait, a.dbg[4] = w_phibuild;
        }
#endif
    } else if (warp == 8) {
        // ===== stage-B issuer (one thread): Phi x gathered rows -> D ==========================================
        // The issuing thread is on the critical path (an MMA of 128 columns takes 65 cycles): descriptors are
        // advanced by integer adds, and the readiness of the next stage is tested while the current MMAs issue.
        // (the whole warp runs the loop with warp-uniform values, one elected lane issues: the descriptors then live in
        //  uniform registers and a tcgen05.mma costs a handful of instructions instead of a divergence-safe loop)
        {
            const uint32_t idB = tc::make_idesc_bf16(128, NC, true, true);
            const uint64_t da0 = tc::make_smem_desc(tc::smem_u32(phi_s), PHI_K8, 128);
            const uint64_t db0 = SW ? (tc::make_smem_desc(tc::smem_u32(x_s), XBLK, 1024) | ((uint64_t)2 << 61))
                                    : tc::make_smem_desc(tc::smem_u32(x_s), XK8, XSB);
            constexpr uint32_t XKSTEP = SW ? 2048 : 2 * XK8;       // bytes per K step of 16 union rows
            uint32_t xit = 0, cc = 0, tt = 0;
            PROF_DECL(m_xfull); PROF_DECL(m_dempty); PROF_DECL(m_phifull);
#ifdef QC_MMA_PROFILE
            const long long tstart = clock64();
#endif
            for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
                const int tile = wi % a.ntiles;
                const int nst = (__ldg(a.tile_uptr + tile + 1) - __ldg(a.tile_uptr + tile)) / KST;
                PROF_WAIT(m_phifull, tc::mbar_wait(phi_full, tt & 1));     // (the producers fenced their Phi stores)
                tc::fence_after_sync();
                for (int c = 0; c < nch; ++c, ++cc) {
                    const uint32_t buf = cc & 1, dph = (cc >> 1) & 1;
                    PROF_WAIT(m_dempty, tc::mbar_wait(&d_empty[buf], dph ^ 1));       // chunk cc-2 has been read out of this accumulator
                    tc::fence_after_sync();
                    const uint32_t dcol = tmem + buf * DCOLS;
                    for (int st = 0; st < nst; ++st, ++xit) {
                        const uint32_t s = xit % NST, ph = (xit / NST) & 1;
                        // (cp.async completion delivered through the mbarrier orders the copies before the MMAs that
                        //  follow the wait; no proxy fence: it costs a MEMBAR per stage and drains the copies in flight)
                        PROF_WAIT(m_xfull, tc::mbar_wait(&x_full[s], ph));
                        tc::fence_after_sync();
                        if (tc::elect_one()) {
                            const uint64_t da = da0 + (uint64_t)((st * (KST / 16) * 2 * PHI_K8) >> 4);
                            const uint64_t db = db0 + (uint64_t)((s * STAGE) >> 4);
                            tc::mma_bf16_ss(dcol, da, db, idB, st > 0);
#pragma unroll
                            for (int k2 = 1; k2 < KST / 16; ++k2)
                                tc::mma_bf16_ss(dcol, da + (uint64_t)((k2 * 2 * PHI_K8) >> 4), db + (uint64_t)((k2 * XKSTEP) >> 4), idB, true);
                            tc::mma_commit(&x_empty[s]);
                        }
                        __syncwarp();
                    }
                    if (tc::elect_one()) tc::mma_commit(&d_full[buf]);
                    __syncwarp();
                }
                if (tc::elect_one()) tc::mma_commit(phi_empty);
                __syncwarp();
            }
#ifdef QC_MMA_PROFILE
            if (a.dbg && blockIdx.x == 0 && lane == 0) a.dbg[8] = clock64() - tstart, a.dbg[9] = m_xfull, a.dbg[10] = m_dempty, a.dbg[12] = m_phifull;
#endif
        }
    } else if (warp == 9 && !DW) {
        // ===== stage-C issuer (one elected lane): regrouped T x W2 -> Y ======================================
        {
            const uint32_t idC = tc::make_idesc_bf16(128, CdP, false, false);
            const uint64_t da0 = tc::make_smem_desc(tc::smem_u32(ac_s), AC_K8, ACR);
            const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(w2_s), 128, CC * 128);
            uint32_t cc = 0, tt = 0;
            PROF_DECL(m_acfull); PROF_DECL(m_yempty);
            for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
                PROF_WAIT(m_yempty, tc::mbar_wait(y_empty, (tt & 1) ^ 1));       // the previous tile's output has left TMEM
                for (int c = 0; c < nch; ++c, ++cc) {
                    PROF_WAIT(m_acfull, tc::mbar_wait(&ac_full[0], cc & 1));
                    tc::fence_after_sync();
                    if (tc::elect_one()) {
                        const uint64_t dbc = db0 + (uint64_t)((c * W2CH) >> 4);
                        for (int mb = 0; mb < MBC; ++mb)
#pragma unroll
                            for (int k2 = 0; k2 < CC / 2; ++k2)
                                tc::mma_bf16_ss(tmem_y + mb * CdP, da0 + (uint64_t)((mb * 16 * ACR + k2 * 2 * AC_K8) >> 4),
                                                dbc + (uint64_t)((k2 * 256) >> 4), idC, c > 0 || k2 > 0);
                        tc::mma_commit(ac_empty);
                    }
                    __syncwarp();
                }
                if (tc::elect_one()) tc::mma_commit(y_full);
                __syncwarp();
            }
#ifdef QC_MMA_PROFILE
            if (a.dbg && blockIdx.x == 0 && lane == 0) a.dbg[11] = m_acfull, a.dbg[13] = m_yempty;
#endif
        }
    } else if (warp == 9 && DW) {
        // ===== dW issuer: D_dw[group][(h,i)][j] += T^T x G over the tile's rows ==================================
        {
            // (issued with M = 128: an M = 64 MMA never retired here; accumulator rows >= MDW read neighbouring
            //  shared-memory bytes -- finite bf16 data -- and are never looked at)
            const uint32_t idW = tc::make_idesc_bf16(128, CgP, true, true);
            // A = the regroup tile read MN-major: 8 k-values (one 16-byte line) contiguous, k-groups AC_K8 apart, K = rows
            const uint64_t da0 = tc::make_smem_desc(tc::smem_u32(ac_s), ACR, AC_K8);
            const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(g_s), 128, GMN8);
            uint32_t cc = 0, tt = 0, slot_uses[2] = {0, 0};
            for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
                tc::mbar_wait(g_full, tt & 1);
                for (int c = 0; c < nch; ++c, ++cc) {
                    const int hf = c % CPP;
                    tc::mbar_wait(&ac_full[hf], slot_uses[hf] & 1);
                    ++slot_uses[hf];
                    tc::fence_after_sync();
                    if ((c % CPP) == CPP - 1 || c == nch - 1) {          // the chunk group is complete
                        if (tc::elect_one()) {
                            const uint32_t dcol = tmem_y + (c / CPP) * CgP;
                            for (int ks = 0; ks < rows / 16; ++ks)
                                tc::mma_bf16_ss(dcol, da0 + (uint64_t)((ks * 2 * ACR) >> 4), db0 + (uint64_t)((ks * 256) >> 4), idW,
                                                tt > 0 || ks > 0);
                            tc::mma_commit(ac_empty);
                        }
                        __syncwarp();
                    }
                }
                if (tc::elect_one()) tc::mma_commit(g_empty);
                __syncwarp();
            }
            if (tc::elect_one()) tc::mma_commit(y_full);                 // every dW MMA of this CTA has retired
            __syncwarp();
        }
    } else if (warp < 4 || warp >= 10) {
        // ===== regroup + output, two groups of 4 warps (thread = TMEM lane m = (point in tile, filter component));
        // group 0 (warps 0-3) takes the first half of a chunk's columns, group 1 (warps 10-13) the second =========
        const int grp = warp >= 10 ? 1 : 0;
        const int m = (warp & 3) * 32 + lane;
        const int ol = m >> 3, h = m & 7;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const size_t yrow = (size_t)qc_cp4(a.Cd) * Bp;
        // element (row = ol*Bp + b, k = h*CC + il) of the K-major stage-C operand
        unsigned char* dst00 = ac_s + (size_t)((h * CC) >> 3) * AC_K8 + ((h * CC) & 7) * 2;
        uint32_t cc = 0, tt = 0, gc = 0;
        PROF_DECL(e_dfull); PROF_DECL(e_acempty); PROF_DECL(e_regroup); PROF_DECL(e_yfull); PROF_DECL(e_out);
#ifdef QC_MMA_PROFILE
        const long long tstart = clock64();
#endif
        for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x, ++tt) {
            const int tile = wi % a.ntiles, blk = wi / a.ntiles;
            for (int c = 0; c < nch; ++c, ++cc) {
                const uint32_t buf = cc & 1, ph = (cc >> 1) & 1;
                PROF_WAIT(e_dfull, tc::mbar_wait(&d_full[buf], ph));
                if (!DW) {
                    PROF_WAIT(e_acempty, tc::mbar_wait(ac_empty, (cc & 1) ^ 1));       // stage C of the previous chunk has read the tile
                } else if ((c % CPP) == 0) {
                    tc::mbar_wait(ac_empty, (gc & 1) ^ 1);                             // ... the dW MMAs of the previous chunk group
                    ++gc;
                }
                tc::fence_after_sync();
                unsigned char* dst0 = dst00 + (DW ? (c % CPP) * CC * AC_K8 : 0);
#ifdef QC_MMA_PROFILE
                const long long tr0 = clock64();
#endif
#pragma unroll 2
                for (int g = grp * (NC / 32); g < (grp + 1) * (NC / 32); ++g) {
                    uint32_t r[16];
                    tc::tmem_ld16(lane_base + tmem + buf * DCOLS + g * 16, r);
                    tc::wait_ld();
                    if (CC == 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int row = ol * Bp + g * 4 + q;
                            const uint2 v = make_uint2(pack_bf16(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])),
                                                       pack_bf16(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                            *reinterpret_cast<uint2*>(dst0 + (size_t)(row >> 3) * ACR + (row & 7) * 16) = v;
                        }
                    } else if (CC == 8) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int row = ol * Bp + g * 2 + q;
                            const uint4 v = make_uint4(pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                                                       pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                                                       pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                                                       pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
                            *reinterpret_cast<uint4*>(dst0 + (size_t)(row >> 3) * ACR + (row & 7) * 16) = v;
                        }
                    } else {
                        const int row = ol * Bp + g;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint4 v = make_uint4(pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                                                       pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                                                       pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                                                       pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
                            *reinterpret_cast<uint4*>(dst0 + (size_t)(row >> 3) * ACR + q * AC_K8 + (row & 7) * 16) = v;
                        }
                    }
                }
                tc::fence_before_sync();
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&d_empty[buf]);
                    tc::mbar_arrive(&ac_full[DW ? c % CPP : 0]);
                }
#ifdef QC_MMA_PROFILE
                e_regroup += clock64() - tr0;
#endif
            }
            if (DW) continue;
            // output rows of this tile: TMEM lane = row % 128, row = (point in tile) * Bp + batch lane
            PROF_WAIT(e_yfull, tc::mbar_wait(y_full, tt & 1));
            tc::fence_after_sync();
#ifdef QC_MMA_PROFILE
            const long long to0 = clock64();
#endif
            for (int mb = (MBC > 1 ? grp : 0); mb < MBC; mb += (MBC > 1 ? 2 : 1)) {
                if (MBC == 1 && grp == 1) break;
                const int row = mb * 128 + m;
                const int o2 = row / Bp, b = row % Bp;
                const int t = tile * TP + o2;
                const int dstp = t < a.Ndst ? (a.order ? a.order[t] : t) : -1;
                float* y = a.Y + ((size_t)blk * a.Ndst + (dstp >= 0 ? dstp : 0)) * yrow + b * 4;
                for (int g = 0; g < CdP / 16; ++g) {
                    uint32_t r[16];
                    tc::tmem_ld16(lane_base + tmem_y + mb * CdP + g * 16, r);
                    tc::wait_ld();
                    if (dstp >= 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (g * 16 + 4 * q < a.Cd)
                                *reinterpret_cast<float4*>(y + (size_t)(g * 4 + q) * Bp * 4) =
                                    make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(y_empty);
#ifdef QC_MMA_PROFILE
            e_out += clock64() - to0;
#endif
        }
        if (DW && grp == 0) {
            // partial d W_last of this CTA: accumulator row (= TMEM lane) m = (chunk in group, h, channel in chunk), m < MDW
            tc::mbar_wait(y_full, 0);
            tc::fence_after_sync();
            const int ngrp = (nch + CPP - 1) / CPP;
            const bool has_row = m < MDW;
            const int mrow = m;
            for (int g = 0; g < ngrp; ++g)
                for (int q = 0; q < CgP / 16; ++q) {
                    uint32_t r[16];
                    tc::tmem_ld16(lane_base + tmem_y + g * CgP + q * 16, r);
                    tc::wait_ld();
                    if (has_row) {
                        float* out = a.dw_part + (((size_t)blockIdx.x * ngrp + g) * MDW + mrow) * CgP + q * 16;
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(out + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                }
        }
#ifdef QC_MMA_PROFILE
        if (a.dbg && blockIdx.x == 0 && tid == 0) {
            a.dbg[16] = clock64() - tstart, a.dbg[17] = e_dfull, a.dbg[18] = e_acempty, a.dbg[19] = e_regroup, a.dbg[20] = e_yfull,
            a.dbg[21] = e_out;
        }
#endif
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// producer variant of the gather (see the VAR bits of k_contract_mma): QC_MMA_VAR read once per process, or set through
// qc_mma_set_variant (tests compare the TMA producers with the cp.async ones in one process)
int g_mma_var = -1;
inline int mma_var()
{
    if (g_mma_var < 0) g_mma_var = getenv("QC_MMA_VAR") ? atoi(getenv("QC_MMA_VAR")) : 4;
    return g_mma_var;
}

inline int mma_grid(const MmaArgs& a)
{
    const int grid = min(qc_num_sms(), a.ntiles * a.nblk);
    return grid < 1 ? 1 : grid;
}

template <int CC, int LB, int VAR = 0, bool DW = false>
int launch_mma(const MmaArgs& a, cudaStream_t st)
{
    const int CdP = max(16, (a.Cd + 15) & ~15);
    const size_t smem = mma_smem_bytes(CC, 1 << LB, a.nch, CdP, a.max_ku, DW, (a.Cg + 15) & ~15);
    if (smem > 227 * 1024) return 0;
    CUtensorMap xmap;
    memset(&xmap, 0, sizeof(xmap));
    if ((VAR & 8) && ((CC << LB) % 64 == 0)) {
        // X as a 2-D tensor [nblk * Nsrc rows][nch * NC columns] of bf16, boxes of 64 columns (128 bytes) x 1 row
        if (!tma::make_map_2d(&xmap, a.X, 2, (uint64_t)a.nblk * a.Nsrc, (uint64_t)a.nch * (CC << LB), 64, 1)) return 0;
    }
    QC_CUDA((qc_set_max_smem<k_contract_mma<CC, LB, VAR, DW>>((int)smem)));
    k_contract_mma<CC, LB, VAR, DW><<<mma_grid(a), NTHREADS, smem, st>>>(a, xmap);
    QC_CHECK_LAUNCH();
    return 1;
}

// d W_last[(i*Cg + j)*H + h] = sum over the CTAs' partial accumulators, in CTA order (deterministic)
__global__ void k_dw_mma_reduce(const float* __restrict__ part, int ncta, int ngrp, int MDW, int CgP, int CC, int cpp, int Cs,
                                int Cg, int H, float* __restrict__ dw)
{
    const int total = Cs * Cg * H;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int h = t % H, j = (t / H) % Cg, i = t / (H * Cg);
        const int c = i / CC, il = i % CC;
        const int g = c / cpp, m = (c % cpp) * 8 * CC + h * CC + il;
        const float* p = part + ((size_t)g * MDW + m) * CgP + j;
        const size_t stride = (size_t)ngrp * MDW * CgP;
        float s0 = 0.f;
        for (int k = 0; k < ncta; ++k) s0 += p[(size_t)k * stride];
        dw[t] = s0;
    }
}

// fp32 [B][C][N] -> bf16 packed [nblk][N][nch][Bp][CC]: 64 points x (CC channels x Bp lanes) per block through a
// padded shared-memory tile; 128-byte row pieces on the [B,C,N] side, CC*2-byte pieces, 2*NC contiguous bytes per
// point, on the packed side.
constexpr int PTB = 64, RSB = 65;
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// TIn = float (the module's fp32 features / grad_out) or __nv_bfloat16 (features handed over in bf16: half the
// host-to-device and HBM read traffic; the values are the ones the tensor cores would see anyway)
template <typename TIn>
__global__ void __launch_bounds__(256) k_pack_bf16(const TIn* __restrict__ x, int B, int C, int N, int Bp, int CC,
                                                   __nv_bfloat16* __restrict__ xp)
{
    __shared__ float tile[128 * RSB];      // [(channel in chunk) * Bp + batch lane][point]
    const int n0 = blockIdx.x * PTB, c = blockIdx.y, blk = blockIdx.z, nch = gridDim.y;
    const int rows = CC * Bp;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (sizeof(TIn) == 4 && (N & 3) == 0) {
        // fp32 source with 16-byte aligned rows: half-warp per row, one 128-bit load per lane (like k_pack)
        const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
        const int n = n0 + hl * 4;
        for (int r = hw; r < rows; r += 16) {
            const int il = r / Bp, bl = r % Bp;
            const int ch = c * CC + il, b = blk * 32 + bl;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch < C && b < B && n < N) v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + ((size_t)b * C + ch) * N + n));
            float* t = tile + r * RSB + hl * 4;
            t[0] = v.x, t[1] = v.y, t[2] = v.z, t[3] = v.w;
        }
    } else {
        for (int r = w; r < rows; r += 8) {
            const int il = r / Bp, bl = r % Bp;
            const int ch = c * CC + il, b = blk * 32 + bl;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int n = n0 + half * 32 + lane;
                float v = 0.f;
                if (ch < C && b < B && n < N) v = to_f32(x[((size_t)b * C + ch) * N + n]);
                tile[r * RSB + half * 32 + lane] = v;
            }
        }
    }
    __syncthreads();
    const int per_pt = rows / 2;           // bf16x2 words per point and chunk
    const int per_pt4 = rows / 8;          // 16-byte pieces (8 bf16) per point and chunk
    for (int idx = threadIdx.x; idx < PTB * per_pt4; idx += 256) {
        const int nl = idx / per_pt4, w4 = idx % per_pt4;
        const int n = n0 + nl;
        if (n < N) {
            uint32_t v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = 8 * w4 + 2 * j, bl = e / CC, il = e % CC;
                v[j] = pack_bf16(tile[(il * Bp + bl) * RSB + nl], tile[((il + 1) * Bp + bl) * RSB + nl]);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(xp) + (((size_t)blk * N + n) * nch + c) * per_pt + 4 * w4) =
                make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
}

}  // namespace

extern "C" {

int qc_mma_get_variant(void) { return mma_var(); }

int qc_mma_set_variant(int var)
{
    const int old = mma_var();
    if (var >= 0) g_mma_var = var;
    return old;
}

int qc_mma_chunk_channels(int Cs, int B)
{
    // channels per column chunk: NC = Bp * CC <= 128 columns, CC in {4, 8, 16}
    const int Bp = qc_bp(B);
    if (Bp < 8) return 0;
    int cc = 128 / Bp;
    while (cc > 4 && cc / 2 >= Cs) cc /= 2;
    return cc;
}

int qc_pack_features_bf16(const void* x, int src_bf16, int B, int C, int N, int CC, void* xp, void* stream)
{
    const int Bp = qc_bp(B);
    if (B <= 0 || C <= 0 || N <= 0 || (CC != 4 && CC != 8 && CC != 16) || Bp * CC > 128) return QC_EINVAL;
    dim3 grid(qc_ceil_div(N, PTB), qc_ceil_div(C, CC), qc_nblk(B));
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(xp);
    if (src_bf16)
        k_pack_bf16<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), B, C, N, Bp, CC, out);
    else
        k_pack_bf16<float><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(x), B, C, N, Bp, CC, out);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_mma_smem_bytes(int Cs, int Cd, int B, int max_ku)
{
    const int CC = qc_mma_chunk_channels(Cs, B);
    if (CC == 0) return 0;
    return (int)mma_smem_bytes(CC, qc_bp(B), qc_ceil_div(Cs, CC), max(16, (Cd + 15) & ~15), max_ku);
}

int qc_contract_mma(const void* Xh, int Nsrc, int Cs, const void* tile_ent, const int* tile_pptr, const int* tile_uptr,
                    const int* union_idx, int ntiles, int max_ku, const float* phi, int H, int HPphi,
                    const float* w_last, int Cd, int transposed, const int* order, int Ndst, int B, float* Yp,
                    void* stream)
{
    const int Bp = qc_bp(B);
    const int CC = qc_mma_chunk_channels(Cs, B);
    if (CC == 0 || H > 8 || H < 1 || (HPphi != 4 && HPphi != 8) || HPphi < H || Cd > 64 || max_ku < KST || (max_ku % KST) ||
        max_ku > 0xffff)
        return QC_EUNSUP;
    if ((TP * Bp / 128) * max(16, (Cd + 15) & ~15) > 256) return QC_EUNSUP;   // TMEM: stage-C accumulators
    if ((size_t)Nsrc * qc_ceil_div(Cs, CC) * Bp * CC * 2 >= 0xffffffffull) return QC_EUNSUP;   // 32-bit row offsets
    MmaArgs a;
    a.X = reinterpret_cast<const __nv_bfloat16*>(Xh);
    a.Y = Yp;
    a.order = order;
    a.tile_ent = reinterpret_cast<const int2*>(tile_ent);
    a.tile_pptr = tile_pptr;
    a.tile_uptr = tile_uptr;
    a.union_idx = union_idx;
    a.phi = phi;
    a.w_last = w_last;
    a.Nsrc = Nsrc, a.Ndst = Ndst, a.Cs = Cs, a.Cd = Cd, a.H = H, a.HPphi = HPphi, a.nch = qc_ceil_div(Cs, CC);
    a.transposed = transposed, a.ntiles = ntiles, a.nblk = qc_nblk(B), a.max_ku = max_ku;
    a.G = nullptr, a.dw_part = nullptr, a.Cg = 0, a.nchg = 0, a.CCg = 4;
    a.dbg = nullptr;
    const cudaStream_t st = (cudaStream_t)stream;
#ifdef QC_MMA_PROFILE
    static long long* dbg = nullptr;
    if (!dbg) cudaMalloc(&dbg, 32 * sizeof(long long));
    cudaMemsetAsync(dbg, 0, 32 * sizeof(long long), st);
    a.dbg = dbg;
#endif
    int rc = 0;
    const int var = mma_var();
    if (var & 8) {   // union rows by TMA gather4 (falls through to the cp.async producers when the map cannot be made)
        if (Bp == 32) rc = launch_mma<4, 5, 12>(a, st);
        else if (Bp == 16) rc = CC == 8 ? launch_mma<8, 4, 12>(a, st) : launch_mma<4, 4, 12>(a, st);
        else if (Bp == 8 && CC >= 8) rc = CC == 16 ? launch_mma<16, 3, 12>(a, st) : launch_mma<8, 3, 12>(a, st);
    }
    if (rc != 0) {
    } else if (Bp == 32) rc = var == 2 ? launch_mma<4, 5, 2>(a, st) : (var == 5 ? launch_mma<4, 5, 5>(a, st) : launch_mma<4, 5, 4>(a, st));
    else if (Bp == 16) rc = CC == 8 ? launch_mma<8, 4, 4>(a, st) : launch_mma<4, 4, 4>(a, st);
    else if (Bp == 8) rc = CC == 16 ? launch_mma<16, 3, 4>(a, st) : (CC == 8 ? launch_mma<8, 3, 4>(a, st) : launch_mma<4, 3, 2>(a, st));
#ifdef QC_MMA_PROFILE
    {
        long long h[32];
        cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        const char* names[32] = {"prod total", "prod wait x_empty", "prod wait phi_empty", "prod cp.async wait", "prod phi build", 0, 0, 0,
                                 "mma total", "mma wait x_full", "mma wait d_empty", "mma wait ac_full", "mma wait phi_full",
                                 "mma wait y_empty", "mma proxy fence", 0, "epi total", "epi wait d_full", "epi wait ac_empty",
                                 "epi regroup", "epi wait y_full", "epi output"};
        fprintf(stderr, "[qc_contract_mma profile, CTA 0, cycles]");
        for (int i = 0; i < 32; ++i)
            if (names[i]) fprintf(stderr, " %s=%lld;", names[i], h[i]);
        fprintf(stderr, "\n");
    }
#endif
    if (rc == 0) return QC_EUNSUP;
    return rc == 1 ? 0 : rc;
}

// d W_last of the reduced-precision mode: same tiles and gather as the forward (sources = in points / features Xh), the
// regrouped stage-B result T contracted with the tile's grad_out rows Gh over (point, batch).  ws: scratch of
// qc_dw_mma_workspace_bytes(...) bytes (per-CTA partial accumulators, summed in CTA order).
size_t qc_dw_mma_workspace_bytes(int Cs, int Cg, int B)
{
    const int CC = qc_mma_chunk_channels(Cs, B);
    if (CC == 0) return 0;
    const int mdw = 8 * CC, cpp = 1, nch = qc_ceil_div(Cs, CC);
    return (size_t)qc_num_sms() * qc_ceil_div(nch, cpp) * mdw * ((Cg + 15) & ~15) * sizeof(float);
}

int qc_dw_mma_smem_bytes(int Cs, int Cg, int B, int max_ku)
{
    const int CC = qc_mma_chunk_channels(Cs, B);
    if (CC == 0) return 0;
    return (int)mma_smem_bytes(CC, qc_bp(B), qc_ceil_div(Cs, CC), 0, max_ku, true, (Cg + 15) & ~15);
}

int qc_dw_mma(const void* Xh, int Nsrc, int Cs, const void* Gh, int Cg, const void* tile_ent, const int* tile_pptr,
              const int* tile_uptr, const int* union_idx, int ntiles, int max_ku, const float* phi, int H, int HPphi,
              const int* order, int Ndst, int B, void* ws, size_t ws_bytes, float* d_w_last, void* stream)
{
    const int Bp = qc_bp(B);
    const int CC = qc_mma_chunk_channels(Cs, B), CCg = qc_mma_chunk_channels(Cg, B);
    if (CC == 0 || CCg == 0 || H > 8 || H < 1 || (HPphi != 4 && HPphi != 8) || HPphi < H || Cg > 64 || max_ku < KST ||
        (max_ku % KST) || max_ku > 0xffff)
        return QC_EUNSUP;
    const int mdw = 8 * CC, cpp = 1, nch = qc_ceil_div(Cs, CC), ngrp = qc_ceil_div(nch, cpp);
    const int CgP = (Cg + 15) & ~15;
    if (ngrp * CgP > 256) return QC_EUNSUP;                                  // TMEM: one accumulator per chunk group
    if ((size_t)Nsrc * nch * Bp * CC * 2 >= 0xffffffffull) return QC_EUNSUP;  // 32-bit row offsets
    if (ws == nullptr || ws_bytes < qc_dw_mma_workspace_bytes(Cs, Cg, B)) return QC_EWORKSPACE;
    MmaArgs a;
    a.X = reinterpret_cast<const __nv_bfloat16*>(Xh);
    a.Y = nullptr;
    a.order = order;
    a.tile_ent = reinterpret_cast<const int2*>(tile_ent);
    a.tile_pptr = tile_pptr;
    a.tile_uptr = tile_uptr;
    a.union_idx = union_idx;
    a.phi = phi;
    a.w_last = nullptr;
    a.G = reinterpret_cast<const __nv_bfloat16*>(Gh);
    a.dw_part = reinterpret_cast<float*>(ws);
    a.Cg = Cg, a.nchg = qc_ceil_div(Cg, CCg), a.CCg = CCg;
    a.Nsrc = Nsrc, a.Ndst = Ndst, a.Cs = Cs, a.Cd = 0, a.H = H, a.HPphi = HPphi, a.nch = nch;
    a.transposed = 0, a.ntiles = ntiles, a.nblk = qc_nblk(B), a.max_ku = max_ku;
    a.dbg = nullptr;
    const cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (mma_var() & 8) {
        if (Bp == 32) rc = launch_mma<4, 5, 12, true>(a, st);
        else if (Bp == 16) rc = CC == 8 ? launch_mma<8, 4, 12, true>(a, st) : launch_mma<4, 4, 12, true>(a, st);
        else if (Bp == 8 && CC >= 8) rc = CC == 16 ? launch_mma<16, 3, 12, true>(a, st) : launch_mma<8, 3, 12, true>(a, st);
    }
    if (rc != 0) {
    } else if (Bp == 32) rc = launch_mma<4, 5, 4, true>(a, st);
    else if (Bp == 16) rc = CC == 8 ? launch_mma<8, 4, 4, true>(a, st) : launch_mma<4, 4, 4, true>(a, st);
    else if (Bp == 8) rc = CC == 16 ? launch_mma<16, 3, 4, true>(a, st) : (CC == 8 ? launch_mma<8, 3, 4, true>(a, st) : launch_mma<4, 3, 2, true>(a, st));
    if (rc == 0) return QC_EUNSUP;
    if (rc != 1) return rc;
    const int total = Cs * Cg * H;
    k_dw_mma_reduce<<<qc_ceil_div(total, 256), 256, 0, st>>>(a.dw_part, mma_grid(a), ngrp, mdw, CgP, CC, cpp, Cs, Cg, H, d_w_last);
    QC_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
