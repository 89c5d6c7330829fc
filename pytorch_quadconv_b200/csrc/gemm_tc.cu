// gemm_tc.cu -- fp32-accurate dense stage of the WIDE layer path on the tcgen05 tensor cores.
//
//     D[m][n] = sum_k A[m][k] * B[n][k]          A [M x K], B [N x K] row-major fp32 ("NT" GEMM)
//
// computed as three kind::tf32 MMAs per product (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, the same error-
// compensated split the fused kernels use: 2e-7 against float64 where one TF32 pass gives 5e-4).
// Both operands are K-major in global memory, which is what contract_wide.cu produces and consumes:
//     Y  = T  * W2^T   (A = T   [B*No x Ci*H], B = W2^T [Co   x Ci*H])      output transposed per batch
//     dT = G^T * W2    (A = G^T [B*No x Co],   B = W2   [Ci*H x Co])        output row-major
//     dX = T' * W2t^T  (A = T'  [B*Ni x Co*H], B = W2t^T[Ci   x Co*H])      output transposed per batch
// (the fourth GEMM of the path, dW = X * T', reduces over the ROW index of T': k_gemm_tn_tf32x3* further down.)
//
// One CTA = 8 warps = a 128-row x BN-column tile with the accumulator in TMEM.  The 32-wide K chunks of
// A and B arrive by cp.async (two chunks ahead; FIFO commit groups, because register prefetching is serialised
// by ptxas' single-scoreboard allocation) straight into 128-byte-swizzled K-major operand tiles: the tensor core
// reads only the upper 19 bits of an fp32 operand, so the raw tile IS the tf32 hi operand (verified: identical
// results with and without masking); every thread then derives the lo tile (v - trunc(v)) from the pieces it
// copied itself, one thread issues the 12 MMAs of the chunk and commits them to the stage's mbarrier.
//
// tcgen05.mma accumulates with truncation (~2e-8 relative per MMA, measured: 1.5e-5 after K = 2048), so the
// TMEM accumulation only runs over groups of GROUP chunks (48 MMAs); two TMEM accumulators alternate, and
// while one group is multiplied the previous group's result is added to fp32 registers (round to nearest).
#include <cstdint>
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int GK = 32;        // K chunk
constexpr int GROUP = 4;      // K chunks accumulated in TMEM before the partial sum moves to registers

struct GemmArgs {
    const float* A;
    const float* B;
    float* D;
    long long M;
    int N, K;
    long long rows_per_batch;   // 0: D[m*N + n];  > 0: D[((m / rpb) * N + n) * rpb + m % rpb]
};

constexpr int GTHREADS = 256;
constexpr int RAW = 4;        // stages of the raw (= hi) operand ring filled by cp.async
constexpr int LOS = 2;        // stages of the lo operand tiles
constexpr int AHEAD = 2;      // K chunks in flight ahead of the one being multiplied

template <int BN>
__global__ void __launch_bounds__(GTHREADS, 1)
k_gemm_tf32x3(const GemmArgs g)
{
    constexpr int A_TILE = 128 * GK;             // floats per A tile
    constexpr int B_TILE = BN * GK;
    constexpr int TILE = A_TILE + B_TILE;        // one stage: A tile then B tile
    constexpr int HN = BN / 2;                   // output columns per thread (two threads share a row)
    static_assert(BN == 64 || BN == 128, "tile widths");
    extern __shared__ __align__(1024) float smem_raw[];
    // operand tiles use the 128-byte swizzle: 1 KB alignment.  [RAW][TILE] raw fp32 (the tensor core reads the
    // upper 19 bits, i.e. the raw tile IS the tf32 "hi" operand), then [LOS][TILE] lo = v - trunc(v)
    float* raw_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* lo_s = raw_s + (size_t)RAW * TILE;
    __shared__ __align__(8) uint64_t bars[RAW];    // MMAs of the chunk that used raw stage i have finished
    __shared__ __align__(8) uint64_t gbars[2];     // one per TMEM accumulator: its group of MMAs has finished
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wq = warp & 3, wh = warp >> 2;     // TMEM lane quadrant (rows wq*32..), column / K half
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * BN;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 2 * BN);
    if (tid == 0) {
        for (int i = 0; i < RAW; ++i) tc::mbar_init(&bars[i], 1);
        tc::mbar_init(&gbars[0], 1);
        tc::mbar_init(&gbars[1], 1);
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    const long long m = m0 + wq * 32 + lane;     // epilogue: this thread's output row (TMEM lane wq*32 + lane)
    const bool mvalid = m < g.M;
    const int nk = (g.K + GK - 1) / GK;

    // Copy mapping: the 8 warps move 128 rows x 128 bytes per chunk with four 16-byte cp.async per lane.  Warp
    // (wq, wh) covers rows wq*32.. and the K half wh; in copy j the lane handles row j*8 + lane%8 of those 32 and
    // the 16-byte piece q = wh*4 + lane/8 (eight lanes share a 64-byte run of one row).  Destination: K-major,
    // SWIZZLE_128B -- a row is one 128-byte line (GK = 32 floats), 8 rows form a 1 KB atom, the piece index is
    // XORed with the row index inside the atom.  Every thread later reads back exactly the pieces it copied.
    const int q = wh * 4 + (lane >> 3);
    uint32_t poff[4];                            // byte offsets of this thread's pieces inside an A or B tile
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int rl = wq * 32 + j * 8 + (lane & 7);
        poff[j] = (uint32_t)((rl >> 3) * 1024 + (rl & 7) * 128 + ((q ^ (rl & 7)) * 16));
    }
    const bool has_b = wq * 32 < BN;
    auto issue = [&](int kc) {
        const int k0 = kc * GK;
        const bool kv = k0 + 4 * q < g.K;        // K is a multiple of 4 (checked by the host)
        const uint32_t st = tc::smem_u32(raw_s + (size_t)(kc % RAW) * TILE);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rl = wq * 32 + j * 8 + (lane & 7);
            const bool av = kv && m0 + rl < g.M;
            tc::cp_async16_zfill(st + poff[j], g.A + (av ? (m0 + rl) * (long long)g.K + k0 + 4 * q : 0), av);
            if (has_b) {
                const bool bv = kv && n0 + rl < g.N;
                tc::cp_async16_zfill(st + A_TILE * 4 + poff[j], g.B + (bv ? (long long)(n0 + rl) * g.K + k0 + 4 * q : 0), bv);
            }
        }
    };
    auto make_lo = [&](int kc) {
        const float* rs = raw_s + (size_t)(kc % RAW) * TILE;
        float* ls = lo_s + (size_t)(kc % LOS) * TILE;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (t == 1 && !has_b) continue;
                const uint32_t o = (t * A_TILE * 4 + poff[j]) / 4;
                const float4 v = *reinterpret_cast<const float4*>(rs + o);
                float4 l;
                l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                *reinterpret_cast<float4*>(ls + o) = l;
            }
        }
    };

    const uint32_t idesc = tc::make_idesc_tf32(128, BN);
    constexpr uint32_t SBO = 1024;                      // between 8-row atoms
    constexpr uint64_t SW128 = (uint64_t)2 << 61;       // descriptor layout type SWIZZLE_128B
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const int ngroups = (nk + GROUP - 1) / GROUP;
    float acc[HN];
#pragma unroll
    for (int j = 0; j < HN; ++j) acc[j] = 0.f;
    int gread = 0;     // groups already added to acc
    auto take_group = [&](int gi) {   // acc += this thread's columns of D[gi & 1] once the group's MMAs have finished
        tc::mbar_wait(&gbars[gi & 1], (gi >> 1) & 1);
        tc::fence_after_sync();
#pragma unroll
        for (int c = 0; c < HN / 16; ++c) {
            uint32_t r[16];
            tc::tmem_ld16(lane_base + tmem + (gi & 1) * BN + wh * HN + c * 16, r);
            tc::wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[j]);
        }
    };

#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
        if (u < nk) issue(u);
        tc::cp_async_commit();
    }
    for (int kc = 0; kc < nk; ++kc) {
        // chunk kc-2 used raw stage (kc+2) % RAW and lo stage kc % LOS: both are free once its MMAs are done
        if (kc >= 2) tc::mbar_wait(&bars[(kc - 2) % RAW], ((kc - 2) / RAW) & 1);
        if (kc + AHEAD < nk) issue(kc + AHEAD);
        tc::cp_async_commit();
        tc::cp_async_wait<AHEAD>();                  // this thread's copies of chunk kc have landed
        make_lo(kc);
        const int gi = kc / GROUP;
        if (kc % GROUP == GROUP / 2 && gi >= 1) {    // the previous group finished long ago: no stall
            take_group(gi - 1);
            gread = gi;
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t ahi = tc::smem_u32(raw_s + (size_t)(kc % RAW) * TILE), bhi = ahi + A_TILE * 4;
            const uint32_t alo = tc::smem_u32(lo_s + (size_t)(kc % LOS) * TILE), blo = alo + A_TILE * 4;
#pragma unroll
            for (int ks = 0; ks < GK / 8; ++ks) {
                // one k-step = 8 tf32 = 32 bytes along the (unswizzled) row; LBO is not used by swizzled K-major tiles
                const uint64_t dah = tc::make_smem_desc(ahi + ks * 32, 16, SBO) | SW128;
                const uint64_t dal = tc::make_smem_desc(alo + ks * 32, 16, SBO) | SW128;
                const uint64_t dbh = tc::make_smem_desc(bhi + ks * 32, 16, SBO) | SW128;
                const uint64_t dbl = tc::make_smem_desc(blo + ks * 32, 16, SBO) | SW128;
                const uint32_t d = tmem + (gi & 1) * BN;
                tc::mma_tf32_ss(d, dah, dbh, idesc, !(kc % GROUP == 0 && ks == 0));
                tc::mma_tf32_ss(d, dal, dbh, idesc, true);
                tc::mma_tf32_ss(d, dah, dbl, idesc, true);
            }
            tc::mma_commit(&bars[kc % RAW]);
            if (kc % GROUP == GROUP - 1 || kc == nk - 1) tc::mma_commit(&gbars[gi & 1]);
        }
    }
    // the groups not yet taken (their commits also cover every stage's last MMAs)
    for (int gi = gread; gi < ngroups; ++gi) take_group(gi);

    // epilogue: thread = (row, column half)
    const int nb = n0 + wh * HN;
    if (mvalid) {
        if (g.rows_per_batch > 0) {
            const long long bb = m / g.rows_per_batch, o = m % g.rows_per_batch;
            float* dst = g.D + bb * g.N * g.rows_per_batch + o;
#pragma unroll
            for (int j = 0; j < HN; ++j)
                if (nb + j < g.N) dst[(long long)(nb + j) * g.rows_per_batch] = acc[j];   // lanes = consecutive rows: coalesced
        } else {
            float* dst = g.D + m * g.N + nb;
#pragma unroll
            for (int q4 = 0; q4 < HN / 4; ++q4) {
                if ((g.N & 3) == 0 && nb + 4 * q4 + 4 <= g.N) {
                    reinterpret_cast<float4*>(dst)[q4] = make_float4(acc[4 * q4], acc[4 * q4 + 1], acc[4 * q4 + 2], acc[4 * q4 + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (nb + 4 * q4 + e < g.N) dst[4 * q4 + e] = acc[4 * q4 + e];
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 2 * BN);
}

// ------------------------------------------------------------------------------------------------------------
// TMA-fed, warp-specialised version of the same GEMM (the default; the cp.async kernel above stays as the fallback
// for hosts whose driver refuses the tensor maps and as the A/B baseline, QC_GEMM_NO_TMA=1).
//
// The loader above is per-thread: every thread copies its pieces, waits for them, derives the lo tile, and the CTA
// meets at a __syncthreads per K chunk -- copy, split and MMA take turns (ncu: 1.2 TB/s of DRAM, tensor pipe 22 %).
// Here the three roles run decoupled, handing stages over through mbarriers only:
//     warp 8  (one lane)  TMA producer: cp.async.bulk.tensor.2d of the A box (128 rows x 32 k) and the B box (BN x 32 k)
//                         per chunk into a TS-stage ring; the copy engine writes the 128-byte-swizzled K-major
//                         layout the MMA descriptors expect and zero-fills ragged M / N / K edges
//     warps 0-7           lo makers: lo = v - trunc_tf32(v) of the landed stage into a 2-stage lo ring (the raw tile IS
//                         the hi operand); they also fold finished TMEM accumulation groups into fp32 registers
//                         (bounded chains, see the truncation note at the top) and write the epilogue
//     warp 9  (one lane)  MMA issuer: 12 kind::tf32 MMAs per chunk, tcgen05.commit frees the raw and lo stages
constexpr int TS = 4;             // raw (= hi) stages filled by TMA
constexpr int TL = 2;             // lo stages
constexpr int TMA_THREADS = 320;  // 8 lo/epilogue warps + producer warp + MMA warp

template <int BN>
__global__ void __launch_bounds__(TMA_THREADS, 1)
k_gemm_tf32x3_tma(const GemmArgs g, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap)
{
    constexpr int A_TILE = 128 * GK;
    constexpr int B_TILE = BN * GK;
    constexpr int TILE = A_TILE + B_TILE;
    constexpr int HN = BN / 2;
    extern __shared__ __align__(1024) float smem_raw[];
    float* raw_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* lo_s = raw_s + (size_t)TS * TILE;
    __shared__ __align__(8) uint64_t full[TS], raw_free[TS], lo_full[TL], lo_free[TL], acc_full[2], acc_free[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * BN;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 2 * BN);
    if (tid == 0) {
        for (int i = 0; i < TS; ++i) tc::mbar_init(&full[i], 1), tc::mbar_init(&raw_free[i], 1);
        for (int i = 0; i < TL; ++i) tc::mbar_init(&lo_full[i], 256), tc::mbar_init(&lo_free[i], 1);
        for (int i = 0; i < 2; ++i) tc::mbar_init(&acc_full[i], 1), tc::mbar_init(&acc_free[i], 256);
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int nk = (g.K + GK - 1) / GK;
    const int ngroups = (nk + GROUP - 1) / GROUP;

    if (warp == 8) {
        // ---- TMA producer ------------------------------------------------------------------------------------
        if (lane == 0) {
            tma::prefetch_map(&amap);
            tma::prefetch_map(&bmap);
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % TS;
                if (kc >= TS) tc::mbar_wait(&raw_free[s], ((kc / TS) - 1) & 1);
                tma::mbar_arrive_expect_tx(&full[s], TILE * 4);
                const uint32_t dst = tc::smem_u32(raw_s + (size_t)s * TILE);
                tma::load_2d(dst, &amap, &full[s], kc * GK, (int)m0);
                tma::load_2d(dst + A_TILE * 4, &bmap, &full[s], kc * GK, n0);
            }
        }
    } else if (warp == 9) {
        // ---- MMA issuer ----------------------------------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(128, BN);
            constexpr uint32_t SBO = 1024;
            constexpr uint64_t SW128 = (uint64_t)2 << 61;
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % TS, l = kc % TL, gi = kc / GROUP;
                if (kc % GROUP == 0 && gi >= 2) tc::mbar_wait(&acc_free[gi & 1], ((gi >> 1) - 1) & 1);   // group gi-2 has been read out
                tc::mbar_wait(&full[s], (kc / TS) & 1);
                tc::mbar_wait(&lo_full[l], (kc / TL) & 1);
                tc::fence_after_sync();
                const uint32_t ahi = tc::smem_u32(raw_s + (size_t)s * TILE), bhi = ahi + A_TILE * 4;
                const uint32_t alo = tc::smem_u32(lo_s + (size_t)l * TILE), blo = alo + A_TILE * 4;
                const uint32_t d = tmem + (gi & 1) * BN;
#pragma unroll
                for (int ks = 0; ks < GK / 8; ++ks) {
                    const uint64_t dah = tc::make_smem_desc(ahi + ks * 32, 16, SBO) | SW128;
                    const uint64_t dal = tc::make_smem_desc(alo + ks * 32, 16, SBO) | SW128;
                    const uint64_t dbh = tc::make_smem_desc(bhi + ks * 32, 16, SBO) | SW128;
                    const uint64_t dbl = tc::make_smem_desc(blo + ks * 32, 16, SBO) | SW128;
                    tc::mma_tf32_ss(d, dah, dbh, idesc, !(kc % GROUP == 0 && ks == 0));
                    tc::mma_tf32_ss(d, dal, dbh, idesc, true);
                    tc::mma_tf32_ss(d, dah, dbl, idesc, true);
                }
                tc::mma_commit(&raw_free[s]);
                tc::mma_commit(&lo_free[l]);
                if (kc % GROUP == GROUP - 1 || kc == nk - 1) tc::mma_commit(&acc_full[gi & 1]);
            }
        }
    } else {
        // ---- lo makers + accumulation + epilogue (warps 0-7) -------------------------------------------------------
        const int wq = warp & 3, wh = warp >> 2;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        float acc[HN];
#pragma unroll
        for (int j = 0; j < HN; ++j) acc[j] = 0.f;
        auto take_group = [&](int gi) {
            tc::mbar_wait(&acc_full[gi & 1], (gi >> 1) & 1);
            tc::fence_after_sync();
#pragma unroll
            for (int c = 0; c < HN / 16; ++c) {
                uint32_t r[16];
                tc::tmem_ld16(lane_base + tmem + (gi & 1) * BN + wh * HN + c * 16, r);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[j]);
            }
            tc::fence_before_sync();
            tc::mbar_arrive(&acc_free[gi & 1]);
        };
        int gread = 0;
        for (int kc = 0; kc < nk; ++kc) {
            const int s = kc % TS, l = kc % TL;
            tc::mbar_wait(&full[s], (kc / TS) & 1);
            if (kc >= TL) tc::mbar_wait(&lo_free[l], ((kc / TL) - 1) & 1);
            const float4* rs = reinterpret_cast<const float4*>(raw_s + (size_t)s * TILE);
            float4* ls = reinterpret_cast<float4*>(lo_s + (size_t)l * TILE);
#pragma unroll
            for (int j = 0; j < TILE / 4 / 256; ++j) {        // elementwise on the landed bytes: the swizzle does not matter
                const float4 v = rs[tid + 256 * j];
                float4 o;
                o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                ls[tid + 256 * j] = o;
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(&lo_full[l]);
            // a finished group moves to registers two chunks into the next one (its MMAs retired long ago)
            const int gi = kc / GROUP;
            if (kc % GROUP == GROUP / 2 && gi >= 1) {
                take_group(gi - 1);
                gread = gi;
            }
        }
        for (int gi = gread; gi < ngroups; ++gi) take_group(gi);

        const long long m = m0 + wq * 32 + lane;
        const int nb = n0 + wh * HN;
        if (m < g.M) {
            if (g.rows_per_batch > 0) {
                const long long bb = m / g.rows_per_batch, o = m % g.rows_per_batch;
                float* dst = g.D + bb * g.N * g.rows_per_batch + o;
#pragma unroll
                for (int j = 0; j < HN; ++j)
                    if (nb + j < g.N) dst[(long long)(nb + j) * g.rows_per_batch] = acc[j];
            }
        }
        if (g.rows_per_batch == 0) {
            // row-major output: a thread holds HN columns of ONE row, so direct stores would put 16 bytes into each of
            // 32 rows per instruction.  The tile goes through shared memory (the operand ring is free: every MMA has
            // retired) and leaves as whole rows, 512 contiguous bytes per warp instruction.
            constexpr int LD = BN + 4;                 // row pitch in floats: quarter-warps hit distinct banks
            float* stg = raw_s;
            float* srow = stg + (size_t)(wq * 32 + lane) * LD + wh * HN;
#pragma unroll
            for (int q4 = 0; q4 < HN / 4; ++q4)
                *reinterpret_cast<float4*>(srow + 4 * q4) = make_float4(acc[4 * q4], acc[4 * q4 + 1], acc[4 * q4 + 2], acc[4 * q4 + 3]);
            tc::named_bar_sync(1, 256);
            constexpr int LPR = BN / 4;                // lanes per row (float4 each)
            constexpr int RPI = 32 / LPR;              // rows per warp instruction
            const int c4 = (lane % LPR) * 4, rsub = lane / LPR;
            const bool vec = (g.N & 3) == 0 && (reinterpret_cast<uintptr_t>(g.D) & 15) == 0;
            for (int r = warp * RPI + rsub; r < 128; r += 8 * RPI) {
                const long long mr = m0 + r;
                if (mr >= g.M || n0 + c4 >= g.N) continue;
                const float4 v = *reinterpret_cast<const float4*>(stg + (size_t)r * LD + c4);
                float* dst = g.D + mr * g.N + n0 + c4;
                if (vec) {
                    *reinterpret_cast<float4*>(dst) = v;
                } else {
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (n0 + c4 + k < g.N) dst[k] = e[k];
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 2 * BN);
}

template <int BN>
int launch_gemm_tma(const GemmArgs& g, cudaStream_t st)
{
    if (g.M > 2147483647LL) return QC_EUNSUP;
    CUtensorMap amap, bmap;
    if (!tma::make_map_2d(&amap, g.A, 4, (uint64_t)g.M, (uint64_t)g.K, GK, 128)) return QC_EUNSUP;
    if (!tma::make_map_2d(&bmap, g.B, 4, (uint64_t)g.N, (uint64_t)g.K, GK, BN)) return QC_EUNSUP;
    const size_t smem = (size_t)(TS + TL) * (128 * GK + BN * GK) * 4 + 1024;
    QC_CUDA((qc_set_max_smem<k_gemm_tf32x3_tma<BN>>((int)smem)));
    const long long mt = (g.M + 127) / 128;
    const int nt = (g.N + BN - 1) / BN;
    if (nt > 65535) return QC_EUNSUP;
    k_gemm_tf32x3_tma<BN><<<dim3((unsigned)mt, (unsigned)nt), TMA_THREADS, smem, st>>>(g, amap, bmap);
    QC_CHECK_LAUNCH();
    return 0;
}

template <int BN>
int launch_gemm(const GemmArgs& g, cudaStream_t st)
{
    const size_t smem = (size_t)(RAW + LOS) * (128 * GK + BN * GK) * 4 + 1024;
    QC_CUDA((qc_set_max_smem<k_gemm_tf32x3<BN>>((int)smem)));
    const long long mt = (g.M + 127) / 128;
    const int nt = (g.N + BN - 1) / BN;
    if (mt > 2147483647LL || nt > 65535) return QC_EUNSUP;
    k_gemm_tf32x3<BN><<<dim3((unsigned)mt, (unsigned)nt), GTHREADS, smem, st>>>(g);
    QC_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" int qc_gemm_nt_tf32x3(const float* A, const float* B, float* D, long long M, int N, int K,
                                  long long rows_per_batch, void* stream)
{
    if (M <= 0 || N <= 0 || K <= 0 || rows_per_batch < 0) return QC_EINVAL;
    if ((K & 3) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(B) & 15) != 0)
        return QC_EUNSUP;   // 128-bit row pieces
    if (rows_per_batch > 0 && M % rows_per_batch != 0) return QC_EINVAL;
    GemmArgs g{A, B, D, M, N, K, rows_per_batch};
    cudaStream_t st = (cudaStream_t)stream;
    if (getenv("QC_GEMM_NO_TMA") == nullptr) {   // read per call: the tests toggle it (these are large launches)
        const int rc = N <= 64 ? launch_gemm_tma<64>(g, st) : launch_gemm_tma<128>(g, st);
        if (rc != QC_EUNSUP) return rc;       // tensor map refused (driver / stride): the cp.async kernel serves it
    }
    return N <= 64 ? launch_gemm<64>(g, st) : launch_gemm<128>(g, st);
}

// ------------------------------------------------------------------------------------------------------------
// "TN" GEMM: the reduction runs over the ROW index of the big operand.
//
//     D[n][m] = sum_k X(n, k) * T[k][m]        T [K x M] row-major,  X [nb][N][Kb] (k = b*Kb + kk),  D [N x M]
//
// This is dW_last of the wide path (dW[i][(j,h)] = sum_{b,p} x[b][i][p] * T'[b][p][(j,h)]): T' (1.3 GB at C4) is read
// exactly once, so the kernel is HBM bound and built around that stream.  kind::tf32 has no MN-major shared-memory
// operands (tc_selftest.cu), so T takes the other road into the tensor core: thread = one column m of T = one TMEM
// lane; it loads 64 consecutive rows of its column with plain coalesced loads (a warp instruction = 128 contiguous
// bytes of one row), splits them into tf32 hi / lo in registers and stores them with tcgen05.st as the A operand
// (lane = m, column = k) -- no shared-memory transpose.  X, the small operand (re-read from L2 by every tile), arrives
// K-major by 4-byte cp.async straight from the reference's [B, C, N] layout into 128-byte-swizzled tiles (raw = hi,
// lo derived in place by the copying thread).  One CTA = one 128-column tile of T x 64 rows of X x one K split; its
// 8 warps are two independent sets of 4 that alternate over 64-row chunks (own TMEM operand + accumulator, own
// mbarrier), so one set's loads are in flight while the other multiplies.  Accumulation chains in TMEM are bounded
// to TN_GROUP chunks (48 MMAs, see the truncation note at the top) and continue in fp32 registers.  K splits write
// partial results that k_sum_slices adds in a fixed order (deterministic).
namespace {

constexpr int TK = 64;            // rows of T per chunk
constexpr int TBN = 64;           // rows of X (output rows) per CTA
constexpr int TN_GROUP = 2;       // chunks per TMEM accumulation chain (48 MMAs, like GROUP above)
constexpr int TN_THREADS = 256;
constexpr int XS_TILE = TBN * TK; // floats per X tile (hi or lo): two 32-k sub-tiles of [64 rows x 128 bytes]

struct TnArgs {
    const float* T;
    const float* X;
    float* out;            // [S][N][M] partial results (D itself when S == 1)
    int nb, Kb, M, N;
    int cpb;               // chunks per batch block = ceil(Kb / TK)
    int nchunks, cps;      // total chunks, chunks per K split
};

__global__ void __launch_bounds__(TN_THREADS, 1)
k_gemm_tn_tf32x3(const TnArgs g)
{
    extern __shared__ __align__(1024) float smem_raw[];
    // [set][stage][hi tile, lo tile], each tile 16 KB, 1 KB aligned (128-byte swizzle)
    float* xs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int set = warp >> 2, wq = warp & 3;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.z * TBN;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    // TMEM columns of this set: A_hi [0,64) A_lo [64,128) D [128,192)
    const uint32_t a_hi = tmem + set * 192, a_lo = a_hi + TK, d_acc = a_hi + 2 * TK;
    const uint32_t idesc = tc::make_idesc_tf32(128, TBN);
    constexpr uint64_t SW128 = (uint64_t)2 << 61;

    const int m = m0 + wq * 32 + lane;
    const bool mvalid = m < g.M;
    const int cbeg = blockIdx.y * g.cps, cend = min(cbeg + g.cps, g.nchunks);
    const int nit = (cend - cbeg - set + 1) / 2;           // chunks cbeg + set, cbeg + set + 2, ...
    float* xset = xs + (size_t)set * 4 * XS_TILE;

    // X copies: warp wq moves rows wq*16 .. wq*16+15 of the tile, 64 k each: lane = k (two instructions per row)
    auto issue_x = [&](int it) {
        const int c = cbeg + set + 2 * it;
        const int b = c / g.cpb, kk0 = (c % g.cpb) * TK;
        const uint32_t st = tc::smem_u32(xset + (size_t)(it & 1) * 2 * XS_TILE);
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const int nl = wq * 16 + (j >> 1), kt = j & 1, k = kt * 32 + lane;
            const bool v = n0 + nl < g.N && kk0 + k < g.Kb;
            const uint32_t off = (uint32_t)(kt * (TBN * 128) + (nl >> 3) * 1024 + (nl & 7) * 128 + (((lane >> 2) ^ (nl & 7)) * 16) + (lane & 3) * 4);
            const float* src = g.X + (v ? ((long long)b * g.N + n0 + nl) * g.Kb + kk0 + k : 0);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(st + off), "l"(src), "r"(v ? 4 : 0) : "memory");
        }
    };
    auto make_lo = [&](int it) {
        float* hs = xset + (size_t)(it & 1) * 2 * XS_TILE;
        float* ls = hs + XS_TILE;
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const int nl = wq * 16 + (j >> 1), kt = j & 1;
            const uint32_t o = (uint32_t)(kt * (TBN * 128) + (nl >> 3) * 1024 + (nl & 7) * 128 + (((lane >> 2) ^ (nl & 7)) * 16) + (lane & 3) * 4) / 4;
            const float v = hs[o];
            ls[o] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        }
    };

    float acc[TBN];
#pragma unroll
    for (int j = 0; j < TBN; ++j) acc[j] = 0.f;
    auto take = [&]() {            // acc += D (the chain's MMAs have finished)
        tc::fence_after_sync();
#pragma unroll
        for (int c = 0; c < TBN / 16; ++c) {
            uint32_t r[16];
            tc::tmem_ld16(lane_base + d_acc + c * 16, r);
            tc::wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[j]);
        }
    };

    if (nit > 0) issue_x(0);
    tc::cp_async_commit();
    for (int it = 0; it < nit; ++it) {
        const int c = cbeg + set + 2 * it;
        const int b = c / g.cpb, kk0 = (c % g.cpb) * TK;
        // this thread's column of the chunk: 64 independent coalesced loads in flight
        float v[TK];
        const float* tp = g.T + ((long long)b * g.Kb + kk0) * g.M + m;
#pragma unroll
        for (int t = 0; t < TK; ++t)
            v[t] = (mvalid && kk0 + t < g.Kb) ? __ldcs(tp + (long long)t * g.M) : 0.f;
        // the previous chunk's MMAs have read the A operand in TMEM and X stage (it+1)&1
        if (it > 0) tc::mbar_wait(&bars[set], (it - 1) & 1);
        if (it > 0 && it % TN_GROUP == 0) take();
        if (it + 1 < nit) issue_x(it + 1);
        tc::cp_async_commit();
        tc::cp_async_wait<1>();
        make_lo(it);
#pragma unroll
        for (int cc = 0; cc < TK / 16; ++cc) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float x = v[cc * 16 + j];
                const float hv = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                hi[j] = __float_as_uint(hv);
                lo[j] = __float_as_uint(x - hv);
            }
            tc::tmem_st16(lane_base + a_hi + cc * 16, hi);
            tc::tmem_st16(lane_base + a_lo + cc * 16, lo);
        }
        tc::wait_st();
        tc::fence_proxy_async();
        tc::fence_before_sync();
        tc::named_bar_sync(1 + set, 128);
        if (wq == 0 && lane == 0) {
            tc::fence_after_sync();
            const uint32_t bhi = tc::smem_u32(xset + (size_t)(it & 1) * 2 * XS_TILE), blo = bhi + XS_TILE * 4;
#pragma unroll
            for (int ks = 0; ks < TK / 8; ++ks) {
                const uint32_t o = (uint32_t)((ks >> 2) * (TBN * 128) + (ks & 3) * 32);
                const uint64_t dh = tc::make_smem_desc(bhi + o, 16, 1024) | SW128;
                const uint64_t dl = tc::make_smem_desc(blo + o, 16, 1024) | SW128;
                tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dh, idesc, !(it % TN_GROUP == 0 && ks == 0));
                tc::mma_tf32_ts(d_acc, a_lo + ks * 8, dh, idesc, true);
                tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dl, idesc, true);
            }
            tc::mma_commit(&bars[set]);
        }
    }
    if (nit > 0) {
        tc::mbar_wait(&bars[set], (nit - 1) & 1);
        take();
    }
    // set 1 hands its sums to set 0 through shared memory (all MMAs have finished: the X tiles are free)
    tc::fence_before_sync();
    __syncthreads();
    float* red = xs;                       // [TBN][128]
    const int row = wq * 32 + lane;
    if (set == 1) {
#pragma unroll
        for (int j = 0; j < TBN; ++j) red[j * 128 + row] = acc[j];
    }
    __syncthreads();
    if (set == 0 && mvalid) {
        float* dst = g.out + (size_t)blockIdx.y * g.N * g.M + m;
#pragma unroll
        for (int j = 0; j < TBN; ++j)
            if (n0 + j < g.N) dst[(size_t)(n0 + j) * g.M] = acc[j] + red[j * 128 + row];
    }
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// TMA-staged version of the TN GEMM (the default when M % 4 == 0; the kernel above is the fallback).  The register
// loads above leave ~40 KB in flight per SM and make each set wait out the full DRAM latency per chunk (1.8 TB/s
// measured).  Here a producer lane keeps TST chunks of T (32 KB each: four 32-column x 64-row boxes, 128-byte swizzle)
// in flight with cp.async.bulk.tensor.2d, and the two consumer sets (chunks alternate between them, as above) read
// their column of the landed stage from shared memory (lane = column: every warp instruction is one 128-byte line,
// conflict free).  X arrives in 16-byte pieces when Kb % 4 == 0 (X16), 4-byte pieces otherwise.
constexpr int TST = 3;             // T stages in flight
constexpr int TN2_THREADS = 288;   // 2 consumer sets of 4 warps + the producer warp

template <bool X16>
__global__ void __launch_bounds__(TN2_THREADS, 1)
k_gemm_tn_tf32x3_tma(const TnArgs g, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(1024) float smem_raw[];
    float* ts = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // [TST][4 boxes][64 x 32]
    float* xs = ts + (size_t)TST * TK * 128;                                                                   // [set][stage][hi, lo][XS_TILE]
    __shared__ __align__(8) uint64_t t_full[TST], t_free[TST], bars[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int m0 = blockIdx.x * 128, n0 = blockIdx.z * TBN;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < TST; ++i) tc::mbar_init(&t_full[i], 1), tc::mbar_init(&t_free[i], 128);
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int cbeg = blockIdx.y * g.cps, cend = min(cbeg + g.cps, g.nchunks);
    const int ntot = max(cend - cbeg, 0);
    float* red = xs;                       // [TBN][128] hand-off of set 1's sums (after every MMA has retired)
    float acc[TBN];
    const int set = (warp >> 2) & 1, wq = warp & 3;
    const int m = m0 + wq * 32 + lane;
    const bool mvalid = m < g.M;

    if (warp == 8) {
        if (lane == 0) {
            tma::prefetch_map(&tmap);
            for (int it = 0; it < ntot; ++it) {
                const int s = it % TST, c = cbeg + it;
                const int b = c / g.cpb, kk0 = (c % g.cpb) * TK;
                if (it >= TST) tc::mbar_wait(&t_free[s], ((it / TST) - 1) & 1);
                tma::mbar_arrive_expect_tx(&t_full[s], TK * 128 * 4);
                const uint32_t dst = tc::smem_u32(ts + (size_t)s * TK * 128);
#pragma unroll
                for (int q = 0; q < 4; ++q) tma::load_2d(dst + q * (TK * 128), &tmap, &t_full[s], m0 + 32 * q, b * g.Kb + kk0);
            }
        }
    } else {
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const uint32_t a_hi = tmem + set * 192, a_lo = a_hi + TK, d_acc = a_hi + 2 * TK;
        const uint32_t idesc = tc::make_idesc_tf32(128, TBN);
        constexpr uint64_t SW128 = (uint64_t)2 << 61;
        const int nit = (ntot - set + 1) / 2;                  // this set's chunks: cbeg + set + 2 i
        float* xset = xs + (size_t)set * 4 * XS_TILE;
        const int ts_tid = wq * 32 + lane;                     // thread index inside the set

        // X tile of this set's chunk i: 64 rows x 64 k, K-major, 128-byte swizzle (two 32-k sub-tiles)
        auto issue_x = [&](int i) {
            const int c = cbeg + set + 2 * i;
            const int b = c / g.cpb, kk0 = (c % g.cpb) * TK;
            const uint32_t st = tc::smem_u32(xset + (size_t)(i & 1) * 2 * XS_TILE);
            const float* xb = g.X + ((long long)b * g.N + n0) * g.Kb + kk0;
            if (X16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int p = ts_tid + 128 * j, nl = p >> 4, q16 = p & 15, kt = q16 >> 3, q = q16 & 7;
                    const bool v = n0 + nl < g.N && kk0 + 4 * q16 < g.Kb;
                    const uint32_t off = (uint32_t)(kt * (TBN * 128) + (nl >> 3) * 1024 + (nl & 7) * 128 + ((q ^ (nl & 7)) << 4));
                    tc::cp_async16_zfill(st + off, v ? xb + (long long)nl * g.Kb + 4 * q16 : g.X, v);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int nl = wq * 16 + (j >> 1), kt = j & 1, k = kt * 32 + lane;
                    const bool v = n0 + nl < g.N && kk0 + k < g.Kb;
                    const uint32_t off = (uint32_t)(kt * (TBN * 128) + (nl >> 3) * 1024 + (nl & 7) * 128 + (((lane >> 2) ^ (nl & 7)) * 16) + (lane & 3) * 4);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(st + off), "l"(v ? xb + (long long)nl * g.Kb + k : g.X), "r"(v ? 4 : 0) : "memory");
                }
            }
        };
        auto make_lo = [&](int i) {          // every thread splits exactly the pieces it copied
            float* hs = xset + (size_t)(i & 1) * 2 * XS_TILE;
            float* ls = hs + XS_TILE;
            if (X16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int p = ts_tid + 128 * j, nl = p >> 4, q16 = p & 15, kt = q16 >> 3, q = q16 & 7;
                    const uint32_t o = (uint32_t)(kt * (TBN * 128) + (nl >> 3) * 1024 + (nl & 7) * 128 + ((q ^ (nl & 7)) << 4)) / 4;
                    const float4 v = *reinterpret_cast<const float4*>(hs + o);
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    *reinterpret_cast<float4*>(ls + o) = l;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int nl = wq * 16 + (j >> 1), kt = j & 1;
                    const uint32_t o = (uint32_t)(kt * (TBN * 128) + (nl >> 3) * 1024 + (nl & 7) * 128 + (((lane >> 2) ^ (nl & 7)) * 16) + (lane & 3) * 4) / 4;
                    const float v = hs[o];
                    ls[o] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
                }
            }
        };
#pragma unroll
        for (int j = 0; j < TBN; ++j) acc[j] = 0.f;
        auto take = [&]() {            // acc += D (the chain's MMAs have finished)
            tc::fence_after_sync();
#pragma unroll
            for (int c = 0; c < TBN / 16; ++c) {
                uint32_t r[16];
                tc::tmem_ld16(lane_base + d_acc + c * 16, r);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[j]);
            }
        };

        if (nit > 0) issue_x(0);
        tc::cp_async_commit();
        for (int i = 0; i < nit; ++i) {
            const int it = set + 2 * i, s = it % TST;
            const int kk0 = ((cbeg + it) % g.cpb) * TK;
            tc::mbar_wait(&t_full[s], (it / TST) & 1);
            // this thread's column of the landed chunk: 64 rows, swizzled 128-byte lines
            float v[TK];
            const float* col = ts + (size_t)s * TK * 128 + wq * (TK * 32);
            const bool full = mvalid && kk0 + TK <= g.Kb;
#pragma unroll
            for (int t = 0; t < TK; ++t) {
                const float x = col[t * 32 + ((((lane >> 2) ^ (t & 7)) << 2) | (lane & 3))];
                v[t] = (full || (mvalid && kk0 + t < g.Kb)) ? x : 0.f;
            }
            // the previous chunk's MMAs of this set have read the A operand in TMEM and X stage (i+1)&1
            if (i > 0) tc::mbar_wait(&bars[set], (i - 1) & 1);
            if (i > 0 && i % TN_GROUP == 0) take();
            if (i + 1 < nit) issue_x(i + 1);
            tc::cp_async_commit();
            tc::cp_async_wait<1>();
            make_lo(i);
#pragma unroll
            for (int cc = 0; cc < TK / 16; ++cc) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x = v[cc * 16 + j];
                    const float hv = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                    hi[j] = __float_as_uint(hv);
                    lo[j] = __float_as_uint(x - hv);
                }
                tc::tmem_st16(lane_base + a_hi + cc * 16, hi);
                tc::tmem_st16(lane_base + a_lo + cc * 16, lo);
            }
            tc::wait_st();
            tc::mbar_arrive(&t_free[s]);                   // this thread's column of the stage has left shared memory
            tc::fence_proxy_async();
            tc::fence_before_sync();
            tc::named_bar_sync(1 + set, 128);
            if (wq == 0 && lane == 0) {
                tc::fence_after_sync();
                const uint32_t bhi = tc::smem_u32(xset + (size_t)(i & 1) * 2 * XS_TILE), blo = bhi + XS_TILE * 4;
#pragma unroll
                for (int ks = 0; ks < TK / 8; ++ks) {
                    const uint32_t o = (uint32_t)((ks >> 2) * (TBN * 128) + (ks & 3) * 32);
                    const uint64_t dh = tc::make_smem_desc(bhi + o, 16, 1024) | SW128;
                    const uint64_t dl = tc::make_smem_desc(blo + o, 16, 1024) | SW128;
                    tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dh, idesc, !(i % TN_GROUP == 0 && ks == 0));
                    tc::mma_tf32_ts(d_acc, a_lo + ks * 8, dh, idesc, true);
                    tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dl, idesc, true);
                }
                tc::mma_commit(&bars[set]);
            }
        }
        if (nit > 0) {
            tc::mbar_wait(&bars[set], (nit - 1) & 1);
            take();
        }
    }
    // set 1 hands its sums to set 0 through shared memory (every MMA has retired: the X tiles are free)
    tc::fence_before_sync();
    __syncthreads();
    const int row = wq * 32 + lane;
    if (warp >= 4 && warp < 8) {
#pragma unroll
        for (int j = 0; j < TBN; ++j) red[j * 128 + row] = acc[j];
    }
    __syncthreads();
    if (warp < 4 && mvalid) {
        float* dst = g.out + (size_t)blockIdx.y * g.N * g.M + m;
#pragma unroll
        for (int j = 0; j < TBN; ++j)
            if (n0 + j < g.N) dst[(size_t)(n0 + j) * g.M] = acc[j] + red[j * 128 + row];
    }
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

__global__ void k_sum_slices(const float* __restrict__ part, int S, long long n, float* __restrict__ D)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = part[i];
    for (int k = 1; k < S; ++k) s += part[(long long)k * n + i];
    D[i] = s;
}

// K splits: fill the SMs in whole waves, at least 8 chunks per split
int tn_splits(int tiles, int nchunks)
{
    const int sms = qc_num_sms();
    int best = 1;
    double best_eff = 0.0;
    for (int S = 1; S <= 160 && (S == 1 || nchunks / S >= 8); ++S) {
        const int ctas = tiles * S;
        const double eff = (double)ctas / ((double)((ctas + sms - 1) / sms) * sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = S; }
    }
    return best;
}

}  // namespace

extern "C" size_t qc_gemm_tn_workspace_bytes(int nb, int Kb, int M, int N)
{
    if (nb <= 0 || Kb <= 0 || M <= 0 || N <= 0) return 0;
    const int cpb = (Kb + TK - 1) / TK;
    const int tiles = ((M + 127) / 128) * ((N + TBN - 1) / TBN);
    const int S = tn_splits(tiles, nb * cpb);
    return S > 1 ? (size_t)S * N * M * sizeof(float) : 0;
}

extern "C" int qc_gemm_tn_tf32x3(const float* T, const float* X, float* D, int nb, int Kb, int M, int N,
                                  void* ws, size_t ws_bytes, void* stream)
{
    if (nb <= 0 || Kb <= 0 || M <= 0 || N <= 0) return QC_EINVAL;
    const int cpb = (Kb + TK - 1) / TK;
    const long long nch = (long long)nb * cpb;
    const int mt = (M + 127) / 128, nt = (N + TBN - 1) / TBN;
    if (nch > 2147483647LL || nt > 65535) return QC_EUNSUP;
    const int S = tn_splits(mt * nt, (int)nch);
    if (S > 1 && (ws == nullptr || ws_bytes < (size_t)S * N * M * sizeof(float))) return QC_EINVAL;
    TnArgs g{T, X, S > 1 ? (float*)ws : D, nb, Kb, M, N, cpb, (int)nch, (int)((nch + S - 1) / S)};
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)2 * 2 * 2 * XS_TILE * 4 + 1024;
    QC_CUDA((qc_set_max_smem<k_gemm_tn_tf32x3>((int)smem)));
    CUtensorMap tmap;
    if (getenv("QC_GEMM_NO_TMA") == nullptr && (M & 3) == 0 && (reinterpret_cast<uintptr_t>(T) & 15) == 0 &&
        (long long)nb * Kb <= 2147483647LL && tma::make_map_2d(&tmap, T, 4, (uint64_t)nb * Kb, (uint64_t)M, 32, TK)) {
        const size_t smem2 = (size_t)TST * TK * 128 * 4 + (size_t)2 * 2 * 2 * XS_TILE * 4 + 1024;
        const bool x16 = (Kb & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
        if (x16) {
            QC_CUDA((qc_set_max_smem<k_gemm_tn_tf32x3_tma<true>>((int)smem2)));
            k_gemm_tn_tf32x3_tma<true><<<dim3((unsigned)mt, (unsigned)S, (unsigned)nt), TN2_THREADS, smem2, st>>>(g, tmap);
        } else {
            QC_CUDA((qc_set_max_smem<k_gemm_tn_tf32x3_tma<false>>((int)smem2)));
            k_gemm_tn_tf32x3_tma<false><<<dim3((unsigned)mt, (unsigned)S, (unsigned)nt), TN2_THREADS, smem2, st>>>(g, tmap);
        }
    } else {
        k_gemm_tn_tf32x3<<<dim3((unsigned)mt, (unsigned)S, (unsigned)nt), TN_THREADS, smem, st>>>(g);
    }
    QC_CHECK_LAUNCH();
    if (S > 1) {
        const long long n = (long long)N * M;
        k_sum_slices<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)ws, S, n, D);
        QC_CHECK_LAUNCH();
    }
    return 0;
}
