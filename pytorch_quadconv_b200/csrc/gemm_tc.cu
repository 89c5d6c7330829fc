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
// (the fourth GEMM of the path, dW = X * T', reduces over the ROW index of T' and stays on cuBLAS.)
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
    for (int S = 1; S <= 64 && (S == 1 || nchunks / S >= 8); ++S) {
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
    k_gemm_tn_tf32x3<<<dim3((unsigned)mt, (unsigned)S, (unsigned)nt), TN_THREADS, smem, st>>>(g);
    QC_CHECK_LAUNCH();
    if (S > 1) {
        const long long n = (long long)N * M;
        k_sum_slices<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)ws, S, n, D);
        QC_CHECK_LAUNCH();
    }
    return 0;
}
