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
