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
// One CTA = 8 warps = a 128-row x BN-column tile with the accumulator in TMEM.  Per 32-wide K chunk
// the warps fetch the A (and B) rows with 128-bit loads three chunks ahead (register ring), split them
// into tf32 hi / lo and write them as 16-byte pieces into the canonical K-major no-swizzle operand tiles
// of a two-stage shared-memory ring; one thread issues the 12 MMAs of the chunk and commits them to the
// stage's mbarrier, which the loaders wait on before they overwrite that stage.
//
// tcgen05.mma accumulates with truncation (~2e-8 relative per MMA, measured: 1.5e-5 after K = 2048), so the
// TMEM accumulation only runs over groups of GROUP chunks (48 MMAs); two TMEM accumulators alternate, and
// while one group is multiplied the previous group's result is added to fp32 registers (round to nearest).
#include <cstdint>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int GK = 32;        // K chunk
constexpr int GSTAGES = 2;
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
constexpr int PD = 3;         // K chunks fetched ahead (register ring)

template <int BN>
__global__ void __launch_bounds__(GTHREADS, 1)
k_gemm_tf32x3(const GemmArgs g)
{
    constexpr int A_TILE = 128 * GK;             // floats per A tile (hi or lo)
    constexpr int B_TILE = BN * GK;
    constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    constexpr int HN = BN / 2;                   // output columns per thread (two threads share a row)
    static_assert(BN == 64 || BN == 128, "tile widths");
    extern __shared__ __align__(1024) float smem_raw[];
    // operand tiles use the 128-byte swizzle: 1 KB alignment
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[GSTAGES];
    __shared__ __align__(8) uint64_t gbars[2];     // one per TMEM accumulator: its group of MMAs has finished
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wq = warp & 3, wh = warp >> 2;     // TMEM lane quadrant (rows wq*32..), column / K half
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * BN;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 2 * BN);
    if (tid == 0) {
        for (int i = 0; i < GSTAGES; ++i) tc::mbar_init(&bars[i], 1);
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

    // Loader mapping: the 8 warps move 128 rows x 128 bytes per chunk with four 128-bit loads per lane.  Warp
    // (wq, wh) covers rows wq*32.. and the K half wh; in load j the lane handles row j*8 + lane%8 of those 32 and
    // the 16-byte piece q = wh*4 + lane/8: eight lanes share a 64-byte run of one row, and the eight lanes of a
    // quarter-warp write one contiguous 128-byte core matrix (no bank conflicts).  PD chunks are in flight.
    float4 ra[PD][4], rb[PD][4];
    auto fetch = [&](float4 (&xa)[4], float4 (&xb)[4], int kc) {
        const int k0 = kc * GK;
        const int q = wh * 4 + (lane >> 3);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rl = wq * 32 + j * 8 + (lane & 7);
            xa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            xb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + 4 * q < g.K) {   // K is a multiple of 4 (checked by the host)
                if (m0 + rl < g.M) xa[j] = __ldg(reinterpret_cast<const float4*>(g.A + (m0 + rl) * (long long)g.K + k0) + q);
                if (rl < BN && n0 + rl < g.N)
                    xb[j] = __ldg(reinterpret_cast<const float4*>(g.B + (long long)(n0 + rl) * g.K + k0) + q);
            }
        }
    };
    auto split_store = [&](float* hi_t, float* lo_t, const float4 (&r)[4]) {
        const int q = wh * 4 + (lane >> 3);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rl = wq * 32 + j * 8 + (lane & 7);
            const float v[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                h[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
                l[e] = v[e] - h[e];
            }
            // K-major, SWIZZLE_128B: a row is one 128-byte line (GK = 32 floats), 8 rows form a 1 KB atom and
            // the 16-byte piece index is XORed with the row index inside the atom (conflict-free operand fetch)
            const uint32_t off = (uint32_t)((rl >> 3) * 1024 + (rl & 7) * 128 + ((q ^ (rl & 7)) * 16)) / 4;
            *reinterpret_cast<float4*>(hi_t + off) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(lo_t + off) = make_float4(l[0], l[1], l[2], l[3]);
        }
    };

    const uint32_t idesc = tc::make_idesc_tf32(128, BN);
    constexpr uint32_t SBO = 1024;                      // between 8-row atoms
    constexpr uint64_t SW128 = (uint64_t)2 << 61;       // descriptor layout type SWIZZLE_128B
    uint32_t phase0 = 0u, phase1 = 0u;
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
    for (int u = 0; u < PD; ++u)
        if (u < nk) fetch(ra[u], rb[u], u);
    for (int kc0 = 0; kc0 < nk; kc0 += PD) {
#pragma unroll
        for (int u = 0; u < PD; ++u) {
            const int kc = kc0 + u;
            if (kc >= nk) break;                 // block-uniform
            const int s = kc % GSTAGES;
            float* st = smem + (size_t)s * STAGE;
            if (kc >= GSTAGES) {   // the MMAs that read this stage two chunks ago have finished
                if (s == 0) { tc::mbar_wait(&bars[0], phase0 & 1); ++phase0; }
                else        { tc::mbar_wait(&bars[1], phase1 & 1); ++phase1; }
            }
            split_store(st, st + A_TILE, ra[u]);
            if (wq * 32 < BN) split_store(st + 2 * A_TILE, st + 2 * A_TILE + B_TILE, rb[u]);
            if (kc + PD < nk) fetch(ra[u], rb[u], kc + PD);     // in flight while the next chunks are multiplied
            const int gi = kc / GROUP;
            if (kc % GROUP == GROUP / 2 && gi >= 1) {   // the previous group finished long ago: no stall
                take_group(gi - 1);
                gread = gi;
            }
            tc::fence_proxy_async();
            tc::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t ahi = tc::smem_u32(st), alo = ahi + A_TILE * 4;
                const uint32_t bhi = ahi + 2 * A_TILE * 4, blo = bhi + B_TILE * 4;
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
                tc::mma_commit(&bars[s]);
                if (kc % GROUP == GROUP - 1 || kc == nk - 1) tc::mma_commit(&gbars[gi & 1]);
            }
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
            for (int q = 0; q < HN / 4; ++q) {
                if ((g.N & 3) == 0 && nb + 4 * q + 4 <= g.N) {
                    reinterpret_cast<float4*>(dst)[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (nb + 4 * q + e < g.N) dst[4 * q + e] = acc[4 * q + e];
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
    const size_t smem = (size_t)GSTAGES * (2 * 128 * GK + 2 * BN * GK) * 4 + 1024;
    QC_CUDA(cudaFuncSetAttribute(k_gemm_tf32x3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
