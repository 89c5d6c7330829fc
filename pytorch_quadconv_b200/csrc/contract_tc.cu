// contract_tc.cu -- tensor-core (tcgen05) version of the fused gather-contract kernel.
//
// Stage B (segmented gather-accumulate over the pair list) is irregular, per-pair-unique work and
// stays on the FP32 pipe: gathers by cp.async into a per-warp shared-memory ring (FIFO commit groups), the
// staged pair n+1 moved to registers under the packed FMAs (fma.rn.f32x2) of pair n, all per-point control
// flow warp-uniform.  Stage C -- the last Linear of the
// reference's filter MLP (quadconv.py:143) folded into the contraction,
//     Y[row][j] = sum_k T[row][k] * W2[k][j],      row = (point, batch lane), k = (channel, h)
// -- is a dense GEMM with a shared weight, and runs on the 5th-generation tensor cores:
//   * each thread owns one row; after a channel chunk (K = 64) it splits its 64 fp32 partial sums
//     into tf32 hi + lo parts and writes them straight from registers into TENSOR MEMORY with
//     tcgen05.st, where they are the A operand (lane = row, column = k) -- no shared-memory A tile;
//   * the repacked weight slices (hi and lo) sit in shared memory for the whole kernel in the
//     canonical K-major no-swizzle layout and are the B operand through matrix descriptors;
//   * one elected thread per 128-row block issues 3 x 8 tcgen05.mma.kind::tf32 (M=128, N=32, K=8):
//     A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  (3xTF32 error compensation: fp32-grade accuracy, needed for
//     the 1e-5 bound); the accumulator D lives in TMEM across all channel chunks of a tile;
//   * tcgen05.commit -> mbarrier tells the row block when A may be overwritten / D may be read;
//     the epilogue reads D with tcgen05.ld and stores the packed output rows with 128-bit stores.
// The MMAs are asynchronous: the FP32 pipe is already gathering the next chunk while they run.
// tcgen05.mma accumulates with truncation (~2e-8 relative per MMA): D only accumulates the chunks of one tile
// (96 MMAs at 32 channels), and the dW_last accumulators of the DW variant, which would otherwise run for the
// whole kernel, are flushed every DW_FLUSH uses (see flush_dw / k_dw_reduce).
#include <cstdlib>
#include "contract_args.cuh"
#include "tc_common.cuh"

namespace {

// dW_last operand tiles (DW variant): K-major, K = the 128 rows of a row block.  A 144-byte step between
// the 4-row core matrices makes the per-row scalar stores of a warp hit 32 distinct banks.
constexpr int OP_LBO = 144;                 // bytes between core matrices adjacent in K (4 rows each)
constexpr int OP_SBO = 32 * OP_LBO;         // bytes between 8-element groups in M / N
constexpr int OP_A_BYTES = 8 * OP_SBO;      // T^T tile: 64 m  x 128 rows
constexpr int OP_B_BYTES = 4 * OP_SBO;      // Z tile:   32 j  x 128 rows
constexpr int OPBUF_BYTES = 2 * OP_A_BYTES + 2 * OP_B_BYTES;
__host__ __device__ constexpr int op_off_words(int mn, int r) { return ((mn >> 3) * OP_SBO + (r >> 2) * OP_LBO + (mn & 7) * 16 + (r & 3) * 4) / 4; }
constexpr int PCH = 64;  // pairs staged per chunk

// Ring depth (gathered pairs in flight per warp) of the two kernel families; NG = 16-byte groups per pair.
// Measured at C3: depth 4 beats 8 by 3-4 % (shorter unrolled body; the shared memory it frees, i.e. the larger
// L1, made no measurable difference on its own), depth 2 loses the overlap (+45 %).
__host__ __device__ constexpr int ctc_ring_depth(bool dw, int kch, int ng) { return (dw ? 8 : (kch == 32 ? 6 : 8)) / ng; }
__host__ __device__ constexpr int dtc_ring_depth(int ng) { return 8 / ng; }
// uses of the dW operand buffer between two flushes of the TMEM dW accumulators (QC_DW_FLUSH overrides): 16 uses = 192
// accumulating MMAs per chain, measured 4.2e-6 on dW_last at C3 (8 uses: 2.4e-6) against the 1e-5 bound
inline int dw_flush_interval()
{
    static const int v = []() { const char* e = getenv("QC_DW_FLUSH"); const int x = e ? atoi(e) : 16; return x < 1 ? 1 : x; }();
    return v;
}

// KCH = K of one channel chunk = channels per chunk x HP.  KCH = 64 runs one CTA per SM (TMEM: 2 row
// blocks x (64 hi + 64 lo + 32 D) = 320 of 512 columns); KCH = 32 halves the register tile and the
// TMEM footprint (2 x 96 = 192 -> a 256-column allocation) so that TWO CTAs share an SM: 16 warps
// hide the latencies that 8 cannot.  The DW variant keeps KCH = 64 (its mini-GEMM tiles assume it).
// NW = warps per CTA: 8 (two 128-row blocks) or 12 (three; 3 x 160 = 480 TMEM columns, 170 registers per
// thread) -- more resident warps per scheduler to cover the FP32 pipe's dependency stalls.
template <int HP, int KCH, int LB, bool DW, int NW>
__global__ void __launch_bounds__(NW * 32, KCH == 32 ? 2 : 1)
k_contract_tc(const ContractArgs a, const int nkc, const int dw_nslots, const int DW_FLUSH)
{
    static_assert(!DW || KCH == 64, "the dW mini-GEMM assumes 64-wide chunks");
    static_assert(!DW || NW == 8, "the dW mini-GEMM assumes 256 threads");
    static_assert(NW == 8 || (NW == 12 && KCH == 64), "supported CTA shapes");
    constexpr int NT = NW * 32;
    constexpr int CXC = KCH / HP, Bp = 1 << LB, PPW = 32 >> LB;
    constexpr int BSLICE = 32 * KCH;            // floats per weight slice (N = 32 rows x K = KCH), hi or lo
    constexpr int TCOLS = KCH == 32 ? 256 : 512;   // TMEM columns allocated by this CTA
    constexpr int RBCOLS = 2 * KCH + 32;           // per row block: A_hi, A_lo, D
    constexpr int NG = CXC / 4, HQ = HP / 4;
    constexpr int D = ctc_ring_depth(DW, KCH, NG);   // gathered pairs in flight per warp (async smem ring)
    constexpr int RING = D * NG * 512;          // bytes per warp: slot = NG groups x 32 lanes x 16 B
    // pairs staged per piece: the dW variant with 8 batch lanes holds 4 points per warp next to the 108 KB
    // operand buffer, so its pieces are shorter
    constexpr int PCHK = (DW && LB == 3) ? 16 : PCH;
    constexpr int PHS = PCHK * HP + 4;
    extern __shared__ __align__(128) float smem[];
    float* Bs = smem;                                              // [nkc][2][BSLICE] (hi, lo), canonical K-major
    float* phi_s = Bs + (size_t)nkc * 2 * BSLICE;                  // [NW*PPW][PHS]
    int* idx_s = reinterpret_cast<int*>(phi_s + NW * PPW * PHS);   // [NW*PPW][PCH]
    uint64_t* bars = reinterpret_cast<uint64_t*>(idx_s + NW * PPW * PCHK);  // [3] mbarriers + tmem base
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 3);
    float* ring_s = reinterpret_cast<float*>(bars + 4);            // [NW warps][RING bytes]
    // DW: operand tiles of the dW_last GEMM, K-major with a 144-byte core-matrix step along K (= rows):
    // T^T hi, T^T lo  [64 m][128 rows]  and  Z hi, Z lo  [32 j][128 rows]; one buffer, used by the two
    // row blocks in turn under a lock.
    float* opbuf = ring_s + NW * RING / 4;
    uint64_t* opbar = reinterpret_cast<uint64_t*>(opbuf + OPBUF_BYTES / 4);
    int* lockw = reinterpret_cast<int*>(opbar + 1);                // [0] lock, [1] commits so far, [2+kc] touched

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // broadcast: the compiler then knows it is warp-uniform
    const int rb = warp >> 2;                                      // 128-row block of this warp
    const int s = lane >> LB, b = lane & (Bp - 1);
    const int blk = blockIdx.y;
    const unsigned xrow = (unsigned)qc_cp4(a.Cx) * Bp;
    const size_t yrow = (size_t)qc_cp4(a.Cy) * Bp;
    const float* Xb = a.X + (size_t)blk * a.Nsrc * xrow + b * 4;
    constexpr int ppt = NW * PPW;
    const int ntiles = (a.Ndst + ppt - 1) / ppt;

    // ---- one-time setup: TMEM, mbarriers, weight slices -------------------------------------
    if (warp == 0) tc::tmem_alloc(tmem_base_s, TCOLS);
    if (tid == 32) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_init(&bars[2], 1);
        if (DW) {
            tc::mbar_init(opbar, 1);
            for (int i = 0; i < 16; ++i) lockw[i] = 0;   // [0] lock, [1] commits, [2+kc] accumulator kc touched
        }
        tc::mbar_fence_init();
    }
    for (int idx = tid; idx < nkc * BSLICE; idx += NT) {
        const int kc = idx / BSLICE, r = idx % BSLICE;
        const int n = r / KCH, k = r % KCH;                        // n = output channel j, k = (i, h)
        const int i = k / HP, h = k % HP;
        const int cx = kc * CXC + i;
        float v = 0.f;
        if (cx < a.Cx && h < a.H && n < a.Cy) v = __ldg(a.W2 + ((size_t)cx * a.H + h) * a.CyP + n);
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const uint32_t off = tc::kmajor_off_bytes(n, k, KCH) / 4;
        Bs[(size_t)(kc * 2 + 0) * BSLICE + off] = hi;
        Bs[(size_t)(kc * 2 + 1) * BSLICE + off] = v - hi;
    }
    for (int i = tid; i < NW * RING / 4; i += NT) ring_s[i] = 0.f;   // stale ring slots are read (and must be finite)
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_base_s;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t a_hi = tmem + rb * RBCOLS, a_lo = a_hi + KCH, d_acc = a_hi + 2 * KCH;
    const uint32_t idesc = tc::make_idesc_tf32(128, 32);
    const uint32_t bs_addr = tc::smem_u32(Bs);
    uint32_t g = 0;  // MMA groups committed so far on this row block (same value in all its threads)

    // DW: the dW2 accumulators live in TMEM (lanes 0..63 = k index m, 32 columns = j, one block per chunk).
    // tcgen05.mma accumulates with truncation, so a long chain of accumulations loses ~2e-8 (relative) per
    // MMA (measured: 1.5e-4 on dW_last at C3 when the chain ran for the whole kernel).  Every DW_FLUSH uses
    // of the operand buffer (96 MMAs per accumulator, ~2e-6) the accumulators are therefore written out as
    // one more partial result (plain stores into this CTA's next slot of the scratch buffer) and restarted;
    // k_dw_reduce sums the slots in fp32 (round to nearest) into dW2.  Called by the warps that own TMEM
    // lanes 0..63 (warp & 3 < 2) while no dW MMA is in flight.
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    float* dw_slots = DW ? a.dW_scratch + (size_t)cta * dw_nslots * ((size_t)nkc * 64 * 32) : nullptr;
    auto flush_dw = [&](int slot) {
        const int m = (warp & 3) * 32 + lane;
        for (int kc = 0; kc < nkc; ++kc) {
            float4* row = reinterpret_cast<float4*>(dw_slots + ((size_t)slot * nkc + kc) * (64 * 32) + (size_t)m * 32);
            if (*reinterpret_cast<volatile int*>(&lockw[2 + kc]) == 0) {   // nothing accumulated since the last flush
#pragma unroll
                for (int q = 0; q < 8; ++q) row[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            uint32_t v0[16], v1[16];
            tc::tmem_ld16(lane_base + tmem + 2 * RBCOLS + kc * 32, v0);
            tc::tmem_ld16(lane_base + tmem + 2 * RBCOLS + kc * 32 + 16, v1);
            tc::wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                row[q] = make_float4(__uint_as_float(v0[4 * q]), __uint_as_float(v0[4 * q + 1]),
                                     __uint_as_float(v0[4 * q + 2]), __uint_as_float(v0[4 * q + 3]));
                row[4 + q] = make_float4(__uint_as_float(v1[4 * q]), __uint_as_float(v1[4 * q + 1]),
                                         __uint_as_float(v1[4 * q + 2]), __uint_as_float(v1[4 * q + 3]));
            }
        }
    };

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

        for (int kc = 0; kc < nkc; ++kc) {
            const int cx0 = kc * CXC;
            const int nval = min(CXC, a.Cx - cx0);
            const int ngv = (nval + 3) >> 2;

            // ---- stage B: segmented gather-accumulate on the FP32 pipe ---------------------------
            float T[CXC][HP];
#pragma unroll
            for (int i = 0; i < CXC; ++i)
#pragma unroll
                for (int h = 0; h < HP; ++h) T[i][h] = 0.f;

            const float* Xc = Xb + (size_t)(cx0 >> 2) * (Bp * 4);
            float* my_phi = phi_s + (warp * PPW + s) * PHS;
            int* my_idx = idx_s + (warp * PPW + s) * PCHK;
            const uint32_t my_idx_sa = tc::smem_u32(my_idx);
            // gathers go global -> shared asynchronously (cp.async, one commit group per pair) into a
            // per-warp ring; each lane copies and later reads back only its own 16 bytes per group
            const uint32_t my_ring = tc::smem_u32(ring_s) + warp * RING + lane * 16;
            const float* my_ring_p = ring_s + (warp * RING + lane * 16) / 4;
            auto issue_u = [&](int slot, unsigned u) {
                const float4* xr = reinterpret_cast<const float4*>(Xc + (size_t)(u * xrow));
#pragma unroll
                for (int q = 0; q < NG; ++q)
                    if (q < ngv) tc::cp_async16(my_ring + (slot * NG + q) * 512, xr + q * Bp);
            };
            auto issue = [&](int slot, int e) { issue_u(slot, (unsigned)my_idx[e]); };
            // registers of one staged pair: its gathered channels and its phi row
            struct Staged { float4 x[NG]; float4 ph[HQ]; };
            auto load = [&](Staged& r, int slot, int e) {   // shared memory -> registers (slot has landed)
                const float4* pr = reinterpret_cast<const float4*>(my_phi + e * HP);
#pragma unroll
                for (int q = 0; q < HQ; ++q) r.ph[q] = pr[q];
#pragma unroll
                for (int q = 0; q < NG; ++q)   // groups past ngv hold stale finite values; their W2 rows are zero
                    r.x[q] = *reinterpret_cast<const float4*>(my_ring_p + (slot * NG + q) * 128);
            };
            auto fma = [&](const Staged& r) {
                float ph[HP];
#pragma unroll
                for (int q = 0; q < HQ; ++q)
                    ph[4 * q] = r.ph[q].x, ph[4 * q + 1] = r.ph[q].y, ph[4 * q + 2] = r.ph[q].z, ph[4 * q + 3] = r.ph[q].w;
#pragma unroll
                for (int q = 0; q < NG; ++q) {
                    const float f[4] = {r.x[q].x, r.x[q].y, r.x[q].z, r.x[q].w};
#pragma unroll
                    for (int e2 = 0; e2 < 4; e2 += 2)
#pragma unroll
                        for (int h = 0; h < HP; ++h)
                            tc::fma2(T[4 * q + e2][h], T[4 * q + e2 + 1][h], f[e2], f[e2 + 1], ph[h], ph[h]);
                }
            };
            for (int c0 = 0; c0 < maxlen; c0 += PCHK) {
                const int clen = min(PCHK, len - c0);
                const int cmax = min(PCHK, maxlen - c0);
                // pair indices and phi rows of this piece of the segment; a segment that fits in one piece
                // (the usual case) is staged once per tile and reused by every channel chunk
                if (kc == 0 || maxlen > PCHK) {
                    __syncwarp();
                    for (int e = b; e < clen; e += Bp) {
                        const int p = c0 + e;
                        my_idx[e] = sidx[p];
                        const int pp = pidx ? pidx[p] : beg + p;
                        const float4* php = reinterpret_cast<const float4*>(a.phi + (size_t)pp * HP);
#pragma unroll
                        for (int q = 0; q < HQ; ++q) reinterpret_cast<float4*>(my_phi + e * HP)[q] = __ldg(php + q);
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    if (d < clen) issue(d, d);
                    tc::cp_async_commit();
                }
                // software pipeline: while pair n is multiplied, pair n+1 moves shared -> registers
                Staged cur, nxt;
                tc::cp_async_wait<D - 1>();
                load(cur, 0, 0);
                for (int it = 0; it < cmax; it += D) {
                    // the source indices of the D pairs issued in this round: ONE 128-bit load a whole round ahead of
                    // their use (the shared-memory pipe is ~50 % busy; a per-pair LDS.32 issued one FMA block ahead of
                    // its address computation was the top stall of the kernel)
                    unsigned unv[D];
                    if constexpr (D == 4) {
                        const uint4 q4 = tc::lds_u128_early(my_idx_sa + (it + D) * 4);      // past the list they are unused
                        unv[0] = q4.x, unv[1] = q4.y, unv[2] = q4.z, unv[3] = q4.w;
                    }
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        tc::cp_async_wait<D - 2>();       // the group of pair it+d+1 (slot d+1) has landed
                        load(nxt, (d + 1) % D, it + d + 1);
                        unsigned un;
                        if constexpr (D == 4) un = unv[d];
                        else un = tc::lds_u32_early(my_idx_sa + (it + d + D) * 4);   // past the list it is unused
                        if (it + d < clen) {
                            fma(cur);
                            if (it + d + D < clen) issue_u(d, un);
                        }
                        tc::cp_async_commit();
                        cur = nxt;
                    }
                }
            }

            // DW: this row of Z, fetched now so that its latency is not paid while holding the operand buffer
            float4 zrow[DW ? 8 : 1];
            if (DW) {
                const float* zg = a.Z + ((size_t)blk * a.Ndst + (pvalid ? dstp : 0)) * yrow + b * 4;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    zrow[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pvalid && 4 * q < a.Cy) zrow[q] = __ldg(reinterpret_cast<const float4*>(zg + (size_t)q * Bp * 4));
                }
            }
            // ---- stage C on the tensor cores: T (hi/lo) -> TMEM, 3xTF32 MMAs into D ---------------
            if (g > 0) {
                tc::mbar_wait(&bars[rb], (g - 1) & 1);   // previous group has finished reading A
                tc::fence_after_sync();
            }
#pragma unroll
            for (int c = 0; c < KCH / 16; ++c) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = T[(c * 16 + j) / HP][(c * 16 + j) % HP];
                    const float hv = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
                    hi[j] = __float_as_uint(hv);
                    lo[j] = __float_as_uint(v - hv);
                }
                tc::tmem_st16(lane_base + a_hi + c * 16, hi);
                tc::tmem_st16(lane_base + a_lo + c * 16, lo);
            }
            tc::wait_st();
            const bool leader = (warp & 3) == 0 && lane == 0;
            int commits = 0;
            if (DW && leader) {
                // take the operand buffer: lock, then wait until the previous user's MMAs have read it
                // dw_ordered: the two row blocks alternate strictly (use 2u of the buffer is block 0's u-th, use 2u + 1
                // block 1's -- both make one use per tile and chunk), which fixes the order of the TMEM accumulations
                // and makes d W_last bit-reproducible; otherwise first come, first served
                for (;;) {
                    if (a.dw_ordered && ((*reinterpret_cast<volatile int*>(&lockw[1]) & 1) != rb)) {
                        __nanosleep(32);
                        continue;
                    }
                    if (atomicCAS(&lockw[0], 0, 1) == 0) break;
                    __nanosleep(64);
                }
                __threadfence_block();
                commits = *reinterpret_cast<volatile int*>(&lockw[1]);
                if (commits > 0) tc::mbar_wait(opbar, (commits - 1) & 1);
            }
            tc::fence_before_sync();
            tc::named_bar_sync(1 + rb, 128);
            if (DW) {
                // (the leader holds the lock and has seen the previous user's MMAs complete)
                const int done = *reinterpret_cast<volatile int*>(&lockw[1]);
                if (done > 0 && done % DW_FLUSH == 0) {
                    tc::fence_after_sync();
                    if ((warp & 3) < 2) flush_dw(done / DW_FLUSH - 1);
                    tc::fence_before_sync();
                    tc::named_bar_sync(1 + rb, 128);
                    if (leader)
                        for (int k2 = 0; k2 < nkc; ++k2) lockw[2 + k2] = 0;   // the next MMA group overwrites
                }
            }
            if (DW) {
                // this thread's row of T (hi/lo) and of Z (hi/lo) -> K-major operand tiles, K = row
                const int r = (warp & 3) * 32 + lane;
                float* Ahi = opbuf;
                float* Alo = opbuf + OP_A_BYTES / 4;
                float* Bhi = opbuf + 2 * OP_A_BYTES / 4;
                float* Blo = Bhi + OP_B_BYTES / 4;
#pragma unroll
                for (int k = 0; k < 64; ++k) {
                    const float v = T[k / HP][k % HP];
                    const float hv = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
                    Ahi[op_off_words(k, r)] = hv;
                    Alo[op_off_words(k, r)] = v - hv;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float z4[4] = {zrow[q].x, zrow[q].y, zrow[q].z, zrow[q].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float hv = __uint_as_float(__float_as_uint(z4[e]) & 0xffffe000u);
                        Bhi[op_off_words(4 * q + e, r)] = hv;
                        Blo[op_off_words(4 * q + e, r)] = z4[e] - hv;
                    }
                }
                tc::fence_proxy_async();
                tc::named_bar_sync(1 + rb, 128);
            }
            if (leader) {
                tc::fence_after_sync();
                if (DW) {
                    // dW2 chunk kc: D_dw[m][j] += sum_r T[r][m] * Z[r][j]  (M = 128: rows 64.. are don't-care).  Issued
                    // BEFORE the stage-C MMAs of this chunk: the tensor pipe retires in order, and the next user of
                    // the operand buffer waits for exactly these MMAs.
                    const uint32_t op = tc::smem_u32(opbuf);
                    const uint32_t d_dw = tmem + 2 * RBCOLS + kc * 32;
                    const bool fresh = lockw[2 + kc] == 0;   // first MMA group of this CTA into accumulator kc (under the lock)
                    lockw[2 + kc] = 1;
#pragma unroll 4
                    for (int ks = 0; ks < 16; ++ks) {   // K = 128 rows, 8 per MMA = two 4-row core matrices
                        const uint64_t ahi = tc::make_smem_desc(op + ks * 2 * OP_LBO, OP_LBO, OP_SBO);
                        const uint64_t alo = tc::make_smem_desc(op + OP_A_BYTES + ks * 2 * OP_LBO, OP_LBO, OP_SBO);
                        const uint64_t zhi = tc::make_smem_desc(op + 2 * OP_A_BYTES + ks * 2 * OP_LBO, OP_LBO, OP_SBO);
                        const uint64_t zlo = tc::make_smem_desc(op + 2 * OP_A_BYTES + OP_B_BYTES + ks * 2 * OP_LBO, OP_LBO, OP_SBO);
                        tc::mma_tf32_ss(d_dw, ahi, zhi, idesc, !(fresh && ks == 0));
                        tc::mma_tf32_ss(d_dw, alo, zhi, idesc, true);
                        tc::mma_tf32_ss(d_dw, ahi, zlo, idesc, true);
                    }
                    tc::mma_commit(opbar);
                    *reinterpret_cast<volatile int*>(&lockw[1]) = commits + 1;
                    __threadfence_block();
                    atomicExch(&lockw[0], 0);
                }
                const uint32_t bhi = bs_addr + (uint32_t)(kc * 2) * BSLICE * 4, blo = bhi + BSLICE * 4;
                constexpr uint32_t SBO = (KCH / 4) * 128;
#pragma unroll
                for (int ks = 0; ks < KCH / 8; ++ks) {
                    const uint64_t dhi = tc::make_smem_desc(bhi + ks * 256, 128, SBO);
                    const uint64_t dlo = tc::make_smem_desc(blo + ks * 256, 128, SBO);
                    tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dhi, idesc, !(kc == 0 && ks == 0));
                    tc::mma_tf32_ts(d_acc, a_lo + ks * 8, dhi, idesc, true);
                    tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dlo, idesc, true);
                }
                tc::mma_commit(&bars[rb]);
            }
            ++g;
        }

        // ---- epilogue: D (TMEM) -> registers -> packed output rows ------------------------------
        tc::mbar_wait(&bars[rb], (g - 1) & 1);
        tc::fence_after_sync();
        uint32_t r0[16], r1[16];
        tc::tmem_ld16(lane_base + d_acc, r0);
        tc::tmem_ld16(lane_base + d_acc + 16, r1);
        tc::wait_ld();
        if (pvalid) {
            float* y = a.Y + ((size_t)blk * a.Ndst + dstp) * yrow + b * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * q < a.Cy)
                    *reinterpret_cast<float4*>(y + (size_t)q * Bp * 4) =
                        make_float4(__uint_as_float(r0[4 * q]), __uint_as_float(r0[4 * q + 1]),
                                    __uint_as_float(r0[4 * q + 2]), __uint_as_float(r0[4 * q + 3]));
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (16 + 4 * q < a.Cy)
                    *reinterpret_cast<float4*>(y + (size_t)(4 + q) * Bp * 4) =
                        make_float4(__uint_as_float(r1[4 * q]), __uint_as_float(r1[4 * q + 1]),
                                    __uint_as_float(r1[4 * q + 2]), __uint_as_float(r1[4 * q + 3]));
        }
    }

    if (DW) {
        // dW2 accumulators (TMEM lanes 0..63 = k index m, 32 columns = j, one block per chunk) -> global
        __syncthreads();
        const int commits = *reinterpret_cast<volatile int*>(&lockw[1]);
        if (commits > 0) tc::mbar_wait(opbar, (commits - 1) & 1);
        tc::fence_after_sync();
        // the remainder goes to the next slot; the number of slots this CTA wrote is recorded for k_dw_reduce
        const int nfl = commits / DW_FLUSH;            // periodic flushes done (the one at `commits` itself has not run)
        const int used = commits > 0 ? (commits % DW_FLUSH == 0 ? nfl - 1 : nfl) : 0;   // slots written so far
        if (warp < 2 && commits > 0) flush_dw(used);
        if (tid == 0) a.dW_counts[cta] = commits > 0 ? used + 1 : 0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

// dW2[(c*H+h)][j] += sum over the slots written by every CTA, in two passes with FIXED summation orders (the result
// is bit-reproducible run to run): k_dw_reduce folds a CTA's slots into its slot 0 (grid (nkc*64*32 / 256, CTAs of the
// main kernel); entry e = (chunk, m, j)), k_dw_sum_ctas adds the CTAs' slot 0 in CTA order.
template <int HP>
__global__ void __launch_bounds__(256)
k_dw_reduce(float* __restrict__ scratch, const int* __restrict__ counts, int nslots, int nkc)
{
    const int e = blockIdx.x * 256 + threadIdx.x, cta = blockIdx.y;
    const int per = nkc * 64 * 32;
    if (e >= per) return;
    const int n = counts[cta];
    float* p = scratch + (size_t)cta * nslots * per + e;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int f = 0;
    for (; f + 4 <= n; f += 4) {
        s0 += p[(size_t)f * per];
        s1 += p[(size_t)(f + 1) * per];
        s2 += p[(size_t)(f + 2) * per];
        s3 += p[(size_t)(f + 3) * per];
    }
    for (; f < n; ++f) s0 += p[(size_t)f * per];
    p[0] = (s0 + s1) + (s2 + s3);            // only this thread touches element e of this CTA's slots
}

template <int HP>
__global__ void __launch_bounds__(256)
k_dw_sum_ctas(const float* __restrict__ scratch, int nctas, int nslots, int nkc, int Cx, int H, int Cy, int CyP,
              float* __restrict__ dW2)
{
    constexpr int CXC = 64 / HP;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int per = nkc * 64 * 32;
    if (e >= per) return;
    const int j = e & 31, m = (e >> 5) & 63, kc = e >> 11;
    const int cx = kc * CXC + m / HP, hh = m % HP;
    if (!(cx < Cx && hh < H && j < Cy)) return;
    const size_t cstride = (size_t)nslots * per;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 4 <= nctas; c += 4) {
        s0 += scratch[(size_t)c * cstride + e];
        s1 += scratch[(size_t)(c + 1) * cstride + e];
        s2 += scratch[(size_t)(c + 2) * cstride + e];
        s3 += scratch[(size_t)(c + 3) * cstride + e];
    }
    for (; c < nctas; ++c) s0 += scratch[(size_t)c * cstride + e];
    dW2[((size_t)cx * H + hh) * CyP + j] += (s0 + s1) + (s2 + s3);
}

template <int HP, int KCH, int LB, bool DW, int NW>
int launch_tc(const ContractArgs& a, int B, cudaStream_t st)
{
    constexpr int CXC = KCH / HP, PPW = 32 >> LB;
    constexpr int BSLICE = 32 * KCH;
    const int nkc = qc_ceil_div(a.Cx, CXC);
    const int nblk = qc_nblk(B);
    constexpr int NG = CXC / 4;
    constexpr int RING = ctc_ring_depth(DW, KCH, NG) * NG * 512;
    constexpr int PCHK = (DW && LB == 3) ? 16 : PCH;
    size_t smem = ((size_t)nkc * 2 * BSLICE + NW * PPW * (PCHK * HP + 4) + NW * PPW * PCHK) * 4 + 32 + NW * RING;
    if (DW) {
        smem += OPBUF_BYTES + 128;
        if (2 * (2 * KCH + 32) + nkc * 32 > 512 || nkc > 12) return 0;   // TMEM: stage C blocks + one dW accumulator per chunk
    }
    constexpr int CTAS = KCH == 32 ? 2 : 1;
    if (CTAS == 2) {
        // exactly two CTAs per SM: each allocates 256 of the 512 TMEM columns, a third resident CTA
        // would spin in tcgen05.alloc forever -> keep the footprint above a third of the SM's smem
        if (smem < 80 * 1024) smem = 80 * 1024;
        if (smem > 113 * 1024) return 0;
    }
    if (smem > 227 * 1024) return 0;   // weight slices do not fit: use the FP32-pipe kernel
    QC_CUDA((qc_set_max_smem<k_contract_tc<HP, KCH, LB, DW, NW>>((int)smem)));
    const int ntiles = qc_ceil_div(a.Ndst, NW * PPW);
    int gx = qc_num_sms() * CTAS;
    gx = (gx + nblk - 1) / nblk;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, nblk);
    if (DW) {
        // slots for the periodic flushes of the TMEM dW accumulators (see flush_dw): stream-ordered scratch
        ContractArgs a2 = a;
        const int tiles_per_cta = qc_ceil_div(ntiles, gx);
        const int max_commits = tiles_per_cta * (NW / 4) * nkc;
        const int DW_FLUSH = dw_flush_interval();
        const int nslots = (max_commits - 1) / DW_FLUSH + 1;
        const int nctas = gx * nblk, per = nkc * 64 * 32;
        const size_t fbytes = (size_t)nctas * nslots * per * sizeof(float);
        char* buf = reinterpret_cast<char*>(a.ws);
        const bool own = buf == nullptr || a.ws_bytes < fbytes + (size_t)nctas * sizeof(int);
        if (own) QC_CUDA(cudaMallocAsync((void**)&buf, fbytes + (size_t)nctas * sizeof(int), st));
        a2.dW_scratch = reinterpret_cast<float*>(buf);
        a2.dW_counts = reinterpret_cast<int*>(buf + fbytes);
        a2.dw_ordered = getenv("QC_DW_FREE_ORDER") == nullptr ? 1 : 0;
        k_contract_tc<HP, KCH, LB, DW, NW><<<grid, NW * 32, smem, st>>>(a2, nkc, nslots, DW_FLUSH);
        cudaError_t le = cudaGetLastError();
        if (le == cudaSuccess) {
            k_dw_reduce<HP><<<dim3(qc_ceil_div(per, 256), nctas), 256, 0, st>>>(a2.dW_scratch, a2.dW_counts, nslots, nkc);
            le = cudaGetLastError();
        }
        if (le == cudaSuccess) {
            k_dw_sum_ctas<HP><<<qc_ceil_div(per, 256), 256, 0, st>>>(a2.dW_scratch, nctas, nslots, nkc, a.Cx, a.H, a.Cy, a.CyP, a.dW2);
            le = cudaGetLastError();
        }
        if (own) cudaFreeAsync(buf, st);
        if (le != cudaSuccess) return (int)le;
        return 1;
    }
    k_contract_tc<HP, KCH, LB, DW, NW><<<grid, NW * 32, smem, st>>>(a, nkc, 0, 1);
    QC_CHECK_LAUNCH();
    return 1;
}


// ---------------------------------------------------------------------------------------------
// d(phi) kernel with the dense stage  dT[row][(c,h)] = sum_j G[row][j] * W3[j][c][h]  on tcgen05:
// the thread's grad_out row (hi/lo) is the A operand in TMEM (written once per tile), the W3 slices
// of all channel chunks are the B operand in shared memory; as soon as every thread of a row block has
// pulled D (64 columns) of chunk kc into registers the MMAs of chunk kc+1 are issued into the same
// TMEM columns and run while the FP32 pipe walks the pair list of chunk kc.  NW = 8 or 16 warps
// (measured at C3: 6.2 ms with 8 warps/SM, 4.8 with 12, 4.2 with 16).
constexpr int WSLICE = 64 * 32;  // floats per W3 slice: N = 64 rows (c,h) x K = 32 (j), hi or lo

template <int HP, int LB, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_dphi_tc(const DphiArgs a, const int nkc, const size_t slice_stride)
{
    constexpr int NT = NW * 32;
    constexpr int CXC = 64 / HP, Bp = 1 << LB, PPW = 32 >> LB;
    constexpr int NG = CXC / 4;
    constexpr int D = dtc_ring_depth(NG);       // gathered pairs in flight per warp (async smem ring); the unrolled
                                                // loop body must stay inside the 32 KB instruction cache
    constexpr int RING = D * NG * 512;
    constexpr int LH = HP == 8 ? 3 : 2;
    // With 32 batch lanes two pairs are reduced together: lane bit 4 picks the pair a half-warp ends up
    // with, the low LBr = 4 bits reduce that pair (16 shuffles per two pairs instead of 2 x 9, and half
    // the dependent stages per pair).
    constexpr bool TWO = LB == 5;
    constexpr int LBr = TWO ? 4 : LB;
    constexpr int NH = LBr < LH ? LBr : LH;
    constexpr int NF = HP >> NH;
    static_assert(NF == 1, "the butterfly ends with one value per lane");
    static_assert(D % 2 == 0, "two-pair steps walk the ring two slots at a time");
    extern __shared__ __align__(128) float smem[];
    float* Ws = smem;                                              // [nkc][2][WSLICE]
    int* idx_s = reinterpret_cast<int*>(Ws + (size_t)nkc * 2 * WSLICE);   // [NW*PPW][PCH]
    uint64_t* bars = reinterpret_cast<uint64_t*>(idx_s + NW * PPW * PCH); // one mbarrier per row block
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 4);
    float* ring_s = reinterpret_cast<float*>(bars + 8);                   // [NW warps][RING bytes]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // broadcast: the compiler then knows it is warp-uniform
    const int rb = warp >> 2;
    const int s = lane >> LB, b = lane & (Bp - 1);
    const int blk = blockIdx.y;
    const unsigned xrow = (unsigned)qc_cp4(a.Cx) * Bp;
    const size_t grow = (size_t)qc_cp4(a.Cg) * Bp;
    const float* Xb = a.X + (size_t)blk * a.Nsrc * xrow + b * 4;
    constexpr int ppt = NW * PPW;
    const int ntiles = (a.Ndst + ppt - 1) / ppt;
    const bool writer = (lane & ((1 << (LBr - NH)) - 1)) == 0;
    const int hsel = (lane >> (LBr - NH)) & ((1 << NH) - 1);   // the h this lane ends up holding after the butterfly
    const int up = TWO ? (lane >> 4) & 1 : 0;                  // which pair of a two-pair step this half-warp keeps

    if (warp == 0) tc::tmem_alloc(tmem_base_s, 512);
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
        tc::mbar_fence_init();
    }
    for (int idx = tid; idx < nkc * WSLICE; idx += NT) {
        const int kc = idx / WSLICE, r = idx % WSLICE;
        const int n = r >> 5, k = r & 31;                          // n = (i,h) of the chunk, k = j
        const int i = n / HP, h = n % HP;
        const int cx = kc * CXC + i;
        float v = 0.f;
        if (cx < a.Cx && k < a.Cg) v = __ldg(a.W3 + ((size_t)k * a.Cx + cx) * HP + h);
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const uint32_t off = tc::kmajor_off_bytes(n, k, 32) / 4;
        Ws[(size_t)(kc * 2 + 0) * WSLICE + off] = hi;
        Ws[(size_t)(kc * 2 + 1) * WSLICE + off] = v - hi;
    }
    for (int i = tid; i < NW * RING / 4; i += NT) ring_s[i] = 0.f;   // stale ring slots are read (and must be finite)
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_base_s;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // TMEM columns per row block: A_hi [0,32) A_lo [32,64) D [64,128)
    const uint32_t a_hi = tmem + rb * 128, a_lo = a_hi + 32, d0 = a_hi + 64;
    const uint32_t idesc = tc::make_idesc_tf32(128, 64);
    const uint32_t ws_addr = tc::smem_u32(Ws);
    uint64_t* mybar = bars + rb;
    uint32_t cnt = 0u;            // completed waits on this row block's barrier (phase bookkeeping)
    const bool issuer = (warp & 3) == 0 && lane == 0;

    auto issue_mma = [&](int kc) {
        // dT chunk kc -> D
        const uint32_t whi = ws_addr + (uint32_t)(kc * 2) * WSLICE * 4, wlo = whi + WSLICE * 4;
        constexpr uint32_t SBO = (32 / 4) * 128;
        const uint32_t d = d0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dhi = tc::make_smem_desc(whi + ks * 256, 128, SBO);
            const uint64_t dlo = tc::make_smem_desc(wlo + ks * 256, 128, SBO);
            tc::mma_tf32_ts(d, a_hi + ks * 8, dhi, idesc, ks != 0);
            tc::mma_tf32_ts(d, a_lo + ks * 8, dhi, idesc, true);
            tc::mma_tf32_ts(d, a_hi + ks * 8, dlo, idesc, true);
        }
        tc::mma_commit(mybar);
    };

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

        // A operand: this thread's grad_out row, hi / lo (all MMAs of the previous tile have been
        // waited for, so A and both D buffers are free)
        {
            const float* gp = a.G + ((size_t)blk * a.Ndst + (pvalid ? dstp : 0)) * grow + b * 4;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pvalid && 16 * c + 4 * q < a.Cg)
                        v = __ldg(reinterpret_cast<const float4*>(gp + (size_t)(4 * c + q) * Bp * 4));
                    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float hv = __uint_as_float(__float_as_uint(f[e]) & 0xffffe000u);
                        hi[4 * q + e] = __float_as_uint(hv);
                        lo[4 * q + e] = __float_as_uint(f[e] - hv);
                    }
                }
                tc::tmem_st16(lane_base + a_hi + c * 16, hi);
                tc::tmem_st16(lane_base + a_lo + c * 16, lo);
            }
            tc::wait_st();
            tc::fence_before_sync();
            tc::named_bar_sync(1 + rb, 128);
            if (issuer) {
                tc::fence_after_sync();
                issue_mma(0);
            }
        }

        for (int kc = 0; kc < nkc; ++kc) {
            const int cx0 = kc * CXC;
            const int nval = min(CXC, a.Cx - cx0);
            const int ngv = (nval + 3) >> 2;
            // dT of this chunk: TMEM -> registers
            tc::mbar_wait(mybar, cnt & 1);
            ++cnt;
            tc::fence_after_sync();
            float dT[CXC][HP];
            {
                const uint32_t d = lane_base + d0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[16];
                    tc::tmem_ld16(d + c * 16, r);
                    tc::wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) dT[(c * 16 + j) / HP][(c * 16 + j) % HP] = __uint_as_float(r[j]);
                }
                // permute the h slots of this lane (slot k <-> h = k ^ hsel) so that the butterfly below
                // always keeps its low slots and sends its high ones: no selects in the per-pair loop
#pragma unroll
                for (int st = 0; st < NH; ++st) {
                    const int bit = HP >> (st + 1);
                    const bool upper = (lane & (1 << (LBr - 1 - st))) != 0;
#pragma unroll
                    for (int c = 0; c < CXC; ++c)
#pragma unroll
                        for (int k = 0; k < HP; ++k)
                            if (!(k & bit)) {
                                const float x = dT[c][k], y = dT[c][k | bit];
                                dT[c][k] = upper ? y : x;
                                dT[c][k | bit] = upper ? x : y;
                            }
                }
            }
            if (kc + 1 < nkc) {
                // D is free again once every thread of the row block has read it
                tc::fence_before_sync();
                tc::named_bar_sync(1 + rb, 128);
                if (issuer) {
                    tc::fence_after_sync();
                    issue_mma(kc + 1);
                }
            }

            const float* Xc = Xb + (size_t)(cx0 >> 2) * (Bp * 4);
            int* my_idx = idx_s + (warp * PPW + s) * PCH;
            const uint32_t my_idx_sa = tc::smem_u32(my_idx);
            const uint32_t my_ring = tc::smem_u32(ring_s) + warp * RING + lane * 16;
            const float* my_ring_p = ring_s + (warp * RING + lane * 16) / 4;
            auto issue_u = [&](int slot, unsigned u) {
                const float4* xr = reinterpret_cast<const float4*>(Xc + (size_t)(u * xrow));
#pragma unroll
                for (int q = 0; q < NG; ++q)
                    if (q < ngv) tc::cp_async16(my_ring + (slot * NG + q) * 512, xr + q * Bp);
            };
            auto issue = [&](int slot, int e) { issue_u(slot, (unsigned)my_idx[e]); };
            for (int c0 = 0; c0 < maxlen; c0 += PCH) {
                const int clen = min(PCH, len - c0);
                const int cmax = min(PCH, maxlen - c0);
                if (kc == 0 || maxlen > PCH) {   // a segment that fits in one piece is staged once per tile
                    __syncwarp();
                    for (int e = b; e < clen; e += Bp) my_idx[e] = sidx[c0 + e];
                    __syncwarp();
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    if (d < clen) issue(d, d);
                    tc::cp_async_commit();
                }
                if constexpr (TWO) {
                    // two pairs per step: this half-warp accumulates the pair it keeps ("own", slot d + up)
                    // into acc0 and the other half-warp's pair into acc1, so the first exchange needs no select
                    const float* own_p = my_ring_p + up * (NG * 128);
                    const float* oth_p = my_ring_p + (1 - up) * (NG * 128);
                    for (int it = 0; it < cmax; it += D) {
                        unsigned unv[D];                          // source indices of this round's issues, one LDS.128 (see k_contract_tc)
                        if constexpr (D == 4) {
                            const uint4 q4 = tc::lds_u128_early(my_idx_sa + (it + D) * 4);
                            unv[0] = q4.x, unv[1] = q4.y, unv[2] = q4.z, unv[3] = q4.w;
                        }
#pragma unroll
                        for (int d = 0; d < D; d += 2) {
                            tc::cp_async_wait<D - 2>();       // slots d and d+1 have landed
                            const bool v0 = it + d < clen;
                            if (v0) {
                                float4 xa[NG], xb[NG];
#pragma unroll
                                for (int q = 0; q < NG; ++q) {   // groups past ngv: stale finite values times dT == 0
                                    xa[q] = *reinterpret_cast<const float4*>(own_p + (d * NG + q) * 128);
                                    xb[q] = *reinterpret_cast<const float4*>(oth_p + (d * NG + q) * 128);
                                }
                                unsigned u0, u1;
                                if constexpr (D == 4) {
                                    u0 = unv[d], u1 = unv[d + 1];
                                } else {
                                    u0 = tc::lds_u32_early(my_idx_sa + (it + d + D) * 4), u1 = tc::lds_u32_early(my_idx_sa + (it + d + 1 + D) * 4);
                                }
                                float acc0[HP], acc1[HP];
#pragma unroll
                                for (int h = 0; h < HP; ++h) acc0[h] = 0.f, acc1[h] = 0.f;
#pragma unroll
                                for (int q = 0; q < NG; ++q) {
                                    const float fa[4] = {xa[q].x, xa[q].y, xa[q].z, xa[q].w};
                                    const float fb[4] = {xb[q].x, xb[q].y, xb[q].z, xb[q].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
#pragma unroll
                                        for (int h = 0; h < HP; h += 2) {
                                            tc::fma2(acc0[h], acc0[h + 1], dT[4 * q + e][h], dT[4 * q + e][h + 1], fa[e], fa[e]);
                                            tc::fma2(acc1[h], acc1[h + 1], dT[4 * q + e][h], dT[4 * q + e][h + 1], fb[e], fb[e]);
                                        }
                                }
#pragma unroll
                                for (int k = 0; k < HP; k += 2) {   // adds in pairs (packed FP32 add)
                                    const float r0 = __shfl_xor_sync(0xffffffffu, acc1[k], 16);
                                    const float r1 = __shfl_xor_sync(0xffffffffu, acc1[k + 1], 16);
                                    tc::add2(acc0[k], acc0[k + 1], r0, r1);
                                }
#pragma unroll
                                for (int st = 0; st < LBr; ++st) {
                                    const int mask = 1 << (LBr - 1 - st);
                                    if (st < NH) {
                                        const int half = HP >> (st + 1);
                                        if (half >= 2) {
#pragma unroll
                                            for (int k = 0; k < half; k += 2) {
                                                const float r0 = __shfl_xor_sync(0xffffffffu, acc0[k + half], mask);
                                                const float r1 = __shfl_xor_sync(0xffffffffu, acc0[k + 1 + half], mask);
                                                tc::add2(acc0[k], acc0[k + 1], r0, r1);
                                            }
                                        } else {
                                            acc0[0] += __shfl_xor_sync(0xffffffffu, acc0[1], mask);
                                        }
                                    } else {
                                        acc0[0] += __shfl_xor_sync(0xffffffffu, acc0[0], mask);
                                    }
                                }
                                if (writer && it + d + up < clen)
                                    a.dphi[(size_t)(blk * nkc + kc) * slice_stride + (size_t)(beg + c0 + it + d + up) * HP + hsel] = acc0[0];
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
                } else {
                    // software pipeline: pair n+1 moves shared -> registers under the dot products of pair n
                    float4 xc[NG], xn[NG];
                    tc::cp_async_wait<D - 1>();
    #pragma unroll
                    for (int q = 0; q < NG; ++q) xc[q] = *reinterpret_cast<const float4*>(my_ring_p + q * 128);
                    for (int it = 0; it < cmax; it += D) {
    #pragma unroll
                        for (int d = 0; d < D; ++d) {
                            tc::cp_async_wait<D - 2>();
    #pragma unroll
                            for (int q = 0; q < NG; ++q)   // groups past ngv: stale finite values times dT == 0
                                xn[q] = *reinterpret_cast<const float4*>(my_ring_p + (((d + 1) % D) * NG + q) * 128);
                            const unsigned un = (unsigned)my_idx[it + d + D];   // read early; past the list it is unused
                            if (it + d < cmax) {   // warp-uniform: the shuffles below need every lane
                                const bool act = it + d < clen;
                                float part[HP];
    #pragma unroll
                                for (int h = 0; h < HP; ++h) part[h] = 0.f;
                                if (act) {
    #pragma unroll
                                    for (int q = 0; q < NG; ++q) {
                                        const float f[4] = {xc[q].x, xc[q].y, xc[q].z, xc[q].w};
    #pragma unroll
                                        for (int e = 0; e < 4; ++e)
    #pragma unroll
                                            for (int h = 0; h < HP; h += 2)   // dT comes out of tcgen05.ld with h in adjacent registers
                                                tc::fma2(part[h], part[h + 1], dT[4 * q + e][h], dT[4 * q + e][h + 1], f[e], f[e]);
                                    }
                                }
                                // reduce over the batch lanes of the sub-group (transposed butterfly)
    #pragma unroll
                                for (int st = 0; st < LB; ++st) {
                                    const int mask = 1 << (LB - 1 - st);
                                    if (st < NH) {
                                        const int half = HP >> (st + 1);
    #pragma unroll
                                        for (int k = 0; k < half; ++k)
                                            part[k] += __shfl_xor_sync(0xffffffffu, part[k + half], mask);
                                    } else {
                                        part[0] += __shfl_xor_sync(0xffffffffu, part[0], mask);
                                    }
                                }
                                if (act && writer) {
                                    float* dst = a.dphi + (size_t)(blk * nkc + kc) * slice_stride +
                                                 (size_t)(beg + c0 + it + d) * HP + hsel * NF;
    #pragma unroll
                                    for (int k = 0; k < NF; ++k) dst[k] = part[k];
                                }
                                if (it + d + D < clen) issue_u(d, un);
                            }
                            tc::cp_async_commit();
    #pragma unroll
                            for (int q = 0; q < NG; ++q) xc[q] = xn[q];
                        }
                    }
                }
            }
        }
        // every MMA group of this tile has been waited for; make sure all threads are past their
        // last TMEM read before the next tile overwrites A / D
        tc::fence_before_sync();
        tc::named_bar_sync(1 + rb, 128);
        tc::fence_after_sync();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int HP, int LB, int NW>
int launch_dtc(const DphiArgs& a, int B, cudaStream_t st)
{
    constexpr int CXC = 64 / HP, PPW = 32 >> LB;
    const int nkc = qc_ceil_div(a.Cx, CXC);
    const int nblk = qc_nblk(B);
    constexpr int RING = dtc_ring_depth(CXC / 4) * (CXC / 4) * 512;
    const size_t smem = ((size_t)nkc * 2 * WSLICE + NW * PPW * PCH) * 4 + 64 + NW * RING;
    if (smem > 227 * 1024) return 0;
    QC_CUDA((qc_set_max_smem<k_dphi_tc<HP, LB, NW>>((int)smem)));
    const int ntiles = qc_ceil_div(a.Ndst, NW * PPW);
    int gx = qc_num_sms();
    gx = (gx + nblk - 1) / nblk;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, nblk);
    k_dphi_tc<HP, LB, NW><<<grid, NW * 32, smem, st>>>(a, nkc, (size_t)a.P * HP);
    QC_CHECK_LAUNCH();
    return 1;
}

}  // namespace

size_t qc_contract_tc_ws_bytes(int Ndst, int Cx, int Cy, int HP, int B)
{
    // mirrors launch_tc<.., DW = true, NW = 8>: slots of partial dW2 accumulators + one counter per CTA
    const int Bp = qc_bp(B), nblk = qc_nblk(B);
    if (Cy > 32 || !(HP == 4 || HP == 8) || !(Bp == 8 || Bp == 32)) return 0;
    const int nkc = qc_ceil_div(Cx, 64 / HP), ppw = 32 / Bp;
    const int ntiles = qc_ceil_div(Ndst, 8 * ppw);
    int gx = qc_ceil_div(qc_num_sms(), nblk);
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    const int max_commits = qc_ceil_div(ntiles, gx) * 2 * nkc;
    const int nslots = (max_commits - 1) / dw_flush_interval() + 1;
    const size_t nctas = (size_t)gx * nblk;
    return nctas * nslots * ((size_t)nkc * 64 * 32) * sizeof(float) + nctas * sizeof(int);
}

int qc_contract_tc(const ContractArgs& a, int HP, int B, bool dw, cudaStream_t st)
{
    const int Bp = qc_bp(B);
    if (a.Cy > 32 || !(HP == 4 || HP == 8) || !(Bp == 8 || Bp == 32)) return 0;
    if ((size_t)a.Nsrc * qc_cp4(a.Cx) * Bp >= ((size_t)1 << 32)) return 0;
    if (QC_ENV_SET("QC_NO_FAST") || QC_ENV_SET("QC_NO_TC")) return 0;
    // Forward-type launches can also run with 32-wide chunks and two CTAs per SM (QC_TC_K32=1).
    // Measured on C3 (profiles/README.md): 16 instead of 8 warps/SM lift issue utilisation from 45 % to
    // 60 %, but the per-pair overhead is amortised over half the FMAs (+23 % instructions) and the two
    // CTAs halve the L1 hit rate, so the time is the same (4.0 ms); the one-CTA variant is the default.
    if (!dw && QC_ENV_SET("QC_TC_K32")) {
        int rc;
        if (HP == 4) rc = Bp == 8 ? launch_tc<4, 32, 3, false, 8>(a, B, st) : launch_tc<4, 32, 5, false, 8>(a, B, st);
        else rc = Bp == 8 ? launch_tc<8, 32, 3, false, 8>(a, B, st) : launch_tc<8, 32, 5, false, 8>(a, B, st);
        if (rc != 0) return rc;
    }
    // forward-type launches with H = 8: three 128-row blocks (12 warps) per CTA unless QC_TC_NW8=1
    if (!dw && HP == 8 && !QC_ENV_SET("QC_TC_NW8")) {
        const int rc = Bp == 8 ? launch_tc<8, 64, 3, false, 12>(a, B, st) : launch_tc<8, 64, 5, false, 12>(a, B, st);
        if (rc != 0) return rc;
    }
#define QC_T(HPV, LBV) (dw ? launch_tc<HPV, 64, LBV, true, 8>(a, B, st) : launch_tc<HPV, 64, LBV, false, 8>(a, B, st))
    if (HP == 4) return Bp == 8 ? QC_T(4, 3) : QC_T(4, 5);
    return Bp == 8 ? QC_T(8, 3) : QC_T(8, 5);
#undef QC_T
}

// Same eligibility as the FP32-pipe fast path (so qc_dphi_num_slices stays valid for both).
int qc_dphi_tc(const DphiArgs& a, int HP, int B, cudaStream_t st)
{
    if (!qc_dphi_fast_ok(a.Nsrc, a.Cx, a.Cg, HP, B)) return 0;
    if (QC_ENV_SET("QC_NO_TC")) return 0;
    const int Bp = qc_bp(B);
    if (HP == 4) return Bp == 8 ? launch_dtc<4, 3, 8>(a, B, st) : launch_dtc<4, 5, 8>(a, B, st);
    if (!QC_ENV_SET("QC_TC_NW8")) {   // H = 8: four 128-row blocks (16 warps, 4 x 128 TMEM columns) per CTA
        const int rc = Bp == 8 ? launch_dtc<8, 3, 16>(a, B, st) : launch_dtc<8, 5, 16>(a, B, st);
        if (rc != 0) return rc;
    }
    return Bp == 8 ? launch_dtc<8, 3, 8>(a, B, st) : launch_dtc<8, 5, 8>(a, B, st);
}
