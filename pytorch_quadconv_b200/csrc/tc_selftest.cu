// tc_selftest.cu -- known-answer test of the tcgen05 building blocks used by the fused kernels:
// D[128 x 32] = A[128 x 64] * B^T (B given as [32 x 64], K-major) in 3xTF32 (hi/lo split), with A
// written to TMEM by the threads (tcgen05.st), B staged in shared memory in the canonical K-major
// layout, the accumulator read back with tcgen05.ld.  tests/test_gpu_tc.py compares against fp32.
#include "common.cuh"
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1) k_tc_selftest(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D)
{
    __shared__ __align__(128) float Bhi[32 * 64];
    __shared__ __align__(128) float Blo[32 * 64];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    // stage B (hi / lo) in the canonical K-major layout
    for (int idx = tid; idx < 32 * 64; idx += 128) {
        const int n = idx / 64, k = idx % 64;
        const float v = Bm[idx];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const uint32_t off = tc::kmajor_off_bytes(n, k, 64) / 4;
        Bhi[off] = hi;
        Blo[off] = v - hi;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t a_hi = tmem + 0, a_lo = tmem + 64, d_acc = tmem + 128;

    // A row of this thread -> TMEM (hi in columns [0,64), lo in [64,128))
    const float* arow = A + (size_t)tid * 64;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float v = arow[c * 16 + j];
            const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(v - h);
        }
        tc::tmem_st16(lane_base + a_hi + c * 16, hi);
        tc::tmem_st16(lane_base + a_lo + c * 16, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_tf32(128, 32);
        const uint32_t bhi = tc::smem_u32(Bhi), blo = tc::smem_u32(Blo);
        const uint32_t sbo = (64 / 4) * 128;
        bool acc = false;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // K = 64 in steps of 8 (two core matrices)
            const uint64_t dhi = tc::make_smem_desc(bhi + ks * 256, 128, sbo);
            const uint64_t dlo = tc::make_smem_desc(blo + ks * 256, 128, sbo);
            tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dhi, idesc, acc);
            acc = true;
            tc::mma_tf32_ts(d_acc, a_lo + ks * 8, dhi, idesc, true);
            tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dlo, idesc, true);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    uint32_t r0[16], r1[16];
    tc::tmem_ld16(lane_base + d_acc, r0);
    tc::tmem_ld16(lane_base + d_acc + 16, r1);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        D[(size_t)tid * 32 + j] = __uint_as_float(r0[j]);
        D[(size_t)tid * 32 + 16 + j] = __uint_as_float(r1[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// Second building block: both operands from shared memory (SS mode), K-major no-swizzle descriptors:
// D[128 x 32] = A[128 x 64] * B[32 x 64]^T in 3xTF32.
__global__ void __launch_bounds__(128, 1) k_tc_selftest_ss(const float* __restrict__ A, const float* __restrict__ Bm,
                                                            float* __restrict__ D)
{
    extern __shared__ __align__(128) float dyn_smem[];
    float* Ahi = dyn_smem;              // [128][64] K-major
    float* Alo = Ahi + 128 * 64;
    float* Bhi = Alo + 128 * 64;        // [32][64]
    float* Blo = Bhi + 32 * 64;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 32);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    for (int k = 0; k < 64; ++k) {
        const float v = A[(size_t)tid * 64 + k];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const uint32_t off = tc::kmajor_off_bytes(tid, k, 64) / 4;
        Ahi[off] = hi;
        Alo[off] = v - hi;
    }
    for (int idx = tid; idx < 32 * 64; idx += 128) {
        const int n = idx / 64, k = idx % 64;
        const float v = Bm[idx];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const uint32_t off = tc::kmajor_off_bytes(n, k, 64) / 4;
        Bhi[off] = hi;
        Blo[off] = v - hi;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, 32);
        constexpr uint32_t SBO = (64 / 4) * 128;
        bool acc = false;
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ahi = tc::make_smem_desc(tc::smem_u32(Ahi) + ks * 256, 128, SBO);
            const uint64_t alo = tc::make_smem_desc(tc::smem_u32(Alo) + ks * 256, 128, SBO);
            const uint64_t bhi = tc::make_smem_desc(tc::smem_u32(Bhi) + ks * 256, 128, SBO);
            const uint64_t blo = tc::make_smem_desc(tc::smem_u32(Blo) + ks * 256, 128, SBO);
            tc::mma_tf32_ss(tmem, ahi, bhi, idesc, acc);
            acc = true;
            tc::mma_tf32_ss(tmem, alo, bhi, idesc, true);
            tc::mma_tf32_ss(tmem, ahi, blo, idesc, true);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    uint32_t q0[16], q1[16];
    const uint32_t lb = (uint32_t)(warp * 32) << 16;
    tc::tmem_ld16(lb + tmem, q0);
    tc::tmem_ld16(lb + tmem + 16, q1);
    tc::wait_ld();
    for (int j = 0; j < 16; ++j) {
        D[(size_t)tid * 32 + j] = __uint_as_float(q0[j]);
        D[(size_t)tid * 32 + 16 + j] = __uint_as_float(q1[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

// Probe (not used by the product path): does MN-major work for 16-bit operands?  D[128 x 32] = A[128 x 32] *
// B[32 x 32]^T with bf16 operands, kind::f16, no swizzle.  mode bit 0: A MN-major, bit 1: B MN-major.
// Canonical layouts in bytes, T = 8 elements per 16 bytes (cute UMMA INTERLEAVE):
//   K-major : off(mn, k) = (mn % 8) * 16 + (k % 8) * 2 + (k / 8) * LBO + (mn / 8) * SBO
//   MN-major: off(mn, k) = (mn % 8) * 2 + (k % 8) * 16 + (mn / 8) * SBO + (k / 8) * LBO
__global__ void __launch_bounds__(128, 1) k_tc_probe_bf16(const float* __restrict__ A, const float* __restrict__ Bm,
                                                           float* __restrict__ D, int mode)
{
    extern __shared__ __align__(128) unsigned char dyn_b[];
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(dyn_b);              // 128 x 32
    __nv_bfloat16* Bs = As + 128 * 32;                                          // 32 x 32
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 32);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    constexpr int K = 32;
    // strides: K-major: LBO = 128 (next 8 k), SBO = (K/8)*128 (next 8 rows); MN-major: SBO = 128 (next 8 mn), LBO = (MN/8)*128
    auto off = [&](bool mnmaj, int MN, int mn, int k) -> uint32_t {
        if (!mnmaj) return (uint32_t)((mn % 8) * 16 + (k % 8) * 2 + (k / 8) * 128 + (mn / 8) * ((K / 8) * 128));
        return (uint32_t)((mn % 8) * 2 + (k % 8) * 16 + (mn / 8) * 128 + (k / 8) * ((MN / 8) * 128));
    };
    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int m = idx / K, k = idx % K;
        As[off(mode & 1, 128, m, k) / 2] = __float2bfloat16_rn(A[idx]);
    }
    for (int idx = tid; idx < 32 * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        Bs[off(mode & 2, 32, n, k) / 2] = __float2bfloat16_rn(Bm[idx]);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        // kind::f16: D fp32 (1 << 4), A/B format bf16 (1 << 7, 1 << 10)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((mode & 1) ? (1u << 15) : 0u) | ((mode & 2) ? (1u << 16) : 0u) |
                               ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < K / 16; ++ks) {   // K = 16 per MMA = two 8-k groups
            const uint32_t a0 = tc::smem_u32(As) + ((mode & 1) ? ks * 2 * (128 / 8) * 128 : ks * 256);
            const uint32_t b0 = tc::smem_u32(Bs) + ((mode & 2) ? ks * 2 * (32 / 8) * 128 : ks * 256);
            const uint64_t da = (mode & 1) ? tc::make_smem_desc(a0, (128 / 8) * 128, 128) : tc::make_smem_desc(a0, 128, (K / 8) * 128);
            const uint64_t db = (mode & 2) ? tc::make_smem_desc(b0, (32 / 8) * 128, 128) : tc::make_smem_desc(b0, 128, (K / 8) * 128);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(ks != 0)) : "memory");
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    uint32_t q0[16], q1[16];
    const uint32_t lb = (uint32_t)(warp * 32) << 16;
    tc::tmem_ld16(lb + tmem, q0);
    tc::tmem_ld16(lb + tmem + 16, q1);
    tc::wait_ld();
    for (int j = 0; j < 16; ++j) {
        D[(size_t)tid * 32 + j] = __uint_as_float(q0[j]);
        D[(size_t)tid * 32 + 16 + j] = __uint_as_float(q1[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 32);
}


// Micro-benchmark (diagnostic, not on the product path): issue rate of back-to-back tcgen05.mma kind::f16 with both
// operands in shared memory, for the operand layouts / shapes used by csrc/contract_mma.cu.  One thread issues `iters`
// MMAs (cycling over 8 K steps of a zeroed operand area), commits, waits; cycles per MMA are written per CTA.
__global__ void __launch_bounds__(128, 1) k_tc_mma_bench(int M, int N, int a_mn, int b_mn, uint32_t a_lbo, uint32_t a_sbo,
                                                          uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep, uint32_t b_kstep,
                                                          int swz, int iters, long long* __restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char dyn_m[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int done_flag;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) done_flag = 0;
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(dyn_m)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init(&bar2, 1 << 20);
        tc::mbar_fence_init();
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (warp > 0 && ((swz >> 8) & 12)) {
        // background traffic while warp 0 issues: mode bit 2 = shared-memory stores (16 B per lane, conflict free) into
        // an unused area, bit 3 = tcgen05.ld of 16 columns in a loop
        const int mode = swz >> 8;
        uint4* area = reinterpret_cast<uint4*>(dyn_m + 144 * 1024) + (tid - 32);
        uint32_t r[16];
        uint32_t acc = 0;
        while (done_flag == 0) {
            if (mode & 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j) area[(j & 3) * 96] = make_uint4(acc, j, 0, 0);
            }
            if (mode & 8) {
                tc::tmem_ld16(((uint32_t)(warp * 32) << 16) + tmem + 384, r);
                tc::wait_ld();
                acc += r[0];
            }
        }
        if (acc == 0x12345678u) out[0] = 1;
    }
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(M, N, a_mn != 0, b_mn != 0);
        const uint32_t a0 = tc::smem_u32(dyn_m), b0 = a0 + 80 * 1024;
        const uint64_t lt = (uint64_t)(swz & 7) << 61;
        const long long t0 = clock64();
        uint64_t da[8], db[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            da[ks] = tc::make_smem_desc(a0 + ks * a_kstep, a_lbo, a_sbo) | lt;
            db[ks] = tc::make_smem_desc(b0 + ks * b_kstep, b_lbo, b_sbo) | lt;
        }
        const int mode = swz >> 8;
        for (int it = 0; it < iters; it += 8) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                if ((mode & 2) && (ks & 1) == 0) {
                    tc::fence_proxy_async();
                    tc::fence_after_sync();
                }
                asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem + (ks & 1) * 256), "l"(da[ks]),
                             "l"(db[ks]), "r"(idesc) : "memory");
                if ((mode & 1) && (ks & 1)) tc::mma_commit(&bar2);
            }
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
        done_flag = 1;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int qc_tc_probe_bf16(const float* A, const float* B, float* D, int mode, void* stream)
{
    const int smem = (128 * 32 + 32 * 32) * 2;
    k_tc_probe_bf16<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, mode);
    QC_CHECK_LAUNCH();
    return 0;
}

extern "C" int qc_tc_selftest_ss(const float* A, const float* B, float* D, void* stream)
{
    const int smem = (2 * 128 * 64 + 2 * 32 * 64) * 4;
    QC_CUDA((qc_set_max_smem<k_tc_selftest_ss>(smem)));
    k_tc_selftest_ss<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D);
    QC_CHECK_LAUNCH();
    return 0;
}

extern "C" int qc_tc_selftest(const float* A, const float* B, float* D, void* stream)
{
    k_tc_selftest<<<1, 128, 0, (cudaStream_t)stream>>>(A, B, D);
    QC_CHECK_LAUNCH();
    return 0;
}

extern "C" int qc_tc_mma_bench(int M, int N, int a_mn, int b_mn, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int a_kstep,
                               int b_kstep, int swz, int iters, int grid, long long* out, void* stream)
{
    const int smem = 160 * 1024;
    QC_CUDA((qc_set_max_smem<k_tc_mma_bench>(smem)));
    k_tc_mma_bench<<<grid, 128, smem, (cudaStream_t)stream>>>(M, N, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep,
                                                             swz, iters, out);
    QC_CHECK_LAUNCH();
    return 0;
}
