// tc_common.cuh -- thin inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path:
// TMEM allocation, tcgen05.st/ld, tcgen05.mma (kind::tf32, A operand in TMEM, B operand in shared
// memory through a K-major no-swizzle matrix descriptor), tcgen05.commit and mbarrier waits.
// Bit layouts follow the PTX ISA "tcgen05" matrix/instruction descriptor tables.
#pragma once
#include <stdint.h>
#include <cstdio>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one full warp executes these) --------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma reads smem there)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05.st / ld, shape 32x32b: thread i of warp w touches TMEM lane 32*(w%4)+i, 16 columns ----
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle.  Canonical layout (PTX ISA / CUTLASS
// make_umma_desc<Major::K>, LayoutType::INTERLEAVE): 8-row x 16-byte core matrices stored as 128
// contiguous bytes; LBO = byte step between core matrices adjacent in K, SBO = byte step between
// 8-row groups in M/N.  Fields: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
// [61,64) layout type (0 = no swizzle).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major:
// [4,6) D format (1 = f32), [7,10) A format (2 = tf32), [10,13) B format (2 = tf32),
// [15] A major (0 = K), [16] B major (0 = K), [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// all previously issued tcgen05.mma of this thread arrive (once) on the mbarrier when they retire
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#ifdef QC_MBAR_WATCHDOG
// debug builds: a wait that lasts longer than ~1 s reports the barrier (shared-memory address, parity, block, thread) and traps
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const long long t0 = clock64();
    bool said = false;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        const long long dt = clock64() - t0;
        if (dt > 2000000000LL && !said && (threadIdx.x & 31) == 0 && blockIdx.x == 0) {
            printf("mbar_wait stuck: bar@%u parity %u block %d warp %d\n", smem_u32(bar), parity, (int)blockIdx.x, (int)(threadIdx.x >> 5));
            said = true;
        }
        if (dt > 8000000000LL) __trap();
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
#endif

// true in exactly one lane of a fully converged warp (the lane that issues tcgen05.mma / commit); every lane must call it
__device__ __forceinline__ bool elect_one()
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok));
    return ok != 0;
}

// one non-blocking test of a phase (true: the phase with this parity has completed)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- asynchronous global->shared copies (LDGSTS): completion is tracked per commit group in FIFO
// order, NOT through the 6 per-warp scoreboards -- ptxas puts every gather LDG of an unrolled loop on
// one scoreboard, which collapses a register-prefetch pipeline to a depth of one (profiles/README.md).
// Packed FP32 FMA (Blackwell FFMA2): (d0, d1) += (a0, a1) * (b0, b1), each half rounded like fmaf.
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1)
{
    asm("{\n"
        ".reg .b64 d, a, b;\n"
        "mov.b64 d, {%0, %1};\n"
        "mov.b64 a, {%2, %3};\n"
        "mov.b64 b, {%4, %5};\n"
        "fma.rn.f32x2 d, a, b, d;\n"
        "mov.b64 {%0, %1}, d;\n"
        "}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// Packed FP32 add: (d0, d1) += (a0, a1), each half rounded like a scalar add.
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1)
{
    asm("{\n"
        ".reg .b64 d, a;\n"
        "mov.b64 d, {%0, %1};\n"
        "mov.b64 a, {%2, %3};\n"
        "add.rn.f32x2 d, d, a;\n"
        "mov.b64 {%0, %1}, d;\n"
        "}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1));
}

// Shared-memory load the compiler may not sink to its use (used to start a dependent address chain early).
__device__ __forceinline__ uint32_t lds_u32_early(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// four consecutive 32-bit words with one LDS.128 (16-byte aligned address), same no-sinking property
__device__ __forceinline__ uint4 lds_u128_early(uint32_t saddr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
// same with zero fill: copies 16 bytes when valid, writes 16 zero bytes otherwise (src-size operand)
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, const void* gsrc, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(uint32_t smem_dst, const void* gsrc, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2-only variant (bypasses L1): for streams that are read once per CTA
__device__ __forceinline__ void cp_async16_cg_zfill(uint32_t smem_dst, const void* gsrc, bool valid)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(valid ? 16 : 0) : "memory");
}
// the mbarrier receives one arrival (not counted in its pending count: .noinc) once all cp.async issued so far by
// the executing thread have completed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Byte offset of element (row n, k) inside a K-major no-swizzle operand tile with K_TILE fp32
// columns: core matrices of 8 rows x 4 fp32, adjacent in K (LBO = 128 B), 8-row groups SBO apart.
__host__ __device__ constexpr uint32_t kmajor_off_bytes(int n, int k, int k_tile)
{
    return (uint32_t)((n >> 3) * (k_tile / 4) * 128 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4);
}

// ---- kind::f16 (bf16 operands, fp32 accumulate), both operands from shared memory -------------------------
// Instruction descriptor: [4,6) D format (1 = f32), [7,10) A format (1 = bf16), [10,13) B format (1 = bf16),
// [15] A major (0 = K, 1 = MN), [16] B major, [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major ? (1u << 15) : 0u) | (b_mn_major ? (1u << 16) : 0u) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// plain arrive (count 1) on a CTA-local mbarrier
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Canonical no-swizzle ("interleave") operand tiles for 16-bit elements, in bytes.  A core matrix is 8 lines of
// 16 bytes (8 elements); K-major: line = one row (M/N index), the 8 elements run along K;  MN-major: line = one
// k, the 8 elements run along M/N.  `stride_mn8` = byte step between groups of 8 rows (M/N), `stride_k8` = byte
// step between groups of 8 k.  Descriptor fields: K-major (LBO = stride_k8, SBO = stride_mn8); MN-major
// (LBO = stride_k8, SBO = stride_mn8) as well -- the leading-dimension field always steps along K here.
__host__ __device__ constexpr uint32_t kmajor16_off(int mn, int k, uint32_t stride_mn8, uint32_t stride_k8)
{
    return (uint32_t)(mn >> 3) * stride_mn8 + (uint32_t)(k >> 3) * stride_k8 + (uint32_t)(mn & 7) * 16 + (uint32_t)(k & 7) * 2;
}
__host__ __device__ constexpr uint32_t mnmajor16_off(int mn, int k, uint32_t stride_mn8, uint32_t stride_k8)
{
    return (uint32_t)(mn >> 3) * stride_mn8 + (uint32_t)(k >> 3) * stride_k8 + (uint32_t)(k & 7) * 16 + (uint32_t)(mn & 7) * 2;
}

}  // namespace tc
