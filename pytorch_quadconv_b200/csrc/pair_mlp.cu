// pair_mlp.cu -- K2a: hidden layers of the filter MLP, one thread per (out,in) pair.
//
// Reference: per pair the reference gathers the quadrature weight (quadconv.py:247), forms the
// offset eval_locs = P_out[idx0] - P_in[idx1] (quadconv.py:250-253) and pushes it through
// self.filter = Sequential(Linear(bias=False), Sin, ..., Linear(bias=False)) (quadconv.py:130-145,
// utils/misc.py:18-19).  The LAST Linear (h_L -> Ci*Co) is linear in its input, so it is folded
// into the contraction kernel (contract.cu); this file evaluates everything before it:
//     phi[p][h] = w[in(p)] * x_L[h],  x_0 = z_p,  x_{l+1} = sin(W_l x_l)
// and its backward.  The hidden widths are tiny (4..32), so this is CUDA-core work: the weights sit
// zero-padded in shared memory and are read as warp-uniform broadcasts; activations stay in
// registers.  sinf/cosf are the accurate versions (no fast-math) to hold the 1e-5 fp32 bound.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

struct MlpDev {
    int bf16;   // operands of every Linear rounded to bf16 (qc_mlp_t::operand_bf16)
    int n_hidden;
    int dims[QC_MAX_MLP_LAYERS + 1];
    const float* w[QC_MAX_MLP_LAYERS];
    float* dw[QC_MAX_MLP_LAYERS];
};

__device__ __forceinline__ float rnd_operand(float v, int bf16)
{
    return bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}

template <int WMAX>
__device__ __forceinline__ void load_weights(const MlpDev& m, float* Wsm)
{
    // Wsm[l][o*WMAX + k], zero padded
    for (int l = 0; l < m.n_hidden; ++l) {
        const int nin = m.dims[l], nout = m.dims[l + 1];
        for (int t = threadIdx.x; t < WMAX * WMAX; t += blockDim.x) {
            int o = t / WMAX, k = t % WMAX;
            Wsm[l * WMAX * WMAX + t] = (o < nout && k < nin) ? rnd_operand(m.w[l][o * nin + k], m.bf16) : 0.f;
        }
    }
}

// Accurate sine / cosine as real calls: the tensor-core kernels below evaluate 64 of them per pair in fully unrolled
// register code, and 64 inlined copies of sincosf (fast path + Payne-Hanek fallback) overflow the instruction cache
// (ncu: 37 % of the backward kernel's stall samples were `no_instruction` on those lines).
__device__ __noinline__ float2 sincos_call(float x)
{
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}
__device__ __noinline__ float sin_call(float x) { return sinf(x); }

// scalar kernels: inlined by default (QC_MLP_CALLS=1 at compile time switches them to the calls for experiments)
#ifdef QC_MLP_CALLS
#define QC_SCALAR_SIN(v) sin_call(v)
#define QC_SCALAR_SINCOS(v, sv_, cv_) do { const float2 sc__ = sincos_call(v); sv_ = sc__.x; cv_ = sc__.y; } while (0)
#else
#define QC_SCALAR_SIN(v) sinf(v)
#define QC_SCALAR_SINCOS(v, sv_, cv_) sincosf(v, &sv_, &cv_)
#endif

template <int WMAX>
__global__ void __launch_bounds__(128)
k_pair_phi_fwd(const float* __restrict__ out_pts, const float* __restrict__ in_pts,
               const float* __restrict__ quad_w, const int* __restrict__ pair_row,
               const int* __restrict__ pair_col, int64_t P, int d, MlpDev m, int HP,
               float* __restrict__ phi)
{
    extern __shared__ float smem[];
    float* Wsm = smem;
    load_weights<WMAX>(m, Wsm);
    __syncthreads();
    const int H = m.dims[m.n_hidden];
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const int r = pair_row[p], c = pair_col[p];
        float a[WMAX];
#pragma unroll
        for (int k = 0; k < WMAX; ++k) a[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < d && k < WMAX) a[k] = __fsub_rn(out_pts[(size_t)r * d + k], in_pts[(size_t)c * d + k]);
#pragma unroll
        for (int k = 0; k < WMAX; ++k) a[k] = rnd_operand(a[k], m.bf16);
        for (int l = 0; l < m.n_hidden; ++l) {
            const float* W = Wsm + l * WMAX * WMAX;
            float t[WMAX];
#pragma unroll
            for (int o = 0; o < WMAX; ++o) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < WMAX; ++k) acc = fmaf(W[o * WMAX + k], a[k], acc);
                t[o] = rnd_operand(QC_SCALAR_SIN(acc), m.bf16);
            }
#pragma unroll
            for (int o = 0; o < WMAX; ++o) a[o] = t[o];
        }
        const float wq = quad_w ? quad_w[c] : 1.f;
        float* dst = phi + (size_t)p * HP;
#pragma unroll
        for (int h = 0; h < WMAX; ++h)
            if (h < HP) dst[h] = h < H ? wq * a[h] : 0.f;
        // HP may exceed WMAX only when H <= WMAX < HP, which the host never configures
    }
}

// Tensor-core forward for wide filter MLPs (hidden width 16 or 32, two or more hidden layers; BASELINE config 4).
// One thread per pair cannot feed the FP32 pipe here: a 32 x 32 Linear is 1024 dependent-free FMAs whose weight
// operands all come from shared memory.  The hidden Linears l >= 1 are therefore run as dense layers on tcgen05 over
// tiles of 128 pairs:  D[128 pairs][W] = A[128][W] * W_l^T  with the tile's activations as the A operand in TMEM
// (thread = pair = TMEM lane: written straight from registers, split into tf32 hi / lo) and the weights (hi, lo) as
// K-major shared-memory operands, three kind::tf32 MMAs per product (A_hi B_hi + A_lo B_hi + A_hi B_lo: ~1e-6, the
// fp32 parity bound is 1e-5).  Layer 0 (spatial_dim <= 3 inputs) and the sines stay on the FP32 / SFU pipes.
template <int W>
__global__ void __launch_bounds__(128, 4)
k_pair_phi_fwd_tc(const float* __restrict__ out_pts, const float* __restrict__ in_pts,
                  const float* __restrict__ quad_w, const int* __restrict__ pair_row,
                  const int* __restrict__ pair_col, int64_t P, int d, MlpDev m, int HP, float* __restrict__ phi)
{
    extern __shared__ __align__(128) float smem[];
    float* W0s = smem;                          // [W][4] layer 0, zero padded
    float* Bs = smem + W * 4;                   // [n_hidden - 1][hi, lo][W * W] canonical K-major tiles
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    constexpr int TCOLS = W == 32 ? 128 : 64;   // A_hi, A_lo, D: 3 W columns
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int L = m.n_hidden;
    for (int t = tid; t < W * 4; t += 128) {
        const int o = t >> 2, k = t & 3;
        W0s[t] = (o < m.dims[1] && k < m.dims[0]) ? rnd_operand(m.w[0][o * m.dims[0] + k], m.bf16) : 0.f;
    }
    for (int l = 1; l < L; ++l) {
        const int nin = m.dims[l], nout = m.dims[l + 1];
        char* hi = reinterpret_cast<char*>(Bs + (size_t)(l - 1) * 2 * W * W);
        char* lo = hi + W * W * 4;
        for (int t = tid; t < W * W; t += 128) {
            const int o = t / W, k = t % W;
            const float v = (o < nout && k < nin) ? rnd_operand(m.w[l][o * nin + k], m.bf16) : 0.f;
            const float hv = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            const uint32_t off = tc::kmajor_off_bytes(o, k, W);
            *reinterpret_cast<float*>(hi + off) = hv;
            *reinterpret_cast<float*>(lo + off) = v - hv;
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t a_hi = tmem, a_lo = tmem + W, d_acc = tmem + 2 * W;
    const uint32_t idesc = tc::make_idesc_tf32(128, W);
    constexpr uint32_t SBO = (W / 4) * 128;
    const int H = m.dims[L];
    uint32_t phase = 0;

    for (int64_t base = (int64_t)blockIdx.x * 128; base < P; base += (int64_t)gridDim.x * 128) {
        const int64_t p = base + tid;
        const bool valid = p < P;
        int r = 0, c = 0;
        if (valid) r = pair_row[p], c = pair_col[p];
        float z[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < d) z[k] = rnd_operand(__fsub_rn(out_pts[(size_t)r * d + k], in_pts[(size_t)c * d + k]), m.bf16);
        }
        float a[W];
#pragma unroll
        for (int o = 0; o < W; ++o) {
            const float4 w4 = *reinterpret_cast<const float4*>(W0s + o * 4);
            // same association order as the scalar kernel: ((w0 z0 + 0) + w1 z1) + w2 z2
            float acc = fmaf(w4.x, z[0], 0.f);
            acc = fmaf(w4.y, z[1], acc);
            acc = fmaf(w4.z, z[2], acc);
            a[o] = rnd_operand(sin_call(acc), m.bf16);
        }
        for (int l = 1; l < L; ++l) {
#pragma unroll
            for (int cc = 0; cc < W / 16; ++cc) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x = a[cc * 16 + j];
                    const float hv = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                    hi[j] = __float_as_uint(hv);
                    lo[j] = __float_as_uint(x - hv);
                }
                tc::tmem_st16(lane_base + a_hi + cc * 16, hi);
                tc::tmem_st16(lane_base + a_lo + cc * 16, lo);
            }
            tc::wait_st();
            tc::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t bhi = tc::smem_u32(Bs + (size_t)(l - 1) * 2 * W * W), blo = bhi + W * W * 4;
#pragma unroll
                for (int ks = 0; ks < W / 8; ++ks) {
                    const uint64_t dhi = tc::make_smem_desc(bhi + ks * 256, 128, SBO);
                    const uint64_t dlo = tc::make_smem_desc(blo + ks * 256, 128, SBO);
                    tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dhi, idesc, ks != 0);
                    tc::mma_tf32_ts(d_acc, a_lo + ks * 8, dhi, idesc, true);
                    tc::mma_tf32_ts(d_acc, a_hi + ks * 8, dlo, idesc, true);
                }
                tc::mma_commit(&bar);
            }
            tc::mbar_wait(&bar, phase);
            phase ^= 1u;
            tc::fence_after_sync();
#pragma unroll
            for (int cc = 0; cc < W / 16; ++cc) {
                uint32_t rr[16];
                tc::tmem_ld16(lane_base + d_acc + cc * 16, rr);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) a[cc * 16 + j] = rnd_operand(sin_call(__uint_as_float(rr[j])), m.bf16);
            }
        }
        if (valid) {
            const float wq = quad_w ? quad_w[c] : 1.f;
            float4* dst = reinterpret_cast<float4*>(phi + (size_t)p * HP);
#pragma unroll
            for (int q = 0; q < W / 4; ++q)
                if (4 * q < HP)
                    dst[q] = make_float4(4 * q < H ? wq * a[4 * q] : 0.f, 4 * q + 1 < H ? wq * a[4 * q + 1] : 0.f,
                                         4 * q + 2 < H ? wq * a[4 * q + 2] : 0.f, 4 * q + 3 < H ? wq * a[4 * q + 3] : 0.f);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

// Backward.  Block = 128 pairs per sweep.  Per layer the block stages (da, x) of its pairs in
// shared memory and every thread owns a fixed set of dW[o][k] entries, so the weight gradient is
// reduced without atomics inside the block; one atomicAdd per entry per block at the end.
// LT > 0: the number of hidden layers is a compile-time constant, so the per-layer activations and
// cosines live in registers (with a runtime layer count they are dynamically indexed and ptxas puts
// them in local memory); LT == 0 is the general version.
template <int WMAX, int LT>
__global__ void __launch_bounds__(128)
k_pair_phi_bwd(const float* __restrict__ out_pts, const float* __restrict__ in_pts,
               const float* __restrict__ quad_w, const int* __restrict__ pair_row,
               const int* __restrict__ pair_col, int64_t P, int d, MlpDev m, int HP,
               const float* __restrict__ dphi, int nslices, float* __restrict__ d_quad_w,
               float* __restrict__ part, float* __restrict__ dwq_out)
{
    // part / dwq_out != nullptr: deterministic mode (qc_pair_phi_bwd_ws) -- the block's weight-gradient sums go to its
    // row of `part`, the per-pair quadrature-weight terms to dwq_out[p]; k_mlp_finish / k_dwq_segsum add them in fixed
    // orders.  Otherwise one atomicAdd per entry per block / per pair.
    constexpr int NT = 128;
    extern __shared__ float smem[];
    float* Wsm = smem;                                  // [n_hidden][WMAX*WMAX]
    float* dWs = Wsm + m.n_hidden * WMAX * WMAX;        // [n_hidden][WMAX*WMAX]
    float* da_s = dWs + m.n_hidden * WMAX * WMAX;       // [NT][WMAX]
    float* x_s = da_s + NT * WMAX;                      // [NT][WMAX]
    float* part_s = x_s + NT * WMAX;                    // [NQR][WMAX*WMAX] partial tiles of the block's q ranges
    load_weights<WMAX>(m, Wsm);
    for (int t = threadIdx.x; t < m.n_hidden * WMAX * WMAX; t += NT) dWs[t] = 0.f;
    __syncthreads();
    const int H = m.dims[m.n_hidden];
    constexpr int LMAX = LT > 0 ? LT : QC_MAX_MLP_LAYERS;
    const int L = LT > 0 ? LT : m.n_hidden;

    // Two-stage software pipeline over the block's sweeps: the pair indices are loaded two sweeps ahead, the data they
    // address (points, quadrature weight) and the pair's d(phi) row one sweep ahead, so the two dependent global-memory
    // latencies of a sweep are paid under the arithmetic of the previous ones instead of at its start.
    struct PairIn { int c; float z[3]; float wq; float g[WMAX]; };
    const int64_t stride = (int64_t)gridDim.x * NT;
    auto load_idx = [&](int64_t base_, int& r_, int& c_) {
        const int64_t p_ = base_ + threadIdx.x;
        r_ = 0, c_ = 0;
        if (p_ < P) r_ = pair_row[p_], c_ = pair_col[p_];
    };
    auto load_in = [&](int64_t base_, int r_, int c_, PairIn& in) {
        const int64_t p_ = base_ + threadIdx.x;
        const bool v_ = p_ < P;
        in.c = c_;
#pragma unroll
        for (int k = 0; k < 3; ++k) in.z[k] = 0.f;
#pragma unroll
        for (int h = 0; h < WMAX; ++h) in.g[h] = 0.f;
        in.wq = 1.f;
        if (v_) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < d && k < WMAX)
                    in.z[k] = rnd_operand(__fsub_rn(out_pts[(size_t)r_ * d + k], in_pts[(size_t)c_ * d + k]), m.bf16);
            if (quad_w) in.wq = quad_w[c_];
            // d loss / d phi of this pair: the slices are summed in a fixed order, 128 bits at a time
            for (int sl = 0; sl < nslices; ++sl) {
                const float4* src = reinterpret_cast<const float4*>(dphi + ((size_t)sl * P + p_) * HP);
#pragma unroll
                for (int q = 0; q < WMAX / 4; ++q)
                    if (4 * q < HP) {
                        const float4 v = __ldg(src + q);
                        in.g[4 * q] += v.x, in.g[4 * q + 1] += v.y, in.g[4 * q + 2] += v.z, in.g[4 * q + 3] += v.w;
                    }
            }
        }
    };
    const int64_t base0 = (int64_t)blockIdx.x * NT;
    int r1, c1, r2 = 0, c2 = 0;
    PairIn nxt;
    load_idx(base0, r1, c1);
    load_in(base0, r1, c1, nxt);
    load_idx(base0 + stride, r1, c1);
    for (int64_t base = base0; base < P; base += stride) {
        const int64_t p = base + threadIdx.x;
        const bool valid = p < P;
        const PairIn cur = nxt;
        load_idx(base + 2 * stride, r2, c2);                 // indices two sweeps ahead
        if (base + stride < P) load_in(base + stride, r1, c1, nxt);   // data one sweep ahead (its indices arrived a sweep ago)
        r1 = r2, c1 = c2;
        const int c = cur.c;
        // forward recompute, keeping every layer input and the cosine of every pre-activation
        float xs[LMAX + 1][WMAX];
        float cs[LMAX][WMAX];
#pragma unroll
        for (int k = 0; k < WMAX; ++k) xs[0][k] = (k < 3 && valid) ? cur.z[k < 3 ? k : 0] : 0.f;
#pragma unroll(LT > 0 ? LMAX : 1)
        for (int l = 0; l < LMAX; ++l) {
            if (l >= L) break;
            const float* W = Wsm + l * WMAX * WMAX;
#pragma unroll
            for (int o = 0; o < WMAX; ++o) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < WMAX; ++k) acc = fmaf(W[o * WMAX + k], xs[l][k], acc);
                float sv, cv;
                QC_SCALAR_SINCOS(acc, sv, cv);
                xs[l + 1][o] = rnd_operand(sv, m.bf16);   // straight-through: the derivative ignores the rounding
                cs[l][o] = cv;
            }
        }
        const float wq = cur.wq;
        float dx[WMAX];
        float dwq = 0.f;
#pragma unroll
        for (int h = 0; h < WMAX; ++h) {
            const float g = h < H ? cur.g[h] : 0.f;
            dwq = fmaf(g, LT > 0 ? xs[LT][h] : xs[L][h], dwq);
            dx[h] = wq * g;
        }
        if (valid && dwq_out) dwq_out[p] = dwq;
        else if (valid && d_quad_w) atomicAdd(&d_quad_w[c], dwq);

#pragma unroll(LT > 0 ? LMAX : 1)
        for (int l = LMAX - 1; l >= 0; --l) {
            if (l >= L) continue;
            const float* W = Wsm + l * WMAX * WMAX;
            float da[WMAX];
#pragma unroll
            for (int o = 0; o < WMAX; ++o) {
                da[o] = dx[o] * cs[l][o];
                da_s[threadIdx.x * WMAX + o] = da[o];
                x_s[threadIdx.x * WMAX + o] = xs[l][o];
            }
            __syncthreads();
            // dW_l[o][k] += sum_q da[q][o] * x[q][k]: a register-tiled mini-GEMM over the block's 128 pairs.
            // Every thread owns a TL x TL tile of (o,k) and a range of q; vector loads of TL values feed
            // TL*TL FMAs; the q ranges are combined in a fixed order (one shared-memory slice per range).
            {
                constexpr int TL = WMAX >= 16 ? 4 : 2;
                constexpr int TPR = WMAX / TL, NTILE = TPR * TPR;
                constexpr int NQR = NT / NTILE, QL = NT / NQR;
                const int tile = threadIdx.x % NTILE, qr = threadIdx.x / NTILE;
                const int o0 = (tile / TPR) * TL, k0 = (tile % TPR) * TL;
                float acc[TL][TL];
#pragma unroll
                for (int i = 0; i < TL; ++i)
#pragma unroll
                    for (int j = 0; j < TL; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                for (int q = qr * QL; q < (qr + 1) * QL; ++q) {
                    float dv[TL], xv[TL];
                    if (TL == 4) {
                        const float4 a4 = *reinterpret_cast<const float4*>(da_s + q * WMAX + o0);
                        const float4 b4 = *reinterpret_cast<const float4*>(x_s + q * WMAX + k0);
                        dv[0] = a4.x, dv[1] = a4.y, dv[TL - 2] = a4.z, dv[TL - 1] = a4.w;
                        xv[0] = b4.x, xv[1] = b4.y, xv[TL - 2] = b4.z, xv[TL - 1] = b4.w;
                    } else {
                        const float2 a2 = *reinterpret_cast<const float2*>(da_s + q * WMAX + o0);
                        const float2 b2 = *reinterpret_cast<const float2*>(x_s + q * WMAX + k0);
                        dv[0] = a2.x, dv[1] = a2.y;
                        xv[0] = b2.x, xv[1] = b2.y;
                    }
#pragma unroll
                    for (int i = 0; i < TL; ++i)
#pragma unroll
                        for (int j = 0; j < TL; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < TL; ++i)
#pragma unroll
                    for (int j = 0; j < TL; ++j) part_s[qr * WMAX * WMAX + (o0 + i) * WMAX + k0 + j] = acc[i][j];
                __syncthreads();
                for (int t = threadIdx.x; t < WMAX * WMAX; t += NT) {
                    float sacc = 0.f;
#pragma unroll
                    for (int q = 0; q < NQR; ++q) sacc += part_s[q * WMAX * WMAX + t];
                    dWs[l * WMAX * WMAX + t] += sacc;
                }
            }
            if (l > 0) {
#pragma unroll
                for (int k = 0; k < WMAX; ++k) {
                    float acc = 0.f;
#pragma unroll
                    for (int o = 0; o < WMAX; ++o) acc = fmaf(W[o * WMAX + k], da[o], acc);
                    dx[k] = acc;
                }
            }
            __syncthreads();
        }
    }
    // flush
    __syncthreads();
    if (part != nullptr) {
        float* row = part + (size_t)blockIdx.x * (L * WMAX * WMAX);
        for (int t = threadIdx.x; t < L * WMAX * WMAX; t += NT) row[t] = dWs[t];
        return;
    }
    for (int l = 0; l < L; ++l) {
        const int nin = m.dims[l], nout = m.dims[l + 1];
        for (int t = threadIdx.x; t < WMAX * WMAX; t += NT) {
            int o = t / WMAX, k = t % WMAX;
            if (o < nout && k < nin) atomicAdd(&m.dw[l][o * nin + k], dWs[l * WMAX * WMAX + t]);
        }
    }
}

// Deterministic finish of qc_pair_phi_bwd_ws, one launch: blocks [0, nb_w) add the blocks' rows of `part`
// ([nblocks][L][W*W]) into dw[l][o][k] -- one warp per entry, lane j adds rows j, j + 32, ... in sequence, then a fixed
// shuffle tree; blocks [nb_w, ...) add the pairs' quadrature-weight terms into d_quad_w[k] in the order of the CSC list
// of in-point k (one thread per in-point).
__global__ void __launch_bounds__(128)
k_mlp_finish(const float* __restrict__ part, int nblocks, int W, MlpDev m, int nb_w, const float* __restrict__ dwq,
             const int* __restrict__ csc_ptr, const int* __restrict__ csc_pair, int Ni, float* __restrict__ d_quad_w)
{
    if ((int)blockIdx.x >= nb_w) {
        const int k = ((int)blockIdx.x - nb_w) * 128 + threadIdx.x;
        if (k >= Ni) return;
        float s = 0.f;
        for (int e = csc_ptr[k]; e < csc_ptr[k + 1]; ++e) s += dwq[csc_pair[e]];
        d_quad_w[k] += s;
        return;
    }
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int L = m.n_hidden;
    if (e >= L * W * W) return;
    const int l = e / (W * W), o = (e % (W * W)) / W, k = e % W;
    if (o >= m.dims[l + 1] || k >= m.dims[l]) return;
    float v = 0.f;
    for (int b = lane; b < nblocks; b += 32) v += part[(size_t)b * (L * W * W) + e];
#pragma unroll
    for (int mk = 16; mk >= 1; mk >>= 1) v += __shfl_xor_sync(0xffffffffu, v, mk);
    if (lane == 0) m.dw[l][o * m.dims[l] + k] += v;
}

// Tensor-core backward for wide filter MLPs with exactly two hidden layers (d -> h1 -> h2, widths <= W = 16 or 32:
// BASELINE config 4).  Per tile of 128 pairs (thread = pair = TMEM lane) four dense products run on tcgen05 in 3xTF32:
//   F  pre1 = x1 W1^T                      recompute of the hidden-to-hidden Linear        A = x1 in TMEM, B = W1
//   X  dx1  = da1 W1                       gradient w.r.t. its input                       A = da1 in TMEM, B = W1^T
//   G  dW1 += da1^T x1   (sum over pairs)  both operands staged K-major (K = pair) in one 128-byte-swizzled
//   H  dW0 += da0^T z                      shared-memory tile: rows [0, W) = da, rows [W, 2W) = x1 / z
// G and H put da in the first rows of the tile and the layer input right behind it, so the B operand is a row offset
// into the A tile (M = 64 rows are multiplied, the rows past W yield products nobody reads).  Their accumulators are
// read out after every tile (48-MMA chains, see the truncation note in DESIGN.md) and carried in fp32 registers by
// the lanes that own those rows (an M = 64 accumulator keeps rows 16 w .. 16 w + 15 in lanes 0..15 of lane quadrant w); one
// atomicAdd per entry per CTA at the end, as in the scalar kernel.
// Sines / cosines, layer 0 and the quadrature-weight gradient stay on the FP32 / SFU pipes.
template <int W>
__global__ void __launch_bounds__(128, 2)
k_pair_phi_bwd_tc(const float* __restrict__ out_pts, const float* __restrict__ in_pts,
                  const float* __restrict__ quad_w, const int* __restrict__ pair_row,
                  const int* __restrict__ pair_col, int64_t P, int d, MlpDev m, int HP,
                  const float* __restrict__ dphi, int nslices, float* __restrict__ d_quad_w,
                  float* __restrict__ part, float* __restrict__ dwq_out)
{
    extern __shared__ __align__(1024) float smem_raw[];
    float* S_hi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // [4][64 rows][32]
    float* S_lo = S_hi + 4 * 64 * 32;
    float* B1 = S_lo + 4 * 64 * 32;              // [hi, lo][W*W]  W1   as [N = o][K = k], canonical K-major
    float* B1T = B1 + 2 * W * W;                 // [hi, lo][W*W]  W1^T as [N = k][K = o]
    float* W0s = B1T + 2 * W * W;                // [W][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    constexpr int TCOLS = W == 32 ? 128 : 64;    // A_hi, A_lo, D (F / X), D (G / H)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n0 = m.dims[0], n1 = m.dims[1], n2 = m.dims[2];
    for (int t = tid; t < 2 * 4 * 64 * 32; t += 128) S_hi[t] = 0.f;      // S_hi and S_lo: rows nobody writes must stay finite
    for (int t = tid; t < W * 4; t += 128) {
        const int o = t >> 2, k = t & 3;
        W0s[t] = (o < n1 && k < n0) ? rnd_operand(m.w[0][o * n0 + k], m.bf16) : 0.f;
    }
    for (int t = tid; t < W * W; t += 128) {
        const int o = t / W, k = t % W;
        const float v = (o < n2 && k < n1) ? rnd_operand(m.w[1][o * n1 + k], m.bf16) : 0.f;
        const float hv = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const uint32_t off = tc::kmajor_off_bytes(o, k, W) / 4, offT = tc::kmajor_off_bytes(k, o, W) / 4;
        B1[off] = hv, B1[W * W + off] = v - hv;
        B1T[offT] = hv, B1T[W * W + offT] = v - hv;
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t a_hi = tmem, a_lo = tmem + W, d_fx = tmem + 2 * W, d_gh = tmem + 3 * W;
    const uint32_t idesc_fx = tc::make_idesc_tf32(128, W), idesc_g = tc::make_idesc_tf32(64, W), idesc_h = tc::make_idesc_tf32(64, 8);
    constexpr uint32_t SBO = (W / 4) * 128;
    constexpr uint64_t SW128 = (uint64_t)2 << 61;
    uint32_t phase = 0;
    float dW1acc[W], dW0acc[8];          // lanes 0..15 of warps 0 (and 1): one row of dW1 / dW0 summed over this CTA's tiles
#pragma unroll
    for (int k = 0; k < W; ++k) dW1acc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) dW0acc[k] = 0.f;

    auto put_a = [&](const float (&v)[W]) {              // this pair's row of the TMEM A operand (hi, lo)
#pragma unroll
        for (int cc = 0; cc < W / 16; ++cc) {
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
    };
    // element (row, pair = this thread) of the staged tile: K-major over the 128 pairs, 128-byte swizzle, four 32-pair sub-tiles
    const uint32_t s_col = (uint32_t)((tid >> 5) * (64 * 32));
    auto put_s = [&](int row, float x) {
        const uint32_t o = s_col + (uint32_t)((row >> 3) * 256 + (row & 7) * 32 + ((((lane >> 2) ^ (row & 7)) << 2) | (lane & 3)));
        const float hv = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        S_hi[o] = hv;
        S_lo[o] = x - hv;
    };
    auto mma_ts = [&](uint32_t dcol, const float* Bt, uint32_t idesc) {     // D[128][W] = A(TMEM) * Bt^T, fresh accumulator
        const uint32_t bhi = tc::smem_u32(Bt), blo = bhi + W * W * 4;
#pragma unroll
        for (int ks = 0; ks < W / 8; ++ks) {
            const uint64_t dhi = tc::make_smem_desc(bhi + ks * 256, 128, SBO);
            const uint64_t dlo = tc::make_smem_desc(blo + ks * 256, 128, SBO);
            tc::mma_tf32_ts(dcol, a_hi + ks * 8, dhi, idesc, ks != 0);
            tc::mma_tf32_ts(dcol, a_lo + ks * 8, dhi, idesc, true);
            tc::mma_tf32_ts(dcol, a_hi + ks * 8, dlo, idesc, true);
        }
    };
    auto mma_ss = [&](uint32_t idesc) {                  // D[64][N] = S[rows 0..63] * S[rows W..W+N)^T over the 128 pairs
        const uint32_t shi = tc::smem_u32(S_hi), slo = tc::smem_u32(S_lo);
        constexpr uint32_t BOFF = (W >> 3) * 1024;       // byte offset of row W inside a sub-tile
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
            const uint32_t o = (uint32_t)((ks >> 2) * (64 * 128) + (ks & 3) * 32);
            const uint64_t ah = tc::make_smem_desc(shi + o, 16, 1024) | SW128, al = tc::make_smem_desc(slo + o, 16, 1024) | SW128;
            const uint64_t bh = tc::make_smem_desc(shi + o + BOFF, 16, 1024) | SW128, bl = tc::make_smem_desc(slo + o + BOFF, 16, 1024) | SW128;
            tc::mma_tf32_ss(d_gh, ah, bh, idesc, ks != 0);
            tc::mma_tf32_ss(d_gh, al, bh, idesc, true);
            tc::mma_tf32_ss(d_gh, ah, bl, idesc, true);
        }
    };
    auto wait_mma = [&]() {
        tc::mbar_wait(&bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
    };
    auto get_d = [&](uint32_t dcol, float (&v)[W]) {
#pragma unroll
        for (int cc = 0; cc < W / 16; ++cc) {
            uint32_t rr[16];
            tc::tmem_ld16(lane_base + dcol + cc * 16, rr);
            tc::wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[cc * 16 + j] = __uint_as_float(rr[j]);
        }
    };

    for (int64_t base = (int64_t)blockIdx.x * 128; base < P; base += (int64_t)gridDim.x * 128) {
        const int64_t p = base + tid;
        const bool valid = p < P;
        int r = 0, c = 0;
        if (valid) r = pair_row[p], c = pair_col[p];
        float z[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < d) z[k] = rnd_operand(__fsub_rn(out_pts[(size_t)r * d + k], in_pts[(size_t)c * d + k]), m.bf16);
        }
        // layer 0 on the FP32 pipe; x1 = sin, c0 = cos of its pre-activation
        float x1[W], c0[W];
#pragma unroll
        for (int o = 0; o < W; ++o) {
            const float4 w4 = *reinterpret_cast<const float4*>(W0s + o * 4);
            float acc = fmaf(w4.x, z[0], 0.f);
            acc = fmaf(w4.y, z[1], acc);
            acc = fmaf(w4.z, z[2], acc);
            const float2 sc = sincos_call(acc);
            x1[o] = rnd_operand(sc.x, m.bf16);
            c0[o] = sc.y;
        }
        // F: pre1 = x1 W1^T
        put_a(x1);
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            mma_ts(d_fx, B1, idesc_fx);
            tc::mma_commit(&bar);
        }
        wait_mma();
        float da[W];                      // pre1, then da1 = d loss / d pre1
        get_d(d_fx, da);
        // d loss / d phi of this pair (slices summed in a fixed order), d quadrature weight, da1
        const float wq = (valid && quad_w) ? quad_w[c] : 1.f;
        float dwq = 0.f;
#pragma unroll
        for (int q = 0; q < W / 4; ++q) {
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && 4 * q < HP)
                for (int sl = 0; sl < nslices; ++sl) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(dphi + ((size_t)sl * P + p) * HP) + q);
                    g4.x += v.x, g4.y += v.y, g4.z += v.z, g4.w += v.w;
                }
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int h = 4 * q + e;
                const float2 sc = sincos_call(da[h]);
                const float gh = h < n2 ? g[e] : 0.f;
                dwq = fmaf(gh, rnd_operand(sc.x, m.bf16), dwq);
                da[h] = wq * gh * sc.y;
            }
        }
        if (valid && dwq_out) dwq_out[p] = dwq;
        else if (valid && d_quad_w) atomicAdd(&d_quad_w[c], dwq);
        // G: dW1 += da1^T x1 (staged tile) and X: dx1 = da1 W1 (TMEM operand), issued together
#pragma unroll
        for (int o = 0; o < W; ++o) put_s(o, da[o]);
#pragma unroll
        for (int k = 0; k < W; ++k) put_s(W + k, x1[k]);
        put_a(da);
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            mma_ss(idesc_g);
            mma_ts(d_fx, B1T, idesc_fx);
            tc::mma_commit(&bar);
        }
        wait_mma();
        if (warp * 16 < W) {               // M = 64 accumulators: rows 16 w .. 16 w + 15 sit in lanes 0..15 of warp w's quadrant
            float t[W];
            get_d(d_gh, t);
#pragma unroll
            for (int k = 0; k < W; ++k) dW1acc[k] += t[k];
        }
        get_d(d_fx, da);                  // dx1
        // H: dW0 += da0^T z
#pragma unroll
        for (int o = 0; o < W; ++o) put_s(o, da[o] * c0[o]);
#pragma unroll
        for (int k = 0; k < 8; ++k) put_s(W + k, k < 3 ? z[k < 3 ? k : 0] : 0.f);
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            mma_ss(idesc_h);
            tc::mma_commit(&bar);
        }
        wait_mma();
        if (warp * 16 < W) {
            uint32_t rr[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7])
                         : "r"(lane_base + d_gh) : "memory");
            tc::wait_ld();
#pragma unroll
            for (int k = 0; k < 8; ++k) dW0acc[k] += __uint_as_float(rr[k]);
        }
        tc::fence_before_sync();          // the next tile's MMAs overwrite what was just read
    }
    // flush: lane l < 16 of warp w owns row 16 w + l
    if (warp * 16 < W && lane < 16 && part != nullptr) {
        // deterministic mode: this CTA's rows of `part` ([2][W*W], same layout as the scalar kernel's)
        const int o = warp * 16 + lane;
        float* row = part + (size_t)blockIdx.x * (2 * W * W);
#pragma unroll
        for (int k = 0; k < 8; ++k) row[o * W + k] = dW0acc[k];
#pragma unroll
        for (int k = 0; k < W; ++k) row[W * W + o * W + k] = dW1acc[k];
    } else if (warp * 16 < W && lane < 16) {
        const int o = warp * 16 + lane;
        if (o < n2)
#pragma unroll
            for (int k = 0; k < W; ++k)
                if (k < n1) atomicAdd(&m.dw[1][o * n1 + k], dW1acc[k]);
        if (o < n1)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < n0) atomicAdd(&m.dw[0][o * n0 + k], dW0acc[k]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

int pick_wmax(const qc_mlp_t* mlp, int d)
{
    int mx = d;
    for (int l = 1; l <= mlp->n_hidden; ++l) mx = mlp->dims[l] > mx ? mlp->dims[l] : mx;
    if (mx <= 4) return 4;
    if (mx <= 8) return 8;
    if (mx <= 16) return 16;
    if (mx <= 32) return 32;
    return -1;
}

int fill_dev(const qc_mlp_t* mlp, int d, MlpDev& m)
{
    if (mlp->n_hidden < 0 || mlp->n_hidden > QC_MAX_MLP_LAYERS) return QC_EUNSUP;
    if (mlp->dims[0] != d) return QC_EINVAL;
    m.n_hidden = mlp->n_hidden;
    m.bf16 = mlp->operand_bf16 != 0;
    for (int l = 0; l <= QC_MAX_MLP_LAYERS; ++l) m.dims[l] = l <= mlp->n_hidden ? mlp->dims[l] : 0;
    for (int l = 0; l < QC_MAX_MLP_LAYERS; ++l) {
        m.w[l] = l < mlp->n_hidden ? mlp->w[l] : nullptr;
        m.dw[l] = nullptr;
    }
    return 0;
}

template <int WMAX>
int launch_fwd(const float* out_pts, const float* in_pts, const float* quad_w, const int* pair_row,
               const int* pair_col, int64_t P, int d, const MlpDev& m, int HP, float* phi, cudaStream_t st)
{
    size_t smem = (size_t)(m.n_hidden > 0 ? m.n_hidden : 1) * WMAX * WMAX * 4;
    QC_CUDA((qc_set_max_smem<k_pair_phi_fwd<WMAX>>((int)smem)));
    int blocks = min(qc_ceil_div(P, 128), qc_num_sms() * 16);
    k_pair_phi_fwd<WMAX><<<blocks, 128, smem, st>>>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi);
    QC_CHECK_LAUNCH();
    return 0;
}

template <int W>
int launch_fwd_tc(const float* out_pts, const float* in_pts, const float* quad_w, const int* pair_row,
                  const int* pair_col, int64_t P, int d, const MlpDev& m, int HP, float* phi, cudaStream_t st)
{
    const size_t smem = ((size_t)W * 4 + (size_t)(m.n_hidden - 1) * 2 * W * W) * 4;
    QC_CUDA((qc_set_max_smem<k_pair_phi_fwd_tc<W>>((int)smem)));
    const int blocks = min(qc_ceil_div(P, 128), qc_num_sms() * 4);     // 4 CTAs per SM share the 512 TMEM columns
    k_pair_phi_fwd_tc<W><<<blocks, 128, smem, st>>>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi);
    QC_CHECK_LAUNCH();
    return 0;
}

// blocks of the backward kernels: one resident wave (the kernels sweep with a grid stride); the deterministic mode's
// workspace has one row per block
inline int bwd_blocks(int64_t P, bool tc) { return min(qc_ceil_div(P, 128), qc_num_sms() * (tc ? 2 : 4)); }

template <int W>
int launch_bwd_tc(const float* out_pts, const float* in_pts, const float* quad_w, const int* pair_row,
                  const int* pair_col, int64_t P, int d, const MlpDev& m, int HP, const float* dphi,
                  int nslices, float* d_quad_w, float* part, float* dwq_out, cudaStream_t st)
{
    const size_t smem = ((size_t)2 * 4 * 64 * 32 + 4 * W * W + W * 4) * 4 + 1024;
    QC_CUDA((qc_set_max_smem<k_pair_phi_bwd_tc<W>>((int)smem)));
    const int blocks = bwd_blocks(P, true);
    k_pair_phi_bwd_tc<W><<<blocks, 128, smem, st>>>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, dphi, nslices, d_quad_w,
                                                    part, dwq_out);
    QC_CHECK_LAUNCH();
    return 0;
}

template <int WMAX, int LT>
int launch_bwd(const float* out_pts, const float* in_pts, const float* quad_w, const int* pair_row,
               const int* pair_col, int64_t P, int d, const MlpDev& m, int HP, const float* dphi,
               int nslices, float* d_quad_w, float* part, float* dwq_out, cudaStream_t st)
{
    constexpr int TL = WMAX >= 16 ? 4 : 2, NQR = 128 / ((WMAX / TL) * (WMAX / TL));
    size_t smem = ((size_t)2 * m.n_hidden * WMAX * WMAX + 2 * 128 * WMAX + NQR * WMAX * WMAX) * 4;
    QC_CUDA((qc_set_max_smem<k_pair_phi_bwd<WMAX, LT>>((int)smem)));
    const int blocks = part ? bwd_blocks(P, false) : min(qc_ceil_div(P, 128), qc_num_sms() * 16);
    k_pair_phi_bwd<WMAX, LT><<<blocks, 128, smem, st>>>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, dphi, nslices, d_quad_w,
                                                        part, dwq_out);
    QC_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" {

int qc_pair_phi_fwd(const float* out_pts, const float* in_pts, const float* quad_w,
                    const int* pair_row, const int* pair_col, int64_t P, int d,
                    const qc_mlp_t* mlp, int HP, float* phi, void* stream)
{
    if (P <= 0) return 0;
    if (d < 1 || d > 3) return QC_EUNSUP;
    MlpDev m;
    int rc = fill_dev(mlp, d, m);
    if (rc) return rc;
    const int H = m.dims[m.n_hidden];
    const int wmax = pick_wmax(mlp, d);
    if (wmax < 0 || HP < H || HP > 32 || (HP & 3)) return QC_EUNSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const int wsel = wmax > HP ? wmax : HP;  // the phi row (HP wide) is written from registers
    // wide MLPs with at least one hidden-to-hidden Linear: those layers run on the tensor cores
    if (wsel >= 16 && m.n_hidden >= 2 && (HP & 3) == 0 && getenv("QC_MLP_NO_TC") == nullptr &&
        (size_t)(m.n_hidden - 1) * 2 * wsel * wsel * 4 <= 48 * 1024) {
        if (wsel == 16) return launch_fwd_tc<16>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
        return launch_fwd_tc<32>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
    }
    switch (wsel) {
        case 4: return launch_fwd<4>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
        case 8: return launch_fwd<8>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
        case 16: return launch_fwd<16>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
        case 32: return launch_fwd<32>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, phi, st);
    }
    return QC_EUNSUP;
}

static bool mlp_bwd_uses_tc(int wsel, int n_hidden) { return wsel >= 16 && n_hidden == 2 && getenv("QC_MLP_NO_TC") == nullptr; }

static int pair_phi_bwd_impl(const float* out_pts, const float* in_pts, const float* quad_w,
                             const int* pair_row, const int* pair_col, int64_t P, int d,
                             const qc_mlp_t* mlp, int HP, const float* dphi, int nslices,
                             float* const* d_mlp_w, float* d_quad_w, float* part, float* dwq_out, int* wsel_out, void* stream)
{
    if (P <= 0) return 0;
    if (d < 1 || d > 3 || nslices < 1) return QC_EUNSUP;
    MlpDev m;
    int rc = fill_dev(mlp, d, m);
    if (rc) return rc;
    for (int l = 0; l < m.n_hidden; ++l) m.dw[l] = d_mlp_w[l];
    const int H = m.dims[m.n_hidden];
    const int wmax = pick_wmax(mlp, d);
    if (wmax < 0 || HP < H || HP > 32 || (HP & 3)) return QC_EUNSUP;
    cudaStream_t st = (cudaStream_t)stream;
    const int wsel = wmax > HP ? wmax : HP;
    if (wsel_out) *wsel_out = wsel;
    // wide MLPs with exactly two hidden layers: the 32 x 32 products run on the tensor cores
    if (mlp_bwd_uses_tc(wsel, m.n_hidden)) {
        if (wsel == 16) return launch_bwd_tc<16>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, dphi, nslices, d_quad_w, part, dwq_out, st);
        return launch_bwd_tc<32>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, dphi, nslices, d_quad_w, part, dwq_out, st);
    }
#define QC_BWD(W, LTV) return launch_bwd<W, LTV>(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, m, HP, dphi, nslices, d_quad_w, part, dwq_out, st)
    const int L = m.n_hidden;
    switch (wsel) {   // register-resident activations when (2L+1)*W floats fit comfortably, else the general kernel
        case 4: if (L == 1) QC_BWD(4, 1); if (L == 2) QC_BWD(4, 2); if (L == 3) QC_BWD(4, 3); if (L == 4) QC_BWD(4, 4); QC_BWD(4, 0);
        case 8: if (L == 1) QC_BWD(8, 1); if (L == 2) QC_BWD(8, 2); if (L == 3) QC_BWD(8, 3); if (L == 4) QC_BWD(8, 4); QC_BWD(8, 0);
        case 16: if (L == 1) QC_BWD(16, 1); if (L == 2) QC_BWD(16, 2); QC_BWD(16, 0);
        case 32: QC_BWD(32, 0);   // 160+ registers per thread: the occupancy loss costs more than the local-memory traffic
    }
#undef QC_BWD
    return QC_EUNSUP;
}

int qc_pair_phi_bwd(const float* out_pts, const float* in_pts, const float* quad_w,
                    const int* pair_row, const int* pair_col, int64_t P, int d,
                    const qc_mlp_t* mlp, int HP, const float* dphi, int nslices,
                    float* const* d_mlp_w, float* d_quad_w, void* stream)
{
    return pair_phi_bwd_impl(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, mlp, HP, dphi, nslices, d_mlp_w, d_quad_w,
                             nullptr, nullptr, nullptr, stream);
}

// workspace layout of the deterministic variant: [blocks][n_hidden][W*W] partial weight gradients, then [P] pair terms
static size_t mlp_ws_rows(int64_t P, int d, const qc_mlp_t* mlp, int HP, int* wsel_out, int* blocks_out)
{
    const int wmax = pick_wmax(mlp, d);
    if (wmax < 0) return 0;
    const int wsel = wmax > HP ? wmax : HP;
    const int blocks = bwd_blocks(P, mlp_bwd_uses_tc(wsel, mlp->n_hidden));
    if (wsel_out) *wsel_out = wsel;
    if (blocks_out) *blocks_out = blocks;
    return (size_t)blocks * (mlp->n_hidden > 0 ? mlp->n_hidden : 1) * wsel * wsel;
}

size_t qc_pair_phi_bwd_workspace_bytes(int64_t P, int d, const qc_mlp_t* mlp, int HP)
{
    if (P <= 0 || mlp == nullptr) return 0;
    return (mlp_ws_rows(P, d, mlp, HP, nullptr, nullptr) + (size_t)P) * sizeof(float);
}

int qc_pair_phi_bwd_ws(const float* out_pts, const float* in_pts, const float* quad_w,
                       const int* pair_row, const int* pair_col, int64_t P, int d,
                       const qc_mlp_t* mlp, int HP, const float* dphi, int nslices,
                       float* const* d_mlp_w, float* d_quad_w, const int* csc_ptr, const int* csc_pair, int Ni,
                       void* ws, size_t ws_bytes, void* stream)
{
    if (P <= 0) return 0;
    if (mlp == nullptr || (d_quad_w != nullptr && (csc_ptr == nullptr || csc_pair == nullptr || Ni <= 0))) return QC_EINVAL;
    int wsel = 0, blocks = 0;
    const size_t rows = mlp_ws_rows(P, d, mlp, HP, &wsel, &blocks);
    if (rows == 0 && mlp->n_hidden > 0) return QC_EUNSUP;
    if (ws == nullptr || ws_bytes < (rows + (size_t)P) * sizeof(float)) return QC_EWORKSPACE;
    float* part = reinterpret_cast<float*>(ws);
    float* dwq = part + rows;
    int rc = pair_phi_bwd_impl(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, mlp, HP, dphi, nslices, d_mlp_w, d_quad_w,
                               part, d_quad_w ? dwq : nullptr, nullptr, stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MlpDev m;
    rc = fill_dev(mlp, d, m);
    if (rc) return rc;
    for (int l = 0; l < m.n_hidden; ++l) m.dw[l] = d_mlp_w[l];
    const int nb_w = qc_ceil_div((int64_t)m.n_hidden * wsel * wsel, 4), nb_q = d_quad_w ? qc_ceil_div(Ni, 128) : 0;
    if (nb_w + nb_q > 0) {
        k_mlp_finish<<<nb_w + nb_q, 128, 0, st>>>(part, blocks, wsel, m, nb_w, dwq, csc_ptr, csc_pair, Ni, d_quad_w);
        QC_CHECK_LAUNCH();
    }
    return 0;
}

}  // extern "C"
