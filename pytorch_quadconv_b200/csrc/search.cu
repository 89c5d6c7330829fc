// search.cu -- K1: grid-binned support-radius neighbour search, bit-exact against
// QuadConv._compute_eval_indices (reference torch_quadconv/quadconv.py:165-199).
//
// The reference forms the dense [No,Ni,d] difference tensor, takes vector_norm over d, compares
// with the fp32 radius and lists the hits with torch.nonzero (row-major).  Here the in-points are
// binned into a uniform grid whose cell edge is >= 1.001 r, so every hit of an out-point lies in
// the 3^d cells around it; one warp per out-point evaluates the SAME fp32 predicate on those
// candidates only, and a rank sort restores the ascending in-index order of torch.nonzero.
//
// HBM-bound integer work: O(N k) instead of O(N^2); coordinates of candidates are read from a
// cell-sorted copy so a warp's loads are contiguous.
#include "common.cuh"

namespace {

struct GridParams {
    double mn[3];
    double inv_h[3];
    int n[3];
    int ncell;
};

struct Ws {
    GridParams* gp;
    int* bbox;        // [6] ordered-int encodings: min[3], max[3]
    int* in_cell;     // [Ni]
    int* cell_count;  // [ncell_max + 1]
    int* cell_start;  // [ncell_max + 1]
    int* cell_cursor; // [ncell_max + 1]
    int* sorted_idx;  // [Ni]
    float* sorted_pts; // [Ni * d]
    int* out_cell;    // [No]
    int* scan_tmp;    // block sums for the scan
    size_t total;
};

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

inline int max_cells_per_dim(int Ni, int d)
{
    double m = pow(4.0 * (double)(Ni > 0 ? Ni : 1), 1.0 / d);
    int cap = d == 1 ? (1 << 20) : (d == 2 ? 2048 : 160);
    int v = (int)m;
    if (v < 1) v = 1;
    if (v > cap) v = cap;
    return v;
}

inline int ncell_max(int Ni, int d)
{
    int m = max_cells_per_dim(Ni, d);
    int64_t c = 1;
    for (int k = 0; k < d; ++k) c *= m;
    return (int)c;
}

inline Ws carve(void* base, int No, int Ni, int d)
{
    Ws w;
    char* p = (char*)base;
    size_t off = 0;
    int nc = ncell_max(Ni, d);
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes); return r; };
    w.gp = (GridParams*)take(sizeof(GridParams));
    w.bbox = (int*)take(6 * sizeof(int));
    w.in_cell = (int*)take((size_t)Ni * 4);
    w.cell_count = (int*)take((size_t)(nc + 1) * 4);
    w.cell_start = (int*)take((size_t)(nc + 1) * 4);
    w.cell_cursor = (int*)take((size_t)(nc + 1) * 4);
    w.sorted_idx = (int*)take((size_t)Ni * 4);
    w.sorted_pts = (float*)take((size_t)Ni * d * 4);
    w.out_cell = (int*)take((size_t)No * 4);
    w.scan_tmp = (int*)take((size_t)(1 << 17) * 4 + 1024);
    w.total = off;
    return w;
}

// order-preserving float <-> int encoding for atomicMin/atomicMax
__device__ __forceinline__ int f2ord(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_bbox_init(int* bbox)
{
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) bbox[threadIdx.x] = (int)0x80000000;
}

__global__ void k_bbox(const float* __restrict__ pts, int N, int d, int* bbox)
{
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        for (int k = 0; k < d; ++k) {
            float v = pts[(size_t)i * d + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    for (int k = 0; k < d; ++k) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], m));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], m));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&bbox[k], f2ord(mn[k]));
            atomicMax(&bbox[3 + k], f2ord(mx[k]));
        }
    }
}

// One thread: choose the grid.  Cell edge h_k = extent_k / n_k >= 1.001 * r, n_k <= maxc.
__global__ void k_grid_setup(const int* bbox, int d, float r32, int maxc, GridParams* gp)
{
    int ncell = 1;
    for (int k = 0; k < 3; ++k) {
        gp->mn[k] = 0.0;
        gp->inv_h[k] = 0.0;
        gp->n[k] = 1;
    }
    for (int k = 0; k < d; ++k) {
        double mn = (double)ord2f(bbox[k]), mx = (double)ord2f(bbox[3 + k]);
        double ext = mx - mn;
        double hmin = (double)r32 * 1.001 + 1e-30;
        int n = 1;
        if (ext > 0.0) {
            double q = ext / hmin;
            n = q >= (double)maxc ? maxc : (int)q;
            if (n < 1) n = 1;
        }
        gp->mn[k] = mn;
        gp->n[k] = n;
        gp->inv_h[k] = ext > 0.0 ? (double)n / ext : 0.0;
        ncell *= n;
    }
    gp->ncell = ncell;
}

// integer cell coordinate (unclamped) of a point along dimension k
__device__ __forceinline__ int cell_coord(const GridParams& gp, float x, int k)
{
    double t = ((double)x - gp.mn[k]) * gp.inv_h[k];
    // clamp to a safe integer range before the conversion
    t = fmin(fmax(t, -2.0), (double)gp.n[k] + 1.0);
    return (int)floor(t);
}

__global__ void k_cell_zero(int* a, int* b, const GridParams* gp)
{
    int n = gp->ncell + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        a[i] = 0;
        b[i] = 0;
    }
}

__global__ void k_cell_assign(const float* __restrict__ pts, int N, int d, const GridParams* gpp,
                              int* __restrict__ cell_of, int* cell_count)
{
    const GridParams gp = *gpp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int c = 0, mul = 1;
        for (int k = 0; k < d; ++k) {
            int ck = cell_coord(gp, pts[(size_t)i * d + k], k);
            ck = min(max(ck, 0), gp.n[k] - 1);
            c += ck * mul;
            mul *= gp.n[k];
        }
        cell_of[i] = c;
        atomicAdd(&cell_count[c], 1);
    }
}

__global__ void k_cell_fill(const float* __restrict__ pts, int N, int d,
                            const int* __restrict__ cell_of, const int* __restrict__ cell_start,
                            int* cursor, int* __restrict__ sorted_idx, float* __restrict__ sorted_pts)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int c = cell_of[i];
        int pos = cell_start[c] + atomicAdd(&cursor[c], 1);
        sorted_idx[pos] = i;
        if (sorted_pts)
            for (int k = 0; k < d; ++k) sorted_pts[(size_t)pos * d + k] = pts[(size_t)i * d + k];
    }
}

// ---- exclusive scan (three small kernels) ---------------------------------------------------
constexpr int SCAN_BLOCK = 1024;

__global__ void k_scan_block_sums(const int* __restrict__ in, int n, const int* n_dev, int* block_sums)
{
    if (n_dev) n = *n_dev;
    __shared__ int red[32];
    int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    int v = i < n ? in[i] : 0;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = red[threadIdx.x];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
    }
}

// single block: exclusive scan of nb block sums in place (nb <= 2^17)
__global__ void k_scan_sums(int* block_sums, int nb)
{
    __shared__ int part[SCAN_BLOCK];
    int per = (nb + SCAN_BLOCK - 1) / SCAN_BLOCK;
    int beg = threadIdx.x * per, end = min(beg + per, nb);
    int s = 0;
    for (int i = beg; i < end; ++i) s += block_sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int off = 1; off < SCAN_BLOCK; off <<= 1) {
        int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = threadIdx.x == 0 ? 0 : part[threadIdx.x - 1];
    for (int i = beg; i < end; ++i) {
        int v = block_sums[i];
        block_sums[i] = run;
        run += v;
    }
}

__global__ void k_scan_final(const int* __restrict__ in, int n, const int* n_dev,
                             const int* __restrict__ block_sums, int* __restrict__ out)
{
    if (n_dev) n = *n_dev;
    __shared__ int wsum[32];
    int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    int v = i < n ? in[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = wsum[lane];
        int sinc = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, sinc, off);
            if (lane >= off) sinc += t;
        }
        wsum[lane] = sinc - s;
    }
    __syncthreads();
    int excl = block_sums[blockIdx.x] + wsum[w] + inc - v;
    if (i < n) out[i] = excl;
    if (i == n - 1) out[n] = excl + v;
    if (n == 0 && i == 0) out[0] = 0;
}

int scan_launch(const int* in, int n, const int* n_dev, int* out, int* tmp, cudaStream_t st)
{
    // n is an upper bound used for the launch; n_dev (optional) holds the true length on device
    int nb = qc_ceil_div(n > 0 ? n : 1, SCAN_BLOCK);
    if (nb > (1 << 17)) return QC_EUNSUP;
    k_scan_block_sums<<<nb, SCAN_BLOCK, 0, st>>>(in, n, n_dev, tmp);
    k_scan_sums<<<1, SCAN_BLOCK, 0, st>>>(tmp, nb);
    k_scan_final<<<nb, SCAN_BLOCK, 0, st>>>(in, n, n_dev, tmp, out);
    QC_CHECK_LAUNCH();
    return 0;
}

// ---- the fp32 predicate, restating quadconv.py:178-182 ---------------------------------------
template <int D>
__device__ __forceinline__ bool qc_hit(const float* __restrict__ o, const float* __restrict__ p, float r32)
{
    if (D == 1) return fabsf(__fsub_rn(o[0], p[0])) <= r32;
    float z0 = __fsub_rn(o[0], p[0]);
    float acc = __fmul_rn(z0, z0);
    float z1 = __fsub_rn(o[1], p[1]);
    acc = __fmaf_rn(z1, z1, acc);
    if (D == 3) {
        float z2 = __fsub_rn(o[2], p[2]);
        acc = __fmaf_rn(z2, z2, acc);
    }
    return __fsqrt_rn(acc) <= r32;
}

// One warp per out-point.  FILL=false: count hits.  FILL=true: write the hits (unsorted) into
// pair_row[row_ptr[o]..) as scratch, rank-sort them into pair_col / eval_indices, then set pair_row.
template <int D, bool FILL>
__global__ void __launch_bounds__(256)
k_search(const float* __restrict__ out_pts, int No, float r32, const GridParams* gpp,
         const int* __restrict__ cell_start, const int* __restrict__ sorted_idx,
         const float* __restrict__ sorted_pts, int* __restrict__ row_counts,
         const int* __restrict__ row_ptr, int64_t* __restrict__ eval_indices,
         int* __restrict__ pair_row, int* __restrict__ pair_col)
{
    const GridParams gp = *gpp;
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < No; o += warps_per_grid) {
        float po[3] = {0.f, 0.f, 0.f};
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        bool empty = false;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            po[k] = out_pts[(size_t)o * D + k];
            int c = cell_coord(gp, po[k], k);
            lo[k] = max(c - 1, 0);
            hi[k] = min(c + 1, gp.n[k] - 1);
            empty |= lo[k] > hi[k];
        }
        int count = 0;
        const int base = FILL ? row_ptr[o] : 0;
        if (!empty) {
            for (int cz = lo[2]; cz <= hi[2]; ++cz)
                for (int cy = lo[1]; cy <= hi[1]; ++cy) {
                    int crow = (cz * gp.n[1] + cy) * gp.n[0];
                    int beg = cell_start[crow + lo[0]], end = cell_start[crow + hi[0] + 1];
                    for (int s = beg; s < end; s += 32) {
                        int q = s + lane;
                        bool hit = false;
                        if (q < end) hit = qc_hit<D>(po, sorted_pts + (size_t)q * D, r32);
                        unsigned m = __ballot_sync(0xffffffffu, hit);
                        if (FILL && hit) pair_row[base + count + __popc(m & ((1u << lane) - 1))] = sorted_idx[q];
                        count += __popc(m);
                    }
                }
        }
        if (!FILL) {
            if (lane == 0) row_counts[o] = count;
        } else {
            __syncwarp();
            // rank sort: in-indices within a row are distinct
            for (int e = lane; e < count; e += 32) {
                int v = pair_row[base + e];
                int rank = 0;
                for (int f = 0; f < count; ++f) rank += pair_row[base + f] < v;
                pair_col[base + rank] = v;
                eval_indices[2 * (size_t)(base + rank)] = o;
                eval_indices[2 * (size_t)(base + rank) + 1] = v;
            }
            __syncwarp();
            for (int e = lane; e < count; e += 32) pair_row[base + e] = o;
        }
    }
}

__global__ void k_order_identity_from_sorted(const int* __restrict__ sorted_idx, int N, int* __restrict__ order)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) order[i] = sorted_idx[i];
}

__global__ void k_pairs_to_csr(const int64_t* __restrict__ ev, int64_t P, int No, int Ni,
                               int* row_counts, int* __restrict__ pair_row, int* __restrict__ pair_col, int* err)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = ev[2 * p], i = ev[2 * p + 1];
        bool bad = o < 0 || o >= No || i < 0 || i >= Ni;
        if (p > 0 && ev[2 * (p - 1)] > o) bad = true;
        if (bad) {
            *err = 1;
            o = 0;
            i = 0;
        }
        pair_row[p] = (int)o;
        pair_col[p] = (int)i;
        atomicAdd(&row_counts[o], 1);
    }
}

__global__ void k_csc_count(const int* __restrict__ pair_col, int64_t P, int* csc_counts)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&csc_counts[pair_col[p]], 1);
}

__global__ void k_csc_scatter(const int* __restrict__ pair_col, int64_t P, const int* __restrict__ csc_ptr,
                              int* cursor, int* __restrict__ csc_pair)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        int c = pair_col[p];
        csc_pair[csc_ptr[c] + atomicAdd(&cursor[c], 1)] = (int)p;
    }
}

// one warp per in-point: sort its pair ids ascending (rank sort through csc_row as scratch)
__global__ void __launch_bounds__(256)
k_csc_sort(const int* __restrict__ csc_ptr, int Ni, const int* __restrict__ pair_row,
           int* __restrict__ csc_pair, int* __restrict__ csc_row)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < Ni; c += warps_per_grid) {
        int beg = csc_ptr[c], n = csc_ptr[c + 1] - beg;
        for (int e = lane; e < n; e += 32) csc_row[beg + e] = csc_pair[beg + e];
        __syncwarp();
        for (int e = lane; e < n; e += 32) {
            int v = csc_row[beg + e];
            int rank = 0;
            for (int f = 0; f < n; ++f) rank += csc_row[beg + f] < v;
            csc_pair[beg + rank] = v;
        }
        __syncwarp();
        for (int e = lane; e < n; e += 32) csc_row[beg + e] = pair_row[csc_pair[beg + e]];
        __syncwarp();
    }
}

int bin_points(const float* pts, int N, int d, Ws& w, int* cell_of, float* sorted_pts, int* sorted_idx, cudaStream_t st)
{
    int nc = 0;  // upper bound for launches
    (void)nc;
    int blocks = min(qc_ceil_div(N > 0 ? N : 1, 256), qc_num_sms() * 8);
    k_cell_zero<<<qc_num_sms() * 4, 256, 0, st>>>(w.cell_count, w.cell_cursor, w.gp);
    k_cell_assign<<<blocks, 256, 0, st>>>(pts, N, d, w.gp, cell_of, w.cell_count);
    QC_CHECK_LAUNCH();
    return 0;
}


// Reorder the processing order of points inside windows of BAL_W consecutive entries by segment length
// (ascending, ties by position: deterministic).  The order stays spatially compact (a window of the
// cell-sorted order), but the four warps of a row block and the CTAs of a wave now work on segments of similar
// length -- measured at C3: step 11.59 -> 11.39 ms.  One block per window, O(W^2) rank sort in shared memory
// (a one-time cost of the cached tables).
constexpr int BAL_W = 2048;
__global__ void __launch_bounds__(256)
k_balance_order(const int* __restrict__ order, const int* __restrict__ seg_ptr, int n, int* __restrict__ out)
{
    __shared__ int len_s[BAL_W];
    const int w0 = blockIdx.x * BAL_W;
    const int cnt = min(BAL_W, n - w0);
    for (int i = threadIdx.x; i < cnt; i += 256) {
        const int p = order ? order[w0 + i] : w0 + i;
        len_s[i] = seg_ptr[p + 1] - seg_ptr[p];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += 256) {
        const int li = len_s[i];
        int rank = 0;
        for (int j = 0; j < cnt; ++j) {
            const int lj = len_s[j];
            rank += (lj < li || (lj == li && j < i)) ? 1 : 0;
        }
        out[w0 + rank] = order ? order[w0 + i] : w0 + i;
    }
}

}  // namespace

extern "C" {

size_t qc_search_workspace_bytes(int No, int Ni, int d)
{
    if (d < 1 || d > 3 || No < 0 || Ni < 0) return 0;
    Ws w = carve(nullptr, No, Ni, d);
    return w.total;
}

int qc_exclusive_scan_i32(const int* in, int n, int* out, void* stream)
{
    if (n < 0) return QC_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    int nb = qc_ceil_div(n > 0 ? n : 1, SCAN_BLOCK);
    int* tmp = nullptr;
    // block sums live in a stream-ordered temporary
    QC_CUDA(cudaMallocAsync((void**)&tmp, (size_t)(nb + 1) * 4, st));
    int rc = scan_launch(in, n, nullptr, out, tmp, st);
    cudaFreeAsync(tmp, st);
    return rc;
}

int qc_search_count(const float* out_pts, const float* in_pts, int No, int Ni, int d, float r32,
                    void* ws, size_t ws_bytes, int* row_counts, int* order_out, int* order_in,
                    void* stream)
{
    if (d < 1 || d > 3 || No <= 0 || Ni <= 0 || !(r32 >= 0.f)) return QC_EINVAL;
    Ws w = carve(ws, No, Ni, d);
    if (ws_bytes < w.total) return QC_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = qc_num_sms();
    const int maxc = max_cells_per_dim(Ni, d);
    const int nc_max = ncell_max(Ni, d);

    k_bbox_init<<<1, 32, 0, st>>>(w.bbox);
    k_bbox<<<min(qc_ceil_div(Ni, 256), sms * 4), 256, 0, st>>>(in_pts, Ni, d, w.bbox);
    k_grid_setup<<<1, 1, 0, st>>>(w.bbox, d, r32, maxc, w.gp);

    // bin the in-points
    int rc = bin_points(in_pts, Ni, d, w, w.in_cell, nullptr, nullptr, st);
    if (rc) return rc;
    rc = scan_launch(w.cell_count, nc_max, &w.gp->ncell, w.cell_start, w.scan_tmp, st);
    if (rc) return rc;
    k_cell_fill<<<min(qc_ceil_div(Ni, 256), sms * 8), 256, 0, st>>>(in_pts, Ni, d, w.in_cell, w.cell_start,
                                                                  w.cell_cursor, w.sorted_idx, w.sorted_pts);
    if (order_in) k_order_identity_from_sorted<<<min(qc_ceil_div(Ni, 256), sms * 8), 256, 0, st>>>(w.sorted_idx, Ni, order_in);

    // count hits, one warp per out-point
    int blocks = min(qc_ceil_div((int64_t)No * 32, 256), sms * 16);
    if (d == 1)
        k_search<1, false><<<blocks, 256, 0, st>>>(out_pts, No, r32, w.gp, w.cell_start, w.sorted_idx, w.sorted_pts, row_counts, nullptr, nullptr, nullptr, nullptr);
    else if (d == 2)
        k_search<2, false><<<blocks, 256, 0, st>>>(out_pts, No, r32, w.gp, w.cell_start, w.sorted_idx, w.sorted_pts, row_counts, nullptr, nullptr, nullptr, nullptr);
    else
        k_search<3, false><<<blocks, 256, 0, st>>>(out_pts, No, r32, w.gp, w.cell_start, w.sorted_idx, w.sorted_pts, row_counts, nullptr, nullptr, nullptr, nullptr);
    QC_CHECK_LAUNCH();

    // processing order of the out-points: cell-sorted with the same grid.  The in-point cell
    // tables must survive for qc_search_fill, so this uses its own count/start/cursor arrays
    // carved after the workspace proper (order_out is produced with a second, small binning pass
    // into scratch that aliases nothing used by the fill).
    if (order_out) {
        if (out_pts == in_pts && No == Ni) {
            k_order_identity_from_sorted<<<min(qc_ceil_div(Ni, 256), sms * 8), 256, 0, st>>>(w.sorted_idx, Ni, order_out);
        } else {
            // temporaries: counts/start/cursor for the out-points
            int* tmp = nullptr;
            size_t n1 = (size_t)nc_max + 1;
            QC_CUDA(cudaMallocAsync((void**)&tmp, (3 * n1 + (1 << 17) + 256) * 4, st));
            int *cnt = tmp, *start = tmp + n1, *cur = tmp + 2 * n1, *stmp = tmp + 3 * n1;
            k_cell_zero<<<sms * 4, 256, 0, st>>>(cnt, cur, w.gp);
            k_cell_assign<<<min(qc_ceil_div(No, 256), sms * 8), 256, 0, st>>>(out_pts, No, d, w.gp, w.out_cell, cnt);
            rc = scan_launch(cnt, nc_max, &w.gp->ncell, start, stmp, st);
            if (rc == 0)
                k_cell_fill<<<min(qc_ceil_div(No, 256), sms * 8), 256, 0, st>>>(out_pts, No, d, w.out_cell, start, cur, order_out, nullptr);
            cudaFreeAsync(tmp, st);
            if (rc) return rc;
        }
    }
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_search_fill(const float* out_pts, const float* in_pts, int No, int Ni, int d, float r32,
                   void* ws, size_t ws_bytes, const int* row_ptr, int64_t* eval_indices,
                   int* pair_row, int* pair_col, void* stream)
{
    (void)in_pts;
    if (d < 1 || d > 3 || No <= 0 || Ni <= 0) return QC_EINVAL;
    Ws w = carve(ws, No, Ni, d);
    if (ws_bytes < w.total) return QC_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = min(qc_ceil_div((int64_t)No * 32, 256), qc_num_sms() * 16);
    if (d == 1)
        k_search<1, true><<<blocks, 256, 0, st>>>(out_pts, No, r32, w.gp, w.cell_start, w.sorted_idx, w.sorted_pts, nullptr, row_ptr, eval_indices, pair_row, pair_col);
    else if (d == 2)
        k_search<2, true><<<blocks, 256, 0, st>>>(out_pts, No, r32, w.gp, w.cell_start, w.sorted_idx, w.sorted_pts, nullptr, row_ptr, eval_indices, pair_row, pair_col);
    else
        k_search<3, true><<<blocks, 256, 0, st>>>(out_pts, No, r32, w.gp, w.cell_start, w.sorted_idx, w.sorted_pts, nullptr, row_ptr, eval_indices, pair_row, pair_col);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_pairs_to_csr(const int64_t* eval_indices, int64_t P, int No, int Ni, int* row_counts,
                    int* pair_row, int* pair_col, int* err_flag, void* stream)
{
    if (P < 0 || No <= 0 || Ni <= 0 || P >= ((int64_t)1 << 31)) return QC_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    QC_CUDA(cudaMemsetAsync(row_counts, 0, (size_t)No * 4, st));
    QC_CUDA(cudaMemsetAsync(err_flag, 0, 4, st));
    if (P > 0) k_pairs_to_csr<<<min(qc_ceil_div(P, 256), qc_num_sms() * 8), 256, 0, st>>>(eval_indices, P, No, Ni, row_counts, pair_row, pair_col, err_flag);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_csc_count(const int* pair_col, int64_t P, int Ni, int* csc_counts, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    QC_CUDA(cudaMemsetAsync(csc_counts, 0, (size_t)Ni * 4, st));
    if (P > 0) k_csc_count<<<min(qc_ceil_div(P, 256), qc_num_sms() * 8), 256, 0, st>>>(pair_col, P, csc_counts);
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_csc_fill(const int* pair_row, const int* pair_col, int64_t P, int Ni, const int* csc_ptr,
                int* cursor, int* csc_pair, int* csc_row, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    QC_CUDA(cudaMemsetAsync(cursor, 0, (size_t)Ni * 4, st));
    if (P > 0) {
        k_csc_scatter<<<min(qc_ceil_div(P, 256), qc_num_sms() * 8), 256, 0, st>>>(pair_col, P, csc_ptr, cursor, csc_pair);
        k_csc_sort<<<min(qc_ceil_div((int64_t)Ni * 32, 256), qc_num_sms() * 16), 256, 0, st>>>(csc_ptr, Ni, pair_row, csc_pair, csc_row);
    }
    QC_CHECK_LAUNCH();
    return 0;
}

int qc_balance_order(const int* order, const int* seg_ptr, int n, int* out, void* stream)
{
    if (n <= 0) return 0;
    k_balance_order<<<qc_ceil_div(n, BAL_W), 256, 0, (cudaStream_t)stream>>>(order, seg_ptr, n, out);
    QC_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
