/*
 * quadconv_b200.h -- C ABI of libquadconv_b200.so (hand-written sm_100a CUDA; no torch types).
 *
 * The reference (AlgorithmicDataReduction/PyTorch-QuadConv) has no FFI layer: its hot path is ATen
 * calls made from torch_quadconv/quadconv.py and torch_quadconv/mesh_pool.py.  Each entry point
 * below replaces the ATen composition at the cited reference lines; the Python nn.Module shells
 * in pytorch_quadconv_b200/ (same constructor/forward signatures as the reference) bind them with
 * ctypes (pytorch_quadconv_b200/_cabi.py).  INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocates), except where a
 *     parameter is documented as host.  Steady-state entry points allocate nothing; two setup-time exceptions use
 *     stream-ordered scratch (cudaMallocAsync/cudaFreeAsync on the caller's stream): the scan / sort scratch of
 *     qc_search_* and qc_balance_order, and qc_contract's dW mode when it is handed no workspace
 *     (qc_contract_workspace_bytes; the modules always pass one, so captured CUDA graphs never allocate);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 on success, a positive cudaError_t, or a negative QC_E* code; no exceptions;
 *   - re-entrant, no global state apart from per-process caches of launch-invariant host state (SM count,
 *     cudaFuncSetAttribute done once per kernel and device, environment switches read once).
 *
 * Internal data layouts (DESIGN.md "Data layout in HBM")
 *   packed features  Xp[nblk][N][CP4/4][Bp][4] fp32; CP4 = C rounded up to 4, Bp = min(32, pow2ceil(B)) lanes,
 *                    nblk = ceil(B/32) batch blocks, padded batch lanes are zero.
 *   CSR              seg_ptr[Ndst+1] int32, src_idx[P] int32 (pairs grouped by destination point)
 *   phi              [P][HP] fp32: last hidden activation of the filter MLP per pair, times the
 *                    quadrature weight of the pair's input point; HP = H rounded up to {4,8,16,32}.
 */
#ifndef QUADCONV_B200_H
#define QUADCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QC_EINVAL (-1)   /* bad argument */
#define QC_EUNSUP (-2)   /* shape outside what the kernels implement (stated in DESIGN.md) */
#define QC_EWORKSPACE (-3)

#define QC_MAX_MLP_LAYERS 8

int qc_version(void);

/* ------------------------------------------------------------------------------------------
 * K1  support-radius neighbour search.
 * Replaces QuadConv._compute_eval_indices, torch_quadconv/quadconv.py:165-199 (dense [No,Ni,d]
 * subtract + vector_norm (:153-154) + `<=` + torch.nonzero).  Bit-exact pair SET and ORDER:
 * predicate sqrtf(fmaf(z2,z2,fmaf(z1,z1,z0*z0))) <= r32 on fp32 differences, rows ascending in
 * the out index then the in index.  d in {1,2,3}.
 *
 * Two calls, because the number of pairs is data dependent (like torch.nonzero):
 *   qc_search_count : bins the in-points into a uniform grid (cell edge >= r), counts the hits of
 *                     every out-point -> row_counts[No]; also emits the cell-sorted processing
 *                     orders order_out[No], order_in[Ni] used by the fused kernels for locality.
 *   (caller: exclusive scan -> row_ptr[No+1] with qc_exclusive_scan_i32, reads P = row_ptr[No])
 *   qc_search_fill  : writes eval_indices[P,2] int64 (reference layout), pair_row[P], pair_col[P]
 *                     int32.
 * `ws` must hold qc_search_workspace_bytes(No, Ni, d) bytes and be passed unchanged to both.
 * ------------------------------------------------------------------------------------------ */
size_t qc_search_workspace_bytes(int No, int Ni, int d);
int qc_search_count(const float* out_pts, const float* in_pts, int No, int Ni, int d, float r32,
                    void* ws, size_t ws_bytes, int* row_counts, int* order_out, int* order_in,
                    void* stream);
int qc_search_fill(const float* out_pts, const float* in_pts, int No, int Ni, int d, float r32,
                   void* ws, size_t ws_bytes, const int* row_ptr, int64_t* eval_indices,
                   int* pair_row, int* pair_col, void* stream);

/* exclusive prefix sum, n <= 2^27; out has n+1 entries (out[n] = total). */
int qc_exclusive_scan_i32(const int* in, int n, int* out, void* stream);

/* Pair list given by the caller (a cached / loaded eval_indices Parameter, quadconv.py:186-188,
 * :240-242): rebuild pair_row/pair_col/row_counts.  err_flag (device int) is set non-zero if the
 * rows are not ascending in the out index or an index is out of range. */
int qc_pairs_to_csr(const int64_t* eval_indices, int64_t P, int No, int Ni, int* row_counts,
                    int* pair_row, int* pair_col, int* err_flag, void* stream);

/* Transposed (CSC) view: for every in-point the pairs that read it, ascending pair id.
 *   csc_counts[Ni] (out, zeroed by the call) ; caller scans -> csc_ptr[Ni+1] ;
 *   qc_csc_fill writes csc_pair[P] (pair id) and csc_row[P] (out index of that pair).
 * `cursor` is an int[Ni] scratch. Needed by the backward pass w.r.t. the input features, whose
 * reduction runs over pairs grouped by INPUT point (autograd of quadconv.py:219, index gather). */
int qc_csc_count(const int* pair_col, int64_t P, int Ni, int* csc_counts, void* stream);
int qc_csc_fill(const int* pair_row, const int* pair_col, int64_t P, int Ni, const int* csc_ptr,
                int* cursor, int* csc_pair, int* csc_row, void* stream);
/* Reorder a processing order (NULL = identity) inside windows of 2048 consecutive entries by segment length
 * (seg_ptr[p+1] - seg_ptr[p], ascending, ties by position) so that the warps that share a CTA work on segments
 * of similar length; `out` must not alias `order`. */
int qc_balance_order(const int* order, const int* seg_ptr, int n, int* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout conversion between the reference's [B,C,N] tensors (quadconv.py:238 docstring) and the
 * packed [nblk][N][CP4/4][Bp][4] layout.  qc_unpack_features optionally adds bias[C][N]
 * (quadconv.py:262-263) and/or accumulates the per-(c,n) batch sum of the source (d bias).
 * ------------------------------------------------------------------------------------------ */
int qc_pack_features(const float* x, int B, int C, int N, float* xp, void* stream);
int qc_unpack_features(const float* xp, int B, int C, int N, const float* bias, float* x,
                       void* stream);
int qc_batch_sum(const float* x, int B, int C, int N, float* out /*[C][N]*/, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2a  hidden layers of the filter MLP per pair.
 * Replaces, per pair p: weights[idx1] gather (quadconv.py:247), eval_locs (quadconv.py:250-253),
 * and every layer of self.filter except the last Linear (quadconv.py:130-145, utils/misc.py:18-19).
 *   phi[p][h] = w[col[p]] * sin(W_L-1 ... sin(W_0 (out_pts[row[p]] - in_pts[col[p]])))[h]
 * mlp_w[l] : device pointer to filter.{2l}.weight, row-major [dims[l+1]][dims[l]], l < n_hidden.
 * dims[0] = d, dims[1..n_hidden] = filter_seq; every dim <= 32.  n_hidden may be 0 (phi = w * z).
 * quad_w may be NULL (weights of one).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int n_hidden;
    int dims[QC_MAX_MLP_LAYERS + 1];
    const float* w[QC_MAX_MLP_LAYERS];
    int operand_bf16;   /* != 0: every Linear of the filter MLP takes bf16 operands (weights as loaded and layer
                           inputs rounded to nearest even), fp32 accumulate; phi is w * bf16(last activation) */
} qc_mlp_t;

int qc_pair_phi_fwd(const float* out_pts, const float* in_pts, const float* quad_w,
                    const int* pair_row, const int* pair_col, int64_t P, int d,
                    const qc_mlp_t* mlp /*host*/, int HP, float* phi, void* stream);

/* Backward of the above.  dphi[nslices][P][HP] holds partial d loss / d phi (summed over the
 * slices here, in fixed order).  Accumulates (+=) into d_mlp_w[l] (same shapes as mlp_w[l]) and
 * d_quad_w[Ni] (may be NULL). */
int qc_pair_phi_bwd(const float* out_pts, const float* in_pts, const float* quad_w,
                    const int* pair_row, const int* pair_col, int64_t P, int d,
                    const qc_mlp_t* mlp /*host*/, int HP, const float* dphi, int nslices,
                    float* const* d_mlp_w /*host array of device ptrs*/, float* d_quad_w,
                    void* stream);

/* Bit-reproducible variants (the plain ones combine per-block partial sums with float atomics, whose order varies run
 * to run): per-block partial sums and per-pair terms go to a caller-owned workspace and are added in fixed orders
 * (block order; the CSC order of each in-point for d_quad_w -- csc_ptr [Ni+1], csc_pair [P] from qc_csc_*). */
size_t qc_pair_phi_bwd_workspace_bytes(int64_t P, int d, const qc_mlp_t* mlp, int HP);
int qc_pair_phi_bwd_ws(const float* out_pts, const float* in_pts, const float* quad_w,
                       const int* pair_row, const int* pair_col, int64_t P, int d,
                       const qc_mlp_t* mlp, int HP, const float* dphi, int nslices,
                       float* const* d_mlp_w, float* d_quad_w, const int* csc_ptr, const int* csc_pair, int Ni,
                       void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2b/K3  fused last-MLP-layer + gather-contract-segmented-reduce.
 * Replaces the last Linear of self.filter + reshape (quadconv.py:143,:91), the feature gather,
 * einsum('n,nij,bin->bjn') and scatter_add_ of QuadConv._integrate (quadconv.py:213-227); the
 * [P,Ci,Co] filter tensor, the gathered [B,Ci,P] features and the [B,Co,P] values never exist.
 * Factorised (SURVEY.md hard-part 3b): for a destination point t with segment S(t)
 *     T[t; b, c, h] = sum_{q in S(t)} phi[phi_idx[q]][h] * Xp[src_idx[q]][c][b]       (stage B)
 *     Yp[t][j][b]   = sum_{c,h} T[t; b, c, h] * W2[(c*H + h)][j]                        (stage C)
 * Forward:  dst = out points, src = in points, W2[(i*H+h)][j] = W_last[i*Co+j][h].
 * d/d features: dst = in points (CSC), src = out points, X = packed grad_out,
 *               W2[(j*H+h)][i] = W_last[i*Co+j][h]; with Zp = packed features the same pass
 *               accumulates dW2[(c*H+h)][j] += sum_{t,b} T[t;b,c,h] * Zp[t][j][b]  (= d W_last).
 * W2 / dW2 have row stride CyP = Cy rounded up to 4 (padding columns zero / ignored).
 * order: processing order of destination points (NULL = identity).  phi_idx NULL = identity.
 * B is the true batch size (Bp / nblk are derived as in the layout note above).
 * ------------------------------------------------------------------------------------------ */
int qc_contract(const float* Xp, int Nsrc, int Cx, const int* seg_ptr, const int* src_idx,
                const int* phi_idx, const float* phi, int H, int HP, const float* W2, int Cy,
                const int* order, int Ndst, int B, float* Yp, const float* Zp, float* dW2,
                void* workspace, size_t workspace_bytes, void* stream);
/* Workspace of the dW2 variant (Zp / dW2 non-NULL): the tensor-core kernel writes partial accumulators there
 * (tcgen05 accumulates with truncation, so its accumulation chains are kept short).  0 = none needed.  A NULL
 * or too small workspace makes qc_contract use a stream-ordered allocation instead (slower, not for graphs). */
size_t qc_contract_workspace_bytes(int Cx, int H, int HP, int Cy, int Ndst, int B);

/* d loss / d phi (grouped by out point):
 *     dT[t; b, c, h] = sum_j Gp[t][j][b] * W3[j][c][h]          W3[j][c][h] = W_last[c*Co+j][h]
 *     dphi[q][h]    += sum_{b,c} Xp[src_idx[q]][c][b] * dT[t; b, c, h]       (q in S(t))
 * W3 is [Cg][Cx][HP] (h padded with zeros).  dphi is [qc_dphi_num_slices(...)][P][HP]: the fast
 * kernels write one slice per (batch block, channel chunk) with plain stores; the slices are
 * summed by qc_pair_phi_bwd.  No initialisation needed by the caller. */
int qc_dphi_num_slices(int Nsrc, int Cx, int Cg, int HP, int B);
int qc_dphi(const float* Xp, int Nsrc, int Cx, const float* Gp, int Cg, const int* seg_ptr,
            const int* src_idx, const float* W3, int H, int HP, const int* order, int Ndst, int B,
            int64_t P, float* dphi, void* stream);

/* Wide layers (last hidden width H > 8, or more than 32 channels on the dense side): the same
 * factorised evaluation of quadconv.py:213-227 with the intermediate written to HBM once, so that the
 * dense stage is a plain row-major GEMM (qc_gemm_nt_tf32x3 / qc_gemm_tn_tf32x3 below, or any fp32 GEMM of the caller):
 *     T[b][t][c*H + h]   = sum_{q in S(t)} phi[phi_idx[q]][h] * Xp[src_idx[q]][c][b]     (accumulate)
 *     Y[b][t][:]         = T[b][t][:] * W2            W2[(c*H+h)][j] = W_last[c*Co+j][h]
 * and, backward, dT = G * W2^T followed by
 *     dphi[s][q][h]      = sum_{b,c in slice s} Xp[src_idx[q]][c][b] * dT[b][t][c*H + h]  (dot)
 * T / dT are [B][Ndst][Cx*H] floats, unpadded.  dphi is [qc_gather_dot_num_slices][P][HP] (one slice per
 * block of 32 batch entries, no initialisation needed; summed by qc_pair_phi_bwd).  qc_gather_supported returns 1 when the shape is handled. */
int qc_gather_supported(int Cx, int H, int HP, int B);
int qc_gather_dot_num_slices(int Cx, int HP, int B);
int qc_gather_accumulate(const float* Xp, int Nsrc, int Cx, const int* seg_ptr, const int* src_idx,
                         const int* phi_idx, const float* phi, int H, int HP, const int* order, int Ndst,
                         int B, float* T, void* stream);
int qc_gather_dot(const float* Xp, int Nsrc, int Cx, const float* dT, const int* seg_ptr, const int* src_idx,
                  int H, int HP, const int* order, int Ndst, int B, int64_t P, float* dphi, void* stream);

/* Dense stage of the wide path on the tensor cores: D[m][n] = sum_k A[m][k] * B[n][k] with A [M x K] and
 * B [N x K] row-major fp32, three kind::tf32 MMAs per product (fp32-accurate: ~2e-7 normwise).
 * rows_per_batch == 0: D is [M x N] row-major.  rows_per_batch = R > 0 (M a multiple of R): D is
 * [M/R][N][R], i.e. every block of R rows is written transposed -- the reference's [B, C, N] layout when
 * the rows are (batch, point).  K must be a multiple of 4 and A, B 16-byte aligned (QC_EUNSUP otherwise). */
int qc_gemm_nt_tf32x3(const float* A, const float* B, float* D, long long M, int N, int K,
                      long long rows_per_batch, void* stream);

/* The fourth GEMM of the wide path, dW_last: the reduction runs over the ROW index of the big operand.
 *     D[n][m] = sum_{b,kk} X[b][n][kk] * T[b*Kb + kk][m]
 * T [nb*Kb x M] row-major fp32 (read exactly once: the kernel is HBM bound), X [nb][N][Kb] (the reference's
 * [B, C, N] feature layout, quadconv.py:238), D [N x M] row-major; 3xTF32 on tcgen05 with T as the TMEM operand
 * (csrc/gemm_tc.cu).  K is split over CTAs; the partial results live in `ws` (qc_gemm_tn_workspace_bytes, may be 0)
 * and are summed in a fixed order (bit-reproducible). */
size_t qc_gemm_tn_workspace_bytes(int nb, int Kb, int M, int N);
int qc_gemm_tn_tf32x3(const float* T, const float* X, float* D, int nb, int Kb, int M, int N,
                      void* ws, size_t ws_bytes, void* stream);

/* Repack the last Linear weight W_last [Ci*Co][H] (filter.{2L}.weight, quadconv.py:143):
 *   mode 0: W2[(i*H+h)*CoP + j]  (forward)        mode 1: W2[(j*H+h)*CiP + i]  (d features)
 *   mode 2: W3[(j*Ci+i)*HP + h]  (d phi)
 * and the inverse scatter of dW2 (mode 1 layout) back to dW_last (+=). */
int qc_repack_wlast(const float* w_last, int Ci, int Co, int H, int HP, int mode, float* out,
                    void* stream);
int qc_unpack_dwlast(const float* dW2, int Ci, int Co, int H, float* d_w_last, void* stream);

/* ------------------------------------------------------------------------------------------
 * The whole layer in one call, fp32 parity mode, for hosts that do not compose the kernels themselves.
 *   qc_layer_forward  replaces QuadConv.forward + _integrate, torch_quadconv/quadconv.py:238-265, :213-227:
 *                     features [B][Ci][Ni] -> out [B][Co][No] (+ bias [Co][No] when given), reference layouts.
 *   qc_layer_backward replaces autograd of the above: grad_out [B][Co][No] -> d_features [B][Ci][Ni], d_quad_w [Ni]
 *                     (may be NULL), d_w_last [Ci*Co][H], d_hidden[l] (host array of device pointers, same shapes as
 *                     mlp.w[l]), d_bias [Co][No] (may be NULL).  All outputs are overwritten.
 * The descriptor holds device pointers to the mesh, the pair tables (from qc_search_* / qc_pairs_to_csr / qc_csc_*)
 * and the filter weights; nothing in it is modified.  `ws` (qc_layer_workspace_bytes) is caller-owned scratch that
 * ALSO carries the saved activations: qc_layer_backward must get the workspace of the matching qc_layer_forward
 * call, untouched.  Last hidden width H <= 32 (QC_EUNSUP otherwise).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int d, Ni, No, Ci, Co, H;           /* spatial dim, points, channels, last hidden width (= mlp.dims[n_hidden]) */
    int64_t P;                          /* number of pairs (rows of eval_indices) */
    const float* out_pts;               /* [No][d] */
    const float* in_pts;                /* [Ni][d] */
    const float* quad_w;                /* [Ni] quadrature weights (MeshHandler.weights()), NULL = ones */
    qc_mlp_t mlp;                       /* hidden layers of the filter MLP (filter.{0,2,..}.weight) */
    const float* w_last;                /* [Ci*Co][H] last Linear (filter.{2L}.weight) */
    const float* bias;                  /* [Co][No] or NULL */
    const int *pair_row, *pair_col;     /* [P] */
    const int *row_ptr;                 /* [No+1] */
    const int *csc_ptr, *csc_pair, *csc_row;   /* [Ni+1], [P], [P] */
    const int *order_out, *order_in;    /* processing orders or NULL */
} qc_layer_t;

size_t qc_layer_workspace_bytes(const qc_layer_t* layer, int B);
int qc_layer_forward(const qc_layer_t* layer, const float* features, int B, float* out, void* ws, size_t ws_bytes,
                     void* stream);
int qc_layer_backward(const qc_layer_t* layer, const float* grad_out, int B, float* d_features, float* d_quad_w,
                      float* d_w_last, float* const* d_hidden, float* d_bias, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5/K6  Mesh_MaxPool.  Replaces scatter_reduce_(amax, include_self=False) (mesh_pool.py:88-89)
 * and gather (mesh_pool.py:93) and their autograd.  Groups are given as CSR over the fine points:
 * grp_ptr[Nout+1], grp_mem[Nin] (members of each group).  Empty groups produce 0 like the
 * reference's zeros() initialisation.  Backward splits the gradient evenly between tied maxima
 * (torch's amax rule).  x, out: [BC][N] fp32 row-major.
 * ------------------------------------------------------------------------------------------ */
int qc_segmax_fwd(const float* x, const int* grp_ptr, const int* grp_mem, int BC, int Nin, int Nout,
                  float* out, void* stream);
int qc_segmax_bwd(const float* x, const float* out, const float* grad_out, const int* group /*[Nin]*/,
                  const int* grp_ptr, const int* grp_mem, int BC, int Nin, int Nout, float* grad_x,
                  void* stream);
int qc_gather_fwd(const float* x, const int* index /*[Nout]*/, int BC, int Nin, int Nout, float* out,
                  void* stream);
/* grad_x[r][g] = sum_{k in members(g)} grad_out[r][k]; grp_* is the CSR inverse of `index`. */
int qc_gather_bwd(const float* grad_out, const int* grp_ptr, const int* grp_mem, int BC, int Nfine,
                  int Ncoarse, float* grad_x, void* stream);

/* The same four kernels with the activation that precedes the pool in the reference's autoencoders fused in (SURVEY 8
 * f-4; the reference composes torch.sin and Mesh_MaxPool as separate modules): act = 0 is exactly the functions above,
 * act = 1 computes pool(sin(x)) / gather(sin(x)) and, backward, the routed gradient times cos(x).  qc_gather_act_bwd
 * reads x (the coarse input of the forward call, [BC][Ncoarse]) only when act != 0. */
int qc_segmax_act_fwd(const float* x, const int* grp_ptr, const int* grp_mem, int BC, int Nin, int Nout, int act,
                      float* out, void* stream);
int qc_segmax_act_bwd(const float* x, const float* out, const float* grad_out, const int* group, const int* grp_ptr,
                      const int* grp_mem, int BC, int Nin, int Nout, int act, float* grad_x, void* stream);
int qc_gather_act_fwd(const float* x, const int* index, int BC, int Nin, int Nout, int act, float* out, void* stream);
int qc_gather_act_bwd(const float* grad_out, const float* x, const int* grp_ptr, const int* grp_mem, int BC, int Nfine,
                      int Ncoarse, int act, float* grad_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * Learned quadrature-weight map.  Replaces MeshHandler.weight_map,
 * torch_quadconv/mesh_handler.py:101-112 (per-simplex Linear/Sigmoid x4 + scatter_add_) and its
 * autograd.  adj: [M][E] int32 simplices, wb: HOST array of 8 device pointers
 * {W0,b0,W1,b1,W2,b2,W3,b3} (Linear(E*d,8), (8,8), (8,8), (8,E), Sigmoid after each).
 *   qc_weight_map_fwd : el_w[M][E]
 *   qc_vertex_sum     : w[v] = sum of el_w over the incidences of vertex v (CSR inc_ptr/inc_idx of
 *                       flat (simplex*E+slot) indices, ascending) -- deterministic scatter_add_
 *   qc_weight_map_bwd : accumulates (+=) d W/b into dwb (8 device pointers) from gw = d loss / d w[N]
 * ------------------------------------------------------------------------------------------ */
int qc_weight_map_fwd(const float* pts, const int* adj, int M, int E, int d, const float* const* wb,
                      float* el_w, void* stream);
int qc_vertex_sum(const float* el_w, const int* inc_ptr, const int* inc_idx, int N, float* w, void* stream);
int qc_weight_map_bwd(const float* pts, const int* adj, int M, int E, int d, const float* const* wb,
                      const float* gw, float* const* dwb, void* stream);
/* bit-reproducible variant: per-block partial sums in `ws`, added in block order */
size_t qc_weight_map_bwd_workspace_bytes(int M);
int qc_weight_map_bwd_ws(const float* pts, const int* adj, int M, int E, int d, const float* const* wb,
                         const float* gw, float* const* dwb, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Reduced-precision tensor-core path (bf16 operands, fp32 accumulation; QuadConv(compute_dtype=torch.bfloat16)).
 * Replaces the same reference lines as qc_contract -- QuadConv._integrate and the last filter Linear,
 * torch_quadconv/quadconv.py:213-227, :143 -- with BOTH stages of the factorised layer on tcgen05
 * (csrc/contract_mma.cu).  Work unit: a tile of 16 destination points (consecutive entries of `order`) and the
 * union of their source points:
 *   tile_uptr[ntiles+1] : offsets into union_idx, every tile padded to a multiple of 32 entries (>= 32)
 *   union_idx[]         : source point of every union slot, -1 = padding
 *   tile_ent[P] (int2)  : the pairs grouped by tile: {phi row, (point in tile << 16) | union slot};
 *   tile_pptr[ntiles+1] : offsets into tile_ent
 *   max_ku              : largest padded union (bounds the shared-memory operand; QC_EUNSUP when it does not fit:
 *                         qc_mma_smem_bytes(...) <= 227 KB)
 * bf16 packed features  Xh[nblk][N][ceil(C/CC)][Bp][CC], CC = qc_mma_chunk_channels(C, B) in {4,8,16}.
 * Output: fp32 packed rows (same layout as qc_contract, feed qc_unpack_features).
 * transposed = 0: forward (sources = in points / in channels);  1: the d-features pass on the CSC view
 * (sources = out points / out channels).
 * Numerics: features, phi, the stage-B result T and W_last are rounded to bf16 (nearest even); both
 * accumulations are fp32.  Stated bound vs the fp32 layer: <= 1e-2 normwise (measured ~3e-3), tests/.
 * ------------------------------------------------------------------------------------------ */
int qc_mma_chunk_channels(int C, int B);          /* 0: batch lanes < 8, not served */
int qc_mma_smem_bytes(int Cs, int Cd, int B, int max_ku);
int qc_pack_features_bf16(const void* x /*[B][C][N] fp32, or bf16 when src_bf16 != 0*/, int src_bf16, int B, int C, int N,
                          int CC, void* xh, void* stream);
int qc_contract_mma(const void* Xh, int Nsrc, int Cs, const void* tile_ent, const int* tile_pptr, const int* tile_uptr,
                    const int* union_idx, int ntiles, int max_ku, const float* phi, int H, int HPphi,
                    const float* w_last, int Cd, int transposed, const int* order, int Ndst, int B, float* Yp,
                    void* stream);

/* d W_last of the reduced-precision mode (csrc/contract_mma.cu, DW variant): d_w_last[(i*Cg + j)*H + h] =
 * sum_{o,b} T[o;b,i,h] * g[b,j,o], T = the stage-B result of the forward recomputed from the same tiles (forward view:
 * Xh = features of the in points, Gh = grad_out of the out points, both bf16 packed).  Written (not accumulated).
 * ws: qc_dw_mma_workspace_bytes() bytes of scratch (per-CTA partial sums, reduced in CTA order: deterministic). */
size_t qc_dw_mma_workspace_bytes(int Cs, int Cg, int B);
int qc_dw_mma_smem_bytes(int Cs, int Cg, int B, int max_ku);
int qc_dw_mma(const void* Xh, int Nsrc, int Cs, const void* Gh, int Cg, const void* tile_ent, const int* tile_pptr,
              const int* tile_uptr, const int* union_idx, int ntiles, int max_ku, const float* phi, int H, int HPphi,
              const int* order, int Ndst, int B, void* ws, size_t ws_bytes, float* d_w_last, void* stream);

/* d(phi) of the reduced-precision mode (csrc/dphi_mma.cu): dphi[p][h] = sum_{b,i} f[b,i,in(p)] * dT[out(p); b,i,h],
 * dT = grad_out x W_last, over the same tiles/unions as qc_contract_mma (forward view: destinations = out points,
 * Xh = bf16 packed features of the in points, Gh = bf16 packed grad_out of the out points).  Autograd of
 * torch_quadconv/quadconv.py:216-219 with respect to the filter output, folded through the last Linear (:143).
 * dphi [P][HPphi] fp32, must be zero on entry when B > 32 (batch blocks accumulate).  QC_EUNSUP: batch lanes x chunk
 * channels < 64, H > 8, more than 64 grad channels, max_ku > 384, operands beyond 227 KB of shared memory. */
int qc_dphi_mma_smem_bytes(int Cs, int Cg, int B, int max_ku);
int qc_dphi_mma(const void* Xh, int Nsrc, int Cs, const void* Gh, int Cg, const void* tile_ent, const int* tile_pptr,
                const int* tile_uptr, const int* union_idx, int ntiles, int max_ku, const float* w_last, int H, int HPphi,
                const int* order, int Ndst, int B, float* dphi, void* stream);

/* Producer variant of the union-row gather in qc_contract_mma / qc_dw_mma / qc_dphi_mma (bit mask; default 4 = per-lane
 * cp.async copies into 128-byte-swizzled rows; 12 = the same layout filled by TMA: cp.async.bulk.tensor.2d tile::gather4,
 * four rows per instruction through a tensor map over the packed features -- measured 2.4x slower at C3, 128-byte boxes
 * are too small for the copy engine, so it is opt-in).  Also read once from QC_MMA_VAR.  set returns the previous value;
 * a negative argument only queries. */
int qc_mma_get_variant(void);
int qc_mma_set_variant(int var);

/* Known-answer self test of the tcgen05 building blocks (TMEM alloc, tcgen05.st/ld, kind::tf32 MMA with
 * A in TMEM and B through a K-major shared-memory descriptor, 3xTF32 split):
 * D[128][32] = A[128][64] * B[32][64]^T.  Used by tests only. */
int qc_tc_selftest(const float* A, const float* B, float* D, void* stream);
/* Same product with BOTH operands in shared memory (SS mode). Used by tests only. */
int qc_tc_selftest_ss(const float* A, const float* B, float* D, void* stream);
/* probe of the operand major-ness for 16-bit operands (kind::f16, bf16): D[128x32] = A[128x32] * B[32x32]^T;
 * mode bit 0: A MN-major, bit 1: B MN-major (K-major otherwise) */
int qc_tc_probe_bf16(const float* A, const float* B, float* D, int mode, void* stream);
/* diagnostic micro-benchmark: cycles for `iters` back-to-back kind::f16 MMAs (SS mode) with the given operand layouts
 * (major-ness, LBO/SBO, K-step advance in bytes, descriptor layout type), per CTA into out[grid]. */
int qc_tc_mma_bench(int M, int N, int a_mn, int b_mn, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int a_kstep,
                    int b_kstep, int swz, int iters, int grid, long long* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUADCONV_B200_H */
