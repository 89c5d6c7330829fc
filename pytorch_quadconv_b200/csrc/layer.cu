// layer.cu -- single-call entry points for the whole QuadConv layer (qc_forward / qc_backward).
//
// A host that is not Python (the C/C++ stub of INTEGRATION.md) calls the layer as two functions:
//     qc_layer_forward   = QuadConv.forward + _integrate          torch_quadconv/quadconv.py:238-265, :213-227
//     qc_layer_backward  = autograd of the above                  (no reference code: SURVEY 3.3)
// on [B, C, N] fp32 tensors in the reference's own layout, with one caller-owned workspace that also carries the
// saved activations (packed features, phi) from the forward call to the backward call.  They compose the fused fp32
// kernels of this library (pair MLP -> pack -> contract -> unpack); nothing is allocated here.
#include "common.cuh"

namespace {

struct Carve {
    char* p;
    size_t used;
    explicit Carve(void* base) : p(reinterpret_cast<char*>(base)), used(0) {}
    template <typename T>
    T* take(size_t n)
    {
        const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
        T* r = p ? reinterpret_cast<T*>(p + used) : nullptr;
        used += bytes;
        return r;
    }
};

inline int hp_of(int H) { return H <= 4 ? 4 : (H <= 8 ? 8 : (H <= 16 ? 16 : 32)); }

struct Plan {
    float *phi, *xp, *yp, *w2, *gp, *w3, *dphi, *w2t, *dw2, *dxp;
    void *cws, *mws;
    size_t cws_bytes, mws_bytes, total;
    int HP, nsl;
};

Plan make_plan(const qc_layer_t* L, int B, void* ws)
{
    Plan pl;
    Carve c(ws);
    const int Bp = qc_bp(B), nblk = qc_nblk(B);
    const int H = L->H, HP = hp_of(H);
    const size_t P = (size_t)(L->P > 0 ? L->P : 1);
    const size_t xin = (size_t)nblk * L->Ni * qc_cp4(L->Ci) * Bp, xout = (size_t)nblk * L->No * qc_cp4(L->Co) * Bp;
    pl.HP = HP;
    pl.phi = c.take<float>(P * HP);                       // saved for backward
    pl.xp = c.take<float>(xin);                           // saved for backward
    pl.yp = c.take<float>(xout);
    pl.w2 = c.take<float>((size_t)L->Ci * H * qc_cp4(L->Co));
    pl.gp = pl.yp;                                        // backward reuses the output staging for packed grad_out
    pl.w3 = c.take<float>((size_t)L->Co * L->Ci * HP);
    pl.nsl = qc_dphi_num_slices(L->Ni, L->Ci, L->Co, HP, B);
    pl.dphi = c.take<float>((size_t)pl.nsl * P * HP);
    pl.w2t = c.take<float>((size_t)L->Co * H * qc_cp4(L->Ci));
    pl.dw2 = c.take<float>((size_t)L->Co * H * qc_cp4(L->Ci));
    pl.dxp = c.take<float>(xin);
    pl.cws_bytes = qc_contract_workspace_bytes(L->Co, H, HP, L->Ci, L->Ni, B);
    pl.cws = c.take<char>(pl.cws_bytes);
    pl.mws_bytes = qc_pair_phi_bwd_workspace_bytes(L->P, L->d, &L->mlp, HP);   // partial sums of the per-pair backward
    pl.mws = c.take<char>(pl.mws_bytes);
    pl.total = c.used;
    return pl;
}

}  // namespace

extern "C" {

size_t qc_layer_workspace_bytes(const qc_layer_t* L, int B)
{
    if (L == nullptr || B <= 0) return 0;
    return make_plan(L, B, nullptr).total;
}

int qc_layer_forward(const qc_layer_t* L, const float* features, int B, float* out, void* ws, size_t ws_bytes,
                     void* stream)
{
    if (L == nullptr || features == nullptr || out == nullptr || B <= 0 || L->H < 1) return QC_EINVAL;
    if (L->H > 32) return QC_EUNSUP;
    const Plan pl = make_plan(L, B, ws);
    if (ws == nullptr || ws_bytes < pl.total) return QC_EWORKSPACE;
    int rc;
    if ((rc = qc_pair_phi_fwd(L->out_pts, L->in_pts, L->quad_w, L->pair_row, L->pair_col, L->P, L->d, &L->mlp, pl.HP, pl.phi,
                              stream)) != 0)
        return rc;
    if ((rc = qc_pack_features(features, B, L->Ci, L->Ni, pl.xp, stream)) != 0) return rc;
    if ((rc = qc_repack_wlast(L->w_last, L->Ci, L->Co, L->H, pl.HP, 0, pl.w2, stream)) != 0) return rc;
    if ((rc = qc_contract(pl.xp, L->Ni, L->Ci, L->row_ptr, L->pair_col, nullptr, pl.phi, L->H, pl.HP, pl.w2, L->Co, L->order_out,
                          L->No, B, pl.yp, nullptr, nullptr, nullptr, 0, stream)) != 0)
        return rc;
    return qc_unpack_features(pl.yp, B, L->Co, L->No, L->bias, out, stream);
}

int qc_layer_backward(const qc_layer_t* L, const float* grad_out, int B, float* d_features, float* d_quad_w,
                      float* d_w_last, float* const* d_hidden, float* d_bias, void* ws, size_t ws_bytes, void* stream)
{
    if (L == nullptr || grad_out == nullptr || d_features == nullptr || d_w_last == nullptr || B <= 0) return QC_EINVAL;
    if (L->H > 32) return QC_EUNSUP;
    const Plan pl = make_plan(L, B, ws);
    if (ws == nullptr || ws_bytes < pl.total) return QC_EWORKSPACE;
    const cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = qc_pack_features(grad_out, B, L->Co, L->No, pl.gp, stream)) != 0) return rc;
    if (d_bias != nullptr && (rc = qc_batch_sum(grad_out, B, L->Co, L->No, d_bias, stream)) != 0) return rc;
    // d phi (grouped by out point), then the per-pair chain rule through the hidden layers
    if ((rc = qc_repack_wlast(L->w_last, L->Ci, L->Co, L->H, pl.HP, 2, pl.w3, stream)) != 0) return rc;
    if ((rc = qc_dphi(pl.xp, L->Ni, L->Ci, pl.gp, L->Co, L->row_ptr, L->pair_col, pl.w3, L->H, pl.HP, L->order_out, L->No, B, L->P,
                      pl.dphi, stream)) != 0)
        return rc;
    // zero the accumulated outputs.  A host that carves them out of one buffer (the Python modules do: pieces in
    // ascending order, at most 16-byte alignment gaps) gets ONE fill instead of one per piece -- a small-mesh layer step
    // is launch bound.  (A gap below 64 bytes cannot hold anybody else's data: device allocators hand out blocks of at
    // least 256 bytes.)
    {
        struct Span { char* p; size_t n; };
        Span sp[QC_MAX_MLP_LAYERS + 2];
        int ns = 0;
        for (int l = 0; l < L->mlp.n_hidden; ++l) {
            if (d_hidden == nullptr || d_hidden[l] == nullptr) return QC_EINVAL;
            sp[ns++] = {reinterpret_cast<char*>(d_hidden[l]), (size_t)L->mlp.dims[l] * L->mlp.dims[l + 1] * sizeof(float)};
        }
        if (d_quad_w != nullptr) sp[ns++] = {reinterpret_cast<char*>(d_quad_w), (size_t)L->Ni * sizeof(float)};
        sp[ns++] = {reinterpret_cast<char*>(d_w_last), (size_t)L->Ci * L->Co * L->H * sizeof(float)};
        bool one = true;
        for (int i = 1; i < ns; ++i) {
            const char* prev_end = sp[i - 1].p + sp[i - 1].n;
            if (sp[i].p < prev_end || sp[i].p - prev_end > 64) one = false;
        }
        if (one) {
            QC_CUDA(cudaMemsetAsync(sp[0].p, 0, (size_t)(sp[ns - 1].p + sp[ns - 1].n - sp[0].p), st));
        } else {
            for (int i = 0; i < ns; ++i) QC_CUDA(cudaMemsetAsync(sp[i].p, 0, sp[i].n, st));
        }
    }
    if (L->P > 0 &&
        (rc = qc_pair_phi_bwd_ws(L->out_pts, L->in_pts, L->quad_w, L->pair_row, L->pair_col, L->P, L->d, &L->mlp, pl.HP, pl.dphi,
                                 pl.nsl, d_hidden, d_quad_w, L->csc_ptr, L->csc_pair, L->Ni, pl.mws, pl.mws_bytes, stream)) != 0)
        return rc;
    // d features (CSC view) and d W_last in the same pass
    if ((rc = qc_repack_wlast(L->w_last, L->Ci, L->Co, L->H, pl.HP, 1, pl.w2t, stream)) != 0) return rc;
    QC_CUDA(cudaMemsetAsync(pl.dw2, 0, (size_t)L->Co * L->H * qc_cp4(L->Ci) * sizeof(float), st));
    if ((rc = qc_contract(pl.gp, L->No, L->Co, L->csc_ptr, L->csc_row, L->csc_pair, pl.phi, L->H, pl.HP, pl.w2t, L->Ci,
                          L->order_in, L->Ni, B, pl.dxp, pl.xp, pl.dw2, pl.cws, pl.cws_bytes, stream)) != 0)
        return rc;
    if ((rc = qc_unpack_dwlast(pl.dw2, L->Ci, L->Co, L->H, d_w_last, stream)) != 0) return rc;
    return qc_unpack_features(pl.dxp, B, L->Ci, L->Ni, nullptr, d_features, stream);
}

}  // extern "C"
