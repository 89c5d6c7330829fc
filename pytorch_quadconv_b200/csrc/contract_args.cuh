// contract_args.cuh -- argument blocks shared by the generic (contract.cu) and the fast
// (contract_fast.cu) gather-contract kernels.
#pragma once
#include "common.cuh"

struct ContractArgs {
    const float* X;
    float* Y;
    const float* Z;
    float* dW2;
    float* dW_scratch;   // tcgen05 dW variant: per-CTA slots of partial dW2 accumulators [cta][slot][chunk][64][32]
    int* dW_counts;      // ... and the number of slots every CTA wrote
    void* ws;            // caller-provided workspace for the above (may be null: stream-ordered allocation)
    size_t ws_bytes;
    const int* seg_ptr;
    const int* src_idx;
    const int* phi_idx;
    const float* phi;
    const float* W2;
    const int* order;
    int Nsrc, Ndst, Cx, Cy, CyP, CyS, H, LB, LWK, KSPLIT;
    int dw_ordered;      // tcgen05 dW variant: the row blocks of a CTA take the dW operand buffer in strict alternation
};

struct DphiArgs {
    const float* X;
    const float* G;
    const float* W3;
    const int* seg_ptr;
    const int* src_idx;
    const int* order;
    float* dphi;
    int Nsrc, Ndst, Cx, Cg, H, LWK;
    int64_t P;
};

// Fast paths (contract_fast.cu).  Return 1 when the shape is handled, 0 when the caller must
// fall back to the generic kernel, <0 / cudaError on failure.
int qc_contract_tc(const ContractArgs& a, int HP, int B, bool dw, cudaStream_t st);
size_t qc_contract_tc_ws_bytes(int Ndst, int Cx, int Cy, int HP, int B);   // dW variant; 0 when the shape is not served
int qc_contract_fast(const ContractArgs& a, int HP, int B, bool dw, cudaStream_t st);
int qc_dphi_tc(const DphiArgs& a, int HP, int B, cudaStream_t st);
int qc_dphi_fast(const DphiArgs& a, int HP, int B, cudaStream_t st);
bool qc_dphi_fast_ok(int Nsrc, int Cx, int Cg, int HP, int B);
