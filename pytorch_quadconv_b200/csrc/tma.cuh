// tma.cuh -- Tensor Memory Accelerator helpers (sm_100a): a tensor map over a row-major 2-D view of a packed
// tensor, and the Blackwell row-gather load (cp.async.bulk.tensor.2d ... tile::gather4): ONE instruction fetches four
// rows at arbitrary row indices (a box of `box_cols` elements each) into four consecutive 128-byte-swizzled
// shared-memory rows and reports the bytes to an mbarrier -- exactly the access pattern of the union-row gather of
// csrc/contract_mma.cu / dphi_mma.cu (16 instructions per 8 KB ring stage instead of 512 per-lane LDGSTS copies).
// The tensor map is encoded on the host through the driver entry point (no link against libcuda) and handed to the
// kernel as a __grid_constant__ parameter.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Row-major [rows][cols] tensor of 2- or 4-byte elements, box = box_cols x box_rows elements, 128-byte swizzle
// (box_cols * element size must be 128), out-of-bounds elements read as zero.  false when the driver call is not
// available or rejects the shape.
static inline bool make_map_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                               uint32_t box_cols, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstr[1] = {cols * (uint64_t)elem_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return fn(map, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// one arrival + `bytes` expected transaction bytes on a CTA-local mbarrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}

// rows r0..r3 of the mapped tensor, columns [col, col + box_cols), into 4 consecutive 128-byte rows at smem_dst
__device__ __forceinline__ void gather4(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int col, int r0, int r1,
                                        int r2, int r3)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
          "r"((uint32_t)__cvta_generic_to_shared(bar))
        : "memory");
}

// plain 2-D tile load: box of the map at (col, row) into smem_dst
__device__ __forceinline__ void load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int col, int row)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"((uint32_t)__cvta_generic_to_shared(bar))
        : "memory");
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

}  // namespace tma
