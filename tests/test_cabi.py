"""CPU: the C-ABI library builds, loads, and exports every symbol include/quadconv_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "quadconv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from pytorch_quadconv_b200 import _cabi, build
    build.build()
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == names
    assert _cabi.lib().qc_version() >= 1
    # pure host helper: workspace sizing needs no device
    assert _cabi.lib().qc_search_workspace_bytes(1000, 1000, 2) > 0
    assert _cabi.lib().qc_search_workspace_bytes(1000, 1000, 4) == 0


def test_host_side_shape_queries():
    """Pure host helpers of the C ABI (no device needed): which shapes the wide path serves, how many d(phi)
    slices the kernels write, and the workspace of the dW variant (without a GPU the SM count falls back to 148)."""
    from pytorch_quadconv_b200 import _cabi
    L = _cabi.lib()
    assert L.qc_gather_supported(64, 32, 32, 8) == 1          # C4: 64 channels, H = 32, batch 8
    assert L.qc_gather_supported(64, 32, 32, 32) == 1
    assert L.qc_gather_supported(64, 32, 32, 2) == 0          # 2 batch lanes: general kernel
    assert L.qc_gather_supported(64, 40, 32, 8) == 0          # H > HP
    assert L.qc_gather_dot_num_slices(64, 32, 8) == 1 and L.qc_gather_dot_num_slices(64, 32, 70) == 3
    assert L.qc_dphi_num_slices(100000, 32, 32, 8, 32) == 4   # C3: one slice per channel chunk
    assert L.qc_dphi_num_slices(1000, 64, 128, 32, 8) == 1    # not served by the fused kernels
    ws = L.qc_contract_workspace_bytes(32, 8, 8, 32, 100000, 32)
    # C3 backward: 85 tiles per CTA x 2 row blocks x 4 chunks = 680 uses of the operand buffer, one slot per 16 uses ->
    # 148 CTAs x 43 slots x 4 chunks x 64 x 32 floats + one counter per CTA
    assert ws == 148 * 43 * 4 * 64 * 32 * 4 + 148 * 4
    assert L.qc_contract_workspace_bytes(64, 32, 32, 128, 10000, 8) == 0   # wide shapes: no tcgen05 dW variant


def test_no_cpu_fallback():
    import pytest
    import torch
    import pytorch_quadconv_b200 as q
    pts = torch.rand(50, 2)
    mesh = q.MeshHandler(pts).construct([50])
    layer = q.QuadConv(spatial_dim=2, in_points=50, out_points=50, in_channels=1, out_channels=1,
                       filter_seq=[4], output_same=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(mesh, torch.randn(1, 1, 50))
    with pytest.raises(RuntimeError, match="no CPU path"):
        q.Mesh_MaxPool()(mesh, torch.randn(1, 1, 50))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorch_quadconv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"
