"""Diagnostic: issue/execution rate of tcgen05.mma kind::f16 (SS mode) for the operand layouts of contract_mma.cu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_quadconv_b200 import _cabi as C
L = C.lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
cases = [
    # name, M, N, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, swz
    ("stageB MN/MN N=128 sbo128", 128, 128, 1, 1, 2048, 128, 2048, 128, 4096, 4096, 0),
    ("stageB MN/MN N=128 B sbo144", 128, 128, 1, 1, 2048, 128, 16 * 144, 144, 4096, 2 * 16 * 144, 0),
    ("  + commit every 2 MMAs", 128, 128, 1, 1, 2048, 128, 16 * 144, 144, 4096, 2 * 16 * 144, 1 << 8),
    ("  + fences every 2 MMAs", 128, 128, 1, 1, 2048, 128, 16 * 144, 144, 4096, 2 * 16 * 144, 2 << 8),
    ("  + commit + fences", 128, 128, 1, 1, 2048, 128, 16 * 144, 144, 4096, 2 * 16 * 144, 3 << 8),
    ("  + background st.shared (3 warps)", 128, 128, 1, 1, 2048, 128, 16 * 144, 144, 4096, 2 * 16 * 144, 4 << 8),
    ("  + background tcgen05.ld (3 warps)", 128, 128, 1, 1, 2048, 128, 16 * 144, 144, 4096, 2 * 16 * 144, 8 << 8),
    ("  + all of the above", 128, 128, 1, 1, 2048, 128, 16 * 144, 144, 4096, 2 * 16 * 144, 15 << 8),
    ("stageB MN/MN N=256 sbo128", 128, 256, 1, 1, 2048, 128, 4096, 128, 4096, 8192, 0),
    ("stageB MN/MN N=64 sbo128", 128, 64, 1, 1, 2048, 128, 1024, 128, 4096, 2048, 0),
    ("K/K N=128 noswz", 128, 128, 0, 0, 128, 256, 128, 256, 256 * 16, 256 * 16, 0),
    ("K/K N=256 noswz", 128, 256, 0, 0, 128, 256, 128, 256, 256 * 16, 256 * 32, 0),
    ("stageC K/K N=32 (144/1168 ; 128/512)", 128, 32, 0, 0, 144, 1168, 128, 512, 288, 256, 0),
    ("stageC K/K N=32 aligned (128/1024)", 128, 32, 0, 0, 128, 1024, 128, 512, 256, 256, 0),
    ("K/K N=128 SW128 (K=64 atoms)", 128, 128, 0, 0, 16, 1024, 16, 1024, 32, 32, 2),
    ("K/K N=256 SW128", 128, 256, 0, 0, 16, 1024, 16, 1024, 32, 32, 2),
    ("MN/MN N=128 SW128", 128, 128, 1, 1, 8192, 1024, 8192, 1024, 2048, 2048, 2),
]
for grid in (1,):
    for c in cases:
        name, args = c[0], c[1:]
        for iters in (64, 512):
            C.check(L.qc_tc_mma_bench(*args, iters, grid, C.ptr(out), C.stream()), name)
            torch.cuda.synchronize()
        cyc = out[:grid].double().mean().item()
        print(f"grid {grid:3d}  {name:42s} {cyc / 512:7.1f} cycles/MMA")
