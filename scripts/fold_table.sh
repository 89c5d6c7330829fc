# kernel tables of the C5 step with and without batch folding (gpurun helper)
cd ${GRAFT_REPO_ROOT:-.}
for mb in 32 8 0; do
for B in 256 128 64 32; do
python - > gpurun_out/fold_mb${mb}_b${B}.txt 2>/dev/null <<PY
import sys, runpy
sys.path.insert(0, ".")
from pytorch_quadconv_b200 import ops
ops.FOLD_MIN_BATCH = $mb
sys.argv = ["k", "$B", "3"]
runpy.run_path("scripts/kernel_table.py", run_name="__main__")
PY
echo "min_batch $mb B $B: $(head -1 gpurun_out/fold_mb${mb}_b${B}.txt)"
done
done
