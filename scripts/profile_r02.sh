set -x
cd $GRAFT_REPO_ROOT
python scripts/run_layer.py --steps 3 > gpurun_out/r02_run_fp32.txt 2>&1 || exit 1
python scripts/run_layer.py --steps 3 --compute-dtype bf16 > gpurun_out/r02_run_bf16.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_r2_fp32 --kernel-name regex:"k_contract_tc|k_dphi_tc" --launch-skip 3 --launch-count 3 python scripts/run_layer.py --steps 2 > gpurun_out/ncu_fp32.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_r2_bf16 --kernel-name regex:"k_contract_mma|k_dphi_mma" --launch-skip 4 --launch-count 4 python scripts/run_layer.py --steps 2 --compute-dtype bf16 > gpurun_out/ncu_bf16.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tc --no-c5 > gpurun_out/r02_plain_fp32.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c3_fp32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tc --no-c5 > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c3_bf16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --compute-dtype bf16 --no-c5 > gpurun_out/ncu_l2.log 2>&1
ls -la gpurun_out/*.ncu-rep
