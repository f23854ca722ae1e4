#!/bin/bash
# round-2 GPU call F (1 GPU): software-pipelined fused DTW, MLPG NA=4 variant, final ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dtw_gpu.py tests/test_gmm_gpu.py -q > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log
timeout 300 python tools/dtw_cfg4_oracle_check.py > gpurun_out/f_dtw_cfg4_oracle_check.log 2>&1; tail -3 gpurun_out/f_dtw_cfg4_oracle_check.log
timeout 900 python tools/ab_mlpg.py nnmnkwii_b200/libnnk_b200.so nnmnkwii_b200/libnnk_b200_na4.so > gpurun_out/f_ab.log 2>&1; cat gpurun_out/f_ab.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step']); print(l['dtw']['exact']['ms_per_batch'], l['dtw']['fastdtw_radius1']['ms_per_batch']); print(l['scale_workload']['value'], l['scale_workload']['ms_per_step'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dtw_fused -c 1 -o gpurun_out/f_dtw_fused python tools/profile_uv_dtw.py exact > gpurun_out/f_ncu_dtw.log 2>&1; echo "ncu dtw rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:uv_fact -c 2 -o gpurun_out/f_uv_fact python tools/profile_uv_dtw.py uv 60 > gpurun_out/f_ncu_uv.log 2>&1; echo "ncu uv rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/f_launch_bench.log 2>&1; echo "launch list rc=$?"
