#!/bin/bash
# round-2 GPU call E (1 GPU): instruction micro-benchmark, pipelined fused DTW, GMM kernels, full parity suite
mkdir -p gpurun_out
./tools/ubench/ubench > gpurun_out/e_ubench.log 2>&1; cat gpurun_out/e_ubench.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log; tail -6 gpurun_out/e_pytest.log
timeout 300 python tools/dtw_cfg4_oracle_check.py > gpurun_out/e_dtw_cfg4_oracle_check.log 2>&1; tail -3 gpurun_out/e_dtw_cfg4_oracle_check.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/e_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['traffic']); print(l['dtw']['exact']['ms_per_batch'], l['dtw']['fastdtw_radius1']['ms_per_batch'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dtw_fused -c 1 -o gpurun_out/e_dtw_fused python tools/profile_uv_dtw.py exact > gpurun_out/e_ncu_dtw.log 2>&1; echo "ncu dtw rc=$?"
