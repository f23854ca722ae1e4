#!/bin/bash
# round-2 GPU call M (1 GPU): FastDTW wavefront / back-track rework: parity, full-size oracle check, timing, phase cycles
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dtw_gpu.py -q > gpurun_out/m_pytest.log 2>&1; tail -3 gpurun_out/m_pytest.log
timeout 300 python tools/dtw_cfg4_oracle_check.py > gpurun_out/m_dtw_cfg4_oracle_check.log 2>&1; tail -3 gpurun_out/m_dtw_cfg4_oracle_check.log
NNK_DTW_PROF=1 timeout 300 python tools/profile_uv_dtw.py fast 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/m_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step']); print(l['dtw']['exact']['ms_per_batch'], l['dtw']['fastdtw_radius1']['ms_per_batch'])"
