#!/bin/bash
# round-2 GPU call A: parity tests, bench (N=1), full-size DTW oracle check, sanitizer logs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/a_bench.err
timeout 300 python tools/dtw_cfg4_oracle_check.py > gpurun_out/a_dtw_cfg4_oracle_check.log 2>&1; tail -3 gpurun_out/a_dtw_cfg4_oracle_check.log
for tool in memcheck synccheck initcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_mlpg.py > gpurun_out/a_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; tail -2 gpurun_out/a_sanitizer_$tool.log
done
