#!/bin/bash
# round-2 GPU call K (1 GPU): sanitizer runs of the round-2 kernels, bench with the UV graph timing
mkdir -p gpurun_out
for tool in memcheck initcheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool python tools/sanitize_others.py > gpurun_out/k_sanitizer_others_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; tail -2 gpurun_out/k_sanitizer_others_$tool.log
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/k_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step']); print(l['other_kernels']['unit_variance_mlpg_fwd_bwd'])"
