#!/bin/bash
# round-2 GPU call L (1 GPU): 48-bit factor scratch A/B against 64-bit, MLPG parity suite, ncu traffic
mkdir -p gpurun_out
timeout 900 python tools/ab_mlpg.py nnmnkwii_b200/libnnk_b200.so nnmnkwii_b200/libnnk_b200_f64ws.so > gpurun_out/l_ab.log 2>&1; cat gpurun_out/l_ab.log
timeout 600 python -m pytest tests/test_mlpg_gpu.py tests/test_autograd_gpu.py tests/test_gmm_gpu.py -q > gpurun_out/l_pytest.log 2>&1; tail -3 gpurun_out/l_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/l_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['parity_max_rel_err_vs_oracle'], l['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlpg_fwd_as -c 1 -s 2 -o gpurun_out/l_mlpg_cfg2 python tools/profile_mlpg.py 3 > gpurun_out/l_ncu_mlpg.log 2>&1; echo "ncu mlpg rc=$?"
