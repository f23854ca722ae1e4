#!/bin/bash
# round-2 final GPU call (2 GPUs): the full parity suite, smoke(), bench N=1 (with the CPU arm) and N=2, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log; tail -4 gpurun_out/z_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; tail -1 gpurun_out/z_smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/z_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['traffic'], l['e2e']['value'], l['cpu_baseline']['value'], l['gpu_launches'])
print(l['dtw']['exact']['ms_per_batch'], l['dtw']['fastdtw_radius1']['ms_per_batch'], l['scale_workload']['value'], l['scale_workload']['ms_per_step'])
print({k:(v.get('ms') or v.get('ms_per_iter')) for k,v in l['other_kernels'].items()})"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/z_bench_n2.json 2> gpurun_out/z_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/z_bench_n2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, l['parity_max_rel_err_vs_oracle'])"
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/z_launch_bench.log 2>&1; echo "launch list rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none -k regex:gmm_ -c 3 -o gpurun_out/z_gmm python -m pytest tests/test_gmm_gpu.py -q -k voice > gpurun_out/z_ncu_gmm.log 2>&1; echo "ncu gmm rc=$?"
