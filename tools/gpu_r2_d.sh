#!/bin/bash
# round-2 GPU call D (2 GPUs): MLPG kernel A/B, ncu of the dominant kernel, fused DTW (rounds of 4), peer transport with split pushes
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/ab_mlpg.py nnmnkwii_b200/libnnk_b200.so nnmnkwii_b200/libnnk_b200_pairs.so nnmnkwii_b200/libnnk_b200_rot.so > gpurun_out/d_ab.log 2>&1; cat gpurun_out/d_ab.log
timeout 600 python -m pytest tests/test_dtw_gpu.py tests/test_sharding_gpu.py tests/test_autograd_gpu.py -q > gpurun_out/d_pytest.log 2>&1; tail -3 gpurun_out/d_pytest.log
NNK_SHARD_TRANSPORT=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/d_bench_n2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, l['allgather']['alone_gbs_per_rank'])"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/d_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step']); print(l['dtw']['exact']['ms_per_batch'], l['dtw']['fastdtw_radius1']['ms_per_batch'])"
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlpg_fwd_as -c 1 -s 2 -o gpurun_out/d_mlpg_cfg2 python tools/profile_mlpg.py 3 > gpurun_out/d_ncu_mlpg.log 2>&1; echo "ncu mlpg rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dtw_fused -c 1 -o gpurun_out/d_dtw_fused python tools/profile_uv_dtw.py exact > gpurun_out/d_ncu_dtw.log 2>&1; echo "ncu dtw rc=$?"
