#!/bin/bash
# round-2 GPU call B (2 GPUs): full parity suite incl. the 2-GPU sharding tests, sharded bench, ncu of UV / DTW kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -8 gpurun_out/b_pytest.log
timeout 300 python tools/dtw_cfg4_oracle_check.py > gpurun_out/b_dtw_cfg4_oracle_check.log 2>&1; tail -3 gpurun_out/b_dtw_cfg4_oracle_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/b_bench_n2.err; tail -c 1500 gpurun_out/b_bench_n2.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:uv_fact -c 2 -o gpurun_out/b_uv_fact python tools/profile_uv_dtw.py uv 60 > gpurun_out/b_ncu_uv.log 2>&1; echo "ncu uv rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dtw_fused -c 1 -o gpurun_out/b_dtw_fused python tools/profile_uv_dtw.py exact > gpurun_out/b_ncu_dtw.log 2>&1; echo "ncu dtw rc=$?"
ls -la gpurun_out | tail -12
