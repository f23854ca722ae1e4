#!/bin/bash
# round-2 GPU call H (2 GPUs): fused p2p stores (sharding tests, N=2 bench for p2p / peer), MLPG NA=4 capped A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharding_gpu.py tests/test_mlpg_gpu.py -q > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
for tr in p2p peer; do
  NNK_SHARD_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/h_bench_n2_$tr.json 2> gpurun_out/h_bench_n2_$tr.err; echo "bench n2 $tr rc=$?"
  tail -c 300 gpurun_out/h_bench_n2_$tr.err
  python -c "
import json;l=json.loads(open('gpurun_out/h_bench_n2_$tr.json').read().strip().splitlines()[-1])
print('$tr',{k:l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, l['parity_max_rel_err_vs_oracle'])"
done
CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/ab_mlpg.py nnmnkwii_b200/libnnk_b200.so nnmnkwii_b200/libnnk_b200_na4.so > gpurun_out/h_ab.log 2>&1; cat gpurun_out/h_ab.log
