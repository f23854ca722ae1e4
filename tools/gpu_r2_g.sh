#!/bin/bash
# round-2 GPU call G (8 GPUs): configs[4] strong scaling at N = 8 and 4 (peer transport), sharding tests
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/g_topo.txt 2>&1
timeout 300 python -m pytest tests/test_sharding_gpu.py -q > gpurun_out/g_pytest.log 2>&1; tail -2 gpurun_out/g_pytest.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/g_bench_n$n.json 2> gpurun_out/g_bench_n$n.err; echo "bench n$n rc=$?"
  tail -c 300 gpurun_out/g_bench_n$n.err
  python -c "
import json;l=json.loads(open('gpurun_out/g_bench_n$n.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, l['allgather']['alone_gbs_per_rank'], l['e2e']['value'], l['parity_max_rel_err_vs_oracle'])"
done
NNK_SHARD_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/g_bench_n8_nccl.json 2> gpurun_out/g_bench_n8_nccl.err; echo "bench n8 nccl rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/g_bench_n8_nccl.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, l['allgather']['alone_gbs_per_rank'])"
