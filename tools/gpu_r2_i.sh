#!/bin/bash
# round-2 GPU call I (8 GPUs): staggered single-stream peer pushes at N = 8 / 4, NCCL at N = 4
mkdir -p gpurun_out
run() { # name nproc transport
  NNK_SHARD_TRANSPORT=$3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 295$2$2 bench.py --gpus $2 --steps 20 --warmup 3 > gpurun_out/i_bench_$1.json 2> gpurun_out/i_bench_$1.err; echo "bench $1 rc=$?"
  python -c "
import json;l=json.loads(open('gpurun_out/i_bench_$1.json').read().strip().splitlines()[-1])
print('$1',{k:round(l[k],3) if isinstance(l[k],float) else l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, round(l['allgather']['alone_gbs_per_rank']), l['parity_max_rel_err_vs_oracle'])"
}
run n8_peer 8 peer
run n4_peer 4 peer
run n4_nccl 4 nccl
