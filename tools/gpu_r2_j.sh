#!/bin/bash
# round-2 GPU call J (8 GPUs): two push streams (two rotating permutations at a time) at N = 8 / 4
mkdir -p gpurun_out
run() { # name nproc streams
  NNK_PEER_STREAMS=$3 NNK_SHARD_TRANSPORT=peer timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 295$2$3 bench.py --gpus $2 --steps 20 --warmup 3 > gpurun_out/j_bench_$1.json 2> gpurun_out/j_bench_$1.err; echo "bench $1 rc=$?"
  python -c "
import json;l=json.loads(open('gpurun_out/j_bench_$1.json').read().strip().splitlines()[-1])
print('$1',{k:round(l[k],3) if isinstance(l[k],float) else l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, round(l['allgather']['alone_gbs_per_rank']), l['parity_max_rel_err_vs_oracle'])"
}
run n8_s2 8 2
run n8_s3 8 3
run n4_s2 4 2
