#!/bin/bash
# round-2: configs[4] at N = 8 with the final code (peer transport, rotating single-stream pushes)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/z_bench_n8.json 2> gpurun_out/z_bench_n8.err; echo "bench n8 rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/z_bench_n8.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, l['allgather']['alone_gbs_per_rank'], l['e2e']['value'], l['parity_max_rel_err_vs_oracle'])"
