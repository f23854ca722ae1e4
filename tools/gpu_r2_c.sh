#!/bin/bash
# round-2 GPU call C (2 GPUs): parity suite, peer-transport sharded bench vs NCCL, MLPG kernel A/B, UV timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -6 gpurun_out/c_pytest.log
for tr in peer nccl; do
  NNK_SHARD_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c_bench_n2_$tr.json 2> gpurun_out/c_bench_n2_$tr.err; echo "bench n2 $tr rc=$?"
  tail -c 300 gpurun_out/c_bench_n2_$tr.err
  python -c "
import json;l=json.loads(open('gpurun_out/c_bench_n2_$tr.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','kernel_ms','allgather_ms','allgather_exposed_ms')}, l['allgather']['alone_gbs_per_rank'], l['e2e']['value'])"
done
CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/ab_mlpg.py nnmnkwii_b200/libnnk_b200.so nnmnkwii_b200/libnnk_b200_norot.so nnmnkwii_b200/libnnk_b200_pairs.so > gpurun_out/c_ab.log 2>&1; cat gpurun_out/c_ab.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
python -c "
import json;l=json.loads(open('gpurun_out/c_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step']); print({k:(v.get('ms') or v.get('ms_per_iter')) for k,v in l['other_kernels'].items()}); print(l['dtw']['exact']['ms_per_batch'], l['dtw']['fastdtw_radius1']['ms_per_batch'])"
