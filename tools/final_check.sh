#!/bin/bash
# Round-end validation on the GPU box: parity tests, smoke, the default bench, the reference arm,
# the ncu launch list of the bench command and one full capture of the dominant kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_final.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlpg_fwd_as -c 1 -s 2 -o gpurun_out/r01_mlpg_v11_as_3a \
    python tools/profile_mlpg.py 2>&1 | tail -1
python - <<'PY'
import json
for f in ("gpurun_out/bench_final.log", "gpurun_out/bench_ref.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "cpu_baseline", "clocks", "gpu_launches", "impl")})
        r = d.get("roofline")
        if r:
            print("  roofline", {k: r[k] for k in ("achieved", "peak", "frac", "kernel_ms")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
