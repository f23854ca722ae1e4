#!/usr/bin/env python
"""A/B timing of MLPG kernel variants: for every library given (NNK_LIB_PATH), the configs[1] device step,
the T=1000 x 256 forward / gradient, and the parity tests of the MLPG family.

    python tools/ab_mlpg.py nnmnkwii_b200/libnnk_b200.so nnmnkwii_b200/libnnk_b200_pairs.so ...
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INNER = r'''
import json, sys, numpy as np, torch
sys.path.insert(0, %r)
import bench
from nnmnkwii_b200 import _device as dev, _lib, paramgen as G
device = torch.device("cuda", 0)
lens, means, variances = bench.make_batch(0)
layout = G.merlin_layout()
wc = _lib.make_windows(bench.WINDOWS)
n_rows = int(lens.sum())
off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(device)
order = torch.from_numpy(np.argsort(-lens, kind="stable").astype(np.int32)).to(device)
dm, dv = torch.from_numpy(means).to(device), torch.from_numpy(variances).to(device)
out = torch.zeros((n_rows, 63), device=device)
chains = dev.chains_on_device(layout.chains, device)
def step():
    return dev.run_mlpg("fwd", means=dm, variances=dv, rhs=None, out=out, offsets=off, lengths=None, order=order, chains=chains,
                        n_chain=63, max_T=int(lens.max()), windows_c=wc, in_ld=187, var_ld=187, go_ld=0, out_ld=63,
                        dtype_code=_lib.NNK_F32, go_f64=0, n_utt=len(lens), device=device, check=False)
res = {}
ms, how = bench._device_ms(step, 50)
res["cfg2_ms"] = ms
import oracle
a, b = 0, int(lens[0])
ref = oracle.mlpg(means[a:b, :180], variances[a:b, :180], bench.WINDOWS)
res["cfg2_err"] = float(np.abs(out[a:b, :60].cpu().numpy() - ref).max() / np.abs(ref).max())
g = torch.Generator(device=device).manual_seed(0)
B2, T2, sd2 = 256, 1000, 60
ch2 = dev.chains_on_device(dev.simple_chains(sd2), device)
off2 = torch.arange(B2 + 1, dtype=torch.int64, device=device) * T2
m2 = torch.rand(B2 * T2, 180, device=device, generator=g); v2 = torch.rand(B2 * T2, 180, device=device, generator=g) + 0.1
go2 = torch.randn(B2 * T2, sd2, device=device, generator=g)
y2 = torch.zeros(B2 * T2, sd2, device=device); g2 = torch.zeros(B2 * T2, 180, device=device)
def run(mode, rhs, o, out_ld):
    return dev.run_mlpg(mode, means=m2, variances=v2, rhs=rhs, out=o, offsets=off2, lengths=None, order=None, chains=ch2, n_chain=sd2,
                        max_T=T2, windows_c=wc, in_ld=180, var_ld=180, go_ld=sd2, out_ld=out_ld, dtype_code=_lib.NNK_F32, go_f64=0,
                        n_utt=B2, device=device, check=False)
res["T1000_fwd_ms"] = bench._device_ms(lambda: run("fwd", None, y2, sd2), 20)[0]
res["T1000_grad_ms"] = bench._device_ms(lambda: run("grad", go2, g2, 180), 20)[0]
print("ABRES " + json.dumps(res))
''' % ROOT


def main():
    run_tests = "--no-tests" not in sys.argv
    for lib in [a for a in sys.argv[1:] if not a.startswith("--")]:
        env = dict(os.environ, NNK_LIB_PATH=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", INNER], env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("ABRES ")]
        print(os.path.basename(lib), line[0][6:] if line else ("FAILED: " + r.stderr[-800:]))
        if not run_tests:
            continue
        t = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_mlpg_gpu.py"), "-q", "-x"], env=env,
                           capture_output=True, text=True, timeout=900)
        print(os.path.basename(lib), "pytest:", t.stdout.strip().splitlines()[-1] if t.stdout.strip() else t.stderr[-300:])
        sys.stdout.flush()


if __name__ == "__main__":
    main()
