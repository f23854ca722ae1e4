"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py <rep> [out.txt]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout


def run(page):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout.splitlines()))


raw = run("raw")
hdr, units = raw[0], raw[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max"]
for vals in raw[2:]:
    print("=" * 100, file=out)
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print("%-70s %-12s %s" % (h, u, v), file=out)
    print("-- warp stall reasons (cycles per issued instruction)", file=out)
    for h, u, v in zip(hdr, units, vals):
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
        if m:
            try:
                if float(v) >= 0.01:
                    print("   %-28s %s" % (m.group(1), v), file=out)
            except ValueError:
                pass
src = run("source")
if len(src) > 2:
    h = src[1]
    ix = {k: i for i, k in enumerate(h)}
    data = [r for r in src[2:] if len(r) == len(h) and r[ix["# Samples"]].strip().isdigit()]
    tot = sum(int(r[ix["# Samples"]]) for r in data)

    def op(r):
        s = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
        return s.split()[0] if s else "?"
    print("-- hottest SASS instructions (samples of %d)" % tot, file=out)
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:16]:
        print("   %6s  exec %9s  %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:80]), file=out)
    ex = collections.Counter(int(r[ix["Instructions Executed"]]) for r in data)
    print("-- static instructions by execution count (top)", file=out)
    for cnt, n in ex.most_common(8):
        print("   executed %9d x : %4d instructions" % (cnt, n), file=out)
    c = collections.Counter()
    for r in data:
        c[op(r)] += int(r[ix["Instructions Executed"]])
    tote = sum(c.values())
    print("-- dynamic opcode mix (warp instructions, total %d)" % tote, file=out)
    for k, v in c.most_common(24):
        print("   %-28s %10d  %5.1f%%" % (k, v, 100.0 * v / tote), file=out)
