"""profiles/<round>_launches_summary.txt from gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum
--clock-control none pass over `bench.py --steps 2 --warmup 3 --no-extras`): per-kernel launch counts
and mean device time, and each of OUR kernels' share of a step.  Cold-cache, serialised timings:
compare shares, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name, v, unit = r[4].split("(")[0], float(r[-1]), r[-2]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit in ("ms", "msecond") else v
    agg.setdefault(name, []).append(v)
ours = {k: v for k, v in agg.items() if "b200::" in k and "add_inplace" not in k}
# The same kernels also run on quarter-height row blocks inside the host-pointer (e2e) leg; the device-path
# step is the full-size class of each kernel: launches within 25 % of that kernel's longest.
full = {k: [x for x in v if x >= 0.75 * max(v)] for k, v in ours.items()}
per_step = {k: sum(v) / len(v) for k, v in full.items()}
tot = sum(per_step.values())
lines = [f"# {os.path.basename(src)}: {len(rows)} launches; per-kernel mean device time (us)"]
for k, v in agg.items():
    lines.append(f"{len(v):5d} x {sum(v) / len(v):10.2f} us  (min {min(v):.2f}, max {max(v):.2f})  {k[:100]}")
lines.append("\n# one bench step (device path, N=1, default mode) = split_f16_rows (A) + col_absmax (B) + split_f16_cols (B) + gemm_tc;")
lines.append("# full-size launches only (the e2e leg runs the same kernels on 1024-row blocks); shares of the step:")
for k, v in per_step.items():
    lines.append(f"  {v:10.2f} us  {100 * v / tot:5.1f} %  {len(full[k])} launches  {k[:100]}")
out = os.path.join(ROOT, "profiles", f"{rnd}_launches_summary.txt")
open(out, "w").write("\n".join(lines) + "\n")
print(open(out).read())
