"""Turn gpurun_out/*.ncu-rep into small tracked summaries under profiles/ (the .ncu-rep files are
scratch).  Usage:  python tools/summarize_ncu.py r01 prof_bf16 prof_tf32 ...
Writes profiles/<round>_<name>.txt (key metrics per captured launch + top stall reasons) and updates
profiles/traffic.json (kernel -> dram bytes per launch) that bench.py reads for roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rnd, names = sys.argv[1], sys.argv[2:]
    out_dir = os.environ.get("B200_SUMMARY_DIR", os.path.join(ROOT, "profiles"))     # on the GPU box: gpurun_out/summaries (only gpurun_out/ travels back)
    os.makedirs(out_dir, exist_ok=True)
    traffic_path = os.path.join(out_dir, "traffic.json")
    if not os.path.exists(traffic_path) and os.path.exists(os.path.join(ROOT, "profiles", "traffic.json")):
        traffic_path_src = os.path.join(ROOT, "profiles", "traffic.json")
    else:
        traffic_path_src = traffic_path
    traffic = json.load(open(traffic_path_src)) if os.path.exists(traffic_path_src) else {}
    for name in names:
        rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
        rows = ncu_csv(rep, "raw")
        hdr, units = rows[0], rows[1]
        lines = [f"# ncu --set full --clock-control none --import-source on  ({name}.ncu-rep, summarised by tools/summarize_ncu.py)"]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            lines.append(f"\nkernel: {d.get('Kernel Name', '?')}")
            for k in KEYS:
                if k in d:
                    lines.append(f"  {k:75s} {d[k]} {u.get(k, '')}")
            if "dram__bytes_read.sum" in d:
                tb = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
                lines.append(f"  dram traffic per launch (read+write)                                        {tb / 1e6:.1f} MB")
                key = name.replace("prof_", "")
                log = os.path.join(ROOT, "gpurun_out", f"ncu_{key}.log")
                if os.path.exists(log):      # tools/run_one.py prints "<kind> <n> <library kernel name>"
                    last = [ln for ln in open(log).read().splitlines() if ln.startswith(key + " ")]
                    if last:
                        key = last[-1].split()[2] + "@" + last[-1].split()[1]
                traffic[key] = tb
        src = ncu_csv(rep, "source")
        if len(src) > 2:
            h = src[1]
            ix = {x: i for i, x in enumerate(h)}
            stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
            tot = {s: 0 for s in stalls}
            samples = 0
            per = []
            for r in src[2:]:
                if len(r) < len(h):
                    continue
                try:
                    n = int(r[ix["# Samples"]] or 0)
                except ValueError:          # a multi-kernel report repeats its header rows
                    continue
                samples += n
                for s in stalls:
                    try:
                        tot[s] += int(r[ix[s]] or 0)
                    except ValueError:
                        pass
                per.append((n, r[ix["Source"]].strip()[:80]))
            lines.append(f"\nwarp stall sampling, {samples} samples:")
            for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
                lines.append(f"  {s:28s} {100.0 * v / max(samples, 1):5.1f} %")
            lines.append("hottest SASS lines:")
            for n, s in sorted(per, key=lambda x: -x[0])[:12]:
                lines.append(f"  {n:7d}  {s}")
        path = os.path.join(out_dir, f"{rnd}_{name}.txt")
        open(path, "w").write("\n".join(lines) + "\n")
        print("wrote", path)
    try:
        head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        head = "?"
    traffic["_source"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of tools/run_one.py, "
                          f"summarised by tools/summarize_ncu.py; last updated {rnd} on top of git {head} (not measured inside bench.py)")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
