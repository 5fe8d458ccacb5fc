"""Summaries of the round's ncu captures: per-kernel key metrics of the --set full capture (raw page CSV) and the
launch-list shares.  Usage: python tools/ncu_extract.py <raw.csv> <launches.csv> <round tag>  -> profiles/<tag>_*"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw_csv, launches_csv, tag = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(raw_csv)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
summary = []
for r in rows[2:]:
    rec = {"kernel": r[ki]}
    for k in KEYS:
        if k in hdr:
            v = r[hdr.index(k)].replace(",", "")
            try:
                rec[k] = float(v)
            except ValueError:
                rec[k] = v
            rec[k + "__unit"] = units[hdr.index(k)]
    summary.append(rec)
json.dump(summary, open(os.path.join(ROOT, "profiles", tag + "_ncu_full_summary.json"), "w"), indent=1)
for rec in summary:
    print(rec["kernel"][:60])
    for k in KEYS:
        if k in rec:
            print("   %-82s %s %s" % (k, rec[k], rec[k + "__unit"]))
    if "ozaki_trmm" in rec["kernel"] and "dram__bytes_read.sum" in rec:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = rec["dram__bytes_read.sum"] * scale.get(rec["dram__bytes_read.sum__unit"], 1.0)
        wr = rec["dram__bytes_write.sum"] * scale.get(rec["dram__bytes_write.sum__unit"], 1.0)
        cand = int(rec["launch__grid_size"]) * 64
        out = {"kernel": "ozaki_trmm_kernel<0, 6> (K2o, 6 digit planes)", "dram_bytes_per_launch": rd + wr,
               "candidates_per_launch": cand, "dram_bytes_per_candidate": (rd + wr) / cand,
               "source": "profiles/%s_ncu_full_summary.json: dram__bytes_read.sum + dram__bytes_write.sum of ONE full-size launch "
                         "(grid %d CTAs x 64 candidates) inside `python bench.py --legs none --no-ask --no-cpu-baseline`, "
                         "ncu --set full --clock-control none (tools/ncu_capture_r02.sh)" % (tag, int(rec["launch__grid_size"]))}
        json.dump(out, open(os.path.join(ROOT, "profiles", tag + "_k2o_dram_bytes_per_launch.json"), "w"), indent=1)
        print("   -> %.0f B DRAM per candidate at %d candidates per launch" % (out["dram_bytes_per_candidate"], cand))
# launch list: share of the step per kernel
lr = [r for r in csv.reader(open(launches_csv)) if len(r) > 10]
h = lr[0]
k_i, v_i = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in lr[1:]:
    name = r[k_i].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[v_i].replace(",", ""))
tot = sum(a[1] for a in agg.values())
with open(os.path.join(ROOT, "profiles", tag + "_launch_shares.txt"), "w") as f:
    f.write("# %s: launches, total us, share of all kernel time (ncu per-launch times are serialised and cold-cache)\n" % launches_csv)
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = "%-70s %5d %12.1f %6.2f %%" % (name[:70], a[0], a[1] / 1e3, 100.0 * a[1] / tot)
        f.write(line + "\n")
        print(line)
