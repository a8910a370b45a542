"""Turns what scripts/gpu_round2.sh prof brought back in gpurun_out/ into the committed summaries under profiles/:
  python scripts/summarise_profiles.py r1d"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

# ---- launch list
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hdr, L = None, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] == "gpu__time_duration.sum":
            L.append((int(d["ID"]), d["Kernel Name"], float(d["Metric Value"].replace(",", "")), d["Grid Size"], d["Block Size"]))
idx = [i for i, l in enumerate(L) if "k_points" in l[1]]
out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none), config 4, one B200", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'emfb|k_' -c 200 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (scripts/gpu_round2.sh prof)",
       "(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes).", "",
       "Last complete steady-state frame of the capture (one `emf_engine_frame` call):", "",
       "| # | kernel | grid | block | us | share |", "|---|---|---|---|---|---|"]
fr = L[idx[-2]:idx[-1]]
tot = sum(l[2] for l in fr)
for l in fr:
    out.append(f"| {l[0]} | `{l[1][:70]}` | {l[3]} | {l[4]} | {l[2] / 1000:.1f} | {l[2] / tot * 100:.1f}% |")
agg = collections.OrderedDict()
for l in L[idx[-4]:idx[-1]]:
    k = l[1].split("(")[0][:60]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += l[2]
t = sum(v[1] for v in agg.values())
out += ["", "Aggregate over the last 3 frames:", "", "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, v in agg.items():
    out.append(f"| `{k}` | {v[0]} | {v[1] / 1000:.1f} | {v[1] / t * 100:.1f}% |")
open(os.path.join(P, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")

# ---- ncu --set full
rep = os.path.join(G, "prof_r2.ncu-rep" if os.path.exists(os.path.join(G, "prof_r2.ncu-rep")) else "prof_r1.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected"]
ix = {h: i for i, h in enumerate(hdr)}
ks = rows[2:8]
out = [f"# ncu --set full, config 4, one B200 ({tag})", "",
       "Command: `ncu --set full --clock-control none --import-source on -k regex:'k_integrate|k_brick|k_raycast|k_assoc|k_composite|k_depth' -s 24 -c 6 python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (scripts/gpu_round2.sh prof)",
       "", "| metric | " + " | ".join(r[ix["Kernel Name"]].split("(")[0].strip() for r in ks) + " |", "|---|" + "---|" * len(ks)]
for w in want:
    out.append(f"| {w} [{units[ix[w]]}] | " + " | ".join(r[ix[w]] for r in ks) + " |")
open(os.path.join(P, f"{tag}_ncu_full_summary.md"), "w").write("\n".join(out) + "\n")
traffic = {"source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full, config 4, one B200, per launch)",
           "source_hashes": {f: subprocess.run(["git", "hash-object", os.path.join(ROOT, "emfusion_b200", "csrc", f)], capture_output=True,
                                               text=True).stdout.strip() for f in ("integrate.cu", "raycast.cu")}}
for r in ks:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
    def mb(key):
        v, u = float(r[ix[key]].replace(",", "")), units[ix[key]]
        return int(v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u])
    traffic[name] = {"dram_bytes_read": mb("dram__bytes_read.sum"), "dram_bytes_write": mb("dram__bytes_write.sum"),
                     "warp_instructions": int(float(r[ix["smsp__inst_executed.sum"]].replace(",", ""))),
                     "duration_us_under_ncu": float(r[ix["gpu__time_duration.sum"]].replace(",", ""))}
json.dump(traffic, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
for a, b in (("bench_ours.json", f"{tag}_bench_ours.json"), ("bench_ref.json", f"{tag}_bench_reference.json"),
             ("bench_unchanged.json", f"{tag}_bench_unchanged_caller.json"), ("bench_cfg2.json", f"{tag}_bench_cfg2.json"),
             ("bench_cfg3.json", f"{tag}_bench_cfg3.json"), ("bench_cfg5.json", f"{tag}_bench_cfg5.json")):
    if not os.path.exists(os.path.join(G, a)):
        continue
    shutil.copy(os.path.join(G, a), os.path.join(P, b))
print(open(os.path.join(P, f"{tag}_ncu_full_summary.md")).read())
print(open(os.path.join(P, f"{tag}_launches.md")).read()[-900:])
