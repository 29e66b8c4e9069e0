"""Turn the scratch ncu output of scripts/gpu_profile.sh (gpurun_out/prof/) into the tracked extracts under profiles/.
Usage (build container, no GPU): python scripts/summarise_profiles.py r02"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC, DST = os.path.join(ROOT, "gpurun_out", "prof"), os.path.join(ROOT, "profiles")


def ncu_rows(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[hi], rows[hi + 1:]


def short(name):
    name = name.replace("void dav2::", "").replace("dav2::", "")
    return name.split("(")[0]


# 1. launch list of one bench step: copy + per-kernel table + traffic json
step = os.path.join(SRC, f"launches_{R}_step.csv")
shutil.copy(step, os.path.join(DST, f"launches_{R}_step.csv"))
h, rows = ncu_rows(step)
ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
per = defaultdict(dict)
for r in rows:
    per[r[ix["ID"]]]["name"] = short(r[ix["Kernel Name"]])
    per[r[ix["ID"]]][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
agg = defaultdict(lambda: [0, 0.0, 0.0])
for v in per.values():
    a = agg[v["name"]]
    a[0] += 1
    a[1] += v.get("gpu__time_duration.sum", 0.0) / 1e6          # ns -> ms
    a[2] += (v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)) / 1e9   # B -> GB  (ncu prints bytes)
tot_ms, tot_gb = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
lines = ["| Kernel | Launches | Time | Share | DRAM bytes |", "|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {a[0]} | {a[1]:.2f} ms | {100 * a[1] / tot_ms:.1f} % | {a[2]:.2f} GB |")
lines.append(f"| **total** | **{sum(a[0] for a in agg.values())}** | **{tot_ms:.1f} ms** | | **{tot_gb:.1f} GB** |")
open(os.path.join(DST, f"step_table_{R}.md"), "w").write("\n".join(lines) + "\n")
gemm = [v for v in per.values() if "gemm" in v["name"] or "conv_halo" in v["name"]]
bp = [v for v in per.values() if v["name"].startswith("backproject")]
from bench import csrc_sha16  # noqa: E402
json.dump({"csrc_sha16": csrc_sha16(), "source": f"profiles/launches_{R}_step.csv",
           "gemm_tcgen05_kernel_bytes_per_launch": sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in gemm) / max(len(gemm), 1),
           "gemm_launches": len(gemm), "step_dram_gb": tot_gb, "step_ms_under_ncu": tot_ms,
           # the fused back-projection pass inside the step: its depth input was just written by the head conv and is partly
           # still in L2, so the DRAM bytes can be BELOW the 21 B/px algorithmic figure
           "backproject_kernel_bytes_per_launch": sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in bp) / max(len(bp), 1),
           "backproject_launches": len(bp)},
          open(os.path.join(DST, f"traffic_{R}.json"), "w"), indent=1)
print("\n".join(lines))

# 2. operator metrics: copy
shutil.copy(os.path.join(SRC, f"ncu_ops_metrics_{R}.csv"), os.path.join(DST, f"ncu_ops_metrics_{R}.csv"))

# 3. details pages of the full captures
for rep, out in ((f"prof_attn_{R}", f"ncu_attn_{R}.txt"), (f"prof_attn5477_{R}", f"ncu_attn5477_{R}.txt"), (f"prof_geom_{R}", f"ncu_geom_{R}.txt"),
                 (f"prof_conv_before_{R}", f"ncu_conv_before_{R}.txt"), (f"prof_conv_{R}", f"ncu_conv_{R}.txt")):
    p = os.path.join(SRC, rep + ".ncu-rep")
    if os.path.exists(p):
        txt = subprocess.run(["ncu", "-i", p, "--page", "details"], capture_output=True, text=True).stdout
        open(os.path.join(DST, out), "w").write(txt)
        print("wrote", out, len(txt.splitlines()), "lines")
