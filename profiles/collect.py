"""Turn the files a `ncu_final_cmd.sh` run left in gpurun_out/ into the tracked artefacts of profiles/:
roofline_r1.json (DRAM bytes per launch from the ncu raw pages), launches_r1.csv + launch_shares.txt (last step of
the launch list), copies of the bench lines and ncu pages."""
import csv
import json
import shutil
from collections import OrderedDict

import os
import sys

SRC, DST = "gpurun_out", "profiles"
RND = sys.argv[1] if len(sys.argv) > 1 else "r2"      # artefacts are named per round


def val(rows, k):
    h = rows[0]
    i = h.index(k)
    u = rows[1][i]
    v = float(rows[2][i].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1)


out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from `ncu --set full --clock-control none` "
                   "(profiles/ncu_%s_cmd.sh, half-size step: 204 pencils per launch, N=1000, k=7: one FULL-WIDTH launch of "
                   "each kernel -- back: first solve, factor: second solve, round: a round with the deflation sum); raw pages in "
                   "profiles/prof_*_%s_raw.csv" % (RND if RND != "r1" else "final", RND)}
for k in ("back", "factor", "round"):
    rows = list(csv.reader(open("%s/prof_%s_%s_raw.csv" % (SRC, k, RND))))
    r, w = val(rows, "dram__bytes_read.sum"), val(rows, "dram__bytes_write.sum")
    out["bsp_%s_kernel" % k] = {
        "dram_bytes": int(r + w), "dram_read": int(r), "dram_write": int(w), "pencils_per_launch": 204,
        "duration_ms": round(val(rows, "gpu__time_duration.sum"), 3),
        "fp64_pipe_active_pct": round(val(rows, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
        "registers_per_thread": int(val(rows, "launch__registers_per_thread")),
        "achieved_occupancy_pct": round(val(rows, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1)}
    print(k, out["bsp_%s_kernel" % k])
    for ext in ("details.txt", "raw.csv"):
        shutil.copy("%s/prof_%s_%s_%s" % (SRC, k, RND, ext), DST)
json.dump(out, open(DST + "/roofline_%s.json" % RND, "w"), indent=1)
for f in ("bench_%s_n1.json", "bench_%s_n1_explin.json", "bench_%s_reference.json", "configs_%s.jsonl", "tridiag_ab_%s.json"):
    if os.path.exists("%s/%s" % (SRC, f % RND)):
        shutil.copy("%s/%s" % (SRC, f % RND), DST)

rows = list(csv.reader(open(SRC + "/launches_all_%s.csv" % RND)))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[start + 1:]:
    if len(r) > vi and r[vi]:
        ms = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1}.get(r[ui], 1e-6)
        seq.append((r[ki].split("(")[0].replace("void ", ""), ms))
idx = [i for i, (k, _) in enumerate(seq) if "assemble" in k]
last = seq[idx[-1]:]
tot = sum(ms for _, ms in last)
agg = OrderedDict()
for k, ms in last:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
txt = ("one full-size step (408 pencils, 1 chunk stream), ncu --metrics gpu__time_duration.sum (cold cache, serialised)\n"
       + "\n".join("%-48s launches %3d  %8.3f ms  %5.1f %%" % (k, c, ms, 100 * ms / tot) for k, (c, ms) in agg.items())
       + "\ntotal %.3f ms\n" % tot
       + "round kernel launches in order (ms): " + ", ".join("%.3f" % ms for k, ms in last if "round" in k) + "\n"
       + "factor kernel launches in order (ms): " + ", ".join("%.3f" % ms for k, ms in last if "factor" in k) + "\n"
       + "back kernel launches in order (ms): " + ", ".join("%.3f" % ms for k, ms in last if "back" in k) + "\n")
open(DST + "/launch_shares_%s.txt" % RND, "w").write(txt)
print(txt)
with open(DST + "/launches_%s.csv" % RND, "w") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel", "gpu__time_duration_ms"])
    for i, (k, ms) in enumerate(last):
        w.writerow([i, k, "%.6f" % ms])
