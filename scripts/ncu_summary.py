"""Summarise gpurun_out/ ncu artefacts into profiles/ (tracked).
  python scripts/ncu_summary.py r01
Reads gpurun_out/launches.csv (gpu__time_duration per launch) and every gpurun_out/prof_*.ncu-rep
(ncu --set full) and writes profiles/<round>_launches.csv, profiles/<round>_kernels.csv and
profiles/<round>_summary.md."""
import csv
import glob
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")


def short(name):
    m = re.match(r"(?:void )?(?:mf6::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


md = [f"# ncu evidence, round {rnd}", "",
      "Workload: `bench.py` C2 (10x1000x1000, 1e7 cells, nja 6.796e7, block-multicolour ILU0 ordering, f64), B200.", ""]
lf = os.path.join(go, "launches.csv")
if os.path.exists(lf):
    rows = [r for r in csv.reader(l for l in open(lf) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    with open(os.path.join(out_dir, f"{rnd}_launches.csv"), "w") as f:
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in rows:
            k = short(r[ki])
            f.write(f"{r[0]},{k},\"{r[hdr.index('Grid Size')]}\",\"{r[hdr.index('Block Size')]}\",{r[vi]}\n")
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi])
    tot = sum(a[1] for a in agg.values())
    md += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400`)", "",
           "Per-launch times are cold-cache and serialised: compare SHARES.", "",
           "| kernel | launches | total us | mean us | share |", "|---|---:|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {c} | {t/1e3:.1f} | {t/c/1e3:.1f} | {100*t/tot:.1f}% |")
    md.append("")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
krows = []
for rep in sorted(glob.glob(os.path.join(go, "prof_*.ncu-rep"))):
    try:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    except Exception as e:
        print("skip", rep, e)
        continue
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"report": os.path.basename(rep), "kernel": short(r[hdr.index("Kernel Name")])}
        for w in want:
            if w in hdr:
                d[w] = r[hdr.index(w)]
                d[w + ".unit"] = units[hdr.index(w)]
        krows.append(d)
if krows:
    with open(os.path.join(out_dir, f"{rnd}_kernels.csv"), "w") as f:
        cols = ["report", "kernel"] + want
        f.write(",".join(cols) + "\n")
        for d in krows:
            f.write(",".join(str(d.get(c, "")) for c in cols) + "\n")
    md += ["## Full captures (`ncu --set full --clock-control none --import-source on`)", "",
           "| kernel | time us | DRAM read MB | DRAM write MB | traffic MB | DRAM % of ncu peak | warps active % | regs | L2 hit % |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---:|"]

    def mb(d, k):
        v = float(d.get(k, "nan"))
        u = d.get(k + ".unit", "")
        return v * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)

    def us(d):
        v = float(d.get("gpu__time_duration.sum", "nan"))
        u = d.get("gpu__time_duration.sum.unit", "us")
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)

    for d in krows:
        rd, wr = mb(d, "dram__bytes_read.sum"), mb(d, "dram__bytes_write.sum")
        md.append(f"| `{d['kernel']}` | {us(d):.1f} | {rd:.1f} | {wr:.1f} | {rd+wr:.1f} | "
                  f"{d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','')} | "
                  f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active','')} | "
                  f"{d.get('launch__registers_per_thread','')} | {d.get('lts__t_sector_hit_rate.pct','')} |")
    md.append("")
    import json
    traffic = {}
    for d in krows:
        t = traffic.setdefault(d["kernel"], [])
        t.append((mb(d, "dram__bytes_read.sum") + mb(d, "dram__bytes_write.sum")) * 1e6)
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()},
              open(os.path.join(out_dir, f"{rnd}_traffic.json"), "w"), indent=1)
open(os.path.join(out_dir, f"{rnd}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md))
