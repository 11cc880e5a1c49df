"""Turns the ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py <round-tag>      e.g. r1

Inputs : gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list of bench.py)
         gpurun_out/model_<tag>.ncu-rep  (ncu --set full of the gemm_tc / flash launches of one forward)
Outputs: profiles/<tag>_launches.md, profiles/<tag>_ncu_kernels.csv, profiles/traffic.json
"""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)


def short(name):
    return name.replace("void ", "").split("(")[0]


# ---------------------------------------------------------------- launch list
lp = os.path.join(GO, f"launches_{tag}.csv")
if os.path.exists(lp):
    lines = open(lp).read().splitlines()
    i = [n for n, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[i:]))))
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PR, f"{tag}_launches.md"), "w") as fh:
        fh.write(f"# ncu launch list, one bench.py step ({len(rows)} launches of fgt:: kernels, {tot:.0f} us serialised)\n\n")
        fh.write("command: `FGT_BENCH_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none "
                 "--kernel-name-base demangled -k regex:fgt:: -s 402 -c 134 --csv python bench.py --steps 2 --warmup 3 "
                 "--no-cpu-baseline` (eager launches so that one step = 134 launches; skip = 3 warm-up steps)\n\n")
        fh.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's "
                 "`kernels` block, not absolutes.\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / tot:.3f} |\n")
    print("wrote", f"{tag}_launches.md")

# ---------------------------------------------------------------- per-launch hardware metrics of one forward
# (ncu --metrics ... --csv long format: one row per (launch, metric); small enough to travel, unlike a
#  --set full .ncu-rep of all 125 launches)
mp = os.path.join(GO, f"model_metrics_{tag}.csv")
if os.path.exists(mp):
    lines = open(mp).read().splitlines()
    i = [n for n, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[i:]))))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"])})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        name = r["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            name = "dram_read_MB" if "read" in name else "dram_write_MB"
        elif name.startswith("gpu__time"):
            v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
            name = "time_us"
        else:
            name = {"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
                    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
                    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
                    "launch__registers_per_thread": "regs"}.get(name, name)
        d[name] = round(v, 3)
    fields = ["kernel", "time_us", "tensor_pipe_pct", "dram_read_MB", "dram_write_MB", "l2_pct", "sm_pct", "regs"]
    with open(os.path.join(PR, f"{tag}_ncu_kernels.csv"), "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=fields, extrasaction="ignore")
        w.writeheader()
        for d in per.values():
            w.writerow(d)
    traffic = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        k = d["kernel"]
        key = "gemm_tc" if "gemm_tc" in k else ("flash" if "flash" in k else k.replace("_kernel", ""))
        t = traffic[key]
        t[0] += 1
        t[1] += (d.get("dram_read_MB", 0) + d.get("dram_write_MB", 0)) * 1e6
        t[2] += d.get("time_us", 0)
        t[3] += d.get("tensor_pipe_pct", 0) * d.get("time_us", 0)
    tj = {}
    for k, (n, b, us, tp) in traffic.items():
        tj[k] = {"launches": n, "dram_bytes_total": b, "dram_bytes_per_launch": b / n, "time_us_total": round(us, 1),
                 "tensor_pipe_pct_time_weighted": round(tp / us, 2) if us else None,
                 "source": f"profiles/{tag}_ncu_kernels.csv (ncu --metrics, one FGT forward 432x240 T=10)"}
    with open(os.path.join(PR, "traffic.json"), "w") as fh:
        json.dump(tj, fh, indent=1)
    print("wrote", f"{tag}_ncu_kernels.csv and traffic.json;", len(per), "launches")
