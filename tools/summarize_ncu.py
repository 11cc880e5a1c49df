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
import subprocess
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
        fh.write("command: `ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled "
                 "-k regex:fgt:: -s 390 -c 130 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n")
        fh.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's "
                 "`kernels` block, not absolutes.\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / tot:.3f} |\n")
    print("wrote", f"{tag}_launches.md")

# ---------------------------------------------------------------- full-set capture
rp = os.path.join(GO, f"model_{tag}.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time_us"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
            ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
            ("launch__registers_per_thread", "regs"), ("sm__cycles_elapsed.avg.per_second", "sm_ghz")]
    units = rows[1]

    def conv(col, val):
        try:
            x = float(val.replace(",", ""))
        except ValueError:
            return val
        u = units[ix[col]]
        if col.startswith("dram__bytes"):
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            return round(x * scale, 3)
        if col.startswith("gpu__time"):
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
            return round(x * scale, 2)
        return round(x, 3)

    out = []
    traffic = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        rec = {}
        for col, name in want:
            if col in ix:
                rec[name] = conv(col, r[ix[col]]) if col not in ("Kernel Name", "Grid Size") else short(r[ix[col]])
        out.append(rec)
        t = traffic[rec["kernel"]]
        t[0] += 1
        t[1] += (rec.get("dram_read_MB", 0) + rec.get("dram_write_MB", 0)) * 1e6
    with open(os.path.join(PR, f"{tag}_ncu_kernels.csv"), "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=[n for _, n in want])
        w.writeheader()
        for rec in out:
            w.writerow(rec)
    tj = {}
    for k, (n, b) in traffic.items():
        key = "gemm_tc" if "gemm_tc" in k else ("flash" if "flash" in k else k)
        tj[key] = {"launches": n, "dram_bytes_total": b, "dram_bytes_per_launch": b / n,
                   "source": f"profiles/{tag}_ncu_kernels.csv (ncu --set full, one FGT forward 432x240 T=10)"}
    with open(os.path.join(PR, "traffic.json"), "w") as fh:
        json.dump(tj, fh, indent=1)
    print("wrote", f"{tag}_ncu_kernels.csv", "and traffic.json;", len(out), "launches")
