"""Turns the round-2 ncu metric captures (tools/r2_ncu.sh -> gpurun_out/ncu_r2/*_metrics.csv) into the tracked summaries
under profiles/: one per-launch CSV and one per-kernel table per target, plus profiles/traffic.json for bench.py's
`roofline.traffic`.

    python tools/summarize_ncu2.py [tag]        (default r2)

ncu's per-launch numbers are cold-cache (caches flushed between replays) and serialised: compare SHARES with bench.py's
CUDA-event `kernels` block, not absolutes. DRAM GB/s here = (dram read + write bytes) / duration of the same launch.
"""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
GO = os.path.join(ROOT, "gpurun_out", f"ncu_{tag}")
PR = os.path.join(ROOT, "profiles")
HBM_PEAK = 6570.3
try:
    HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass

TARGETS = [
    ("model", "one FGT forward, 432x240 T=10 (bench.py step, eager launches)"),
    ("raft", "RAFT pair 480x864, 4 refinement iterations (tools/ncu_targets.py raft)"),
    ("lafc", "LAFC call 240x432 (tools/ncu_targets.py lafc)"),
    ("prop", "get_flowNN_gradient 240x432x10 (tools/ncu_targets.py prop)"),
    ("fill", "flow diffusion / region fill, 36 Laplace solves 240x432 (first 600 launches)"),
    ("poisson", "Poisson blend, 10 frames 240x432 (first 600 launches)"),
    ("splat", "flow_prop forward splat 2x64x240x432"),
    ("tail", "decoder final conv 64->3 + tanh, 10x240x432 (fgt_conv_tail)"),
]


def short(name):
    n = name.replace("void ", "").split("(")[0]
    return n.replace("fgt::", "")


def load(path):
    lines = open(path).read().splitlines()
    idx = [n for n, l in enumerate(lines) if l.startswith('"ID"')]
    if not idx:
        return []
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[idx[0]:]))))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u, name = r["Metric Unit"], r["Metric Name"]
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
    return list(per.values())


def main():
    os.makedirs(PR, exist_ok=True)
    md = [f"# ncu per-kernel summaries, round {tag[1:]} (B200, `--clock-control none`, cold-cache serialised launches)\n",
          "Captured by `tools/r2_ncu.sh` with `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput…,sm__throughput…,"
          "launch__registers_per_thread`; per-launch rows in `profiles/" + tag + "_ncu_<target>.csv`.\n",
          f"DRAM GB/s = (read + write bytes) / duration of the same launch; HBM peak {HBM_PEAK:.0f} GB/s (measured copy).\n"]
    traffic = {}
    for tgt, title in TARGETS:
        path = os.path.join(GO, f"{tgt}_metrics.csv")
        if not os.path.exists(path):
            continue
        per = load(path)
        if not per:
            continue
        fields = ["kernel", "time_us", "tensor_pipe_pct", "dram_read_MB", "dram_write_MB", "l2_pct", "sm_pct", "regs", "grid", "block"]
        with open(os.path.join(PR, f"{tag}_ncu_{tgt}.csv"), "w", newline="") as fh:
            w = csv.DictWriter(fh, fieldnames=fields, extrasaction="ignore")
            w.writeheader()
            for d in per:
                w.writerow(d)
        agg = collections.OrderedDict()
        for d in per:
            a = agg.setdefault(d["kernel"], dict(n=0, us=0.0, mb=0.0, tp=0.0, regs=d.get("regs", 0)))
            a["n"] += 1
            a["us"] += d.get("time_us", 0.0)
            a["mb"] += d.get("dram_read_MB", 0.0) + d.get("dram_write_MB", 0.0)
            a["tp"] += d.get("tensor_pipe_pct", 0.0) * d.get("time_us", 0.0)
        tot = sum(a["us"] for a in agg.values())
        md.append(f"\n## {tgt}: {title}\n\n{len(per)} launches, {tot:.0f} us serialised.\n\n"
                  "| kernel | launches | us total | share | us / launch | tensor pipe % (time-weighted) | DRAM MB / launch | DRAM GB/s | of HBM peak | regs |\n"
                  "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            gbs = a["mb"] * 1e6 / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
            md.append(f"| {k} | {a['n']} | {a['us']:.1f} | {a['us'] / tot:.3f} | {a['us'] / a['n']:.1f} | "
                      f"{(a['tp'] / a['us'] if a['us'] else 0):.1f} | {a['mb'] / a['n']:.2f} | {gbs:.0f} | {gbs / HBM_PEAK:.2f} | {int(a['regs'])} |")
        if tgt == "model":
            for k, a in agg.items():
                key = "gemm_tc" if "gemm_tc" in k else ("flash" if "flash" in k else k.replace("_kernel", ""))
                t = traffic.setdefault(key, dict(launches=0, dram_bytes_total=0.0, time_us_total=0.0, tp=0.0))
                t["launches"] += a["n"]
                t["dram_bytes_total"] += a["mb"] * 1e6
                t["time_us_total"] += a["us"]
                t["tp"] += a["tp"]
    if traffic:
        tj = {}
        for k, t in traffic.items():
            tj[k] = {"launches": t["launches"], "dram_bytes_total": t["dram_bytes_total"],
                     "dram_bytes_per_launch": t["dram_bytes_total"] / t["launches"], "time_us_total": round(t["time_us_total"], 1),
                     "tensor_pipe_pct_time_weighted": round(t["tp"] / t["time_us_total"], 2) if t["time_us_total"] else None,
                     "source": f"profiles/{tag}_ncu_model.csv (ncu --metrics, one FGT forward 432x240 T=10)"}
        with open(os.path.join(PR, "traffic.json"), "w") as fh:
            json.dump(tj, fh, indent=1)
    with open(os.path.join(PR, f"{tag}_ncu_kernels.md"), "w") as fh:
        fh.write("\n".join(md) + "\n")
    print("wrote", f"profiles/{tag}_ncu_kernels.md")


if __name__ == "__main__":
    main()
