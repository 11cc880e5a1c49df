"""Per-kernel SASS evidence for the Blackwell-native paths: `cuobjdump -sass fgt_b200/libfgt_sm100a.so`, counted per
entry function (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tile load / store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync (must be 0)). Runs without a GPU.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fgt_b200", "libfgt_sm100a.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UTCBAR", "UTCATOMSWS",
             "SYNCS", "HMMA", "HGMMA", "LDGSTS", "ATOMG", "REDG", "RED", "MUFU", "SHFL", "LDG", "STG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = res.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            cur["_instr"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    cur[mn] += 1
                    break
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (arch: {sorted(set(arch))})")
    print(f"# {'kernel':58s} {'instr':>6s} " + " ".join(f"{m:>7s}" for m in MNEMONICS[:14]))
    tot = collections.Counter()
    for k, c in res.items():
        tot.update(c)
        print(f"{k[:60]:60s} {c['_instr']:6d} " + " ".join(f"{c[m]:7d}" for m in MNEMONICS[:14]))
    print(f"{'TOTAL':60s} {tot['_instr']:6d} " + " ".join(f"{tot[m]:7d}" for m in MNEMONICS[:14]))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core instructions found"
    return 0


if __name__ == "__main__":
    sys.exit(main())
