"""Per-kernel SASS mnemonic counts of the in-tree library (no GPU needed): what proves the Blackwell-native data
movement (B200_PROFILING.md: UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDTM/STTM = tcgen05.ld/st, DMMA = FP64 tensor
cores; LDGSTS = Ampere-style cp.async).

    python profiles/sass_summary.py > profiles/r2_sass_summary.txt
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "scythe_jl_b200" / "libscythe_b200.so"
KEYS = ["DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "UTMALDG", "SYNCS", "LDTM", "STTM", "LDGSTS", "LDS", "STS", "LDG", "STG", "BAR", "ATOMG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*$", "", name).replace("void sb::", "")
            per[name] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            per[name][m.group(1)] += 1
    print(f"SASS mnemonic counts per kernel of {LIB.name} (cuobjdump -sass; static counts)")
    print(f"{'kernel':58s} " + " ".join(f"{k:>7s}" for k in KEYS))
    tot = Counter()
    for n, c in per.items():
        if not any(c[k] for k in ("DMMA", "UBLKCP", "SYNCS", "LDTM", "STTM", "LDGSTS")) and sum(c.values()) < 400:
            continue
        print(f"{n[:58]:58s} " + " ".join(f"{c[k]:7d}" for k in KEYS))
        tot.update(c)
    print(f"{'TOTAL (all kernels listed)':58s} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
    bulk = [n for n, c in per.items() if c["UBLKCP"]]
    tm = [n for n, c in per.items() if c["LDTM"]]
    print("\nkernels with bulk asynchronous copies (UBLKCP) + mbarrier (SYNCS):", ", ".join(bulk))
    print("kernels reading per-thread tables from Tensor Memory (LDTM):", ", ".join(tm))


if __name__ == "__main__":
    sys.exit(main())
