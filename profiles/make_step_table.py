"""Per-kernel table of ONE timestep from an ncu launch list.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv \
        --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-tcbl --no-e2e \
        --no-cpu-baseline --no-materialised
    python profiles/make_step_table.py gpurun_out/launches.csv profiles/<round>_step_kernels.json

The step taken is the last one in the capture: the launches after the previous step's spline solve up to and
including the spline solve that follows the last equation-set kernel.  Times under ncu are cold-cache and
serialised: use the SHARES, not the absolute values (bench.py's CUDA-event times are the absolute ones).
"""
from __future__ import annotations

import csv
import json
import re
import sys
from collections import OrderedDict

EQ_KERNELS = ("k_pointwise", "k_heightresolved_bl", "k_euler_test", "k_inv_z_advection")


def short(name: str) -> str:
    return re.sub(r"\(.*$", "", name).strip()


def main(src: str, dst: str, source_note: str = ""):
    rows = OrderedDict()
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rec = rows.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"])})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            rec["ms"] = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        elif m.startswith("dram__bytes"):
            gb = val * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit, 1e-9)
            rec["dram_read_GB" if "read" in m else "dram_write_GB"] = gb
        elif m.startswith("sm__pipe_fp64"):
            rec["fp64_pipe_pct"] = val
    launches = [rows[k] for k in sorted(rows)]
    eq = [i for i, r in enumerate(launches) if any(e in r["name"] for e in EQ_KERNELS)]
    if not eq:
        raise SystemExit("no equation-set kernel in the capture")
    is_solve = lambda r: "k_spline" in r["name"]  # noqa: E731
    last = eq[-1]
    lo = last
    while lo > 0 and not is_solve(launches[lo - 1]):
        lo -= 1
    hi = last
    while hi < len(launches) - 1 and not is_solve(launches[hi]):
        hi += 1
    while hi + 1 < len(launches) and is_solve(launches[hi + 1]):   # per-variable solves of one K2
        hi += 1
    step = launches[lo:hi + 1]
    total = sum(r["ms"] for r in step)
    table = OrderedDict()
    for r in step:
        t = table.setdefault(r["name"], {"launches": 0, "ms": 0.0, "share": 0.0, "dram_read_GB": 0.0, "dram_write_GB": 0.0,
                                         "fp64_pipe_pct": 0.0})
        t["launches"] += 1
        t["fp64_pipe_pct"] += r.get("fp64_pipe_pct", 0.0) * r["ms"]
        t["ms"] += r["ms"]
        t["dram_read_GB"] += r.get("dram_read_GB", 0.0)
        t["dram_write_GB"] += r.get("dram_write_GB", 0.0)
    for t in table.values():
        t["fp64_pipe_pct"] = round(t["fp64_pipe_pct"] / t["ms"], 2) if t["ms"] else 0.0
        t["share"] = round(t["ms"] / total, 4)
        t["ms"] = round(t["ms"], 4)
        t["dram_read_GB"] = round(t["dram_read_GB"], 4)
        t["dram_write_GB"] = round(t["dram_write_GB"], 4)
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    from bench import kernel_source_hash
    out = {"source": source_note or f"profiles/make_step_table.py {src}", "kernel_src_sha16": kernel_source_hash(),
           "one_step": table, "step_ms_under_ncu": round(total, 4),
           "launches_per_step": len(step),
           "dram_GB_per_step": round(sum(t["dram_read_GB"] + t["dram_write_GB"] for t in table.values()), 3)}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "one_step"}))
    for k, v in table.items():
        print(f"{v['ms']:9.3f} ms {100 * v['share']:5.1f}%  R {v['dram_read_GB']:7.3f} W {v['dram_write_GB']:7.3f} GB  fp64 {v['fp64_pipe_pct']:5.1f}%  x{v['launches']}  {k}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
