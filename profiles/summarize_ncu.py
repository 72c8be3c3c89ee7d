"""Text summary of an `ncu --set full --import-source on` capture, for committing under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/<round>_ncu_full_<kernel>.txt [launch index]

Per captured launch: the headline raw metrics (duration, registers, occupancy limits, DRAM bytes, L2 / L1 hit rates,
FP64 pipe and issue activity, shared-memory wavefronts and bank conflicts, warp-state stall ratios), then the source
lines ranked by sampled stalls with their three largest stall reasons (needs -lineinfo at compile time).
Runs here (no GPU needed): `ncu -i` only reads the report.
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def ncu_csv(rep: str, page: str) -> list[list[str]]:
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def main(rep: str, dst: str, which: int | None = None):
    rows = ncu_csv(rep, "raw")
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    out = []
    for li, r in enumerate(body):
        if which is not None and li != which:
            continue
        out.append(f"=== launch {li}: {r[col['Kernel Name']]}")
        for m in RAW:
            if m in col:
                out.append(f"{m} = {r[col[m]]} {units[col[m]]}".rstrip())
        stalls = []
        for n, i in col.items():
            if n.startswith(STALL_PREFIX) and n.endswith(STALL_SUFFIX):
                try:
                    stalls.append((float(r[i].replace(",", "")), n[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
                except ValueError:
                    pass
        out.append("stalls (warps per issue-active cycle): " + str(sorted(stalls, reverse=True)[:6]))
    # source page, CUDA lines correlated with SASS ("cuda,sass" view): sections per file, "Line No" rows carry the
    # per-line sums; the source text itself may contain commas that the CSV does not quote, so metrics are taken from
    # the right-hand end of each row
    out_src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                             capture_output=True, text=True).stdout
    fname, head, agg, total, kernel = "", None, {}, 0.0, ""
    sections = []
    for r in csv.reader(io.StringIO(out_src)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if r[0] == "Function Name":
            if r[1] != kernel and agg:
                sections.append((kernel, agg, total))
                agg, total = {}, 0.0
            kernel = r[1]
            continue
        if r[0] == "Line No":
            head = r
            continue
        if head is None or not r[0].strip().isdigit():
            continue
        nmet = len(head) - 4
        met = dict(zip(head[4:], r[-nmet:]))
        try:
            v = float(met.get("# Samples", "0").replace(",", "") or 0)
        except ValueError:
            continue
        if v <= 0:
            continue
        text = ",".join(r[1:len(r) - nmet - 2]).strip()[:100]
        st = {}
        for k, x in met.items():
            if k.startswith("stall_") and "(" not in k:
                try:
                    st[k] = float(x.replace(",", "") or 0)
                except ValueError:
                    pass
        key = f"{fname}:{int(r[0]):4d} {text}"
        if key in agg:
            agg[key][0] += v
            for k, x in st.items():
                agg[key][1][k] = agg[key][1].get(k, 0) + x
        else:
            agg[key] = [v, st]
        total += v
    if agg:
        sections.append((kernel, agg, total))
    for ti, (kname, ag, tot) in enumerate(sections):
        if which is not None and ti != which:
            continue
        out.append(f"--- source lines by stall samples: {kname[:90]} (total {tot:.0f})")
        for key, (v, st) in sorted(ag.items(), key=lambda kv: -kv[1][0])[:40]:
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            out.append(f"{v:8.0f} {100 * v / max(tot, 1):5.1f}% {key}  {[(k, int(x)) for k, x in top]}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
