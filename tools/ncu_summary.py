#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run here, no GPU needed).
  python tools/ncu_summary.py launches gpurun_out/launches_r1a.csv [last_n]   -> per-kernel time shares
  python tools/ncu_summary.py rep gpurun_out/prof_x.ncu-rep                   -> key metrics of a --set full capture
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg", "smsp__inst_executed.sum"]


def launches(path, last_n=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            name = r["Kernel Name"].split("(")[0].replace("sblk::", "").replace("void ", "")
            rows.append((name, r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
    if last_n == -1:  # last forward pass only: from the last clip-prep / stem launch (the first kernel of a forward) to the end
        first = "prep_clip" if any(r[0].startswith("prep_clip") for r in rows) else "stem_t_kernel"
        start = max(i for i, r in enumerate(rows) if r[0].startswith(first))
        rows = rows[start:]
    elif last_n:
        rows = rows[-last_n:]
    tot = sum(r[3] for r in rows)
    agg = {}
    for name, g, b, us in rows:
        a = agg.setdefault((name, g), [0, 0.0])
        a[0] += 1
        a[1] += us
    print(f"# {len(rows)} launches, total {tot:.1f} us (ncu per-launch times: cold-cache, serialised -> compare SHARES)")
    print("kernel,grid,launches,total_us,avg_us,share")
    for (name, g), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name},{g},{n},{us:.1f},{us / n:.2f},{us / tot:.4f}")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("# kernel:", d.get("Kernel Name", "?")[:100])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"{h},{units[i]},{vals[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    else:
        rep(sys.argv[2])
