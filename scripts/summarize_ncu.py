"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
  python scripts/summarize_ncu.py launches <csv> <out.md>
  python scripts/summarize_ncu.py raw <ncu-rep> <out.csv>      (selected metrics per profiled launch)
"""
import collections
import csv
import subprocess
import sys

METRICS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__cycles_elapsed.avg.per_second", "launch__shared_mem_per_block_dynamic"]


def launches(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for row in r:
        v = float(row[vi].replace(",", ""))
        v = v / 1000 if row[ui] == "ns" else (v * 1000 if row[ui] == "ms" else v)
        agg.setdefault(row[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary of `{path}`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                f"(cold-cache, serialised launches: compare SHARES, not absolutes)\n\ntotal {tot:.1f} us over "
                f"{sum(len(v) for v in agg.values())} launches\n\n| kernel | launches | mean us | min us | max us | share |\n"
                f"|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {min(v):.2f} | {max(v):.2f} | {100 * sum(v) / tot:.1f}% |\n")
    print(open(out).read())


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(m) for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for row in data:
            w.writerow([row[i] for i in idx])
    print(open(out).read()[:8000])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
