"""Condense an .ncu-rep (ncu --set full) into the small JSON kept under profiles/.
  python tools/ncu_summary.py <report.ncu-rep> <out.json> "<source / command description>"
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and keeps the metrics the
roofline discussion in DESIGN.md uses.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def main():
    rep, out, source = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    kernels = []
    for r in data:
        rec = dict(zip(header, r))
        k = {"Kernel Name": rec.get("Kernel Name"), "Grid Size": rec.get("Grid Size"), "Block Size": rec.get("Block Size")}
        for m in KEEP:
            if m in rec and rec[m] != "":
                k[m] = "%s %s" % (rec[m], units[header.index(m)])
        kernels.append(k)
    json.dump({"source": source, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"][:60], k.get("gpu__time_duration.sum"), "dram R", k.get("dram__bytes_read.sum"), "W", k.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main()
