"""Condense an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline discussion uses.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv [layer,layer,...]
"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    labels = sys.argv[3].split(",") if len(sys.argv) > 3 else []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "layer"] + [f"{n} [{units[i]}]" if units[i] else n for n, i in cols])
        for j, r in enumerate(data):
            name = labels[j % len(labels)] if labels else ""
            vals = [r[i] for _, i in cols]
            vals[0] = vals[0].split("(")[0].replace("unnamed>::", "")
            w.writerow([j, name] + vals)
    print(open(out).read())


if __name__ == "__main__":
    main()
