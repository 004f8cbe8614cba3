"""Condense an .ncu-rep (ncu --set full) into the per-kernel CSV kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "capture description" > profiles/x.csv"""
import csv
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def main():
    rep, desc = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow(["capture", "Kernel Name"] + KEYS)
    for r in data:
        out = [desc, r[cols["Kernel Name"]]]
        for k in KEYS:
            i = cols.get(k)
            out.append(f"{r[i]} {units[i]}".strip() if i is not None else "")
        w.writerow(out)


if __name__ == "__main__":
    main()
