"""Reduce `ncu -i X.ncu-rep --page raw --csv` output to the columns quoted in DESIGN.md / bench.py.
usage: python tools/ncu_summarise.py raw.csv out.csv [skip_first_n_launches]"""
import csv
import sys

COLS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
idx = [names.index(c) for c in COLS if c in names]
kn = names.index("Kernel Name")
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel"] + [names[i] for i in idx])
    w.writerow(["unit", ""] + [units[i] for i in idx])
    for j, r in enumerate(data[skip:]):
        if len(r) > kn:
            w.writerow([j, r[kn]] + [r[i] for i in idx])
