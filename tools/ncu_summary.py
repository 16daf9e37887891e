"""Condense an `ncu --page raw --csv` export into a small markdown table for profiles/."""
import csv
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, data = rows[0], rows[2:]
cols = {"kernel": "Kernel Name", "duration": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum",
        "dram_write": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "grid": "launch__grid_size"}
units = dict(zip(hdr, rows[1]))
idx = {k: hdr.index(v) for k, v in cols.items() if v in hdr}
with open(dst, "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none` (cold-cache, serialised replays: compare shares, not absolutes).\n")
    f.write("Column units are the ones ncu exported (shown in brackets).\n\n")
    f.write("| " + " | ".join(k if k == "kernel" or not units[cols[k]] else f"{k} [{units[cols[k]]}]" for k in idx) + " |\n|" + "---|" * len(idx) + "\n")
    for r in data:
        vals = []
        for k, i in idx.items():
            v = r[i]
            if k == "kernel":
                v = v.replace("void ", "").replace("<unnamed>::", "")[:60]
            else:
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
            vals.append(v)
        f.write("| " + " | ".join(vals) + " |\n")
print("wrote", dst)
