"""Condense `ncu -i X.ncu-rep --page raw --csv` exports into the per-kernel table kept under profiles/.
usage: python tools/ncu_csv_summary.py a.csv [b.csv ...] > profiles/rNN_ncu_topkernels.txt"""
import csv, re, sys
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem")]
print("%-58s %-14s %-10s %s" % ("kernel", "grid", "block", "  ".join(n for _, n in COLS)))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(anonymous namespace\)::|void |s2c::|<unnamed>::", "", r[idx["Kernel Name"]]).split("(")[0]
        vals = []
        for c, _ in COLS:
            if c in idx:
                v, u = r[idx[c]], units[idx[c]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                vals.append(v + (u if u not in ("", "inst", "register/thread") else ""))
            else:
                vals.append("-")
        print("%-58s %-14s %-10s %s" % (name[:58], r[idx["Grid Size"]].replace(" ", ""), r[idx["Block Size"]].replace(" ", ""), "  ".join(vals)))
