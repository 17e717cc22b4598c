"""Key metrics per kernel from an `ncu --set full` report (ncu -i X.ncu-rep --page raw --csv)."""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf")]
print("%-58s " % "kernel" + " ".join("%12s" % w[1] for w in want))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    name = name.split("(")[0][:58]
    vals = []
    for k, _ in want:
        v = r[idx[k]] if k in idx else ""
        u = units[idx[k]] if k in idx else ""
        try:
            f = float(v.replace(",", ""))
            v = "%.4g%s" % (f, {"Mbyte": "MB", "Gbyte": "GB", "Kbyte": "KB", "us": "us", "ms": "ms", "%": "", "byte": "B"}.get(u, ""))
        except ValueError:
            pass
        vals.append("%12s" % v[:12])
    print("%-58s " % name + " ".join(vals))
