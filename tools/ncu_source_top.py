"""Top source lines by warp-stall samples of an `ncu --set full --import-source on` report:
   python tools/ncu_source_top.py report.ncu-rep   (run on the box: the reports are too big to bring back)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# find header
hi = next(i for i, l in enumerate(lines) if "Source" in l and "Sampling" in l)
rows = list(csv.reader(lines[hi:]))
hdr = rows[0]
print("columns:", hdr[:12], file=sys.stderr)
def col(name):
    for i, h in enumerate(hdr):
        if h.strip().startswith(name):
            return i
    return None
ci_src = col("Source"); ci_samp = col("Warp Stall Sampling (All"); ci_addr = col("Address")
agg = collections.Counter(); text = {}
tot = 0
for r in rows[1:]:
    if len(r) <= max(ci_src, ci_samp): continue
    try: v = float(r[ci_samp].replace(",", ""))
    except ValueError: continue
    key = r[ci_src].strip()[:110]
    agg[key] += v; tot += v
print("total samples", tot)
for k, v in agg.most_common(45):
    print("%6.2f%%  %s" % (100 * v / tot, k))
