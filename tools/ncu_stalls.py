"""Warp-stall sampling of ONE launch of an `ncu --set full --import-source on` report, per SASS instruction:
   python tools/ncu_stalls.py report.ncu-rep <launch index> [top N]
Prints the stall-reason totals and the hottest instructions (with executed counts): how the round-2 findings on the
CTA-pair GEMM epilogues were made (profiles/r02_gemm_stalls.txt)."""
import collections
import csv
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = raw.splitlines()
print(lines[0][:160])
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
rows = [r for r in rows[1:] if len(r) >= len(hdr) and r[0] != "Address"]
rows = rows[:len(rows) // 2]            # the page lists the function twice
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
samp = "Warp Stall Sampling (All Samples)"
n = sum(float(r[ix[samp]]) for r in rows)
tot = collections.Counter()
for r in rows:
    for c in stall_cols:
        tot[c] += float(r[ix[c]])
print("samples %d" % n)
for c, v in tot.most_common(8):
    print("  %-26s %5.1f%%" % (c, 100 * v / n))
order = sorted(range(len(rows)), key=lambda i: -float(rows[i][ix[samp]]))[:top_n]
for i in sorted(order):
    r = rows[i]
    reasons = sorted(((float(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% exec=%-9s %-70s %s" % (i, 100 * float(r[ix[samp]]) / n, r[ix["Instructions Executed"]],
                                              r[ix["Source"]].strip()[:70], " ".join("%s:%d" % (c, v) for v, c in reasons)))
