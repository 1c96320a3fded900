"""Top stall locations of one kernel from an ncu report (SASS view), with a coarse position (pct of the way through the code).
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Kernel Name" or r[iN] == "Instructions Executed": break
    body.append(r)
n = len(body); tot = sum(int(r[iSamp]) for r in body)
# samples by decile of the code
dec = [0] * 20
for k, r in enumerate(body): dec[k * 20 // n] += int(r[iSamp])
print("samples by 5%% code position:", " ".join("%d" % d for d in dec), " total", tot)
idx = sorted(range(n), key=lambda k: -int(body[k][iSamp]))[:top]
for k in sorted(idx):
    r = body[k]
    st = sorted(((int(r[i]), h) for i, h in stall_cols), reverse=True)[:2]
    print("%5d (%4.1f%%) samp %4d exec %8s  %-60s %s" % (k, 100.0 * k / n, int(r[iSamp]), r[iN], r[iS].strip()[:60], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
