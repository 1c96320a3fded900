"""Aggregate an `ncu --page source --csv` dump (SASS view): instruction mix by opcode and stall reasons.
usage: python tools/ncu_src.py report.ncu-rep kernel_regex [launch_index]"""
import csv, subprocess, sys, collections, io
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "-s", skip, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:150])
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    if r[iN] == 'Instructions Executed' or r[0] == 'Kernel Name': break
    toks = r[iS].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "STG", "LDS", "STS", "MUFU")) and "." in op else "")
    n = int(r[iN]); ops[op] += n; tot += n; samp[op] += int(r[iSamp])
    for i, h in stall_cols: stalls[h] += int(r[i])
print("total warp instr", tot, " static SASS", len(rows) - 2)
for op, n in ops.most_common(28): print("  %-14s %10d %5.1f%%   samples %6d" % (op, n, 100.0 * n / tot, samp[op]))
ts = sum(stalls.values())
print("stall samples", ts)
for h, n in stalls.most_common(10): print("  %-26s %7d %5.1f%%" % (h, n, 100.0 * n / max(ts, 1)))
