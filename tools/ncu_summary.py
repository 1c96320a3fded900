"""Summarise an ncu --set full report: per kernel launch duration, DRAM bytes (read+write), DRAM/L2/SM throughput, occupancy,
registers. usage: python tools/ncu_summary.py report.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_inst", "sm__inst_executed_pipe_fma.sum": "fma_inst",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conflicts", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
}
idx = {h: i for i, h in enumerate(hdr)}
res = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    d = {"kernel": r[idx["Kernel Name"]].split("<")[0].split("(")[0].replace("void ", "").replace("ow::", "")}
    for m, k in want.items():
        if m in idx:
            try:
                v = float(r[idx[m]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[m]]
            if k == "duration": v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0); k2 = "duration_us"
            elif k in ("dram_read", "dram_write"): v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0); k2 = k + "_bytes"
            else: k2 = k
            d[k2] = v
    if "dram_read_bytes" in d: d["dram_total_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
    res.append(d)
for d in res:
    print(" ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in d.items()))
if len(sys.argv) > 2: json.dump(res, open(sys.argv[2], "w"), indent=1)
