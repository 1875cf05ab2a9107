"""Digest of an .ncu-rep (read here, no GPU): one line per captured launch with the counters DESIGN.md quotes.
usage: python scripts/ncu_digest.py gpurun_out/x.ncu-rep [out.json]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = {"name": "Kernel Name", "us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "inst": "smsp__inst_executed.sum", "grid": "launch__grid_size",
        "waves": "launch__waves_per_multiprocessor", "smem_per_block": "launch__shared_mem_per_block_dynamic"}
out = []
for r in data:
    d = {}
    for k, h in want.items():
        if h in col:
            v = r[col[h]]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            d[k] = v
            if k in ("dram_rd", "dram_wr"):
                d[k + "_unit"] = units[col[h]]
    d["stalls_per_issue"] = {}
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.15:
                d["stalls_per_issue"][h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
    out.append(d)
for d in out:
    print(json.dumps(d))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
