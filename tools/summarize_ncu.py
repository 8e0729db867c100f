"""Summarise an `ncu --set full` report into a small text/JSON table for profiles/ (per captured launch)."""
import csv, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
def col(n): return hdr.index(n) if n in hdr else None
want = {
    "duration_us": "gpu__time_duration.sum", "grid": "launch__grid_size", "block": "launch__block_size", "regs": "launch__registers_per_thread",
    "smem_B": "launch__shared_mem_per_block_static", "waves_per_sm": "launch__waves_per_multiprocessor",
    "dram_read_B": "dram__bytes_read.sum", "dram_write_B": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "inst_executed": "smsp__inst_executed.sum",
    "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio", "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct", "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
}
units = rows[1]
res = []
for r in rows[2:]:
    d = {"kernel": r[col("Kernel Name")].replace("ngicp::<unnamed>::", "").split("(")[0].replace("void ", "")}
    for k, m in want.items():
        c = col(m)
        if c is None: continue
        v = r[c].replace(",", "")
        try: v = float(v)
        except ValueError: pass
        if k in ("dram_read_B", "dram_write_B") and isinstance(v, float):
            u = units[c]
            v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        if k == "duration_us" and isinstance(v, float):
            v *= {"ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}.get(units[c], 1)
        d[k] = v
    if "dram_read_B" in d and "dram_write_B" in d: d["dram_traffic_B"] = d["dram_read_B"] + d["dram_write_B"]
    res.append(d)
print(json.dumps(res, indent=1))
