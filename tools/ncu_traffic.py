"""Extract per-launch DRAM traffic of the hot kernels from ncu captures into profiles/r02_ncu_traffic.json (read by bench.py):
    python tools/ncu_traffic.py <workload>=<rep.ncu-rep> [...]      e.g.  cfg2=gpurun_out/prof_r2_final.ncu-rep cfg3=gpurun_out/prof_r2_bpr.ncu-rep"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"source": "ncu --set full --clock-control none captures summarised in profiles/r02_ncu_*.txt (tools/ncu_traffic.py)", "kernels": {}}
for arg in sys.argv[1:]:
    wl, rep = arg.split("=")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    def col(name): return hdr.index(name)
    def to_bytes(v, unit):
        f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        return float(v.replace(",", "")) * f
    units = rows[1]
    for r in rows[2:]:
        name = r[col("Kernel Name")]
        short = [k for k in ("score_filter2_kernel", "score_filter_kernel", "select_topk_kernel", "bpr_step_kernel", "mmr_warp_kernel", "rescue_kernel", "rescue_score_kernel", "row_mask_kernel") if k in name]
        if not short: continue
        e = {"dram_bytes_read": to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")]),
             "dram_bytes_write": to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")]),
             "gpu_time_ms_under_ncu": float(r[col("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col("gpu__time_duration.sum")], 1.0),
             "report": os.path.basename(rep)}
        out["kernels"].setdefault(wl, {})[short[0]] = e   # the last launch of a kernel in the report wins
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
