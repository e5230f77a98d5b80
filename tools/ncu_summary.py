"""Summarise an .ncu-rep here on the CPU box: headline raw metrics, stall mix, hottest SASS.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "lts__t_bytes.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:90])
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS: print(f"  {h:70s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
for si, s in enumerate(start):
    end = start[si + 1] - 1 if si + 1 < len(start) else len(lines)
    rd = list(csv.reader(io.StringIO("\n".join(lines[s:end]))))
    h = rd[0]; body = [r for r in rd[1:] if len(r) == len(h)]
    ci = {n: h.index(n) for n in h}
    tot_s = sum(int(r[ci["# Samples"]] or 0) for r in body); tot_i = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
    print(f"\n== kernel {si}: {len(body)} SASS instrs, samples {tot_s}, warp-instrs {tot_i}")
    st = collections.Counter()
    for r in body:
        for n in h:
            if n.startswith("stall_") and "Not Issued" not in n: st[n] += int(r[ci[n]] or 0)
    print("  stall mix:", ", ".join(f"{k[6:]} {100*v/max(1,tot_s):.1f}%" for k, v in st.most_common(8)))
    print("  hottest by samples:")
    for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]] or 0))[:top]:
        print(f"   {int(r[ci['# Samples']]):7d} smp {int(r[ci['Instructions Executed']]):10d} ex  {r[ci['Source']].strip()[:100]}")
    ops = collections.Counter()
    for r in body:
        ops[r[ci["Source"]].strip().split()[0 if not r[ci["Source"]].strip().startswith("@") else 1].split(".")[0]] += int(r[ci["Instructions Executed"]] or 0)
    print("  executed by opcode:", ", ".join(f"{k} {100*v/max(1,tot_i):.1f}%" for k, v in ops.most_common(14)))
