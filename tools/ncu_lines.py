"""Per-CUDA-source-line samples and executed warp instructions of one kernel of an .ncu-rep (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n] [file_substr]"""
import csv, io, subprocess, sys, collections

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fsub = sys.argv[4] if len(sys.argv) > 4 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
cur_file, hdr = None, None
agg = collections.OrderedDict()
first_fn = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name":
        if first_fn is None: first_fn = r[1]
        elif r[1] != first_fn: cur_file = None  # only the first matching kernel
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or cur_file is None or len(r) != len(hdr) or r[0] == "": continue
    i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
    try: s, e = int(r[i_s]), int(r[i_e])
    except ValueError: continue
    key = (cur_file.split("/")[-1], int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
    a[0] += s; a[1] += e
tot_s = sum(a[0] for a in agg.values()); tot_e = sum(a[1] for a in agg.values())
print(f"{first_fn}\n total samples {tot_s}, warp instructions {tot_e}")
rows = [(k, a) for k, a in agg.items() if fsub in k[0]]
print("-- in source order (lines with >= 0.3 % of samples or instructions)")
for k, a in sorted(rows):
    if a[0] >= 0.003 * tot_s or a[1] >= 0.003 * tot_e:
        print(f" {k[0]:>18s}:{k[1]:<4d} {100*a[0]/max(1,tot_s):5.1f}% smp {100*a[1]/max(1,tot_e):5.1f}% ins  {a[2]}")
