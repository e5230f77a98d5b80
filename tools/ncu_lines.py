"""Attribute an ncu source-page (SASS) profile to CUDA source lines using nvdisasm -g line markers.
usage: python tools/ncu_lines.py <rep> <kernel-regex> <cubin> <function-substring> [top]
(The SASS of the profiled kernel and of the cubin are matched by instruction order.)"""
import csv, io, re, subprocess, sys, collections
rep, kre, cubin, fsub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
rd = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = rd[0]; body = [r for r in rd[1:] if len(r) == len(h)]
ci = {n: h.index(n) for n in h}
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section
infn = False; cur = None; mapping = []
for l in dis:
    if l.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", l):
        infn = fsub in l
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): mapping.append(cur)
print(f"ncu SASS rows {len(body)}  nvdisasm instrs {len(mapping)}")
n = min(len(body), len(mapping))
agg = collections.defaultdict(lambda: [0, 0])
for r, loc in zip(body[:n], mapping[:n]):
    agg[loc][0] += int(r[ci["# Samples"]] or 0); agg[loc][1] += int(r[ci["Instructions Executed"]] or 0)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
srcs = {}
for (loc, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    if loc is None: print(f"{100*v[1]/ti:5.1f}% instr {100*v[0]/ts:5.1f}% smp  <none>"); continue
    f, ln = loc
    if f not in srcs:
        try: srcs[f] = open(f"/root/repo/diversity-recommendations_b200/csrc/{f}").read().splitlines()
        except OSError: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
    print(f"{100*v[1]/ti:5.1f}% instr {100*v[0]/ts:5.1f}% smp  {f}:{ln}  {text}")
