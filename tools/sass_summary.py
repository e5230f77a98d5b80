"""Per-kernel counts of the Blackwell-native SASS mnemonics of libdivrec_b200.so (cuobjdump -sass):
    python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "diversity-recommendations_b200", "libdivrec_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = [("UTCHMMA", r"\bUTCHMMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"), ("UTCBAR", r"\bUTCBAR\b"),
        ("SYNCS", r"\bSYNCS\b"), ("REDG.F32", r"\bREDG\.E\.ADD\.F32"), ("STG.256", r"\bSTG\.E\.ENL2\.256"), ("ATOMS", r"\bATOMS\b"),
        ("HMMA(legacy)", r"\bHMMA\.")]
print("# SASS summary of libdivrec_b200.so (final round-2 build): `cuobjdump -sass`, per-kernel counts of the Blackwell-native mnemonics")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,")
print("# REDG.F32 = red.global.add(.v4).f32 (BPR backward incl. the NVLink peer-memory scatter of bpr_step_kernel<*, true>),")
print("# STG.256 = 256-bit candidate-record store, HMMA(legacy) = mma.sync / wmma (none: the contraction is tcgen05 only).\n")
print(f"{'kernel':62s}" + f"{'instrs':>12s}" + "".join(f"{c:>13s}" for c, _ in COLS))
for blk in sass.split("\t\tFunction : ")[1:]:
    mangled = blk.split("\n", 1)[0].strip()
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = name.replace("divrec::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("divrec::", "")
    name = re.sub(r"\(.*$", "", name)
    lines = [l for l in blk.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]
    print(f"{name[:60]:62s}{len(lines):12d}" + "".join(f"{sum(1 for l in lines if re.search(rx, l)):13d}" for _, rx in COLS))
