"""A/B timing of library builds on one cfg4 shard (1.25 M x 256 bf16 items, 65 536 queries, k = 100), in one process:
    python tools/ab_cfg4.py libdivrec_b200.so libdivrec_b200.x.so ..."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "diversity-recommendations_b200")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
from divrec import _native as N
libs = sys.argv[1:] or ["libdivrec_b200.so"]
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(1)
V = torch.randn((1_250_000, 256), generator=g, device=dev).to(torch.bfloat16)
U = torch.randn((65536, 256), generator=g, device=dev).to(torch.bfloat16)
handles = {name: N.load_library(os.path.join(PKG, name)) for name in libs}
res = {name: [] for name in libs}
ref = None
for rnd in range(3):
    for name in libs:
        N._lib = handles[name]
        N.set_profiling(True)
        for i in range(3):
            s, ii = N.score_topk(U, V, 100)
            torch.cuda.synchronize()
            if i >= 1: res[name].append(N.profile_read())
        if rnd == 0:
            if ref is None: ref = (ii.clone(), s.clone())
            else: print(name, "same items:", bool(torch.equal(ref[0], ii)), "same scores:", bool(torch.equal(ref[1], s)))
        N.set_profiling(False)
for name in libs:
    med = {key: sorted(r[key] for r in res[name])[len(res[name]) // 2] for key in res[name][0]}
    print(name, {key: round(v, 3) for key, v in med.items() if v >= 0})
