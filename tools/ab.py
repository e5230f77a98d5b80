"""A/B timing of library builds in ONE process on ONE GPU (box-to-box variation is a few percent, larger than most
single optimisations):  python tools/ab.py libdivrec_b200.so libdivrec_b200.expA.so ...   (paths relative to the package dir)
Prints the median phase times of the cfg2 step for every build, interleaved over several rounds."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "diversity-recommendations_b200")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
from divrec import _native as N
from oracle import divrec_oracle as O
libs = sys.argv[1:] or ["libdivrec_b200.so"]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
nU, nI, d, k, nnz = 138493, 26744, 128, 100, 20000263
U, V = O.make_embeddings(nU, d, "gauss", g).to(dev), O.make_embeddings(nI, d, "gauss", g).to(dev)
seen = O.make_seen(nU, nI, nnz, g)
rp, it = O.seen_csr(seen, nU); rp, it = rp.to(dev), it.to(dev)
counts = torch.zeros(nI, dtype=torch.int32, device=dev)
handles = {name: N.load_library(os.path.join(PKG, name)) for name in libs}
res = {name: [] for name in libs}
ref = None
for rnd in range(4):
    for name in libs:
        N._lib = handles[name]
        N.set_profiling(True)
        for i in range(6):
            counts.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s, ii, ild, _ = N.score_topk_metrics(U, V, k, rp, it, counts=counts)
            e1.record(); torch.cuda.synchronize()
            ph = N.profile_read(); ph["step_ms"] = e0.elapsed_time(e1)
            if i >= 2: res[name].append(ph)
        _s, _i, st = N.score_topk(U, V, k, rp, it, return_stats=True)
        if rnd == 0:
            if ref is None: ref = (ii.clone(), s.clone())
            else: print(name, "same items:", bool(torch.equal(ref[0], ii)), "same scores:", bool(torch.equal(ref[1], s)))
            print(name, "rescued", st)
        N.set_profiling(False)
for name in libs:
    med = {key: sorted(r[key] for r in res[name])[len(res[name]) // 2] for key in res[name][0]}
    print(name, {key: round(v, 4) for key, v in med.items() if v >= 0})
