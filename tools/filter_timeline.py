"""Coarse timeline of CTA 0's MMA-issuing thread: its clock at every schedule position of every user-block pair
(diagnostic build -DDIVREC_DIAG_TRACE2).  usage: python tools/filter_timeline.py libdivrec_b200.t2.so [mode]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
from oracle import divrec_oracle as O
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
nU, nI, d, k, nnz = 138493, 26744, 128, 100, 20000263
U, V = O.make_embeddings(nU, d, "gauss", g).to(dev), O.make_embeddings(nI, d, "gauss", g).to(dev)
seen = O.make_seen(nU, nI, nnz, g)
rp, it = O.seen_csr(seen, nU); rp, it = rp.to(dev), it.to(dev)
lib = N.load_library(os.path.join(ROOT, "diversity-recommendations_b200", sys.argv[1])); N._lib = lib
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
buf = torch.zeros((148 * 8 + 8 * 256,), dtype=torch.int64, device=dev)
lib.divrec_debug_filter_mode(mode)
for _ in range(2): N.score_topk(U, V, k, rp, it)
lib.divrec_debug_filter_stats(buf.data_ptr())
N.score_topk(U, V, k, rp, it); torch.cuda.synchronize()
lib.divrec_debug_filter_stats(None); lib.divrec_debug_filter_mode(0)
tr = buf[148 * 8:].view(8, 256).cpu()
n_sched = int((tr[0] > 0).sum())
t0 = int(tr[0, 0])
print("schedule positions per pair:", n_sched)
for p in range(8):
    if int(tr[p, 0]) == 0: break
    row = tr[p, :n_sched] - t0
    d = (row[1:] - row[:-1]).tolist()
    hint, main = d[:15], d[16:]
    main.sort()
    print(f"pair {p}: starts at {int(row[0]):9d}  hint phase (16 tiles) {int(row[16] - row[0]):7d}  main {int(row[-1] - row[16]):8d} cycles over {n_sched - 17} tiles: "
          f"median {main[len(main) // 2]:.0f}, p90 {main[int(len(main) * 0.9)]:.0f}, max {main[-1]:.0f}; tiles slower than 2x median: {sum(1 for x in main if x > 2 * main[len(main) // 2])}")
    med = main[len(main) // 2]
    print("         slow tiles (schedule position: cycles):", ", ".join(f"{i + 1}: {int(x)}" for i, x in enumerate(d) if x > 1.6 * med))
    if p + 1 < 8 and int(tr[p + 1, 0]) > 0:
        print(f"         gap from this pair's last tile start to the next pair's first: {int(tr[p + 1, 0] - tr[p, n_sched - 1])}")
