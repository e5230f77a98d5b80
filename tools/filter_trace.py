"""Timestamps of the MMA-issuing thread over 64 consecutive tiles.  Needs a diagnostic build of score_filter2.cu with the TRACE
hooks of the round-2 investigation (git history: the commit that introduced this tool); the release kernel has none.
usage: python tools/filter_trace.py libdivrec_b200.t.so [mode]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
from oracle import divrec_oracle as O
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
nU, nI, d, k = 138493, 26744, 128, 100
U, V = O.make_embeddings(nU, d, "gauss", g).to(dev), O.make_embeddings(nI, d, "gauss", g).to(dev)
lib = N.load_library(os.path.join(ROOT, "diversity-recommendations_b200", sys.argv[1])); N._lib = lib
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
buf = torch.zeros((148 * 8 + 64 * 8 + 64,), dtype=torch.int64, device=dev)
lib.divrec_debug_filter_mode(mode)
for _ in range(2): N.score_topk(U, V, k)
lib.divrec_debug_filter_stats(buf.data_ptr())
N.score_topk(U, V, k); torch.cuda.synchronize()
lib.divrec_debug_filter_stats(None); lib.divrec_debug_filter_mode(0)
tr = buf[148 * 8:148 * 8 + 512].view(64, 8).cpu()
prod = buf[148 * 8 + 512:].cpu()
names = ["top", "acc0", "B", "mma0", "commit0", "acc1", "mma1", "commit1"]
print("tile  period | " + " ".join(f"{n:>8s}" for n in names[1:]) + "   (cycles since the previous mark)")
for i in range(1, 64):
    row = tr[i]; prev = tr[i - 1]
    d = [int(row[j] - (row[j - 1] if j else prev[7])) for j in range(8)]
    print(f"{i:4d} {int(row[0] - prev[0]):7d} | " + " ".join(f"{x:8d}" for x in d[1:]) + f"   producer issued this tile {int(row[0] - prod[i])} cycles before the MMA thread reached it")
