"""Time the MMR re-rank at the cfg5 shape (C = 500, k = 100, d = 128) on a user sample: python tools/mmr_time.py [users]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
dev = torch.device("cuda:0")
nU = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
g = torch.Generator(device=dev).manual_seed(9)
V = torch.randn((26744, 128), generator=g, device=dev)
U = torch.randn((nU, 128), generator=g, device=dev)
sc, cand = N.score_topk(U, V, 500)
for lam in (0.5, 0.2):
    for _ in range(2):
        out, rel = N.mmr_rerank(V, cand, sc, 100, lam)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out, rel = N.mmr_rerank(V, cand, sc, 100, lam)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"lambda {lam}: {ms:.2f} ms for {nU} users -> {nU / ms * 1e3 / 1e6:.2f} M users/s; changed rows vs top-k: {int((out != cand[:, :100]).any(1).sum())}")
