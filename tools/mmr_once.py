"""One MMR re-rank launch at the cfg5 shape on a small user count (ncu target)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
dev = torch.device("cuda:0")
nU = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device=dev).manual_seed(9)
V = torch.randn((26744, 128), generator=g, device=dev)
U = torch.randn((nU, 128), generator=g, device=dev)
sc, cand = torch.topk(U @ V.T, 500, dim=1, sorted=True)
for _ in range(2):
    N.mmr_rerank(V, cand, sc, 100, 0.5)
torch.cuda.synchronize()
print("ok")
