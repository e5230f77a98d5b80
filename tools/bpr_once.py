"""A few BPR step launches at the cfg3 shape (ncu target): python tools/bpr_once.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
U = torch.randn((1_000_000, 128), generator=g, device=dev); V = torch.randn((1_000_000, 128), generator=g, device=dev)
u, p, n = (torch.randint(0, 1_000_000, (65536,), device=dev) for _ in range(3))
gU, gV = torch.zeros_like(U), torch.zeros_like(V)
for _ in range(4):
    N.bpr_step(U, V, u, p, n, 1.0 / 65536, gU, gV, want_scores=False)
torch.cuda.synchronize(); print("ok")
