import os, sys, torch
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
dev=torch.device("cuda:0"); g=torch.Generator(device=dev).manual_seed(1)
V=torch.randn((1_250_000,256),generator=g,device=dev).to(torch.bfloat16)
U=torch.randn((65536,256),generator=g,device=dev).to(torch.bfloat16)
for _ in range(2):
    N.score_topk(U,V,100); torch.cuda.synchronize()
print("ok")
