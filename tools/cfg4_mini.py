import os, sys, torch
#os.environ["DIVREC_DEBUG_NO_RESCUE"]="1"
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
dev=torch.device("cuda:0")
g=torch.Generator(device=dev).manual_seed(1)
nI=600_000; nQ=2048; k=100
V=torch.randn((nI,256),generator=g,device=dev).to(torch.bfloat16)
U=torch.randn((nQ,256),generator=g,device=dev).to(torch.bfloat16)
s,i,st=N.score_topk(U,V,k,return_stats=True); torch.cuda.synchronize(); print(st)
bad=(i[:,0]==-1002).nonzero().flatten()
print("failed rows", bad.numel())
S=(U.float()@V.float().T)
for r in bad[:6].tolist():
    T,m2,t0,t1,t2,t3,ne,nv,n0,n1,n2,n3,clo=s[r,:13].tolist()
    srt=S[r].sort(descending=True).values
    print(f"row {r}: T={T:.4f} m2={m2:.4f} thr={t0:.4f},{t1:.4f},{t2:.4f},{t3:.4f} ne={ne} nv={nv} counts={n0},{n1},{n2},{n3} c_lo={clo} | true 100th={srt[99]:.4f} 108th={srt[107]:.4f} 120th={srt[119]:.4f} count>=maxthr+m2: {int((S[r]>=max(t0,t1,t2,t3)+m2).sum())}")
