"""Where does the time of one cfg4 shard (1.25M items x 256 bf16, 65 536 queries) go?  full vs no-store vs pipeline-only."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nI, nQ = 1_250_000, 65536
V = torch.randn((nI, 256), generator=g, device=dev).to(torch.bfloat16)
U = torch.randn((nQ, 256), generator=g, device=dev).to(torch.bfloat16)
lib = N.load_library(); N.set_profiling(True)
for mode in (0,):
    lib.divrec_debug_filter_mode(mode)
    N.score_topk(U, V, 100); torch.cuda.synchronize()
    print(mode, N.profile_read(), flush=True)
lib.divrec_debug_filter_mode(0)
buf = torch.zeros((148, 8), dtype=torch.int64, device=dev)
lib.divrec_debug_filter_stats(buf.data_ptr())
N.score_topk(U, V, 100); torch.cuda.synchronize()
b = buf.double().cpu()
print('MMA thread: total %.0f  A %.0f  acc %.0f  B %.0f | tighten events/CTA %.0f  cycles/event %.0f  (sum over warps %.0f cycles/CTA)' % (b[:,0].mean(), b[:,1].mean(), b[:,2].mean(), b[:,3].mean(), b[:,4].mean(), (b[:,5].sum()/max(1,b[:,4].sum())), b[:,5].mean()))
lib.divrec_debug_filter_stats(None)
