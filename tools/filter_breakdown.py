"""Attribute the score+filter kernel's time: full kernel vs filter-without-stores vs pipeline-only (cfg2 shape)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
from oracle import divrec_oracle as O
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
nU, nI, d, k = 138493, 26744, 128, 100
U, V = O.make_embeddings(nU, d, "gauss", g).to(dev), O.make_embeddings(nI, d, "gauss", g).to(dev)
lib = N.load_library(os.path.join(ROOT, "diversity-recommendations_b200", sys.argv[1])) if len(sys.argv) > 1 else N.load_library()
N._lib = lib  # (an experiment build named on the command line)
N.set_profiling(True)
out = {}
for mode in (0, 1):
    lib.divrec_debug_filter_mode(mode)
    ms = []
    for _ in range(6):
        N.score_topk(U, V, k)
        torch.cuda.synchronize()
        ms.append(N.profile_read()["score_filter_ms"])
    out[f"mode{mode}"] = sorted(ms)[len(ms) // 2]
    print(mode, out[f"mode{mode}"], flush=True)
lib.divrec_debug_filter_mode(0)
buf = torch.zeros((148, 8), dtype=torch.int64, device=dev)
for mode in (0, 1):
    lib.divrec_debug_filter_mode(mode)
    lib.divrec_debug_filter_stats(buf.data_ptr())
    N.score_topk(U, V, k); torch.cuda.synchronize()
    b = buf.double().cpu()
    if float(b.abs().sum()) == 0.0:  # the two-block kernel's lean issue loop carries no counters (tools/filter_trace.py traces it)
        continue
    print("mode", mode, "MMA thread cycles: total %.0f  blocked on A %.0f  on acc %.0f  on B %.0f  (CTA 0: %s; CTA 147: %s)" % (
        b[:, 0].mean(), b[:, 1].mean(), b[:, 2].mean(), b[:, 3].mean(), b[0].tolist(), b[147].tolist()), flush=True)
    out[f"stats_mode{mode}"] = b.mean(0).tolist()
lib.divrec_debug_filter_stats(None)
lib.divrec_debug_filter_mode(0)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "filter_breakdown.json"), "w"))
