"""Small shapes of every kernel of libdivrec_b200.so: a quick whole-library parity run, and the input of a
compute-sanitizer pass where the pool allows one (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
    CUDA_LAUNCH_BLOCKING=1 python tools/sanitize_small.py      # attributes a fault to the launch that caused it
Results are compared with the oracle where that is cheap."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))
from divrec import _native as N
from oracle import divrec_oracle as O

dev = torch.device("cuda:0")
only = sys.argv[1] if len(sys.argv) > 1 else "all"


def topk_case(nU, nI, d, k, nnz, kind, dtype=torch.float32, mode=None, metrics=False):
    gen = torch.Generator().manual_seed(nU + nI + d)
    U, V = O.make_embeddings(nU, d, kind, gen), O.make_embeddings(nI, d, kind, gen)
    seen = O.make_seen(nU, nI, nnz, gen) if nnz else torch.zeros((0, 2), dtype=torch.int64)
    rowptr, items = O.seen_csr(seen, nU)
    Ud, Vd = U.to(dev).to(dtype), V.to(dev).to(dtype)
    kw = {} if mode is None else {"mode": mode}
    if metrics:
        s, i, ild, cnt = N.score_topk_metrics(Ud, Vd, k, rowptr.to(dev), items.to(dev) if nnz else None, **kw)
    else:
        s, i = N.score_topk(Ud, Vd, k, rowptr.to(dev), items.to(dev) if nnz else None, **kw)
    torch.cuda.synchronize()
    if dtype == torch.float32 and kind == "grid":
        o_i, o_s = O.recommendations_dense(U, V, rowptr, items, k)
        assert torch.equal(i.cpu(), o_i) and torch.equal(s.cpu(), o_s), (nU, nI, d, k)
    print("ok topk", nU, nI, d, k, nnz, kind, dtype, mode, flush=True)
    return i


if only in ("all", "topk"):
    topk_case(300, 1682, 64, 10, 9000, "grid")                    # two-block filter, no hint phase
    topk_case(300, 6000, 128, 100, 30000, "grid", metrics=True)   # two-block filter with hint + ladder + fused metrics
    topk_case(200, 5000, 192, 64, 8000, "grid")                   # one-block filter (three k-blocks)
    topk_case(130, 4500, 128, 300, 5000, "grid")                  # large-k buffers
    topk_case(150, 5000, 256, 50, 0, "gauss", dtype=torch.bfloat16)
    topk_case(60, 3001, 72, 33, 4000, "grid", mode=N.MODE_SIMT)   # every row through the rescue kernels
if only in ("all", "metrics"):
    gen = torch.Generator().manual_seed(5)
    V = O.make_embeddings(2000, 64, "gauss", gen).to(dev)
    recs = torch.stack([torch.randperm(2000, generator=gen)[:20] for _ in range(300)]).to(dev)
    N.ild_cosine(V, recs); N.ild_cosine(V, recs, mode=N.ILD_PAIRWISE)
    N.ild_matrix(torch.rand(2000, 2000, device=dev), recs)
    N.ild_labels(torch.randint(0, 5, (2000,), device=dev), recs)
    c = N.coverage_hist(recs, 2000); N.entropy_from_counts(c); N.coverage_hist_ranks(recs, 2000)
    pos = O.make_seen(300, 2000, 4000, gen); rp, it = O.seen_csr(pos, 300)
    raw = torch.bincount(pos[:, 0], minlength=300).to(torch.int32)
    N.relevance_metrics(recs, rp.to(dev), it.to(dev), raw.to(dev))
    torch.cuda.synchronize(); print("ok metrics", flush=True)
if only in ("all", "bpr"):
    gen = torch.Generator().manual_seed(6)
    U, V = O.make_embeddings(500, 64, "gauss", gen).to(dev), O.make_embeddings(700, 64, "gauss", gen).to(dev)
    B = 2048
    u, p, n = (torch.randint(0, m, (B,), generator=gen).to(dev) for m in (500, 700, 700))
    pos, neg, ls, au = N.bpr_fwd(U, V, u, p, n)
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    N.bpr_bwd(U, V, u, p, n, pos, neg, 1.0 / B, gU, gV)
    rows = torch.empty((B, 64), device=dev)
    N.bpr_step(U, V, u, p, n, 1.0 / B, gU, gV, item_rows=rows)
    N.bpr_apply_rows(rows, p, n, gV)
    s = N.pair_scores(U, V, u, p); N.pair_scores_bwd(U, V, u, p, torch.ones_like(s), gU, gV)
    N.bpr_loss_fwd(pos, neg); N.auc_indicator(pos, neg); N.bpr_loss_bwd(pos, neg, torch.ones_like(pos))
    sp = O.make_seen(100, 700, 900, gen); rp, it = O.seen_csr(sp, 100)
    N.bpr_sample_triples(rp.to(dev), it.to(dev), rp.to(dev), it.to(dev), 700, 4, 1)
    torch.cuda.synchronize(); print("ok bpr", flush=True)
if only in ("all", "mmr"):
    gen = torch.Generator().manual_seed(7)
    V = O.make_embeddings(3000, 64, "gauss", gen).to(dev)
    U = O.make_embeddings(64, 64, "gauss", gen).to(dev)
    sc, cand = torch.topk(U @ V.T, 200, dim=1)
    N.mmr_rerank(V, cand, sc, 50, 0.5)
    torch.cuda.synchronize(); print("ok mmr", flush=True)
if only in ("all", "merge"):
    gen = torch.Generator().manual_seed(8)
    U, V = O.make_embeddings(100, 64, "grid", gen), O.make_embeddings(4000, 64, "grid", gen)
    parts = []
    for a in (0, 2000):
        s, i = N.score_topk(U.to(dev), V[a:a + 2000].to(dev), 20, item_id_offset=a)
        parts.append((s, i))
    gs, gi = torch.stack([x[0] for x in parts]), torch.stack([x[1] for x in parts])
    N.topk_merge(gs, gi)
    N.topk_merge_packed(torch.stack([N.topk_pack(s, i) for s, i in parts]))
    torch.cuda.synchronize(); print("ok merge", flush=True)
print("sanitize_small: done")
