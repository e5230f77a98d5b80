"""First GPU probe: validates the tcgen05 scorer against torch on the B200 and times cfg2."""
import json, os, sys, time, traceback
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "diversity-recommendations_b200"))
from divrec import _native as N

out = {}
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False

def grid(n, d, g):
    return torch.randint(-8, 9, (n, d), generator=g).float() / 8

def ref_topk(S, k, seen=None):
    S = S.clone()
    if seen is not None:
        S[seen[:, 0], seen[:, 1]] = float("-inf")
    v, i = torch.sort(S, dim=1, descending=True, stable=True)
    return v[:, :k], i[:, :k]

def step(name, fn):
    try:
        t = time.time(); r = fn(); torch.cuda.synchronize(); out[name] = r; print(name, r, f"{time.time()-t:.2f}s", flush=True)
    except Exception as e:
        out[name] = "ERR " + repr(e); traceback.print_exc(); print(name, "ERR", flush=True)
        raise

g = torch.Generator().manual_seed(0)

def t_dense_grid():
    U = grid(300, 64, g).to(dev); V = grid(1000, 64, g).to(dev)
    S = N.score_dense(U, V); R = U.double() @ V.double().T
    return {"max_abs_diff": (S.double() - R).abs().max().item(), "exact": bool((S.double() == R).all().item())}
step("dense_grid_d64", t_dense_grid)

def t_dense_grid128():
    U = grid(257, 128, g).to(dev); V = grid(1300, 128, g).to(dev)
    S = N.score_dense(U, V); R = U.double() @ V.double().T
    return {"max_abs_diff": (S.double() - R).abs().max().item(), "exact": bool((S.double() == R).all().item())}
step("dense_grid_d128", t_dense_grid128)

def t_dense_gauss():
    U = torch.randn(512, 128, generator=g).to(dev); V = torch.randn(3000, 128, generator=g).to(dev)
    S = N.score_dense(U, V); R = U.double() @ V.double().T
    nrm = U.double().norm(dim=1)[:, None] * V.double().norm(dim=1)[None, :]
    S32 = U @ V.T
    return {"max_rel_to_norm": ((S.double() - R).abs() / nrm).max().item(),
            "torch_fp32_max_rel_to_norm": ((S32.double() - R).abs() / nrm).max().item(),
            "max_abs": (S.double() - R).abs().max().item()}
step("dense_gauss_d128", t_dense_gauss)

def t_dense_bf16():
    U = torch.randn(300, 256, generator=g).bfloat16().to(dev); V = torch.randn(2000, 256, generator=g).bfloat16().to(dev)
    S = N.score_dense(U, V); R = U.double() @ V.double().T
    nrm = U.double().norm(dim=1)[:, None] * V.double().norm(dim=1)[None, :]
    return {"max_rel_to_norm": ((S.double() - R).abs() / nrm).max().item()}
step("dense_bf16_d256", t_dense_bf16)

def make_seen(nU, nI, nnz, g):
    pairs = torch.stack([torch.randint(0, nU, (nnz,), generator=g), torch.randint(0, nI, (nnz,), generator=g)], 1)
    pairs = torch.unique(pairs, dim=0)  # sorted by (user,item)
    counts = torch.bincount(pairs[:, 0], minlength=nU)
    rowptr = torch.zeros(nU + 1, dtype=torch.int64); rowptr[1:] = torch.cumsum(counts, 0)
    return pairs, rowptr, pairs[:, 1].int()

def t_topk(nU, nI, d, k, nnz, kind):
    def f():
        if kind == "grid":
            U = grid(nU, d, g); V = grid(nI, d, g)
        else:
            U = torch.randn(nU, d, generator=g); V = torch.randn(nI, d, generator=g)
        pairs, rowptr, items = make_seen(nU, nI, nnz, g)
        s, i = N.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev))
        S = (U.double() @ V.double().T)
        rv, ri = ref_topk(S, k, pairs)
        i = i.cpu(); s = s.cpu()
        same_idx = (i == ri).all(dim=1)
        same_set = (i.sort(1).values == ri.sort(1).values).all(dim=1)
        return {"rows_idx_equal": int(same_idx.sum()), "rows_set_equal": int(same_set.sum()), "rows": nU,
                "max_score_abs_diff": (s.double() - rv).abs().max().item()}
    return f
step("topk_grid_small", t_topk(943, 1682, 64, 10, 90570, "grid"))
step("topk_gauss_small", t_topk(943, 1682, 64, 10, 90570, "gauss"))
step("topk_grid_mid", t_topk(1000, 26744, 128, 100, 150000, "grid"))
step("topk_gauss_mid", t_topk(1000, 26744, 128, 100, 150000, "gauss"))

def t_metrics():
    V = torch.randn(5000, 128, generator=g); recs = torch.stack([torch.randperm(5000, generator=g)[:100] for _ in range(64)])
    En = V / V.norm(dim=1, keepdim=True)
    ref = []
    for r in recs:
        G = En[r] @ En[r].T
        iu = torch.triu_indices(100, 100, 1)
        ref.append((1 - G[iu[0], iu[1]]).sum() / (100 * 99))
    ref = torch.stack(ref)
    a = N.ild_cosine(V.to(dev), recs.to(dev), N.ILD_SUMVEC).cpu()
    b = N.ild_cosine(V.to(dev), recs.to(dev), N.ILD_PAIRWISE).cpu()
    c = N.coverage_hist(recs.to(dev), 5000).cpu()
    cref = torch.bincount(recs.flatten(), minlength=5000).int()
    return {"ild_sumvec_rel": ((a - ref).abs() / ref.abs()).max().item(), "ild_pair_rel": ((b - ref).abs() / ref.abs()).max().item(),
            "cov_exact": bool((c == cref).all())}
step("metrics", t_metrics)

def t_bpr():
    nU, nI, d, B = 2000, 3000, 128, 4096
    U = torch.randn(nU, d, generator=g, requires_grad=True); V = torch.randn(nI, d, generator=g, requires_grad=True)
    u = torch.randint(0, nU, (B,), generator=g); p = torch.randint(0, nI, (B,), generator=g); n = torch.randint(0, nI, (B,), generator=g)
    pos = (U[u] * V[p]).sum(1); neg = (U[u] * V[n]).sum(1)
    loss = (-torch.nn.functional.logsigmoid(pos - neg)).mean(); loss.backward()
    Ud, Vd = U.detach().to(dev), V.detach().to(dev)
    ps, ng, ls, au = N.bpr_fwd(Ud, Vd, u.to(dev), p.to(dev), n.to(dev))
    gU = torch.zeros_like(Ud); gV = torch.zeros_like(Vd)
    N.bpr_bwd(Ud, Vd, u.to(dev), p.to(dev), n.to(dev), ps, ng, 1.0 / B, gU, gV)
    return {"loss_rel": abs(ls.item() / B - loss.item()) / loss.item(), "auc": au.item(), "auc_ref": int((pos >= neg).sum()),
            "gU_max_abs": (gU.cpu() - U.grad).abs().max().item(), "gV_max_abs": (gV.cpu() - V.grad).abs().max().item(),
            "gU_scale": U.grad.abs().max().item()}
step("bpr", t_bpr)

def t_time_cfg2():
    nU, nI, d, k = 138493, 26744, 128, 100
    U = torch.randn(nU, d, generator=g).to(dev); V = torch.randn(nI, d, generator=g).to(dev)
    nnz = 20000263
    us = torch.randint(0, nU, (nnz,), generator=g); its = torch.randint(0, nI, (nnz,), generator=g)
    key = torch.unique(us * nI + its)
    us, its = key // nI, key % nI
    counts = torch.bincount(us, minlength=nU); rowptr = torch.zeros(nU + 1, dtype=torch.int64); rowptr[1:] = torch.cumsum(counts, 0)
    rp, it = rowptr.to(dev), its.int().to(dev)
    res = {}
    for name, args in (("masked", (rp, it)), ("nomask", (None, None))):
        for _ in range(2):
            s, i = N.score_topk(U, V, k, *args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            s, i = N.score_topk(U, V, k, *args)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[name] = {"ms": ms, "users_per_s": nU / ms * 1e3}
    # ILD + coverage on the result
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N.ild_cosine(V, i); torch.cuda.synchronize()
    e0.record(); ild = N.ild_cosine(V, i); e1.record(); torch.cuda.synchronize(); res["ild_ms"] = e0.elapsed_time(e1)
    e0.record(); c = N.coverage_hist(i, nI); e1.record(); torch.cuda.synchronize(); res["cov_ms"] = e0.elapsed_time(e1)
    e0.record(); ild2 = N.ild_cosine(V, i, N.ILD_PAIRWISE); e1.record(); torch.cuda.synchronize(); res["ild_pairwise_ms"] = e0.elapsed_time(e1)
    res["ild_modes_max_diff"] = (ild - ild2).abs().max().item()
    # spot check 256 users against torch
    S = U[:256].double() @ V.double().T
    seen_mask_rows = torch.repeat_interleave(torch.arange(nU), counts)[: int(rowptr[256])]
    S[seen_mask_rows.to(dev), its[: int(rowptr[256])].to(dev)] = float("-inf")
    s, i = N.score_topk(U, V, k, rp, it)
    rv, ri = torch.sort(S, dim=1, descending=True, stable=True)
    res["spot_rows_idx_equal"] = int((i[:256] == ri[:, :k]).all(dim=1).sum())
    res["spot_rows_set_equal"] = int((i[:256].sort(1).values == ri[:, :k].sort(1).values).all(dim=1).sum())
    return res
step("time_cfg2", t_time_cfg2)
out["launches"] = N.launch_count()
json.dump(out, open("gpurun_out/probe0.json", "w"), indent=1)
print("DONE")
