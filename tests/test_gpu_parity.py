"""Parity of the CUDA path -- called through the C ABI (ctypes binding divrec._native) -- against the
oracle and against golden vectors produced by the real reference.

Bars (BASELINE.json north_star): top-k indices and coverage counts bit-exact (ties -> lowest item id);
scores, ILD and loss within 1e-5 relative in fp32.  "Relative" for a dot product means relative to
|u| |v| (single scores cancel to ~0).  Two tiers: grid-valued inputs are exact in
every arithmetic, so everything must be bit-exact; Gaussian inputs must agree within tolerance and
any index disagreement must sit on a near-tie of the oracle's scores.
"""
import math

import pytest
import torch

from conftest import load_golden
from oracle import divrec_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-5


def csr_dev(seen, nU, dev):
    rowptr, items = O.seen_csr(seen, nU)
    return rowptr.to(dev), items.to(dev)


def check_topk_against_oracle(items, scores, o_items, o_scores, U, V, exact):
    items, scores = items.cpu(), scores.cpu()
    if exact:
        assert torch.equal(items, o_items)
        assert torch.equal(scores, o_scores)
        return 0
    scale = (U.norm(dim=1).max() * V.norm(dim=1).max()).item()
    tol = REL * scale
    finite = torch.isfinite(o_scores)
    assert torch.equal(torch.isfinite(scores), finite)
    assert (scores[finite] - o_scores[finite]).abs().max().item() <= tol
    differ = (items != o_items).any(dim=1).nonzero().flatten().tolist()
    for u in differ:  # every disagreement must be a swap inside a near-tie, or a near-tie at the k boundary
        pos = (items[u] != o_items[u]).nonzero().flatten()
        for p in pos.tolist():
            s_here = o_scores[u, p]
            j = (o_items[u] == items[u, p]).nonzero().flatten()
            if j.numel():
                assert abs(float(o_scores[u, j[0]] - s_here)) <= 2 * tol, (u, p)
            else:
                assert abs(float(scores[u, p] - o_scores[u, -1])) <= 2 * tol, (u, p)
    return len(differ)


# ------------------------------------------------------------------------------------------------
# score + mask + top-k
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name,exact", [("cfg1_grid", True), ("cfg1_gauss", False), ("small_k100_gauss", False)])
def test_score_topk_vs_reference_golden(native, dev, name, exact):
    g = load_golden(name)
    k = int(g["k"])
    nU = g["U"].shape[0]
    rowptr, items = csr_dev(g["seen"], nU, dev)
    s, i = native.score_topk(g["U"].to(dev), g["V"].to(dev), k, rowptr, items)
    if exact:
        # the reference's order inside exact ties is arbitrary: compare score rows bit for bit and the
        # canonical order against the oracle
        assert torch.equal(s.cpu(), g["rec_scores"])
        o_i, o_s = O.recommendations_dense(g["U"], g["V"], *O.seen_csr(g["seen"], nU), k)
        assert torch.equal(i.cpu(), o_i)
    else:
        n_diff = check_topk_against_oracle(i, s, g["recs"], g["rec_scores"], g["U"], g["V"], False)
        assert n_diff <= max(1, nU // 100)


@pytest.mark.parametrize("nU,nI,d,k,nnz,kind", [
    (300, 1000, 64, 10, 3000, "grid"),
    (257, 1300, 128, 100, 20000, "grid"),     # nU, nI not multiples of the 128-wide tiles
    (129, 127, 32, 16, 500, "grid"),          # fewer items than one tile, d < 64
    (64, 5000, 100, 128, 10000, "grid"),      # d not a multiple of 64, k = the small-k class limit
    (500, 26744, 128, 100, 60000, "gauss"),   # cfg2 catalogue
    (200, 3000, 48, 7, 0, "gauss"),           # no interactions at all
    (1, 300, 16, 5, 10, "grid"),              # a single user
    (128, 4096, 8, 1, 0, "grid"),             # exact tile multiples, tiny d, k = 1
    (300, 6000, 192, 64, 9000, "gauss"),      # d = 192: three k-blocks -> the one-block-per-CTA kernel
    (150, 5000, 512, 32, 0, "gauss"),         # d = 512: the widest tensor-core shape (two accumulator stages)
    (140, 4000, 320, 20, 3000, "gauss"),      # d = 320: five k-blocks (every k-block count has its own issue loop)
    (140, 4000, 384, 20, 0, "grid"),          # six
    (140, 4000, 448, 20, 3000, "gauss"),      # seven
    (300, 26744, 128, 500, 40000, "gauss"),   # k = 500 (BASELINE.json config 5's candidate lists): the large-k buffers
    (130, 9000, 128, 500, 20000, "grid"),     # k = 500 with exact ties
    (90, 3000, 256, 300, 4000, "gauss"),      # k > 128 on the one-block kernel (d = 256)
    (70, 2000, 50, 512, 0, "grid"),           # k = MAX_K, scalar re-scoring path (d % 4 != 0), catalogue without a hint
    (1500, 60000, 128, 100, 150000, "gauss"), # 469 item tiles: threshold-ladder checkpoints every 59 tiles
])
def test_score_topk_vs_oracle(native, dev, nU, nI, d, k, nnz, kind):
    gen = torch.Generator().manual_seed(nU * 7 + nI)
    U, V = O.make_embeddings(nU, d, kind, gen), O.make_embeddings(nI, d, kind, gen)
    seen = O.make_seen(nU, nI, nnz, gen) if nnz else torch.zeros((0, 2), dtype=torch.int64)
    rowptr, items = O.seen_csr(seen, nU)
    o_i, o_s = O.recommendations_dense(U, V, rowptr, items, k)
    s, i = native.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev) if nnz else None)
    check_topk_against_oracle(i, s, o_i, o_s, U, V, kind == "grid")


def test_score_topk_without_mask_arguments(native, dev):
    gen = torch.Generator().manual_seed(11)
    U, V = O.make_embeddings(150, 64, "grid", gen), O.make_embeddings(700, 64, "grid", gen)
    o_i, o_s = O.recommendations_dense(U, V, None, None, 20)
    s, i = native.score_topk(U.to(dev), V.to(dev), 20)
    assert torch.equal(i.cpu(), o_i) and torch.equal(s.cpu(), o_s)


def test_ragged_rows_fewer_candidates_than_k(native, dev):
    """Users whose unmasked catalogue is smaller than k (the reference itself crashes on ragged lists,
    train/utils.py:77): padded with (-inf, -1)."""
    gen = torch.Generator().manual_seed(13)
    nU, nI, k = 40, 200, 50
    U, V = O.make_embeddings(nU, 32, "grid", gen), O.make_embeddings(nI, 32, "grid", gen)
    pairs = []
    for u in range(nU):
        n_seen = nI - (u % 60)  # user 0 has nothing left, user 1 one item, ...
        pairs += [(u, i) for i in torch.randperm(nI, generator=gen)[:n_seen].tolist()]
    seen = torch.tensor(pairs)
    rowptr, items = O.seen_csr(seen, nU)
    o_i, o_s = O.recommendations_dense(U, V, rowptr, items, k)
    s, i = native.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev))
    assert torch.equal(i.cpu(), o_i) and torch.equal(s.cpu(), o_s)
    assert (i.cpu()[0] == -1).all() and torch.isinf(s.cpu()[0]).all()


def test_all_scores_equal_ties_broken_by_lowest_id(native, dev):
    U, V = torch.ones(70, 64), torch.ones(900, 64)
    seen = torch.tensor([[u, i] for u in range(70) for i in range(0, u)])
    rowptr, items = O.seen_csr(seen, 70)
    s, i = native.score_topk(U.to(dev), V.to(dev), 10, rowptr.to(dev), items.to(dev))
    expect = torch.stack([torch.arange(u, u + 10) for u in range(70)])
    assert torch.equal(i.cpu(), expect) and (s.cpu() == 64.0).all()


def test_simt_mode_large_k(native, dev):
    """k = 400 through the exact SIMT kernel (rescue path: 512-entry sort buffer)."""
    gen = torch.Generator().manual_seed(43)
    nU, nI, d, k = 40, 5000, 64, 400
    U, V = O.make_embeddings(nU, d, "grid", gen), O.make_embeddings(nI, d, "grid", gen)
    seen = O.make_seen(nU, nI, 3000, gen)
    rowptr, items = O.seen_csr(seen, nU)
    o_i, o_s = O.recommendations_dense(U, V, rowptr, items, k)
    s, i = native.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev), mode=native.MODE_SIMT)
    assert torch.equal(i.cpu(), o_i) and torch.equal(s.cpu(), o_s)


@pytest.mark.parametrize("kind", ["grid", "gauss"])
def test_simt_mode_equals_oracle(native, dev, kind):
    """DIVREC_MODE_SIMT: every row through the exact SIMT scorer + radix select (the rescue kernel)."""
    gen = torch.Generator().manual_seed(41)
    nU, nI, d, k = 77, 3001, 72, 33
    U, V = O.make_embeddings(nU, d, kind, gen), O.make_embeddings(nI, d, kind, gen)
    seen = O.make_seen(nU, nI, 5000, gen)
    rowptr, items = O.seen_csr(seen, nU)
    o_i, o_s = O.recommendations_dense(U, V, rowptr, items, k)
    s, i, st = native.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev), mode=native.MODE_SIMT, return_stats=True)
    assert st["rescued_rows"] == nU
    check_topk_against_oracle(i, s, o_i, o_s, U, V, kind == "grid")


def test_degenerate_rows_are_rescued_exactly(native, dev):
    """Zero user vectors (every score ties at 0), duplicated item vectors and users that have seen almost
    everything: the tensor-core filter cannot certify these rows, the exact kernel must take over."""
    gen = torch.Generator().manual_seed(43)
    nU, nI, d, k = 200, 9000, 64, 50
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    U[::7] = 0.0                       # untrained users
    V[1000:6000] = V[999]              # 5000 identical (cold-start) items
    pairs = [O.make_seen(nU, nI, 20000, gen)]
    heavy = torch.arange(3, nU, 50)
    for u in heavy.tolist():           # users who have seen all but 30 items
        keep = torch.randperm(nI, generator=gen)[: nI - 30]
        pairs.append(torch.stack([torch.full_like(keep, u), keep], 1))
    seen = torch.cat(pairs)
    rowptr, items = O.seen_csr(seen, nU)
    o_i, o_s = O.recommendations_dense(U, V, rowptr, items, k)
    s, i, st = native.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev), return_stats=True)
    assert st["rescued_rows"] >= 29    # at least the zero rows
    check_topk_against_oracle(i, s, o_i, o_s, U, V, False)
    z = torch.arange(0, nU, 7)
    assert (s.cpu()[z][torch.isfinite(o_s[z])] == 0).all() and torch.equal(i.cpu()[z], o_i[z])


def test_hint_path_rarely_needs_the_rescue_kernel(native, dev):
    gen = torch.Generator().manual_seed(47)
    nU, nI, d, k = 4096, 26744, 128, 100
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    seen = O.make_seen(nU, nI, 600000, gen)
    rowptr, items = O.seen_csr(seen, nU)
    s, i, st = native.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev), return_stats=True)
    assert st["rescued_rows"] <= nU // 50
    sample = torch.arange(0, nU, 16)
    sub_rp = torch.zeros(sample.numel() + 1, dtype=torch.int64)
    segs = []
    for n_, u in enumerate(sample.tolist()):
        seg = items[rowptr[u]:rowptr[u + 1]]
        segs.append(seg)
        sub_rp[n_ + 1] = sub_rp[n_] + seg.numel()
    o_i, o_s = O.recommendations_dense(U[sample], V, sub_rp, torch.cat(segs), k)
    check_topk_against_oracle(i[sample.to(dev)], s[sample.to(dev)], o_i, o_s, U, V, False)


@pytest.mark.parametrize("nU,nI,d,k,kind,dt", [
    (943, 1682, 64, 10, "gauss", torch.float32),
    (600, 26744, 128, 100, "gauss", torch.float32),
    (300, 9000, 256, 100, "gauss", torch.bfloat16),
    (257, 5000, 100, 37, "grid", torch.float32),
])
def test_fused_metrics_equal_the_separate_kernels_and_the_oracle(native, dev, nU, nI, d, k, kind, dt):
    gen = torch.Generator().manual_seed(nU + nI + d)
    U, V = O.make_embeddings(nU, d, kind, gen).to(dt), O.make_embeddings(nI, d, kind, gen).to(dt)
    U[5] = 0  # a row that has to go through the rescue kernel (fused metrics must cover it too)
    seen = O.make_seen(nU, nI, 8 * nU, gen)
    rowptr, items = O.seen_csr(seen, nU)
    Ud, Vd, rp, it = U.to(dev), V.to(dev), rowptr.to(dev), items.to(dev)
    s0, i0 = native.score_topk(Ud, Vd, k, rp, it)
    s1, i1, ild, counts = native.score_topk_metrics(Ud, Vd, k, rp, it)
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    torch.testing.assert_close(ild, native.ild_cosine(Vd, i1, native.ILD_SUMVEC), rtol=REL, atol=1e-7)
    torch.testing.assert_close(ild.cpu(), O.ild_cosine(i1.cpu(), V.float()), rtol=REL, atol=1e-7)
    assert torch.equal(counts, native.coverage_hist(i1, nI))
    assert torch.equal(counts.cpu().long(), O.coverage_counts(i1.cpu(), nI))


def test_large_catalogue_tightens_lists_online(native, dev):
    """2M items: the sampled hint cannot resolve a 5e-5 quantile, every candidate list fills several times and is
    tightened in place by the filter kernel (validated bound X - 2 delta).  Results must still be exact and the
    rescue kernel must stay (almost) idle."""
    gen = torch.Generator().manual_seed(59)
    nU, nI, d, k = 128, 2_000_000, 32, 100
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    seen = O.make_seen(nU, nI, 2000, gen)   # ~16 seen items per user: need = k + seen + 8 must fit a 256-group list
    rowptr, items = O.seen_csr(seen, nU)
    s, i, st = native.score_topk(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev), return_stats=True)
    assert st["rescued_rows"] <= 2
    o_i, o_s = O.recommendations_dense(U, V, rowptr, items, k, chunk=32)
    check_topk_against_oracle(i, s, o_i, o_s, U, V, False)
    # the item-sharded route over the same catalogue (4 shards + merge) must agree bit for bit
    Ud, Vd, rp, it = U.to(dev), V.to(dev), rowptr.to(dev), items.to(dev)
    per = nI // 4
    gs, gi = [], []
    for sh in range(4):
        a, b = native.score_topk(Ud, Vd[sh * per:(sh + 1) * per].contiguous(), k, rp, it, item_id_offset=sh * per)
        gs.append(a)
        gi.append(b)
    ms, mi = native.topk_merge(torch.stack(gs), torch.stack(gi))
    assert torch.equal(mi, i) and torch.equal(ms, s)
    ps, pi = native.topk_merge_packed(torch.stack([native.topk_pack(a, b) for a, b in zip(gs, gi)]))
    assert torch.equal(pi, i) and torch.equal(ps, s)


def test_bf16_tables(native, dev):
    gen = torch.Generator().manual_seed(17)
    U = O.make_embeddings(200, 256, "gauss", gen).bfloat16()
    V = O.make_embeddings(5000, 256, "gauss", gen).bfloat16()
    o_i, o_s = O.recommendations_dense(U.float(), V.float(), None, None, 100)
    s, i = native.score_topk(U.to(dev), V.to(dev), 100)
    check_topk_against_oracle(i, s, o_i, o_s, U.float(), V.float(), False)
    Ug, Vg = O.make_embeddings(100, 256, "grid", gen), O.make_embeddings(3000, 256, "grid", gen)
    o_i, o_s = O.recommendations_dense(Ug, Vg, None, None, 100)
    s, i = native.score_topk(Ug.bfloat16().to(dev), Vg.bfloat16().to(dev), 100)
    assert torch.equal(i.cpu(), o_i) and torch.equal(s.cpu(), o_s)


def test_wide_bf16_tables_and_simt_beyond_512(native, dev):
    gen = torch.Generator().manual_seed(67)
    U = O.make_embeddings(100, 512, "gauss", gen).bfloat16()
    V = O.make_embeddings(4500, 512, "gauss", gen).bfloat16()
    o_i, o_s = O.recommendations_dense(U.float(), V.float(), None, None, 50)
    s, i = native.score_topk(U.to(dev), V.to(dev), 50)
    check_topk_against_oracle(i, s, o_i, o_s, U.float(), V.float(), False)
    # d > 512 is outside the tensor-core path: explicit error there, exact SIMT mode works
    U2, V2 = O.make_embeddings(20, 600, "gauss", gen), O.make_embeddings(700, 600, "gauss", gen)
    with pytest.raises(ValueError):
        native.score_topk(U2.to(dev), V2.to(dev), 10)
    o_i, o_s = O.recommendations_dense(U2, V2, None, None, 10)
    s, i = native.score_topk(U2.to(dev), V2.to(dev), 10, mode=native.MODE_SIMT)
    check_topk_against_oracle(i, s, o_i, o_s, U2, V2, False)


def test_item_sharded_merge_equals_unsharded(native, dev):
    """cfg4 in miniature: 8 catalogue shards scored separately (item_id_offset), lists merged by
    divrec_topk_merge == one unsharded call, bit for bit."""
    gen = torch.Generator().manual_seed(19)
    nU, nI, d, k, n_sh = 300, 8000, 64, 100, 8
    U, V = O.make_embeddings(nU, d, "grid", gen), O.make_embeddings(nI, d, "grid", gen)
    seen = O.make_seen(nU, nI, 20000, gen)
    rowptr, items = csr_dev(seen, nU, dev)
    Ud, Vd = U.to(dev), V.to(dev)
    full_s, full_i = native.score_topk(Ud, Vd, k, rowptr, items)
    per = nI // n_sh
    gs, gi = [], []
    for sh in range(n_sh):
        s, i = native.score_topk(Ud, Vd[sh * per:(sh + 1) * per].contiguous(), k, rowptr, items, item_id_offset=sh * per)
        gs.append(s)
        gi.append(i)
    ms, mi = native.topk_merge(torch.stack(gs), torch.stack(gi))
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    o_s, o_i = O.topk_merge(torch.stack(gs).cpu(), torch.stack(gi).cpu(), k)
    assert torch.equal(mi.cpu(), o_i) and torch.equal(ms.cpu(), o_s)
    # the packed exchange format (one 64-bit composite per pair) merges to the same bits
    ps, pi = native.topk_merge_packed(torch.stack([native.topk_pack(a, b) for a, b in zip(gs, gi)]))
    assert torch.equal(pi, full_i) and torch.equal(ps, full_s)


def test_packed_merge_with_padded_shard_lists(native, dev):
    """Shards smaller than k: every per-shard list ends in (-inf, -1) padding, which packs to 0 and must never win;
    negative scores, exact ties across shards and rows with fewer than k items overall are all in here."""
    gen = torch.Generator().manual_seed(23)
    nU, nI, d, k, n_sh = 70, 160, 16, 50, 8
    U, V = O.make_embeddings(nU, d, "grid", gen), O.make_embeddings(nI, d, "grid", gen)
    seen = O.make_seen(nU, nI, 20000, gen)  # ~83 % of the 160 items seen per user: rows end with fewer than k
    rowptr, items = csr_dev(seen, nU, dev)
    Ud, Vd = U.to(dev), V.to(dev)
    per = nI // n_sh
    gs, gi = [], []
    for sh in range(n_sh):
        s, i = native.score_topk(Ud, Vd[sh * per:(sh + 1) * per].contiguous(), k, rowptr, items, item_id_offset=sh * per)
        gs.append(s)
        gi.append(i)
    o_i, o_s = O.recommendations_dense(U, V, *O.seen_csr(seen, nU), k)
    assert (o_i < 0).any() and (torch.stack(gi) < 0).any()
    packed = [native.topk_pack(a, b) for a, b in zip(gs, gi)]
    assert all(torch.equal(pk.cpu(), O.topk_pack(a.cpu(), b.cpu())) for pk, a, b in zip(packed, gs, gi))
    ps, pi = native.topk_merge_packed(torch.stack(packed))
    assert torch.equal(pi.cpu(), o_i) and torch.equal(ps.cpu(), o_s)
    ms, mi = native.topk_merge(torch.stack(gs), torch.stack(gi))
    assert torch.equal(mi, pi) and torch.equal(ms, ps)


# ------------------------------------------------------------------------------------------------
# ILD / coverage / entropy
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name", ["cfg1_gauss", "cfg1_grid", "small_k100_gauss"])
def test_ild_and_coverage_vs_reference_golden(native, dev, name):
    g = load_golden(name)
    n = int(g["ild_users"])
    V, recs = g["V"].to(dev), g["recs"].to(dev)
    for mode in (native.ILD_SUMVEC, native.ILD_PAIRWISE):
        ild = native.ild_cosine(V, recs[:n], mode).cpu()
        torch.testing.assert_close(ild, g["ild_none"], rtol=REL, atol=1e-7)
    D = O.cosine_distance_matrix(g["V"]).to(dev)
    torch.testing.assert_close(native.ild_matrix(D, recs[:n]).cpu(), g["ild_none"], rtol=REL, atol=1e-7)
    counts = native.coverage_hist(recs, g["V"].shape[0]).cpu()
    nz = counts.nonzero().flatten()
    assert torch.equal(nz, g["cov_values"]) and torch.equal(counts[nz].long(), g["cov_counts"])
    ent = native.entropy_from_counts(counts.to(dev)).cpu()
    k = recs.shape[1]
    val = (float(ent[0]) - math.log(k)) / (math.log(g["V"].shape[0]) - math.log(k))
    assert val == pytest.approx(float(g["entropy"]), rel=REL)
    assert int(ent[1]) == nz.numel()


def test_ild_known_answers_and_padding(native, dev):
    V = torch.eye(8).repeat(1, 4)  # d = 32, orthogonal rows
    recs = torch.tensor([[0, 1, 2, 3], [4, 5, 6, 7]])
    for mode in (native.ILD_SUMVEC, native.ILD_PAIRWISE):
        out = native.ild_cosine(V.to(dev), recs.to(dev), mode).cpu()
        torch.testing.assert_close(out, torch.full((2,), 0.5))
        same = native.ild_cosine(torch.ones(5, 32, device=dev), torch.tensor([[0, 1, 2, 3]], device=dev), mode).cpu()
        assert abs(float(same[0])) < 1e-6
    gen = torch.Generator().manual_seed(23)
    Vg = O.make_embeddings(500, 64, "gauss", gen)
    r = torch.stack([torch.randperm(500, generator=gen)[:20] for _ in range(33)])
    r[::3, 15:] = -1  # ragged lists padded with -1
    ref = O.ild_cosine(r, Vg)
    for mode in (native.ILD_SUMVEC, native.ILD_PAIRWISE):
        torch.testing.assert_close(native.ild_cosine(Vg.to(dev), r.to(dev), mode).cpu(), ref, rtol=REL, atol=1e-7)


def test_ild_asymmetric_and_integer_matrix(native, dev):
    gen = torch.Generator().manual_seed(29)
    D = torch.rand(60, 60, generator=gen)
    recs = torch.stack([torch.randperm(60, generator=gen)[:9] for _ in range(21)])
    torch.testing.assert_close(native.ild_matrix(D.to(dev), recs.to(dev)).cpu(), O.ild_matrix(recs, D), rtol=REL, atol=0)
    torch.testing.assert_close(O.ild_matrix(recs[:4], D), O.ild_matrix_loop(recs[:4], D), rtol=REL, atol=0)
    labels = torch.randint(0, 3, (60,), generator=gen)
    Di = (labels[None, :] == labels[:, None]).int()
    want = O.ild_matrix(recs, Di)
    torch.testing.assert_close(native.ild_matrix(Di.to(dev), recs.to(dev)).cpu(), want, rtol=0, atol=1e-7)
    torch.testing.assert_close(native.ild_labels(labels.to(dev), recs.to(dev)).cpu(), want, rtol=0, atol=1e-7)


def test_coverage_sharded_accumulation(native, dev):
    gen = torch.Generator().manual_seed(31)
    recs = torch.randint(0, 999, (1001, 7), generator=gen)
    recs[5, 3] = -1
    counts = torch.zeros(999, dtype=torch.int32, device=dev)
    for part in recs.to(dev).chunk(3):  # user shards accumulate into one histogram
        native.coverage_hist(part.contiguous(), 999, counts)
    assert torch.equal(counts.cpu().long(), O.coverage_counts(recs, 999))


# ------------------------------------------------------------------------------------------------
# BPR
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name", ["bpr_small", "bpr_d64"])
def test_bpr_vs_reference_golden(native, dev, name):
    g = load_golden(name)
    U, V = g["U"].to(dev), g["V"].to(dev)
    u, p, n = g["u"].to(dev), g["p"].to(dev), g["n"].to(dev)
    B = u.numel()
    pos, neg, loss_sum, auc_sum = native.bpr_fwd(U, V, u, p, n)
    scale = float(g["U"].norm(dim=1).max() * g["V"].norm(dim=1).max())
    assert (pos.cpu() - g["pos"]).abs().max().item() <= REL * scale
    assert (neg.cpu() - g["neg"]).abs().max().item() <= REL * scale
    assert float(loss_sum) / B == pytest.approx(float(g["loss"]), rel=REL)
    assert int(auc_sum) / B == pytest.approx(float(g["auc"]), abs=1.5 / B)
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    native.bpr_bwd(U, V, u, p, n, pos, neg, 1.0 / B, gU, gV)
    for got, want in ((gU.cpu(), g["gU"]), (gV.cpu(), g["gV"])):
        assert (got - want).abs().max().item() <= REL * float(want.abs().max())
        assert torch.equal(got == 0, want == 0)  # untouched rows stay exactly zero
    # paired scorer (MatrixFactorization.forward)
    out = native.pair_scores(U, V, u, p).cpu()
    assert (out - g["pos"]).abs().max().item() <= REL * scale
    lv = native.bpr_loss_fwd(pos, neg).cpu()
    torch.testing.assert_close(lv, -torch.nn.functional.logsigmoid(g["pos"] - g["neg"]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["bpr_small", "bpr_d64"])
def test_bpr_one_kernel_step_vs_reference_golden(native, dev, name):
    """divrec_bpr_step (forward + backward in one pass over the rows) against the reference's loss.backward()."""
    g = load_golden(name)
    U, V = g["U"].to(dev), g["V"].to(dev)
    u, p, n = g["u"].to(dev), g["p"].to(dev), g["n"].to(dev)
    B = u.numel()
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    pos, neg, loss_sum, auc_sum = native.bpr_step(U, V, u, p, n, 1.0 / B, gU, gV)
    pos2, neg2, loss2, auc2 = native.bpr_fwd(U, V, u, p, n)
    assert torch.equal(pos, pos2) and torch.equal(neg, neg2) and int(auc_sum) == int(auc2)   # same summation order
    assert float(loss_sum) / B == pytest.approx(float(g["loss"]), rel=REL)
    assert int(auc_sum) / B == pytest.approx(float(g["auc"]), abs=1.5 / B)
    for got, want in ((gU.cpu(), g["gU"]), (gV.cpu(), g["gV"])):
        assert (got - want).abs().max().item() <= REL * float(want.abs().max())
        assert torch.equal(got == 0, want == 0)
    # accumulation semantics (a second call doubles the gradient) and the scores-off form
    _, _, l3, _ = native.bpr_step(U, V, u, p, n, 1.0 / B, gU, gV, want_scores=False)
    assert (gU.cpu() - 2 * g["gU"]).abs().max().item() <= 2 * REL * float(g["gU"].abs().max())
    assert float(l3) == pytest.approx(float(loss_sum), rel=1e-6)
    # sparse form of the item gradient (multi-GPU exchange): rows g_b U[u_b], applied by divrec_bpr_apply_rows
    rows = torch.empty((B, U.shape[1]), dtype=torch.float32, device=dev)
    gU2, gV2 = torch.zeros_like(U), torch.zeros_like(V)
    native.bpr_step(U, V, u, p, n, 1.0 / B, gU2, None, want_scores=False, item_rows=rows)
    native.bpr_apply_rows(rows, p, n, gV2)
    assert (gV2.cpu() - g["gV"]).abs().max().item() <= REL * float(g["gV"].abs().max())
    assert (gU2.cpu() - g["gU"]).abs().max().item() <= REL * float(g["gU"].abs().max())
    assert torch.equal(gV2.cpu() == 0, g["gV"] == 0)


def test_bpr_permutation_invariance(native, dev):
    g = load_golden("bpr_d64")
    U, V = g["U"].to(dev), g["V"].to(dev)
    perm = torch.randperm(g["u"].numel(), generator=torch.Generator().manual_seed(1))
    outs = []
    for idx in (torch.arange(perm.numel()), perm):
        u, p, n = g["u"][idx].to(dev), g["p"][idx].to(dev), g["n"][idx].to(dev)
        pos, neg, _, _ = native.bpr_fwd(U, V, u, p, n)
        gU, gV = torch.zeros_like(U), torch.zeros_like(V)
        native.bpr_bwd(U, V, u, p, n, pos, neg, 1.0 / perm.numel(), gU, gV)
        outs.append((gU.cpu(), gV.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert (a - b).abs().max().item() <= 1e-6 * float(a.abs().max())


def test_device_triple_sampler_structure_and_distribution(native, dev):
    """divrec_bpr_sample_triples against the sets the reference's sampler draws from (pinned by the golden of the real
    PairWiseDataset in tests/test_oracle_golden.py): iteration order and cross-product structure exactly, membership
    exactly, uniformity statistically (the random stream itself cannot match CPython's)."""
    from divrec.datasets.base_datasets import interactions_to_csr
    g = load_golden("pairwise_small")
    nU, nI, m = int(g["n_users"]), int(g["n_items"]), int(g["max_sampled"])
    inter, frozen = g["interactions"], g["frozen"]
    prp, pit = interactions_to_csr(inter.to(dev), nU)
    erp, eit = interactions_to_csr(torch.cat([inter, frozen]).to(dev), nU)
    pos_sets = [set(inter[inter[:, 0] == u, 1].tolist()) for u in range(nU)]
    excl_sets = [pos_sets[u] | set(frozen[frozen[:, 0] == u, 1].tolist()) for u in range(nU)]
    u, p, n = (t.cpu() for t in native.bpr_sample_triples(prp, pit, erp, eit, nI, m, seed=5))
    assert torch.equal(u, torch.arange(nU).repeat_interleave(m * m))                 # user-major, m^2 triples each
    P, N = p.view(nU, m, m), n.view(nU, m, m)
    assert torch.equal(P, P[:, :, :1].expand(-1, -1, m)) and torch.equal(N, N[:, :1, :].expand(-1, m, -1))  # positive-major cross product
    assert torch.equal(g["triples"][:, 0], u)                                        # same layout as the reference's iteration
    for uu in range(nU):
        assert set(P[uu, :, 0].tolist()) <= pos_sets[uu]
        assert not (set(N[uu, 0, :].tolist()) & excl_sets[uu])
        assert all(0 <= x < nI for x in N[uu, 0, :].tolist())
    again = native.bpr_sample_triples(prp, pit, erp, eit, nI, m, seed=5)
    other = native.bpr_sample_triples(prp, pit, erp, eit, nI, m, seed=6)
    assert torch.equal(again[1].cpu(), p) and torch.equal(again[2].cpu(), n) and not torch.equal(other[2].cpu(), n)
    # uniformity over many seeds: every positive / every allowed negative of a user equally often (5 sigma)
    S = 600
    pc = torch.zeros((nU, nI)); nc = torch.zeros((nU, nI))
    for seed in range(S):
        _, p2, n2 = native.bpr_sample_triples(prp, pit, erp, eit, nI, m, seed=1000 + seed)
        P2, N2 = p2.view(nU, m, m)[:, :, 0].cpu(), n2.view(nU, m, m)[:, 0, :].cpu()
        pc.scatter_add_(1, P2, torch.ones_like(P2, dtype=torch.float32))
        nc.scatter_add_(1, N2, torch.ones_like(N2, dtype=torch.float32))
    for uu in range(nU):
        for counts, allowed in ((pc[uu], sorted(pos_sets[uu])), (nc[uu], sorted(set(range(nI)) - excl_sets[uu]))):
            q = 1.0 / len(allowed)
            exp, sd = S * m * q, math.sqrt(S * m * q * (1 - q)) + 1e-9
            got = counts[allowed]
            assert float(counts.sum()) == S * m and float(got.sum()) == S * m        # nothing outside the allowed set
            assert float((got - exp).abs().max()) <= 5 * sd + 1
    # a user without positives, and one who has seen everything, are skipped by the shim (the reference raises)
    from divrec.datasets import PairWiseDataset, UserItemInteractionsDataset
    inter2 = torch.cat([inter[inter[:, 0] != 3], torch.stack([torch.full((nI,), 7), torch.arange(nI)], 1)])
    ds = PairWiseDataset(UserItemInteractionsDataset(interactions=inter2, number_of_users=nU, number_of_items=nI), max_sampled=m)
    u3, p3, n3 = ds.sample_triples(dev, seed=1)
    assert u3.numel() == (nU - 2) * m * m and 3 not in u3.tolist() and 7 not in u3.tolist()


def test_train_loop_on_the_device_sampler(native, dev):
    """pair_wise_train_loop on a feature-less dataset (where the reference's own loop raises in default_collate): triples
    from the device sampler, one-kernel BPR steps, loss falls over a few epochs."""
    from divrec.datasets import PairWiseDataset, UserItemInteractionsDataset
    from divrec.losses import LogSigmoidDifferenceLoss
    from divrec.metrics import AUCScore
    from divrec.models import MatrixFactorization
    from divrec.train import pair_wise_train_loop
    g = load_golden("pairwise_small")
    nU, nI = int(g["n_users"]), int(g["n_items"])
    torch.manual_seed(0)
    data = UserItemInteractionsDataset(interactions=g["interactions"], number_of_users=nU, number_of_items=nI)
    ds = PairWiseDataset(data, max_sampled=8)
    model = MatrixFactorization(nU, nI, 16).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    hist = [pair_wise_train_loop(ds, model, LogSigmoidDifferenceLoss(), opt, [AUCScore()], batch_size=64) for _ in range(6)]
    losses, aucs = [h[0] for h in hist], [h[1][0] for h in hist]
    assert all(math.isfinite(x) for x in losses) and losses[-1] < losses[0] and aucs[-1] > aucs[0]


# ------------------------------------------------------------------------------------------------
# the reference-facing Python API on top of the C ABI
# ------------------------------------------------------------------------------------------------


def test_shim_end_to_end_matches_reference_golden(native, dev):
    from divrec.datasets import RankingDataset, UserItemInteractionsDataset
    from divrec.losses import IntraListDiversityScore, LogSigmoidDifferenceLoss
    from divrec.metrics import AUCScore, EntropyDiversityScore
    from divrec.models import MatrixFactorization
    from divrec.train import bpr_step, get_model_recommendations, recommendations_score_loop

    g = load_golden("cfg1_gauss")
    nU, nI, d = 943, 1682, 64
    model = MatrixFactorization(nU, nI, d)
    model.load_state_dict({"user_embeddings.weight": g["U"], "item_embeddings.weight": g["V"]})
    model.to(dev)
    data = UserItemInteractionsDataset(interactions=torch.tensor([[nU - 1, nI - 1]]), number_of_users=nU, number_of_items=nI)
    frozen = UserItemInteractionsDataset(interactions=g["seen"], number_of_users=nU, number_of_items=nI)
    ds = RankingDataset(data, frozen=frozen)
    recs = get_model_recommendations(ds, model, 10)
    assert recs.dtype == torch.int64 and tuple(recs.shape) == (943, 10)
    assert (recs.cpu() != g["recs"]).any(dim=1).sum().item() <= 9
    ild = IntraListDiversityScore(item_embeddings=model.item_embeddings.weight, reduction="mean")
    ent = EntropyDiversityScore(dataset=data)
    v_ild, v_ent = recommendations_score_loop(ds, model, [ild, ent], 10)
    assert float(v_ild) == pytest.approx(float(g["ild_mean"]), rel=1e-4)
    assert float(v_ent) == pytest.approx(float(g["entropy"]), rel=1e-4)
    # model.forward + loss + autograd route == reference autograd
    b = load_golden("bpr_d64")
    m2 = MatrixFactorization(300, 200, 64)
    m2.load_state_dict({"user_embeddings.weight": b["U"], "item_embeddings.weight": b["V"]})
    m2.to(dev)
    u, p, n = b["u"].to(dev), b["p"].to(dev), b["n"].to(dev)
    pos, neg = m2(u, p), m2(u, n)
    loss = LogSigmoidDifferenceLoss()(pos, neg)
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(b["loss"]), rel=REL)
    assert float(AUCScore()(pos, neg)) == pytest.approx(float(b["auc"]), abs=2e-3)
    gU_generic = m2.user_embeddings.weight.grad.clone()
    gV_generic = m2.item_embeddings.weight.grad.clone()
    m2.zero_grad()
    loss2, _, _, _ = bpr_step(m2, u, p, n)
    loss2.backward()
    for got, want in ((gU_generic, b["gU"]), (m2.user_embeddings.weight.grad, b["gU"]),
                      (gV_generic, b["gV"]), (m2.item_embeddings.weight.grad, b["gV"])):
        assert (got.cpu() - want).abs().max().item() <= REL * float(want.abs().max())
    # the train loop's one-kernel step: loss + gradients accumulated straight into .grad
    from divrec.train import bpr_train_step
    m2.zero_grad()
    loss3, _, _, auc3 = bpr_train_step(m2, u, p, n)
    assert float(loss3) == pytest.approx(float(b["loss"]), rel=REL)
    for got, want in ((m2.user_embeddings.weight.grad, b["gU"]), (m2.item_embeddings.weight.grad, b["gV"])):
        assert (got.cpu() - want).abs().max().item() <= REL * float(want.abs().max())


@pytest.mark.parametrize("nU,nI,d,C,k,lams,table", [
    (24, 4000, 64, 120, 30, (1.0, 0.5, 0.2), "f32"),
    (10, 6000, 128, 500, 100, (0.5, 0.02), "f32"),    # BASELINE.json config 5: top-500 -> 100; lam = 0.02: diversity dominates
    (8, 3000, 256, 300, 40, (0.3,), "bf16"),          # bf16 table, two float4 chunks per lane
    (6, 2000, 36, 700, 25, (0.1,), "f32"),            # C > 512 (8 slots per lane), d not a multiple of 32
    (5, 1500, 30, 90, 20, (0.4,), "f32"),             # d % 4 != 0: the eager kernel
])
def test_mmr_rerank_vs_oracle(native, dev, nU, nI, d, C, k, lams, table):
    """Greedy MMR re-rank (no reference implementation: parity is against the oracle).  Greedy arg-max chains are
    sensitive to last-bit differences, so every pick is checked to be an (almost) maximiser of the oracle's
    objective given the picks made so far, and identical whenever the oracle's margin is clear."""
    from divrec.rerank import mmr_rerank
    gen = torch.Generator().manual_seed(61 + C)
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    if table == "bf16":
        V = V.bfloat16().float()
    if lams[-1] < 0.05:
        U = U * 0.05   # small relevance scores: the similarity term decides most picks
    cand, sc = O.recommendations_dense(U, V, None, None, C)
    cand[3, C - 20:] = -1   # a ragged candidate list
    sc[3, C - 20:] = float("-inf")
    Vd = V.to(dev).bfloat16() if table == "bf16" else V.to(dev)
    for lam in lams:
        want = O.mmr_rerank(cand, sc, V, k, lam)
        got, rel = mmr_rerank(Vd, cand.to(dev), sc.to(dev), k, lam)
        got, rel = got.cpu(), rel.cpu()
        assert torch.equal(got[:, 0], cand[:, 0])
        if lam == 1.0:
            assert torch.equal(got, cand[:, :k])
        En = (V / V.norm(dim=1, keepdim=True)).double()
        n_diff = 0
        for u in range(nU):
            ids = cand[u]
            valid = ids >= 0
            E = En[ids.clamp(min=0)]
            picked = []
            for r in range(k):
                j = int((ids == got[u, r]).nonzero()[0]) if got[u, r] >= 0 else -1
                assert j >= 0 and valid[j] and j not in picked
                if r > 0:
                    maxsim = (E @ E[picked].T).max(dim=1).values
                    obj = lam * sc[u].double() - (1 - lam) * maxsim
                    obj[picked] = float("-inf")
                    obj[~valid] = float("-inf")
                    assert float(obj[j]) >= float(obj.max()) - 1e-5     # an (almost) maximiser
                picked.append(j)
                assert float(rel[u, r]) == float(sc[u, j])
            n_diff += int((got[u] != want[u]).any())
        assert n_diff <= 2   # identical lists except where two objectives tie to ~1e-7
        if lam < 1.0:        # re-ranking must not reduce the intra-list diversity
            assert float(O.ild_cosine(got, V).mean()) >= float(O.ild_cosine(cand[:, :k], V).mean()) - 1e-6


def test_host_buffer_pipeline_equals_resident_path(native, dev):
    """divrec.train.host_recommend_and_score (pinned host buffers, chunked and overlapped copies) must return
    exactly what the device-resident fused call returns."""
    from divrec.train import host_recommend_and_score
    gen = torch.Generator().manual_seed(53)
    nU, nI, d, k = 40000, 5000, 64, 20   # more than one 18 944-user wave -> several chunks
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    seen = O.make_seen(nU, nI, 400000, gen)
    rowptr, items = O.seen_csr(seen, nU)
    s, i, ild, counts = native.score_topk_metrics(U.to(dev), V.to(dev), k, rowptr.to(dev), items.to(dev))
    for chunks in (1, 3):
        res = host_recommend_and_score(U.pin_memory(), V.pin_memory(), rowptr.pin_memory(), items.pin_memory(), k, dev, chunks=chunks)
        assert torch.equal(res["recommendations"], i.cpu())
        assert torch.equal(res["ild_per_user"], ild.cpu())
        assert torch.equal(res["coverage_counts"], counts.cpu())
        assert res["entropy_diversity"] == pytest.approx(float(O.entropy_diversity(i.cpu(), nI)), rel=1e-5)
        assert res["h2d_bytes"] == U.numel() * 4 + V.numel() * 4 + rowptr.numel() * 8 + items.numel() * 4
        assert res["d2h_bytes"] == nU * k * 8 + nU * 4 + nI * 4 + 8
    # 32-bit recommendation ids: the same lists, narrowed on the device, half of the table's bytes over the link
    res = host_recommend_and_score(U.pin_memory(), V.pin_memory(), rowptr.pin_memory(), items.pin_memory(), k, dev, chunks=3,
                                   index_dtype=torch.int32)
    assert res["recommendations"].dtype == torch.int32 and torch.equal(res["recommendations"].long(), i.cpu())
    assert res["d2h_bytes"] == nU * k * 4 + nU * 4 + nI * 4 + 8


def test_relevance_metrics_vs_reference_formulas(native, dev):
    """precision / recall / AP / NDCG against a direct restatement of the reference loops
    (precision_at_k.py:16-22, recall_at_k.py:16-22, average_precision_at_k.py:18-24, ndcg :16-23)."""
    from divrec.metrics import average_precision_at_k, normalized_discounted_cumulative_gain, precision_at_k, recall_at_k
    gen = torch.Generator().manual_seed(37)
    nU, nI, k = 50, 80, 12
    inter = torch.stack([torch.randint(0, nU, (600,), generator=gen), torch.randint(0, nI, (600,), generator=gen)], 1)
    inter[:nU, 0] = torch.arange(nU)
    recs = torch.stack([torch.randperm(nI, generator=gen)[:k] for _ in range(nU)])
    want = torch.zeros(4, nU)
    disc = torch.arange(1, k + 1)
    dlog = torch.log2(torch.arange(2, k + 2))
    for u in range(nU):
        positives = inter[inter[:, 0] == u, 1]
        rel = torch.isin(recs[u], positives).float()
        want[0, u] = rel.sum() / k
        want[1, u] = rel.sum() / positives.size(0)
        want[2, u] = torch.sum(torch.cumsum(rel, 0) / disc) / k
        want[3, u] = torch.sum(rel / dlog) / torch.sum(1 / dlog)
    r = recs.to(dev)
    got = torch.stack([precision_at_k(inter, r), recall_at_k(inter, r), average_precision_at_k(inter, r),
                       normalized_discounted_cumulative_gain(inter, r)]).cpu()
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# full size (BASELINE.json configs[1]): size-independent properties + sampled oracle rows
# ------------------------------------------------------------------------------------------------


def test_cfg2_full_size_properties(native, dev):
    nU, nI, d, k = 138493, 26744, 128, 100
    gen = torch.Generator().manual_seed(0)
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    seen = O.make_seen(nU, nI, 2_000_000, gen)
    rowptr, items = O.seen_csr(seen, nU)
    Ud, Vd, rp, it = U.to(dev), V.to(dev), rowptr.to(dev), items.to(dev)
    s, i = native.score_topk(Ud, Vd, k, rp, it)
    # sortedness in the canonical total order
    ok = (s[:, :-1] > s[:, 1:]) | ((s[:, :-1] == s[:, 1:]) & (i[:, :-1] < i[:, 1:]))
    assert bool(ok.all())
    # no duplicates, ids in range
    assert bool((i >= 0).all()) and bool((i < nI).all())
    assert bool((i.sort(dim=1).values.diff(dim=1) > 0).all())
    # masked items never appear: membership test of (user, item) keys
    keys = (torch.arange(nU, device=dev)[:, None] * nI + i).flatten()
    seen_keys = (seen[:, 0] * nI + seen[:, 1]).to(dev)
    assert not bool(torch.isin(keys, seen_keys).any())
    # reported scores are the dot products of the reported items (rescoring in fp64)
    rows = torch.arange(0, nU, 997, device=dev)
    ref = torch.einsum("ud,ukd->uk", Ud[rows].double(), Vd[i[rows]].double())
    scale = float(U.norm(dim=1).max() * V.norm(dim=1).max())
    assert (s[rows].double() - ref).abs().max().item() <= REL * scale
    # sampled rows against the oracle
    sample = torch.arange(0, nU, 541)
    sub_rowptr = torch.zeros(sample.numel() + 1, dtype=torch.int64)
    sub_items = []
    for n_, u in enumerate(sample.tolist()):
        seg = items[rowptr[u]:rowptr[u + 1]]
        sub_items.append(seg)
        sub_rowptr[n_ + 1] = sub_rowptr[n_] + seg.numel()
    o_i, o_s = O.recommendations_dense(U[sample], V, sub_rowptr, torch.cat(sub_items), k)
    check_topk_against_oracle(i[sample.to(dev)], s[sample.to(dev)], o_i, o_s, U, V, False)
    # user-sharded == unsharded (checksum of checksums): two halves concatenated give the same lists
    h = nU // 2
    s0, i0 = native.score_topk(Ud[:h].contiguous(), Vd, k, rp[:h + 1].contiguous(), it)
    assert torch.equal(i0, i[:h]) and torch.equal(s0, s[:h])
    # coverage: counts sum to U k and match a bincount; ILD modes agree and stay in [0, 1]
    counts = native.coverage_hist(i, nI)
    assert int(counts.sum()) == nU * k
    assert torch.equal(counts.long(), torch.bincount(i.flatten(), minlength=nI))
    a = native.ild_cosine(Vd, i, native.ILD_SUMVEC)
    b = native.ild_cosine(Vd[:, :], i[:4096].contiguous(), native.ILD_PAIRWISE)
    torch.testing.assert_close(a[:4096], b, rtol=REL, atol=1e-7)
    assert bool((a > 0).all()) and bool((a < 1).all())
    torch.testing.assert_close(a[sample.to(dev)].cpu(), O.ild_cosine(i[sample.to(dev)].cpu(), V), rtol=REL, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# list metrics against goldens produced by the real reference (tests/golden/make_golden_metrics.py)
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name", ["metrics_small", "metrics_k50"])
def test_relevance_pri_unfairness_vs_reference_golden(native, dev, name):
    from divrec.datasets import UserItemInteractionsDataset
    from divrec.datasets.storages import Features
    from divrec.losses.intra_list_diversity_score import IntraListBinaryUnfairnessScore
    from divrec.metrics import (PRI, MeanAveragePrecisionAtKScore, NDCGScore, PrecisionAtKScore, RecallAtKScore,
                                average_precision_at_k, avg_rank, normalized_discounted_cumulative_gain, precision_at_k, recall_at_k)
    g = load_golden(name)
    inter, recs, nI = g["interactions"], g["recs"].to(dev), int(g["nI"])
    for fn, key in ((precision_at_k, "precision"), (recall_at_k, "recall"), (average_precision_at_k, "ap"),
                    (normalized_discounted_cumulative_gain, "ndcg")):
        torch.testing.assert_close(fn(inter, recs).cpu(), g[key], rtol=1e-5, atol=1e-7)
    for cls, key in ((PrecisionAtKScore, "precision_mean"), (RecallAtKScore, "recall_mean"), (MeanAveragePrecisionAtKScore, "map_mean"),
                     (NDCGScore, "ndcg_mean")):
        assert cls().name  # snake-cased class name, as in the reference
        assert float(cls()(inter, recs)) == pytest.approx(float(g[key]), rel=1e-5, abs=1e-7)
    # avg_rank / coverage_hist_ranks: items bit-exact, mean positions within 1e-6
    items, ranks = avg_rank(recs, nI)
    assert torch.equal(items.cpu(), g["avg_rank_items"])
    torch.testing.assert_close(ranks.cpu(), g["avg_rank_values"], rtol=1e-6, atol=1e-6)
    counts, sums = native.coverage_hist_ranks(recs, nI)
    assert torch.equal(counts.cpu().long(), torch.bincount(g["recs"].flatten(), minlength=nI))
    pos = torch.arange(recs.shape[1]).expand_as(g["recs"]).flatten()
    assert torch.equal(sums.cpu(), torch.zeros(nI, dtype=torch.int64).index_add_(0, g["recs"].flatten(), pos))
    # PRI and the binary-unfairness class through the shim's own classes
    feats = Features(feature_names=["partition"], features=g["labels"].reshape(-1, 1))
    data = UserItemInteractionsDataset(interactions=inter, item_features=feats, number_of_users=recs.shape[0], number_of_items=nI)
    pri = PRI(dataset=data)
    assert pri.name == "pri"
    assert float(pri(inter, recs)) == pytest.approx(float(g["pri"]), rel=1e-4, abs=1e-6)
    unfair = IntraListBinaryUnfairnessScore(dataset=data, reduction="none")
    assert unfair.name == "intra_list_binary_unfairness_score"
    torch.testing.assert_close(unfair(inter, recs).cpu(), g["unfairness"], rtol=1e-6, atol=1e-7)
    assert torch.equal(unfair.get_distance_matrix(), (g["labels"][None] == g["labels"][:, None]).int())


# ------------------------------------------------------------------------------------------------
# row-gather kernels: any embedding_dim, ids out of range
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("d", [10, 30, 50])
def test_bpr_and_pair_scores_with_odd_embedding_dims(native, dev, d):
    """nn.Embedding takes any embedding_dim (the reference's configs use 10 / 30 / 50): rows that cannot move as 128-bit
    vectors take the scalar form of the same kernels."""
    gen = torch.Generator().manual_seed(d)
    nU, nI, B = 120, 150, 700
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    u, p, n = (torch.randint(0, m, (B,), generator=gen) for m in (nU, nI, nI))
    Ud, Vd, ud, pd, nd = U.to(dev), V.to(dev), u.to(dev), p.to(dev), n.to(dev)
    o_pos, o_neg, o_loss, o_auc = O.bpr_forward(U, V, u, p, n)
    o_gU, o_gV = O.bpr_backward(U, V, u, p, n)
    tol = 1e-5 * float(U.norm(dim=1).max() * V.norm(dim=1).max())
    s = native.pair_scores(Ud, Vd, ud, pd).cpu()
    assert (s - o_pos).abs().max().item() <= tol
    gU, gV = torch.zeros_like(Ud), torch.zeros_like(Vd)
    pos, neg, loss_sum, auc_sum = native.bpr_step(Ud, Vd, ud, pd, nd, 1.0 / B, gU, gV)
    assert (pos.cpu() - o_pos).abs().max().item() <= tol and (neg.cpu() - o_neg).abs().max().item() <= tol
    assert abs(float(loss_sum) / B - float(o_loss)) <= 1e-5 * abs(float(o_loss))
    assert (gU.cpu() - o_gU).abs().max().item() <= 1e-5 * float(o_gU.abs().max())
    assert (gV.cpu() - o_gV).abs().max().item() <= 1e-5 * float(o_gV.abs().max())
    # the model shim end to end (forward + autograd backward)
    from divrec.models import MatrixFactorization
    m = MatrixFactorization(nU, nI, d).to(dev)
    with torch.no_grad():
        m.user_embeddings.weight.copy_(Ud); m.item_embeddings.weight.copy_(Vd)
    out = m(ud, pd)
    out.sum().backward()
    ref_gU = torch.zeros(nU, d).index_add_(0, u, V[p])
    torch.testing.assert_close(m.user_embeddings.weight.grad.cpu(), ref_gU, rtol=1e-5, atol=1e-5)


def test_out_of_range_ids_are_skipped_and_reported(native, dev):
    """The reference's nn.Embedding raises IndexError for an id outside the table; here the triple is skipped (no
    out-of-range access) and the shim raises IndexError at its next synchronisation point."""
    gen = torch.Generator().manual_seed(3)
    nU, nI, d, B = 50, 60, 32, 64
    U, V = O.make_embeddings(nU, d, "gauss", gen).to(dev), O.make_embeddings(nI, d, "gauss", gen).to(dev)
    u, p, n = (torch.randint(0, m, (B,), generator=gen).to(dev) for m in (nU, nI, nI))
    native.bad_index_count(dev)  # reset
    u_bad, p_bad = u.clone(), p.clone()
    u_bad[5] = nU + 7
    p_bad[9] = -1
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    pos, neg, _, _ = native.bpr_step(U, V, u_bad, p_bad, n, 1.0 / B, gU, gV)
    good = torch.ones(B, dtype=torch.bool); good[5] = False; good[9] = False
    o_gU, o_gV = O.bpr_backward(U.cpu(), V.cpu(), u.cpu()[good], p.cpu()[good], n.cpu()[good], upstream=(B - 2) / B)
    assert (gU.cpu() - o_gU).abs().max().item() <= 1e-5 * float(o_gU.abs().max())
    assert (gV.cpu() - o_gV).abs().max().item() <= 1e-5 * float(o_gV.abs().max())
    assert float(pos[5]) == 0.0 and float(pos[9]) == 0.0
    assert native.bad_index_count(dev, reset=False) == 2
    with pytest.raises(IndexError):
        native.raise_on_bad_indices(dev)
    assert native.bad_index_count(dev) == 0
    native.pair_scores(U, V, u_bad, p)
    with pytest.raises(IndexError):
        native.raise_on_bad_indices(dev)


# ------------------------------------------------------------------------------------------------
# owner-sharded BPR step over peer memory (2 GPUs of one node; skipped on a 1-GPU box)
# ------------------------------------------------------------------------------------------------


def _owner_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from divrec import parallel
    gen = torch.Generator().manual_seed(0)
    nU, nI, d, B = 4000, 5003, 128, 6000
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    a, b = parallel.shard_range(nU, rank, world)
    us, ps, ns = [], [], []
    for r in range(world):  # every rank's triples, generated identically everywhere
        ra, rb = parallel.shard_range(nU, r, world)
        us.append(torch.randint(ra, rb, (B,), generator=gen)); ps.append(torch.randint(0, nI, (B,), generator=gen)); ns.append(torch.randint(0, nI, (B,), generator=gen))
    Ud, Vd = U.to(dev), V.to(dev)
    state = parallel.BprShardedState(Ud, Vd, world, rank, B, route="owner")
    for _ in range(2):  # gradients accumulate across steps
        loss_sum, auc_hits = parallel.bpr_step_sharded(state, us[rank].to(dev), ps[rank].to(dev), ns[rank].to(dev), 1.0 / (world * B))
    torch.cuda.synchronize()
    ua, pa, na = torch.cat(us), torch.cat(ps), torch.cat(ns)
    _, _, o_loss, o_auc = O.bpr_forward(U, V, ua, pa, na)
    o_gU, o_gV = O.bpr_backward(U, V, ua, pa, na)
    lo, hi = state.owned_range()
    ok = abs(float(loss_sum) / (world * B) - float(o_loss)) <= 1e-5 * abs(float(o_loss)) and abs(float(auc_hits) / (world * B) - float(o_auc)) <= 1e-6
    ok = ok and (state.gV_slice[:hi - lo].cpu() - 2 * o_gV[lo:hi]).abs().max().item() <= 1e-5 * 2 * float(o_gV.abs().max())
    ok = ok and (state.gU[a:b].cpu() - 2 * o_gU[a:b]).abs().max().item() <= 1e-5 * 2 * float(o_gU.abs().max())
    # and the dense all-reduce route gives the same item gradient
    l2, a2, gU2, gV2 = parallel.bpr_step_once(Ud, Vd, us[rank].to(dev), ps[rank].to(dev), ns[rank].to(dev), world * B, item_grad="dense")
    ok = ok and (gV2.cpu() - o_gV).abs().max().item() <= 1e-5 * float(o_gV.abs().max())
    state.close()
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, bool(ok)))


def test_bpr_owner_sharded_step_over_peer_memory():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one node (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_owner_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    res = sorted(out.get(timeout=5) for _ in range(2))
    assert res == [(0, True), (1, True)], res


# ------------------------------------------------------------------------------------------------
# parity at the BENCHED shapes (BASELINE.json configs[1], [2], [3]): sampled oracle rows + size-independent properties
# ------------------------------------------------------------------------------------------------


def _oracle_rows(U, V, rowptr, items, sample, k):
    sub_rowptr = torch.zeros(sample.numel() + 1, dtype=torch.int64)
    segs = []
    for j, u in enumerate(sample.tolist()):
        seg = items[rowptr[u]:rowptr[u + 1]]
        segs.append(seg)
        sub_rowptr[j + 1] = sub_rowptr[j] + seg.numel()
    return O.recommendations_dense(U[sample], V, sub_rowptr, torch.cat(segs) if segs else torch.zeros(0, dtype=torch.int32), k)


def test_cfg2_bench_inputs_2000_oracle_rows(native, dev):
    """The bench's own cfg2 inputs (138 493 x 26 744, d = 128, k = 100, ~20M seen pairs): 2 000+ sampled rows against the
    oracle, the fused metrics against the separate kernels, no masked item in any list."""
    nU, nI, d, k, nnz = 138493, 26744, 128, 100, 20000263
    gen = torch.Generator().manual_seed(0)
    U, V = O.make_embeddings(nU, d, "gauss", gen), O.make_embeddings(nI, d, "gauss", gen)
    seen = O.make_seen(nU, nI, nnz, gen)
    rowptr, items = O.seen_csr(seen, nU)
    Ud, Vd, rp, it = U.to(dev), V.to(dev), rowptr.to(dev), items.to(dev)
    s, i, ild, counts = native.score_topk_metrics(Ud, Vd, k, rp, it)
    _s, _i, stats = native.score_topk(Ud, Vd, k, rp, it, return_stats=True)
    assert torch.equal(_i, i) and torch.equal(_s, s)
    assert stats["rescued_rows"] < 64, stats   # the whole-chip rescue path covers the usual handful
    sample = torch.arange(0, nU, 67)   # 2 068 rows
    o_i, o_s = _oracle_rows(U, V, rowptr, items, sample, k)
    n_diff = check_topk_against_oracle(i[sample.to(dev)], s[sample.to(dev)], o_i, o_s, U, V, False)
    assert n_diff <= sample.numel() // 100
    keys = (torch.arange(nU, device=dev)[:, None] * nI + i).flatten()
    seen_keys = (seen[:, 0] * nI + seen[:, 1]).to(dev)
    assert not bool(torch.isin(keys, seen_keys).any())
    assert torch.equal(counts.long(), torch.bincount(i.flatten(), minlength=nI))
    torch.testing.assert_close(ild, native.ild_cosine(Vd, i), rtol=REL, atol=1e-7)
    torch.testing.assert_close(ild[sample.to(dev)].cpu(), O.ild_cosine(i[sample.to(dev)].cpu(), V), rtol=REL, atol=1e-7)


def test_cfg4_shard_shape_bf16(native, dev):
    """One shard of BASELINE.json configs[3]: 1.25M items x d = 256 bf16, 65 536 queries, k = 100 -- the benched shape.
    512+ sampled rows against the oracle (exact fp32 dot products of the bf16 tables) and two-way shard merge == unsharded."""
    nI, d, k, nQ = 1_250_000, 256, 100, 65536
    gen = torch.Generator().manual_seed(4)
    V = torch.randn((nI, d), generator=gen).to(torch.bfloat16)
    U = torch.randn((nQ, d), generator=gen).to(torch.bfloat16)
    Ud, Vd = U.to(dev), V.to(dev)
    s, i, stats = native.score_topk(Ud, Vd, k, return_stats=True)
    assert stats["rescued_rows"] <= 8, stats
    ok = (s[:, :-1] > s[:, 1:]) | ((s[:, :-1] == s[:, 1:]) & (i[:, :-1] < i[:, 1:]))
    assert bool(ok.all()) and bool((i >= 0).all()) and bool((i < nI).all())
    sample = torch.arange(0, nQ, 127)   # 517 rows
    o_i, o_s = O.recommendations_dense(U[sample].float(), V.float(), None, None, k)
    check_topk_against_oracle(i[sample.to(dev)], s[sample.to(dev)], o_i, o_s, U.float(), V.float(), False)
    # item-sharded in two halves + packed merge == unsharded, bit for bit
    h = nI // 2
    parts = [native.score_topk(Ud, Vd[a:b], k, item_id_offset=a) for a, b in ((0, h), (h, nI))]
    ms, mi = native.topk_merge_packed(torch.stack([native.topk_pack(ps, pi) for ps, pi in parts]))
    assert torch.equal(mi, i) and torch.equal(ms, s)


def test_cfg3_shape_sampled_gradient_rows(native, dev):
    """BASELINE.json configs[2] at its shape: 1M users x 1M items, d = 128, 65 536 triples: the dense gradients of the
    one-kernel step on every touched row against the oracle's formula (fp64), untouched rows exactly zero."""
    nU = nI = 1_000_000
    d, B = 128, 65536
    g = torch.Generator(device=dev).manual_seed(5)
    U = torch.randn((nU, d), generator=g, device=dev)
    V = torch.randn((nI, d), generator=g, device=dev)
    gh = torch.Generator().manual_seed(100)
    u, p, n = (torch.randint(0, m, (B,), generator=gh).to(dev) for m in (nU, nI, nI))
    u[:64] = u[0]; p[:32] = p[0]      # duplicates accumulate
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    pos, neg, loss_sum, auc_sum = native.bpr_step(U, V, u, p, n, 1.0 / B, gU, gV)
    Uu, Vp, Vn = U[u].double(), V[p].double(), V[n].double()
    x = ((Uu * Vp).sum(1) - (Uu * Vn).sum(1))
    c = (-torch.sigmoid(-x) / B)[:, None]
    ref_loss = torch.nn.functional.softplus(-x).mean()
    assert abs(float(loss_sum) / B - float(ref_loss)) <= 1e-5 * float(ref_loss)
    assert int(auc_sum) == int(((Uu * Vp).sum(1).float() >= (Uu * Vn).sum(1).float()).sum()) or abs(int(auc_sum) - int((x >= 0).sum())) <= 2
    ref_gU = torch.zeros((nU, d), dtype=torch.float64, device=dev).index_add_(0, u, c * (Vp - Vn))
    ref_gV = torch.zeros((nI, d), dtype=torch.float64, device=dev).index_add_(0, p, c * Uu).index_add_(0, n, -c * Uu)
    assert (gU.double() - ref_gU).abs().max().item() <= 1e-5 * float(ref_gU.abs().max())
    assert (gV.double() - ref_gV).abs().max().item() <= 1e-5 * float(ref_gV.abs().max())
    touched = torch.zeros(nI, dtype=torch.bool, device=dev); touched[p] = True; touched[n] = True
    assert float(gV[~touched].abs().sum()) == 0.0


# ------------------------------------------------------------------------------------------------
# seen-mask CSR builder (divrec_build_csr) against the oracle's sort-based restatement
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("nU,nI,nnz,shape", [
    (943, 1682, 90570, "uniform"),          # cfg1's mask
    (300, 26744, 60000, "duplicates"),      # repeated pairs, users / items out of range, empty rows
    (50, 400000, 30000, "uniform"),         # catalogue wider than one id window of the bitmap path
    (40, 20000, 200000, "long_rows"),       # rows of ~10 000 items: beyond the shared-memory sort
    (1, 100, 0, "uniform"),                 # no interactions at all
])
def test_csr_builder_vs_oracle(native, dev, nU, nI, nnz, shape):
    gen = torch.Generator().manual_seed(nU + nnz)
    if nnz == 0:
        pairs = torch.zeros((0, 2), dtype=torch.int64)
    elif shape == "long_rows":
        pairs = torch.stack([torch.randint(0, nU // 2, (nnz,), generator=gen) * 2, torch.randint(0, nI, (nnz,), generator=gen)], 1)
    else:
        pairs = torch.stack([torch.randint(0, nU, (nnz,), generator=gen), torch.randint(0, nI, (nnz,), generator=gen)], 1)
    if shape == "duplicates":
        pairs[: nnz // 3] = pairs[nnz // 3: 2 * (nnz // 3)]                 # exact repeats
        pairs[-50:, 0] = nU + 5                                             # users beyond the scored set
        pairs[-100:-50, 0] = -1
        pairs[pairs[:, 0] % 7 == 3, 0] = 0                                  # some users end up with no row at all
    keep = (pairs[:, 0] >= 0) & (pairs[:, 0] < nU)
    o_rp, o_it = O.seen_csr(pairs[keep], nU)
    rp, it = native.build_csr(pairs.to(dev), nU, nI)
    assert torch.equal(rp.cpu(), o_rp) and torch.equal(it.cpu(), o_it)
    # and through the dataset shim (the path get_model_recommendations takes)
    from divrec.datasets.base_datasets import interactions_to_csr
    rp2, it2 = interactions_to_csr(pairs.to(dev), nU, nI)
    assert torch.equal(rp2.cpu(), o_rp) and torch.equal(it2.cpu(), o_it)
