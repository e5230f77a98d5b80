"""CPU oracle for the divrec hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in numpy / CPU torch, of what amtsyplov/diversity-recommendations ("divrec")
computes on the path  score -> seen-item mask -> top-k -> ILD / coverage  and the BPR
gather / scatter-add step.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module; nothing under diversity-recommendations_b200/ does.

Parity pinning.  The reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference itself*:
tests/golden/make_golden.py imports the real package from /root/reference, runs
get_model_recommendations / IntraListDiversityScore / EntropyDiversityScore /
LogSigmoidDifferenceLoss + autograd on seeded inputs and commits inputs + outputs as
tests/golden/*.npz; tests/test_oracle_golden.py checks every function below against them.
The arithmetic itself lives in torch (reference pin torch==2.1.0, divrec/requirements.txt:1;
2.11.0 here); the call sites are cited per function.  Two items of the north-star path have no
reference implementation at all (cosine distance for ILD, greedy MMR re-rank): their semantics
are defined here and in DESIGN.md and are marked "parity unpinned by reference" -- the
reference-anchored check for the cosine ILD is IntraListDiversityScore fed with D = 1 - En En^T.

All file:line citations are relative to the reference repository root.
"""
from __future__ import annotations

import math
from itertools import combinations
from typing import Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# scoring
# --------------------------------------------------------------------------------------------


def mf_forward(U: torch.Tensor, V: torch.Tensor, user_id: torch.Tensor, item_id: torch.Tensor) -> torch.Tensor:
    """MatrixFactorization.forward -- divrec/models/matrix_factorization.py:26-28:
    u = user_embeddings(user_id); i = item_embeddings(item_id); sum(u * i, dim=1).  Paired ids."""
    return torch.sum(U[user_id] * V[item_id], dim=1)


def seen_csr(frozen_interactions: torch.Tensor, n_users: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[nnz,2] int64 (user,item) pairs (any order, duplicates allowed) -> CSR (rowptr int64 [U+1],
    items int32 ascending per row, de-duplicated).  This is the set `frozen` that
    RankingDataset.__iter__ builds per user at divrec/datasets/base_datasets.py:143-149."""
    if frozen_interactions is None or frozen_interactions.numel() == 0:
        return torch.zeros(n_users + 1, dtype=torch.int64), torch.zeros(0, dtype=torch.int32)
    pairs = torch.unique(frozen_interactions.to(torch.int64), dim=0)  # sorted by (user, item)
    counts = torch.bincount(pairs[:, 0], minlength=n_users)
    rowptr = torch.zeros(n_users + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, pairs[:, 1].to(torch.int32)


def ranking_candidates(n_items: int, frozen_interactions: torch.Tensor, user: int) -> torch.Tensor:
    """Candidate ids of one user: list(frozenset(range(I)) - frozen_u), ascending item id --
    divrec/datasets/base_datasets.py:137,143-149,168.  (CPython iterates small-int sets in
    ascending order)  The dataset's own positives are NOT removed (:125)."""
    frozen = set(frozen_interactions[frozen_interactions[:, 0] == user, 1].tolist())
    return torch.tensor([i for i in range(n_items) if i not in frozen], dtype=torch.int64)


def recommendations_loop(U: torch.Tensor, V: torch.Tensor, frozen_interactions: torch.Tensor, k: int,
                         users: Optional[range] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_model_recommendations, literally: divrec/train/utils.py:59-77 over
    RankingDataset.__iter__ (base_datasets.py:136-171).  One user at a time: score the candidates
    with mf_forward, argsort descending, keep k.  The reference's argsort is not stable, so its
    order on exact ties is arbitrary; the oracle fixes the canonical order (score desc, item id
    asc) = argsort(stable=True) over ascending candidates, equal to the reference whenever the
    reference is tie-free.  Returns (items [U,k] int64, scores [U,k] fp32).  Small cases only."""
    n_items = V.shape[0]
    users = range(U.shape[0]) if users is None else users
    rec, sc = [], []
    for u in users:
        cand = ranking_candidates(n_items, frozen_interactions, u)
        s = mf_forward(U, V, torch.full((cand.numel(),), u, dtype=torch.int64), cand)
        order = torch.argsort(s, descending=True, stable=True)[:k]
        rec.append(cand[order])
        sc.append(s[order])
    return torch.stack(rec), torch.stack(sc)


def recommendations_dense(U: torch.Tensor, V: torch.Tensor, rowptr: Optional[torch.Tensor],
                          seen_items: Optional[torch.Tensor], k: int, chunk: int = 2048,
                          exact_scores: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """The same result as recommendations_loop for large shapes: S = U V^T chunk by chunk
    (fp64 accumulate when exact_scores, so that the order is the order of the exactly rounded
    fp32 scores rather than of one particular summation order), S[u, seen_u] = -inf, stable
    descending sort, first k.  Rows with fewer than k candidates are padded (-inf, -1).
    Restates divrec/train/utils.py:66-76 + base_datasets.py:143-151 as dense algebra."""
    nU, nI = U.shape[0], V.shape[0]
    items = torch.empty((nU, k), dtype=torch.int64)
    scores = torch.empty((nU, k), dtype=torch.float32)
    Vd = V.double() if exact_scores else V.float()
    for a in range(0, nU, chunk):
        b = min(nU, a + chunk)
        Ub = U[a:b].double() if exact_scores else U[a:b].float()
        S = (Ub @ Vd.T).float()
        if rowptr is not None and seen_items is not None and seen_items.numel() > 0:
            lo, hi = int(rowptr[a]), int(rowptr[b])
            if hi > lo:
                cnt = (rowptr[a + 1:b + 1] - rowptr[a:b])
                rows = torch.repeat_interleave(torch.arange(b - a), cnt)
                S[rows, seen_items[lo:hi].long()] = float("-inf")
        v, i = torch.sort(S, dim=1, descending=True, stable=True)
        v, i = v[:, :k], i[:, :k]
        if v.shape[1] < k:  # fewer items than k
            pad = k - v.shape[1]
            v = torch.cat([v, torch.full((b - a, pad), float("-inf"))], 1)
            i = torch.cat([i, torch.full((b - a, pad), -1, dtype=torch.int64)], 1)
        i = torch.where(torch.isinf(v) & (v < 0), torch.full_like(i, -1), i)
        items[a:b], scores[a:b] = i, v
    return items, scores


def topk_merge(gathered_scores: torch.Tensor, gathered_items: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """n-way merge of per-shard (score desc, id asc) lists [n_shards, U, k] into the global top-k.
    Restates the argsort at divrec/train/utils.py:73-75 applied to the union of shard lists."""
    n, nU, kk = gathered_scores.shape
    s = gathered_scores.permute(1, 0, 2).reshape(nU, n * kk)
    i = gathered_items.permute(1, 0, 2).reshape(nU, n * kk)
    s = torch.where(i < 0, torch.full_like(s, float("-inf")), s)
    # total order: score desc, id asc (padding last) -> sort by id asc first, then stable by score
    big = torch.where(i < 0, torch.full_like(i, torch.iinfo(torch.int64).max), i)
    o1 = torch.argsort(big, dim=1, stable=True)
    s1, i1 = torch.gather(s, 1, o1), torch.gather(i, 1, o1)
    o2 = torch.argsort(s1, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(s1, 1, o2), torch.gather(i1, 1, o2)


def topk_pack(scores: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
    """The packed exchange format of the item-sharded all-gather (include/divrec_b200.h, divrec_topk_pack):
    (order-preserving 32-bit key of the fp32 score) << 32 | ~uint32(id); padding (id < 0) -> 0.  Descending
    unsigned order of the composites is the canonical order of train/utils.py:73-75 (score desc, id asc).
    Returned as int64 (the bit pattern of the uint64 composite)."""
    bits = (scores.contiguous() + 0.0).view(torch.int32).numpy().view(np.uint32).astype(np.uint64)
    key = np.where(bits & np.uint64(0x80000000), ~bits & np.uint64(0xFFFFFFFF), bits | np.uint64(0x80000000))
    ids = items.numpy()
    inv = (~ids.astype(np.uint64)) & np.uint64(0xFFFFFFFF)
    c = np.where(ids >= 0, (key << np.uint64(32)) | inv, np.uint64(0))
    return torch.from_numpy(c.view(np.int64).copy())


def topk_merge_packed(gathered: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge [n_shards, U, k] composites: the k largest (unsigned) per user, unpacked to (scores, items)."""
    n, nU, k = gathered.shape
    c = gathered.permute(1, 0, 2).reshape(nU, n * k).numpy().view(np.uint64)
    c = np.sort(c, axis=1)[:, ::-1][:, :k]
    key = (c >> np.uint64(32)).astype(np.uint32)
    bits = np.where(key & np.uint32(0x80000000), key & np.uint32(0x7FFFFFFF), ~key)
    scores = torch.from_numpy(bits.astype(np.uint32).view(np.float32).copy())
    items = torch.from_numpy(((~c) & np.uint64(0xFFFFFFFF)).astype(np.int64))
    pad = torch.from_numpy(c == 0)
    scores = torch.where(pad, torch.full_like(scores, float("-inf")), scores)
    items = torch.where(pad, torch.full_like(items, -1), items)
    return scores, items


# --------------------------------------------------------------------------------------------
# intra-list diversity
# --------------------------------------------------------------------------------------------


def ild_matrix_loop(recs: torch.Tensor, D: torch.Tensor) -> torch.Tensor:
    """IntraListDiversityScore.recommendations_loss with reduction="none", literally:
    divrec/losses/intra_list_diversity_score.py:25-32,40-42 -- per user
    sum(D[i, j] for i, j in combinations(recs_u, 2)) as a Python sum, collected with
    torch.Tensor([...]) (fp32), divided by k (k - 1).  Small cases only."""
    k = recs.size(1)
    vals = torch.Tensor([sum(D[i, j] for i, j in combinations(r, 2)) for r in recs])
    vals /= k * (k - 1)
    return vals


def ild_matrix(recs: torch.Tensor, D: torch.Tensor) -> torch.Tensor:
    """Vectorised ild_matrix_loop: pairs a<b in rank order, D[rec_a, rec_b] (so an asymmetric D is
    honoured as D[earlier, later]); fp64 accumulation then fp32, / (k (k-1)).  ids < 0 are skipped."""
    nU, k = recs.shape
    iu = torch.triu_indices(k, k, 1)
    a, b = recs[:, iu[0]], recs[:, iu[1]]
    ok = (a >= 0) & (b >= 0)
    vals = D[a.clamp(min=0), b.clamp(min=0)].double() * ok
    return (vals.sum(1) / (k * (k - 1))).float()


def cosine_distance_matrix(V: torch.Tensor) -> torch.Tensor:
    """D = 1 - En En^T, En = V / |V|_2.  NOT in the reference (no cosine function anywhere); this
    is the distance the north-star names.  Reference anchor: IntraListDiversityScore accepts any
    [I,I] distance matrix (intra_list_diversity_score.py:16-18)."""
    Vd = V.double()
    En = Vd / Vd.norm(dim=1, keepdim=True)
    return (1.0 - En @ En.T).float()


def ild_cosine(recs: torch.Tensor, V: torch.Tensor) -> torch.Tensor:
    """Per-user cosine ILD without the [I,I] matrix: sum_{a<b} (1 - cos(V[rec_a], V[rec_b])) / (k (k-1)),
    fp64 internally.  Equals ild_matrix(recs, cosine_distance_matrix(V)) to fp32 rounding."""
    nU, k = recs.shape
    Vd = V.double()
    En = Vd / Vd.norm(dim=1, keepdim=True)
    out = torch.empty(nU, dtype=torch.float32)
    iu = torch.triu_indices(k, k, 1)
    for a in range(0, nU, 4096):
        r = recs[a:a + 4096]
        ok = (r >= 0)
        E = En[r.clamp(min=0)] * ok[..., None]
        G = E @ E.transpose(1, 2)
        pair_ok = (ok[:, iu[0]] & ok[:, iu[1]]).double()
        out[a:a + 4096] = ((pair_ok - G[:, iu[0], iu[1]] * pair_ok).sum(1) / (k * (k - 1))).float()
    return out


def reduce_loss_values(vals: torch.Tensor, reduction: str) -> torch.Tensor:
    """ScoreWithReduction.reduce_loss_values -- divrec/losses/base_losses.py:22-27."""
    if reduction == "none":
        return vals
    if reduction == "mean":
        return torch.sum(vals, dim=0) / vals.size(0)
    return torch.sum(vals, dim=0)


# --------------------------------------------------------------------------------------------
# coverage / entropy
# --------------------------------------------------------------------------------------------


def coverage_counts(recs: torch.Tensor, n_items: int) -> torch.Tensor:
    """Dense [I] int64 histogram of recommended ids; its non-zeros are exactly the (values, counts)
    of torch.unique(recommendations, return_counts=True) -- divrec/metrics/entropy_diversity_score.py:22."""
    flat = recs.reshape(-1)
    flat = flat[(flat >= 0) & (flat < n_items)]
    return torch.bincount(flat, minlength=n_items)


def entropy_diversity(recs: torch.Tensor, n_items: int) -> torch.Tensor:
    """EntropyDiversityScore.recommendations_loss -- divrec/metrics/entropy_diversity_score.py:19-26
    with max_entropy = log(number_of_items) (:12) and min_entropy = log(recommendations.size(1)) (:25)."""
    _, counts = torch.unique(recs, return_counts=True)
    p = counts / counts.sum()
    h = -torch.sum(p * torch.log(p))
    mn = math.log(recs.size(1))
    return (h - mn) / (math.log(n_items) - mn)


# --------------------------------------------------------------------------------------------
# BPR
# --------------------------------------------------------------------------------------------


def bpr_forward(U, V, u, p, n):
    """pos = model(u, p), neg = model(u, n) (divrec/train/utils.py:145-146),
    loss = mean(-logsigmoid(pos - neg)) (log_sigmoid_difference_loss.py:14 + base_losses.py:25-26;
    the ctor bug at :8 makes the reduction always "mean"), auc = mean((pos >= neg).int())
    (auc_score.py:10 + base_losses.py:25-26)."""
    pos, neg = mf_forward(U, V, u, p), mf_forward(U, V, u, n)
    lv = -torch.nn.functional.logsigmoid(pos - neg)
    loss = torch.sum(lv, dim=0) / lv.size(0)
    hit = (pos >= neg).int()
    auc = torch.sum(hit, dim=0) / hit.size(0)
    return pos, neg, loss, auc


def bpr_backward(U, V, u, p, n, upstream: float = 1.0):
    """Dense gradients of bpr_forward's loss w.r.t. both tables, written out by hand (what autograd's
    embedding_dense_backward produces behind loss.backward() at divrec/train/utils.py:149):
      x = pos - neg; c = -sigmoid(-x) * upstream / B
      gU[u] += c (V[p] - V[n]);  gV[p] += c U[u];  gV[n] -= c U[u]   (duplicates accumulate), fp64 sums."""
    B = u.numel()
    Ud, Vd = U.double(), V.double()
    x = (Ud[u] * Vd[p]).sum(1) - (Ud[u] * Vd[n]).sum(1)
    c = (-torch.sigmoid(-x) * upstream / B)[:, None]
    gU = torch.zeros_like(Ud)
    gV = torch.zeros_like(Vd)
    gU.index_add_(0, u, c * (Vd[p] - Vd[n]))
    gV.index_add_(0, p, c * Ud[u])
    gV.index_add_(0, n, -c * Ud[u])
    return gU.float(), gV.float()


# --------------------------------------------------------------------------------------------
# greedy MMR re-rank (north-star extension; NOT in the reference -> parity unpinned by reference)
# --------------------------------------------------------------------------------------------


def mmr_rerank(cand_items: torch.Tensor, cand_scores: torch.Tensor, V: torch.Tensor, k: int, lam: float) -> torch.Tensor:
    """Greedy maximal-marginal-relevance selection of k out of C candidates per user.
    First pick = candidate 0 (best score; candidates come sorted by (score desc, id asc)); then
    repeatedly pick argmax_i  lam * score_i - (1 - lam) * max_{j in S} cos(V_i, V_j)
    (fp32 arithmetic, ties -> lowest candidate rank).  Padding candidates (id < 0) are never picked
    before real ones.  The only reference anchor is README.md:3-5 ("tools for increasing ...
    diversity"); semantics defined in SURVEY.md 8f N1."""
    nU, C = cand_items.shape
    out = torch.full((nU, k), -1, dtype=torch.int64)
    Vf = V.float()
    En = Vf / Vf.norm(dim=1, keepdim=True)
    for u in range(nU):
        ids = cand_items[u]
        valid = ids >= 0
        E = En[ids.clamp(min=0)]
        sc = cand_scores[u].float()
        maxsim = torch.full((C,), float("-inf"))
        picked = torch.zeros(C, dtype=torch.bool)
        for r in range(min(k, int(valid.sum()))):
            if r == 0:
                j = 0
            else:
                obj = lam * sc - (1.0 - lam) * maxsim
                obj = torch.where(picked | ~valid, torch.full_like(obj, float("-inf")), obj)
                j = int(torch.argmax(obj))  # first maximal index = lowest candidate rank
            picked[j] = True
            out[u, r] = ids[j]
            maxsim = torch.maximum(maxsim, E @ E[j])
    return out


def pairwise_triples(interactions: torch.Tensor, frozen: Optional[torch.Tensor], n_users: int, n_items: int,
                     max_sampled: int, rng) -> torch.Tensor:
    """PairWiseDataset.__iter__ (divrec/datasets/base_datasets.py:70-107) restated: per user, the positives are the
    user's items (a frozenset filled in interaction order, :73-77), the negatives are all items minus the positives
    minus the user's frozen items (:79-86); max_sampled of each are drawn with replacement by `rng.choices` over the
    sets' iteration order (:88-90, `rng` = the `random` module or a random.Random), and the cross product is emitted
    positive-major (:92-107).  Returns int64 [U * max_sampled^2, 3]."""
    universe = frozenset(range(n_items))
    users = interactions[:, 0].tolist()
    its = interactions[:, 1].tolist()
    f_users = frozen[:, 0].tolist() if frozen is not None else []
    f_its = frozen[:, 1].tolist() if frozen is not None else []
    out = []
    for user in range(n_users):
        liked = frozenset([i for uu, i in zip(users, its) if uu == user])
        if frozen is not None:
            blocked = frozenset([i for uu, i in zip(f_users, f_its) if uu == user])
            rest = universe - liked - blocked
        else:
            rest = universe - liked
        if max_sampled > 0:
            liked = rng.choices(list(liked), k=max_sampled)
            rest = rng.choices(list(rest), k=max_sampled)
        for a in liked:
            for b in rest:
                out.append((user, a, b))
    return torch.tensor(out, dtype=torch.int64).reshape(-1, 3)


# --------------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests, smoke and bench so oracle and kernels see the same bits
# --------------------------------------------------------------------------------------------


def make_embeddings(n: int, d: int, kind: str, gen: torch.Generator) -> torch.Tensor:
    """kind "gauss": N(0,1) fp32 (= nn.Embedding default init, matrix_factorization.py:16-17);
    kind "grid": randint(-8, 9) / 8 -- exact in bf16/fp16/fp32 with exact partial sums, so every
    GEMM mode must be bit-exact and ties are abundant (Tier-E)."""
    if kind == "grid":
        return torch.randint(-8, 9, (n, d), generator=gen).float() / 8
    return torch.randn(n, d, generator=gen)


def make_seen(n_users: int, n_items: int, nnz: int, gen: torch.Generator) -> torch.Tensor:
    """Uniform random (user, item) pairs, de-duplicated, every user >= 1 pair; [nnz', 2] int64."""
    u = torch.randint(0, n_users, (nnz,), generator=gen)
    i = torch.randint(0, n_items, (nnz,), generator=gen)
    u[:n_users] = torch.arange(n_users)  # every user has at least one interaction
    key = torch.unique(u * n_items + i)
    return torch.stack([key // n_items, key % n_items], 1)
