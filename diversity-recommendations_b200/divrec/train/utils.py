"""Score and train loops with the reference's signatures (divrec/train/utils.py:10-216).

The bodies of the two hot loops change shape:
  * get_model_recommendations: the reference iterates RankingDataset in Python (one model call and one
    full argsort per user, :59-77).  With a MatrixFactorization model the whole thing is ONE call of the
    fused scorer (tcgen05 GEMM + seen mask + per-user top-k, divrec_score_topk); other RankingModels are
    scored candidate list by candidate list and ranked on the device.
  * pair_wise_train_loop: with MatrixFactorization + LogSigmoidDifferenceLoss each batch is ONE fused
    gather + loss + AUC kernel and ONE scatter-add backward kernel (bpr_step); other combinations take
    the generic forward/backward route through the same native ops.
"""
from typing import List, Optional, Tuple

import torch

from .. import _native
from ..datasets import PairWiseDataset, PointWiseDataset, RankingDataset
from ..losses import LogSigmoidDifferenceLoss, PairWiseLoss, PointWiseLoss, RecommendationsAwareLoss
from ..metrics.auc_score import AUCScore
from ..models import MatrixFactorization, RankingModel


def _model_device(model: torch.nn.Module) -> torch.device:
    for p in model.parameters():
        return p.device
    return torch.device("cuda", torch.cuda.current_device())


def _dataset_users(dataset, model) -> int:
    n_users = int(dataset.data.number_of_users)
    if n_users > model.no_users:
        raise IndexError(f"dataset has {n_users} users but the model's user table has {model.no_users} rows")
    return n_users


def _reduce_all(losses, chunks):
    return [loss.reduce_loss_values(torch.concatenate(chunks[i])) for i, loss in enumerate(losses)]


def point_wise_score_loop(dataset: PointWiseDataset, model: RankingModel, losses: List[PointWiseLoss], **loader_params):
    model.eval()
    dev = _model_device(model)
    chunks = [[] for _ in losses]
    for user_id, item_id, user_features, item_features, true_relevance in dataset.loader(**loader_params):
        predicted = model(user_id.to(dev), item_id.to(dev), user_features, item_features)
        for i, loss in enumerate(losses):
            chunks[i].append(loss.point_wise(true_relevance.to(dev), predicted))
    return _reduce_all(losses, chunks)


def pair_wise_score_loop(dataset: PairWiseDataset, model: RankingModel, losses: List[PairWiseLoss], **loader_params):
    model.eval()
    dev = _model_device(model)
    chunks = [[] for _ in losses]
    for user_id, pos, neg, user_features, pos_features, neg_features in dataset.loader(**loader_params):
        user_id = user_id.to(dev)
        positives = model(user_id, pos.to(dev), user_features, pos_features)
        negatives = model(user_id, neg.to(dev), user_features, neg_features)
        for i, loss in enumerate(losses):
            chunks[i].append(loss.pair_wise(positives, negatives))
    return _reduce_all(losses, chunks)


@torch.no_grad()
def get_model_recommendations(dataset: RankingDataset, model: RankingModel, number_of_recommendations: int) -> torch.LongTensor:
    """[U,k] int64 item ids by (score desc, item id asc), on the model's device."""
    dev = _model_device(model)
    if dev.type != "cuda":
        raise TypeError("divrec_b200 computes on a B200 only: move the model with .to('cuda')")
    k = int(number_of_recommendations)
    if isinstance(model, MatrixFactorization):
        # one row per DATASET user, as the reference's loop over RankingDataset yields (the model's table may be larger)
        n_users = _dataset_users(dataset, model)
        rowptr, items = dataset.seen_csr(dev)
        recs, _ = model.recommend(k, rowptr, items if items.numel() else None, user_range=(0, n_users))
        return recs
    # generic RankingModel: score each user's candidates with the model, rank on the device
    out = []
    for user_rep, _pos, candidates, user_features, item_features in dataset:
        scores = model(user_rep.to(dev), candidates.to(dev), user_features, item_features).float()
        order = torch.argsort(scores, descending=True, stable=True)[:k]
        out.append(candidates.to(dev)[order])
    return torch.stack(out)


def recommendations_score_loop(dataset: RankingDataset, model: RankingModel, losses: List[RecommendationsAwareLoss],
                               number_of_recommendations: int):
    """[loss(interactions, recommendations) for loss in losses] as in the reference (train/utils.py:80-91).
    With a MatrixFactorization model, a cosine IntraListDiversityScore over the model's own item table and an
    EntropyDiversityScore are not computed from the [U,k] lists afterwards: their per-user ILD and coverage
    histogram come out of the fused selection kernel (divrec_score_topk_metrics) in the same pass."""
    import math
    from ..losses import IntraListDiversityScore
    from ..metrics import EntropyDiversityScore
    model.eval()
    interactions = dataset.data.interactions
    k = int(number_of_recommendations)
    dev = _model_device(model)
    fused = {}
    if isinstance(model, MatrixFactorization) and dev.type == "cuda":
        V = model.item_embeddings.weight
        for i, loss in enumerate(losses):
            if (type(loss) is IntraListDiversityScore and loss.item_embeddings is not None and loss.ild_mode == "sumvec"
                    and loss.item_embeddings.data_ptr() == V.data_ptr() and k >= 2 and V.shape[1] % 4 == 0 and V.shape[1] <= 256):
                fused[i] = "ild"
            elif type(loss) is EntropyDiversityScore and int(loss.dataset.number_of_items) == V.shape[0]:
                fused[i] = "entropy"
    if not fused:
        recommendations = get_model_recommendations(dataset, model, k)
        return [loss(interactions, recommendations) for loss in losses]
    with torch.no_grad():
        rowptr, items = dataset.seen_csr(dev)
        n_users = _dataset_users(dataset, model)
        _scores, recommendations, ild, counts = _native.score_topk_metrics(
            model.user_embeddings.weight[:n_users], model.item_embeddings.weight, k, rowptr, items if items.numel() else None,
            want_ild="ild" in fused.values(), want_coverage="entropy" in fused.values())
    out = []
    for i, loss in enumerate(losses):
        if fused.get(i) == "ild":
            out.append(loss.reduce_loss_values(ild))
        elif fused.get(i) == "entropy":
            entropy = _native.entropy_from_counts(counts)[0]
            out.append((entropy - math.log(k)) / (loss.max_entropy - math.log(k)))
        else:
            out.append(loss(interactions, recommendations))
    return out


def point_wise_train_loop(dataset: PointWiseDataset, model: RankingModel, loss: PointWiseLoss,
                          optimizer: torch.optim.Optimizer, scores: Optional[List[PointWiseLoss]] = None,
                          **loader_params) -> Tuple[float, List[float]]:
    assert scores is None or all(score.reduce for score in scores)
    model.train()
    dev = _model_device(model)
    scores = scores or []
    n, loss_total, score_totals = 0, 0.0, [0.0] * len(scores)
    for user_id, item_id, user_features, item_features, true_relevance in dataset.loader(**loader_params):
        true_relevance = true_relevance.to(dev)
        predicted = model(user_id.to(dev), item_id.to(dev), user_features, item_features)
        value = loss(true_relevance, predicted)
        value.backward()
        optimizer.step()
        optimizer.zero_grad()
        n += 1
        loss_total += value.item()
        for i, score in enumerate(scores):
            score_totals[i] += score(true_relevance, predicted).item()
    return loss_total / n, [t / n for t in score_totals]


class _FusedBpr(torch.autograd.Function):
    """mean(-logsigmoid(<U[u],V[p]> - <U[u],V[n]>)) with the gather fused into the loss kernel and the
    scatter-add fused into the backward kernel (dense .grad, duplicates accumulate)."""

    @staticmethod
    def forward(ctx, U, V, u, p, n):
        pos, neg, loss_sum, auc_sum = _native.bpr_fwd(U, V, u, p, n)
        ctx.save_for_backward(U, V, u, p, n, pos, neg)
        ctx.mark_non_differentiable(pos, neg, auc_sum)
        return loss_sum[0] / u.numel(), pos, neg, auc_sum

    @staticmethod
    def backward(ctx, g_loss, _gp, _gn, _ga):
        U, V, u, p, n, pos, neg = ctx.saved_tensors
        gU, gV = torch.zeros_like(U), torch.zeros_like(V)
        _native.bpr_bwd(U, V, u, p, n, pos, neg, float(g_loss) / u.numel(), gU, gV)
        return gU, gV, None, None, None


def bpr_step(model: MatrixFactorization, user_id, positive_item_id, negative_item_id):
    """(loss 0-dim with grad, pos [B], neg [B], auc_sum int32 [1]) for one batch of triples."""
    U, V = model.user_embeddings.weight, model.item_embeddings.weight
    dev = U.device
    to = dict(device=dev, dtype=torch.int64)
    return _FusedBpr.apply(U, V, user_id.to(**to), positive_item_id.to(**to), negative_item_id.to(**to))


def bpr_train_step(model: MatrixFactorization, user_id, positive_item_id, negative_item_id):
    """`loss = mean(-logsigmoid(pos - neg)); loss.backward()` (divrec/train/utils.py:145-149) as ONE kernel: the
    upstream gradient of the loop's backward() is 1, so the rows of a triple are read once and the gradients are
    accumulated straight into the dense `.grad` tables (created zero-filled when absent, as autograd would).
    Returns (loss 0-dim, detached; pos [B]; neg [B]; auc_sum int32 [1])."""
    U, V = model.user_embeddings.weight, model.item_embeddings.weight
    to = dict(device=U.device, dtype=torch.int64)
    u, p, n = user_id.to(**to), positive_item_id.to(**to), negative_item_id.to(**to)
    grads = []
    for w in (U, V):
        if w.requires_grad and w.grad is None:
            w.grad = torch.zeros_like(w)
        grads.append(w.grad if w.requires_grad else None)
    pos, neg, loss_sum, auc_sum = _native.bpr_step(U.detach(), V.detach(), u, p, n, 1.0 / u.numel(), grads[0], grads[1])
    return loss_sum[0] / u.numel(), pos, neg, auc_sum


def pair_wise_train_loop(dataset: PairWiseDataset, model: RankingModel, loss: PairWiseLoss,
                         optimizer: torch.optim.Optimizer, scores: Optional[List[PairWiseLoss]] = None,
                         **loader_params) -> Tuple[float, List[float]]:
    assert scores is None or all(score.reduce for score in scores)
    model.train()
    dev = _model_device(model)
    scores = scores or []
    fused = isinstance(model, MatrixFactorization) and type(loss) is LogSigmoidDifferenceLoss
    n, loss_total, score_totals = 0, 0.0, [0.0] * len(scores)
    # Datasets without features (on which the reference's own loop raises in default_collate) with a MatrixFactorization
    # model take the device path end to end: triples sampled on the GPU (divrec_bpr_sample_triples), cut into batches in
    # the reference's iteration order.  Anything else walks the Python sampler through a DataLoader as the reference does.
    device_sampler = (fused and dev.type == "cuda" and not dataset.data.has_user_features() and not dataset.data.has_item_features()
                      and 0 < dataset.max_sampled <= 256 and not loader_params.get("shuffle", False))
    if device_sampler:
        tu, tp, tn = dataset.sample_triples(dev, seed=int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        bs = int(loader_params.get("batch_size", 1) or 1)
        batches = ((tu[a:a + bs], tp[a:a + bs], tn[a:a + bs], None, None, None) for a in range(0, tu.numel(), bs))
        if loader_params.get("drop_last", False):
            batches = (b for b in batches if b[0].numel() == bs)
    else:
        batches = dataset.loader(**loader_params)
    for user_id, pos, neg, user_features, pos_features, neg_features in batches:
        if fused:  # forward + backward in one pass over the rows
            value, positives, negatives, auc_sum = bpr_train_step(model, user_id, pos, neg)
        else:
            user_id = user_id.to(dev)
            positives = model(user_id, pos.to(dev), user_features, pos_features)
            negatives = model(user_id, neg.to(dev), user_features, neg_features)
            value = loss(positives, negatives)
            value.backward()
        optimizer.step()
        optimizer.zero_grad()
        n += 1
        loss_total += value.item()
        if dev.type == "cuda":  # the loop's synchronisation point: ids outside the tables raise here (nn.Embedding's IndexError)
            _native.raise_on_bad_indices(dev)
        for i, score in enumerate(scores):
            if fused and type(score) is AUCScore and score.reduction == "mean":
                score_totals[i] += auc_sum.item() / positives.numel()  # free epilogue of the fused forward
            else:
                score_totals[i] += score(positives, negatives).item()
    return loss_total / n, [t / n for t in score_totals]


def recommendations_train_loop(dataset: RankingDataset, model: RankingModel, loss: RecommendationsAwareLoss,
                               number_of_recommendations: int, optimizer: torch.optim.Optimizer,
                               scores: Optional[List[RecommendationsAwareLoss]] = None) -> Tuple[float, List[float]]:
    """Kept for import compatibility.  The reference's version (train/utils.py:167-216) cannot run: the
    recommendations it hands to the loss are integer ids with no path back to the model parameters, so
    `loss_value.backward()` fails with every shipped loss."""
    raise NotImplementedError(
        "recommendations_train_loop is dead code in the reference (no differentiable path from the "
        "recommended item ids to the model); use pair_wise_train_loop")


_host_pipe = {}


def host_recommend_and_score(U_host: torch.Tensor, V_host: torch.Tensor, seen_rowptr_host: torch.Tensor,
                             seen_items_host: torch.Tensor, k: int, device, want_recommendations: bool = True,
                             chunks: int = 4, index_dtype: torch.dtype = torch.int64):
    """The whole evaluation pass from HOST buffers: copy the embedding tables and the seen-item CSR to the
    device, run score + mask + top-k with the cosine ILD and the coverage histogram fused, and bring the
    results back to pinned host memory.  Equivalent of recommendations_score_loop(RankingDataset(test,
    frozen=train), MatrixFactorization, [IntraListDiversityScore, EntropyDiversityScore], k) for a caller
    whose model and interactions live in host memory.

    Users are processed in `chunks` contiguous ranges on three streams -- host->device copies of the next
    range, kernels of the current one, device->host copies of the previous one -- so the PCIe transfers hide
    behind the kernels.  Pass pinned tensors (pageable memory makes the copies synchronous).
    index_dtype=torch.int32 brings the [U,k] recommendations back as 32-bit ids (narrowed on the device): the int64 table
    the reference API returns is 40 % of the pass's PCIe traffic, and the link is what bounds this call.
    Returns a dict with host tensors and the bytes moved."""
    import math
    dev = torch.device(device)
    nU, d = U_host.shape
    nI = V_host.shape[0]
    if index_dtype not in (torch.int64, torch.int32):
        raise ValueError("index_dtype must be torch.int64 or torch.int32")
    if index_dtype == torch.int32 and nI >= 2 ** 31:
        raise ValueError("index_dtype=torch.int32 needs fewer than 2^31 items")
    key = (dev.index, nU, nI, d, k, seen_items_host.numel(), chunks, index_dtype)
    st = _host_pipe.get(key)
    if st is None:
        wave = 148 * 128  # users scored by one wave of the persistent filter kernel
        per = nU if nU <= wave else -(-(-(-nU // max(1, chunks))) // wave) * wave  # ceil(ceil(nU / chunks) / wave) * wave
        bounds = list(range(0, nU, per)) + [nU]
        st = {
            "U": torch.empty_like(U_host, device=dev), "V": torch.empty_like(V_host, device=dev),
            "rp": torch.empty_like(seen_rowptr_host, device=dev), "it": torch.empty_like(seen_items_host, device=dev),
            "counts": torch.zeros(nI, dtype=torch.int32, device=dev),
            "recs": torch.empty((nU, k), dtype=torch.int64, device=dev),
            "scores": torch.empty((nU, k), dtype=torch.float32, device=dev),
            "ild": torch.empty(nU, dtype=torch.float32, device=dev),
            "recs_h": torch.empty((nU, k), dtype=index_dtype).pin_memory(),
            "recs_n": torch.empty((nU, k), dtype=index_dtype, device=dev) if index_dtype != torch.int64 else None,
            "ild_h": torch.empty(nU, dtype=torch.float32).pin_memory(),
            "counts_h": torch.empty(nI, dtype=torch.int32).pin_memory(),
            "ent_h": torch.empty(2, dtype=torch.float32).pin_memory(),
            "bounds": bounds,
            "s_in": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
        }
        _host_pipe.clear()
        _host_pipe[key] = st
    bounds = st["bounds"]
    # seen-item ranges of the user ranges, from THIS call's row pointers (the cached state is keyed on shapes only)
    rp_at = [int(seen_rowptr_host[b]) for b in bounds]
    main = torch.cuda.current_stream(dev)
    s_in, s_out = st["s_in"], st["s_out"]
    s_in.wait_stream(main)
    s_out.wait_stream(main)
    h2d = d2h = 0
    # ---- host -> device: the item table and the row pointers first, then user rows + seen items range by range ----
    ready = []
    with torch.cuda.stream(s_in):
        st["V"].copy_(V_host, non_blocking=True)
        st["rp"].copy_(seen_rowptr_host, non_blocking=True)
        h2d += V_host.numel() * V_host.element_size() + seen_rowptr_host.numel() * 8
        for c in range(len(bounds) - 1):
            a, b = bounds[c], bounds[c + 1]
            st["U"][a:b].copy_(U_host[a:b], non_blocking=True)
            h2d += (b - a) * d * U_host.element_size()
            if seen_items_host.numel():
                lo, hi = rp_at[c], rp_at[c + 1]
                if hi > lo:
                    st["it"][lo:hi].copy_(seen_items_host[lo:hi], non_blocking=True)
                    h2d += (hi - lo) * 4
            ev = torch.cuda.Event()
            ev.record(s_in)
            ready.append(ev)
    # ---- kernels on the caller's stream, device -> host behind them ----
    st["counts"].zero_()
    items_dev = st["it"] if st["it"].numel() else None
    for c in range(len(bounds) - 1):
        a, b = bounds[c], bounds[c + 1]
        main.wait_event(ready[c])
        _native.score_topk_metrics(st["U"][a:b], st["V"], k, st["rp"][a:b + 1], items_dev, counts=st["counts"],
                                   out=(st["scores"][a:b], st["recs"][a:b], st["ild"][a:b]))
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            if want_recommendations:
                src = st["recs"]
                if st["recs_n"] is not None:  # narrow on the device (on the copy stream, behind the kernels)
                    st["recs_n"][a:b].copy_(st["recs"][a:b])
                    src = st["recs_n"]
                st["recs_h"][a:b].copy_(src[a:b], non_blocking=True)
                d2h += (b - a) * k * src.element_size()
            st["ild_h"][a:b].copy_(st["ild"][a:b], non_blocking=True)
            d2h += (b - a) * 4
    ent = _native.entropy_from_counts(st["counts"])
    st["counts_h"].copy_(st["counts"], non_blocking=True)
    st["ent_h"].copy_(ent, non_blocking=True)
    d2h += nI * 4 + 8
    main.wait_stream(s_out)
    main.synchronize()
    entropy = (float(st["ent_h"][0]) - math.log(k)) / (math.log(nI) - math.log(k))
    return {"recommendations": st["recs_h"] if want_recommendations else None, "ild_per_user": st["ild_h"],
            "coverage_counts": st["counts_h"], "entropy_diversity": entropy, "h2d_bytes": h2d, "d2h_bytes": d2h}
