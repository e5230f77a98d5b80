"""Multi-GPU host logic: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

The path shards two ways:
  * users sharded, item table replicated -- no data-path collective; the metric reductions are one all_reduce
    of the [I] coverage histogram and one of the (ILD sum, user count) pair;
  * item table sharded (large catalogues) -- every rank scores all query users against its slice, one
    all_gather of the per-shard [U,k] (score, id) lists, then the n-way merge kernel;
  * BPR -- triples partitioned by user shard (user gradients stay local), item gradient all_reduce.

`kernels` is the compute backend (default: the ctypes binding of libdivrec_b200.so).  It is a parameter only
so that the sharding / collective logic can be exercised on CPU with the gloo backend in tests; the product
path never substitutes it.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def _kernels(kernels):
    if kernels is not None:
        return kernels
    from . import _native
    return _native


def _world(group) -> Tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced ranges: the first n % world shards get one extra element."""
    base, extra = divmod(n, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def recommend_user_sharded(U: torch.Tensor, V: torch.Tensor, seen_rowptr: Optional[torch.Tensor],
                           seen_items: Optional[torch.Tensor], k: int, group=None, kernels=None):
    """Score this rank's contiguous range of users against the replicated item table.
    Returns (items [n_local,k], scores [n_local,k], ild_local [n_local], ild_mean_global, coverage_counts_global, (a,b))."""
    K = _kernels(kernels)
    rank, world = _world(group)
    nU = U.shape[0]
    a, b = shard_range(nU, rank, world)
    rp = None
    if seen_rowptr is not None:
        rp = seen_rowptr[a:b + 1].contiguous()  # absolute offsets into the (replicated) seen_items array
    scores, items, ild, counts = K.score_topk_metrics(U[a:b].contiguous(), V, k, rp, seen_items)
    stat = torch.stack([ild.sum(dtype=torch.float64), torch.tensor(float(b - a), dtype=torch.float64, device=ild.device)])
    if world > 1:
        dist.all_reduce(counts, group=group)
        dist.all_reduce(stat, group=group)
    return items, scores, ild, (stat[0] / stat[1]).to(torch.float32), counts, (a, b)


_route_cache = {}


def _all_gather_stacked(t: torch.Tensor, world: int, group) -> torch.Tensor:
    """all_gather of equally shaped [n, k] tensors straight into one [world, n, k] buffer (no per-rank copies)."""
    flat = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(flat, t.contiguous(), group=group)
    return flat.view((world,) + tuple(t.shape))


def recommend_item_sharded(U: torch.Tensor, V_shard: torch.Tensor, shard_offset: int, k: int,
                           seen_rowptr: Optional[torch.Tensor] = None, seen_items: Optional[torch.Tensor] = None,
                           group=None, kernels=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank holds items [shard_offset, shard_offset + len(V_shard)) and all query users.
    Local top-k with global ids -> all_gather -> merge (ties -> lowest id; deterministic because the merge
    uses the total order (score desc, id asc)).  Returns the global (scores [U,k], items [U,k]) on every rank."""
    K = _kernels(kernels)
    rank, world = _world(group)
    s, i = K.score_topk(U, V_shard, k, seen_rowptr, seen_items, item_id_offset=shard_offset)
    if world == 1:
        return s, i
    # the exchange step.  Packed form (one 64-bit composite per pair, one all-gather, integer merge) when the
    # backend has it and the sizes allow; otherwise the two-tensor form.
    packed = (hasattr(K, "topk_pack") and world * k <= getattr(K, "MERGE_PACKED_MAX", 0)
              and shard_offset + V_shard.shape[0] < 2 ** 32)
    key = (id(group), world, k, int(shard_offset), int(V_shard.shape[0]), bool(packed))
    route = _route_cache.get(key)
    if route is None:  # every rank must take the same route: agreed once per configuration (a host sync), then cached
        flag = torch.tensor([1 if packed else 0], dtype=torch.int32, device=s.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        route = _route_cache[key] = int(flag)
    if route == 1:
        mine = K.topk_pack(s, i)
        return K.topk_merge_packed(_all_gather_stacked(mine, world, group))
    return K.topk_merge(_all_gather_stacked(s, world, group), _all_gather_stacked(i, world, group))


def bpr_step_sharded(U: torch.Tensor, V: torch.Tensor, u: torch.Tensor, p: torch.Tensor, n: torch.Tensor,
                     global_batch: int, group=None, kernels=None, item_grad: str = "dense"):
    """One BPR forward/backward on this rank's triples (its users only).  Gradients are those of the GLOBAL
    mean loss: every rank scales by 1 / global_batch, the item-table gradient is summed over ranks, the user
    gradient needs no communication because a user's triples live on one rank.

    item_grad: "dense" (default) all-reduces the [nI, d] item gradient (BASELINE.json config 3 as stated).
    "sparse" all-gathers the per-triple rows g_b U[u_b] with their (positive, negative) ids and every rank
    scatter-adds all of them -- the same dense result on every rank for world * B * (d + 4) * 4 gathered bytes instead
    of the ring all-reduce's 2 nI d 4; the caller guarantees EQUAL batch sizes on all ranks (no check, no host sync).
    "auto" agrees on the batch sizes first (one tiny all-reduce and a host sync per call) and takes the sparse route
    when they are equal and the gathered rows are at most half the all-reduce volume.
    Returns (loss_global, auc_global, gU_local_rows_dense, gV_global)."""
    K = _kernels(kernels)
    rank, world = _world(group)
    B, d = u.numel(), U.shape[1]
    sparse = False
    if world > 1 and item_grad != "dense" and hasattr(K, "bpr_apply_rows"):
        if item_grad == "sparse":
            sparse = True
        else:
            mm = torch.tensor([B, -B], dtype=torch.int64, device=U.device)
            dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
            bmax, bmin = int(mm[0]), -int(mm[1])
            sparse = bmax == bmin and world * B * (d + 4) * 4 <= V.shape[0] * d * 4
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    if sparse:
        rows = torch.empty((B, d), dtype=torch.float32, device=U.device)
        _pos, _neg, loss_sum, auc_sum = K.bpr_step(U, V, u, p, n, 1.0 / global_batch, gU, None, want_scores=False, item_rows=rows)
        ids = torch.stack([p.to(torch.int64), n.to(torch.int64)])                       # [2, B]
        all_rows = torch.empty((world * B, d), dtype=torch.float32, device=U.device)
        all_ids = torch.empty((world * 2, B), dtype=torch.int64, device=U.device)
        dist.all_gather_into_tensor(all_rows, rows, group=group)
        dist.all_gather_into_tensor(all_ids, ids.contiguous(), group=group)
        all_ids = all_ids.view(world, 2, B)
        K.bpr_apply_rows(all_rows, all_ids[:, 0].reshape(-1), all_ids[:, 1].reshape(-1), gV)
    elif hasattr(K, "bpr_step"):  # forward + backward in one pass over the rows
        _pos, _neg, loss_sum, auc_sum = K.bpr_step(U, V, u, p, n, 1.0 / global_batch, gU, gV, want_scores=False)
    else:
        pos, neg, loss_sum, auc_sum = K.bpr_fwd(U, V, u, p, n)
        K.bpr_bwd(U, V, u, p, n, pos, neg, 1.0 / global_batch, gU, gV)
    stat = torch.stack([loss_sum.reshape(()).to(torch.float64), auc_sum.reshape(()).to(torch.float64)])
    if world > 1:
        if not sparse:
            dist.all_reduce(gV, group=group)
        dist.all_reduce(stat, group=group)
    return (stat[0] / global_batch).to(torch.float32), (stat[1] / global_batch).to(torch.float32), gU, gV
