"""Iteration protocols over UserItemInteractionsDataset with the reference's class names
(divrec/datasets/base_datasets.py:41-171).

RankingDataset is the seen-item mask producer of the hot path.  The reference yields, per user,
the Python list `items - frozen_u` (an O(nnz) scan plus an O(I) set difference per user -- 78 % of
its evaluation time).  Here the mask is one CSR built once (`seen_csr`), sorted by
(user, item) and de-duplicated, which the fused scorer consumes directly; `__iter__` still yields
the reference's row tuples for code that walks the dataset itself.
"""
import random
from typing import Iterator, Optional, Tuple

import torch
from torch.utils.data import DataLoader, Dataset, IterableDataset

from .storages import UserItemInteractionsDataset, get_item_features, get_user_features

PointWiseRow = Tuple[int, int, Optional[torch.Tensor], Optional[torch.Tensor], float]
PairWiseRow = Tuple[int, int, int, Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]
RankingRow = Tuple[torch.LongTensor, torch.LongTensor, torch.LongTensor, Optional[torch.Tensor], Optional[torch.Tensor]]


def interactions_to_csr(interactions: Optional[torch.Tensor], n_users: int, n_items: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[nnz,2] (user,item) pairs -> (rowptr int64 [U+1], items int32 ascending and unique per row).
    Interactions on a CUDA device go through the library's sort-free builder (divrec_build_csr: count, scan, bucket,
    per-row sort + de-duplication, compaction); host tensors (the reference-compatible `__iter__` path) are built with a
    sort on the host.  Users outside [0, n_users) have no row (a frozen / train set may hold users the scored set never
    mentions: the reference filters `frozen` per user, base_datasets.py:145-147)."""
    if interactions is None or interactions.numel() == 0:
        dev = interactions.device if interactions is not None else "cpu"
        return torch.zeros(n_users + 1, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int32, device=dev)
    if interactions.is_cuda:
        from .. import _native
        if n_items is None:
            n_items = int(interactions[:, 1].max()) + 1
        return _native.build_csr(interactions.to(torch.int64), int(n_users), int(n_items))
    pairs = interactions.to(torch.int64)
    pairs = pairs[(pairs[:, 0] >= 0) & (pairs[:, 0] < n_users)]
    pairs = torch.unique(pairs, dim=0)
    counts = torch.bincount(pairs[:, 0], minlength=n_users)
    rowptr = torch.zeros(n_users + 1, dtype=torch.int64, device=pairs.device)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, pairs[:, 1].to(torch.int32)


def _user_items(data: UserItemInteractionsDataset, user_id: int) -> torch.Tensor:
    return data.interactions[data.interactions[:, 0] == user_id, 1]


class PointWiseDataset(Dataset):
    def __init__(self, data: UserItemInteractionsDataset):
        self.data = data

    def __len__(self) -> int:
        return self.data.number_of_interactions

    def __getitem__(self, index: int) -> PointWiseRow:
        user_id, item_id = self.data.interactions[index]
        # the reference indexes interaction_scores with item_id (base_datasets.py:52), a bug; the score
        # of interaction `index` is what is meant
        return (user_id, item_id, get_user_features(self.data, user_id), get_item_features(self.data, item_id),
                self.data.interaction_scores[index])

    def loader(self, **loader_params) -> DataLoader:
        return DataLoader(self, **loader_params)


class PairWiseDataset(IterableDataset):
    """BPR triple sampler: per user, max_sampled positives x max_sampled negatives drawn with
    replacement (reference: base_datasets.py:70-107).  `__iter__` is host-side Python, as in the reference;
    `sample_triples` draws the same thing for every user at once on the GPU (divrec_bpr_sample_triples)."""

    def __init__(self, data: UserItemInteractionsDataset, frozen: Optional[UserItemInteractionsDataset] = None,
                 max_sampled: int = 100):
        self.data = data
        self.frozen = frozen
        self.max_sampled = max_sampled
        self._csr = {}

    def sample_triples(self, device, seed: int = 0):
        """One pass of `__iter__` as three int64 tensors (user, positive, negative) on `device`, in the reference's
        iteration order (user-major, positive-major): max_sampled^2 triples per user, sampled with replacement from
        the user's positives and from items - positives - frozen.  Users without a positive or a negative are dropped
        (the reference raises IndexError on them).  Deterministic for a seed; not CPython's random stream."""
        from .. import _native
        if not (0 < self.max_sampled <= 256):
            raise ValueError("sample_triples needs 0 < max_sampled <= 256")
        key = str(device)
        if key not in self._csr:
            n_users = int(self.data.number_of_users)
            pos = self.data.interactions.to(device)
            excl = pos if self.frozen is None else torch.cat([pos, self.frozen.interactions.to(device)])
            self._csr[key] = interactions_to_csr(pos, n_users) + interactions_to_csr(excl, n_users)
        prp, pit, erp, eit = self._csr[key]
        u, p, n = _native.bpr_sample_triples(prp, pit, erp, eit, int(self.data.number_of_items), self.max_sampled, seed)
        keep = u >= 0
        if not bool(keep.all()):
            u, p, n = u[keep], p[keep], n[keep]
        return u, p, n

    def __iter__(self) -> Iterator[PairWiseRow]:
        everything = frozenset(range(int(self.data.number_of_items)))
        for user_id in range(int(self.data.number_of_users)):
            pos = frozenset(_user_items(self.data, user_id).tolist())
            blocked = pos if self.frozen is None else pos | frozenset(_user_items(self.frozen, user_id).tolist())
            pos, neg = list(pos), list(everything - blocked)
            if self.max_sampled > 0:
                pos = random.choices(pos, k=self.max_sampled)
                neg = random.choices(neg, k=self.max_sampled)
            uf = get_user_features(self.data, user_id)
            for p in pos:
                pf = get_item_features(self.data, p)
                for n in neg:
                    yield (user_id, p, n, uf, pf, get_item_features(self.data, n))

    def loader(self, **loader_params) -> DataLoader:
        return DataLoader(self, **loader_params)


class RankingDataset(IterableDataset):
    """One row per user: the candidate items (everything except the user's `frozen` interactions;
    the dataset's own positives stay recommendable, reference docstring base_datasets.py:125)."""

    def __init__(self, data: UserItemInteractionsDataset, frozen: Optional[UserItemInteractionsDataset] = None):
        self.data = data
        self.frozen = frozen
        self._csr = {}

    def seen_csr(self, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """CSR of the mask on `device`, built once per device: (rowptr int64 [U+1], items int32)."""
        key = str(device)
        if key not in self._csr:
            inter = None if self.frozen is None else self.frozen.interactions
            if inter is not None:
                inter = inter.to(device)
            rowptr, items = interactions_to_csr(inter, int(self.data.number_of_users), int(self.data.number_of_items))
            self._csr[key] = (rowptr.to(device), items.to(device))
        return self._csr[key]

    def __iter__(self) -> Iterator[RankingRow]:
        n_items = int(self.data.number_of_items)
        rowptr, items = self.seen_csr("cpu")
        for user_id in range(int(self.data.number_of_users)):
            allowed = torch.ones(n_items, dtype=torch.bool)
            allowed[items[int(rowptr[user_id]):int(rowptr[user_id + 1])].long()] = False
            candidates = torch.nonzero(allowed).flatten()  # ascending item id, like the reference
            uf = itf = None
            if self.data.has_user_features():
                uf = self.data.user_features.features[user_id].expand(candidates.numel(), -1)
            if self.data.has_item_features():
                itf = self.data.item_features.features[candidates]
            yield (torch.full((candidates.numel(),), user_id), _user_items(self.data, user_id), candidates, uf, itf)
