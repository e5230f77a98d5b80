"""Dot-product matrix factorisation with the reference's constructor, attributes and state_dict keys
(divrec/models/matrix_factorization.py:8-28): two nn.Embedding tables, score = <U[user], V[item]>.

forward() is the paired scorer `divrec_pair_scores` of libdivrec_b200.so with a hand-written backward
(`divrec_pair_scores_bwd`, vectorised red.global.add scatter) producing the dense .grad tensors torch
optimisers expect.  The full-catalogue path does not go through forward(): see `recommend`.
"""
from typing import Optional, Tuple

import torch

from .. import _native
from .base_models import RankingModel


class _PairScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, V, user_id, item_id):
        ctx.save_for_backward(U, V, user_id, item_id)
        return _native.pair_scores(U, V, user_id, item_id)

    @staticmethod
    def backward(ctx, g):
        U, V, user_id, item_id = ctx.saved_tensors
        gU = torch.zeros_like(U) if ctx.needs_input_grad[0] else None
        gV = torch.zeros_like(V) if ctx.needs_input_grad[1] else None
        _native.pair_scores_bwd(U, V, user_id, item_id, g.contiguous(), gU, gV)
        return gU, gV, None, None


class MatrixFactorization(RankingModel):
    def __init__(self, no_users: int, no_items: int, embedding_dim: int):
        super().__init__()
        self.no_users = no_users
        self.no_items = no_items
        self.embedding_dim = embedding_dim
        self.user_embeddings = torch.nn.Embedding(no_users, embedding_dim)
        self.item_embeddings = torch.nn.Embedding(no_items, embedding_dim)

    def forward(self, user_id: torch.LongTensor, item_id: torch.LongTensor,
                user_features: Optional[torch.Tensor] = None, item_features: Optional[torch.Tensor] = None) -> torch.Tensor:
        U, V = self.user_embeddings.weight, self.item_embeddings.weight
        if not U.is_cuda:
            raise TypeError("divrec_b200.MatrixFactorization computes on a B200 only: move the model with .to('cuda')")
        user_id = user_id.to(device=U.device, dtype=torch.int64)
        item_id = item_id.to(device=U.device, dtype=torch.int64)
        return _PairScores.apply(U, V, user_id, item_id)

    @torch.no_grad()
    def recommend(self, k: int, seen_rowptr: Optional[torch.Tensor] = None, seen_items: Optional[torch.Tensor] = None,
                  user_range: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Full-catalogue scoring + seen mask + top-k in one fused pass (divrec_score_topk).
        Returns (items [U,k] int64, scores [U,k] fp32) in (score desc, item id asc) order."""
        U, V = self.user_embeddings.weight, self.item_embeddings.weight
        if user_range is not None:
            a, b = user_range
            U = U[a:b]
            if seen_rowptr is not None:
                seen_rowptr = seen_rowptr[a:b + 1]
        scores, items = _native.score_topk(U, V, k, seen_rowptr, seen_items)
        return items, scores
