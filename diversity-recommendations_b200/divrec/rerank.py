"""Diversity-increasing re-rank (BASELINE.json config 5): greedy MMR over each user's top-C candidates.

Not part of the reference (its README announces "tools for increasing ... diversity" but ships none that run);
the semantics are fixed by oracle/divrec_oracle.py::mmr_rerank and DESIGN.md.  Typical use:

    scores, cands = divrec._native.score_topk(U, V, 500, seen_rowptr, seen_items)     # top-500 by relevance
    items, rel = mmr_rerank(V, cands, scores, k=100, lam=0.5)                         # 100 diverse picks
"""
import torch

from . import _native


def mmr_rerank(item_embeddings: torch.Tensor, candidates: torch.LongTensor, candidate_scores: torch.Tensor,
               k: int, lam: float = 0.5):
    """(items [U,k] int64, relevance scores [U,k] fp32).  First pick = best-scoring candidate, then repeatedly
    argmax  lam * score - (1 - lam) * max cosine similarity to the items already picked."""
    return _native.mmr_rerank(item_embeddings.detach(), candidates, candidate_scores, k, lam)
