"""Relevance metrics over [U,k] recommendation lists: precision / recall / AP / NDCG @k
(divrec/metrics/precision_at_k.py:6-32, recall_at_k.py:6-29, average_precision_at_k.py:6-38,
normalized_discounted_cumulative_gain.py:6-30).

The reference loops over users in Python, each iteration scanning all interactions
(`interactions[:,0] == user`) and calling torch.isin.  Here the interactions become one CSR and a
single kernel (divrec_relevance_metrics, one warp per user, binary-search membership) produces all
four per-user values at once.  Reference quirks kept: recall divides by the raw number of the user's
interaction rows (duplicates counted, 0 -> nan/inf); "average precision" sums cumsum(rel)/rank over
ALL ranks (not only the relevant ones) and divides by k.
"""
import torch

from .. import _native
from ..datasets.base_datasets import interactions_to_csr
from ..losses.base_losses import RecommendationsAwareLoss


_last = {"key": None, "value": None}


def _per_user(interactions: torch.LongTensor, recommendations: torch.LongTensor) -> torch.Tensor:
    """[4, U] fp32: precision, recall, average precision, ndcg.  The four metric functions of one evaluation are called on
    the same two tensors: the CSR build and the kernel run once and are shared (keyed on the tensors' storage and version)."""
    key = (interactions.data_ptr(), interactions._version, tuple(interactions.shape), recommendations.data_ptr(),
           recommendations._version, tuple(recommendations.shape), str(recommendations.device))
    if _last["key"] == key:
        return _last["value"]
    dev = recommendations.device
    n_users = recommendations.size(0)
    inter = interactions.to(dev)
    inter = inter[inter[:, 0] < n_users]
    rowptr, items = interactions_to_csr(inter, n_users)
    raw_counts = torch.bincount(inter[:, 0], minlength=n_users).to(torch.int32)
    out = _native.relevance_metrics(recommendations, rowptr, items, raw_counts)
    _last["key"], _last["value"] = key, out
    return out


def precision_at_k(interactions: torch.LongTensor, recommendations: torch.LongTensor):
    return _per_user(interactions, recommendations)[0]


def recall_at_k(interactions: torch.LongTensor, recommendations: torch.LongTensor):
    return _per_user(interactions, recommendations)[1]


def average_precision_at_k(interactions: torch.LongTensor, recommendations: torch.LongTensor):
    return _per_user(interactions, recommendations)[2]


def normalized_discounted_cumulative_gain(interactions: torch.LongTensor, recommendations: torch.LongTensor):
    return _per_user(interactions, recommendations)[3]


class PrecisionAtKScore(RecommendationsAwareLoss):
    def recommendations_loss(self, interactions, recommendations) -> torch.Tensor:
        return precision_at_k(interactions, recommendations)


HitRateScore = PrecisionAtKScore


class RecallAtKScore(RecommendationsAwareLoss):
    def recommendations_loss(self, interactions, recommendations) -> torch.Tensor:
        return recall_at_k(interactions, recommendations)


class AveragePrecisionAtKScore(RecommendationsAwareLoss):
    def recommendations_loss(self, interactions, recommendations) -> torch.Tensor:
        return average_precision_at_k(interactions, recommendations)


class MeanAveragePrecisionAtKScore(AveragePrecisionAtKScore):
    def forward(self, interactions, recommendations):
        return torch.mean(average_precision_at_k(interactions, recommendations))


class NDCGScore(RecommendationsAwareLoss):
    def recommendations_loss(self, interactions, recommendations) -> torch.Tensor:
        return normalized_discounted_cumulative_gain(interactions, recommendations)
