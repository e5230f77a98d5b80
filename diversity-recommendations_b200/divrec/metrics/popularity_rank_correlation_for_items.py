"""Popularity-rank correlation of the recommended items, PRI
(divrec/metrics/popularity_rank_correlation_for_items.py:6-63).

avg_rank: the reference builds a Python dict item -> list of positions; here one histogram kernel
(divrec_coverage_hist_ranks) accumulates per item the count and the sum of positions.  The final
Pearson correlation over the distinct recommended items is a handful of reductions on [I] vectors.
"""
import torch

from .. import _native
from ..losses.base_losses import DatasetAwareLoss, RecommendationsAwareLoss


def rank(a: torch.Tensor, dim=-1, descending=False, stable=False):
    order = torch.argsort(a, dim=dim, descending=descending, stable=stable)
    return torch.argsort(order, dim=dim, stable=stable)


def spearman_rank_correlation(a: torch.Tensor, b: torch.Tensor, evaluate_rank: bool = True) -> torch.Tensor:
    assert a.size(0) == b.size(0)
    n = a.size(0)
    if evaluate_rank:
        diff = rank(a) - rank(b)
        return 1 - 6 * torch.sum(diff**2) / n / (n**2 - 1)
    a_std, a_mean = torch.std_mean(a)
    b_std, b_mean = torch.std_mean(b)
    return torch.mean((a - a_mean) * (b - b_mean)) / a_std / b_std


def avg_rank(recommendations: torch.LongTensor, number_of_items: int = None):
    """(items ascending, mean position of each item over all lists it appears in)"""
    n_items = int(recommendations.max()) + 1 if number_of_items is None else int(number_of_items)
    counts, rank_sums = _native.coverage_hist_ranks(recommendations, n_items)
    items = torch.nonzero(counts).flatten()
    return items, rank_sums[items].float() / counts[items].float()


class PRI(RecommendationsAwareLoss, DatasetAwareLoss):
    def __init__(self, *args, **kwargs):
        DatasetAwareLoss.__init__(self, *args, **kwargs)
        RecommendationsAwareLoss.__init__(self, *args, **kwargs)
        n_items = int(self.dataset.number_of_items)
        self.popularity = torch.bincount(self.dataset.interactions[:, 1], minlength=n_items).float()
        self.popularity_rank = rank(self.popularity, descending=True).float()

    def forward(self, interactions, recommendations):
        return self.recommendations_loss(interactions, recommendations)

    def recommendations_loss(self, interactions, recommendations) -> torch.Tensor:
        items, ranks = avg_rank(recommendations, int(self.dataset.number_of_items))
        return spearman_rank_correlation(self.popularity_rank.to(items.device)[items], ranks, evaluate_rank=False)
