"""Normalised entropy of the recommended-item distribution (divrec/metrics/entropy_diversity_score.py:8-26).

The reference's torch.unique(recommendations, return_counts=True) -- a sort -- becomes an int32
histogram over item ids (divrec_coverage_hist, atomics in L2) followed by a single-CTA entropy
reduction (divrec_entropy_from_counts).  `coverage_counts` exposes the histogram itself: its non-zero
entries are exactly the reference's (values, counts).
"""
import math

import torch

from .. import _native
from ..losses.base_losses import DatasetAwareLoss, RecommendationsAwareLoss


class EntropyDiversityScore(RecommendationsAwareLoss, DatasetAwareLoss):
    def __init__(self, *args, **kwargs):
        DatasetAwareLoss.__init__(self, *args, **kwargs)
        RecommendationsAwareLoss.__init__(self, *args, **kwargs)
        self.max_entropy = math.log(int(self.dataset.number_of_items))

    def coverage_counts(self, recommendations: torch.LongTensor) -> torch.Tensor:
        """int32 [number_of_items]: how often each item was recommended."""
        return _native.coverage_hist(recommendations, int(self.dataset.number_of_items))

    def forward(self, interactions: torch.LongTensor, recommendations: torch.LongTensor):
        return self.recommendations_loss(interactions, recommendations)

    def recommendations_loss(self, interactions: torch.LongTensor, recommendations: torch.LongTensor) -> torch.Tensor:
        counts = self.coverage_counts(recommendations)
        actual_entropy = _native.entropy_from_counts(counts)[0]
        min_entropy = math.log(recommendations.size(1))
        return (actual_entropy - min_entropy) / (self.max_entropy - min_entropy)
