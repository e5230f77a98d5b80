"""Loss / metric contracts of the reference (divrec/losses/base_losses.py:9-85): keyword arguments
`reduce=True, reduction in {"none","mean","sum"}`, `.name` = snake-cased class name, and the three
call shapes point-wise / pair-wise / recommendations-aware."""
from typing import Optional

import torch

from ..datasets import UserItemInteractionsDataset
from ..utils import to_camel_case

_REDUCTIONS = ("none", "mean", "sum")


class DatasetAwareLoss:
    def __init__(self, *args, dataset: Optional[UserItemInteractionsDataset] = None, **kwargs):
        self.dataset = dataset


class ScoreWithReduction:
    def __init__(self, *args, reduce: bool = True, reduction: str = "mean", **kwargs):
        assert reduction in _REDUCTIONS
        self.reduction = reduction
        self.reduce = bool(reduce) and reduction != "none"

    def reduce_loss_values(self, loss_values: torch.Tensor):
        if not self.reduce:
            return loss_values
        total = loss_values.sum(dim=0)
        return total / loss_values.size(0) if self.reduction == "mean" else total

    @property
    def name(self):
        return to_camel_case(type(self).__name__)


def _missing(obj, method: str):
    return NotImplementedError(f'Score [{type(obj).__name__}] is missing the required "{method}" function')


class PointWiseLoss(torch.nn.Module, ScoreWithReduction):
    def __init__(self, *args, **kwargs):
        torch.nn.Module.__init__(self)
        ScoreWithReduction.__init__(self, *args, **kwargs)

    def forward(self, true_relevance: torch.Tensor, predicted_relevance: torch.Tensor):
        return self.reduce_loss_values(self.point_wise(true_relevance, predicted_relevance))

    def point_wise(self, true_relevance: torch.Tensor, predicted_relevance: torch.Tensor) -> torch.Tensor:
        raise _missing(self, "point_wise")


class PairWiseLoss(torch.nn.Module, ScoreWithReduction):
    def __init__(self, *args, **kwargs):
        torch.nn.Module.__init__(self)
        ScoreWithReduction.__init__(self, *args, **kwargs)

    def forward(self, positives: torch.Tensor, negatives: torch.Tensor):
        return self.reduce_loss_values(self.pair_wise(positives, negatives))

    def pair_wise(self, positives: torch.Tensor, negatives: torch.Tensor) -> torch.Tensor:
        raise _missing(self, "pair_wise")


class RecommendationsAwareLoss(torch.nn.Module, ScoreWithReduction):
    def __init__(self, *args, **kwargs):
        torch.nn.Module.__init__(self)
        ScoreWithReduction.__init__(self, *args, **kwargs)

    def forward(self, interactions: torch.LongTensor, recommendations: torch.LongTensor):
        return self.reduce_loss_values(self.recommendations_loss(interactions, recommendations))

    def recommendations_loss(self, interactions: torch.LongTensor, recommendations: torch.LongTensor) -> torch.Tensor:
        raise _missing(self, "recommendations_loss")
