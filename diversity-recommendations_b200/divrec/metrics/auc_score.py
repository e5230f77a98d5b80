"""AUC indicator (pos >= neg) as int32 (divrec/metrics/auc_score.py:6-10); the mean comes from
ScoreWithReduction (integer sum / count -> float, base_losses.py:25-26)."""
import torch

from .. import _native
from ..losses.base_losses import PairWiseLoss


class AUCScore(PairWiseLoss):
    def pair_wise(self, positives: torch.Tensor, negatives: torch.Tensor) -> torch.Tensor:
        return _native.auc_indicator(positives.detach(), negatives.detach())
