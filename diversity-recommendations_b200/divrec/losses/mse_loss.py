"""Squared error, kept for import compatibility (divrec/losses/mse_loss.py:6-10); the point-wise path is
outside the accelerated hot path, so this is plain tensor arithmetic."""
import torch

from .base_losses import PointWiseLoss


class MSELoss(PointWiseLoss):
    def point_wise(self, true_relevance: torch.Tensor, predicted_relevance: torch.Tensor) -> torch.Tensor:
        return torch.square(true_relevance - predicted_relevance)
