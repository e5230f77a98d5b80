"""Random scorer kept for import compatibility (divrec/models/random_model.py:8-21); no arithmetic."""
from typing import Optional

import torch

from .base_models import RankingModel


class RandomModel(RankingModel):
    def __init__(self, no_users: int, no_items: int):
        super().__init__()
        self.no_users = no_users
        self.no_items = no_items

    def forward(self, user_id, item_id, user_features: Optional[torch.Tensor] = None,
                item_features: Optional[torch.Tensor] = None) -> torch.Tensor:
        return torch.randn(user_id.size(), device=user_id.device)
