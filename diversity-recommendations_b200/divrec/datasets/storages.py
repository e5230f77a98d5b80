"""Interaction storage with the reference's field names and validation
(divrec/datasets/storages.py:7-94): `interactions` is an [nnz, 2] int64 tensor of (user, item)
pairs in any order; the counters default to max id + 1.

Differences from the reference, both deliberate: the inferred counters are cast to Python ints
(the reference leaves 0-dim tensors behind), and the id-range check is
a min/max test instead of torch.isin against an arange (same accept/reject behaviour)."""
from dataclasses import dataclass, field
from typing import List, Optional

import torch


@dataclass
class Features:
    """A table with named columns: features [rows, n], feature_names of length n."""

    features: torch.Tensor = None
    feature_names: List[str] = field(default_factory=list)
    number_of_features: int = 0

    def __post_init__(self):
        self.number_of_features = max(self.number_of_features, self.features.size(1))
        assert len(self.feature_names) == self.number_of_features

    def __len__(self):
        return self.features.size(0)

    def __getitem__(self, name):
        return self.features[:, self.feature_names.index(name)]

    def __contains__(self, name):
        return name in self.feature_names


@dataclass
class UserItemInteractionsDataset:
    interactions: Optional[torch.LongTensor] = None
    interaction_scores: Optional[torch.Tensor] = None
    user_features: Optional[Features] = None
    item_features: Optional[Features] = None
    number_of_interactions: int = 0
    number_of_users: int = 0
    number_of_items: int = 0

    def __post_init__(self):
        if self.interaction_scores is not None:
            assert self.interactions is not None
        if self.interactions is not None:
            nnz, cols = self.interactions.size()
            assert cols == 2
            self.number_of_interactions = nnz
            if self.interaction_scores is None:
                self.interaction_scores = torch.ones(nnz)
            assert self.interaction_scores.size(0) == nnz
            if nnz > 0:
                lo = self.interactions.min(dim=0).values
                hi = self.interactions.max(dim=0).values
                assert int(lo[0]) >= 0 and int(lo[1]) >= 0
                self.number_of_users = max(int(self.number_of_users), int(hi[0]) + 1)
                self.number_of_items = max(int(self.number_of_items), int(hi[1]) + 1)
        for table, attr in ((self.user_features, "number_of_users"), (self.item_features, "number_of_items")):
            if table is not None:
                if getattr(self, attr) == 0:
                    setattr(self, attr, len(table))
                assert getattr(self, attr) == len(table)

    def has_interactions(self) -> bool:
        return self.interactions is not None

    def has_user_features(self) -> bool:
        return self.user_features is not None

    def has_item_features(self) -> bool:
        return self.item_features is not None


def get_user_features(data: UserItemInteractionsDataset, user_id: int) -> Optional[torch.Tensor]:
    return None if data.user_features is None else data.user_features.features[user_id]


def get_item_features(data: UserItemInteractionsDataset, item_id: int) -> Optional[torch.Tensor]:
    return None if data.item_features is None else data.item_features.features[item_id]
