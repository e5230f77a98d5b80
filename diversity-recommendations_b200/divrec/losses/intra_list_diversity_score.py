"""Intra-list diversity (divrec/losses/intra_list_diversity_score.py:8-63).

Per user:  sum_{a<b} D[rec_a, rec_b] / (k (k - 1))  with pairs taken in rank order.  The reference
walks itertools.combinations in Python over a caller-supplied dense [I,I] matrix; here
  * `distance_matrix=D`      -> divrec_ild_matrix  (one warp per user over the same D), or
  * `item_embeddings=V`      -> divrec_ild_cosine  (D = 1 - cos is never materialised: rows of V are
                                gathered with 128-bit loads and reduced per user),
  * `item_labels=l`          -> divrec_ild_labels  (D[i,j] = (l_i == l_j), the binary-unfairness distance).
`item_embeddings` / `item_labels` are additions; a caller that passes `distance_matrix` gets the
reference's numbers.

Deliberate fix: the reference reduces inside recommendations_loss AND again
in forward, which makes the default reduction="mean" raise IndexError; here forward applies the
reduction once and equals the reference's recommendations_loss for all three reductions.
"""
from typing import Optional

import torch

from .. import _native
from .base_losses import DatasetAwareLoss, RecommendationsAwareLoss


class IntraListDiversityScore(RecommendationsAwareLoss):
    def __init__(self, *args, distance_matrix: Optional[torch.Tensor] = None,
                 item_embeddings: Optional[torch.Tensor] = None, item_labels: Optional[torch.Tensor] = None,
                 ild_mode: str = "sumvec", **kwargs):
        RecommendationsAwareLoss.__init__(self, *args, **kwargs)
        given = [x is not None for x in (distance_matrix, item_embeddings, item_labels)]
        if sum(given) != 1:
            raise TypeError("IntraListDiversityScore needs exactly one of distance_matrix=, item_embeddings=, item_labels=")
        if ild_mode not in ("sumvec", "pairwise"):
            raise ValueError("ild_mode must be 'sumvec' or 'pairwise'")
        self.distance_matrix = distance_matrix
        self.item_embeddings = item_embeddings
        self.item_labels = item_labels
        self.ild_mode = ild_mode

    def per_user(self, recommendations: torch.LongTensor) -> torch.Tensor:
        """[U] fp32 on the device of the recommendations (which must be CUDA)."""
        if self.item_embeddings is not None:
            V = self.item_embeddings.detach().to(recommendations.device)
            mode = _native.ILD_SUMVEC if self.ild_mode == "sumvec" else _native.ILD_PAIRWISE
            return _native.ild_cosine(V, recommendations, mode)
        if self.item_labels is not None:
            return _native.ild_labels(self.item_labels.to(recommendations.device), recommendations)
        D = self.distance_matrix.to(recommendations.device)
        if D.dtype not in (torch.float32, torch.int32):
            D = D.to(torch.int32 if not D.dtype.is_floating_point else torch.float32)
        return _native.ild_matrix(D, recommendations)

    def recommendations_loss(self, interactions: torch.LongTensor, recommendations: torch.LongTensor) -> torch.Tensor:
        return self.reduce_loss_values(self.per_user(recommendations))

    def forward(self, interactions: torch.LongTensor, recommendations: torch.LongTensor):
        return self.recommendations_loss(interactions, recommendations)

    @staticmethod
    def user_ild(user_recommendations: torch.Tensor, distance_matrix: torch.Tensor) -> float:
        """Un-normalised pair sum of one user (reference staticmethod, :36-42)."""
        recs = user_recommendations.reshape(1, -1).to(distance_matrix.device)
        k = recs.shape[1]
        D = distance_matrix if distance_matrix.dtype in (torch.float32, torch.int32) else distance_matrix.float()
        return float(_native.ild_matrix(D, recs)[0]) * k * (k - 1)


class IntraListBinaryUnfairnessScore(IntraListDiversityScore, DatasetAwareLoss):
    """D[i,j] = (label_i == label_j) from the item-feature column `items_partition_feature`
    (reference :45-63).  The [I,I] int matrix of the reference is replaced by the label vector."""

    def __init__(self, *args, items_partition_feature: str = "partition", **kwargs):
        DatasetAwareLoss.__init__(self, *args, **kwargs)
        self.items_partition_feature = items_partition_feature
        kwargs.pop("dataset", None)
        assert self.items_partition_feature in self.dataset.item_features
        IntraListDiversityScore.__init__(self, *args, item_labels=self.dataset.item_features[items_partition_feature], **kwargs)

    def get_distance_matrix(self):
        labels = self.dataset.item_features[self.items_partition_feature]
        return (labels.unsqueeze(0) == labels.unsqueeze(1)).int()
