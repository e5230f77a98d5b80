from ..losses.intra_list_diversity_score import IntraListDiversityScore
from .auc_score import AUCScore
from .entropy_diversity_score import EntropyDiversityScore
from .relevance import (
    AveragePrecisionAtKScore,
    HitRateScore,
    MeanAveragePrecisionAtKScore,
    NDCGScore,
    PrecisionAtKScore,
    RecallAtKScore,
    average_precision_at_k,
    normalized_discounted_cumulative_gain,
    precision_at_k,
    recall_at_k,
)
from .popularity_rank_correlation_for_items import PRI, avg_rank, rank

__all__ = [
    "AUCScore",
    "AveragePrecisionAtKScore",
    "EntropyDiversityScore",
    "HitRateScore",
    "IntraListDiversityScore",
    "MeanAveragePrecisionAtKScore",
    "NDCGScore",
    "avg_rank",
    "rank",
    "PRI",
    "PrecisionAtKScore",
    "RecallAtKScore",
    "average_precision_at_k",
    "normalized_discounted_cumulative_gain",
    "precision_at_k",
    "recall_at_k",
]
