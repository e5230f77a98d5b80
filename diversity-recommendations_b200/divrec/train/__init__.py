from .utils import (
    bpr_step,
    bpr_train_step,
    get_model_recommendations,
    host_recommend_and_score,
    pair_wise_score_loop,
    pair_wise_train_loop,
    point_wise_score_loop,
    point_wise_train_loop,
    recommendations_score_loop,
    recommendations_train_loop,
)

__all__ = [
    "point_wise_score_loop",
    "pair_wise_score_loop",
    "get_model_recommendations",
    "recommendations_score_loop",
    "point_wise_train_loop",
    "pair_wise_train_loop",
    "recommendations_train_loop",
    "bpr_step",
    "bpr_train_step",
    "host_recommend_and_score",
]
