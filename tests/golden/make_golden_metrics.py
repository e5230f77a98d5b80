"""Golden fixtures for the list metrics, produced by the REAL reference (dev container only: needs /root/reference):

    python tests/golden/make_golden_metrics.py

Reference entry points exercised (file:line relative to the reference repository):
  precision_at_k / PrecisionAtKScore                 divrec/metrics/precision_at_k.py:6-32
  recall_at_k                                        divrec/metrics/recall_at_k.py:6-22
  average_precision_at_k / MeanAveragePrecision...   divrec/metrics/average_precision_at_k.py:6-38
  normalized_discounted_cumulative_gain / NDCGScore  divrec/metrics/normalized_discounted_cumulative_gain.py:6-30
  avg_rank, rank, PRI                                divrec/metrics/popularity_rank_correlation_for_items.py:6-63
  IntraListBinaryUnfairnessScore                     divrec/losses/intra_list_diversity_score.py:45-63
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from divrec.datasets import UserItemInteractionsDataset  # noqa: E402  (the reference)
from divrec.datasets.storages import Features  # noqa: E402
from divrec.losses.intra_list_diversity_score import IntraListBinaryUnfairnessScore  # noqa: E402
from divrec.metrics import (PRI, MeanAveragePrecisionAtKScore, NDCGScore, PrecisionAtKScore, RecallAtKScore,  # noqa: E402
                            average_precision_at_k, avg_rank, normalized_discounted_cumulative_gain, precision_at_k, recall_at_k)

assert os.path.realpath(sys.modules["divrec"].__file__).startswith(REF), "must import the reference, not the shim"


def case(name, nU, nI, k, nnz, seed):
    g = torch.Generator().manual_seed(seed)
    inter = torch.stack([torch.randint(0, nU, (nnz,), generator=g), torch.randint(0, nI, (nnz,), generator=g)], 1)
    inter[:nU, 0] = torch.arange(nU)          # every user has at least one interaction (recall's denominator)
    inter[-40:] = inter[:40]                  # duplicated rows: the reference counts them in recall's denominator
    recs = torch.stack([torch.randperm(nI, generator=g)[:k] for _ in range(nU)])
    labels = torch.randint(0, 4, (nI,), generator=g)
    feats = Features(feature_names=["partition"], features=labels.reshape(-1, 1))
    data = UserItemInteractionsDataset(interactions=inter, item_features=feats, number_of_users=nU, number_of_items=nI)
    items, ranks = avg_rank(recs)
    pri = PRI(dataset=data)(inter, recs)
    unfair = IntraListBinaryUnfairnessScore(dataset=data, reduction="none").recommendations_loss(inter, recs)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        interactions=inter.numpy(), recs=recs.numpy(), labels=labels.numpy(), nI=nI,
        precision=precision_at_k(inter, recs).numpy(), recall=recall_at_k(inter, recs).numpy(),
        ap=average_precision_at_k(inter, recs).numpy(), ndcg=normalized_discounted_cumulative_gain(inter, recs).numpy(),
        precision_mean=PrecisionAtKScore()(inter, recs).numpy(), recall_mean=RecallAtKScore()(inter, recs).numpy(),
        map_mean=MeanAveragePrecisionAtKScore()(inter, recs).numpy(), ndcg_mean=NDCGScore()(inter, recs).numpy(),
        avg_rank_items=items.numpy(), avg_rank_values=ranks.numpy(), pri=pri.numpy(), unfairness=unfair.numpy(),
    )
    print(name, "precision", float(precision_at_k(inter, recs).mean()), "pri", float(pri), "unfairness", float(unfair.mean()))


if __name__ == "__main__":
    case("metrics_small", 60, 90, 12, 900, 11)
    case("metrics_k50", 40, 400, 50, 3000, 12)
