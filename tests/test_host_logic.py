"""Host-side mirror of the reference interface: names, reductions, dataset -> CSR, error behaviour."""
import pytest
import torch

from divrec.datasets import Features, RankingDataset, UserItemInteractionsDataset
from divrec.datasets.base_datasets import interactions_to_csr
from divrec.losses import (IntraListDiversityScore, LogSigmoidDifferenceLoss, MSELoss, PairWiseLoss,
                           RecommendationsAwareLoss, ScoreWithReduction)
from divrec.metrics import (AUCScore, EntropyDiversityScore, HitRateScore, MeanAveragePrecisionAtKScore, NDCGScore, PRI,
                            PrecisionAtKScore, RecallAtKScore)
from divrec.models import MatrixFactorization, RankingModel
from divrec.utils import to_camel_case
from oracle import divrec_oracle as O


def test_metric_names_match_reference():
    # SURVEY.md appendix item 7 (probed on the reference)
    ds = UserItemInteractionsDataset(interactions=torch.tensor([[0, 1], [1, 0]]))
    assert IntraListDiversityScore(distance_matrix=torch.zeros(2, 2)).name == "intra_list_diversity_score"
    assert EntropyDiversityScore(dataset=ds).name == "entropy_diversity_score"
    assert LogSigmoidDifferenceLoss().name == "log_sigmoid_difference_loss"
    assert AUCScore().name == "auc_score"
    assert RecallAtKScore().name == "recall_atk_score"
    assert PrecisionAtKScore().name == "precision_atk_score"
    assert MeanAveragePrecisionAtKScore().name == "mean_average_precision_atk_score"
    assert NDCGScore().name == "ndcg_score"
    assert PRI(dataset=ds).name == "pri"
    assert HitRateScore is PrecisionAtKScore
    assert to_camel_case("MSELoss") == "mse_loss"


def test_reduction_contract():
    s = ScoreWithReduction(reduction="mean")
    v = torch.tensor([1.0, 2.0, 6.0])
    assert float(s.reduce_loss_values(v)) == 3.0
    assert float(ScoreWithReduction(reduction="sum").reduce_loss_values(v)) == 9.0
    assert ScoreWithReduction(reduction="none").reduce_loss_values(v) is v
    assert ScoreWithReduction(reduce=False).reduce_loss_values(v) is v
    with pytest.raises(AssertionError):
        ScoreWithReduction(reduction="median")
    # reference quirk: LogSigmoidDifferenceLoss ignores reduction kwargs (always mean)
    loss = LogSigmoidDifferenceLoss(reduction="sum")
    assert loss.reduction == "mean" and loss.reduce
    with pytest.raises(NotImplementedError):
        PairWiseLoss().pair_wise(v, v)
    with pytest.raises(NotImplementedError):
        RecommendationsAwareLoss().recommendations_loss(None, None)
    assert torch.equal(MSELoss(reduction="none")(v, v + 2), torch.full((3,), 4.0))


def test_dataset_validation_and_counters():
    ds = UserItemInteractionsDataset(interactions=torch.tensor([[0, 3], [2, 1], [2, 1]]))
    assert (ds.number_of_users, ds.number_of_items, ds.number_of_interactions) == (3, 4, 3)
    assert isinstance(ds.number_of_users, int)
    ds2 = UserItemInteractionsDataset(interactions=torch.tensor([[0, 3]]), number_of_users=10, number_of_items=20)
    assert (ds2.number_of_users, ds2.number_of_items) == (10, 20)
    with pytest.raises(AssertionError):
        UserItemInteractionsDataset(interactions=torch.zeros(3, 3, dtype=torch.long))
    with pytest.raises(AssertionError):
        UserItemInteractionsDataset(interactions=torch.tensor([[0, -1]]))
    f = Features(features=torch.arange(6.0).reshape(3, 2), feature_names=["a", "partition"])
    assert "partition" in f and torch.equal(f["partition"], torch.tensor([1.0, 3.0, 5.0])) and len(f) == 3


def test_csr_matches_oracle_and_reference_candidates():
    gen = torch.Generator().manual_seed(3)
    seen = O.make_seen(40, 60, 500, gen)
    seen = torch.cat([seen, seen[:50]])[torch.randperm(seen.shape[0] + 50, generator=gen)]  # unsorted, duplicates
    rowptr, items = interactions_to_csr(seen, 40)
    orow, oitems = O.seen_csr(seen, 40)
    assert torch.equal(rowptr, orow) and torch.equal(items, oitems)
    data = UserItemInteractionsDataset(interactions=torch.tensor([[39, 59]]), number_of_users=40, number_of_items=60)
    frozen = UserItemInteractionsDataset(interactions=seen, number_of_users=40, number_of_items=60)
    rows = list(RankingDataset(data, frozen=frozen))
    assert len(rows) == 40
    for u in (0, 7, 39):
        user_rep, positives, cand, uf, itf = rows[u]
        assert torch.equal(cand, O.ranking_candidates(60, seen, u))
        assert user_rep.numel() == cand.numel() and int(user_rep[0]) == u and uf is None and itf is None
    # no frozen dataset -> nothing masked (the reference crashes here)
    assert list(RankingDataset(data))[5][2].numel() == 60


def test_model_surface_and_cpu_is_refused():
    m = MatrixFactorization(7, 9, 4)
    assert isinstance(m, RankingModel) and (m.no_users, m.no_items, m.embedding_dim) == (7, 9, 4)
    assert set(m.state_dict()) == {"user_embeddings.weight", "item_embeddings.weight"}
    assert m.user_embeddings.weight.shape == (7, 4) and m.item_embeddings.weight.shape == (9, 4)
    with pytest.raises(TypeError):  # no CPU path
        m(torch.tensor([0, 1]), torch.tensor([2, 3]))
    with pytest.raises(NotImplementedError):
        RankingModel()(None, None, None, None)


def test_native_wrappers_refuse_cpu_tensors():
    from divrec import _native
    with pytest.raises(TypeError):
        _native.score_topk(torch.zeros(4, 8), torch.zeros(5, 8), 2)
    with pytest.raises(TypeError):
        _native.ild_cosine(torch.zeros(5, 8), torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(TypeError):
        _native.coverage_hist(torch.zeros(2, 3, dtype=torch.long), 5)
    with pytest.raises(TypeError):
        IntraListDiversityScore(distance_matrix=torch.zeros(4, 4))(None, torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(TypeError):
        IntraListDiversityScore()
