"""Generate the golden fixtures in this directory by running the REAL reference.

    python tests/golden/make_golden.py          (dev container only: needs /root/reference)

The reference (amtsyplov/diversity-recommendations) has no tests or golden vectors of its own, so
parity is pinned on outputs of the reference itself: this script imports the unmodified package from
/root/reference, runs the hot-path entry points on seeded synthetic inputs and stores inputs +
outputs as small .npz files.  /root/reference does not exist on the GPU box, so tests only read
the committed .npz files.  Reference entry points exercised (file:line relative to the reference):
  get_model_recommendations            divrec/train/utils.py:53-77  over RankingDataset (base_datasets.py:113-171)
  MatrixFactorization.forward          divrec/models/matrix_factorization.py:19-28
  IntraListDiversityScore              divrec/losses/intra_list_diversity_score.py:8-42
  EntropyDiversityScore                divrec/metrics/entropy_diversity_score.py:8-26
  LogSigmoidDifferenceLoss + backward  divrec/losses/log_sigmoid_difference_loss.py:6-14, train/utils.py:145-149
  AUCScore                             divrec/metrics/auc_score.py:6-10
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from divrec.datasets import RankingDataset, UserItemInteractionsDataset  # noqa: E402  (the reference)
from divrec.losses import IntraListDiversityScore, LogSigmoidDifferenceLoss  # noqa: E402
from divrec.metrics import AUCScore, EntropyDiversityScore  # noqa: E402
from divrec.models import MatrixFactorization  # noqa: E402
from divrec.train import get_model_recommendations  # noqa: E402

from oracle import divrec_oracle as O  # noqa: E402  (input generators only)

assert os.path.realpath(sys.modules["divrec"].__file__).startswith(REF), "must import the reference, not the shim"


def make_model(U, V):
    m = MatrixFactorization(U.shape[0], V.shape[0], U.shape[1])
    with torch.no_grad():
        m.user_embeddings.weight.copy_(U)
        m.item_embeddings.weight.copy_(V)
    return m


def eval_case(name, nU, nI, d, k, nnz, kind, seed, ild_users=None):
    g = torch.Generator().manual_seed(seed)
    U = O.make_embeddings(nU, d, kind, g)
    V = O.make_embeddings(nI, d, kind, g)
    seen = O.make_seen(nU, nI, nnz, g)
    # `data` (the test interactions) only supplies number_of_users / number_of_items and the positives
    # the reference returns but does not use; `frozen` is the mask.
    data = UserItemInteractionsDataset(interactions=torch.tensor([[nU - 1, nI - 1]]), number_of_users=nU, number_of_items=nI)
    frozen = UserItemInteractionsDataset(interactions=seen, number_of_users=nU, number_of_items=nI)
    model = make_model(U, V)
    model.eval()
    with torch.no_grad():
        recs = get_model_recommendations(RankingDataset(data, frozen=frozen), model, k)
        rec_scores = torch.stack([model(torch.full((k,), u), recs[u]) for u in range(nU)])
    Vd = V.double()
    En = Vd / Vd.norm(dim=1, keepdim=True)
    D = (1.0 - En @ En.T).float()
    nild = nU if ild_users is None else ild_users
    ild_none = IntraListDiversityScore(distance_matrix=D, reduction="none").recommendations_loss(None, recs[:nild])
    ild_sum = IntraListDiversityScore(distance_matrix=D, reduction="sum").recommendations_loss(None, recs[:nild])
    ild_mean = IntraListDiversityScore(distance_matrix=D, reduction="mean").recommendations_loss(None, recs[:nild])
    ent = EntropyDiversityScore(dataset=data)(None, recs)
    uniq, counts = torch.unique(recs, return_counts=True)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        U=U.numpy(), V=V.numpy(), seen=seen.numpy(), k=k, kind=kind,
        recs=recs.numpy(), rec_scores=rec_scores.numpy(),
        ild_users=nild, ild_none=ild_none.numpy(), ild_sum=ild_sum.numpy(), ild_mean=ild_mean.numpy(),
        entropy=ent.numpy(), cov_values=uniq.numpy(), cov_counts=counts.numpy(),
    )
    print(name, "recs", tuple(recs.shape), "ild_mean", float(ild_mean), "entropy", float(ent), "distinct", uniq.numel())


def bpr_case(name, nU, nI, d, B, seed):
    g = torch.Generator().manual_seed(seed)
    U = O.make_embeddings(nU, d, "gauss", g)
    V = O.make_embeddings(nI, d, "gauss", g)
    u = torch.randint(0, nU, (B,), generator=g)
    p = torch.randint(0, nI, (B,), generator=g)
    n = torch.randint(0, nI, (B,), generator=g)
    u[: B // 4] = u[0]  # heavy duplicates: the scatter-add must accumulate
    p[: B // 8] = p[0]
    model = make_model(U, V)
    model.train()
    pos = model(u, p, None, None)
    neg = model(u, n, None, None)
    loss = LogSigmoidDifferenceLoss()(pos, neg)
    loss.backward()
    auc = AUCScore()(pos, neg)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        U=U.numpy(), V=V.numpy(), u=u.numpy(), p=p.numpy(), n=n.numpy(),
        pos=pos.detach().numpy(), neg=neg.detach().numpy(), loss=loss.detach().numpy(), auc=auc.numpy(),
        gU=model.user_embeddings.weight.grad.numpy(), gV=model.item_embeddings.weight.grad.numpy(),
    )
    print(name, "loss", float(loss), "auc", float(auc))


if __name__ == "__main__":
    torch.manual_seed(0)
    # cfg1 = BASELINE.json configs[0]: MovieLens-100k shape, the reference's own CPU-runnable case
    eval_case("cfg1_gauss", 943, 1682, 64, 10, 90570, "gauss", 0)
    # Tier-E twin: ties everywhere.  The reference's argsort is NOT stable, so its order on ties is
    # arbitrary; tests compare the multiset of scores and the tie-free prefix for this file.
    eval_case("cfg1_grid", 943, 1682, 64, 10, 90570, "grid", 1)
    # wider lists / d=128 / k=100 on a small catalogue (ILD over 64 users keeps the Python loop short)
    eval_case("small_k100_gauss", 96, 3000, 128, 100, 20000, "gauss", 2, ild_users=64)
    bpr_case("bpr_small", 500, 700, 128, 4096, 3)
    bpr_case("bpr_d64", 300, 200, 64, 1000, 4)
