"""Golden fixture for the BPR triple sampler: the REAL reference's PairWiseDataset.__iter__
(divrec/datasets/base_datasets.py:70-107) under a fixed `random.seed`.

    python tests/golden/make_golden_pairwise.py      (dev container only: needs /root/reference)

Writes pairwise_small.npz: the interactions, the frozen interactions, sizes, max_sampled, the Python seed and the
(user, positive, negative) triples in iteration order.  tests/test_oracle_golden.py replays the oracle's restatement
under the same seed and requires the identical sequence.
"""
import os
import random
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from divrec.datasets import PairWiseDataset, UserItemInteractionsDataset  # noqa: E402  (the reference)

assert os.path.realpath(sys.modules["divrec"].__file__).startswith(REF), "must import the reference, not the shim"

gen = torch.Generator().manual_seed(77)
nU, nI, m, seed = 23, 61, 4, 20241018
inter = torch.unique(torch.stack([torch.randint(0, nU, (160,), generator=gen), torch.randint(0, nI, (160,), generator=gen)], 1), dim=0)
inter = torch.cat([inter, torch.stack([torch.arange(nU), torch.randint(0, nI, (nU,), generator=gen)], 1)])  # every user >= 1 positive
inter = inter[torch.randperm(inter.shape[0], generator=gen)]                                              # unsorted, with duplicates
frozen = torch.unique(torch.stack([torch.randint(0, nU, (90,), generator=gen), torch.randint(0, nI, (90,), generator=gen)], 1), dim=0)
data = UserItemInteractionsDataset(interactions=inter, number_of_users=nU, number_of_items=nI)
froz = UserItemInteractionsDataset(interactions=frozen, number_of_users=nU, number_of_items=nI)
random.seed(seed)
rows = [(u, p, n) for (u, p, n, _uf, _pf, _nf) in PairWiseDataset(data, frozen=froz, max_sampled=m)]
t = np.asarray(rows, dtype=np.int64)
assert t.shape == (nU * m * m, 3)
np.savez_compressed(os.path.join(HERE, "pairwise_small.npz"), interactions=inter.numpy(), frozen=frozen.numpy(), n_users=nU, n_items=nI,
                    max_sampled=m, py_seed=seed, triples=t)
print("wrote pairwise_small.npz", t.shape)
