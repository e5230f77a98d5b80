"""Run the UNMODIFIED reference (oracle/_ref/divrec, see make_ref.py) on a synthetic workload and print one JSON line.

    python oracle/ref_run.py --users 200 [--workload cfg2|cfg1] [--threads N] [--check]

Runs in its own process because the reference's package is also called `divrec` (absolute imports), like the shim.
What is timed, literally as the reference's own experiment code calls it (divrec/train/utils.py:80-91):
    recommendations = get_model_recommendations(RankingDataset(data, frozen=seen), MatrixFactorization, k)   # train/utils.py:53-77
    IntraListDiversityScore(distance_matrix=D, reduction="none").recommendations_loss(interactions, recommendations)  # losses/intra_list_diversity_score.py:20-42
    EntropyDiversityScore(dataset=data)(interactions, recommendations)                                          # metrics/entropy_diversity_score.py:14-26
on the FIRST `--users` users of the workload: the scored dataset and the model's user table hold that prefix, while the
mask dataset (`frozen`) holds the WHOLE workload's interactions, so every user pays the reference's O(nnz) scan of the
full interaction table (base_datasets.py:139-147) exactly as a full run would.
D = 1 - En En^T (cosine distance over the model's item table) is built once outside the timed region: the reference
takes it as a constructor argument.  --check also compares the recommendations with the oracle restatement.
TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(1, os.path.dirname(HERE))

import torch  # noqa: E402

WORKLOADS = {"cfg2": (138493, 26744, 128, 100, 20000263), "cfg1": (943, 1682, 64, 10, 90570)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=200)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--budget", type=float, default=200.0, help="seconds the timed passes may take in total (the sample shrinks to fit)")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(HERE, "_ref", "divrec")):
        print(json.dumps({"unavailable": "oracle/_ref/divrec is missing (run oracle/make_ref.py where /root/reference exists)"}))
        return
    import divrec  # the reference
    assert os.path.realpath(divrec.__file__).startswith(os.path.realpath(os.path.join(HERE, "_ref"))), "not the reference copy"
    from divrec.datasets import RankingDataset, UserItemInteractionsDataset
    from divrec.losses import IntraListDiversityScore
    from divrec.metrics import EntropyDiversityScore
    from divrec.models import MatrixFactorization
    from divrec.train import get_model_recommendations
    from oracle import divrec_oracle as O  # input generators (+ the restatement for --check)

    cores = args.threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    nU, nI, d, k, nnz = WORKLOADS[args.workload]
    n = min(args.users, nU)
    gen = torch.Generator().manual_seed(0)
    U = O.make_embeddings(n, d, "gauss", gen)
    V = O.make_embeddings(nI, d, "gauss", gen)
    gen_s = torch.Generator().manual_seed(1)
    seen = O.make_seen(nU, nI, nnz, gen_s)  # the whole workload's mask
    data = UserItemInteractionsDataset(interactions=torch.tensor([[n - 1, nI - 1]]), number_of_users=n, number_of_items=nI)
    frozen = UserItemInteractionsDataset(interactions=seen, number_of_users=nU, number_of_items=nI)
    model = MatrixFactorization(n, nI, d)
    with torch.no_grad():
        model.user_embeddings.weight.copy_(U)
        model.item_embeddings.weight.copy_(V)
    model.eval()
    Vd = V.double()
    En = Vd / Vd.norm(dim=1, keepdim=True)
    D = (1.0 - En @ En.T).float()  # [I,I]: 2.9 GB at cfg2 (the reference's own interface)
    ild = IntraListDiversityScore(distance_matrix=D, reduction="none")
    ent = EntropyDiversityScore(dataset=data)
    interactions = data.interactions

    def one_step(n_step):
        """One timed pass of the reference over the first n_step users."""
        if n_step != n:
            d_ = UserItemInteractionsDataset(interactions=torch.tensor([[n_step - 1, nI - 1]]), number_of_users=n_step, number_of_items=nI)
            m_ = MatrixFactorization(n_step, nI, d)
            with torch.no_grad():
                m_.user_embeddings.weight.copy_(U[:n_step])
                m_.item_embeddings.weight.copy_(V)
            m_.eval()
            e_ = EntropyDiversityScore(dataset=d_)
        else:
            d_, m_, e_ = data, model, ent
        t0 = time.perf_counter()
        with torch.no_grad():
            r_ = get_model_recommendations(RankingDataset(d_, frozen=frozen), m_, k)
        t1 = time.perf_counter()
        iv = ild.recommendations_loss(d_.interactions, r_)
        t2 = time.perf_counter()
        en = e_(d_.interactions, r_)
        t3 = time.perf_counter()
        return r_, iv, en, (t1 - t0, t2 - t1, t3 - t2)

    # the first pass sizes the sample: the whole run (warm-up + steps) has to end within --budget seconds
    recs, ild_values, entropy, tt = one_step(n)
    total_steps = args.warmup + args.steps
    n_run = n
    if total_steps > 1 and sum(tt) * total_steps > args.budget:
        n_run = max(32, int(n * args.budget / (sum(tt) * total_steps)))
    times = [tt] if args.warmup == 0 else []
    done = 1
    while done < total_steps:
        r2, iv2, en2, tt2 = one_step(n_run)
        if done >= args.warmup:
            times.append(tt2)
        done += 1
    n_timed = n_run if total_steps > 1 else n
    if not times:
        times = [tt]
    sec = sorted(sum(t) for t in times)[len(times) // 2]
    rec_s = sorted(t[0] for t in times)[len(times) // 2]
    ild_s = sorted(t[1] for t in times)[len(times) // 2]
    out = {"users": n_timed, "workload": args.workload, "cores": cores, "seconds": sec, "users_per_s": n_timed / sec,
           "recommend_s": rec_s, "ild_s": ild_s, "steps": len(times), "warmup": args.warmup,
           "ild_mean": float(ild_values.mean()), "entropy": float(entropy), "torch": torch.__version__, "mask_nnz": int(seen.shape[0])}
    if args.check:
        rowptr, items = O.seen_csr(seen[seen[:, 0] < n], n)
        o_i, _ = O.recommendations_dense(U, V, rowptr, items, k)
        out["rows_equal_to_oracle"] = int((o_i == recs).all(dim=1).sum())
        out["ild_max_abs_diff_vs_oracle"] = float((O.ild_cosine(recs, V) - ild_values).abs().max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
