"""The N > 1 host logic (sharding, collectives, merge order) with 2 ranks over gloo on CPU.  The product has no
backend parameter: the workers below REPLACE the module attribute divrec.parallel._native with an oracle-backed
stand-in (tests may use the oracle); the collectives, shard ranges and result layout are the product code in
divrec/parallel.py.  The owner-sharded BPR route lives in a CUDA kernel over peer memory and is covered by the
2-GPU test in test_gpu_parity.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import divrec_oracle as O


class OracleKernels:
    """CPU stand-in with the signatures of divrec._native used by divrec.parallel."""

    @staticmethod
    def score_topk(U, V, k, seen_rowptr=None, seen_items=None, item_id_offset=0):
        rp, it = seen_rowptr, seen_items
        if rp is not None and it is not None:
            # keep only the seen items that fall into this shard, re-based to local ids
            rp = rp - rp[0]
            lo, hi = int(seen_rowptr[0]), int(seen_rowptr[-1])
            seg = it[lo:hi].long() - item_id_offset
            rows = torch.repeat_interleave(torch.arange(U.shape[0]), (rp[1:] - rp[:-1]))
            ok = (seg >= 0) & (seg < V.shape[0])
            cnt = torch.bincount(rows[ok], minlength=U.shape[0])
            rp = torch.zeros(U.shape[0] + 1, dtype=torch.int64)
            rp[1:] = torch.cumsum(cnt, 0)
            it = seg[ok].int()
        i, s = O.recommendations_dense(U, V, rp, it, k)
        return s, torch.where(i >= 0, i + item_id_offset, i)

    @staticmethod
    def score_topk_metrics(U, V, k, seen_rowptr=None, seen_items=None):
        rp = None if seen_rowptr is None else seen_rowptr - seen_rowptr[0]
        it = None if seen_items is None else seen_items[int(seen_rowptr[0]):int(seen_rowptr[-1])]
        i, s = O.recommendations_dense(U, V, rp, it, k)
        return s, i, O.ild_cosine(i, V), O.coverage_counts(i, V.shape[0]).int()

    @staticmethod
    def topk_merge(gs, gi):
        return O.topk_merge(gs, gi, gs.shape[2])

    MERGE_PACKED_MAX = 6144

    @staticmethod
    def topk_pack(s, i):
        return O.topk_pack(s, i)

    @staticmethod
    def topk_merge_packed(g):
        return O.topk_merge_packed(g)

    @staticmethod
    def bpr_fwd(U, V, u, p, n):
        pos, neg, loss, auc = O.bpr_forward(U, V, u, p, n)
        B = u.numel()
        return pos, neg, (loss * B).reshape(1), (auc * B).round().int().reshape(1)

    @staticmethod
    def bpr_step(U, V, u, p, n, scale, gU, gV, want_scores=True, loss_sum=None, auc_sum=None, item_rows=None):
        pos, neg, loss, auc = O.bpr_forward(U, V, u, p, n)
        B = u.numel()
        a, b = O.bpr_backward(U, V, u, p, n, upstream=scale * B)
        if gU is not None:
            gU += a
        if gV is not None:
            gV += b
        if item_rows is not None:   # row b = g_b U[u_b], g_b = -sigmoid(-(pos - neg)) * scale
            item_rows.copy_((-torch.sigmoid(-(pos - neg).double()) * scale)[:, None].float() * U[u])
        ls, au = (loss * B).reshape(1), (auc * B).round().int().reshape(1)
        if loss_sum is not None:    # caller-owned accumulators, as in the C ABI
            loss_sum += ls; ls = loss_sum
        if auc_sum is not None:
            auc_sum += au; au = auc_sum
        return pos, neg, ls, au

    @staticmethod
    def bpr_apply_rows(rows, p, n, gV):
        gV.index_add_(0, p, rows)
        gV.index_add_(0, n, -rows)

    @staticmethod
    def bpr_bwd(U, V, u, p, n, pos, neg, scale, gU, gV):
        a, b = O.bpr_backward(U, V, u, p, n, upstream=scale * u.numel())
        gU += a
        gV += b


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from divrec import parallel
    gen = torch.Generator().manual_seed(0)
    nU, nI, d, k = 101, 600, 32, 20
    U, V = O.make_embeddings(nU, d, "grid", gen), O.make_embeddings(nI, d, "grid", gen)
    seen = O.make_seen(nU, nI, 3000, gen)
    rowptr, items = O.seen_csr(seen, nU)
    full_i, full_s = O.recommendations_dense(U, V, rowptr, items, k)
    parallel._native = OracleKernels   # test-only substitution of the compute backend (no CUDA here)
    # users sharded
    it, sc, ild, ild_mean, counts, (a, b) = parallel.recommend_user_sharded(U, V, rowptr, items, k)
    assert (a, b) == parallel.shard_range(nU, rank, world)
    assert torch.equal(it, full_i[a:b]) and torch.equal(sc, full_s[a:b])
    assert torch.equal(counts.long(), O.coverage_counts(full_i, nI))          # all-reduced histogram = global
    assert abs(float(ild_mean) - float(O.ild_cosine(full_i, V).double().mean())) < 1e-6
    # items sharded
    ia, ib = parallel.shard_range(nI, rank, world)
    ms, mi = parallel.recommend_item_sharded(U, V[ia:ib].contiguous(), ia, k, rowptr, items)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)

    class TwoTensorKernels(OracleKernels):   # a backend without the packed exchange: two all-gathers + topk_merge
        MERGE_PACKED_MAX = 0
    parallel._native = TwoTensorKernels
    ms2, mi2 = parallel.recommend_item_sharded(U, V[ia:ib].contiguous(), ia, k, rowptr, items)
    assert torch.equal(mi2, full_i) and torch.equal(ms2, full_s)
    parallel._native = OracleKernels
    # BPR: triples partitioned by user shard
    B = 400
    u = torch.randint(0, nU, (B,), generator=gen); p = torch.randint(0, nI, (B,), generator=gen); n = torch.randint(0, nI, (B,), generator=gen)
    mine = (u >= a) & (u < b)
    Ug = O.make_embeddings(nU, d, "gauss", gen); Vg = O.make_embeddings(nI, d, "gauss", gen)
    loss, auc, gU, gV = parallel.bpr_step_once(Ug, Vg, u[mine], p[mine], n[mine], B)
    _, _, o_loss, o_auc = O.bpr_forward(Ug, Vg, u, p, n)
    o_gU, o_gV = O.bpr_backward(Ug, Vg, u, p, n)
    assert abs(float(loss) - float(o_loss)) < 1e-5 * abs(float(o_loss)) and abs(float(auc) - float(o_auc)) < 1e-6
    torch.testing.assert_close(gV, o_gV, rtol=0, atol=1e-6 * float(o_gV.abs().max()))
    torch.testing.assert_close(gU[a:b], o_gU[a:b], rtol=0, atol=1e-6 * float(o_gU.abs().max()))
    assert float(gU[:a].abs().sum()) == 0 and float(gU[b:].abs().sum()) == 0    # other shards' users untouched
    # sparse item-gradient exchange (rows + ids all-gathered, every rank scatter-adds all of them): equal batch sizes
    per_rank = []
    for r in range(world):
        ra, rb = parallel.shard_range(nU, r, world)
        per_rank.append(((u >= ra) & (u < rb)).nonzero().flatten())
    m = min(t.numel() for t in per_rank)
    sel = torch.cat([t[:m] for t in per_rank])
    me = per_rank[rank][:m]
    loss_s, auc_s, gU_s, gV_s = parallel.bpr_step_once(Ug, Vg, u[me], p[me], n[me], world * m, item_grad="sparse")
    _, _, o_loss, o_auc = O.bpr_forward(Ug, Vg, u[sel], p[sel], n[sel])
    o_gU, o_gV = O.bpr_backward(Ug, Vg, u[sel], p[sel], n[sel])
    assert abs(float(loss_s) - float(o_loss)) < 1e-5 * abs(float(o_loss)) and abs(float(auc_s) - float(o_auc)) < 1e-6
    torch.testing.assert_close(gV_s, o_gV, rtol=0, atol=1e-6 * float(o_gV.abs().max()))
    torch.testing.assert_close(gU_s[a:b], o_gU[a:b], rtol=0, atol=1e-6 * float(o_gU.abs().max()))
    loss_d, _, _, gV_d = parallel.bpr_step_once(Ug, Vg, u[me], p[me], n[me], world * m, item_grad="dense")
    torch.testing.assert_close(gV_d, gV_s, rtol=0, atol=1e-6 * float(o_gV.abs().max()))
    # "auto": equal batches and a small volume -> sparse; unequal batches (the first call's split) -> dense, no mismatch
    _, _, _, gV_a = parallel.bpr_step_once(Ug, Vg, u[me], p[me], n[me], world * m, item_grad="auto")
    torch.testing.assert_close(gV_a, gV_s, rtol=0, atol=1e-6 * float(o_gV.abs().max()))
    _, _, _, gV_u = parallel.bpr_step_once(Ug, Vg, u[mine], p[mine], n[mine], B, item_grad="auto")
    torch.testing.assert_close(gV_u, gV, rtol=0, atol=1e-6 * float(gV.abs().max()))
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_two_rank_sharding_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    assert sorted(out.get(timeout=5) for _ in range(world)) == [0, 1]


def test_shard_range_covers_everything():
    from divrec.parallel import shard_range
    for n in (0, 1, 7, 100, 138493):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_bind_host_to_gpu_is_harmless_without_topology(tmp_path, monkeypatch):
    """bind_host_to_gpu: no CUDA device / no sysfs / one NUMA node -> nothing changes; with a (faked) two-node topology the
    process is pinned to the GPU's node and unbind_host restores the previous affinity."""
    import os
    from types import SimpleNamespace
    from divrec import parallel
    before = os.sched_getaffinity(0)
    info = parallel.bind_host_to_gpu(0, sysfs=str(tmp_path))
    assert info["node"] is None and os.sched_getaffinity(0) == before
    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    if len(before) < 2:
        return
    cpus = sorted(before)
    half = cpus[: len(cpus) // 2]
    monkeypatch.setattr(parallel.torch.cuda, "get_device_properties",
                        lambda i: SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0))
    d = tmp_path / "bus/pci/devices/0000:1b:00.0"; d.mkdir(parents=True)
    (d / "numa_node").write_text("1\n")
    n = tmp_path / "devices/system/node/node1"; n.mkdir(parents=True)
    (n / "cpulist").write_text(",".join(str(c) for c in half) + "\n")
    try:
        info = parallel.bind_host_to_gpu(0, sysfs=str(tmp_path))
        assert info["node"] == 1 and info["cpus"] == len(half) and os.sched_getaffinity(0) == set(half)
    finally:
        parallel.unbind_host(info)
    assert os.sched_getaffinity(0) == before
