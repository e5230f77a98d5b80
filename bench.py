#!/usr/bin/env python
"""bench.py -- the headline measurement of the divrec hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1]

Metric (BASELINE.json): users/sec scored (score + seen mask + top-k + ILD + coverage).  A "step" is one
pass of the hot path over one batch of synthetic users: BASELINE.json configs[1], the MovieLens-20M
shape (138 493 users x 26 744 items, d = 128, k = 100, 20 000 263 seen interactions).  At N > 1 every
rank scores its own batch of that size against the replicated item table (weak scaling, users
sharded); the only collectives are the metric reductions (coverage histogram + ILD sum all-reduce).

One JSON line on stdout (rank 0).  `value` = users/s with inputs resident in HBM; `e2e` = the same
through the reference-facing Python API with HOST buffers (pinned), host<->device copies inside the
timed region; `roofline` = the dominant kernel by time (selection: mask + exact top-k + ILD gather) against the
measured HBM bandwidth, `roofline_filter` = the tcgen05 score + filter kernel against the measured bf16 tensor peak;
`cpu_baseline` = the CPU oracle port timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU algorithm (oracle port, all host threads) on a bounded
sample of the same workload; under torchrun only rank 0 works.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diversity-recommendations_b200"))

WORKLOADS = {
    # name: (users, items, d, k, nnz seen)
    "cfg2": (138493, 26744, 128, 100, 20000263),
    "cfg1": (943, 1682, 64, 10, 90570),
}
METRIC = "users/sec scored (score+top-k+ILD)"
# dram__bytes_read.sum + dram__bytes_write.sum per launch at cfg2, from the committed ncu --set full capture
# (profiles/r01_ncu_filter_select_v6.txt, taken before the masked hint: ~500 instead of ~360 captured elements per row,
# so the current build moves somewhat less); bench.py cannot run ncu itself
NCU_DRAM_TRAFFIC_CFG2 = {"score_filter2_kernel": 0.109167e9 + 2.306356e9, "select_topk_kernel": 2.518991e9 + 0.165455e9}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def make_inputs(name, seed):
    """Synthetic inputs on the CPU: N(0,1) fp32 embeddings = nn.Embedding's default init,
    uniform unique (user,item) seen pairs as CSR."""
    from oracle import divrec_oracle as O  # input generators only (shared with the tests)
    nU, nI, d, k, nnz = WORKLOADS[name]
    gen = torch.Generator().manual_seed(seed)
    U = O.make_embeddings(nU, d, "gauss", gen)
    V = O.make_embeddings(nI, d, "gauss", gen)
    seen = O.make_seen(nU, nI, nnz, gen)
    rowptr, items = O.seen_csr(seen, nU)
    return U, V, rowptr, items, k


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # "under load": the upper half of the samples (idle samples before/after the region pull the median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_step_device(N, model_U, V, rowptr, items, k, nI, counts):
    """One pass with everything resident: score + mask + top-k with the cosine ILD and the coverage
    histogram fused into the selection kernel (divrec_score_topk_metrics)."""
    counts.zero_()
    s, i, ild, _ = N.score_topk_metrics(model_U, V, k, rowptr, items, counts=counts)
    return s, i, ild


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout, but C libraries write there too (NCCL prints its version banner on the
    first communicator whatever NCCL_DEBUG says on some boxes).  From here on file descriptor 1 points at stderr; the
    JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        _JSON_OUT = os.fdopen(saved, "w")


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _quiet_nccl():
    """NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/INFO; the contract is ONE JSON line."""
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("DIVREC_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"


def bench_ours(args):
    _quiet_nccl()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from divrec import _native as N
    N.load_library()

    nU, nI, d, k, nnz = WORKLOADS[args.workload]
    U, V, rowptr, items, k = make_inputs(args.workload, seed=rank)
    Uh, Vh, rph, ith = (t.pin_memory() for t in (U, V, rowptr, items))
    Ud, Vd, rpd, itd = (t.to(dev) for t in (U, V, rowptr, items))
    counts = torch.zeros(nI, dtype=torch.int32, device=dev)
    ild_sum = torch.zeros(1, dtype=torch.float64, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        s, i, ild = run_step_device(N, Ud, Vd, rpd, itd, k, nI, counts)
        ild_sum.copy_(ild.sum(dtype=torch.float64).reshape(1))
        if dist is not None:  # the path's only exchange: the metric reductions (as divrec.parallel.recommend_user_sharded)
            dist.all_reduce(counts)
            dist.all_reduce(ild_sum)
        return s, i, ild

    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    # ---- device-resident timing ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    N.set_profiling(True)
    launches0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    filter_ms = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
        filter_ms.append(None)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = N.launch_count() - launches0
    # per-kernel time of the dominant kernel: events recorded by the library around the score+filter
    # launch, on the launching stream; read back per step in a second short loop (same inputs)
    kern, phases = [], []
    for _ in range(min(args.steps, 10)):
        step()
        torch.cuda.synchronize()
        phases.append(N.profile_read())
        kern.append(phases[-1]["score_filter_ms"])
    _s, _i, stats = N.score_topk(Ud, Vd, k, rpd, itd, return_stats=True)
    N.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / args.steps
    value = nU * world / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers ----
    from divrec.train import host_recommend_and_score
    for _ in range(2):
        host_recommend_and_score(Uh, Vh, rph, ith, k, dev, chunks=args.e2e_chunks)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        res = host_recommend_and_score(Uh, Vh, rph, ith, k, dev, chunks=args.e2e_chunks)
    t1.record()
    barrier()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = nU * world / (float(e2e_ms) / args.steps * 1e-3)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    kms = sorted(kern)[len(kern) // 2]
    flop = 2.0 * nU * nI * d
    achieved = flop / (kms * 1e-3) / 1e12
    peak = pk["bf16_tflops_sustained"]
    cpu = cpu_baseline(args.workload, sample_users=args.cpu_sample) if world == 1 and not args.no_cpu else None
    line = {
        "metric": METRIC, "value": value, "unit": "users/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16 tensor-core filter + exact fp32 rescoring, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"{args.workload}: MovieLens-20M shape" if args.workload == "cfg2" else args.workload,
                   "users_per_gpu": nU, "items": nI, "d": d, "k": k, "seen_nnz": int(items.numel()),
                   "stages": "score+mask+top-k, ILD (cosine), coverage histogram",
                   "sharding": "users sharded, item table replicated" if world > 1 else "single GPU",
                   "l2": "inputs + candidate workspace (>1 GB) exceed the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "users/s", "h2d_bytes_per_step": res["h2d_bytes"], "d2h_bytes_per_step": res["d2h_bytes"],
                "api": "divrec.train.host_recommend_and_score (pinned host buffers -> C ABI -> host results)"},
        "gpu_launches": int(launches),
        # the dominant kernel by time (ncu launch list profiles/r01_ncu_launches_v7.txt: 51 % of the launches) is the selection
        # kernel (mask + exact top-k + ILD gather + coverage): HBM/gather-bound stage S2+S3 of SURVEY.md 8d
        "roofline": select_roofline(phases, nU, d, k, pk, args.workload, float(items.numel()) / nU, ms_per_step),
        # the tcgen05 score + filter kernel (35 % of the launches) against the measured tensor peak
        "roofline_filter": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            "traffic": NCU_DRAM_TRAFFIC_CFG2["score_filter2_kernel"] if args.workload == "cfg2" else None,
                            "kernel": "score_filter2_kernel" if d <= 128 else "score_filter_kernel", "kernel_ms": kms,
                            "share_of_step": kms / ms_per_step, "algorithmic_flop_per_launch": flop,
                            "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)"},
        "cpu_baseline": cpu,
        "phases_ms": {name: sorted(ph[name] for ph in phases)[len(phases) // 2] for name in phases[0]},
        "rescued_rows_per_step": stats["rescued_rows"],
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def bench_cfg4(args):
    """BASELINE.json configs[3]: 10M-item x d=256 bf16 item table sharded over the GPUs, 65 536 query users per step,
    per-shard score + top-100, NCCL all-gather of the per-shard lists, n-way merge.  Strong scaling: the
    catalogue is fixed, every rank holds 1/N of it.  Synthetic N(0,1) bf16 tables generated on the device."""
    _quiet_nccl()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from divrec import _native as N, parallel
    N.load_library()
    nI_total, d, k, nQ = args.cfg4_items, 256, 100, args.cfg4_queries
    a, b = parallel.shard_range(nI_total, rank, world)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    V = torch.empty((b - a, d), dtype=torch.bfloat16, device=dev)
    for i in range(0, b - a, 1 << 20):
        V[i:i + (1 << 20)] = torch.randn((min(1 << 20, b - a - i), d), generator=g, device=dev).to(torch.bfloat16)
    gq = torch.Generator().manual_seed(7)
    Uh = torch.randn((nQ, d), generator=gq).to(torch.bfloat16).pin_memory()
    Ud = Uh.to(dev)
    out_s = torch.empty((nQ, k), dtype=torch.float32).pin_memory()
    out_i = torch.empty((nQ, k), dtype=torch.int64).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(U):
        return parallel.recommend_item_sharded(U, V, a, k)

    for _ in range(max(1, min(args.warmup, 2))):
        step(Ud)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    N.set_profiling(True)
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        s, i = step(Ud)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = N.launch_count() - l0
    phases = N.profile_read()
    _s, _i, stats = N.score_topk(Ud, V, k, item_id_offset=a, return_stats=True)
    N.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    # end to end: queries from pinned host memory, results back to pinned host memory
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        Ud.copy_(Uh, non_blocking=True)
        s, i = step(Ud)
        out_s.copy_(s, non_blocking=True)
        out_i.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t1.record()
    barrier()
    e2e = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    if rank == 0:
        pk = peaks()
        flop = 2.0 * nQ * (b - a) * d
        kms = phases["score_filter_ms"]
        ach = flop / (kms * 1e-3) / 1e12
        line = {"metric": METRIC.replace("+ILD", ""), "value": nQ / (ms_per_step * 1e-3), "unit": "users/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16 tables (tensor-core filter, exact fp32 rescoring of the candidates)", "data": "synthetic",
                "config": {"workload": "cfg4: item table sharded, per-shard top-k + all-gather + merge", "items_total": nI_total,
                           "items_per_gpu": b - a, "d": d, "k": k, "queries_per_step": nQ, "l2": "item shard exceeds the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": nQ / (float(e2e) / args.steps * 1e-3), "unit": "users/s", "h2d_bytes_per_step": Uh.numel() * 2,
                        "d2h_bytes_per_step": nQ * k * 12, "api": "divrec.parallel.recommend_item_sharded (queries from pinned host memory)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": ach / pk["bf16_tflops_sustained"], "traffic": None, "kernel": "score_filter_kernel", "kernel_ms": kms,
                             "algorithmic_flop_per_launch": flop, "peak_source": pk["source"] + " bf16 sustained"},
                "cpu_baseline": None, "phases_ms": phases, "rescued_rows_per_step": stats["rescued_rows"]}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def bench_cfg3(args):
    """BASELINE.json configs[2]: BPR training step, 1M users x 1M items, d = 128 fp32, B = 65 536 triples per GPU:
    fused gather + loss + AUC forward and scatter-add backward into dense gradient tables (the reference's
    autograd semantics).  Users sharded: each rank draws its triples' users from its own range; the item-table
    gradient is all-reduced.  Gradient buffers persist across steps (zero_grad belongs to the optimiser: out of scope)."""
    _quiet_nccl()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from divrec import _native as N, parallel
    N.load_library()
    nU = nI = 1_000_000
    d, B = 128, 65536
    g = torch.Generator(device=dev).manual_seed(5)
    U = torch.randn((nU, d), generator=g, device=dev)
    V = torch.randn((nI, d), generator=g, device=dev)
    a, b = parallel.shard_range(nU, rank, world)
    gh = torch.Generator().manual_seed(100 + rank)
    uh = torch.randint(a, b, (B,), generator=gh).pin_memory()
    ph = torch.randint(0, nI, (B,), generator=gh).pin_memory()
    nh = torch.randint(0, nI, (B,), generator=gh).pin_memory()
    u, p_, n_ = uh.to(dev), ph.to(dev), nh.to(dev)
    gU, gV = torch.zeros_like(U), torch.zeros_like(V)
    Bg = B * world

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
    auc_buf = torch.zeros(1, dtype=torch.int32, device=dev)

    # item-gradient exchange: the dense all-reduce BASELINE.json names, or the sparse form (all-gather of the per-triple
    # rows g_b U[u_b] + ids, every rank scatter-adds all of them: the same dense gV on every rank) when it moves less
    sparse = world > 1 and (args.cfg3_item_grad == "sparse" or
                            (args.cfg3_item_grad == "auto" and world * B * (d + 4) * 4 <= nI * d * 4))
    if sparse:
        rows = torch.empty((B, d), dtype=torch.float32, device=dev)
        all_rows = torch.empty((world * B, d), dtype=torch.float32, device=dev)
        ids = torch.empty((2, B), dtype=torch.int64, device=dev)
        all_ids = torch.empty((world * 2, B), dtype=torch.int64, device=dev)

    def step(uu, pp, nn):
        # forward + backward in ONE kernel (divrec_bpr_step): the rows of a triple are read once
        loss_buf.zero_(); auc_buf.zero_()
        if sparse:
            N.bpr_step(U, V, uu, pp, nn, 1.0 / Bg, gU, None, want_scores=False, loss_sum=loss_buf, auc_sum=auc_buf, item_rows=rows)
            ids[0].copy_(pp); ids[1].copy_(nn)
            dist.all_gather_into_tensor(all_rows, rows)
            dist.all_gather_into_tensor(all_ids, ids)
            g3 = all_ids.view(world, 2, B)
            N.bpr_apply_rows(all_rows, g3[:, 0].reshape(-1), g3[:, 1].reshape(-1), gV)
        else:
            N.bpr_step(U, V, uu, pp, nn, 1.0 / Bg, gU, gV, want_scores=False, loss_sum=loss_buf, auc_sum=auc_buf)
            if dist is not None:
                dist.all_reduce(gV)
        if dist is not None:
            dist.all_reduce(loss_buf)
        return loss_buf

    for _ in range(max(3, args.warmup)):
        step(u, p_, n_)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(u, p_, n_)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = N.launch_count() - l0
    # kernels alone (no collective) for the roofline
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        N.bpr_step(U, V, u, p_, n_, 1.0 / Bg, gU, gV, want_scores=False, loss_sum=loss_buf, auc_sum=auc_buf)
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    loss_h = torch.empty(1, dtype=torch.float32).pin_memory()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        u.copy_(uh, non_blocking=True); p_.copy_(ph, non_blocking=True); n_.copy_(nh, non_blocking=True)
        ls = step(u, p_, n_)
        loss_h.copy_(ls, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t1.record()
    barrier()
    e2e = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    if rank == 0:
        pk = peaks()
        ms_per_step = float(ms) / args.steps
        bytes_per_triple = 3 * d * 4 + 3 * 2 * d * 4 + 24   # 3 row reads + 3 row read-modify-writes + ids
        ach = B * bytes_per_triple / (kms * 1e-3) / 1e9
        line = {"metric": "BPR triples/sec (gather + loss + scatter-add backward)", "value": Bg / (ms_per_step * 1e-3), "unit": "triples/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg3: BPR step, 1M users x 1M items", "d": d, "triples_per_gpu": B,
                           "sharding": ("triples partitioned by user shard; item gradient: " +
                                        ("all-gather of per-triple rows + scatter-add on every rank (sparse form of the all-reduce)" if sparse
                                         else "dense all-reduce")) if world > 1 else "single GPU",
                           "l2": "tables + dense gradients (2 GB) exceed the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": Bg / (float(e2e) / args.steps * 1e-3), "unit": "triples/s", "h2d_bytes_per_step": 3 * B * 8,
                        "d2h_bytes_per_step": 4, "api": "divrec._native.bpr_step (triples from pinned host memory, loss back)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                             "traffic": (194.578176e6 + 47.259392e6) if B == 65536 else None,  # ncu --set full, profiles/r01_ncu_bpr_step.txt
                             "kernel": "bpr_step_kernel", "kernel_ms": kms,
                             "algorithmic_bytes_per_launch": B * bytes_per_triple, "peak_source": pk["source"] + " copy bandwidth"},
                "cpu_baseline": None}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def bench_cfg5(args):
    """BASELINE.json configs[4]: greedy MMR re-rank of each user's top-500 candidates down to k = 100 (d = 128,
    lambda = 0.5), 1M users in total sharded over the GPUs with no collective (strong scaling).  The candidate lists
    (top-500 by relevance, sorted) are synthetic INPUT of this path, generated untimed with torch.topk over U V^T."""
    _quiet_nccl()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from divrec import _native as N, parallel
    N.load_library()
    nU_total, nI, d, C, k, lam = args.cfg5_users, 26744, 128, args.cfg5_candidates, 100, 0.5
    ua, ub = parallel.shard_range(nU_total, rank, world)
    nU = ub - ua
    g = torch.Generator(device=dev).manual_seed(9)
    V = torch.randn((nI, d), generator=g, device=dev)
    g.manual_seed(100 + rank)
    cand = torch.empty((nU, C), dtype=torch.int64, device=dev)
    sc = torch.empty((nU, C), dtype=torch.float32, device=dev)
    for a in range(0, nU, 8192):
        Uc = torch.randn((min(8192, nU - a), d), generator=g, device=dev)
        sc[a:a + 8192], cand[a:a + 8192] = torch.topk(Uc @ V.T, C, dim=1, sorted=True)
    del Uc
    ch, sh = cand.cpu().pin_memory(), sc.cpu().pin_memory()
    out_h = torch.empty((nU, k), dtype=torch.int64).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        N.mmr_rerank(V, cand, sc, k, lam)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        items, rel = N.mmr_rerank(V, cand, sc, k, lam)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = N.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        cand.copy_(ch, non_blocking=True); sc.copy_(sh, non_blocking=True)
        items, rel = N.mmr_rerank(V, cand, sc, k, lam)
        out_h.copy_(items, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t1.record()
    barrier()
    e2e = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    if rank == 0:
        pk = peaks()
        ms_per_step = float(ms) / args.steps
        algo = nU * (C * d * 4 + C * 12 + k * 8)  # gather C rows + candidate ids/scores in + k ids out
        ach = algo / (ms_per_step * 1e-3) / 1e9
        line = {"metric": "users/sec re-ranked (greedy MMR)", "value": nU_total / (ms_per_step * 1e-3), "unit": "users/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg5: greedy MMR re-rank of top-C candidates", "users_total": nU_total, "users_per_gpu": nU, "candidates": C,
                           "k": k, "lambda": lam, "d": d, "candidates_from": "torch.topk(U V^T, C), N(0,1) embeddings (untimed input generation)",
                           "l2": "candidate lists + scratch exceed the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": nU_total / (float(e2e) / args.steps * 1e-3), "unit": "users/s", "h2d_bytes_per_step": nU * C * 12,
                        "d2h_bytes_per_step": nU * k * 8, "api": "divrec.rerank.mmr_rerank (candidates from pinned host memory)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                             "kernel": "mmr_kernel", "kernel_ms": ms_per_step, "algorithmic_bytes_per_launch": algo,
                             "note": "lazy greedy: only candidates whose upper bound can still win are gathered and evaluated, so the kernel "
                                     "moves fewer bytes than the algorithmic figure; the rest is per-step latency (two block barriers per pick)"},
                "cpu_baseline": None}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def select_roofline(phases, nU, d, k, pk, workload, avg_seen, ms_per_step):
    """HBM view of select_topk_kernel with the ALGORITHMIC bytes of its stage (SURVEY.md 8d, S2 + S3 and the top-k side
    traffic of S1): per user the k item rows gathered for the exact scores and the cosine ILD (k (8 + 4 d) + 4), the
    top-k outputs (12 k), the seen list (4 seen_u + 8).  The candidate groups handed over by the filter are this
    implementation's own intermediate data and are NOT counted; `traffic` is the measured DRAM traffic (the 13.7 MB item
    table is L2-resident, so the row gathers are mostly L2 hits)."""
    ms = sorted(ph["select_ms"] for ph in phases)[len(phases) // 2]
    algo = nU * (k * (8 + 4 * d) + 4 + 12 * k + 4 * avg_seen + 8)
    ach = algo / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "traffic": NCU_DRAM_TRAFFIC_CFG2["select_topk_kernel"] if workload == "cfg2" else None,
            "kernel": "select_topk_kernel", "kernel_ms": ms, "share_of_step": ms / ms_per_step,
            "algorithmic_bytes_per_launch": algo, "peak_source": pk["source"] + " copy bandwidth",
            "note": "issue-bound rather than bandwidth-bound (ncu: issue slots 60 % busy, 2.36 inst/cycle/SM; row gathers hit L2)"}


def cpu_baseline(workload, sample_users):
    """The oracle port (dense torch-CPU restatement of the reference loop) on this box's host cores."""
    from oracle import divrec_oracle as O
    nU, nI, d, k, nnz = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(sample_users, nU)
    gen = torch.Generator().manual_seed(0)
    U = O.make_embeddings(sample, d, "gauss", gen)
    V = O.make_embeddings(nI, d, "gauss", gen)
    seen = O.make_seen(sample, nI, int(nnz * sample / nU), gen)
    rowptr, items = O.seen_csr(seen, sample)
    t0 = time.perf_counter()
    recs, _ = O.recommendations_dense(U, V, rowptr, items, k, exact_scores=False)
    ild = O.ild_cosine(recs, V)
    O.coverage_counts(recs, nI)
    dt = time.perf_counter() - t0
    # the literal per-user reference loop (Python set difference + argsort + combinations ILD) on a few users
    lit_users = 8 if nI > 5000 else 64
    t1 = time.perf_counter()
    r2, _ = O.recommendations_loop(U, V, seen, k, range(lit_users))
    D_free = O.ild_cosine(r2, V)
    lit = lit_users / (time.perf_counter() - t1)
    del ild, D_free
    return {"value": sample / dt, "unit": "users/s", "cores": cores, "kind": "port",
            "sample": f"{sample} users of {workload} (dense torch-CPU oracle: GEMM + -inf mask + stable sort + Gram ILD + bincount), {dt:.1f} s",
            "literal_reference_loop_users_per_s": lit}


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nU, nI, d, k, nnz = WORKLOADS[args.workload]
    vals = []
    for s in range(max(1, args.warmup) + args.steps):
        c = cpu_baseline(args.workload, sample_users=args.cpu_sample)
        if s >= max(1, args.warmup):
            vals.append(c)
        if len(vals) >= 3:  # bounded: the whole arm stays within minutes
            break
    v = sum(c["value"] for c in vals) / len(vals)
    c = dict(vals[-1]); c["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "users/s", "n_gpus": world, "steps": len(vals),
            "warmup": max(1, args.warmup), "ms_per_step": args.cpu_sample / v * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: MovieLens-20M shape" if args.workload == "cfg2" else args.workload,
                       "users_per_step": min(args.cpu_sample, nU), "items": nI, "d": d, "k": k},
            "cpu_baseline": c,
            "e2e": {"value": v, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--cfg5-users", type=int, default=1_000_000)
    ap.add_argument("--cfg5-candidates", type=int, default=500)
    ap.add_argument("--cfg3-item-grad", default="auto", choices=["auto", "dense", "sparse"])
    ap.add_argument("--cfg4-items", type=int, default=10_000_000)
    ap.add_argument("--cfg4-queries", type=int, default=65536)
    ap.add_argument("--cpu-sample", type=int, default=32768)
    ap.add_argument("--e2e-chunks", type=int, default=4, help="user ranges the host-buffer pipeline overlaps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-baseline leg (profiling runs)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        bench_reference(args)
    else:
        if not torch.cuda.is_available():
            sys.exit("bench.py: no CUDA device -- divrec_b200 has no CPU path (use --impl reference for the CPU oracle)")
        if args.workload == "cfg4":
            bench_cfg4(args)
        elif args.workload == "cfg3":
            bench_cfg3(args)
        elif args.workload == "cfg5":
            bench_cfg5(args)
        else:
            bench_ours(args)


if __name__ == "__main__":
    main()
