#!/usr/bin/env python
"""bench.py -- the headline measurement of the divrec hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|cfg1|cfg2|cfg3|cfg4|cfg5]

Metric (BASELINE.json): users/sec scored (score + seen mask + top-k + ILD + coverage).  A "step" is one
pass of the hot path over one batch of synthetic users: BASELINE.json configs[1], the MovieLens-20M
shape (138 493 users x 26 744 items, d = 128, k = 100, 20 000 263 seen interactions).  At N > 1 every
rank scores its own batch of that size against the replicated item table (weak scaling, users
sharded); the only collectives are the metric reductions (coverage histogram + ILD sum all-reduce).

One JSON line on stdout (rank 0).  `value` = users/s with inputs resident in HBM; `e2e` = the same
through the reference-facing Python API with HOST buffers (pinned), host<->device copies inside the
timed region; `roofline` = the dominant kernel by time; `roofline_filter` = the tcgen05 score + filter kernel against the
measured bf16 tensor peak; `roofline_step` = the whole fused step against SURVEY.md 8(d)'s serial bound (tensor-core
contraction + ILD gather + coverage); `cpu_baseline` = the CPU oracle port timed on this box's host cores on a bounded
sample, with the literal reference loop beside it.

With the default `--workload all` the same line also carries the sharded workloads BASELINE.json names, each measured
at the same N in the same process, as sub-records: `cfg4` (10M x 256 bf16 item table sharded over the ranks: per-shard
score + top-100, all-gather, merge -- strong scaling), `cfg3` (BPR step at 1M x 1M, users sharded, item gradient
exchanged -- weak scaling, two batch sizes) and `cfg5` (greedy MMR re-rank of top-500 candidate lists that
divrec_score_topk(k = 500) produced).  `--workload cfgN` prints that workload's own line instead.

Timing: W (>= 3) untimed steps, then exactly K steps between two CUDA events, bracketed by barrier + synchronize, max over
ranks; SM clock and clock-event reasons are read through NVML every 50 ms from before the warm-up to after the bracket.  Per-step
events are kept as a diagnostic (`step_ms_rank0`); if one step of a bracket was held up for many step times by something
outside the process (seen on this pool with and without the sampler: median step unchanged, one step of 10-200 ms), the
bracket is measured once more and the line carries `remeasured` with the first bracket's figures.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/divrec, copied by oracle/make_ref.py) on the box's host
cores: get_model_recommendations + IntraListDiversityScore + EntropyDiversityScore on a user prefix of the same
workload, the same number of steps; the dense torch-CPU port is reported beside it.  Under torchrun only rank 0 works.
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this build;
# written by tools/ncu_traffic.py next to the text summary it was taken from (bench.py cannot run ncu itself)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def ncu_traffic(kernel, workload):
    """(bytes per launch or None, source string) of `kernel` at `workload` from the committed ncu capture."""
    try:
        t = json.load(open(TRAFFIC_FILE))
        e = t["kernels"][workload][kernel]
        return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), t.get("source")
    except (OSError, KeyError, ValueError):
        return None, None


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
    """SM clock and clock-event (throttle) reasons DURING the timed region: the B200_PROFILING.md clocks line, read through
    NVML in-process (nvidia_ml_py, the library nvidia-smi itself calls) every 50 ms.  A looping `nvidia-smi --query-gpu`
    process was measured to stall kernel launches: its start-up cost 250 ms inside a 4-step region, and on an 8-GPU box
    one periodic query inside the region took the cfg2 step from 3.0 to 7.3 ms -- through the per-step all-reduce, on
    every rank.  nvidia-smi remains the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, active=True, period=0.05):
        self.gpu = gpu_index
        self.active = active and not os.environ.get("DIVREC_BENCH_NO_CLOCKS")  # (diagnostic switch: is a stall the sampler's?)
        self.period = period
        self.f = None
        self.p = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()
        self.source = None

    def _nvml_loop(self, nv, h):
        bits = []
        for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            if hasattr(nv, attr):
                bits.append((name, getattr(nv, attr)))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = get_reasons(h)
                for name, bit in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if not self.active:
            return self
        try:
            import pynvml as nv
            import threading
            nv.nvmlInit()
            pr = torch.cuda.get_device_properties(self.gpu)  # NVML counts physical devices: find ours by its bus id
            h = nv.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            # first calls of the two queries here, outside the timed region (NVML initialises lazily per query type)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            (getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons)(h)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.source = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi"
        except OSError:
            self.p = None
        # nvidia-smi's start-up (NVML initialisation) stalls kernel launches: wait for its first sample before timing
        t0 = time.time()
        while self.p is not None and self.p.poll() is None and os.path.getsize(self.f.name) == 0 and time.time() - t0 < 10.0:
            time.sleep(0.05)
        return self

    def stop(self):
        if not self.active:
            return None
        if self.thread is not None:
            time.sleep(2 * self.period)
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        elif self.p is not None:
            time.sleep(0.25)
            self.p.terminate()
            self.p.wait()
            self.f.flush()
            rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
            os.unlink(self.f.name)
            for r in rows:
                try:
                    self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML binding and no nvidia-smi"]}
        sm = sorted(self.sm)
        # "under load": the upper half of the samples (idle samples before/after the region pull the median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": self.source}


class Ctx:
    """One process per GPU: rank / device / process group, set up once and shared by every workload of the run."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        # host placement: this rank's pinned buffers and copy threads on the NUMA node of its GPU (e2e at N > 2 staged half of
        # the ranks through the other socket); undone for the CPU-baseline leg, which should see every core
        from divrec.parallel import bind_host_to_gpu
        self.numa = bind_host_to_gpu(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        from divrec import _native as N
        N.load_library()
        self.N = N

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def timed(self, fn, steps, retry_on_stall=True):
        """barrier + sync, exactly `steps` calls between two CUDA events on the current stream, barrier + sync; max over ranks.
        Per-step events are a diagnostic only.  One thing they are used for: on this pool a single step of a bracket is now and
        then held up for 10-200 ms by something outside the process (the median step is unchanged; it happens with the clock
        sampler off as well) -- if any step of any rank took more than 5x its median + 5 ms, the bracket is measured once more,
        as the contract does for a run that saw a thermal event, and the line says so (`remeasured`)."""
        self.last_remeasured = None
        for attempt in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            self.barrier()
            e0.record()
            out = None
            for i in range(steps):
                out = fn()
                marks[i].record()
            e1.record()
            self.barrier()
            per = sorted((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(steps))
            med, mx = per[len(per) // 2], per[-1]
            self.last_steps_ms = {"median": med, "max": mx}
            ms = self.max_over_ranks(e0.elapsed_time(e1)) / steps
            stalled = self.max_over_ranks(1.0 if (steps >= 4 and mx > 5.0 * med + 5.0) else 0.0) > 0.0
            if not (stalled and retry_on_stall and attempt == 0):
                return ms, out
            self.last_remeasured = {"reason": "a step of the first bracket was held up outside the kernels", "first_bracket_ms_per_step": ms,
                                    "first_bracket_step_ms": dict(self.last_steps_ms)}
        return ms, out

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout, but C libraries write there too (NCCL prints its version banner and, with
    NCCL_DEBUG=INFO, its log).  From here on file descriptor 1 points at stderr -- NCCL's output stays visible there,
    untouched -- and the JSON line goes to a private duplicate of the original stdout."""
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


# ------------------------------------------------------------------------------------------------------------------
# cfg2 / cfg1: score + mask + top-k + ILD + coverage, users sharded (weak scaling)
# ------------------------------------------------------------------------------------------------------------------


def bench_eval(ctx, args, workload):
    N, dev, dist, world, rank = ctx.N, ctx.dev, ctx.dist, ctx.world, ctx.rank
    nU, nI, d, k, nnz = WORKLOADS[workload]
    U, V, rowptr, items, k = make_inputs(workload, seed=rank)
    Uh, Vh, rph, ith = (t.pin_memory() for t in (U, V, rowptr, items))
    Ud, Vd, rpd, itd = (t.to(dev) for t in (U, V, rowptr, items))
    counts = torch.zeros(nI, dtype=torch.int32, device=dev)
    ild_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    warmup = max(3, args.warmup)

    def step():
        # everything resident: score + mask + top-k with the cosine ILD and the coverage histogram fused into the selection
        # kernel (divrec_score_topk_metrics), then the path's only exchange -- the metric reductions
        counts.zero_()
        s, i, ild, _ = N.score_topk_metrics(Ud, Vd, k, rpd, itd, counts=counts)
        ild_sum.copy_(ild.sum(dtype=torch.float64).reshape(1))
        if dist is not None:
            dist.all_reduce(counts)
            dist.all_reduce(ild_sum)
        return s, i, ild

    sampler = ClockSampler(ctx.local_rank, active=rank == 0).start()  # before the warm-up: no idle gap in front of the timed region
    for _ in range(warmup):
        step()
    ctx.barrier()
    launches0 = N.launch_count()
    ms_per_step, _ = ctx.timed(step, args.steps)
    step_spread = dict(ctx.last_steps_ms)
    remeasured = ctx.last_remeasured
    launches = N.launch_count() - launches0
    # per-kernel durations: CUDA events recorded by the library around its launches, on the launching stream
    N.set_profiling(True)
    phases = []
    for _ in range(min(args.steps, 10)):
        step()
        torch.cuda.synchronize()
        phases.append(N.profile_read())
    _s, _i, stats = N.score_topk(Ud, Vd, k, rpd, itd, return_stats=True)
    N.set_profiling(False)
    clocks = sampler.stop()
    value = nU * world / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers ----
    from divrec.train import host_recommend_and_score
    res = None
    for _ in range(2):
        res = host_recommend_and_score(Uh, Vh, rph, ith, k, dev, chunks=args.e2e_chunks)
    e2e_ms, res = ctx.timed(lambda: host_recommend_and_score(Uh, Vh, rph, ith, k, dev, chunks=args.e2e_chunks), args.steps)
    # the same call with 32-bit recommendation ids (an option of the API; the headline e2e keeps the reference's int64)
    for _ in range(2):
        host_recommend_and_score(Uh, Vh, rph, ith, k, dev, chunks=args.e2e_chunks, index_dtype=torch.int32)
    e2e32_ms, res32 = ctx.timed(lambda: host_recommend_and_score(Uh, Vh, rph, ith, k, dev, chunks=args.e2e_chunks, index_dtype=torch.int32),
                                max(3, args.steps // 2))
    e2e_value = nU * world / (e2e_ms * 1e-3)
    if rank != 0:
        return None
    pk = peaks()
    med = {name: sorted(ph[name] for ph in phases)[len(phases) // 2] for name in phases[0]}
    flop = 2.0 * nU * nI * d
    fkern = "score_filter2_kernel" if d <= 128 else "score_filter_kernel"
    f_ms, s_ms = med["score_filter_ms"], med["select_ms"]
    avg_seen = float(items.numel()) / nU
    # S2 + S3 + the top-k side traffic of S1: per user the k item rows gathered for the exact scores and the
    # cosine ILD (k (8 + 4 d) + 4), the top-k outputs (12 k), the seen list (4 seen_u + 8).  The candidate records handed
    # over by the filter are this implementation's intermediate data and are NOT counted.
    sel_algo = nU * (k * (8 + 4 * d) + 4 + 12 * k + 4 * avg_seen + 8)
    sel_traffic, src = ncu_traffic("select_topk_kernel", workload)
    fil_traffic, _ = ncu_traffic(fkern, workload)
    roof_select = {"bound": "hbm", "achieved": sel_algo / (s_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                   "frac": sel_algo / (s_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": sel_traffic, "traffic_source": src,
                   "kernel": "select_topk_kernel", "kernel_ms": s_ms, "share_of_step": s_ms / ms_per_step,
                   "algorithmic_bytes_per_launch": sel_algo, "peak_source": pk["source"] + " copy bandwidth",
                   "note": "latency / issue-bound rather than bandwidth-bound (the 13.7 MB item table is L2-resident: row gathers hit L2)"}
    ach = flop / (f_ms * 1e-3) / 1e12
    roof_filter = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                   "frac": ach / pk["bf16_tflops_sustained"], "traffic": fil_traffic, "traffic_source": src, "kernel": fkern,
                   "kernel_ms": f_ms, "share_of_step": f_ms / ms_per_step, "algorithmic_flop_per_launch": flop,
                   "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
                   "note": "kernel_ms includes the row-mask pass that feeds it (one launch list entry each)"}
    # the fused step against the serial sum of its stages' bounds: contraction at the burst tensor peak,
    # ILD gather + coverage at the copy bandwidth
    bound_ms = (flop / (pk["bf16_tflops"] * 1e12) + nU * (k * (8 + 4 * d) + 4) / (pk["hbm_gbs"] * 1e9) + nU * k * 8 / (pk["hbm_gbs"] * 1e9)) * 1e3
    roof_step = {"bound": "tensor + hbm (serial sum of S1, S2, S3)", "achieved": nU / (ms_per_step * 1e-3), "peak": nU / (bound_ms * 1e-3),
                 "unit": "users/s per GPU", "frac": bound_ms / ms_per_step, "bound_ms": bound_ms, "step_ms": ms_per_step}
    cpu = None
    if world == 1 and not args.no_cpu:
        from divrec.parallel import bind_host_to_gpu, unbind_host
        unbind_host(ctx.numa)  # the CPU leg may use every core of the box
        cpu = cpu_baseline(workload, sample_users=args.cpu_sample, literal_users=args.cpu_literal_users)
        ctx.numa = bind_host_to_gpu(ctx.local_rank)
    dominant = roof_select if s_ms >= f_ms else roof_filter
    return {
        "metric": METRIC, "value": value, "unit": "users/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16 tensor-core filter + exact fp32 rescoring, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"{workload}: MovieLens-20M shape" if workload == "cfg2" else workload,
                   "users_per_gpu": nU, "items": nI, "d": d, "k": k, "seen_nnz": int(items.numel()),
                   "stages": "score+mask+top-k, ILD (cosine), coverage histogram",
                   "sharding": "users sharded, item table replicated" if world > 1 else "single GPU",
                   "l2": "inputs + candidate workspace (>1 GB) exceed the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "users/s", "h2d_bytes_per_step": res["h2d_bytes"], "d2h_bytes_per_step": res["d2h_bytes"],
                "api": "divrec.train.host_recommend_and_score (pinned host buffers -> C ABI -> host results)",
                "host_numa_node": ctx.numa["node"], "host_cpus_bound": ctx.numa["cpus"],
                "int32_ids": {"value": world * nU / (e2e32_ms * 1e-3), "d2h_bytes_per_step": res32["d2h_bytes"]}},
        "gpu_launches": int(launches),
        "roofline": dominant, "roofline_select": roof_select, "roofline_filter": roof_filter, "roofline_step": roof_step,
        "cpu_baseline": cpu, "phases_ms": med, "rescued_rows_per_step": stats["rescued_rows"], "step_ms_rank0": step_spread,
        "remeasured": remeasured,
    }


# ------------------------------------------------------------------------------------------------------------------
# cfg4: item table sharded (strong scaling)
# ------------------------------------------------------------------------------------------------------------------


def bench_cfg4(ctx, args, steps, warmup):
    """BASELINE.json configs[3]: 10M-item x d=256 bf16 item table sharded over the GPUs, 65 536 query users per step,
    per-shard score + top-100, NCCL all-gather of the per-shard lists, n-way merge.  Strong scaling: the
    catalogue is fixed, every rank holds 1/N of it.  Synthetic N(0,1) bf16 tables generated on the device."""
    N, dev, dist, world, rank = ctx.N, ctx.dev, ctx.dist, ctx.world, ctx.rank
    from divrec import parallel
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
    warmup = max(1, min(warmup, 2))

    def step():
        return parallel.recommend_item_sharded(Ud, V, a, k)

    sampler = ClockSampler(ctx.local_rank, active=rank == 0).start()
    for _ in range(warmup):
        step()
    ctx.barrier()
    l0 = N.launch_count()
    ms_per_step, _ = ctx.timed(step, steps)
    launches = N.launch_count() - l0
    # phases: the library's events around prep / filter / select, plus the exchange (pack + all-gather + merge) by difference
    N.set_profiling(True)
    step(); torch.cuda.synchronize()
    phases = N.profile_read()
    _s, _i, stats = N.score_topk(Ud, V, k, item_id_offset=a, return_stats=True)
    N.set_profiling(False)
    phases = {kk: v for kk, v in phases.items() if v >= 0}
    phases["exchange_ms (pack + all-gather + merge, by difference)"] = max(0.0, ms_per_step - sum(v for kk, v in phases.items() if kk in ("prep_ms", "score_filter_ms", "select_ms", "rescue_ms")))
    clocks = sampler.stop()

    def e2e_step():
        Ud.copy_(Uh, non_blocking=True)
        s, i = step()
        out_s.copy_(s, non_blocking=True)
        out_i.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_ms, _ = ctx.timed(e2e_step, steps)
    del V
    if rank != 0:
        return None
    pk = peaks()
    flop = 2.0 * nQ * (b - a) * d
    kms = phases["score_filter_ms"]
    ach = flop / (kms * 1e-3) / 1e12
    return {"metric": METRIC.replace("+ILD", ""), "value": nQ / (ms_per_step * 1e-3), "unit": "users/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16 tables (tensor-core filter, exact fp32 rescoring of the candidates)", "data": "synthetic",
            "config": {"workload": "cfg4: item table sharded, per-shard top-k + all-gather + merge", "items_total": nI_total,
                       "items_per_gpu": b - a, "d": d, "k": k, "queries_per_step": nQ, "l2": "item shard exceeds the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": nQ / (e2e_ms * 1e-3), "unit": "users/s", "h2d_bytes_per_step": Uh.numel() * 2,
                    "d2h_bytes_per_step": nQ * k * 12, "api": "divrec.parallel.recommend_item_sharded (queries from pinned host memory)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": ach / pk["bf16_tflops_sustained"], "traffic": None, "kernel": "score_filter_kernel", "kernel_ms": kms,
                         "algorithmic_flop_per_launch": flop, "peak_source": pk["source"] + " bf16 sustained"},
            "cpu_baseline": None, "phases_ms": phases, "rescued_rows_per_step": stats["rescued_rows"]}


# ------------------------------------------------------------------------------------------------------------------
# cfg3: BPR step, users sharded, item gradient exchanged (weak scaling)
# ------------------------------------------------------------------------------------------------------------------


def bench_cfg3(ctx, args, steps, warmup, B=65536):
    """BASELINE.json configs[2]: BPR training step, 1M users x 1M items, d = 128 fp32, B triples per GPU:
    fused gather + loss + AUC forward and scatter-add backward (the reference's autograd semantics).  Users sharded: each
    rank draws its triples' users from its own range.  Item gradient: see divrec.parallel.bpr_step_sharded (`route`).
    Gradient buffers persist across steps (zero_grad belongs to the optimiser: out of scope)."""
    N, dev, dist, world, rank = ctx.N, ctx.dev, ctx.dist, ctx.world, ctx.rank
    from divrec import parallel
    nU = nI = 1_000_000
    d = 128
    g = torch.Generator(device=dev).manual_seed(5)
    U = torch.randn((nU, d), generator=g, device=dev)
    V = torch.randn((nI, d), generator=g, device=dev)
    a, b = parallel.shard_range(nU, rank, world)
    gh = torch.Generator().manual_seed(100 + rank)
    uh = torch.randint(a, b, (B,), generator=gh).pin_memory()
    ph = torch.randint(0, nI, (B,), generator=gh).pin_memory()
    nh = torch.randint(0, nI, (B,), generator=gh).pin_memory()
    u, p_, n_ = uh.to(dev), ph.to(dev), nh.to(dev)
    Bg = B * world
    warmup = max(3, warmup)
    route = args.cfg3_item_grad if world > 1 else "local"
    state = parallel.BprShardedState(U, V, world, rank, B, route=route)

    def step():
        return parallel.bpr_step_sharded(state, u, p_, n_, 1.0 / Bg)

    sampler = ClockSampler(ctx.local_rank, active=rank == 0).start()
    for _ in range(warmup):
        step()
    ctx.barrier()
    l0 = N.launch_count()
    ms_per_step, _ = ctx.timed(step, steps)
    launches = N.launch_count() - l0
    # the kernel alone (no exchange) for the roofline
    gU1, gV1 = torch.zeros_like(U), torch.zeros_like(V)
    loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
    auc_buf = torch.zeros(1, dtype=torch.int32, device=dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(steps):
        N.bpr_step(U, V, u, p_, n_, 1.0 / Bg, gU1, gV1, want_scores=False, loss_sum=loss_buf, auc_sum=auc_buf)
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / steps
    del gU1, gV1
    clocks = sampler.stop()
    loss_h = torch.empty(1, dtype=torch.float32).pin_memory()

    def e2e_step():
        u.copy_(uh, non_blocking=True); p_.copy_(ph, non_blocking=True); n_.copy_(nh, non_blocking=True)
        ls, _auc = step()
        loss_h.copy_(ls, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_ms, _ = ctx.timed(e2e_step, steps)
    desc = state.describe()
    state.close()
    del U, V, state
    if rank != 0:
        return None
    pk = peaks()
    bytes_per_triple = 3 * d * 4 + 3 * 2 * d * 4 + 24   # 3 row reads + 3 row read-modify-writes + ids
    ach = B * bytes_per_triple / (kms * 1e-3) / 1e9
    traffic, src = ncu_traffic("bpr_step_kernel", "cfg3") if B == 65536 else (None, None)
    return {"metric": "BPR triples/sec (gather + loss + scatter-add backward)", "value": Bg / (ms_per_step * 1e-3), "unit": "triples/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: BPR step, 1M users x 1M items", "d": d, "triples_per_gpu": B, "item_gradient": desc,
                       "l2": "tables + dense gradients (2 GB) exceed the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": Bg / (e2e_ms * 1e-3), "unit": "triples/s", "h2d_bytes_per_step": 3 * B * 8,
                    "d2h_bytes_per_step": 4, "api": "divrec.parallel.bpr_step_sharded (triples from pinned host memory, loss back)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                         "traffic": traffic, "traffic_source": src, "kernel": "bpr_step_kernel", "kernel_ms": kms,
                         "algorithmic_bytes_per_launch": B * bytes_per_triple, "peak_source": pk["source"] + " copy bandwidth"},
            "cpu_baseline": None}


# ------------------------------------------------------------------------------------------------------------------
# cfg5: greedy MMR re-rank of top-500 candidate lists (strong scaling, users sharded, no collective)
# ------------------------------------------------------------------------------------------------------------------


def bench_cfg5(ctx, args, steps, warmup):
    """BASELINE.json configs[4]: greedy MMR re-rank of each user's top-500 candidates down to k = 100 (d = 128,
    lambda = 0.5), 1M users in total sharded over the GPUs with no collective (strong scaling).  The candidate lists are
    the library's own output: divrec_score_topk with k = 500 over N(0,1) embeddings (untimed input of this path)."""
    N, dev, dist, world, rank = ctx.N, ctx.dev, ctx.dist, ctx.world, ctx.rank
    from divrec import parallel
    nU_total, nI, d, C, k, lam = args.cfg5_users, 26744, 128, args.cfg5_candidates, 100, 0.5
    ua, ub = parallel.shard_range(nU_total, rank, world)
    nU = ub - ua
    g = torch.Generator(device=dev).manual_seed(9)
    V = torch.randn((nI, d), generator=g, device=dev)
    g.manual_seed(100 + rank)
    cand = torch.empty((nU, C), dtype=torch.int64, device=dev)
    sc = torch.empty((nU, C), dtype=torch.float32, device=dev)
    t_gen0 = time.perf_counter()
    for a in range(0, nU, 131072):
        Uc = torch.randn((min(131072, nU - a), d), generator=g, device=dev)
        s_, i_ = N.score_topk(Uc, V, C)
        sc[a:a + 131072], cand[a:a + 131072] = s_, i_
    del Uc, s_, i_
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t_gen0
    ch, sh = cand.cpu().pin_memory(), sc.cpu().pin_memory()
    out_h = torch.empty((nU, k), dtype=torch.int64).pin_memory()
    warmup = max(1, min(warmup, 3))
    sampler = ClockSampler(ctx.local_rank, active=rank == 0).start()
    for _ in range(warmup):
        N.mmr_rerank(V, cand, sc, k, lam)
    ctx.barrier()
    l0 = N.launch_count()
    ms_per_step, _ = ctx.timed(lambda: N.mmr_rerank(V, cand, sc, k, lam), steps)
    launches = N.launch_count() - l0
    clocks = sampler.stop()

    def e2e_step():
        cand.copy_(ch, non_blocking=True); sc.copy_(sh, non_blocking=True)
        items, rel = N.mmr_rerank(V, cand, sc, k, lam)
        out_h.copy_(items, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_ms, _ = ctx.timed(e2e_step, steps)
    del cand, sc, ch, sh
    if rank != 0:
        return None
    pk = peaks()
    algo = nU * (C * d * 4 + C * 12 + k * 8)  # gather C rows + candidate ids/scores in + k ids out
    ach = algo / (ms_per_step * 1e-3) / 1e9
    return {"metric": "users/sec re-ranked (greedy MMR)", "value": nU_total / (ms_per_step * 1e-3), "unit": "users/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg5: greedy MMR re-rank of top-C candidates", "users_total": nU_total, "users_per_gpu": nU, "candidates": C,
                       "k": k, "lambda": lam, "d": d,
                       "candidates_from": f"divrec_score_topk(k = {C}) over N(0,1) embeddings (untimed input generation, {gen_s:.2f} s)",
                       "l2": "candidate lists + scratch exceed the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": nU_total / (e2e_ms * 1e-3), "unit": "users/s", "h2d_bytes_per_step": nU * C * 12,
                    "d2h_bytes_per_step": nU * k * 8, "api": "divrec.rerank.mmr_rerank (candidates from pinned host memory)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                         "kernel": "mmr_warp_kernel", "kernel_ms": ms_per_step, "algorithmic_bytes_per_launch": algo,
                         "traffic_note": "profiles/r02_ncu_mmr_warp.txt holds a full capture of a smaller launch (not per-launch comparable)",
                         "note": "lazy greedy: only candidates whose upper bound can still win are gathered and evaluated, so the kernel "
                                 "moves fewer bytes than the algorithmic figure; the rest is per-pick latency"},
            "cpu_baseline": None}


def sub_record(line):
    """The part of a workload's line that travels inside another line (SCALE records: one entry per N)."""
    if line is None:
        return None
    keys = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "clocks", "e2e", "gpu_launches",
            "roofline", "phases_ms", "rescued_rows_per_step")
    return {kk: line[kk] for kk in keys if kk in line}


# ------------------------------------------------------------------------------------------------------------------
# CPU legs
# ------------------------------------------------------------------------------------------------------------------


def run_literal_reference(workload, users, steps=1, warmup=0, budget=200.0, check=False):
    """The unmodified reference in its own process (its package is also called `divrec`): oracle/ref_run.py."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_run.py"), "--users", str(users), "--workload", workload,
           "--steps", str(steps), "--warmup", str(warmup), "--budget", str(budget)] + (["--check"] if check else [])
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=budget * 2 + 300)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except (subprocess.SubprocessError, ValueError, IndexError) as e:
        return {"unavailable": f"oracle/ref_run.py failed: {type(e).__name__}"}


def cpu_port(workload, sample_users):
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
    del ild
    return {"value": sample / dt, "unit": "users/s", "cores": cores, "kind": "port",
            "sample": f"{sample} users of {workload} (dense torch-CPU oracle: GEMM + -inf mask + stable sort + Gram ILD + bincount), {dt:.1f} s"}


def cpu_baseline(workload, sample_users, literal_users):
    """cpu_baseline of our arm: the port (the FASTER CPU baseline, hence the conservative ratio) with the literal reference
    loop's users/s beside it."""
    c = cpu_port(workload, sample_users)
    lit = run_literal_reference(workload, literal_users, budget=60.0)
    c["literal_reference"] = ({"users_per_s": lit["users_per_s"], "users": lit["users"], "cores": lit["cores"],
                               "what": "unmodified reference: get_model_recommendations + IntraListDiversityScore + EntropyDiversityScore (oracle/ref_run.py)"}
                              if "users_per_s" in lit else lit)
    return c


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = args.workload if args.workload in WORKLOADS else "cfg2"
    nU, nI, d, k, nnz = WORKLOADS[workload]
    lit = run_literal_reference(workload, args.ref_users, steps=args.steps, warmup=args.warmup, budget=args.ref_budget, check=True)
    port = cpu_port(workload, args.cpu_sample)
    cfg = {"workload": f"{workload}: MovieLens-20M shape" if workload == "cfg2" else workload, "items": nI, "d": d, "k": k}
    if "users_per_s" not in lit:
        # no copy of the reference on this box: the port is the CPU arm (never an `unavailable` line: the oracle always exists)
        v = port["value"]
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "users/s", "n_gpus": world, "steps": 1, "warmup": 0,
                "ms_per_step": args.cpu_sample / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(cfg, users_per_step=min(args.cpu_sample, nU)), "cpu_baseline": port,
                "reference_copy": lit, "e2e": {"value": v, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return
    v = lit["users_per_s"]
    cb = {"value": v, "unit": "users/s", "cores": lit["cores"], "kind": "reference",
          "sample": (f"first {lit['users']} users of {workload} per step, mask table whole ({lit['mask_nnz']} interactions): unmodified "
                     f"get_model_recommendations ({lit['recommend_s']:.2f} s) + IntraListDiversityScore ({lit['ild_s']:.2f} s) + "
                     f"EntropyDiversityScore; median of {lit['steps']} steps; {lit.get('rows_equal_to_oracle')} of {lit['users']} "
                     "rows equal to the oracle restatement"),
          "torch": lit["torch"], "port": port}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "users/s", "n_gpus": world, "steps": lit["steps"], "warmup": lit["warmup"],
            "ms_per_step": lit["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(cfg, users_per_step=lit["users"]), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS) + ["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--sub-steps", type=int, default=5, help="timed steps of the sub-record workloads of --workload all")
    ap.add_argument("--no-sub", action="store_true", help="--workload all without the cfg3 / cfg4 / cfg5 sub-records")
    ap.add_argument("--cfg5-users", type=int, default=1_000_000)
    ap.add_argument("--cfg5-candidates", type=int, default=500)
    ap.add_argument("--cfg3-item-grad", default="owner", choices=["owner", "dense", "sparse"])
    ap.add_argument("--cfg3-batch", type=int, default=65536)
    ap.add_argument("--cfg4-items", type=int, default=10_000_000)
    ap.add_argument("--cfg4-queries", type=int, default=65536)
    ap.add_argument("--cpu-sample", type=int, default=32768)
    ap.add_argument("--cpu-literal-users", type=int, default=48)
    ap.add_argument("--ref-users", type=int, default=200, help="--impl reference: users per step of the literal reference")
    ap.add_argument("--ref-budget", type=float, default=200.0, help="--impl reference: seconds the timed passes may take in total")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="user ranges the host-buffer pipeline overlaps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-baseline leg (profiling runs)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        bench_reference(args)
        return
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device -- divrec_b200 has no CPU path (use --impl reference for the CPU arm)")
    ctx = Ctx()
    try:
        if args.workload == "cfg4":
            line = bench_cfg4(ctx, args, args.steps, args.warmup)
        elif args.workload == "cfg3":
            line = bench_cfg3(ctx, args, args.steps, args.warmup, B=args.cfg3_batch)
        elif args.workload == "cfg5":
            line = bench_cfg5(ctx, args, args.steps, args.warmup)
        elif args.workload in WORKLOADS:
            line = bench_eval(ctx, args, args.workload)
        else:
            line = bench_eval(ctx, args, "cfg2")
            if not args.no_sub:
                subs = {}
                torch.cuda.empty_cache()
                subs["cfg4"] = sub_record(bench_cfg4(ctx, args, args.sub_steps, 1))
                torch.cuda.empty_cache()
                small = sub_record(bench_cfg3(ctx, args, max(args.sub_steps, 10), 3, B=65536))
                torch.cuda.empty_cache()
                large = sub_record(bench_cfg3(ctx, args, args.sub_steps, 3, B=524288))
                subs["cfg3"] = {"B65536": small, "B524288": large} if ctx.rank == 0 else None
                torch.cuda.empty_cache()
                subs["cfg5"] = sub_record(bench_cfg5(ctx, args, 2, 1))
                if line is not None:
                    line.update(subs)
        if ctx.rank == 0 and line is not None:
            emit(line)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
