"""Multi-GPU host logic: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

The path shards two ways:
  * users sharded, item table replicated -- no data-path collective; the metric reductions are one all_reduce
    of the [I] coverage histogram and one of the (ILD sum, user count) pair;
  * item table sharded (large catalogues) -- every rank scores all query users against its slice, one
    all_gather of the per-shard [U,k] (score, id) lists, then the n-way merge kernel;
  * BPR -- triples partitioned by user shard (user gradients stay local); the item gradient is either sharded by
    OWNER and exchanged inside the step kernel over NVLink peer memory (divrec_bpr_step_owner, the default on GPUs), or
    all-reduced dense, or all-gathered as per-triple rows (BprShardedState.route).

The compute backend is the ctypes binding of libdivrec_b200.so, always: there is no backend parameter.  (The 2-rank gloo
tests on CPU replace this module's `_native` attribute with an oracle-backed stand-in -- a monkeypatch in the test, not
a seam in the product.)
"""
from typing import Optional, Tuple

import os

import torch
import torch.distributed as dist

from . import _native


def _kernels():
    return _native


def _world(group) -> Tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced ranges: the first n % world shards get one extra element."""
    base, extra = divmod(n, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def recommend_user_sharded(U: torch.Tensor, V: torch.Tensor, seen_rowptr: Optional[torch.Tensor],
                           seen_items: Optional[torch.Tensor], k: int, group=None):
    """Score this rank's contiguous range of users against the replicated item table.
    Returns (items [n_local,k], scores [n_local,k], ild_local [n_local], ild_mean_global, coverage_counts_global, (a,b))."""
    K = _kernels()
    rank, world = _world(group)
    nU = U.shape[0]
    a, b = shard_range(nU, rank, world)
    rp = None
    if seen_rowptr is not None:
        rp = seen_rowptr[a:b + 1].contiguous()  # absolute offsets into the (replicated) seen_items array
    scores, items, ild, counts = K.score_topk_metrics(U[a:b].contiguous(), V, k, rp, seen_items)
    stat = torch.stack([ild.sum(dtype=torch.float64), torch.tensor(float(b - a), dtype=torch.float64, device=ild.device)])
    if world > 1:
        dist.all_reduce(counts, group=group)
        dist.all_reduce(stat, group=group)
    return items, scores, ild, (stat[0] / stat[1]).to(torch.float32), counts, (a, b)


_route_cache = {}


def _all_gather_stacked(t: torch.Tensor, world: int, group) -> torch.Tensor:
    """all_gather of equally shaped [n, k] tensors straight into one [world, n, k] buffer (no per-rank copies)."""
    flat = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(flat, t.contiguous(), group=group)
    return flat.view((world,) + tuple(t.shape))


def recommend_item_sharded(U: torch.Tensor, V_shard: torch.Tensor, shard_offset: int, k: int,
                           seen_rowptr: Optional[torch.Tensor] = None, seen_items: Optional[torch.Tensor] = None,
                           group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank holds items [shard_offset, shard_offset + len(V_shard)) and all query users.
    Local top-k with global ids -> all_gather -> merge (ties -> lowest id; deterministic because the merge
    uses the total order (score desc, id asc)).  Returns the global (scores [U,k], items [U,k]) on every rank."""
    K = _kernels()
    rank, world = _world(group)
    s, i = K.score_topk(U, V_shard, k, seen_rowptr, seen_items, item_id_offset=shard_offset)
    if world == 1:
        return s, i
    # the exchange step.  Packed form (one 64-bit composite per pair, one all-gather, integer merge) when the
    # backend has it and the sizes allow; otherwise the two-tensor form.
    packed = (hasattr(K, "topk_pack") and world * k <= getattr(K, "MERGE_PACKED_MAX", 0)
              and shard_offset + V_shard.shape[0] < 2 ** 32)
    key = (id(group), world, k, int(shard_offset), int(V_shard.shape[0]), bool(packed))
    route = _route_cache.get(key)
    if route is None:  # every rank must take the same route: agreed once per configuration (a host sync), then cached
        flag = torch.tensor([1 if packed else 0], dtype=torch.int32, device=s.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        route = _route_cache[key] = int(flag)
    if route == 1:
        mine = K.topk_pack(s, i)
        return K.topk_merge_packed(_all_gather_stacked(mine, world, group))
    return K.topk_merge(_all_gather_stacked(s, world, group), _all_gather_stacked(i, world, group))


class BprShardedState:
    """Buffers and exchange route of the user-sharded BPR step (BASELINE.json config 3), built once and reused every step.

    route:
      "local"   one GPU: dense gU / gV, no exchange;
      "owner"   (default for world > 1) the item gradient is SHARDED BY OWNER: rank r owns items [r S, (r+1) S),
                S = ceil(nI / world), in a peer-mapped buffer (PeerBuffer, CUDA IPC); divrec_bpr_step_owner scatters every
                item-row update straight to its owner over NVLink inside the step kernel.  After the step's one collective
                (the loss / AUC all-reduce, which is also the barrier behind the remote reductions) `gV_slice` holds the
                globally summed gradient of the rank's own items -- reduce-scatter semantics, 1/world of the table per rank;
      "dense"   the [nI, d] item gradient is all-reduced (BASELINE.json config 3 as literally stated);
      "sparse"  the per-triple rows g_b U[u_b] and their ids are all-gathered and every rank scatter-adds all of them
                (same dense result on every rank; needs equal batch sizes on all ranks).
    Gradient buffers accumulate across steps (zero_grad belongs to the optimiser)."""

    def __init__(self, U: torch.Tensor, V: torch.Tensor, world: int, rank: int, batch: int, route: str = "owner", group=None):
        K = _kernels()
        if world == 1:
            route = "local"
        if route not in ("local", "owner", "dense", "sparse"):
            raise ValueError(f"unknown route {route!r}")
        self.U, self.V, self.world, self.rank, self.batch, self.route, self.group = U, V, world, rank, batch, route, group
        dev, (nI, d) = U.device, V.shape
        self.gU = torch.zeros_like(U)
        # step statistics.  owner route: ONE float pair (loss sum, AUC hits) -> one all-reduce; the others: the kernel's float
        # loss and int32 AUC accumulators side by side in one 8-byte buffer (one fill kernel clears both)
        self.stats2 = torch.zeros(2, dtype=torch.float32, device=dev)
        self._acc = torch.zeros(2, dtype=torch.int32, device=dev)
        self.loss_sum = self._acc[0:1].view(torch.float32)
        self.auc_sum = self._acc[1:2]
        self.gV = self.peer = self.gV_slice = None
        self.slice_rows = -(-nI // world)
        if route == "owner":
            self.peer = K.PeerBuffer(self.slice_rows, d, dev, group=group)
            self.gV_slice = self.peer.local
            self.gV_slice.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)  # every slice is zeroed before any rank's kernel adds into it
        else:
            self.gV = torch.zeros_like(V)
        if route == "sparse":
            self.rows = torch.empty((batch, d), dtype=torch.float32, device=dev)
            self.all_rows = torch.empty((world * batch, d), dtype=torch.float32, device=dev)
            self.ids = torch.empty((2, batch), dtype=torch.int64, device=dev)
            self.all_ids = torch.empty((world * 2, batch), dtype=torch.int64, device=dev)

    def describe(self) -> str:
        return {"local": "single GPU (dense gradient tables)",
                "owner": "sharded by owner: row updates scattered to the owning rank's peer-mapped slice inside the step kernel "
                         "(red.global.add.v4.f32 over NVLink), one all-reduce of the loss / AUC scalars",
                "dense": "dense all-reduce of the [nI, d] item gradient",
                "sparse": "all-gather of per-triple rows + ids, every rank scatter-adds all of them"}[self.route]

    def owned_range(self) -> Tuple[int, int]:
        a = self.rank * self.slice_rows
        return a, min(self.V.shape[0], a + self.slice_rows)

    def close(self):
        if self.peer is not None:
            dist.barrier(group=self.group)  # no rank may still be adding into a slice that is about to be unmapped
            self.gV_slice = None
            self.peer.close()
            self.peer = None


def bpr_step_sharded(state: BprShardedState, u: torch.Tensor, p: torch.Tensor, n: torch.Tensor, grad_scale: float):
    """One BPR forward/backward on this rank's triples (its users only), gradients accumulated into the state's buffers.
    grad_scale = 1 / global_batch for the GLOBAL mean loss.  Returns (loss_sum [1], auc_hits [1]) device tensors summed
    over ALL ranks' triples of this step (divide by the global batch for the means); no host synchronisation."""
    K = _kernels()
    st = state
    if st.route == "owner":
        st.stats2.zero_()
        K.bpr_step_owner(st.U, st.V, u, p, n, grad_scale, st.gU, st.peer.ptrs, st.slice_rows, st.stats2)
        # the step's one collective, and the barrier of the remote reductions: when it completes, every rank's step kernel
        # -- and with it every update of this rank's slice -- has finished
        dist.all_reduce(st.stats2, group=st.group)
        return st.stats2[0:1], st.stats2[1:2]
    st._acc.zero_()
    if st.route == "sparse":
        K.bpr_step(st.U, st.V, u, p, n, grad_scale, st.gU, None, want_scores=False, loss_sum=st.loss_sum, auc_sum=st.auc_sum, item_rows=st.rows)
        st.ids[0].copy_(p); st.ids[1].copy_(n)
        dist.all_gather_into_tensor(st.all_rows, st.rows, group=st.group)
        dist.all_gather_into_tensor(st.all_ids, st.ids, group=st.group)
        g3 = st.all_ids.view(st.world, 2, st.batch)
        K.bpr_apply_rows(st.all_rows, g3[:, 0].reshape(-1), g3[:, 1].reshape(-1), st.gV)
    else:
        K.bpr_step(st.U, st.V, u, p, n, grad_scale, st.gU, st.gV, want_scores=False, loss_sum=st.loss_sum, auc_sum=st.auc_sum)
        if st.route == "dense":
            dist.all_reduce(st.gV, group=st.group)
    if st.world > 1:
        dist.all_reduce(st.loss_sum, group=st.group)
        dist.all_reduce(st.auc_sum, group=st.group)
    return st.loss_sum, st.auc_sum


def bpr_step_once(U: torch.Tensor, V: torch.Tensor, u: torch.Tensor, p: torch.Tensor, n: torch.Tensor,
                  global_batch: int, group=None, item_grad: str = "dense"):
    """One step on fresh buffers (tests, one-off calls): returns (loss_global, auc_global, gU_local_rows_dense, gV) where gV
    is the dense global item gradient, or -- item_grad="owner" -- this rank's slice of it.
    item_grad="auto": the ranks agree on their batch sizes first (one tiny all-reduce and a host sync) and take the sparse
    route when they are equal and the gathered rows are at most half the all-reduce volume, else the dense one."""
    rank, world = _world(group)
    B, d = u.numel(), U.shape[1]
    route = item_grad
    if world == 1:
        route = "local"
    elif item_grad == "auto":
        mm = torch.tensor([B, -B], dtype=torch.int64, device=U.device)
        dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
        bmax, bmin = int(mm[0]), -int(mm[1])
        route = "sparse" if (bmax == bmin and world * B * (d + 4) * 4 <= V.shape[0] * d * 4) else "dense"
    state = BprShardedState(U, V, world, rank, B, route=route, group=group)
    loss_sum, auc_hits = bpr_step_sharded(state, u, p, n, 1.0 / global_batch)
    loss, auc = (loss_sum[0] / global_batch).to(torch.float32), (auc_hits[0].to(torch.float64) / global_batch).to(torch.float32)
    gV = state.gV_slice.clone() if route == "owner" else state.gV
    gU = state.gU
    state.close()
    return loss, auc, gU, gV


# ---- host placement ------------------------------------------------------------------------------
def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_gpu(device_index: int, sysfs: str = "/sys") -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated afterwards
    (first touch) and the copy-issuing threads are local to the GPU's PCIe root.  One process per GPU on an 8-GPU box
    otherwise leaves half of the ranks staging their host<->device traffic through the other socket.  Returns
    {"node", "cpus", "saved"}; node is None (and nothing is changed) when the topology cannot be read -- a container
    without sysfs, one NUMA node, no CUDA device.  Undo with unbind_host(info)."""
    info = {"node": None, "cpus": 0, "saved": None}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(os.path.join(sysfs, "devices/system/node", "node%d" % node, "cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        saved = os.sched_getaffinity(0)
        cpus &= saved
        if not cpus or cpus == saved:
            return info
        os.sched_setaffinity(0, cpus)
        info.update(node=node, cpus=len(cpus), saved=saved)
    except (OSError, ValueError, RuntimeError, AssertionError, AttributeError):
        pass
    return info


def unbind_host(info: dict) -> None:
    """Restore the CPU affinity bind_host_to_gpu() replaced (e.g. before a CPU leg that should use every core)."""
    if info.get("saved"):
        os.sched_setaffinity(0, info["saved"])
