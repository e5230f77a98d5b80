"""ctypes binding of libdivrec_b200.so (the C ABI declared in include/divrec_b200.h).

PyTorch is used here for device memory and streams only: every wrapper takes CUDA tensors,
passes raw device pointers and the current CUDA stream to the library, and returns tensors it
allocated.  There is no CPU path: a CPU tensor raises TypeError, a missing library raises
ImportError at first use (the product must fail loudly when the CUDA extension is absent).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, "libdivrec_b200.so")

DIVREC_F32, DIVREC_BF16 = 0, 1
MODE_TENSOR, MODE_SIMT = 0, 2
ILD_SUMVEC, ILD_PAIRWISE = 0, 1
MAX_K = 512

_c = ctypes
_vp, _i64, _i32, _f32, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/divrec_b200.h one to one
SIGNATURES = {
    "divrec_version": (_i32, []),
    "divrec_last_error": (_c.c_char_p, []),
    "divrec_launch_count": (_i64, []),
    "divrec_set_profiling": (None, [_i32]),
    "divrec_debug_filter_mode": (None, [_i32]),
    "divrec_debug_filter_stats": (None, [_vp]),
    "divrec_profile_read": (_i32, [_c.POINTER(_c.c_float), _i32]),
    "divrec_score_topk_workspace": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "divrec_score_topk": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "divrec_score_topk_metrics": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "divrec_score_topk_rescued": (_i32, [_vp, _i64, _i64, _i32, _i32, _c.POINTER(_c.c_int32), _vp]),
    "divrec_topk_merge": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    "divrec_topk_pack": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "divrec_topk_merge_packed": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    "divrec_ild_cosine": (_i32, [_vp, _i64, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "divrec_ild_matrix": (_i32, [_vp, _i32, _i64, _vp, _i64, _i32, _vp, _vp]),
    "divrec_coverage_hist": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "divrec_entropy_from_counts": (_i32, [_vp, _i64, _vp, _vp]),
    "divrec_bpr_fwd": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "divrec_bpr_bwd": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i32, _vp, _vp, _vp]),
    "divrec_bpr_step": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "divrec_bpr_step_owner": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "divrec_bad_index_count": (_i32, [_c.POINTER(_c.c_int64), _i32, _vp]),
    "divrec_peer_alloc": (_i32, [_sz, _c.POINTER(_vp), _c.POINTER(_c.c_ubyte)]),
    "divrec_peer_open": (_i32, [_c.POINTER(_c.c_ubyte), _c.POINTER(_vp)]),
    "divrec_peer_close": (_i32, [_vp]),
    "divrec_peer_free": (_i32, [_vp]),
    "divrec_bpr_apply_rows": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "divrec_bpr_sample_triples": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _c.c_uint64, _vp, _vp, _vp, _vp]),
    "divrec_pair_scores": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i32, _vp, _vp]),
    "divrec_pair_scores_bwd": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "divrec_ild_labels": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    "divrec_coverage_hist_ranks": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _vp]),
    "divrec_relevance_metrics": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "divrec_bpr_loss_fwd": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "divrec_bpr_loss_bwd": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "divrec_build_csr_workspace": (_sz, [_i64, _i64]),
    "divrec_build_csr": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "divrec_mmr_workspace": (_sz, [_i32, _i32]),
    "divrec_mmr_rerank": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


class DivrecNativeError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen the shared library and set the prototypes.  Does not touch the GPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            f"{p} is missing: build it with `python diversity-recommendations_b200/build.py` "
            "(nvcc, sm_100a). divrec_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().divrec_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise DivrecNativeError(f"{what} failed (status {rc}): {msg}")


def _cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor: divrec_b200 has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _table_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DIVREC_F32
    if t.dtype == torch.bfloat16:
        return DIVREC_BF16
    raise TypeError(f"embedding tables must be float32 or bfloat16, got {t.dtype}")


def launch_count() -> int:
    return int(load_library().divrec_launch_count())


def set_profiling(on: bool) -> None:
    load_library().divrec_set_profiling(1 if on else 0)


PROFILE_SLOTS = ("prep_ms", "score_filter_ms", "select_ms", "ild_ms", "coverage_ms", "rescue_ms")


def profile_read() -> dict:
    """Durations (ms) of the library's last phases, from CUDA events on the launching stream."""
    buf = (_c.c_float * len(PROFILE_SLOTS))()
    _check(load_library().divrec_profile_read(buf, len(PROFILE_SLOTS)), "divrec_profile_read")
    return {name: float(buf[i]) for i, name in enumerate(PROFILE_SLOTS)}


_ws_cache = {}


def _workspace(nbytes: int, dev) -> torch.Tensor:
    """One cached byte buffer per device, grown on demand, handed out as a 1024-byte aligned view.
    The buffer is shared by every call on the device: calls must be issued on ONE stream at a time (the wrappers use
    the current stream; callers that overlap streams bring their own workspace through the C ABI)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    # the caching allocator only guarantees 512-byte alignment: keep room for the worst-case offset of the aligned view
    if buf is None or buf.numel() < nbytes + 1024:
        buf = None
        _ws_cache.pop(key, None)
        buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    off = (-buf.data_ptr()) % 1024
    view = buf[off:off + nbytes]
    assert view.numel() == nbytes
    return view


def score_topk(U: torch.Tensor, V: torch.Tensor, k: int, seen_rowptr: Optional[torch.Tensor] = None,
               seen_items: Optional[torch.Tensor] = None, item_id_offset: int = 0, mode: int = MODE_TENSOR,
               return_stats: bool = False):
    """(scores [nU,k] fp32, items [nU,k] int64): the k best unmasked items per user, (score desc, id asc).
    With return_stats=True a third value {"rescued_rows": n} is returned (synchronises the stream)."""
    lib = load_library()
    U = _cuda(U, "U")
    V = _cuda(V, "V", U.dtype)
    dt = _table_dtype(U)
    if U.dim() != 2 or V.dim() != 2 or U.shape[1] != V.shape[1]:
        raise ValueError("U and V must be [n,d] with the same d")
    nU, d = U.shape
    nI = V.shape[0]
    if seen_rowptr is not None:
        seen_rowptr = _cuda(seen_rowptr, "seen_rowptr", torch.int64)
        if seen_rowptr.numel() != nU + 1:
            raise ValueError("seen_rowptr must have nU + 1 entries")
        if seen_items is not None:
            seen_items = _cuda(seen_items, "seen_items", torch.int32)
            if seen_items.numel() == 0:
                seen_items = None
    with torch.cuda.device(U.device):
        ws_bytes = lib.divrec_score_topk_workspace(nU, nI, d, dt, k)
        ws = _workspace(ws_bytes, U.device)
        out_s = torch.empty((nU, k), dtype=torch.float32, device=U.device)
        out_i = torch.empty((nU, k), dtype=torch.int64, device=U.device)
        rc = lib.divrec_score_topk(_ptr(U), _ptr(V), nU, nI, d, dt, mode, _ptr(seen_rowptr), _ptr(seen_items),
                                   item_id_offset, k, _ptr(out_s), _ptr(out_i), _ptr(ws), ws_bytes, _stream(U.device))
        _check(rc, "divrec_score_topk")
        if return_stats:
            n = (_c.c_int32 * 5)()
            _check(lib.divrec_score_topk_rescued(_ptr(ws), nU, nI, d, dt, n, _stream(U.device)), "divrec_score_topk_rescued")
            return out_s, out_i, {"rescued_rows": int(n[0]), "list_overflow": int(n[1]), "too_many_candidates": int(n[2]),
                                  "hint_not_certified": int(n[3]), "too_few_captured": int(n[4])}
    return out_s, out_i


def score_topk_metrics(U: torch.Tensor, V: torch.Tensor, k: int, seen_rowptr: Optional[torch.Tensor] = None,
                       seen_items: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                       want_ild: bool = True, want_coverage: bool = True, mode: int = MODE_TENSOR, out=None):
    """score + mask + top-k with the cosine ILD and the coverage histogram fused into the selection kernel.
    Returns (scores [nU,k], items [nU,k], ild [nU] or None, counts int32 [nI] or None).
    `out` = (scores, items, ild) preallocated contiguous views to write into (optional)."""
    lib = load_library()
    U = _cuda(U, "U")
    V = _cuda(V, "V", U.dtype)
    dt = _table_dtype(U)
    if U.dim() != 2 or V.dim() != 2 or U.shape[1] != V.shape[1]:
        raise ValueError("U and V must be [n,d] with the same d")
    nU, d = U.shape
    nI = V.shape[0]
    if seen_rowptr is not None:
        seen_rowptr = _cuda(seen_rowptr, "seen_rowptr", torch.int64)
        if seen_rowptr.numel() != nU + 1:
            raise ValueError("seen_rowptr must have nU + 1 entries")
        if seen_items is not None:
            seen_items = _cuda(seen_items, "seen_items", torch.int32)
            if seen_items.numel() == 0:
                seen_items = None
    with torch.cuda.device(U.device):
        ws_bytes = lib.divrec_score_topk_workspace(nU, nI, d, dt, k)
        ws = _workspace(ws_bytes, U.device)
        if out is not None:
            out_s, out_i, ild = out
            assert out_s.is_contiguous() and out_i.is_contiguous() and (ild is None or ild.is_contiguous())
            assert out_s.shape == (nU, k) and out_i.shape == (nU, k) and out_s.dtype == torch.float32 and out_i.dtype == torch.int64
        else:
            out_s = torch.empty((nU, k), dtype=torch.float32, device=U.device)
            out_i = torch.empty((nU, k), dtype=torch.int64, device=U.device)
            ild = torch.empty(nU, dtype=torch.float32, device=U.device) if want_ild else None
        if want_coverage and counts is None:
            counts = torch.zeros(nI, dtype=torch.int32, device=U.device)
        rc = lib.divrec_score_topk_metrics(_ptr(U), _ptr(V), nU, nI, d, dt, mode, _ptr(seen_rowptr), _ptr(seen_items), k,
                                           _ptr(out_s), _ptr(out_i), _ptr(ild), _ptr(counts) if want_coverage else None,
                                           _ptr(ws), ws_bytes, _stream(U.device))
    _check(rc, "divrec_score_topk_metrics")
    return out_s, out_i, ild, (counts if want_coverage else None)


def topk_merge(gathered_scores: torch.Tensor, gathered_items: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = load_library()
    gs = _cuda(gathered_scores, "gathered_scores", torch.float32)
    gi = _cuda(gathered_items, "gathered_items", torch.int64)
    if gs.dim() != 3 or gs.shape != gi.shape:
        raise ValueError("gathered tensors must be [n_shards, nU, k]")
    n_shards, nU, k = gs.shape
    out_s = torch.empty((nU, k), dtype=torch.float32, device=gs.device)
    out_i = torch.empty((nU, k), dtype=torch.int64, device=gs.device)
    with torch.cuda.device(gs.device):
        rc = lib.divrec_topk_merge(_ptr(gs), _ptr(gi), n_shards, nU, k, _ptr(out_s), _ptr(out_i), _stream(gs.device))
    _check(rc, "divrec_topk_merge")
    return out_s, out_i


MERGE_PACKED_MAX = 6144  # n_shards * k the shared-memory merge takes


def topk_pack(scores: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
    """[nU,k] (fp32 score, int64 global id < 2^32) -> [nU,k] int64 composites whose descending UNSIGNED order is
    (score desc, id asc); padding (-inf, -1) -> 0.  The exchange format of the item-sharded all-gather."""
    lib = load_library()
    s = _cuda(scores, "scores", torch.float32)
    i = _cuda(items, "items", torch.int64)
    if s.shape != i.shape:
        raise ValueError("scores and items must have the same shape")
    out = torch.empty(s.shape, dtype=torch.int64, device=s.device)
    with torch.cuda.device(s.device):
        rc = lib.divrec_topk_pack(_ptr(s), _ptr(i), s.numel(), _ptr(out), _stream(s.device))
    _check(rc, "divrec_topk_pack")
    return out


def topk_merge_packed(gathered: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n_shards, nU, k] composites (topk_pack) -> global (scores [nU,k] fp32, items [nU,k] int64)."""
    lib = load_library()
    g = _cuda(gathered, "gathered", torch.int64)
    if g.dim() != 3:
        raise ValueError("gathered must be [n_shards, nU, k]")
    n_shards, nU, k = g.shape
    out_s = torch.empty((nU, k), dtype=torch.float32, device=g.device)
    out_i = torch.empty((nU, k), dtype=torch.int64, device=g.device)
    with torch.cuda.device(g.device):
        rc = lib.divrec_topk_merge_packed(_ptr(g), n_shards, nU, k, _ptr(out_s), _ptr(out_i), _stream(g.device))
    _check(rc, "divrec_topk_merge_packed")
    return out_s, out_i


def ild_cosine(V: torch.Tensor, recs: torch.Tensor, mode: int = ILD_SUMVEC) -> torch.Tensor:
    """per-user sum_{a<b}(1 - cos(V[rec_a], V[rec_b])) / (k (k-1)), fp32 [nU]"""
    lib = load_library()
    V = _cuda(V, "V")
    dt = _table_dtype(V)
    recs = _cuda(recs, "recommendations", torch.int64)
    if recs.dim() != 2:
        raise ValueError("recommendations must be [U,k]")
    nU, k = recs.shape
    nI, d = V.shape
    inv = torch.empty(nI, dtype=torch.float32, device=V.device)
    out = torch.empty(nU, dtype=torch.float32, device=V.device)
    with torch.cuda.device(V.device):
        rc = lib.divrec_ild_cosine(_ptr(V), nI, d, dt, _ptr(recs), nU, k, mode, _ptr(inv), _ptr(out), _stream(V.device))
    _check(rc, "divrec_ild_cosine")
    return out


def ild_matrix(D: torch.Tensor, recs: torch.Tensor) -> torch.Tensor:
    lib = load_library()
    if D.dtype not in (torch.float32, torch.int32):
        raise TypeError("distance_matrix must be float32 or int32")
    D = _cuda(D, "distance_matrix")
    recs = _cuda(recs, "recommendations", torch.int64)
    nU, k = recs.shape
    nI = D.shape[0]
    if D.dim() != 2 or D.shape[1] != nI:
        raise ValueError("distance_matrix must be square [I,I]")
    out = torch.empty(nU, dtype=torch.float32, device=D.device)
    with torch.cuda.device(D.device):
        rc = lib.divrec_ild_matrix(_ptr(D), int(D.dtype == torch.int32), nI, _ptr(recs), nU, k, _ptr(out), _stream(D.device))
    _check(rc, "divrec_ild_matrix")
    return out


def coverage_hist(recs: torch.Tensor, n_items: int, counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int32 [n_items] histogram of the recommended ids (accumulates into `counts` if given)."""
    lib = load_library()
    recs = _cuda(recs, "recommendations", torch.int64)
    if counts is None:
        counts = torch.zeros(n_items, dtype=torch.int32, device=recs.device)
    else:
        counts = _cuda(counts, "counts", torch.int32)
    with torch.cuda.device(recs.device):
        rc = lib.divrec_coverage_hist(_ptr(recs), recs.numel(), n_items, _ptr(counts), _stream(recs.device))
    _check(rc, "divrec_coverage_hist")
    return counts


def entropy_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """fp32 [2]: (-sum p log p, number of non-zero counts)"""
    lib = load_library()
    counts = _cuda(counts, "counts", torch.int32)
    out = torch.empty(2, dtype=torch.float32, device=counts.device)
    with torch.cuda.device(counts.device):
        rc = lib.divrec_entropy_from_counts(_ptr(counts), counts.numel(), _ptr(out), _stream(counts.device))
    _check(rc, "divrec_entropy_from_counts")
    return out


def build_csr(pairs: torch.Tensor, n_users: int, n_items: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[nnz,2] int64 (user,item) pairs on the device -> (rowptr int64 [U+1], items int32 ascending and unique per row).
    Pairs outside [0,U) x [0,I) are dropped.  One host synchronisation at the end (the number of distinct pairs sizes the
    returned item array)."""
    lib = load_library()
    pairs = _cuda(pairs, "interactions", torch.int64)
    if pairs.dim() != 2 or pairs.shape[1] != 2:
        raise ValueError("interactions must be [nnz, 2]")
    nnz, dev = pairs.shape[0], pairs.device
    rowptr = torch.empty(n_users + 1, dtype=torch.int64, device=dev)
    items = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.divrec_build_csr_workspace(nnz, n_users)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        rc = lib.divrec_build_csr(_ptr(pairs) if nnz else None, nnz, int(n_users), int(n_items), _ptr(rowptr), _ptr(items),
                                  ws.data_ptr() + off, nbytes, _stream(dev))
    _check(rc, "divrec_build_csr")
    return rowptr, items[:int(rowptr[-1])]


def bpr_fwd(U, V, u, p, n):
    """returns pos [B], neg [B], loss_sum [1] fp32, auc_sum [1] int32"""
    lib = load_library()
    U = _cuda(U, "U", torch.float32)
    V = _cuda(V, "V", torch.float32)
    u = _cuda(u, "user_id", torch.int64)
    p = _cuda(p, "positive_item_id", torch.int64)
    n = _cuda(n, "negative_item_id", torch.int64)
    B = u.numel()
    d = U.shape[1]
    pos = torch.empty(B, dtype=torch.float32, device=U.device)
    neg = torch.empty(B, dtype=torch.float32, device=U.device)
    loss_sum = torch.zeros(1, dtype=torch.float32, device=U.device)
    auc_sum = torch.zeros(1, dtype=torch.int32, device=U.device)
    with torch.cuda.device(U.device):
        rc = lib.divrec_bpr_fwd(_ptr(U), _ptr(V), U.shape[0], V.shape[0], _ptr(u), _ptr(p), _ptr(n), B, d, _ptr(pos), _ptr(neg),
                                _ptr(loss_sum), _ptr(auc_sum), _stream(U.device))
    _check(rc, "divrec_bpr_fwd")
    return pos, neg, loss_sum, auc_sum


def bpr_step(U, V, u, p, n, grad_scale: float, gU: Optional[torch.Tensor], gV: Optional[torch.Tensor],
             want_scores: bool = True, loss_sum: Optional[torch.Tensor] = None, auc_sum: Optional[torch.Tensor] = None,
             item_rows: Optional[torch.Tensor] = None):
    """Forward + backward of one batch in one kernel (upstream gradient of the loss = grad_scale * B, i.e. grad_scale = 1/B
    for `mean_loss.backward()`): gradients accumulate into the dense tables gU / gV.
    item_rows [B,d] (optional) receives the sparse form of the item gradient (row b = g_b U[u_b]; see bpr_apply_rows).
    Returns pos [B] | None, neg [B] | None, loss_sum [1] fp32, auc_sum [1] int32."""
    lib = load_library()
    U = _cuda(U, "U", torch.float32)
    V = _cuda(V, "V", torch.float32)
    u = _cuda(u, "user_id", torch.int64)
    p = _cuda(p, "positive_item_id", torch.int64)
    n = _cuda(n, "negative_item_id", torch.int64)
    for name, t in (("gU", gU), ("gV", gV), ("item_rows", item_rows)):
        if t is not None and (not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32):
            raise TypeError(f"{name} must be a contiguous float32 CUDA tensor")
    B = u.numel()
    d = U.shape[1]
    if item_rows is not None and tuple(item_rows.shape) != (B, d):
        raise ValueError("item_rows must be [B, d]")
    want_scores = want_scores or d % 4 != 0  # the two-pass form of odd row widths needs the score buffers
    pos = torch.empty(B, dtype=torch.float32, device=U.device) if want_scores else None
    neg = torch.empty(B, dtype=torch.float32, device=U.device) if want_scores else None
    if loss_sum is None:
        loss_sum = torch.zeros(1, dtype=torch.float32, device=U.device)
    if auc_sum is None:
        auc_sum = torch.zeros(1, dtype=torch.int32, device=U.device)
    with torch.cuda.device(U.device):
        rc = lib.divrec_bpr_step(_ptr(U), _ptr(V), U.shape[0], V.shape[0], _ptr(u), _ptr(p), _ptr(n), B, d, float(grad_scale), _ptr(pos), _ptr(neg),
                                 _ptr(loss_sum), _ptr(auc_sum), _ptr(gU), _ptr(gV), _ptr(item_rows), _stream(U.device))
    _check(rc, "divrec_bpr_step")
    return pos, neg, loss_sum, auc_sum


def bad_index_count(device=None, reset: bool = True) -> int:
    """Triples / pairs the row-gather kernels skipped since the last reset because an id was outside its table
    (synchronises the current stream).  The reference's nn.Embedding raises IndexError for such ids."""
    lib = load_library()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = _c.c_int64(0)
    with torch.cuda.device(dev):
        _check(lib.divrec_bad_index_count(_c.byref(out), 1 if reset else 0, _stream(dev)), "divrec_bad_index_count")
    return int(out.value)


def raise_on_bad_indices(device=None) -> None:
    n_bad = bad_index_count(device)
    if n_bad:
        raise IndexError(f"{n_bad} (user, item) ids were outside the embedding tables (index out of range in self)")


class PeerBuffer:
    """A float32 [rows, cols] device buffer allocated by the library (cudaMalloc) and exported over CUDA IPC, plus the
    peers' buffers mapped into this process: `ptrs` is an int64 device tensor of world addresses (this rank's own buffer at
    index `rank`), `local` a torch view of the own buffer.  Used by the owner-sharded BPR step (divrec_bpr_step_owner)."""

    def __init__(self, rows: int, cols: int, device, group=None):
        import torch.distributed as dist
        lib = load_library()
        self.dev = torch.device(device)
        self.rows, self.cols = int(rows), int(cols)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._own = _vp()
        handle = (_c.c_ubyte * 64)()
        with torch.cuda.device(self.dev):
            _check(lib.divrec_peer_alloc(self.rows * self.cols * 4, _c.byref(self._own), handle), "divrec_peer_alloc")
        self._opened = []
        addrs = [0] * self.world
        addrs[self.rank] = int(self._own.value)
        if self.world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.dev)
            allh = torch.empty((self.world, 64), dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu()
            with torch.cuda.device(self.dev):
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    hb = (_c.c_ubyte * 64)(*allh[r].tolist())
                    ptr = _vp()
                    _check(lib.divrec_peer_open(hb, _c.byref(ptr)), "divrec_peer_open")
                    self._opened.append(ptr)
                    addrs[r] = int(ptr.value)
        self.ptrs = torch.tensor(addrs, dtype=torch.int64, device=self.dev)

        class _Raw:  # __cuda_array_interface__ view of the own buffer
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.rows, self.cols), "typestr": "<f4", "data": (int(self._own.value), False),
                                        "version": 3, "strides": None}
        self._raw = raw
        self.local = torch.as_tensor(raw, device=self.dev)

    def close(self):
        lib = load_library()
        if self._own is None:
            return
        torch.cuda.synchronize(self.dev)
        with torch.cuda.device(self.dev):
            for ptr in self._opened:
                lib.divrec_peer_close(ptr)
            self.local = None
            lib.divrec_peer_free(self._own)
        self._opened, self._own = [], None


def bpr_step_owner(U, V, u, p, n, grad_scale: float, gU: Optional[torch.Tensor], slices: torch.Tensor, slice_rows: int,
                   stats2: torch.Tensor):
    """One-pass BPR step whose item-gradient updates go straight to the owning rank's slice (`slices`: int64 device tensor
    of per-rank slice addresses, PeerBuffer.ptrs).  stats2 (float32 [2], caller-zeroed) += (loss sum, AUC hit count)."""
    lib = load_library()
    U = _cuda(U, "U", torch.float32)
    V = _cuda(V, "V", torch.float32)
    u = _cuda(u, "user_id", torch.int64)
    p = _cuda(p, "positive_item_id", torch.int64)
    n = _cuda(n, "negative_item_id", torch.int64)
    slices = _cuda(slices, "slices", torch.int64)
    stats2 = _cuda(stats2, "stats2", torch.float32)
    with torch.cuda.device(U.device):
        rc = lib.divrec_bpr_step_owner(_ptr(U), _ptr(V), U.shape[0], V.shape[0], _ptr(u), _ptr(p), _ptr(n), u.numel(), U.shape[1],
                                       float(grad_scale), None, None, _ptr(stats2), _ptr(gU), _ptr(slices), slices.numel(),
                                       int(slice_rows), _stream(U.device))
    _check(rc, "divrec_bpr_step_owner")
    return stats2


def bpr_apply_rows(rows: torch.Tensor, p: torch.Tensor, n: torch.Tensor, gV: torch.Tensor) -> None:
    """gV[p_j] += rows_j; gV[n_j] -= rows_j (duplicates accumulate): applies gathered item-gradient rows."""
    lib = load_library()
    rows = _cuda(rows, "rows", torch.float32)
    p = _cuda(p, "positive_item_id", torch.int64)
    n = _cuda(n, "negative_item_id", torch.int64)
    if not gV.is_cuda or not gV.is_contiguous() or gV.dtype != torch.float32:
        raise TypeError("gV must be a contiguous float32 CUDA tensor")
    N, d = rows.shape
    if p.numel() != N or n.numel() != N or gV.shape[1] != d:
        raise ValueError("rows [N,d], p [N], n [N], gV [nI,d] expected")
    with torch.cuda.device(rows.device):
        rc = lib.divrec_bpr_apply_rows(_ptr(rows), _ptr(p), _ptr(n), N, gV.shape[0], d, _ptr(gV), _stream(rows.device))
    _check(rc, "divrec_bpr_apply_rows")


def bpr_sample_triples(pos_rowptr: torch.Tensor, pos_items: torch.Tensor, excl_rowptr: torch.Tensor, excl_items: torch.Tensor,
                       n_items: int, max_sampled: int, seed: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """PairWiseDataset's sampler on the device: (user, positive, negative) int64 [nU * max_sampled^2] in the
    reference's iteration order; rows of users without a positive or a negative are (-1, -1, -1)."""
    lib = load_library()
    pos_rowptr = _cuda(pos_rowptr, "pos_rowptr", torch.int64)
    excl_rowptr = _cuda(excl_rowptr, "excl_rowptr", torch.int64)
    nU = pos_rowptr.numel() - 1
    if excl_rowptr.numel() != nU + 1:
        raise ValueError("pos_rowptr and excl_rowptr must have the same length")
    pos_items = _cuda(pos_items, "pos_items", torch.int32) if pos_items.numel() else pos_items
    excl_items = _cuda(excl_items, "excl_items", torch.int32) if excl_items.numel() else excl_items
    n = nU * max_sampled * max_sampled
    dev = pos_rowptr.device
    out = [torch.empty(n, dtype=torch.int64, device=dev) for _ in range(3)]
    with torch.cuda.device(dev):
        rc = lib.divrec_bpr_sample_triples(_ptr(pos_rowptr), _ptr(pos_items) if pos_items.numel() else None, _ptr(excl_rowptr),
                                           _ptr(excl_items) if excl_items.numel() else None, nU, int(n_items), int(max_sampled),
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream(dev))
    _check(rc, "divrec_bpr_sample_triples")
    return out[0], out[1], out[2]


def bpr_bwd(U, V, u, p, n, pos, neg, grad_scale: float, gU: torch.Tensor, gV: torch.Tensor, g: Optional[torch.Tensor] = None):
    lib = load_library()
    U = _cuda(U, "U", torch.float32)
    V = _cuda(V, "V", torch.float32)
    for name, t in (("gU", gU), ("gV", gV)):
        if t is not None and (not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32):
            raise TypeError(f"{name} must be a contiguous float32 CUDA tensor")
    B = u.numel()
    d = U.shape[1]
    with torch.cuda.device(U.device):
        rc = lib.divrec_bpr_bwd(_ptr(U), _ptr(V), U.shape[0], V.shape[0], _ptr(u), _ptr(p), _ptr(n), _ptr(pos), _ptr(neg), _ptr(g),
                                float(grad_scale), B, d, _ptr(gU), _ptr(gV), _stream(U.device))
    _check(rc, "divrec_bpr_bwd")


def pair_scores(U, V, user_id, item_id) -> torch.Tensor:
    lib = load_library()
    U = _cuda(U, "U", torch.float32)
    V = _cuda(V, "V", torch.float32)
    user_id = _cuda(user_id, "user_id", torch.int64)
    item_id = _cuda(item_id, "item_id", torch.int64)
    if user_id.shape != item_id.shape:
        raise ValueError("user_id and item_id must have the same shape")
    L = user_id.numel()
    out = torch.empty(L, dtype=torch.float32, device=U.device)
    if L == 0:
        return out
    with torch.cuda.device(U.device):
        rc = lib.divrec_pair_scores(_ptr(U), _ptr(V), U.shape[0], V.shape[0], _ptr(user_id), _ptr(item_id), L, U.shape[1], _ptr(out), _stream(U.device))
    _check(rc, "divrec_pair_scores")
    return out


def pair_scores_bwd(U, V, user_id, item_id, g, gU: Optional[torch.Tensor], gV: Optional[torch.Tensor]):
    lib = load_library()
    g = _cuda(g, "grad_output", torch.float32)
    L = user_id.numel()
    if L == 0:
        return
    with torch.cuda.device(U.device):
        rc = lib.divrec_pair_scores_bwd(_ptr(U), _ptr(V), U.shape[0], V.shape[0], _ptr(user_id), _ptr(item_id), _ptr(g), L, U.shape[1],
                                        _ptr(gU), _ptr(gV), _stream(U.device))
    _check(rc, "divrec_pair_scores_bwd")


def ild_labels(labels: torch.Tensor, recs: torch.Tensor) -> torch.Tensor:
    """per-user sum_{a<b} (label[rec_a] == label[rec_b]) / (k (k-1)), fp32 [nU]"""
    lib = load_library()
    recs = _cuda(recs, "recommendations", torch.int64)
    labels = _cuda(labels.to(torch.int64) if labels.dtype != torch.int64 else labels, "item_labels", torch.int64)
    nU, k = recs.shape
    out = torch.empty(nU, dtype=torch.float32, device=recs.device)
    with torch.cuda.device(recs.device):
        rc = lib.divrec_ild_labels(_ptr(labels), labels.numel(), _ptr(recs), nU, k, _ptr(out), _stream(recs.device))
    _check(rc, "divrec_ild_labels")
    return out


def coverage_hist_ranks(recs: torch.Tensor, n_items: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(counts int32 [n_items], rank_sums int64 [n_items])"""
    lib = load_library()
    recs = _cuda(recs, "recommendations", torch.int64)
    nU, k = recs.shape
    counts = torch.zeros(n_items, dtype=torch.int32, device=recs.device)
    sums = torch.zeros(n_items, dtype=torch.int64, device=recs.device)
    with torch.cuda.device(recs.device):
        rc = lib.divrec_coverage_hist_ranks(_ptr(recs), nU, k, n_items, _ptr(counts), _ptr(sums), _stream(recs.device))
    _check(rc, "divrec_coverage_hist_ranks")
    return counts, sums


def relevance_metrics(recs: torch.Tensor, pos_rowptr: torch.Tensor, pos_items: torch.Tensor, raw_counts: torch.Tensor) -> torch.Tensor:
    """fp32 [4, nU]: precision, recall, average precision, ndcg @k"""
    lib = load_library()
    recs = _cuda(recs, "recommendations", torch.int64)
    pos_rowptr = _cuda(pos_rowptr, "pos_rowptr", torch.int64)
    pos_items = _cuda(pos_items, "pos_items", torch.int32)
    raw_counts = _cuda(raw_counts, "raw_counts", torch.int32)
    nU, k = recs.shape
    out = torch.empty((4, nU), dtype=torch.float32, device=recs.device)
    with torch.cuda.device(recs.device):
        rc = lib.divrec_relevance_metrics(_ptr(recs), nU, k, _ptr(pos_rowptr), _ptr(pos_items) if pos_items.numel() else None,
                                          _ptr(raw_counts), _ptr(out), _stream(recs.device))
    _check(rc, "divrec_relevance_metrics")
    return out


def bpr_loss_fwd(pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
    lib = load_library()
    pos = _cuda(pos.detach(), "positives", torch.float32)
    neg = _cuda(neg.detach(), "negatives", torch.float32)
    out = torch.empty_like(pos)
    if pos.numel() == 0:
        return out
    with torch.cuda.device(pos.device):
        rc = lib.divrec_bpr_loss_fwd(_ptr(pos), _ptr(neg), pos.numel(), _ptr(out), None, _stream(pos.device))
    _check(rc, "divrec_bpr_loss_fwd")
    return out


def auc_indicator(pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
    lib = load_library()
    pos = _cuda(pos, "positives", torch.float32)
    neg = _cuda(neg, "negatives", torch.float32)
    out = torch.empty(pos.shape, dtype=torch.int32, device=pos.device)
    if pos.numel() == 0:
        return out
    with torch.cuda.device(pos.device):
        rc = lib.divrec_bpr_loss_fwd(_ptr(pos), _ptr(neg), pos.numel(), None, _ptr(out), _stream(pos.device))
    _check(rc, "divrec_bpr_loss_fwd")
    return out


def bpr_loss_bwd(pos: torch.Tensor, neg: torch.Tensor, g: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = load_library()
    pos = _cuda(pos.detach(), "positives", torch.float32)
    neg = _cuda(neg.detach(), "negatives", torch.float32)
    g = _cuda(g, "grad_output", torch.float32)
    gp, gn = torch.empty_like(pos), torch.empty_like(pos)
    if pos.numel() == 0:
        return gp, gn
    with torch.cuda.device(pos.device):
        rc = lib.divrec_bpr_loss_bwd(_ptr(pos), _ptr(neg), _ptr(g), pos.numel(), _ptr(gp), _ptr(gn), _stream(pos.device))
    _check(rc, "divrec_bpr_loss_bwd")
    return gp, gn


def mmr_rerank(V: torch.Tensor, cand_items: torch.Tensor, cand_scores: torch.Tensor, k: int, lam: float):
    """Greedy MMR re-rank of each user's candidate list: (items [nU,k] int64, scores [nU,k] fp32 of the picks)."""
    lib = load_library()
    V = _cuda(V, "V")
    dt = _table_dtype(V)
    cand_items = _cuda(cand_items, "cand_items", torch.int64)
    cand_scores = _cuda(cand_scores, "cand_scores", torch.float32)
    if cand_items.dim() != 2 or cand_items.shape != cand_scores.shape:
        raise ValueError("cand_items / cand_scores must be [nU, C]")
    nU, C = cand_items.shape
    nI, d = V.shape
    out_i = torch.empty((nU, k), dtype=torch.int64, device=V.device)
    out_s = torch.empty((nU, k), dtype=torch.float32, device=V.device)
    with torch.cuda.device(V.device):
        nbytes = lib.divrec_mmr_workspace(C, d)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=V.device)
        rc = lib.divrec_mmr_rerank(_ptr(V), nI, d, dt, _ptr(cand_items), _ptr(cand_scores), nU, C, k, float(lam),
                                   _ptr(out_i), _ptr(out_s), _ptr(ws), nbytes, _stream(V.device))
    _check(rc, "divrec_mmr_rerank")
    return out_i, out_s
