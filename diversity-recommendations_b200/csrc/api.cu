// api.cu -- the extern "C" boundary (include/divrec_b200.h): argument checks, workspace carving,
// kernel sequencing.  No torch types, no allocation, no device synchronisation.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/divrec_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace divrec {

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return DIVREC_ERR_CUDA;
}

// ---- optional per-kernel timing (bench.py roofline): CUDA events on the launching stream ----
enum ProfSlot { PROF_PREP = 0, PROF_FILTER = 1, PROF_SELECT = 2, PROF_ILD = 3, PROF_COVERAGE = 4, PROF_RESCUE = 5, PROF_SLOTS = 6 };
static std::atomic<int> g_prof_on{0};
static std::atomic<int> g_debug_mode{0};
static std::atomic<long long*> g_filter_stats{nullptr};
static cudaEvent_t g_prof_ev[PROF_SLOTS][2];
static bool g_prof_made = false;
static bool g_prof_set[PROF_SLOTS];
static std::mutex g_prof_mu;
static void prof_mark(int slot, int which, cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_made) {
    for (int i = 0; i < PROF_SLOTS; ++i) { cudaEventCreate(&g_prof_ev[i][0]); cudaEventCreate(&g_prof_ev[i][1]); g_prof_set[i] = false; }
    g_prof_made = true;
  }
  cudaEventRecord(g_prof_ev[slot][which], st);
  if (which == 1) g_prof_set[slot] = true;
}

struct DevInfo { int sm_count = 0; int cc_major = 0; bool ok = false; };
static DevInfo device_info() {
  static std::mutex mu;
  static DevInfo cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DevInfo{};
  std::lock_guard<std::mutex> lk(mu);
  if (!cache[dev].ok) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return DevInfo{};
    cache[dev].sm_count = prop.multiProcessorCount;
    cache[dev].cc_major = prop.major;
    cache[dev].ok = true;
  }
  return cache[dev];
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int64_t CHUNK_USERS = 148 * 8 * 128;  // users per internal pass: 8 user blocks per SM, bounds the workspace

// workspace layout of score_topk (one chunk of users at a time; the item plane is prepared once)
struct WsLayout {
  int d_pad, K, cap;
  bool convA, convB;  // operands have to be rewritten into the workspace (fp32 -> fp16, or bf16 padding)
  int64_t chunk;
  int rescue_grid;
  size_t offA, offB, offNormU, offRowExp, offInvV, offListS, offCounts, offThr, offFail, offScalars, offRescue, offHintMask, total;
  bool two_blocks;
};
static WsLayout ws_layout(int64_t nU, int64_t nI, int d, int dtype, int k) {
  WsLayout w{};
  w.d_pad = (int)align_up((size_t)d, 64);
  w.K = w.d_pad;
  w.convA = w.convB = (dtype == DIVREC_F32) || (d != w.d_pad);
  w.chunk = nU < CHUNK_USERS ? nU : CHUNK_USERS;
  // large k (above MAX_K_SMALL) captures several times more groups per row: the larger lists
  w.cap = (nI >= TIGHTEN_MIN_ITEMS || k > MAX_K_SMALL) ? LIST_CAP_MAX : LIST_CAP;
  const size_t cpad = align_up((size_t)w.chunk, 128);
  size_t off = 0;
  w.offScalars = off; off += 1024;  // u32[0..1] amax bits (U,V), i32[2..3] scale exps, u32[4] max item norm, i32[5] fail count, i32[6..11] statistics
  w.offA = off; if (w.convA) off += align_up(cpad * w.K * 2, 1024);
  w.offB = off; if (w.convB) off += align_up((size_t)nI * w.K * 2, 1024);
  w.offNormU = off; off += align_up(cpad * sizeof(float), 1024);
  w.offRowExp = off; off += align_up(cpad * sizeof(int32_t), 1024);
  w.offInvV = off; off += align_up((size_t)nI * sizeof(float), 1024);
  w.offListS = off; off += align_up(cpad * LISTS_PER_ROW * w.cap * REC_FLOATS * sizeof(float), 1024);
  w.offCounts = off; off += align_up(cpad * LISTS_PER_ROW * sizeof(int32_t), 1024);
  w.offThr = off; off += align_up(cpad * LISTS_PER_ROW * sizeof(float), 1024);
  w.offFail = off; off += align_up(cpad * sizeof(int32_t), 1024);
  w.offRescue = off; off += align_up(rescue_scratch_bytes(nI, &w.rescue_grid), 1024);
  // the common shape (d <= 128, catalogue below the online-tightening threshold) scores two user blocks per CTA and
  // takes its hint on unmasked sample maxima: the rows' seen pattern over the sample tiles
  w.two_blocks = w.K <= 128 && nI < TIGHTEN_MIN_ITEMS;
  w.offHintMask = off; if (w.two_blocks) off += align_up(hint_mask_bytes((int64_t)cpad, (int32_t)nI), 1024);
  w.total = off;
  return w;
}

static int check_score_args(const void* U, const void* V, int64_t nU, int64_t nI, int d, int dtype, int mode) {
  if (!U || !V) return fail(DIVREC_ERR_ARG, "score: null embedding table");
  if (nU <= 0 || nI <= 0 || d <= 0) return fail(DIVREC_ERR_ARG, "score: nU, nI, d must be positive");
  if (nI >= (int64_t)1 << 31) return fail(DIVREC_ERR_ARG, "score: nI must be < 2^31");
  if (dtype != DIVREC_F32 && dtype != DIVREC_BF16) return fail(DIVREC_ERR_ARG, "score: unknown dtype");
  if (mode != DIVREC_MODE_TENSOR && mode != DIVREC_MODE_SIMT) return fail(DIVREC_ERR_ARG, "score: unknown mode");
  if (mode == DIVREC_MODE_TENSOR && d > 512) return fail(DIVREC_ERR_ARG, "score: the tensor-core path supports d <= 512 (use DIVREC_MODE_SIMT)");
  if (d > 8192) return fail(DIVREC_ERR_ARG, "score: d must be <= 8192");
  if ((reinterpret_cast<uintptr_t>(U) & 15) || (reinterpret_cast<uintptr_t>(V) & 15))
    return fail(DIVREC_ERR_ARG, "score: embedding tables must be 16-byte aligned");
  DevInfo di = device_info();
  if (!di.ok) return fail(DIVREC_ERR_CUDA, "score: cannot query the CUDA device");
  if (di.cc_major != 10) return fail(DIVREC_ERR_DEVICE, "score: this library only runs on sm_100 (B200)");
  return DIVREC_OK;
}

}  // namespace divrec

using namespace divrec;

extern "C" {

int divrec_version(void) { return DIVREC_B200_VERSION; }
const char* divrec_last_error(void) { return g_err; }
int64_t divrec_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void divrec_set_profiling(int on) { g_prof_on.store(on ? 1 : 0); }
void divrec_debug_filter_mode(int mode) { g_debug_mode.store(mode); }
void divrec_debug_filter_stats(void* dev_buf) { g_filter_stats.store(reinterpret_cast<long long*>(dev_buf)); }

int divrec_profile_read(float* ms_out, int n) {
  if (!ms_out || n <= 0) return fail(DIVREC_ERR_ARG, "profile_read: bad argument");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < n; ++i) {
    ms_out[i] = -1.0f;
    if (i < PROF_SLOTS && g_prof_made && g_prof_set[i]) {
      if (cudaEventSynchronize(g_prof_ev[i][1]) == cudaSuccess) cudaEventElapsedTime(&ms_out[i], g_prof_ev[i][0], g_prof_ev[i][1]);
    }
  }
  return DIVREC_OK;
}

size_t divrec_score_topk_workspace(int64_t nU, int64_t nI, int d, int dtype, int k) {
  if (nU <= 0 || nI <= 0 || d <= 0) return 0;
  return ws_layout(nU, nI, d, dtype, k).total;
}

static int score_topk_impl(const void* U, const void* V, int64_t nU, int64_t nI, int d, int dtype, int mode,
                           const int64_t* seen_rowptr, const int32_t* seen_items, int64_t item_id_offset, int k,
                           float* out_scores, int64_t* out_items, float* out_ild, int32_t* coverage, void* workspace,
                           size_t ws_bytes, void* stream) {
  int rc = check_score_args(U, V, nU, nI, d, dtype, mode);
  if (rc != DIVREC_OK) return rc;
  if (out_ild != nullptr && ((d & 3) != 0 || d > 256 || k < 2))
    return fail(DIVREC_ERR_ARG, "score_topk_metrics: the fused ILD needs d % 4 == 0, d <= 256 and k >= 2");
  if ((out_ild != nullptr || coverage != nullptr) && item_id_offset != 0)
    return fail(DIVREC_ERR_ARG, "score_topk_metrics: fused metrics need the whole catalogue (item_id_offset == 0)");
  if (k <= 0 || k > MAX_K) return fail(DIVREC_ERR_ARG, "score_topk: k must be in [1,512]");
  if (!out_scores || !out_items) return fail(DIVREC_ERR_ARG, "score_topk: null output");
  if (seen_rowptr == nullptr && seen_items != nullptr)
    return fail(DIVREC_ERR_ARG, "score_topk: seen_items without seen_rowptr");
  const WsLayout w = ws_layout(nU, nI, d, dtype, k);
  if (!workspace || ws_bytes < w.total) return fail(DIVREC_ERR_WORKSPACE, "score_topk: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(DIVREC_ERR_ARG, "score_topk: workspace must be 1024-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (seen_items == nullptr) seen_rowptr = nullptr;
  const bool tensor = mode == DIVREC_MODE_TENSOR;
  const size_t esz = dtype == DIVREC_F32 ? 4 : 2;

  uint32_t* scal = reinterpret_cast<uint32_t*>(ws + w.offScalars);
  uint32_t* amax = scal;                                      // [0] users, [1] items
  int32_t* sexp = reinterpret_cast<int32_t*>(scal + 2);       // [0] users, [1] items
  uint32_t* nvmax = scal + 4;
  int32_t* fail_count = reinterpret_cast<int32_t*>(scal + 5);
  float* norms_u = reinterpret_cast<float*>(ws + w.offNormU);
  float* inv_v = reinterpret_cast<float*>(ws + w.offInvV);
  cudaError_t e = cudaMemsetAsync(scal, 0, 64, st);
  if (e != cudaSuccess) return cuda_fail(e, "memset scalars");

  // item plane: once per call
  const void* B16 = V;
  prof_mark(PROF_PREP, 0, st);
  if (tensor) {
    if (dtype == DIVREC_F32) {
      e = launch_prep_f32(reinterpret_cast<const float*>(V), nI, d, w.d_pad, reinterpret_cast<__half*>(ws + w.offB), amax + 1,
                          sexp + 1, nullptr, inv_v, nvmax, st);
      B16 = ws + w.offB;
    } else {
      e = launch_prep_bf16(V, nI, d, w.d_pad, w.convB ? ws + w.offB : nullptr, nullptr, inv_v, nvmax, st);
      if (w.convB) B16 = ws + w.offB;
    }
    if (e != cudaSuccess) return cuda_fail(e, "prepare item table");
  }

  for (int64_t u0 = 0; u0 < nU; u0 += w.chunk) {
    const int64_t cu = (nU - u0 < w.chunk) ? nU - u0 : w.chunk;
    const uint8_t* Uc = reinterpret_cast<const uint8_t*>(U) + (size_t)u0 * d * esz;
    const int64_t* rp = seen_rowptr ? seen_rowptr + u0 : nullptr;
    float* outs = out_scores + (size_t)u0 * k;
    int64_t* outi = out_items + (size_t)u0 * k;

    SelectParams sp{};
    sp.nU = cu; sp.nI = (int32_t)nI; sp.k = k; sp.d = d; sp.dtype = dtype; sp.cap = w.cap;
    sp.U = Uc; sp.V = V;
    sp.lists = reinterpret_cast<float*>(ws + w.offListS);
    sp.list_count = reinterpret_cast<int32_t*>(ws + w.offCounts);
    sp.thr = reinterpret_cast<float*>(ws + w.offThr);
    sp.seen_rowptr = rp; sp.seen_items = seen_items; sp.item_id_offset = item_id_offset;
    sp.norms_u = norms_u; sp.nvmax_bits = nvmax; sp.inv_norm_v = inv_v;
    sp.out_scores = outs; sp.out_items = outi;
    sp.fail_rows = reinterpret_cast<int32_t*>(ws + w.offFail); sp.fail_count = fail_count;
    sp.fail_reasons = reinterpret_cast<int32_t*>(scal) + 8;
    sp.force_fail = tensor ? 0 : 1;
    sp.id_bytes = rec_id_bytes(nI);
    const bool no_rescue = debug_env("DIVREC_DEBUG_NO_RESCUE") != nullptr;  // debugging aid: leave uncertified rows undone
    sp.debug_dump = no_rescue ? 1 : 0;
    sp.out_ild = out_ild ? out_ild + u0 : nullptr; sp.coverage = coverage;

    if (u0 > 0) {
      e = cudaMemsetAsync(fail_count, 0, sizeof(int32_t), st);
      if (e != cudaSuccess) return cuda_fail(e, "memset fail count");
    }
    if (tensor) {
      const void* A16 = Uc;
      if (u0 > 0) prof_mark(PROF_PREP, 0, st);
      if (dtype == DIVREC_F32) {
        int32_t* row_exps = reinterpret_cast<int32_t*>(ws + w.offRowExp);
        e = launch_prep_f32_rows(reinterpret_cast<const float*>(Uc), cu, d, w.d_pad, reinterpret_cast<__half*>(ws + w.offA), row_exps,
                                 norms_u, st);
        A16 = ws + w.offA;
        sp.scale_exps = sexp; sp.amax_bits = amax; sp.row_exps = row_exps;
        sp.c_margin = 1.1f / 1024.0f;
      } else {
        e = launch_prep_bf16(Uc, cu, d, w.d_pad, w.convA ? ws + w.offA : nullptr, norms_u, nullptr, nullptr, st);
        if (w.convA) A16 = ws + w.offA;
        sp.scale_exps = nullptr; sp.amax_bits = nullptr; sp.row_exps = nullptr;
        sp.c_margin = 1.0f / 4096.0f;
      }
      if (e != cudaSuccess) return cuda_fail(e, "prepare user table");
      prof_mark(PROF_PREP, 1, st);

      FilterParams fp{};
      fp.nU = cu; fp.nI = (int32_t)nI; fp.kblocks = w.K / 64; fp.fmt = dtype == DIVREC_F32 ? 0 : 1; fp.k = k;
      fp.use_hint = 1; fp.cap = w.cap; fp.seen_rowptr = rp; fp.debug_mode = g_debug_mode.load(); fp.stats = g_filter_stats.load();
      fp.lists = reinterpret_cast<float*>(ws + w.offListS);
      fp.list_count = reinterpret_cast<int32_t*>(ws + w.offCounts);
      fp.thr_out = reinterpret_cast<float*>(ws + w.offThr);
      fp.norms_u = norms_u; fp.nvmax_bits = nvmax; fp.amax_bits = sp.amax_bits; fp.scale_exps = sp.scale_exps; fp.row_exps = sp.row_exps;
      fp.c_margin = sp.c_margin; fp.d = d;
      const char* err = "";
      prof_mark(PROF_FILTER, 0, st);
      // the common shape (d <= 128, catalogue below the online-tightening threshold) scores two user blocks per CTA
      const bool two_blocks = w.two_blocks && !debug_env("DIVREC_DEBUG_ONE_BLOCK");
      sp.merged_lists = two_blocks ? 1 : 0;
      if (two_blocks && rp != nullptr && !debug_env("DIVREC_DEBUG_NO_HINT_MASK")) {
        unsigned long long* hmask = reinterpret_cast<unsigned long long*>(ws + w.offHintMask);
        const int64_t rows_padded = (int64_t)align_up((size_t)w.chunk, 128);
        e = launch_hint_mask(rp, seen_items, cu, rows_padded, (int32_t)nI, item_id_offset, hmask, st);
        if (e != cudaSuccess) return cuda_fail(e, "hint mask launch");
        fp.hint_mask = hmask; fp.hint_mask_stride = rows_padded;
      }
      e = two_blocks ? launch_score_filter2(A16, B16, fp, device_info().sm_count, st, &err)
                     : launch_score_filter(A16, B16, fp, device_info().sm_count, st, &err);
      if (e != cudaSuccess) return cuda_fail(e, err[0] ? err : "score_filter launch");
      prof_mark(PROF_FILTER, 1, st);
    } else {
      // DIVREC_MODE_SIMT: no candidate lists; select only enqueues every row for the exact kernel
      e = cudaMemsetAsync(ws + w.offCounts, 0, align_up((size_t)cu, 128) * LISTS_PER_ROW * sizeof(int32_t), st);
      if (e != cudaSuccess) return cuda_fail(e, "memset counts");
    }
    prof_mark(PROF_SELECT, 0, st);
    e = launch_select_topk(sp, st);
    if (e != cudaSuccess) return cuda_fail(e, "select_topk launch");
    prof_mark(PROF_SELECT, 1, st);

    RescueParams rp2{};
    rp2.nU = cu; rp2.nI = (int32_t)nI; rp2.k = k; rp2.d = d; rp2.dtype = dtype; rp2.U = Uc; rp2.V = V;
    rp2.seen_rowptr = rp; rp2.seen_items = seen_items; rp2.item_id_offset = item_id_offset;
    rp2.fail_rows = sp.fail_rows; rp2.fail_count = fail_count;
    rp2.scratch = reinterpret_cast<uint32_t*>(ws + w.offRescue); rp2.grid = w.rescue_grid;
    rp2.out_scores = outs; rp2.out_items = outi; rp2.out_ild = sp.out_ild; rp2.coverage = coverage;
    rp2.total_fail = reinterpret_cast<int32_t*>(scal) + 6;  // statistics: rows re-done exactly
    prof_mark(PROF_RESCUE, 0, st);
    e = no_rescue ? cudaSuccess : launch_rescue(rp2, st);  // (a debug build that skips it also skips the statistics)
    if (e != cudaSuccess) return cuda_fail(e, "rescue launch");
    prof_mark(PROF_RESCUE, 1, st);
  }
  return DIVREC_OK;
}

int divrec_score_topk(const void* U, const void* V, int64_t nU, int64_t nI, int d, int dtype, int mode,
                      const int64_t* seen_rowptr, const int32_t* seen_items, int64_t item_id_offset, int k,
                      float* out_scores, int64_t* out_items, void* workspace, size_t ws_bytes, void* stream) {
  return score_topk_impl(U, V, nU, nI, d, dtype, mode, seen_rowptr, seen_items, item_id_offset, k, out_scores, out_items,
                         nullptr, nullptr, workspace, ws_bytes, stream);
}

int divrec_score_topk_metrics(const void* U, const void* V, int64_t nU, int64_t nI, int d, int dtype, int mode,
                              const int64_t* seen_rowptr, const int32_t* seen_items, int k, float* out_scores,
                              int64_t* out_items, float* out_ild, int32_t* coverage_counts, void* workspace,
                              size_t ws_bytes, void* stream) {
  return score_topk_impl(U, V, nU, nI, d, dtype, mode, seen_rowptr, seen_items, 0, k, out_scores, out_items, out_ild,
                         coverage_counts, workspace, ws_bytes, stream);
}

int divrec_score_topk_rescued(const void* workspace, int64_t nU, int64_t nI, int d, int dtype, int32_t* host_out, void* stream) {
  if (!workspace || !host_out || nU <= 0 || nI <= 0 || d <= 0) return fail(DIVREC_ERR_ARG, "score_topk_rescued: bad argument");
  (void)dtype;
  const size_t off_scalars = 0;  // the scalar block leads the workspace whatever the shape (ws_layout)
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // host_out[0] = rows re-done exactly; host_out[1..4] = by reason: list overflow / forced, too many candidates,
  // hint not certified, too few captured
  int32_t tmp[6];
  cudaError_t e = cudaMemcpyAsync(tmp, reinterpret_cast<const uint8_t*>(workspace) + off_scalars + 6 * sizeof(int32_t),
                                  6 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) { host_out[0] = tmp[0]; for (int i = 0; i < 4; ++i) host_out[1 + i] = tmp[2 + i]; }
  if (e != cudaSuccess) return cuda_fail(e, "score_topk_rescued");
  return DIVREC_OK;
}

int divrec_topk_merge(const float* gathered_scores, const int64_t* gathered_items, int n_shards, int64_t nU, int k,
                      float* out_scores, int64_t* out_items, void* stream) {
  if (!gathered_scores || !gathered_items || !out_scores || !out_items) return fail(DIVREC_ERR_ARG, "topk_merge: null pointer");
  if (n_shards <= 0 || nU <= 0 || k <= 0) return fail(DIVREC_ERR_ARG, "topk_merge: n_shards, nU, k must be positive");
  cudaError_t e = launch_topk_merge(gathered_scores, gathered_items, n_shards, nU, k, out_scores, out_items,
                                    reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "topk_merge launch");
  return DIVREC_OK;
}

int divrec_topk_pack(const float* scores, const int64_t* items, int64_t n, uint64_t* out_packed, void* stream) {
  if (!scores || !items || !out_packed) return fail(DIVREC_ERR_ARG, "topk_pack: null pointer");
  if (n <= 0) return fail(DIVREC_ERR_ARG, "topk_pack: n must be positive");
  cudaError_t e = launch_topk_pack(scores, items, n, out_packed, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "topk_pack launch");
  return DIVREC_OK;
}

int divrec_topk_merge_packed(const uint64_t* gathered, int n_shards, int64_t nU, int k, float* out_scores,
                             int64_t* out_items, void* stream) {
  if (!gathered || !out_scores || !out_items) return fail(DIVREC_ERR_ARG, "topk_merge_packed: null pointer");
  if (n_shards <= 0 || nU <= 0 || k <= 0) return fail(DIVREC_ERR_ARG, "topk_merge_packed: n_shards, nU, k must be positive");
  if ((size_t)n_shards * k > 6144) return fail(DIVREC_ERR_ARG, "topk_merge_packed: n_shards * k must be <= 6144 (use divrec_topk_merge)");
  cudaError_t e = launch_topk_merge_packed(gathered, n_shards, nU, k, out_scores, out_items, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "topk_merge_packed launch");
  return DIVREC_OK;
}

int divrec_ild_cosine(const void* V, int64_t nI, int d, int dtype, const int64_t* recs, int64_t nU, int k,
                      int ild_mode, float* inv_norm, float* out_per_user, void* stream) {
  if (!V || !recs || !inv_norm || !out_per_user) return fail(DIVREC_ERR_ARG, "ild_cosine: null pointer");
  if (nI <= 0 || nU <= 0 || k < 2) return fail(DIVREC_ERR_ARG, "ild_cosine: need nI > 0, nU > 0, k >= 2");
  if (d <= 0 || (d & 3) || d > 1024) return fail(DIVREC_ERR_ARG, "ild_cosine: d must be a multiple of 4, <= 1024");
  if (dtype != DIVREC_F32 && dtype != DIVREC_BF16) return fail(DIVREC_ERR_ARG, "ild_cosine: unknown dtype");
  if (reinterpret_cast<uintptr_t>(V) & 15) return fail(DIVREC_ERR_ARG, "ild_cosine: V must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  prof_mark(PROF_ILD, 0, st);
  cudaError_t e = launch_inv_norm(V, nI, d, dtype, inv_norm, st);
  if (e != cudaSuccess) return cuda_fail(e, "inv_norm launch");
  if (ild_mode == DIVREC_ILD_SUMVEC) e = launch_ild_sumvec(V, nI, d, dtype, recs, nU, k, inv_norm, out_per_user, st);
  else if (ild_mode == DIVREC_ILD_PAIRWISE) e = launch_ild_pairwise(V, nI, d, dtype, recs, nU, k, inv_norm, out_per_user, st);
  else return fail(DIVREC_ERR_ARG, "ild_cosine: unknown ild_mode");
  if (e != cudaSuccess) return cuda_fail(e, "ild launch");
  prof_mark(PROF_ILD, 1, st);
  return DIVREC_OK;
}

int divrec_ild_matrix(const void* D, int d_is_int32, int64_t nI, const int64_t* recs, int64_t nU, int k,
                      float* out_per_user, void* stream) {
  if (!D || !recs || !out_per_user) return fail(DIVREC_ERR_ARG, "ild_matrix: null pointer");
  if (nI <= 0 || nU <= 0 || k < 2) return fail(DIVREC_ERR_ARG, "ild_matrix: need nI > 0, nU > 0, k >= 2");
  cudaError_t e = launch_ild_matrix(D, d_is_int32, nI, recs, nU, k, out_per_user, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ild_matrix launch");
  return DIVREC_OK;
}

int divrec_coverage_hist(const int64_t* recs, int64_t n, int64_t nI, int32_t* counts, void* stream) {
  if (!recs || !counts) return fail(DIVREC_ERR_ARG, "coverage_hist: null pointer");
  if (n < 0 || nI <= 0) return fail(DIVREC_ERR_ARG, "coverage_hist: bad sizes");
  if (n == 0) return DIVREC_OK;
  if (reinterpret_cast<uintptr_t>(recs) & 15) return fail(DIVREC_ERR_ARG, "coverage_hist: recs must be 16-byte aligned");
  prof_mark(PROF_COVERAGE, 0, reinterpret_cast<cudaStream_t>(stream));
  cudaError_t e = launch_coverage_hist(recs, n, nI, counts, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "coverage_hist launch");
  prof_mark(PROF_COVERAGE, 1, reinterpret_cast<cudaStream_t>(stream));
  return DIVREC_OK;
}

int divrec_entropy_from_counts(const int32_t* counts, int64_t nI, float* out2, void* stream) {
  if (!counts || !out2 || nI <= 0) return fail(DIVREC_ERR_ARG, "entropy_from_counts: bad argument");
  cudaError_t e = launch_entropy(counts, nI, out2, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "entropy launch");
  return DIVREC_OK;
}

static int check_rows(const float* U, const float* V, int64_t nU, int64_t nI, int d, const char* who) {
  if (!U || !V) return fail(DIVREC_ERR_ARG, "%s: null embedding table", who);
  if (nU <= 0 || nI <= 0) return fail(DIVREC_ERR_ARG, "%s: nU and nI must be positive", who);
  if (d <= 0 || d > 1024) return fail(DIVREC_ERR_ARG, "%s: d must be in [1, 1024]", who);
  return DIVREC_OK;
}
static bool vec_ok(int d, const void* a, const void* b = nullptr, const void* c = nullptr) {
  return (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

int divrec_bpr_fwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p, const int64_t* n,
                   int64_t B, int d, float* pos, float* neg, float* loss_sum, int32_t* auc_sum, void* stream) {
  int rc = check_rows(U, V, nU, nI, d, "bpr_fwd");
  if (rc != DIVREC_OK) return rc;
  if (!u || !p || !n || !pos || !neg) return fail(DIVREC_ERR_ARG, "bpr_fwd: null pointer");
  if (B <= 0) return fail(DIVREC_ERR_ARG, "bpr_fwd: B must be positive");
  cudaError_t e = launch_bpr_fwd(U, V, nU, nI, u, p, n, B, d, pos, neg, loss_sum, auc_sum, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bpr_fwd launch");
  return DIVREC_OK;
}

int divrec_bpr_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p, const int64_t* n,
                   const float* pos, const float* neg, const float* g, float grad_scale, int64_t B, int d, float* gU,
                   float* gV, void* stream) {
  int rc = check_rows(U, V, nU, nI, d, "bpr_bwd");
  if (rc != DIVREC_OK) return rc;
  if (!u || !p || !n) return fail(DIVREC_ERR_ARG, "bpr_bwd: null index pointer");
  if (!g && (!pos || !neg)) return fail(DIVREC_ERR_ARG, "bpr_bwd: need g or (pos,neg)");
  if (B <= 0) return fail(DIVREC_ERR_ARG, "bpr_bwd: B must be positive");
  cudaError_t e = launch_bpr_bwd(U, V, nU, nI, u, p, n, pos, neg, g, grad_scale, B, d, gU, gV, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bpr_bwd launch");
  return DIVREC_OK;
}

static int bpr_step_impl(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                         const int64_t* n, int64_t B, int d, float grad_scale, float* pos, float* neg, float* loss_sum,
                         int32_t* auc_sum, float* gU, float* gV, float* item_rows, float* const* slices, int64_t slice_rows,
                         float* auc_f, void* stream) {
  int rc = check_rows(U, V, nU, nI, d, "bpr_step");
  if (rc != DIVREC_OK) return rc;
  if (!u || !p || !n) return fail(DIVREC_ERR_ARG, "bpr_step: null index pointer");
  if (B <= 0) return fail(DIVREC_ERR_ARG, "bpr_step: B must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!vec_ok(d, U, V) || !vec_ok(d, gU, gV, item_rows)) {
    // rows that cannot move as 128-bit vectors (d % 4 != 0 or unaligned tables): forward, then backward (two passes)
    if (slices || item_rows) return fail(DIVREC_ERR_ARG, "bpr_step: the sparse / owner-sharded item gradient needs d % 4 == 0 and 16-byte aligned tables");
    if (!pos || !neg) return fail(DIVREC_ERR_ARG, "bpr_step: pos and neg are required when d % 4 != 0 (two-pass form)");
    cudaError_t e = launch_bpr_fwd(U, V, nU, nI, u, p, n, B, d, pos, neg, loss_sum, auc_sum, st);
    if (e == cudaSuccess) e = launch_bpr_bwd(U, V, nU, nI, u, p, n, pos, neg, nullptr, grad_scale, B, d, gU, gV, st);
    if (e != cudaSuccess) return cuda_fail(e, "bpr_step launch");
    return DIVREC_OK;
  }
  cudaError_t e = launch_bpr_step(U, V, nU, nI, u, p, n, B, d, grad_scale, pos, neg, loss_sum, auc_sum, gU, gV, item_rows, slices,
                                  slice_rows, auc_f, st);
  if (e != cudaSuccess) return cuda_fail(e, "bpr_step launch");
  return DIVREC_OK;
}

int divrec_bpr_step(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p, const int64_t* n,
                    int64_t B, int d, float grad_scale, float* pos, float* neg, float* loss_sum, int32_t* auc_sum, float* gU,
                    float* gV, float* item_rows, void* stream) {
  return bpr_step_impl(U, V, nU, nI, u, p, n, B, d, grad_scale, pos, neg, loss_sum, auc_sum, gU, gV, item_rows, nullptr, 1, nullptr, stream);
}

int divrec_bpr_step_owner(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                          const int64_t* n, int64_t B, int d, float grad_scale, float* pos, float* neg, float* stats2,
                          float* gU, float* const* gV_slices, int n_slices, int64_t slice_rows, void* stream) {
  if (!gV_slices || n_slices <= 0 || slice_rows <= 0) return fail(DIVREC_ERR_ARG, "bpr_step_owner: need the slice table, n_slices > 0 and slice_rows > 0");
  if ((int64_t)n_slices * slice_rows < nI) return fail(DIVREC_ERR_ARG, "bpr_step_owner: n_slices * slice_rows must cover nI");
  return bpr_step_impl(U, V, nU, nI, u, p, n, B, d, grad_scale, pos, neg, stats2, nullptr, gU, nullptr, nullptr, gV_slices, slice_rows,
                       stats2 ? stats2 + 1 : nullptr, stream);
}

int divrec_bad_index_count(int64_t* host_out, int reset, void* stream) {
  if (!host_out) return fail(DIVREC_ERR_ARG, "bad_index_count: null pointer");
  unsigned long long v = 0;
  cudaError_t e = read_bad_index_count(&v, reset, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bad_index_count");
  *host_out = (int64_t)v;
  return DIVREC_OK;
}

// ---- peer buffers (CUDA IPC): device memory another process of the same node can map ----
int divrec_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out) {
  if (!ptr_out || !handle_out || bytes == 0) return fail(DIVREC_ERR_ARG, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == DIVREC_PEER_HANDLE_BYTES, "handle size");
  void* ptr = nullptr;
  cudaError_t e = cudaMalloc(&ptr, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "peer_alloc: cudaMalloc");
  cudaIpcMemHandle_t hd;
  e = cudaIpcGetMemHandle(&hd, ptr);
  if (e != cudaSuccess) { cudaFree(ptr); return cuda_fail(e, "peer_alloc: cudaIpcGetMemHandle"); }
  memcpy(handle_out, &hd, sizeof(hd));
  *ptr_out = ptr;
  return DIVREC_OK;
}
int divrec_peer_open(const unsigned char* handle, void** ptr_out) {
  if (!handle || !ptr_out) return fail(DIVREC_ERR_ARG, "peer_open: bad argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, sizeof(hd));
  cudaError_t e = cudaIpcOpenMemHandle(ptr_out, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "peer_open: cudaIpcOpenMemHandle");
  return DIVREC_OK;
}
int divrec_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) return cuda_fail(e, "peer_close");
  return DIVREC_OK;
}
int divrec_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) return cuda_fail(e, "peer_free");
  return DIVREC_OK;
}

int divrec_bpr_sample_triples(const int64_t* pos_rowptr, const int32_t* pos_items, const int64_t* excl_rowptr,
                              const int32_t* excl_items, int64_t nU, int64_t nI, int max_sampled, uint64_t seed,
                              int64_t* out_u, int64_t* out_p, int64_t* out_n, void* stream) {
  if (!pos_rowptr || !excl_rowptr || !out_u || !out_p || !out_n) return fail(DIVREC_ERR_ARG, "bpr_sample_triples: null pointer");
  if (nU <= 0 || nI <= 0) return fail(DIVREC_ERR_ARG, "bpr_sample_triples: nU, nI must be positive");
  if (max_sampled <= 0 || max_sampled > 256) return fail(DIVREC_ERR_ARG, "bpr_sample_triples: max_sampled must be in [1, 256]");
  cudaError_t e = launch_bpr_sample(pos_rowptr, pos_items, excl_rowptr, excl_items, nU, nI, max_sampled, seed, out_u, out_p,
                                    out_n, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bpr_sample_triples launch");
  return DIVREC_OK;
}

int divrec_bpr_apply_rows(const float* rows, const int64_t* p, const int64_t* n, int64_t N, int64_t nI, int d, float* gV, void* stream) {
  if (!rows || !p || !n || !gV) return fail(DIVREC_ERR_ARG, "bpr_apply_rows: null pointer");
  if (N <= 0 || nI <= 0) return fail(DIVREC_ERR_ARG, "bpr_apply_rows: N and nI must be positive");
  if (d <= 0 || (d & 3) || d > 1024) return fail(DIVREC_ERR_ARG, "bpr_apply_rows: d must be a multiple of 4, <= 1024");
  if ((reinterpret_cast<uintptr_t>(rows) & 15) || (reinterpret_cast<uintptr_t>(gV) & 15))
    return fail(DIVREC_ERR_ARG, "bpr_apply_rows: rows and gV must be 16-byte aligned");
  cudaError_t e = launch_bpr_apply_rows(rows, p, n, N, nI, d, gV, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bpr_apply_rows launch");
  return DIVREC_OK;
}

int divrec_pair_scores(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* user_id, const int64_t* item_id,
                       int64_t L, int d, float* out, void* stream) {
  int rc = check_rows(U, V, nU, nI, d, "pair_scores");
  if (rc != DIVREC_OK) return rc;
  if (!user_id || !item_id || !out) return fail(DIVREC_ERR_ARG, "pair_scores: null pointer");
  if (L <= 0) return fail(DIVREC_ERR_ARG, "pair_scores: L must be positive");
  cudaError_t e = launch_pair_scores(U, V, nU, nI, user_id, item_id, L, d, out, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "pair_scores launch");
  return DIVREC_OK;
}

int divrec_pair_scores_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* user_id, const int64_t* item_id,
                           const float* g, int64_t L, int d, float* gU, float* gV, void* stream) {
  int rc = check_rows(U, V, nU, nI, d, "pair_scores_bwd");
  if (rc != DIVREC_OK) return rc;
  if (!user_id || !item_id || !g) return fail(DIVREC_ERR_ARG, "pair_scores_bwd: null pointer");
  if (L <= 0) return fail(DIVREC_ERR_ARG, "pair_scores_bwd: L must be positive");
  cudaError_t e = launch_pair_scores_bwd(U, V, nU, nI, user_id, item_id, g, L, d, gU, gV, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "pair_scores_bwd launch");
  return DIVREC_OK;
}

int divrec_ild_labels(const int64_t* labels, int64_t nI, const int64_t* recs, int64_t nU, int k, float* out_per_user,
                      void* stream) {
  if (!labels || !recs || !out_per_user) return fail(DIVREC_ERR_ARG, "ild_labels: null pointer");
  if (nI <= 0 || nU <= 0 || k < 2) return fail(DIVREC_ERR_ARG, "ild_labels: need nI > 0, nU > 0, k >= 2");
  cudaError_t e = launch_ild_labels(labels, nI, recs, nU, k, out_per_user, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ild_labels launch");
  return DIVREC_OK;
}

int divrec_coverage_hist_ranks(const int64_t* recs, int64_t nU, int k, int64_t nI, int32_t* counts, uint64_t* rank_sums,
                               void* stream) {
  if (!recs || !counts || !rank_sums) return fail(DIVREC_ERR_ARG, "coverage_hist_ranks: null pointer");
  if (nU < 0 || k <= 0 || nI <= 0) return fail(DIVREC_ERR_ARG, "coverage_hist_ranks: bad sizes");
  if (nU == 0) return DIVREC_OK;
  cudaError_t e = launch_coverage_hist_ranks(recs, nU, k, nI, counts, reinterpret_cast<unsigned long long*>(rank_sums),
                                             reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "coverage_hist_ranks launch");
  return DIVREC_OK;
}

int divrec_relevance_metrics(const int64_t* recs, int64_t nU, int k, const int64_t* pos_rowptr, const int32_t* pos_items,
                             const int32_t* raw_counts, float* out, void* stream) {
  if (!recs || !pos_rowptr || !raw_counts || !out) return fail(DIVREC_ERR_ARG, "relevance_metrics: null pointer");
  if (nU <= 0 || k <= 0) return fail(DIVREC_ERR_ARG, "relevance_metrics: nU, k must be positive");
  cudaError_t e = launch_relevance(recs, nU, k, pos_rowptr, pos_items, raw_counts, out, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "relevance_metrics launch");
  return DIVREC_OK;
}

int divrec_bpr_loss_fwd(const float* pos, const float* neg, int64_t B, float* loss_values, int32_t* hits, void* stream) {
  if (!pos || !neg) return fail(DIVREC_ERR_ARG, "bpr_loss_fwd: null pointer");
  if (B <= 0) return fail(DIVREC_ERR_ARG, "bpr_loss_fwd: B must be positive");
  cudaError_t e = launch_bpr_loss_fwd(pos, neg, B, loss_values, hits, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bpr_loss_fwd launch");
  return DIVREC_OK;
}

int divrec_bpr_loss_bwd(const float* pos, const float* neg, const float* g, int64_t B, float* gpos, float* gneg,
                        void* stream) {
  if (!pos || !neg || !g || !gpos || !gneg) return fail(DIVREC_ERR_ARG, "bpr_loss_bwd: null pointer");
  if (B <= 0) return fail(DIVREC_ERR_ARG, "bpr_loss_bwd: B must be positive");
  cudaError_t e = launch_bpr_loss_bwd(pos, neg, g, B, gpos, gneg, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bpr_loss_bwd launch");
  return DIVREC_OK;
}

size_t divrec_build_csr_workspace(int64_t nnz, int64_t nU) {
  if (nnz < 0 || nU <= 0) return 0;
  return build_csr_workspace_bytes(nnz, nU);
}

int divrec_build_csr(const int64_t* pairs, int64_t nnz, int64_t nU, int64_t nI, int64_t* rowptr, int32_t* items, void* workspace,
                     size_t ws_bytes, void* stream) {
  if (!rowptr || (nnz > 0 && (!pairs || !items))) return fail(DIVREC_ERR_ARG, "build_csr: null pointer");
  if (nnz < 0 || nU <= 0 || nI <= 0 || nI >= ((int64_t)1 << 31)) return fail(DIVREC_ERR_ARG, "build_csr: need nnz >= 0, nU > 0, 0 < nI < 2^31");
  if (reinterpret_cast<uintptr_t>(pairs) & 15) return fail(DIVREC_ERR_ARG, "build_csr: pairs must be 16-byte aligned");
  if (!workspace || ws_bytes < build_csr_workspace_bytes(nnz, nU)) return fail(DIVREC_ERR_WORKSPACE, "build_csr: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(DIVREC_ERR_ARG, "build_csr: workspace must be 256-byte aligned");
  cudaError_t e = launch_build_csr(pairs, nnz, nU, nI, rowptr, items, workspace, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "build_csr launch");
  return DIVREC_OK;
}

size_t divrec_mmr_workspace(int C, int d) {
  if (C <= 0 || d <= 0) return 0;
  return mmr_scratch_bytes(C, d, nullptr);
}

int divrec_mmr_rerank(const void* V, int64_t nI, int d, int dtype, const int64_t* cand_items, const float* cand_scores,
                      int64_t nU, int C, int k, float lambda, int64_t* out_items, float* out_scores, void* workspace,
                      size_t ws_bytes, void* stream) {
  if (!V || !cand_items || !cand_scores || !out_items) return fail(DIVREC_ERR_ARG, "mmr_rerank: null pointer");
  if (nI <= 0 || nU <= 0 || d <= 0 || C <= 0 || k <= 0 || k > C) return fail(DIVREC_ERR_ARG, "mmr_rerank: need nI, nU, d > 0 and 0 < k <= C");
  if (dtype != DIVREC_F32 && dtype != DIVREC_BF16) return fail(DIVREC_ERR_ARG, "mmr_rerank: unknown dtype");
  if (!(lambda >= 0.0f && lambda <= 1.0f)) return fail(DIVREC_ERR_ARG, "mmr_rerank: lambda must be in [0, 1]");
  if ((size_t)((d + 3) & ~3) * 4 + (size_t)C * 12 > 200 * 1024) return fail(DIVREC_ERR_ARG, "mmr_rerank: C and d too large for shared memory");
  if (!workspace || ws_bytes < mmr_scratch_bytes(C, d, nullptr)) return fail(DIVREC_ERR_WORKSPACE, "mmr_rerank: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return fail(DIVREC_ERR_ARG, "mmr_rerank: workspace must be 16-byte aligned");
  MmrParams mp{};
  mp.V = V; mp.nI = nI; mp.d = d; mp.dtype = dtype; mp.cand_items = cand_items; mp.cand_scores = cand_scores; mp.nU = nU;
  mp.C = C; mp.k = k; mp.lambda = lambda; mp.scratch = reinterpret_cast<float*>(workspace); mp.out_items = out_items;
  mp.out_scores = out_scores;
  cudaError_t e = launch_mmr(mp, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "mmr launch");
  return DIVREC_OK;
}

}  // extern "C"
