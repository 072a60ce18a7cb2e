// C ABI (include/qforest_b200.h): forest residency in HBM, launch orchestration, and the
// host-buffer end-to-end path.  All compute is in the kernels of apply_kernel.cuh and
// quantile_kernel.cuh; there is no CPU fallback anywhere in this file.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "apply_kernel.cuh"
#include "qf_internal.h"
#include "quantile_kernel.cuh"
#include "select_kernel.cuh"
#include "select2_kernel.cuh"
#include "compact_lists.cuh"
#include "select3_kernel.cuh"
#include "select4_kernel.cuh"
#include "csr_builder.cuh"

namespace qf {

std::string& last_error_slot() {
    static thread_local std::string slot;
    return slot;
}

}  // namespace qf

namespace {

constexpr int64_t kAlign = 256;
inline int64_t align_up(int64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Host-buffer path: three streams (H2D, kernels, D2H) over a ring of kPipeSlots chunk buffers.
// The copies of one chunk run while the kernels of its neighbours do; the kernels themselves
// stay on ONE stream, chunk after chunk, and share one workspace.
constexpr int kPipeSlots = 4;
struct HostPipe {
    cudaStream_t h2d = nullptr, compute = nullptr, d2h = nullptr;
    float* x[kPipeSlots] = {};
    double* out[kPipeSlots] = {};
    cudaEvent_t x_ready[kPipeSlots] = {}, computed[kPipeSlots] = {}, out_free[kPipeSlots] = {};
    int64_t x_bytes = 0, out_bytes = 0;     // per slot
    void* ws = nullptr;
    int64_t ws_bytes = 0;
    // page-locked staging for callers that hand over pageable memory (an ordinary numpy array): a
    // pageable cudaMemcpyAsync blocks the calling thread until the copy is done, and a pageable D2H
    // even until the chunk's kernels are, which would serialise copies and kernels
    float* hx[kPipeSlots] = {};
    double* hout[kPipeSlots] = {};
    int64_t hx_bytes = 0, hout_bytes = 0;   // per slot
};

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace

struct qf_forest {
    int device = 0;
    int T = 0, F = 0, fbits = 0;
    int64_t N = 0, n_nodes = 0, n_members = 0, max_row_members = 0;
    double expected_row_members = 0.0;
    uint2* d_nodes = nullptr;
    uint32_t* d_roots = nullptr;
    int32_t* d_node_ids = nullptr;
    uint2* d_members = nullptr;
    float* d_y_sorted = nullptr;
    double* d_values = nullptr;
    double* d_impurity = nullptr;          // tree_.impurity per packed node (qf_forest_set_leaf_impurity)
    int64_t device_bytes = 0;
    std::vector<uint32_t> tree_max_leaf;   // members of the largest leaf of every tree
    bool unit_leaves = false;              // every leaf of every tree holds exactly one member
    int64_t row_weight_bound = 0;          // upper bound on the total weight one row can gather (T for normalised leaf weights)
    int sm_count = 148;
    // options (0 = auto where noted)
    int apply_row_warps = 0;      // warps along rows per CTA (0 auto)
    int apply_ilp = 4;
    int apply_levels = qf::kApplyLevels;   // visits between two warp votes (1, 2, 3, 4; ILP 4 only)
    int sort_rows = -1;           // walk rows in locality order: -1 auto, 0 never, 1 always
    int key_shift = 0, key_bins = 1;   // locality buckets = first member of the leaf in tree 0 >> key_shift
    int key_hist = qf::kKeyBins;       // histogram counters (>= key_bins; option "key_bins")
    uint32_t root0 = 0;
    int64_t rows_per_chunk = 262144;    // rows per kernel chunk (denser chunks = closer neighbours in the locality order: C3-quarter K1 2.66 -> 2.52 ms against 131072)
    int cta_capacity = 0;         // cap on the members per row of the first bucket stage (0 = its capacity)
    int select_bits = qf::kSelectL2Bits;   // log2 refinement slots of the selection kernel (tests lower it)
    int select4_hits = qf::kSelect4Hits;   // hit-list entries a row may use before select4 narrows instead (tests shrink it)
    int select4_ctas = 4;                  // CTAs (of 8 rows) per SM select4 is launched with, at most
    int compact_align = 0;                 // 1: compact lists of two or more octets start on an even octet (before the lists are built)
    int select4_hints = 0;                 // 1: L2 eviction hints on select4's octet loads (keep after pass 1, drop after pass 2)
    int select4_unroll = 2;                // octets a lane keeps in flight per step of a pass (2 or 4)
    int select4_list = 0;                  // entries of select4's per-row octet list (0 = auto; tests shrink it)
    int select3_capw = 0;                  // octets per warp segment of the select3 stage (0 = auto; tests shrink it)
    int select3_variant = 1;               // 0: 3 select warps, 2 stages, 2 CTAs per SM; 1: 1 select warp, 1 stage, 3 CTAs per SM
    int select_impl = 0;                   // 0: auto (select2 when the training set allows, else the multi-level kernel);
                                           // 1: always the multi-level kernel; 3 / 4: select3 (CTA per row) / select4 (warp per
                                           // row) on compact leaf lists when the forest and the call allow (<= 3 percentiles),
                                           // else as 0
    int cta2_capacity = 8192;     // same for the 512-thread overflow stage (0 = stage not used)
    int warp_elems = -1;          // warp kernel: keys per lane (2, 4, 8), 0 = never, -1 = auto
    int warp_unit = 1;            // 0: never take the one-member-per-leaf gather of the warp kernel
    int dense_ctas = 0;           // 0 = auto
    // compact leaf lists (compact_lists.cuh), built on the first QF_MODE_FAST call that can use them
    std::mutex compact_mu;
    int compact_state = 0;                 // 0: not tried, 1: ready, -1: this forest is not compactable
    uint2* d_nodes_c = nullptr;            // copy of d_nodes whose leaves carry the compact list word
    uint4* d_lists = nullptr;              // two uint4 per octet
    int64_t n_octets = 0;
    std::vector<double> tree_octets;       // octets a row expects to read from every tree
    // The handle may be used from several host threads at once (one stream / workspace per
    // thread): everything a call mutates lives in its CallStats and is published under `mu` when
    // the call returns; calls that share the host pipeline or the event pool serialise on a mutex.
    struct Span { int kind; cudaEvent_t a, b; };
    struct CallStats {
        int launches = 0;
        std::vector<Span> spans;
    };
    std::mutex mu;                         // guards `last` and the event pool
    CallStats last;                        // counters / event spans of the most recent call
    std::mutex pipe_mu;                    // qf_forest_predict_quantiles_host: one call at a time per handle
    HostPipe pipe;
    // per-kernel CUDA-event timing of the most recent call (option "profile"; calls then serialise on prof_mu)
    int profile = 0;
    std::mutex prof_mu;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
};

namespace {

enum { SPAN_APPLY = 0, SPAN_QUANTILE = 1, SPAN_OTHER = 2 };

using CallStats = qf_forest::CallStats;

// One per entry-point call: its counters, and (profiling on) the lock that keeps the event pool
// to one call at a time.  The destructor publishes the counters as the handle's "last call".
struct CallScope {
    qf_forest* f;
    CallStats cs;
    std::unique_lock<std::mutex> prof;
    explicit CallScope(qf_forest* f_) : f(f_) {
        if (f->profile) {
            prof = std::unique_lock<std::mutex>(f->prof_mu);
            std::lock_guard<std::mutex> g(f->mu);
            f->events_used = 0;
        }
    }
    ~CallScope() {
        std::lock_guard<std::mutex> g(f->mu);
        f->last = std::move(cs);
    }
};

// RAII: records an event pair around a launch when profiling is on
struct SpanTimer {
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    SpanTimer(qf_forest* f, CallStats& cs, int kind, cudaStream_t st_) : st(st_) {
        if (!f->profile) return;
        cudaEvent_t ev[2];
        {
            std::lock_guard<std::mutex> g(f->mu);
            for (cudaEvent_t& e : ev) {
                if (f->events_used == f->event_pool.size()) {
                    cudaEvent_t n;
                    if (cudaEventCreate(&n) != cudaSuccess) return;
                    f->event_pool.push_back(n);
                }
                e = f->event_pool[f->events_used++];
            }
        }
        cudaEventRecord(ev[0], st);
        b = ev[1];
        cs.spans.push_back({kind, ev[0], ev[1]});
    }
    ~SpanTimer() {
        if (b) cudaEventRecord(b, st);
    }
};

// Which kernels a call runs, and how the caller's workspace is carved up for one chunk.
// The weight+percentile kernels form a chain; a row that does not fit a stage is appended
// to that stage's overflow list and picked up by the next one:
//   warp kernel (P <= 32*E, registers)  ->  bucket kernel (P <= 16 * threads, shared memory)
//   ->  bucket kernel with 512 threads (P <= 8192)  ->  dense kernel (any P, global scratch)
struct Plan {
    int64_t chunk_rows;
    int warp_elems;                // 0 = warp kernel not used
    int s1_threads, s1_limit;      // first bucket stage (0 threads = not used)
    int s2_limit;                  // second bucket stage, 512 threads (0 = not used)
    bool dense_needed;
    int dense_ctas;
    bool sort_rows;                // rows are walked in locality order (bucket/order/hist in use)
    int64_t off_seg, off_counts, off_rows[3], off_bucket, off_order, off_hist, off_dense, total;
};

constexpr int kMaxBucketThreads = 512;

// locality buckets: first member of the leaf reached in tree 0 (members of the first tree start at 0
// and there are at most n_train of them) >> key_shift, at most `counters` buckets
void set_key_bins(qf_forest* f, int counters) {
    const int64_t n0 = f->N;
    int shift = 0;
    while (((n0 - 1) >> shift) >= counters) ++shift;
    f->key_shift = shift;
    f->key_bins = static_cast<int>(((n0 - 1) >> shift) + 1);
    f->key_hist = counters;
}

Plan make_plan(const qf_forest* f, int64_t n_rows, int64_t rows_per_chunk = 0) {
    Plan pl;
    if (rows_per_chunk <= 0) rows_per_chunk = f->rows_per_chunk;
    pl.chunk_rows = std::max<int64_t>(1, std::min<int64_t>(n_rows, rows_per_chunk));
    const int64_t max_members = std::max<int64_t>(f->max_row_members, 1);
    const double typical = std::max(1.0, f->expected_row_members) * 1.25;
    // warp kernel: ranks must fit 32 - log2(32*E) bits, and the typical row must fit 32*E
    int we = f->warp_elems;
    if (we < 0) {
        we = typical <= 64 ? 2 : (typical <= 128 ? 4 : (typical <= 256 ? 8 : 0));
        if (max_members <= 64) we = 2;
        else if (max_members <= 128 && we != 2) we = 4;
    }
    if (we != 0 && f->N > (int64_t(1) << 24)) we = 0;
    pl.warp_elems = we;
    int64_t covered = 32 * we;                                    // largest row handled so far
    pl.s1_threads = pl.s1_limit = pl.s2_limit = 0;
    if (we == 0 || max_members > covered) {
        // the typical row must fit the first stage; rarer, larger rows go on to the 512-thread stage
        const int64_t want = std::min<int64_t>(max_members, static_cast<int64_t>(f->expected_row_members * 1.15) + 1);
        int threads = 64;
        while (threads < kMaxBucketThreads && qf::bucket_kernel_capacity(threads) < want) threads <<= 1;
        pl.s1_threads = threads;
        pl.s1_limit = qf::bucket_kernel_capacity(threads);
        if (f->cta_capacity > 0) pl.s1_limit = std::min(pl.s1_limit, f->cta_capacity);
        covered = pl.s1_limit;
        if (max_members > covered && f->cta2_capacity > 0 &&
            (threads < kMaxBucketThreads || pl.s1_limit < qf::bucket_kernel_capacity(threads))) {
            pl.s2_limit = std::min(qf::bucket_kernel_capacity(kMaxBucketThreads), f->cta2_capacity);
            covered = std::max<int64_t>(covered, pl.s2_limit);
        }
    }
    pl.dense_needed = max_members > covered;
    pl.dense_ctas = 0;
    if (pl.dense_needed) {
        const int64_t per = qf::dense_scratch_bytes(f->N);         // bound the scratch to ~2 GiB
        int64_t g = f->dense_ctas > 0 ? f->dense_ctas : f->sm_count;
        g = std::min<int64_t>(g, std::max<int64_t>(1, (int64_t(2) << 30) / per));
        pl.dense_ctas = static_cast<int>(std::min<int64_t>(g, pl.chunk_rows));
    }
    int64_t o = 0;
    // K1 output: (start, count) pairs row-major, or one 32-bit list word per (tree, row) tree-major
    // with rows padded to a multiple of four (select3), or row-major (select4)
    pl.off_seg = o;    o += align_up(std::max<int64_t>(pl.chunk_rows * f->T * 8, ((pl.chunk_rows + 3) & ~int64_t(3)) * f->T * 4));
    pl.off_counts = o; o += kAlign;                               // one int counter per overflow list
    for (int k = 0; k < 3; ++k) { pl.off_rows[k] = o; o += align_up(pl.chunk_rows * 4); }
    // locality order is worth its three small launches once there are enough rows for
    // neighbours in the order to be neighbours in feature space
    pl.sort_rows = f->sort_rows > 0 || (f->sort_rows < 0 && n_rows >= 2048);
    pl.off_bucket = o; o += align_up(pl.chunk_rows * 4);
    pl.off_order = o;  o += align_up(pl.chunk_rows * 4);
    pl.off_hist = o;   o += align_up(static_cast<int64_t>(f->key_hist) * 4);
    pl.off_dense = o;  o += pl.dense_needed ? qf::dense_scratch_bytes(f->N) * pl.dense_ctas : 0;
    pl.total = o;
    return pl;
}

int pick_row_warps(const qf_forest* f) {
    if (f->apply_row_warps > 0) return f->apply_row_warps;
    // 1 row-warp x 8 tree-warps measured best at configs 2 and 3 (32 rows per CTA: the most
    // CTAs, the smallest tile); fewer trees than tree-warps would leave warps idle
    int rw = 1;
    while (rw < 8 && (8 / rw) > f->T) rw <<= 1;
    return rw;
}

template <typename Kernel>
int set_smem(Kernel kernel, size_t smem) {
    // static + dynamic shared memory may not exceed 48 KB without the opt-in: leave room for the
    // kernels' static arrays (a few hundred bytes)
    if (smem > 40 * 1024)
        QF_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    return QF_OK;
}

// One launch of the leaf-lookup kernel over trees [t_begin, t_end).  `order` (device,
// nullable) is the locality order of the rows.
template <int EMIT>
int launch_apply(qf_forest* f, CallStats& cs, const float* X, int64_t n, int64_t ldx, void* out, const int* order,
                 int t_begin, int t_end, cudaStream_t st, int64_t ld_out = 0, int64_t ld_row = 1) {
    qf::ApplyParams ap;
    ap.nodes = EMIT == qf::APPLY_EMIT_COMPACT ? f->d_nodes_c : f->d_nodes;
    ap.ld_out = ld_out;
    ap.ld_row = ld_row;
    ap.roots = f->d_roots;
    ap.X = X;
    ap.order = order;
    ap.ldx = ldx;
    ap.n = n;
    ap.T = f->T;
    ap.t_begin = t_begin;
    ap.t_end = t_end;
    ap.F = f->F;
    ap.fbits = f->fbits;
    // a single tree cannot feed several tree-warps: all 8 warps along rows
    ap.row_warps = (t_end - t_begin == 1) ? 8 : pick_row_warps(f);
    while (ap.row_warps > 1 && static_cast<size_t>(f->F) * 32 * ap.row_warps * sizeof(float) > 200 * 1024)
        ap.row_warps >>= 1;
    const int rows_per_cta = 32 * ap.row_warps;
    const size_t smem = static_cast<size_t>(f->F) * rows_per_cta * sizeof(float);
    if (smem > 200 * 1024) {                            // rows wider than the shared-memory tile: one thread per row
        {
            SpanTimer span(f, cs, SPAN_APPLY, st);
            qf::apply_wide_rows_kernel<EMIT><<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(ap, out);
        }
        QF_CUDA_TRY(cudaGetLastError());
        ++cs.launches;
        return QF_OK;
    }
    const int64_t grid = (n + rows_per_cta - 1) / rows_per_cta;
    auto run = [&](auto kernel) -> int {
        int rc = set_smem(kernel, smem);
        if (rc != QF_OK) return rc;
        {
            SpanTimer span(f, cs, SPAN_APPLY, st);
            kernel<<<static_cast<unsigned>(grid), qf::kApplyThreads, smem, st>>>(ap, out);
        }
        QF_CUDA_TRY(cudaGetLastError());
        ++cs.launches;
        return QF_OK;
    };
    if (f->apply_ilp == 4 && f->apply_levels != qf::kApplyLevels) {
        switch (f->apply_levels) {
            case 1: return run(qf::apply_kernel<4, EMIT, 1>);
            case 3: return run(qf::apply_kernel<4, EMIT, 3>);
            case 4: return run(qf::apply_kernel<4, EMIT, 4>);
            default: break;
        }
    }
    switch (f->apply_ilp) {
        case 1: return run(qf::apply_kernel<1, EMIT>);
        case 2: return run(qf::apply_kernel<2, EMIT>);
        case 8: return run(qf::apply_kernel<8, EMIT>);
        default: return run(qf::apply_kernel<4, EMIT>);
    }
}

// Locality order of one chunk of rows (apply_kernel.cuh): bucket every row by its leaf in
// tree 0, exclusive scan of the bucket populations, scatter.  Returns the device order
// array, or nullptr when the plan does not sort.
int build_locality_order(qf_forest* f, CallStats& cs, const float* X, int64_t n, int64_t ldx, const Plan& pl,
                         unsigned char* base, cudaStream_t st, const int** order_out) {
    *order_out = nullptr;
    if (!pl.sort_rows) return QF_OK;
    int* bucket = reinterpret_cast<int*>(base + pl.off_bucket);
    int* order = reinterpret_cast<int*>(base + pl.off_order);
    int* hist = reinterpret_cast<int*>(base + pl.off_hist);
    QF_CUDA_TRY(cudaMemsetAsync(hist, 0, static_cast<size_t>(f->key_hist) * sizeof(int), st));
    {
        SpanTimer span(f, cs, SPAN_OTHER, st);
        qf::locality_key_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(
            f->d_nodes, f->root0, X, ldx, n, f->fbits, f->key_shift, f->key_bins, bucket, hist);
    }
    QF_CUDA_TRY(cudaGetLastError());
    {
        SpanTimer span(f, cs, SPAN_OTHER, st);
        qf::exclusive_scan_kernel<<<1, qf::kScanThreads, 0, st>>>(hist, f->key_hist);
    }
    QF_CUDA_TRY(cudaGetLastError());
    {
        SpanTimer span(f, cs, SPAN_OTHER, st);
        qf::order_scatter_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(bucket, n, hist, order);
    }
    QF_CUDA_TRY(cudaGetLastError());
    cs.launches += 3;
    *order_out = order;
    return QF_OK;
}

int check_forest_call(const qf_forest* f, const void* X, int64_t n, int64_t ldx, const void* out,
                      const char* who) {
    if (!f) return qf::fail(QF_ERR_INVALID, std::string(who) + ": null forest");
    if (n < 0 || (n > 0 && (!X || !out)))
        return qf::fail(QF_ERR_INVALID, std::string(who) + ": null buffer or negative row count");
    if (ldx < f->F) return qf::fail(QF_ERR_INVALID, std::string(who) + ": ldx < n_features");
    return QF_OK;
}

int check_quantiles(const double* q_host, int nq, int mode, const char* who) {
    if (nq <= 0 || !q_host) return qf::fail(QF_ERR_INVALID, std::string(who) + ": no quantiles");
    for (int i = 0; i < nq; ++i)
        if (!(q_host[i] >= 0.0 && q_host[i] <= 100.0))          // utils.py:42-44
            return qf::fail(QF_ERR_INVALID, "q should be in-between 0 and 100");
    if (mode != QF_MODE_EXACT && mode != QF_MODE_FAST)
        return qf::fail(QF_ERR_INVALID, std::string(who) + ": unknown mode");
    return QF_OK;
}

int launch_bucket_kernel(int threads, unsigned grid, const qf::QuantileParams& qp, const qf::QuantileList& ql,
                         int T, cudaStream_t st) {
    const size_t smem = qf::bucket_kernel_smem(threads, T);
    if (smem > 220 * 1024)
        return qf::fail(QF_ERR_LIMIT, "predict_quantiles: n_estimators too large for shared memory");
    auto run = [&](auto kernel) -> int {
        int rc = set_smem(kernel, smem);
        if (rc != QF_OK) return rc;
        kernel<<<grid, threads, smem, st>>>(qp, ql);
        return QF_OK;
    };
    switch (threads) {
        case 64: return run(qf::quantile_bucket_kernel<64>);
        case 128: return run(qf::quantile_bucket_kernel<128>);
        case 256: return run(qf::quantile_bucket_kernel<256>);
        default: return run(qf::quantile_bucket_kernel<512>);
    }
}

// QF_MODE_FAST: the streaming selection kernel (select_kernel.cuh) takes every row the warp
// kernel does not, whatever its number of members
int launch_select_kernel(const qf_forest* f, unsigned grid, const qf::QuantileParams& qp,
                         const qf::QuantileList& ql, cudaStream_t st) {
    const size_t smem = qf::select_kernel_smem();
    int k = 0;                                          // weights of a row sum to <= bound < 2^k
    const int64_t bound = std::max<int64_t>(f->T, f->row_weight_bound);   // T when every leaf's weights sum to one
    while ((int64_t(1) << k) < bound + 1) ++k;
    if (k > 20) return qf::fail(QF_ERR_LIMIT, "predict_quantiles: total weight per row too large for fixed-point weights");
    const float fx_scale = static_cast<float>(int64_t(1) << (31 - k));
    int nbits = 0;                                      // ranks are < 2^nbits
    while ((int64_t(1) << nbits) < f->N) ++nbits;
    // training sets of up to 2M samples: one refinement level is enough (select2_kernel.cuh)
    if (f->select_impl != 1 && f->select_bits == qf::kSelectL2Bits && nbits <= 12 + qf::kSelect2SlotBits &&
        f->T <= qf::kSelect2MaxIPT * qf::kSelect2Threads && f->n_members < (int64_t(1) << 31) &&
        static_cast<int>(f->tree_max_leaf.size()) == f->T) {
        // pair-table segment of every warp: its leaves are 32 * ipt consecutive trees, a leaf of c
        // members spans at most c / 2 + 1 aligned pairs
        qf::Select2Layout lay;
        int ipt = (f->T + qf::kSelect2Threads - 1) / qf::kSelect2Threads;   // leaves per lane: 1, 2 or 4
        if (ipt == 3) ipt = 4;
        uint64_t pairs = 0;
        for (int w = 0; w < qf::kSelect2Warps; ++w) {
            lay.warp_off[w] = static_cast<uint32_t>(pairs);
            for (int t = w * 32 * ipt; t < std::min(f->T, (w + 1) * 32 * ipt); ++t)
                pairs += f->tree_max_leaf[t] / 2 + 1;
        }
        lay.warp_off[qf::kSelect2Warps] = static_cast<uint32_t>(pairs);
        if (pairs <= static_cast<uint64_t>(qf::kSelect2MaxPairs)) {
            const int log2nb = nbits <= 11 + qf::kSelect2SlotBits ? 11 : 12;
            const int shift1 = std::max(0, nbits - log2nb);
            const size_t smem2 = qf::select2_kernel_smem(log2nb, shift1, static_cast<uint32_t>(pairs));
            const unsigned ctas = static_cast<unsigned>(f->sm_count) * qf::kSelect2CtasPerSm;
            auto run2 = [&](auto kernel) -> int {
                int rc = set_smem(kernel, smem2);
                if (rc != QF_OK) return rc;
                kernel<<<std::min(grid, ctas), qf::kSelect2Threads, smem2, st>>>(qp, ql, lay, fx_scale, shift1);
                const cudaError_t e = cudaPeekAtLastError();
                if (e != cudaSuccess) {
                    char msg[256];
                    std::snprintf(msg, sizeof(msg), "quantile_select2_kernel launch failed: %s (grid %u, %zu bytes of "
                                  "shared memory, %llu pairs)", cudaGetErrorString(e), std::min(grid, ctas), smem2,
                                  static_cast<unsigned long long>(pairs));
                    cudaGetLastError();
                    return qf::fail(QF_ERR_CUDA, msg);
                }
                return QF_OK;
            };
            if (log2nb == 11)
                return ipt == 1 ? run2(qf::quantile_select2_kernel<11, 1>)
                     : ipt == 2 ? run2(qf::quantile_select2_kernel<11, 2>) : run2(qf::quantile_select2_kernel<11, 4>);
            return ipt == 1 ? run2(qf::quantile_select2_kernel<12, 1>)
                 : ipt == 2 ? run2(qf::quantile_select2_kernel<12, 2>) : run2(qf::quantile_select2_kernel<12, 4>);
        }
    }
    const int threads = f->T <= 128 ? 128 : 256;        // 4 lanes per leaf list
    auto run = [&](auto kernel) -> int {
        int rc = set_smem(kernel, smem);
        if (rc != QF_OK) return rc;
        kernel<<<grid, threads, smem, st>>>(qp, ql, fx_scale, nbits, f->select_bits);
        return QF_OK;
    };
    return threads == 128 ? run(qf::quantile_select_kernel<128>) : run(qf::quantile_select_kernel<256>);
}

// Compact leaf lists (compact_lists.cuh) for the select3 kernel, built once per forest on the
// first call that can use them.  Synchronous (cudaMalloc + a handful of launches on the legacy
// stream); returns with f->compact_state = 1 (ready) or -1 (forest not compactable: the 8-byte
// kernels stay in charge).
int ensure_compact(qf_forest* f) {
    std::lock_guard<std::mutex> guard(f->compact_mu);
    if (f->compact_state != 0) return QF_OK;
    f->compact_state = -1;
    if (f->N > (int64_t(1) << 24) || f->n_members <= 0) return QF_OK;     // ranks live in 24 bits
    // fixed-point units per unit of weight (as in launch_select_kernel): the octet headers carry
    // rn(scale / leaf count sum), the weight of one bootstrap count
    int kbits = 0;
    const int64_t bound = std::max<int64_t>(f->T, f->row_weight_bound);
    while ((int64_t(1) << kbits) < bound + 1) ++kbits;
    if (kbits > 20) return QF_OK;
    const float fx_scale = static_cast<float>(int64_t(1) << (31 - kbits));
    struct Tmp {
        void* p[4] = {};
        ~Tmp() { for (void* q : p) cudaFree(q); }
    } tmp;
    const int64_t n_tiles = (f->n_members + qf::kCompactScanTile - 1) / qf::kCompactScanTile;
    QF_CUDA_TRY(cudaMalloc(&tmp.p[0], static_cast<size_t>(f->n_members) * 4));          // octets per leaf -> first octet
    QF_CUDA_TRY(cudaMalloc(&tmp.p[1], static_cast<size_t>(n_tiles + 1) * 8));           // tile sums + total
    QF_CUDA_TRY(cudaMalloc(&tmp.p[2], 16));                                             // flags
    QF_CUDA_TRY(cudaMalloc(&tmp.p[3], static_cast<size_t>(f->T) * 16));                 // per-tree stats
    uint32_t* oct_at = static_cast<uint32_t*>(tmp.p[0]);
    uint64_t* tile_sum = static_cast<uint64_t*>(tmp.p[1]);
    int* flags = static_cast<int*>(tmp.p[2]);
    unsigned long long* stats = static_cast<unsigned long long*>(tmp.p[3]);
    QF_CUDA_TRY(cudaMemset(oct_at, 0, static_cast<size_t>(f->n_members) * 4));
    QF_CUDA_TRY(cudaMemset(flags, 0, 16));
    QF_CUDA_TRY(cudaMemset(stats, 0, static_cast<size_t>(f->T) * 16));
    const unsigned node_blocks = static_cast<unsigned>((f->n_nodes + 255) / 256);
    qf::compact_mark_kernel<<<node_blocks, 256>>>(f->d_nodes, f->n_nodes, f->d_members, oct_at, flags, f->compact_align);
    qf::scan_tile_sums_kernel<<<static_cast<unsigned>(n_tiles), 256>>>(oct_at, f->n_members, tile_sum);
    qf::scan_tile_offsets_kernel<<<1, 1024>>>(tile_sum, n_tiles, tile_sum + n_tiles);
    QF_CUDA_TRY(cudaGetLastError());
    int h_flags = 0;
    uint64_t n_octets = 0;
    QF_CUDA_TRY(cudaMemcpy(&h_flags, flags, 4, cudaMemcpyDeviceToHost));
    QF_CUDA_TRY(cudaMemcpy(&n_octets, tile_sum + n_tiles, 8, cudaMemcpyDeviceToHost));
    if (h_flags != 0 || n_octets == 0 || n_octets > qf::kCompactOffMask) {               // not compactable
        char why[160];
        std::snprintf(why, sizeof(why), "compact leaf lists not built: flags %d (2 weights are not count / sum, "
                      "4 rank >= 2^24), %llu octets", h_flags, static_cast<unsigned long long>(n_octets));
        qf::last_error_slot() = why;                    // informational: the 8-byte kernels take over
        return QF_OK;
    }
    qf::scan_apply_kernel<<<static_cast<unsigned>(n_tiles), 256>>>(oct_at, f->n_members, tile_sum);
    uint2* nodes_c = nullptr;
    uint4* lists = nullptr;
    cudaError_t e = cudaMalloc(&nodes_c, static_cast<size_t>(f->n_nodes) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&lists, static_cast<size_t>(n_octets) * 32);
    if (e == cudaSuccess) {
        qf::compact_fill_kernel<<<node_blocks, 256>>>(f->d_nodes, f->n_nodes, f->d_members, oct_at, fx_scale, static_cast<uint32_t>(f->N), nodes_c, lists, f->compact_align);
        qf::compact_tree_stats_kernel<<<f->T, 256>>>(f->d_nodes, f->d_roots, f->n_nodes, f->T, stats);
        e = cudaGetLastError();
    }
    std::vector<unsigned long long> h_stats(static_cast<size_t>(f->T) * 2);
    if (e == cudaSuccess) e = cudaMemcpy(h_stats.data(), stats, h_stats.size() * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        cudaFree(nodes_c);
        cudaFree(lists);
        return qf::fail(QF_ERR_CUDA, std::string("compact leaf lists: ") + cudaGetErrorString(e));
    }
    f->tree_octets.assign(f->T, 1.0);
    for (int t = 0; t < f->T; ++t)
        if (h_stats[2 * t + 1]) f->tree_octets[t] = static_cast<double>(h_stats[2 * t]) / static_cast<double>(h_stats[2 * t + 1]);
    f->d_nodes_c = nodes_c;
    f->d_lists = lists;
    f->n_octets = static_cast<int64_t>(n_octets);
    f->device_bytes += f->n_nodes * 8 + static_cast<int64_t>(n_octets) * 32;
    f->compact_state = 1;
    return QF_OK;
}

// Geometry of the select3 kernel for this forest, or ok = false when it cannot take it (then
// select2 / the multi-level kernel do).
struct Select3Geometry {
    bool ok = false;
    int log2nb = 11, shift1 = 0, ipt = 1, capw = 0, stages = 2, selw = 3;
    float fx_scale = 0.0f;
    size_t smem = 0;
};

Select3Geometry select3_geometry(const qf_forest* f, int nq) {
    Select3Geometry g;
    if (f->select_impl != 3 || f->select_bits != qf::kSelectL2Bits || nq > qf::kSelect2QB) return g;
    if (f->T > qf::kSelect3MaxIPT * qf::kSelect3Workers) return g;
    int k = 0;                                          // weights of a row sum to <= bound < 2^k
    const int64_t bound = std::max<int64_t>(f->T, f->row_weight_bound);
    while ((int64_t(1) << k) < bound + 1) ++k;
    if (k > 20) return g;
    int nbits = 0;                                      // ranks are < 2^nbits
    while ((int64_t(1) << nbits) < f->N) ++nbits;
    if (nbits > 12 + qf::kSelect2SlotBits) return g;
    g.fx_scale = static_cast<float>(int64_t(1) << (31 - k));
    g.log2nb = nbits <= 11 + qf::kSelect2SlotBits ? 11 : 12;
    g.shift1 = std::max(0, nbits - g.log2nb);
    g.ipt = (f->T + qf::kSelect3Workers - 1) / qf::kSelect3Workers;
    if (g.ipt == 3) g.ipt = 4;
    return g;
}

// stage capacity per warp: 1.25 x the octets the busiest warp expects + 8, a multiple of 4
bool select3_finish_geometry(const qf_forest* f, Select3Geometry& g) {
    if (f->compact_state != 1 || static_cast<int>(f->tree_octets.size()) != f->T) return false;
    double busiest = 0.0;
    for (int w = 0; w < qf::kSelect3Warps; ++w) {
        double e = 0.0;
        for (int i = 0; i < g.ipt; ++i)
            for (int l = 0; l < 32; ++l) {
                const int t = w * 32 + l + i * qf::kSelect3Workers;
                if (t < f->T) e += f->tree_octets[t];
            }
        busiest = std::max(busiest, e);
    }
    int capw = static_cast<int>(busiest * 1.25) + 8;
    if (f->select3_capw > 0) capw = f->select3_capw;
    capw = (capw + 3) & ~3;
    g.capw = capw;
    g.stages = f->select3_variant == 0 ? 2 : 1;
    g.selw = f->select3_variant == 0 ? 3 : 1;
    g.smem = qf::select3_kernel_smem(g.log2nb, g.shift1, capw, g.stages);
    // 2 (3) CTAs per SM is what the kernel is tuned for; one oversized stage (big leaves) is not worth it
    g.ok = g.smem <= static_cast<size_t>(g.stages == 2 ? 112 : 75) * 1024;
    return g.ok;
}

// select4 (warp per row): geometry or ok = false
struct Select4Geometry {
    bool ok = false;
    int list_cap = 0;
    float fx_scale = 0.0f;
    size_t smem = 0;
};

Select4Geometry select4_geometry(const qf_forest* f, int nq) {
    Select4Geometry g;
    if (f->select_impl != 4 || nq > qf::kSelect2QB || f->T > qf::kSelect4MaxTrees) return g;
    int k = 0;
    const int64_t bound = std::max<int64_t>(f->T, f->row_weight_bound);
    while ((int64_t(1) << k) < bound + 1) ++k;
    if (k > 20) return g;
    int nbits = 0;
    while ((int64_t(1) << nbits) < f->N) ++nbits;
    if (nbits > 24) return g;
    g.fx_scale = static_cast<float>(int64_t(1) << (31 - k));
    return g;
}

bool select4_finish_geometry(const qf_forest* f, Select4Geometry& g) {
    if (f->compact_state != 1 || static_cast<int>(f->tree_octets.size()) != f->T) return false;
    double octets = 0.0;                                // octets a row expects
    for (int t = 0; t < f->T; ++t) octets += f->tree_octets[t];
    int cap = static_cast<int>(octets * 1.25) + 96;     // rows beyond it walk their leaves one by one (slow, correct)
    if (f->select4_list > 0) cap = f->select4_list;
    cap = std::min((cap + 31) & ~31, 4096);
    g.list_cap = cap;
    g.smem = qf::select4_warp_smem(cap) * qf::kSelect4Warps;
    g.ok = g.smem <= 200 * 1024;
    return g.ok;
}

int launch_select4_kernel(const qf_forest* f, const Select4Geometry& g, const qf::Select4Params& sp,
                          const qf::QuantileList& ql, cudaStream_t st) {
    const int64_t want = (sp.n + qf::kSelect4Warps - 1) / qf::kSelect4Warps;
    const int wanted = std::min(std::max(f->select4_ctas, 1), qf::kSelect4CtasPerSm);
    const int per_sm = std::max<int>(1, std::min<int>(wanted, static_cast<int>((220 * 1024) / std::max<size_t>(g.smem + 1024, 1))));
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(want, int64_t(f->sm_count) * per_sm));
    auto run = [&](auto kernel) -> int {
        int rc = set_smem(kernel, g.smem);
        if (rc != QF_OK) return rc;
        kernel<<<grid, qf::kSelect4Threads, g.smem, st>>>(sp, ql);
        return QF_OK;
    };
    if (f->select4_unroll == 4) return run(qf::quantile_select4_kernel<4, false>);
    return f->select4_hints ? run(qf::quantile_select4_kernel<2, true>) : run(qf::quantile_select4_kernel<2, false>);
}

int launch_select3_kernel(const qf_forest* f, const Select3Geometry& g, const qf::Select3Params& sp,
                          const qf::QuantileList& ql, cudaStream_t st) {
    const int64_t n_groups = (sp.n + qf::kSelect3RowGroup - 1) / qf::kSelect3RowGroup;
    const unsigned grid = static_cast<unsigned>(
        std::min<int64_t>(n_groups, int64_t(f->sm_count) * qf::select3_ctas_per_sm(g.stages)));
    auto run = [&](auto kernel) -> int {
        int rc = set_smem(kernel, g.smem);
        if (rc != QF_OK) return rc;
        kernel<<<grid, qf::select3_threads(g.selw), g.smem, st>>>(sp, ql);
        return QF_OK;
    };
    auto pick_ipt = [&](auto log2nb_c, auto selw_c, auto stages_c) -> int {
        constexpr int L = decltype(log2nb_c)::value, W = decltype(selw_c)::value, S = decltype(stages_c)::value;
        return g.ipt == 1 ? run(qf::quantile_select3_kernel<L, 1, W, S>)
             : g.ipt == 2 ? run(qf::quantile_select3_kernel<L, 2, W, S>) : run(qf::quantile_select3_kernel<L, 4, W, S>);
    };
    using I11 = std::integral_constant<int, 11>;
    using I12 = std::integral_constant<int, 12>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;
    if (g.stages == 2)
        return g.log2nb == 11 ? pick_ipt(I11{}, I3{}, I2{}) : pick_ipt(I12{}, I3{}, I2{});
    return g.log2nb == 11 ? pick_ipt(I11{}, I1{}, I1{}) : pick_ipt(I12{}, I1{}, I1{});
}

// QF_MODE_FAST adds weights as 32-bit fixed point with 2^(31-k) units per unit of weight
// (2^k > total weight of a row).  A member of a leaf of L samples weighs about 1/L, i.e.
// 2^(31-k) / L units: below ~2^14 units the rounding of each weight (the same direction for a
// whole leaf) shifts tail percentiles visibly (tools/fuzz_parity.py: 3e-3 at q = 99.9 with
// leaves of thousands of members).  Such forests take the exact kernels in both modes.
bool fast_mode_is_precise(const qf_forest* f) {
    if (f->row_weight_bound > f->T) return true;        // unit weights (tree-level estimators): nothing to round
    int k = 0;
    while ((int64_t(1) << k) < int64_t(f->T) + 1) ++k;
    uint32_t largest = 1;
    for (uint32_t m : f->tree_max_leaf) largest = std::max(largest, m);
    return (static_cast<int64_t>(largest) << k) <= (int64_t(1) << 17);
}

// `pl` is the caller's carve-up of `ws` (made once for the largest chunk, so that the
// zeroed dense scratch stays where the kernels expect it for every chunk).
int quantiles_on_stream(qf_forest* f, CallStats& cs, const float* X, int64_t n, int64_t ldx, const double* q_host,
                        int nq, double* out, int mode, const Plan& pl, void* ws, int64_t ws_bytes,
                        cudaStream_t st) {
    // QF_MODE_FAST swaps the shared-memory sorting stages for histogram selection; the warp
    // kernel (small member sets) and the dense fallback are exact in both modes
    const bool select = mode == QF_MODE_FAST && nq <= qf::kSelectMaxQ && fast_mode_is_precise(f);
    if (ws_bytes < pl.total || (!ws && pl.total > 0))
        return qf::fail(QF_ERR_WORKSPACE, "predict_quantiles: workspace too small");
    unsigned char* base = static_cast<unsigned char*>(ws);
    uint2* seg = reinterpret_cast<uint2*>(base + pl.off_seg);
    int* counts = reinterpret_cast<int*>(base + pl.off_counts);   // [k]: rows in overflow list k
    int* lists[3];
    for (int k = 0; k < 3; ++k) lists[k] = reinterpret_cast<int*>(base + pl.off_rows[k]);

    // QF_MODE_FAST on forests whose typical row is too large for the warp kernel: every row goes
    // through the select3 kernel on compact leaf lists (one 32-bit list word per (tree, row) out of
    // K1, tree-major)
    Select3Geometry g3;
    if (select && pl.warp_elems == 0) {
        g3 = select3_geometry(f, nq);
        if (g3.fx_scale > 0.0f) {
            int rc = ensure_compact(f);
            if (rc != QF_OK) return rc;
            select3_finish_geometry(f, g3);
        }
    }
    const int64_t ld3 = (pl.chunk_rows + 3) & ~int64_t(3);
    Select4Geometry g4;
    if (select && pl.warp_elems == 0) {
        g4 = select4_geometry(f, nq);
        if (g4.fx_scale > 0.0f) {
            int rc = ensure_compact(f);
            if (rc != QF_OK) return rc;
            select4_finish_geometry(f, g4);
        }
    }
    const int64_t ldT = (f->T + 3) & ~3;

    for (int64_t r0 = 0; r0 < n; r0 += pl.chunk_rows) {
        const int64_t rows = std::min(pl.chunk_rows, n - r0);
        const int* order = nullptr;
        int rc = build_locality_order(f, cs, X + r0 * ldx, rows, ldx, pl, base, st, &order);
        if (rc != QF_OK) return rc;
        if (g4.ok) {                                    // warp per row: K1 writes the list words row-major
            rc = launch_apply<qf::APPLY_EMIT_COMPACT>(f, cs, X + r0 * ldx, rows, ldx, seg, order, 0, f->T, st, 1, ldT);
            if (rc != QF_OK) return rc;
            qf::QuantileList ql;
            std::memset(&ql, 0, sizeof(ql));
            for (int i = 0; i < nq; ++i) ql.q[i] = q_host[i];
            qf::Select4Params sp;
            sp.seg = reinterpret_cast<const uint32_t*>(seg);
            sp.ldT = ldT;
            sp.order = order;
            sp.lists = f->d_lists;
            sp.y_sorted = f->d_y_sorted;
            sp.out = out + r0 * nq;
            sp.n = rows;
            sp.ldo = nq;
            sp.T = f->T;
            sp.Q = nq;
            sp.fx_scale = g4.fx_scale;
            sp.n_train = static_cast<uint32_t>(f->N);
            sp.list_cap = g4.list_cap;
            sp.hits_cap = std::min(f->select4_hits, qf::kSelect4Hits);
            {
                SpanTimer span(f, cs, SPAN_QUANTILE, st);
                rc = launch_select4_kernel(f, g4, sp, ql, st);
            }
            if (rc != QF_OK) return rc;
            QF_CUDA_TRY(cudaGetLastError());
            ++cs.launches;
            continue;
        }
        if (g3.ok) {
            rc = launch_apply<qf::APPLY_EMIT_COMPACT>(f, cs, X + r0 * ldx, rows, ldx, seg, order, 0, f->T, st, ld3);
            if (rc != QF_OK) return rc;
            qf::QuantileList ql;
            std::memset(&ql, 0, sizeof(ql));
            for (int i = 0; i < nq; ++i) ql.q[i] = q_host[i];
            qf::Select3Params sp;
            sp.seg = reinterpret_cast<const uint32_t*>(seg);
            sp.ld = ld3;
            sp.order = order;
            sp.lists = f->d_lists;
            sp.y_sorted = f->d_y_sorted;
            sp.out = out + r0 * nq;
            sp.n = rows;
            sp.ldo = nq;
            sp.T = f->T;
            sp.Q = nq;
            sp.fx_scale = g3.fx_scale;
            sp.shift1 = g3.shift1;
            sp.capw = g3.capw;
            {
                SpanTimer span(f, cs, SPAN_QUANTILE, st);
                rc = launch_select3_kernel(f, g3, sp, ql, st);
            }
            if (rc != QF_OK) return rc;
            QF_CUDA_TRY(cudaGetLastError());
            ++cs.launches;
            continue;
        }
        rc = launch_apply<qf::APPLY_EMIT_SEGMENT>(f, cs, X + r0 * ldx, rows, ldx, seg, order, 0, f->T, st);
        if (rc != QF_OK) return rc;
        for (int q0 = 0; q0 < nq; q0 += qf::kMaxQuantilesPerLaunch) {
            const int qn = std::min(qf::kMaxQuantilesPerLaunch, nq - q0);
            qf::QuantileList ql;
            std::memset(&ql, 0, sizeof(ql));
            for (int i = 0; i < qn; ++i) ql.q[i] = q_host[q0 + i];
            qf::QuantileParams qp;
            qp.seg = seg;
            qp.order = order;
            qp.members = f->d_members;
            qp.y_sorted = f->d_y_sorted;
            qp.out = out + r0 * nq + q0;
            qp.n = rows;
            qp.ldo = nq;
            qp.T = f->T;
            qp.Q = qn;
            qp.cap = 0;
            qp.row_list = nullptr;                       // first stage: every row of the chunk
            qp.row_count = nullptr;
            QF_CUDA_TRY(cudaMemsetAsync(counts, 0, 4 * sizeof(int), st));
            int stage = 0;                               // stage k appends to lists[k]
            auto chain = [&]() {                         // wire this stage's overflow, next stage's input
                qp.overflow_count = counts + stage;
                qp.overflow_rows = lists[stage];
            };
            auto advance = [&]() {
                qp.row_list = lists[stage];
                qp.row_count = counts + stage;
                ++stage;
            };
            if (pl.warp_elems) {                         // one warp per row
                chain();
                const unsigned grid =
                    static_cast<unsigned>((rows + qf::kWarpKernelWarps - 1) / qf::kWarpKernelWarps);
                {
                    SpanTimer span(f, cs, SPAN_QUANTILE, st);
                    // unit_leaves (deep bootstrap trees, min_samples_leaf=1): member t of the row is
                    // the single member of its leaf in tree t -- no offsets to scan
                    const bool unit = f->unit_leaves && f->warp_unit != 0;
                    if (pl.warp_elems == 2) {
                        if (unit) qf::quantile_warp_kernel<2, true><<<grid, qf::kWarpKernelWarps * 32, 0, st>>>(qp, ql);
                        else qf::quantile_warp_kernel<2, false><<<grid, qf::kWarpKernelWarps * 32, 0, st>>>(qp, ql);
                    } else if (pl.warp_elems == 4) {
                        if (unit) qf::quantile_warp_kernel<4, true><<<grid, qf::kWarpKernelWarps * 32, 0, st>>>(qp, ql);
                        else qf::quantile_warp_kernel<4, false><<<grid, qf::kWarpKernelWarps * 32, 0, st>>>(qp, ql);
                    } else {
                        if (unit) qf::quantile_warp_kernel<8, true><<<grid, qf::kWarpKernelWarps * 32, 0, st>>>(qp, ql);
                        else qf::quantile_warp_kernel<8, false><<<grid, qf::kWarpKernelWarps * 32, 0, st>>>(qp, ql);
                    }
                }
                QF_CUDA_TRY(cudaGetLastError());
                ++cs.launches;
                advance();
            }
            if (select && (pl.s1_threads || pl.dense_needed)) {   // every remaining row, any member count
                chain();                                 // (never overflows)
                const unsigned grid = qp.row_list ? static_cast<unsigned>(std::min<int64_t>(rows, int64_t(f->sm_count) * 8))
                                                  : static_cast<unsigned>(rows);
                {
                    SpanTimer span(f, cs, SPAN_QUANTILE, st);
                    rc = launch_select_kernel(f, grid, qp, ql, st);
                }
                if (rc != QF_OK) return rc;
                QF_CUDA_TRY(cudaGetLastError());
                ++cs.launches;
                continue;
            }
            const int bucket_threads[2] = {pl.s1_threads, pl.s2_limit ? kMaxBucketThreads : 0};
            const int bucket_limit[2] = {pl.s1_limit, pl.s2_limit};
            for (int b = 0; b < 2; ++b) {
                if (!bucket_threads[b]) continue;
                chain();
                qp.cap = bucket_limit[b];
                // the first stage gets a CTA per row; overflow stages a persistent grid over their list
                const unsigned grid = qp.row_list ? static_cast<unsigned>(std::min<int64_t>(rows, int64_t(f->sm_count) * 4))
                                                  : static_cast<unsigned>(rows);
                {
                    SpanTimer span(f, cs, SPAN_QUANTILE, st);
                    rc = launch_bucket_kernel(bucket_threads[b], grid, qp, ql, f->T, st);
                }
                if (rc != QF_OK) return rc;
                QF_CUDA_TRY(cudaGetLastError());
                ++cs.launches;
                advance();
            }
            if (pl.dense_needed) {
                chain();                                 // the dense kernel never overflows
                {
                    SpanTimer span(f, cs, SPAN_QUANTILE, st);
                    qf::quantile_dense_kernel<256><<<pl.dense_ctas, 256, 0, st>>>(qp, ql, base + pl.off_dense, f->N);
                }
                QF_CUDA_TRY(cudaGetLastError());
                ++cs.launches;
            }
        }
    }
    return QF_OK;
}

// the dense scratch must start zeroed (the kernel restores it after every row)
int zero_dense_scratch(const qf_forest* f, const Plan& pl, void* ws, cudaStream_t st) {
    if (!pl.dense_needed) return QF_OK;
    QF_CUDA_TRY(cudaMemsetAsync(static_cast<unsigned char*>(ws) + pl.off_dense, 0,
                                qf::dense_scratch_bytes(f->N) * pl.dense_ctas, st));
    return QF_OK;
}

}  // namespace

extern "C" {

int qf_abi_version(void) { return QF_ABI_VERSION; }

const char* qf_last_error(void) { return qf::last_error_slot().c_str(); }

int qf_forest_create(const qf_forest_desc* d, qf_forest** out) {
    if (!d || !out) return qf::fail(QF_ERR_INVALID, "qf_forest_create: null argument");
    *out = nullptr;
    if (d->n_trees <= 0 || d->n_features <= 0 || d->n_train <= 0 || d->n_nodes <= 0 ||
        d->n_members <= 0 || !d->nodes || !d->tree_offsets || !d->members || !d->y_sorted)
        return qf::fail(QF_ERR_INVALID, "qf_forest_create: empty forest or null array");
    if (d->feature_bits < 1 || d->feature_bits > 24 || (int64_t(1) << d->feature_bits) < d->n_features)
        return qf::fail(QF_ERR_INVALID, "qf_forest_create: feature_bits does not cover n_features");
    if (d->n_nodes >= (int64_t(1) << 31) || d->n_members > 0xFFFFFFFFll || d->n_train >= (int64_t(1) << 31))
        return qf::fail(QF_ERR_LIMIT, "qf_forest_create: forest exceeds the 32-bit packed layout");
    int count = 0;
    QF_CUDA_TRY(cudaGetDeviceCount(&count));
    if (d->device < 0 || d->device >= count)
        return qf::fail(QF_ERR_CUDA, "qf_forest_create: no such CUDA device");
    DeviceGuard guard(d->device);
    if (!guard.ok) return qf::fail(QF_ERR_CUDA, "qf_forest_create: cudaSetDevice failed");
    cudaDeviceProp prop;
    QF_CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major != 10)
        return qf::fail(QF_ERR_CUDA, "qf_forest_create: device is not sm_100 (this library has sm_100a code only)");

    qf_forest* f = new qf_forest();
    f->device = d->device;
    f->T = d->n_trees;
    f->F = d->n_features;
    f->fbits = d->feature_bits;
    f->N = d->n_train;
    f->n_nodes = d->n_nodes;
    f->n_members = d->n_members;
    f->max_row_members = d->max_row_members > 0 ? d->max_row_members : d->n_members;
    f->expected_row_members = d->expected_row_members > 0 ? d->expected_row_members
                                                          : static_cast<double>(f->max_row_members);
    {   // locality buckets: first member of the leaf reached in tree 0 (members of the first tree
        // start at 0 and there are at most n_train of them), at most kKeyBins buckets
        f->root0 = static_cast<uint32_t>(d->tree_offsets[0]);
        set_key_bins(f, qf::kKeyBins);
    }
    f->sm_count = prop.multiProcessorCount;

    std::vector<uint32_t> roots(d->n_trees);
    for (int t = 0; t < d->n_trees; ++t) roots[t] = static_cast<uint32_t>(d->tree_offsets[t]);

    auto upload = [&](void** dst, const void* src, int64_t bytes) -> int {
        // + 16: kernels read members in aligned 128-bit pairs, the last pair may straddle the end
        QF_CUDA_TRY(cudaMalloc(dst, static_cast<size_t>(bytes) + 16));
        QF_CUDA_TRY(cudaMemset(static_cast<unsigned char*>(*dst) + bytes, 0, 16));
        // cudaMemcpyDefault: the source may be host memory or (UVA) device memory, e.g. a
        // buffer that was just broadcast to this rank over NCCL
        QF_CUDA_TRY(cudaMemcpy(*dst, src, static_cast<size_t>(bytes), cudaMemcpyDefault));
        f->device_bytes += bytes;
        return QF_OK;
    };
    int rc = upload(reinterpret_cast<void**>(&f->d_nodes), d->nodes, d->n_nodes * 8);
    if (rc == QF_OK) rc = upload(reinterpret_cast<void**>(&f->d_roots), roots.data(), d->n_trees * 4);
    if (rc == QF_OK && d->node_ids)
        rc = upload(reinterpret_cast<void**>(&f->d_node_ids), d->node_ids, d->n_nodes * 4);
    if (rc == QF_OK) rc = upload(reinterpret_cast<void**>(&f->d_members), d->members, d->n_members * 8);
    if (rc == QF_OK) rc = upload(reinterpret_cast<void**>(&f->d_y_sorted), d->y_sorted, d->n_train * 4);
    if (rc == QF_OK && d->leaf_values)
        rc = upload(reinterpret_cast<void**>(&f->d_values), d->leaf_values, d->n_nodes * 8);
    if (rc == QF_OK) {                                  // largest leaf of every tree (sizes the pair table of select2)
        uint32_t* d_max = nullptr;
        cudaError_t e = cudaMalloc(&d_max, static_cast<size_t>(d->n_trees) * 8);
        if (e == cudaSuccess) {
            qf::tree_max_leaf_kernel<<<d->n_trees, 256>>>(f->d_nodes, f->d_roots, f->n_nodes, f->T, d_max);
            std::vector<uint32_t> both(static_cast<size_t>(d->n_trees) * 2);
            e = cudaMemcpy(both.data(), d_max, both.size() * 4, cudaMemcpyDeviceToHost);
            cudaFree(d_max);
            f->tree_max_leaf.assign(both.begin(), both.begin() + d->n_trees);
            f->unit_leaves = true;
            for (int t = 0; t < d->n_trees; ++t)
                f->unit_leaves = f->unit_leaves && both[t] == 1u && both[d->n_trees + t] == 1u;
        }
        // largest total weight a row can gather: scales the fixed-point weights of QF_MODE_FAST
        uint32_t* d_w = nullptr;
        if (e == cudaSuccess) e = cudaMalloc(&d_w, static_cast<size_t>(d->n_trees) * 4);
        if (e == cudaSuccess) e = cudaMemset(d_w, 0, static_cast<size_t>(d->n_trees) * 4);
        if (e == cudaSuccess) {
            qf::tree_max_leaf_weight_kernel<<<d->n_trees, 256>>>(f->d_nodes, f->d_roots, f->n_nodes, f->T, f->d_members, d_w);
            std::vector<float> w(static_cast<size_t>(d->n_trees));
            e = cudaMemcpy(w.data(), d_w, w.size() * 4, cudaMemcpyDeviceToHost);
            double bound = 0.0;
            for (float v : w) bound += v <= 1.0f + 1e-5f ? 1.0 : static_cast<double>(v) * (1.0 + 1e-6);
            f->row_weight_bound = static_cast<int64_t>(std::ceil(bound));
        }
        cudaFree(d_w);
        if (e != cudaSuccess) rc = qf::fail(QF_ERR_CUDA, std::string("qf_forest_create: ") + cudaGetErrorString(e));
    }
    if (rc != QF_OK) {
        const std::string keep = qf::last_error_slot();
        qf_forest_destroy(f);
        return qf::fail(rc, keep);
    }
    *out = f;
    return QF_OK;
}

void qf_forest_destroy(qf_forest* f) {
    if (!f) return;
    DeviceGuard guard(f->device);
    cudaFree(f->d_nodes);
    cudaFree(f->d_roots);
    cudaFree(f->d_node_ids);
    cudaFree(f->d_members);
    cudaFree(f->d_y_sorted);
    cudaFree(f->d_values);
    cudaFree(f->d_impurity);
    cudaFree(f->d_nodes_c);
    cudaFree(f->d_lists);
    for (cudaEvent_t e : f->event_pool) cudaEventDestroy(e);
    {
        HostPipe& hp = f->pipe;
        for (int s = 0; s < kPipeSlots; ++s) {
            cudaFree(hp.x[s]);
            cudaFree(hp.out[s]);
            cudaFreeHost(hp.hx[s]);
            cudaFreeHost(hp.hout[s]);
            if (hp.x_ready[s]) cudaEventDestroy(hp.x_ready[s]);
            if (hp.computed[s]) cudaEventDestroy(hp.computed[s]);
            if (hp.out_free[s]) cudaEventDestroy(hp.out_free[s]);
        }
        cudaFree(hp.ws);
        if (hp.h2d) cudaStreamDestroy(hp.h2d);
        if (hp.compute) cudaStreamDestroy(hp.compute);
        if (hp.d2h) cudaStreamDestroy(hp.d2h);
    }
    delete f;
}

int64_t qf_forest_device_bytes(const qf_forest* f) { return f ? f->device_bytes : 0; }

int64_t qf_forest_workspace_bytes(const qf_forest* f, int64_t n_rows, int32_t n_quantiles) {
    (void)n_quantiles;
    if (!f || n_rows < 0) return 0;
    return make_plan(f, n_rows).total;
}

int qf_forest_plan(const qf_forest* f, int64_t n_rows, int32_t* out) {
    if (!f || !out || n_rows < 0) return qf::fail(QF_ERR_INVALID, "qf_forest_plan: bad argument");
    const Plan pl = make_plan(f, n_rows);
    out[0] = static_cast<int32_t>(std::min<int64_t>(pl.chunk_rows, INT32_MAX));
    out[1] = pl.warp_elems;
    out[2] = pl.s1_threads;
    out[3] = pl.s1_limit;
    out[4] = pl.s2_limit;
    out[5] = pl.dense_needed ? 1 : 0;
    out[6] = pl.sort_rows ? 1 : 0;
    out[7] = 0;
    return QF_OK;
}

int qf_forest_quantile_kernel_name(qf_forest* f, int64_t n_rows, int32_t nq, int32_t mode, char* buf, int32_t len) {
    if (!f || !buf || len <= 0 || n_rows < 0) return qf::fail(QF_ERR_INVALID, "qf_forest_quantile_kernel_name: bad argument");
    const Plan pl = make_plan(f, n_rows);
    char name[96];
    const bool select = mode == QF_MODE_FAST && nq <= qf::kSelectMaxQ && fast_mode_is_precise(f);
    if (pl.warp_elems) {
        std::snprintf(name, sizeof(name), "quantile_warp_kernel<%d>", pl.warp_elems);
    } else if (select) {
        Select3Geometry g3 = select3_geometry(f, nq);
        Select4Geometry g4 = select4_geometry(f, nq);
        if (g3.fx_scale > 0.0f || g4.fx_scale > 0.0f) {
            DeviceGuard guard(f->device);
            if (ensure_compact(f) == QF_OK) {
                if (g3.fx_scale > 0.0f) select3_finish_geometry(f, g3);
                if (g4.fx_scale > 0.0f) select4_finish_geometry(f, g4);
            }
        }
        int nbits = 0;
        while ((int64_t(1) << nbits) < f->N) ++nbits;
        if (g4.ok)
            std::snprintf(name, sizeof(name), "quantile_select4_kernel");
        else if (g3.ok)
            std::snprintf(name, sizeof(name), "quantile_select3_kernel<%d, %d, %d, %d>", g3.log2nb, g3.ipt, g3.selw, g3.stages);
        else if (f->select_impl != 1 && f->select_bits == qf::kSelectL2Bits && nbits <= 12 + qf::kSelect2SlotBits &&
                 f->T <= qf::kSelect2MaxIPT * qf::kSelect2Threads)
            std::snprintf(name, sizeof(name), "quantile_select2_kernel<%d>", nbits <= 11 + qf::kSelect2SlotBits ? 11 : 12);
        else
            std::snprintf(name, sizeof(name), "quantile_select_kernel<%d>", f->T <= 128 ? 128 : 256);
    } else {
        std::snprintf(name, sizeof(name), "quantile_bucket_kernel<%d>", pl.s1_threads);
    }
    std::snprintf(buf, static_cast<size_t>(len), "%s", name);
    return QF_OK;
}

int qf_forest_last_launches(const qf_forest* f) {
    if (!f) return 0;
    qf_forest* m = const_cast<qf_forest*>(f);
    std::lock_guard<std::mutex> g(m->mu);
    return m->last.launches;
}

int qf_forest_last_timings(qf_forest* f, double* apply_ms, double* quantile_ms) {
    if (!f) return qf::fail(QF_ERR_INVALID, "qf_forest_last_timings: null forest");
    DeviceGuard guard(f->device);
    double acc[3] = {0.0, 0.0, 0.0};
    std::vector<qf_forest::Span> spans;
    {
        std::lock_guard<std::mutex> g(f->mu);
        spans = f->last.spans;
    }
    for (const qf_forest::Span& s : spans) {
        QF_CUDA_TRY(cudaEventSynchronize(s.b));
        float ms = 0.0f;
        QF_CUDA_TRY(cudaEventElapsedTime(&ms, s.a, s.b));
        acc[s.kind] += ms;
    }
    if (apply_ms) *apply_ms = acc[SPAN_APPLY];
    if (quantile_ms) *quantile_ms = acc[SPAN_QUANTILE];
    return QF_OK;
}

int qf_forest_set_option(qf_forest* f, const char* key, int64_t value) {
    if (!f || !key) return qf::fail(QF_ERR_INVALID, "qf_forest_set_option: null argument");
    const std::string k(key);
    if (k == "sort_rows") {
        if (value < -1 || value > 1) return qf::fail(QF_ERR_INVALID, "sort_rows must be -1 (auto), 0 or 1");
        f->sort_rows = static_cast<int>(value);
    } else if (k == "apply_row_warps") {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8)
            return qf::fail(QF_ERR_INVALID, "apply_row_warps must be 0 (auto), 1, 2, 4 or 8");
        f->apply_row_warps = static_cast<int>(value);
    } else if (k == "apply_levels") {
        if (value < 1 || value > 4) return qf::fail(QF_ERR_INVALID, "apply_levels must be 1, 2, 3 or 4");
        f->apply_levels = static_cast<int>(value);
    } else if (k == "apply_ilp") {
        if (value != 1 && value != 2 && value != 4 && value != 8)
            return qf::fail(QF_ERR_INVALID, "apply_ilp must be 1, 2, 4 or 8");
        f->apply_ilp = static_cast<int>(value);
    } else if (k == "rows_per_chunk") {
        if (value < 1) return qf::fail(QF_ERR_INVALID, "rows_per_chunk must be >= 1");
        f->rows_per_chunk = value;
    } else if (k == "cta_capacity") {
        if (value < 0 || value > 8192) return qf::fail(QF_ERR_INVALID, "cta_capacity must be in [0, 8192]");
        f->cta_capacity = static_cast<int>(value);
    } else if (k == "cta2_capacity") {
        if (value < 0 || value > 8192) return qf::fail(QF_ERR_INVALID, "cta2_capacity must be in [0, 8192]");
        f->cta2_capacity = static_cast<int>(value);
    } else if (k == "warp_unit") {
        if (value != 0 && value != 1) return qf::fail(QF_ERR_INVALID, "warp_unit must be 0 or 1");
        f->warp_unit = static_cast<int>(value);
    } else if (k == "select_impl") {
        if (value < 0 || value > 4) return qf::fail(QF_ERR_INVALID, "select_impl must be 0 (auto), 1 (multi-level), 2 (select2), 3 (select3) or 4 (select4)");
        f->select_impl = static_cast<int>(value);
    } else if (k == "select_bits") {
        if (value < 1 || value > qf::kSelectL2Bits)
            return qf::fail(QF_ERR_INVALID, "select_bits must be in [1, 10]");
        f->select_bits = static_cast<int>(value);
    } else if (k == "key_bins") {
        if (value < qf::kKeyBins || value > qf::kKeyBinsMax || value % qf::kKeyBins != 0)
            return qf::fail(QF_ERR_INVALID, "key_bins must be a multiple of 8192 in [8192, 262144]");
        set_key_bins(f, static_cast<int>(value));
    } else if (k == "select4_hits") {
        if (value < 0 || value > qf::kSelect4Hits) return qf::fail(QF_ERR_INVALID, "select4_hits must be in [0, 256]");
        f->select4_hits = static_cast<int>(value);
    } else if (k == "select4_ctas") {
        if (value < 1 || value > 4) return qf::fail(QF_ERR_INVALID, "select4_ctas must be in [1, 4]");
        f->select4_ctas = static_cast<int>(value);
    } else if (k == "compact_align") {
        if (value != 0 && value != 1) return qf::fail(QF_ERR_INVALID, "compact_align must be 0 or 1");
        if (f->compact_state != 0 && value != f->compact_align)
            return qf::fail(QF_ERR_INVALID, "compact_align must be set before the compact lists are built");
        f->compact_align = static_cast<int>(value);
    } else if (k == "select4_hints") {
        if (value != 0 && value != 1) return qf::fail(QF_ERR_INVALID, "select4_hints must be 0 or 1");
        f->select4_hints = static_cast<int>(value);
    } else if (k == "select4_unroll") {
        if (value != 2 && value != 4) return qf::fail(QF_ERR_INVALID, "select4_unroll must be 2 or 4");
        f->select4_unroll = static_cast<int>(value);
    } else if (k == "select4_list") {
        if (value < 0 || value > 4096) return qf::fail(QF_ERR_INVALID, "select4_list must be in [0, 4096]");
        f->select4_list = static_cast<int>(value);
    } else if (k == "select3_variant") {
        if (value != 0 && value != 1) return qf::fail(QF_ERR_INVALID, "select3_variant must be 0 or 1");
        f->select3_variant = static_cast<int>(value);
    } else if (k == "select3_capw") {
        if (value < 0 || value > 4096) return qf::fail(QF_ERR_INVALID, "select3_capw must be in [0, 4096]");
        f->select3_capw = static_cast<int>(value);
    } else if (k == "warp_elems") {
        if (value != -1 && value != 0 && value != 2 && value != 4 && value != 8)
            return qf::fail(QF_ERR_INVALID, "warp_elems must be -1 (auto), 0 (off), 2, 4 or 8");
        f->warp_elems = static_cast<int>(value);
    } else if (k == "profile") {
        f->profile = value != 0;
    } else if (k == "dense_ctas") {
        if (value < 0) return qf::fail(QF_ERR_INVALID, "dense_ctas must be >= 0");
        f->dense_ctas = static_cast<int>(value);
    } else {
        return qf::fail(QF_ERR_INVALID, "qf_forest_set_option: unknown key " + k);
    }
    return QF_OK;
}

int qf_forest_apply(qf_forest* f, const float* X, int64_t n, int64_t ldx, int64_t* leaves,
                    void* ws, int64_t ws_bytes, void* stream) {
    int rc = check_forest_call(f, X, n, ldx, leaves, "qf_forest_apply");
    if (rc != QF_OK) return rc;
    CallScope scope(f);
    CallStats& cs = scope.cs;
    if (n == 0) return QF_OK;
    DeviceGuard guard(f->device);
    const Plan pl = make_plan(f, n);
    if (ws_bytes < pl.total || !ws) return qf::fail(QF_ERR_WORKSPACE, "qf_forest_apply: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t* idx = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(ws) + pl.off_seg);
    for (int64_t r0 = 0; r0 < n; r0 += pl.chunk_rows) {
        const int64_t rows = std::min(pl.chunk_rows, n - r0);
        const int* order = nullptr;
        rc = build_locality_order(f, cs, X + r0 * ldx, rows, ldx, pl, static_cast<unsigned char*>(ws), st, &order);
        if (rc != QF_OK) return rc;
        rc = launch_apply<qf::APPLY_EMIT_NODE>(f, cs, X + r0 * ldx, rows, ldx, idx, order, 0, f->T, st);
        if (rc != QF_OK) return rc;
        const int64_t total = rows * f->T;
        const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, f->sm_count * 16));
        qf::node_to_sklearn_id_kernel<<<blocks, 256, 0, st>>>(idx, total, f->T, f->d_roots, f->d_node_ids,
                                                             leaves + r0 * f->T);
        QF_CUDA_TRY(cudaGetLastError());
        ++cs.launches;
    }
    return QF_OK;
}

int qf_forest_predict_mean(qf_forest* f, const float* X, int64_t n, int64_t ldx, double* out,
                           void* ws, int64_t ws_bytes, void* stream) {
    int rc = check_forest_call(f, X, n, ldx, out, "qf_forest_predict_mean");
    if (rc != QF_OK) return rc;
    if (!f->d_values)
        return qf::fail(QF_ERR_INVALID, "qf_forest_predict_mean: forest was created without leaf_values");
    CallScope scope(f);
    CallStats& cs = scope.cs;
    if (n == 0) return QF_OK;
    DeviceGuard guard(f->device);
    const Plan pl = make_plan(f, n);
    if (ws_bytes < pl.total || !ws)
        return qf::fail(QF_ERR_WORKSPACE, "qf_forest_predict_mean: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t* idx = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(ws) + pl.off_seg);
    for (int64_t r0 = 0; r0 < n; r0 += pl.chunk_rows) {
        const int64_t rows = std::min(pl.chunk_rows, n - r0);
        const int* order = nullptr;
        rc = build_locality_order(f, cs, X + r0 * ldx, rows, ldx, pl, static_cast<unsigned char*>(ws), st, &order);
        if (rc != QF_OK) return rc;
        rc = launch_apply<qf::APPLY_EMIT_NODE>(f, cs, X + r0 * ldx, rows, ldx, idx, order, 0, f->T, st);
        if (rc != QF_OK) return rc;
        qf::mean_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(idx, rows, f->T, f->d_values,
                                                                                   out + r0);
        QF_CUDA_TRY(cudaGetLastError());
        ++cs.launches;
    }
    return QF_OK;
}

int qf_forest_set_leaf_impurity(qf_forest* f, const double* impurity) {
    if (!f || !impurity) return qf::fail(QF_ERR_INVALID, "qf_forest_set_leaf_impurity: null argument");
    DeviceGuard guard(f->device);
    if (!f->d_impurity) {
        QF_CUDA_TRY(cudaMalloc(&f->d_impurity, static_cast<size_t>(f->n_nodes) * 8));
        f->device_bytes += f->n_nodes * 8;
    }
    QF_CUDA_TRY(cudaMemcpy(f->d_impurity, impurity, static_cast<size_t>(f->n_nodes) * 8, cudaMemcpyDefault));
    return QF_OK;
}

int qf_forest_predict_mean_std(qf_forest* f, const float* X, int64_t n, int64_t ldx, double min_variance,
                               double* mean_out, double* std_out, void* ws, int64_t ws_bytes, void* stream) {
    int rc = check_forest_call(f, X, n, ldx, mean_out, "qf_forest_predict_mean_std");
    if (rc != QF_OK) return rc;
    if (!std_out) return qf::fail(QF_ERR_INVALID, "qf_forest_predict_mean_std: null output");
    if (!f->d_values || !f->d_impurity)
        return qf::fail(QF_ERR_INVALID, "qf_forest_predict_mean_std: forest has no leaf_values / leaf impurities");
    CallScope scope(f);
    CallStats& cs = scope.cs;
    if (n == 0) return QF_OK;
    DeviceGuard guard(f->device);
    const Plan pl = make_plan(f, n);
    if (ws_bytes < pl.total || !ws)
        return qf::fail(QF_ERR_WORKSPACE, "qf_forest_predict_mean_std: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t* idx = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(ws) + pl.off_seg);
    for (int64_t r0 = 0; r0 < n; r0 += pl.chunk_rows) {
        const int64_t rows = std::min(pl.chunk_rows, n - r0);
        const int* order = nullptr;
        rc = build_locality_order(f, cs, X + r0 * ldx, rows, ldx, pl, static_cast<unsigned char*>(ws), st, &order);
        if (rc != QF_OK) return rc;
        rc = launch_apply<qf::APPLY_EMIT_NODE>(f, cs, X + r0 * ldx, rows, ldx, idx, order, 0, f->T, st);
        if (rc != QF_OK) return rc;
        qf::mean_std_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(
            idx, rows, f->T, f->d_values, f->d_impurity, min_variance, mean_out + r0, std_out + r0);
        QF_CUDA_TRY(cudaGetLastError());
        ++cs.launches;
    }
    return QF_OK;
}

int qf_forest_predict_quantiles(qf_forest* f, const float* X, int64_t n, int64_t ldx,
                                const double* q_host, int32_t nq, double* out, int32_t mode,
                                void* ws, int64_t ws_bytes, void* stream) {
    int rc = check_forest_call(f, X, n, ldx, out, "qf_forest_predict_quantiles");
    if (rc != QF_OK) return rc;
    rc = check_quantiles(q_host, nq, mode, "qf_forest_predict_quantiles");
    if (rc != QF_OK) return rc;
    CallScope scope(f);
    CallStats& cs = scope.cs;
    if (n == 0) return QF_OK;
    DeviceGuard guard(f->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Plan pl = make_plan(f, n);
    if (ws_bytes < pl.total || !ws)
        return qf::fail(QF_ERR_WORKSPACE, "qf_forest_predict_quantiles: workspace too small");
    rc = zero_dense_scratch(f, pl, ws, st);
    if (rc != QF_OK) return rc;
    return quantiles_on_stream(f, cs, X, n, ldx, q_host, nq, out, mode, pl, ws, ws_bytes, st);
}

int qf_forest_predict_quantiles_host(qf_forest* f, const float* X_host, int64_t n, int64_t ldx,
                                     const double* q_host, int32_t nq, double* out_host, int32_t mode,
                                     int64_t chunk_rows) {
    int rc = check_forest_call(f, X_host, n, ldx, out_host, "qf_forest_predict_quantiles_host");
    if (rc != QF_OK) return rc;
    rc = check_quantiles(q_host, nq, mode, "qf_forest_predict_quantiles_host");
    if (rc != QF_OK) return rc;
    CallScope scope(f);
    CallStats& cs = scope.cs;
    if (n == 0) return QF_OK;
    DeviceGuard guard(f->device);
    // measured at 100k x 20 rows: splitting below rows_per_chunk only adds per-chunk launches
    // (1 chunk 0.74 ms, 2 chunks 0.76 ms, 4 chunks 0.91 ms); the copies overlap the kernels of
    // neighbouring chunks once there are several full-size chunks
    if (chunk_rows <= 0) chunk_rows = f->rows_per_chunk;
    chunk_rows = std::min(chunk_rows, n);
    const Plan pl = make_plan(f, chunk_rows, chunk_rows);  // one kernel chunk per pipeline slot
    std::lock_guard<std::mutex> pipe_guard(f->pipe_mu);    // the pipeline buffers are per handle
    const int64_t x_bytes = chunk_rows * f->F * 4, out_bytes = chunk_rows * nq * 8;
    HostPipe& hp = f->pipe;
    if (!hp.h2d) {
        QF_CUDA_TRY(cudaStreamCreateWithFlags(&hp.h2d, cudaStreamNonBlocking));
        QF_CUDA_TRY(cudaStreamCreateWithFlags(&hp.compute, cudaStreamNonBlocking));
        QF_CUDA_TRY(cudaStreamCreateWithFlags(&hp.d2h, cudaStreamNonBlocking));
        for (int s = 0; s < kPipeSlots; ++s) {
            QF_CUDA_TRY(cudaEventCreateWithFlags(&hp.x_ready[s], cudaEventDisableTiming));
            QF_CUDA_TRY(cudaEventCreateWithFlags(&hp.computed[s], cudaEventDisableTiming));
            QF_CUDA_TRY(cudaEventCreateWithFlags(&hp.out_free[s], cudaEventDisableTiming));
        }
    }
    if (hp.x_bytes < x_bytes) {
        for (int s = 0; s < kPipeSlots; ++s) { cudaFree(hp.x[s]); hp.x[s] = nullptr; }
        hp.x_bytes = 0;
        for (int s = 0; s < kPipeSlots; ++s) QF_CUDA_TRY(cudaMalloc(&hp.x[s], x_bytes));
        hp.x_bytes = x_bytes;
    }
    if (hp.out_bytes < out_bytes) {
        for (int s = 0; s < kPipeSlots; ++s) { cudaFree(hp.out[s]); hp.out[s] = nullptr; }
        hp.out_bytes = 0;
        for (int s = 0; s < kPipeSlots; ++s) QF_CUDA_TRY(cudaMalloc(&hp.out[s], out_bytes));
        hp.out_bytes = out_bytes;
    }
    if (hp.ws_bytes < pl.total) {
        cudaFree(hp.ws); hp.ws = nullptr; hp.ws_bytes = 0;
        QF_CUDA_TRY(cudaMalloc(&hp.ws, pl.total)); hp.ws_bytes = pl.total;
    }
    rc = zero_dense_scratch(f, pl, hp.ws, hp.compute);
    if (rc != QF_OK) return rc;
    // pageable buffers go through page-locked staging slots (several chunks only: one chunk has nothing to overlap)
    const bool stage_x = n > chunk_rows && is_pageable(X_host);
    const bool stage_out = n > chunk_rows && is_pageable(out_host);
    if (stage_x && hp.hx_bytes < x_bytes) {
        for (int s = 0; s < kPipeSlots; ++s) { cudaFreeHost(hp.hx[s]); hp.hx[s] = nullptr; }
        hp.hx_bytes = 0;
        for (int s = 0; s < kPipeSlots; ++s) QF_CUDA_TRY(cudaMallocHost(&hp.hx[s], x_bytes));
        hp.hx_bytes = x_bytes;
    }
    if (stage_out && hp.hout_bytes < out_bytes) {
        for (int s = 0; s < kPipeSlots; ++s) { cudaFreeHost(hp.hout[s]); hp.hout[s] = nullptr; }
        hp.hout_bytes = 0;
        for (int s = 0; s < kPipeSlots; ++s) QF_CUDA_TRY(cudaMallocHost(&hp.hout[s], out_bytes));
        hp.hout_bytes = out_bytes;
    }
    // result of the chunk that used slot s before: wait for its D2H, then hand it to the caller
    auto drain = [&](int64_t done_chunk) -> int {
        const int s = static_cast<int>(done_chunk % kPipeSlots);
        QF_CUDA_TRY(cudaEventSynchronize(hp.out_free[s]));
        const int64_t q0 = done_chunk * chunk_rows, rows = std::min(chunk_rows, n - q0);
        std::memcpy(out_host + q0 * nq, hp.hout[s], static_cast<size_t>(rows) * nq * 8);
        return QF_OK;
    };
    int64_t chunk = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, ++chunk) {
        const int64_t rows = std::min(chunk_rows, n - r0);
        const int s = static_cast<int>(chunk % kPipeSlots);
        if (chunk >= kPipeSlots) {
            if (stage_out) {
                rc = drain(chunk - kPipeSlots);
                if (rc != QF_OK) return rc;
            } else if (stage_x) {                       // the H2D out of hx[s] is long done once that chunk's D2H is
                QF_CUDA_TRY(cudaEventSynchronize(hp.out_free[s]));
            }
        }
        const float* src = X_host + r0 * ldx;
        int64_t src_ld = ldx;
        if (stage_x) {                                  // host-side copy of chunk k runs while the GPU is on chunk k - 1
            if (ldx == f->F) {
                std::memcpy(hp.hx[s], src, static_cast<size_t>(rows) * f->F * 4);
            } else {
                for (int64_t r = 0; r < rows; ++r)
                    std::memcpy(hp.hx[s] + r * f->F, src + r * ldx, static_cast<size_t>(f->F) * 4);
            }
            src = hp.hx[s];
            src_ld = f->F;
        }
        // H2D: the slot's X buffer is free once the kernels of the chunk that used it are done.
        // Rows of X are compacted to F columns on the device (ldx may exceed F on the host).
        if (chunk >= kPipeSlots) QF_CUDA_TRY(cudaStreamWaitEvent(hp.h2d, hp.computed[s], 0));
        if (src_ld == f->F) {
            QF_CUDA_TRY(cudaMemcpyAsync(hp.x[s], src, rows * f->F * 4, cudaMemcpyHostToDevice, hp.h2d));
        } else {
            QF_CUDA_TRY(cudaMemcpy2DAsync(hp.x[s], f->F * 4, src, src_ld * 4, f->F * 4, rows,
                                          cudaMemcpyHostToDevice, hp.h2d));
        }
        QF_CUDA_TRY(cudaEventRecord(hp.x_ready[s], hp.h2d));
        // kernels: one stream, chunk after chunk
        QF_CUDA_TRY(cudaStreamWaitEvent(hp.compute, hp.x_ready[s], 0));
        if (chunk >= kPipeSlots) QF_CUDA_TRY(cudaStreamWaitEvent(hp.compute, hp.out_free[s], 0));
        rc = quantiles_on_stream(f, cs, hp.x[s], rows, f->F, q_host, nq, hp.out[s], mode, pl, hp.ws, hp.ws_bytes, hp.compute);
        if (rc != QF_OK) return rc;
        QF_CUDA_TRY(cudaEventRecord(hp.computed[s], hp.compute));
        // D2H
        QF_CUDA_TRY(cudaStreamWaitEvent(hp.d2h, hp.computed[s], 0));
        QF_CUDA_TRY(cudaMemcpyAsync(stage_out ? hp.hout[s] : out_host + r0 * nq, hp.out[s], rows * nq * 8,
                                    cudaMemcpyDeviceToHost, hp.d2h));
        QF_CUDA_TRY(cudaEventRecord(hp.out_free[s], hp.d2h));
    }
    if (stage_out) {
        for (int64_t c = std::max<int64_t>(0, chunk - kPipeSlots); c < chunk; ++c) {
            rc = drain(c);
            if (rc != QF_OK) return rc;
        }
    }
    QF_CUDA_TRY(cudaStreamSynchronize(hp.d2h));
    QF_CUDA_TRY(cudaStreamSynchronize(hp.compute));
    return QF_OK;
}

// K4: leaf -> member lists on the device (csr_builder.cuh).  Synchronous.
int qf_build_csr(const qf_csr_build_desc* d, int64_t* n_members_out) {
    if (!d || !n_members_out) return qf::fail(QF_ERR_INVALID, "qf_build_csr: null argument");
    if (d->n_trees <= 0 || d->n_features <= 0 || d->n_train <= 0 || d->n_nodes <= 0 || !d->nodes_dev ||
        !d->tree_offsets || !d->X_train_dev || !d->rank_dev || !d->members_dev || !d->inbag_out ||
        !d->max_leaf_members_out || !d->sum_sq_out)
        return qf::fail(QF_ERR_INVALID, "qf_build_csr: empty forest or null array");
    if (d->n_nodes >= (int64_t(1) << 31) || d->n_train >= (int64_t(1) << 31))
        return qf::fail(QF_ERR_LIMIT, "qf_build_csr: forest exceeds the 32-bit packed layout");
    int count = 0;
    QF_CUDA_TRY(cudaGetDeviceCount(&count));
    if (d->device < 0 || d->device >= count) return qf::fail(QF_ERR_CUDA, "qf_build_csr: no such CUDA device");
    DeviceGuard guard(d->device);
    if (!guard.ok) return qf::fail(QF_ERR_CUDA, "qf_build_csr: cudaSetDevice failed");
    cudaDeviceProp prop;
    QF_CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major != 10)
        return qf::fail(QF_ERR_CUDA, "qf_build_csr: device is not sm_100 (this library has sm_100a code only)");

    const int T = d->n_trees;
    const int64_t N = d->n_train, n_nodes = d->n_nodes, total = N * T;
    const int64_t n_leaves = n_nodes / 2 + 1;            // a tree of k packed slots has k / 2 leaves
    const int64_t n_blocks = (n_leaves + qf::kCsrScanBlock - 1) / qf::kCsrScanBlock;
    // a minimal forest object for the leaf-lookup launcher (no member lists yet)
    qf_forest f;
    f.device = d->device;
    f.T = T;
    f.F = d->n_features;
    f.fbits = d->feature_bits;
    f.N = N;
    f.n_nodes = n_nodes;
    f.sm_count = prop.multiProcessorCount;
    f.d_nodes = reinterpret_cast<uint2*>(d->nodes_dev);

    struct Scratch {
        void* p[11] = {};
        ~Scratch() { for (void* q : p) cudaFree(q); }
    } sc;
    auto alloc = [&](int k, size_t bytes) -> cudaError_t { return cudaMalloc(&sc.p[k], bytes ? bytes : 16); };
    QF_CUDA_TRY(alloc(0, static_cast<size_t>(T) * 4));                      // roots
    QF_CUDA_TRY(alloc(1, static_cast<size_t>(total) * 4));                  // leaf node of every (sample, tree)
    QF_CUDA_TRY(alloc(2, static_cast<size_t>(n_nodes) * 4));                // leaf_cnt
    QF_CUDA_TRY(alloc(3, static_cast<size_t>(n_nodes) * 4));                // leaf_sum
    QF_CUDA_TRY(alloc(4, static_cast<size_t>(n_nodes) * 4));                // start
    QF_CUDA_TRY(alloc(5, static_cast<size_t>(n_blocks + 1) * 8));           // block sums + total
    QF_CUDA_TRY(alloc(6, static_cast<size_t>((n_nodes + T) / 2 + 2) * 4 + 16));   // long lists (<= one per leaf: a tree of k nodes has (k + 1) / 2) + counters
    QF_CUDA_TRY(alloc(7, static_cast<size_t>(T + 1) * 8));                  // tree offsets
    QF_CUDA_TRY(alloc(8, static_cast<size_t>(T) * 24));                     // per-tree stats
    QF_CUDA_TRY(alloc(9, static_cast<size_t>(n_leaves) * 4));               // member counts in left-to-right leaf order
    QF_CUDA_TRY(alloc(10, static_cast<size_t>(n_leaves) * 4));              // packed node of every leaf ordinal
    uint32_t* d_roots = static_cast<uint32_t*>(sc.p[0]);
    int32_t* d_idx = static_cast<int32_t*>(sc.p[1]);
    uint32_t* leaf_cnt = static_cast<uint32_t*>(sc.p[2]);
    uint32_t* leaf_sum = static_cast<uint32_t*>(sc.p[3]);
    uint32_t* start = static_cast<uint32_t*>(sc.p[4]);
    uint64_t* block_sum = static_cast<uint64_t*>(sc.p[5]);
    int* counters = static_cast<int*>(sc.p[6]);                             // [0] long lists, [1] too long
    int32_t* big = reinterpret_cast<int32_t*>(counters + 4);
    int64_t* d_offsets = static_cast<int64_t*>(sc.p[7]);
    unsigned long long* d_inbag = static_cast<unsigned long long*>(sc.p[8]);
    uint32_t* d_maxleaf = reinterpret_cast<uint32_t*>(d_inbag + T);
    double* d_sumsq = reinterpret_cast<double*>(d_inbag + 2 * T);
    uint32_t* cnt_ord = static_cast<uint32_t*>(sc.p[9]);
    int32_t* node_of = static_cast<int32_t*>(sc.p[10]);

    std::vector<uint32_t> roots(T);
    for (int t = 0; t < T; ++t) roots[t] = static_cast<uint32_t>(d->tree_offsets[t]);
    QF_CUDA_TRY(cudaMemcpy(d_roots, roots.data(), static_cast<size_t>(T) * 4, cudaMemcpyHostToDevice));
    QF_CUDA_TRY(cudaMemcpy(d_offsets, d->tree_offsets, static_cast<size_t>(T + 1) * 8, cudaMemcpyHostToDevice));
    f.d_roots = d_roots;
    QF_CUDA_TRY(cudaMemset(leaf_cnt, 0, static_cast<size_t>(n_nodes) * 4));
    QF_CUDA_TRY(cudaMemset(leaf_sum, 0, static_cast<size_t>(n_nodes) * 4));
    QF_CUDA_TRY(cudaMemset(counters, 0, 16));
    QF_CUDA_TRY(cudaMemset(start, 0, static_cast<size_t>(n_nodes) * 4));
    QF_CUDA_TRY(cudaMemset(cnt_ord, 0, static_cast<size_t>(n_leaves) * 4));
    QF_CUDA_TRY(cudaMemset(node_of, 0xFF, static_cast<size_t>(n_leaves) * 4));

    cudaStream_t st = nullptr;
    int rc = QF_OK;
    CallStats build_stats;
    // tree.py:101 for every tree: leaf of every training row (rows in their own order)
    const int64_t chunk = 131072;
    for (int64_t r0 = 0; r0 < N && rc == QF_OK; r0 += chunk) {
        const int64_t rows = std::min(chunk, N - r0);
        rc = launch_apply<qf::APPLY_EMIT_NODE>(&f, build_stats, d->X_train_dev + r0 * d->n_features, rows, d->n_features,
                                                d_idx + r0 * T, nullptr, 0, T, st);
    }
    f.d_nodes = nullptr;                                 // not owned
    f.d_roots = nullptr;
    if (rc != QF_OK) return rc;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, int64_t(prop.multiProcessorCount) * 32));
    qf::csr_count_kernel<<<grid, 256, 0, st>>>(d_idx, d->counts_dev, total, leaf_cnt, leaf_sum);
    uint2* nodes = reinterpret_cast<uint2*>(d->nodes_dev);
    qf::csr_leaf_order_kernel<<<static_cast<unsigned>((n_nodes + 255) / 256), 256, 0, st>>>(nodes, n_nodes, leaf_cnt, n_leaves,
                                                                                          cnt_ord, node_of, counters + 2);
    qf::csr_block_sum_kernel<<<static_cast<unsigned>(n_blocks), 256, 0, st>>>(cnt_ord, n_leaves, block_sum);
    qf::csr_scan_sums_kernel<<<1, 1024, 0, st>>>(block_sum, n_blocks, block_sum + n_blocks);
    QF_CUDA_TRY(cudaGetLastError());
    uint64_t n_members = 0;
    QF_CUDA_TRY(cudaMemcpy(&n_members, block_sum + n_blocks, 8, cudaMemcpyDeviceToHost));
    int bad_ordinal = 0;
    QF_CUDA_TRY(cudaMemcpy(&bad_ordinal, counters + 2, 4, cudaMemcpyDeviceToHost));
    if (bad_ordinal)
        return qf::fail(QF_ERR_INVALID, "qf_build_csr: leaves do not carry their left-to-right ordinal "
                                        "(nodes must come from qf_pack_tree with train_leaves == NULL and member_base = slot / 2)");
    if (n_members > 0xFFFFFFFFull)
        return qf::fail(QF_ERR_LIMIT, "qf_build_csr: more than 2^32-1 member entries in the forest");
    if (static_cast<int64_t>(n_members) > d->members_capacity)
        return qf::fail(QF_ERR_INVALID, "qf_build_csr: members_capacity too small");
    uint2* members = reinterpret_cast<uint2*>(d->members_dev);
    qf::csr_leaf_record_kernel<<<static_cast<unsigned>(n_blocks), 256, 0, st>>>(nodes, cnt_ord, node_of, n_leaves, block_sum, start);
    qf::csr_scatter_kernel<<<grid, 256, 0, st>>>(d_idx, d->counts_dev, total, T, d->rank_dev, start, leaf_cnt, leaf_sum, members);
    qf::csr_sort_small_kernel<<<static_cast<unsigned>((n_nodes + 255) / 256), 256, 0, st>>>(nodes, n_nodes, members, big,
                                                                                          counters, counters + 1);
    QF_CUDA_TRY(cudaGetLastError());
    int host_counters[2] = {0, 0};
    QF_CUDA_TRY(cudaMemcpy(host_counters, counters, 8, cudaMemcpyDeviceToHost));
    if (host_counters[1])
        return qf::fail(QF_ERR_LIMIT, "qf_build_csr: a leaf holds more than 4096 members; use the host flattener");
    if (host_counters[0] > 0)
        qf::csr_sort_big_kernel<<<static_cast<unsigned>(host_counters[0]), 256, 0, st>>>(nodes, big, members);
    qf::csr_tree_stats_kernel<<<T, 256, 0, st>>>(nodes, d_offsets, d_inbag, d_maxleaf, d_sumsq);
    QF_CUDA_TRY(cudaGetLastError());
    std::vector<unsigned long long> h_inbag(T);
    std::vector<uint32_t> h_max(T);
    QF_CUDA_TRY(cudaMemcpy(h_inbag.data(), d_inbag, static_cast<size_t>(T) * 8, cudaMemcpyDeviceToHost));
    QF_CUDA_TRY(cudaMemcpy(h_max.data(), d_maxleaf, static_cast<size_t>(T) * 4, cudaMemcpyDeviceToHost));
    QF_CUDA_TRY(cudaMemcpy(d->sum_sq_out, d_sumsq, static_cast<size_t>(T) * 8, cudaMemcpyDeviceToHost));
    for (int t = 0; t < T; ++t) {
        d->inbag_out[t] = static_cast<int64_t>(h_inbag[t]);
        d->max_leaf_members_out[t] = static_cast<int32_t>(h_max[t]);
    }
    *n_members_out = static_cast<int64_t>(n_members);
    return QF_OK;
}

}  // extern "C"
