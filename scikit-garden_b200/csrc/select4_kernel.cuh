// K2, QF_MODE_FAST, WARP per row on compact leaf lists (compact_lists.cuh) -- fused leaf co-membership
// weighting (ensemble.py:137-140) + weighted percentile (utils.py:46-81) by histogram selection.
//
// Same arithmetic as select3_kernel.cuh (32-bit fixed-point member weight = count * U, U in the octet
// header; exact integer sums; utils.py:72-81 in float32 on the three entries around the crossing), so
// the two kernels return the same bits.  What is different is the shape: the round-2 profiles of the
// CTA-per-row kernels (select2, select3: DESIGN.md section 3) say they are bound by shared memory --
// 60-100 KB of it per row in flight (2-4 rows and ~25 warps per SM) and 2-3 k shared-memory wavefronts
// per row -- and by the block barriers between their phases.  Here a row belongs to ONE warp:
//
//   state    shared memory per row: a 1024-bucket histogram (4 KB, re-used as the hit list) and the
//            flat list of the row's octets (3.7 KB at config 3); 24 rows in flight per SM, each on its
//            own warp, no block barrier anywhere (only __syncwarp).
//   members  read straight from global memory / L2, one 256-bit load per octet (predicated: a lane
//            without an octet loads nothing), no staging.  The octet list (eight leaf words per lane
//            loaded together, one warp scan per 256 leaves) makes every pass one balanced loop of
//            32 * UNR octets per step.  The second pass walks the list in the REVERSE order of the
//            first, so that what it re-reads first is what the L2 saw last.
//   window   the members of a row sit in a narrow band of the rank axis (the conditional distribution
//            of y).  The first 128 octets of the list give [lo, hi); the 1024 buckets cover that band only
//            (bucket = clamp((rank - lo) >> s)), ranks outside fall into the two edge buckets.
//   pass 1   one shared-memory RED per member into the histogram.
//   scan     lane totals of 32 consecutive buckets (rotated 128-bit reads), one warp scan; per
//            percentile the group that holds the crossing is scanned across the lanes: crossing bucket
//            bq, nearest non-empty neighbours bp / bn from ballots.  The buckets between bp and bn
//            are empty, so "member of a marked bucket" is one rank interval per percentile.
//   pass 2   members of marked buckets (a few dozen of a row's 3300 at config 3) are appended to the
//            hit list (shared-memory counter; their order does not matter, see select).
//   select   the hits are sorted by rank in registers (bitonic network, 4 keys per lane; 8 beyond 128 hits), weights
//            prefix-summed in that order, and lane q finds percentile q's crossing entry and its
//            neighbours by binary search: because every member of bp, bq and bn is in the list, the
//            neighbours of the crossing entry are simply the adjacent rank groups of the sorted list.
//   overflow a row with more than 256 hits narrows each percentile exactly instead: sub-histograms of
//            the crossing bucket until it is one rank wide, then two more passes for the neighbours and
//            the merged weights.  Slow (a dozen passes) and rare; tests force it with `select4_hits`.
//
// Measured at config 3 ("select4"): 41 ms per 1M rows for the first, fully unrolled
// version (instruction fetch bound: 17.8 k SASS instructions), 21 ms for this one -- level with select2 and
// select3, bound by the random 64-byte DRAM accesses of TWO passes (75 MB of octets in flight between them do
// not stay in the L2).  Opt-in: `select_impl = 4`.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "compact_lists.cuh"
#include "quantile_kernel.cuh"
#include "select2_kernel.cuh"
#include "select3_kernel.cuh"

namespace qf {

constexpr int kSelect4Threads = 256;
constexpr int kSelect4Warps = kSelect4Threads / 32;     // rows in flight per CTA
constexpr int kSelect4CtasPerSm = 4;
constexpr int kSelect4LogNB = 10;
constexpr int kSelect4NB = 1 << kSelect4LogNB;          // buckets of the row's window of the rank axis
constexpr int kSelect4Hits = 256;                       // hit list entries per row (8 keys per lane)
constexpr int kSelect4MaxTrees = 4096;

struct Select4Params {
    const uint32_t* __restrict__ seg;      // [n][ldT] compact list words, ROW-major
    int64_t ldT;
    const int* __restrict__ order;         // [n] locality position -> output row; nullptr = identity
    const uint4* __restrict__ lists;       // two uint4 per octet (32-byte aligned)
    const float* __restrict__ y_sorted;    // [N]
    double* __restrict__ out;              // [n][ldo], indexed by output row
    int64_t n;
    int64_t ldo;
    int T;
    int Q;                                 // <= kSelect2QB
    float fx_scale;
    uint32_t n_train;
    int list_cap;                          // entries of the per-warp list of the row's octets
    int hits_cap;                          // <= kSelect4Hits
};

__host__ __device__ inline size_t select4_warp_smem(int list_cap) {
    return static_cast<size_t>(kSelect4NB) * 4 + static_cast<size_t>(list_cap) * 4 + 32;
}

// One octet = one 256-bit load, issued only by the lanes that have an octet (`ok`); the others get zeros, which
// read as seven padding slots (count 0) under a zero unit weight.  HINT 1 / 2: bypass L1 and tell the L2 to
// keep the line (first touch of a row's octets) / to drop it first (their last touch): the rows in flight
// hold 50-100 MB of octets between their two passes.
#define QF_LD_OCTET(HINTSTR)                                                                                    \
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %9, 0;\n\t"                                                      \
        "@p ld.global.nc" HINTSTR ".v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t}"                          \
        : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])       \
        : "l"(src), "r"(static_cast<uint32_t>(ok)))
template <int HINT>
__device__ __forceinline__ void ld_octet(const uint4* __restrict__ lists, uint32_t octet, bool ok, uint32_t (&o)[8]) {
    const uint4* src = lists + 2 * static_cast<size_t>(octet);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0u;
    if (HINT == 1) QF_LD_OCTET(".L1::no_allocate.L2::evict_last");
    else if (HINT == 2) QF_LD_OCTET(".L1::no_allocate.L2::evict_first");
    else QF_LD_OCTET("");
}
#undef QF_LD_OCTET

template <int H>
struct Select4Hint { static constexpr int value = H; };

// The hits of a row (hit_r[v] = rank, hit_w[v] = weight, v < nh <= 32 E) sorted by rank, in place: hit_r
// becomes the sorted ranks (0xffffff in the unused slots), hit_w the running sums of the weights in that order.
template <int E>
__device__ __forceinline__ void select4_sort_hits(uint32_t* hit_r, uint32_t* hit_w, uint32_t nh, int lane) {
    uint32_t key[E], wv[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const uint32_t v = static_cast<uint32_t>(lane * E + i);
        key[i] = v < nh ? ((hit_r[v] << 8) | v) : 0xffffffffu;
    }
    warp_bitonic_sort<E>(key, lane);
    uint32_t run = 0u;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        wv[i] = key[i] != 0xffffffffu ? hit_w[key[i] & 0xffu] : 0u;
        run += wv[i];
        wv[i] = run;
    }
    uint32_t lincl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, lincl, d);
        if (lane >= d) lincl += u;
    }
    __syncwarp();                                       // every lane has read the unsorted list
#pragma unroll
    for (int i = 0; i < E; ++i) {
        hit_r[lane * E + i] = key[i] >> 8;
        hit_w[lane * E + i] = wv[i] + lincl - run;
    }
    __syncwarp();
}

// UNR: octets a lane has in flight per step of a pass.  The loops are deliberately NOT unrolled further: the
// first versions of this kernel (17.8 k and 8.7 k SASS instructions, mostly straight-line) ran at one warp
// instruction per SM per cycle whatever the occupancy -- 24-32 warps at different places of 50+ KB of hot
// code live on instruction fetches from the L2 (L0 ~6 KB, L1.5 32 KB).
template <int UNR, bool HINTS>
__global__ void __launch_bounds__(kSelect4Threads, kSelect4CtasPerSm)
quantile_select4_kernel(const Select4Params p, const __grid_constant__ QuantileList ql) {
    constexpr int NB = kSelect4NB;
    constexpr uint32_t NONE = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const size_t wbytes = select4_warp_smem(p.list_cap);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + warp * wbytes);   // [NB], zero between rows
    uint32_t* hit_r = hist;                                                   // [256] ranks -> sorted ranks   } in the histogram's
    uint32_t* hit_w = hist + kSelect4Hits;                                    // [256] weights -> running sums } place, after the scan
    uint32_t* olist = hist + NB;                                              // [list_cap] the row's octets
    uint32_t* n_hits = olist + p.list_cap;                                    // hit counter, zero between rows
    const uint32_t hist_s = static_cast<uint32_t>(__cvta_generic_to_shared(hist));
    const float inv = __fdiv_rn(1.0f, p.fx_scale);
    const int n_slots = (p.T + 31) >> 5;                // leaves per lane

    for (int i = lane; i < NB; i += 32) hist[i] = 0u;
    if (lane == 0) *n_hits = 0u;
    __syncwarp();

    const int64_t n_warps = static_cast<int64_t>(gridDim.x) * kSelect4Warps;
    for (int64_t pos = static_cast<int64_t>(blockIdx.x) * kSelect4Warps + warp; pos < p.n; pos += n_warps) {
        const uint32_t* __restrict__ words = p.seg + pos * p.ldT;             // leaf t of this row: words[t]
        // first member octet and octet count of a leaf (a long list starts with an info octet)
        auto leaf = [&](int t, uint32_t& first, uint32_t& noct) {
            const uint32_t word = t < p.T ? __ldg(words + t) : kCompactEmpty;
            first = word & kCompactOffMask;
            noct = word == kCompactEmpty ? 0u : (word >> 28) + 1u;
            if (noct > static_cast<uint32_t>(kCompactMaxOctets)) {
                noct = __ldg(p.lists + 2 * static_cast<size_t>(first)).x;
                ++first;
            }
        };
        // ---- the row's octets -> flat list: every pass is then one balanced loop over it.  Eight leaves per
        // lane at a time (their words are loaded together), one warp scan per group of 256 leaves.
        uint32_t n_oct = 0u;
#pragma unroll 1
        for (int i0 = 0; i0 < n_slots; i0 += 8) {
            uint32_t first[8], noct[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int t = (i0 + k) * 32 + lane;
                const uint32_t word = t < p.T ? __ldg(words + t) : kCompactEmpty;
                first[k] = word & kCompactOffMask;
                noct[k] = word == kCompactEmpty ? 0u : (word >> 28) + 1u;
            }
            uint32_t mine = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (noct[k] > static_cast<uint32_t>(kCompactMaxOctets)) {       // long list: info octet first
                    noct[k] = __ldg(p.lists + 2 * static_cast<size_t>(first[k])).x;
                    ++first[k];
                }
                mine += noct[k];
            }
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            const uint32_t group = __shfl_sync(0xffffffffu, incl, 31);
            if (n_oct + group <= static_cast<uint32_t>(p.list_cap)) {           // (once false it stays false)
                uint32_t at = n_oct + incl - mine;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    for (uint32_t j = 0; j < noct[k]; ++j) olist[at++] = first[k] + j;
            }
            n_oct += group;
        }
        __syncwarp();
        const bool flat = n_oct <= static_cast<uint32_t>(p.list_cap);

        // f(ok, o) on octets [j0, j0 + 32 UNR) of the list, all 32 lanes in step (f may shuffle); ok = false:
        // a lane without an octet in this step (it sees octet 0 of the lists, to be ignored)
        auto step = [&](auto hint, uint32_t j0, auto&& f) {
            constexpr int H = HINTS ? decltype(hint)::value : 0;
            uint32_t o[UNR][8];
            bool ok[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const uint32_t j = j0 + k * 32 + lane;
                ok[k] = j < n_oct;
                ld_octet<H>(p.lists, ok[k] ? olist[j] : 0u, ok[k], o[k]);
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) f(ok[k], o[k]);
        };
        // f on every octet of the row; `reverse` walks them in the opposite order
        auto each_octet = [&](auto hint, bool reverse, auto&& f) {
            constexpr int H = HINTS ? decltype(hint)::value : 0;
            if (flat) {
                const uint32_t steps = (n_oct + 32u * UNR - 1u) / (32u * UNR);
#pragma unroll 1
                for (uint32_t si = 0; si < steps; ++si) step(hint, (reverse ? steps - 1u - si : si) * (32u * UNR), f);
            } else {                                    // list too small for this row: leaf by leaf, lanes in step
#pragma unroll 1
                for (int i = 0; i < n_slots; ++i) {
                    uint32_t first, noct;
                    leaf(i * 32 + lane, first, noct);
                    const uint32_t most = __reduce_max_sync(0xffffffffu, noct);
#pragma unroll 1
                    for (uint32_t k = 0; k < most; ++k) {
                        const bool ok = k < noct;
                        uint32_t o[8];
                        ld_octet<H>(p.lists, ok ? first + k : 0u, ok, o);
                        f(ok, o);
                    }
                }
            }
        };

        // ---- the row's window of the rank axis, from the first octets of its list (widened by a quarter each side)
        uint32_t lo = NONE, wsh = 0u;                 // window: first rank, log2(ranks per bucket)
        {
            uint32_t hi = 0u;
            auto sample = [&](bool ok, const uint32_t (&o)[8]) {
#pragma unroll
                for (int e = 1; e < 8; ++e) {
                    const bool real = ok && o[0] != 0u && (o[e] >> 24) != 0u;    // (a real octet has a non-zero unit weight)
                    lo = min(lo, real ? (o[e] & 0xffffffu) : NONE);
                    hi = max(hi, real ? (o[e] & 0xffffffu) : 0u);
                }
            };
            if (flat) {
                step(Select4Hint<1>{}, 0u, sample);
                if (UNR < 4 && n_oct > 32u * UNR) step(Select4Hint<1>{}, 32u * UNR, sample);
            } else {                                    // the first octet of the lane's first leaf
                uint32_t first, noct;
                leaf(lane, first, noct);
                uint32_t o[8];
                ld_octet<0>(p.lists, noct ? first : 0u, noct != 0u, o);
                sample(noct != 0u, o);
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lo > hi) { lo = 0u; hi = p.n_train - 1u; }          // nothing in the sample
            const uint32_t margin = (hi - lo) >> 2;
            lo = lo > margin ? lo - margin : 0u;
            hi = min(hi + margin, p.n_train - 1u);
            const int bits = 32 - __clz(hi - lo);                   // hi - lo < 2^bits
            wsh = static_cast<uint32_t>(max(bits - kSelect4LogNB, 0));
        }
        // bucket of a rank: monotone, buckets 0 and NB - 1 also take what lies outside the window
        auto bucket_of = [&](uint32_t rank) -> uint32_t { return min((max(rank, lo) - lo) >> wsh, static_cast<uint32_t>(NB - 1)); };
        auto bucket_lo = [&](uint32_t b) -> uint32_t { return b == 0u ? 0u : lo + (b << wsh); };          // first rank
        auto bucket_hi = [&](uint32_t b) -> uint32_t {                                                  // last rank + 1
            return b == static_cast<uint32_t>(NB - 1) ? max(p.n_train, lo + (b << wsh) + 1u) : lo + ((b + 1u) << wsh);
        };

        // ---- pass 1: bucket weights
        each_octet(Select4Hint<1>{}, false, [&](bool ok, const uint32_t (&o)[8]) {
            const uint32_t U = ok ? o[0] : 0u;
#pragma unroll
            for (int e = 1; e < 8; ++e)
                red_add_shared(hist_s + (bucket_of(o[e] & 0xffffffu) << 2), (o[e] >> 24) * U);
        });
        __syncwarp();

        // ---- scan: lane l owns buckets [32 l, 32 l + 32); 128-bit reads rotated by the lane (no bank conflicts)
        uint32_t mine = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 v = ld_shared_v4(hist_s + (lane * 32 + ((j + lane) & 7) * 4) * 4);
            mine += (v.x + v.y) + (v.z + v.w);
        }
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t groups = __ballot_sync(0xffffffffu, mine != 0u);
        const int64_t out_row = p.order ? static_cast<int64_t>(__ldg(p.order + pos)) : pos;
        double* __restrict__ out = p.out + out_row * p.ldo;
        if (total == 0u) {                              // no member with a non-zero weight: cannot happen for a fitted forest
            if (lane < p.Q) out[lane] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        uint32_t tau_l = 0u;
        if (lane < p.Q) {
            const double t = ceil(ql.q[lane] * static_cast<double>(total) / 100.0);
            tau_l = static_cast<uint32_t>(fmin(fmax(t, 1.0), static_cast<double>(total)));
        }
        // last / first non-empty bucket of group G (G has at least one)
        auto edge_of_group = [&](int G, bool want_last) -> uint32_t {
            const uint32_t y = ld_shared_u32(hist_s + (G * 32 + lane) * 4);
            const uint32_t m = __ballot_sync(0xffffffffu, y != 0u);
            return static_cast<uint32_t>(G * 32 + (want_last ? 31 - __clz(m) : __ffs(m) - 1));
        };
        uint32_t k_bq = 0u, k_base = 0u, k_tau = 0u;    // lane q: percentile q
        uint32_t k_from = 0u, k_len = 0u;               // lane q: the ranks of bp, bq and bn -- one interval, the buckets between them are empty
#pragma unroll 1
        for (int q = 0; q < p.Q; ++q) {
            const uint32_t tau = __shfl_sync(0xffffffffu, tau_l, q);
            const int G = __ffs(__ballot_sync(0xffffffffu, incl - mine < tau && tau <= incl)) - 1;
            const uint32_t run = __shfl_sync(0xffffffffu, incl - mine, G);
            const uint32_t x = ld_shared_u32(hist_s + (G * 32 + lane) * 4);
            uint32_t in = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, in, d);
                if (lane >= d) in += u;
            }
            const int ln = __ffs(__ballot_sync(0xffffffffu, run + in - x < tau && tau <= run + in)) - 1;
            const uint32_t base = __shfl_sync(0xffffffffu, run + in - x, ln);
            const uint32_t ne = __ballot_sync(0xffffffffu, x != 0u);
            const uint32_t below = ne & ((1u << ln) - 1u);
            const uint32_t above = ln == 31 ? 0u : (ne >> (ln + 1));
            const uint32_t lower = groups & ((1u << G) - 1u);
            const uint32_t upper = G == 31 ? 0u : (groups >> (G + 1));
            const uint32_t bq = static_cast<uint32_t>(G * 32 + ln);
            uint32_t bp = NONE, bn = NONE;
            if (below) bp = static_cast<uint32_t>(G * 32 + 31 - __clz(below));
            else if (lower) bp = edge_of_group(31 - __clz(lower), true);
            if (above) bn = static_cast<uint32_t>(G * 32 + ln + __ffs(above));
            else if (upper) bn = edge_of_group(G + __ffs(upper), false);
            if (lane == q) {
                k_bq = bq; k_base = base; k_tau = tau;
                k_from = bucket_lo(bp != NONE ? bp : bq);
                k_len = bucket_hi(bn != NONE ? bn : bq) - k_from;
            }
        }
        const uint32_t from0 = __shfl_sync(0xffffffffu, k_from, 0), len0 = __shfl_sync(0xffffffffu, k_len, 0);
        const uint32_t from1 = __shfl_sync(0xffffffffu, k_from, 1), len1 = __shfl_sync(0xffffffffu, k_len, 1);
        const uint32_t from2 = __shfl_sync(0xffffffffu, k_from, 2), len2 = __shfl_sync(0xffffffffu, k_len, 2);   // len 0 beyond Q
        // the histogram goes back to zero
#pragma unroll
        for (int j = 0; j < 8; ++j) st_shared_v4(hist_s + (lane * 32 + ((j + lane) & 7) * 4) * 4, make_uint4(0u, 0u, 0u, 0u));
        __syncwarp();

        // ---- pass 2: members of marked buckets -> hit list (padding slots have count 0: skipped)
        each_octet(Select4Hint<2>{}, true, [&](bool ok, const uint32_t (&o)[8]) {
#pragma unroll
            for (int e = 1; e < 8; ++e) {
                const uint32_t rank = o[e] & 0xffffffu;
                const bool marked = (rank - from0 < len0) | (rank - from1 < len1) | (rank - from2 < len2);
                if (marked && (o[e] >> 24) != 0u) {                // (lanes without an octet hold zeros: count 0)
                    const uint32_t at = atomicAdd(n_hits, 1u);
                    if (at < static_cast<uint32_t>(kSelect4Hits)) {
                        hit_r[at] = rank;
                        hit_w[at] = (o[e] >> 24) * o[0];
                    }
                }
            }
        });
        __syncwarp();
        const uint32_t nh = *n_hits;
        __syncwarp();
        if (lane == 0) *n_hits = 0u;

        // per percentile (lane q): ranks of the crossing entry, its neighbours, and the weights utils.py:72 needs
        uint32_t r_e = 0u, r_s = NONE, r_p = NONE, Ce = 0u, We = 0u, Ws = 0u, Wp = 0u;
        if (nh <= static_cast<uint32_t>(p.hits_cap)) {
            // ---- sort the hits by rank (key = rank << 8 | slot), running sum of the weights in that order
            int half = 64;                              // half the slots of the sorted list
            if (nh <= 128u) {
                select4_sort_hits<4>(hit_r, hit_w, nh, lane);
            } else {
                select4_sort_hits<8>(hit_r, hit_w, nh, lane);
                half = 128;
            }
            if (lane < p.Q) {                           // scalar work of percentile `lane` on the sorted list
                const uint32_t lo_rank = bucket_lo(k_bq);
                int a = 0;                              // first slot with rank >= lo_rank
                for (int sp = half; sp > 0; sp >>= 1)
                    if (a + sp <= static_cast<int>(nh) && hit_r[a + sp - 1] < lo_rank) a += sp;
                const uint32_t cstart = a ? hit_w[a - 1] : 0u;
                const uint32_t need = k_tau - k_base + cstart;             // first slot whose running sum reaches it
                int i = 0;
                for (int sp = half; sp > 0; sp >>= 1)
                    if (i + sp <= static_cast<int>(nh) && hit_w[i + sp - 1] < need) i += sp;
                r_e = hit_r[i];
                int gs = i, ge = i;
                while (gs > 0 && hit_r[gs - 1] == r_e) --gs;
                while (ge + 1 < static_cast<int>(nh) && hit_r[ge + 1] == r_e) ++ge;
                const uint32_t c_before = gs ? hit_w[gs - 1] : 0u;
                We = hit_w[ge] - c_before;
                Ce = k_base + hit_w[ge] - cstart;
                if (gs > 0) {                           // the rank group before: last group of bq below e, or of bp
                    r_p = hit_r[gs - 1];
                    int ps = gs - 1;
                    while (ps > 0 && hit_r[ps - 1] == r_p) --ps;
                    Wp = c_before - (ps ? hit_w[ps - 1] : 0u);
                }
                if (ge + 1 < static_cast<int>(nh)) {
                    r_s = hit_r[ge + 1];
                    int se = ge + 1;
                    while (se + 1 < static_cast<int>(nh) && hit_r[se + 1] == r_s) ++se;
                    Ws = hit_w[se] - hit_w[ge];
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) st_shared_v4(hist_s + (lane * 16 + ((j + lane) & 3) * 4) * 4, make_uint4(0u, 0u, 0u, 0u));   // hit list -> zero
            __syncwarp();
        } else {
            // ---- too many hits for the list: narrow every percentile exactly, the whole warp on one at a time
#pragma unroll
            for (int j = 0; j < 4; ++j) st_shared_v4(hist_s + (lane * 16 + ((j + lane) & 3) * 4) * 4, make_uint4(0u, 0u, 0u, 0u));
            __syncwarp();
#pragma unroll 1
            for (int q = 0; q < p.Q; ++q) {
                const uint32_t tau = __shfl_sync(0xffffffffu, k_tau, q);
                const uint32_t bq = __shfl_sync(0xffffffffu, k_bq, q);
                uint32_t rlo = bucket_lo(bq);                               // ranks [rlo, rlo + 2^wbits) hold the crossing
                uint32_t before = __shfl_sync(0xffffffffu, k_base, q);      // weight of the ranks below rlo
                uint32_t wbits = static_cast<uint32_t>(32 - __clz(bucket_hi(bq) - rlo - 1u));
#pragma unroll 1
                while (wbits > 0u) {
                    const uint32_t sbits = wbits > static_cast<uint32_t>(kSelect4LogNB) ? static_cast<uint32_t>(kSelect4LogNB) : wbits;
                    const uint32_t sshift = wbits - sbits;                  // sub-bucket = (rank - rlo) >> sshift, 2^sbits of them
                    each_octet(Select4Hint<0>{}, false, [&](bool ok, const uint32_t (&o)[8]) {
#pragma unroll
                        for (int e = 1; e < 8; ++e) {
                            const uint32_t d = (o[e] & 0xffffffu) - rlo;
                            if (ok && (d >> wbits) == 0u) red_add_shared(hist_s + ((d >> sshift) << 2), (o[e] >> 24) * o[0]);
                        }
                    });
                    __syncwarp();
                    uint32_t run = before, k_sub = 0u;
                    bool found = false;
                    for (uint32_t t0 = 0; t0 < (1u << sbits); t0 += 32u) {  // tiles of 32 sub-buckets
                        const bool in_range = t0 + lane < (1u << sbits);
                        const uint32_t x = in_range ? ld_shared_u32(hist_s + (t0 + lane) * 4) : 0u;
                        if (in_range) st_shared_u32(hist_s + (t0 + lane) * 4, 0u);
                        uint32_t in = x;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t u = __shfl_up_sync(0xffffffffu, in, d);
                            if (lane >= d) in += u;
                        }
                        const uint32_t hitm = __ballot_sync(0xffffffffu, !found && run + in - x < tau && tau <= run + in);
                        if (hitm && !found) {
                            const int ln = __ffs(hitm) - 1;
                            k_sub = t0 + ln;
                            before = __shfl_sync(0xffffffffu, run + in - x, ln);
                            found = true;
                        }
                        run += __shfl_sync(0xffffffffu, in, 31);
                    }
                    __syncwarp();
                    rlo += k_sub << sshift;
                    wbits = sshift;
                }
                // rlo is the crossing rank: its neighbours, then the merged weights of all three
                uint32_t pr = 0u, sr = NONE, we = 0u;   // pr: rank + 1 of the largest rank below rlo (0: none)
                each_octet(Select4Hint<0>{}, false, [&](bool ok, const uint32_t (&o)[8]) {
#pragma unroll
                    for (int e = 1; e < 8; ++e) {
                        const uint32_t rank = o[e] & 0xffffffu, c = ok ? (o[e] >> 24) : 0u;
                        if (c != 0u && rank < rlo) pr = max(pr, rank + 1u);
                        if (c != 0u && rank > rlo) sr = min(sr, rank);
                        if (rank == rlo) we += c * o[0];
                    }
                });
                pr = __reduce_max_sync(0xffffffffu, pr);
                sr = __reduce_min_sync(0xffffffffu, sr);
                we = __reduce_add_sync(0xffffffffu, we);
                uint32_t wp = 0u, ws = 0u;
                each_octet(Select4Hint<0>{}, false, [&](bool ok, const uint32_t (&o)[8]) {
#pragma unroll
                    for (int e = 1; e < 8; ++e) {
                        const uint32_t rank = o[e] & 0xffffffu, wgt = ok ? (o[e] >> 24) * o[0] : 0u;
                        if (pr != 0u && rank == pr - 1u) wp += wgt;
                        if (rank == sr) ws += wgt;
                    }
                });
                wp = __reduce_add_sync(0xffffffffu, wp);
                ws = __reduce_add_sync(0xffffffffu, ws);
                if (lane == q) {
                    r_e = rlo; r_p = pr ? pr - 1u : NONE; r_s = sr;
                    We = we; Ce = before + we; Wp = wp; Ws = ws;
                }
            }
        }

        // ---- utils.py:69-81 on (pred, e, succ), float32: lane q interpolates percentile q
        if (lane < p.Q) {
            const double qv = ql.q[lane];
            const float s = __fdiv_rn(100.0f, __fmul_rn(static_cast<float>(total), inv));         // utils.py:69,72
            const float a_e = __ldg(p.y_sorted + r_e);
            const float a_s = __ldg(p.y_sorted + (r_s != NONE ? r_s : r_e));
            const float a_p = __ldg(p.y_sorted + (r_p != NONE ? r_p : r_e));
            const float p_e = plotting_position(s, __fmul_rn(static_cast<float>(Ce), inv),
                                                __fmul_rn(static_cast<float>(We), inv));
            float result = a_e;
            if (static_cast<double>(p_e) < qv) {                            // interval (e, succ)
                if (r_s != NONE) {                      // else q lies beyond the last entry: a_last (utils.py:74-75)
                    const float p_s = plotting_position(s, __fmul_rn(static_cast<float>(Ce + Ws), inv),
                                                        __fmul_rn(static_cast<float>(Ws), inv));
                    const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_e), __fsub_rn(p_s, p_e));
                    result = __fadd_rn(a_e, __fmul_rn(frac, __fsub_rn(a_s, a_e)));
                }
            } else if (r_p != NONE) {                   // interval (pred, e); else before the first entry: a_0 (utils.py:76-77)
                const float p_p = plotting_position(s, __fmul_rn(static_cast<float>(Ce - We), inv),
                                                    __fmul_rn(static_cast<float>(Wp), inv));
                const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_p), __fsub_rn(p_e, p_p));
                result = __fadd_rn(a_p, __fmul_rn(frac, __fsub_rn(a_e, a_p)));
            }
            out[lane] = static_cast<double>(result);
        }
        __syncwarp();
    }
}

}  // namespace qf
