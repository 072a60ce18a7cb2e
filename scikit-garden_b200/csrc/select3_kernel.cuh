// K2, QF_MODE_FAST on compact leaf lists (compact_lists.cuh) -- fused leaf co-membership weighting
// (ensemble.py:137-140) + weighted percentile (utils.py:46-81) by histogram selection.
//
// Same arithmetic contract as select2_kernel.cuh (32-bit fixed-point weights: exact sums whatever
// the order, duplicates of a training sample merge by landing in the same slot; utils.py:72-81 in
// float32 on the three entries around the crossing).  The fixed-point weight of a member is
// count * U with U = rn(scale / s) per leaf (the octet header) instead of rn(f32(count / s) * scale):
// both are within 2^-21 relative of count / s.
//
// ncu on select2 at config 3 (profiles/r2_a_*): 11.9 k warp instructions per row, issue slots 51 %
// busy, 3.5 warps per issue waiting at block barriers and 3.6 on global loads.  This kernel is built
// around those three numbers:
//
//   input    K1 leaves ONE 32-bit word per (tree, row): first octet | (octets - 1) << 28, tree-major,
//            so that a CTA reads the words of FOUR consecutive rows of a tree with one 128-bit load
//            (prefetched one group of rows ahead into registers).
//   roles    a CTA is 8 WORKER warps + one SELECT warp per percentile (3).  Workers own everything
//            that touches members (copies, both passes) and the short parallel scan in between; a
//            select warp owns the serial tail of its percentile (crossing entry, neighbours,
//            interpolation, the y loads, the output) and runs it while the workers are already in
//            the next row.  Workers meet each other at three named barriers per row (bar.sync on 256
//            threads); workers and select warps only ever ARRIVE for each other (bar.arrive on one
//            side, bar.sync on the other): nobody waits for work that is not its own input.
//   copies   every worker thread owns IPT trees.  A warp scan of the octet counts places the lists of
//            the warp's 32 * IPT leaves back to back in the warp's segment of a shared-memory stage
//            and every lane moves its own lists there with 16-byte cp.async (LDGSTS).  Two stages:
//            the copies of row r + 1 are issued BEFORE row r is touched and complete under it
//            (cp.async.wait_group 1), so nothing ever waits on a global load.  A row whose lists
//            exceed the segment is read from global memory directly (same code, synchronous loads).
//   pass 1   the warp waits for ITS OWN copies (wait_group + __syncwarp: no block barrier) and walks
//            its segment octet by octet, all lanes busy whatever the leaf sizes: seven members -> one
//            shared-memory RED each into the NB-bucket histogram (padding slots add zero).
//   scan     4 warps x 16 buckets per thread in registers, one warp scan + 4 partials; the thread
//            that finds a percentile's crossing bucket bq among its own looks for the nearest
//            non-empty neighbours bp / bn in its registers or, through a 128-bit map of non-empty
//            16-bucket groups, in one other group of the (not yet cleared) histogram.  A 16-bit table
//            says which slot arrays a bucket feeds.
//   pass 2   second walk over the staged octets: the seven table lookups of an octet are issued
//            together; the ~50 members of marked buckets are added at single-rank resolution to the
//            slot arrays [bp | bq | bn] of their percentile(s).
//   select   one warp per percentile: crossing entry e in bq's slots, its neighbours inside bq or,
//            failing that, the largest rank of bp / the smallest rank of bn; lane 0 interpolates.
//   Histogram and table are double-buffered by row parity, so that clearing them never races with
//   the next row's writes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "compact_lists.cuh"
#include "quantile_kernel.cuh"
#include "select2_kernel.cuh"

namespace qf {

constexpr int kSelect3Workers = 256;                    // worker threads (8 warps)
constexpr int kSelect3Warps = kSelect3Workers / 32;
// Two shapes of the CTA (template parameters SELW, STAGES of the kernel):
//   SELW = 3, STAGES = 2: one select warp per percentile, two stages (the copies of row r + 1 complete under
//                          row r), 100 KB of shared memory at config 3 -> 2 CTAs (22 warps) per SM
//   SELW = 1, STAGES = 1: one select warp for all percentiles, one stage (the copies of row r + 1 are issued
//                          when the warp is done with row r and complete under the other CTAs of the SM),
//                          71 KB -> 3 CTAs (27 warps) per SM
__host__ __device__ constexpr int select3_threads(int selw) { return kSelect3Workers + 32 * selw; }
__host__ __device__ constexpr int select3_ctas_per_sm(int stages) { return stages == 2 ? 2 : 3; }
constexpr int kSelect3MaxIPT = 4;                       // trees per worker thread: T <= 1024
constexpr int kSelect3RowGroup = 4;                     // rows per 128-bit load of list words
constexpr int kSelect3ScanThreads = 128;                // threads that scan the histogram

struct Select3Params {
    const uint32_t* __restrict__ seg;      // [T][ld] compact list words from K1 (APPLY_EMIT_COMPACT)
    int64_t ld;                            // multiple of 4
    const int* __restrict__ order;         // [n] locality position -> output row; nullptr = identity
    const uint4* __restrict__ lists;       // two uint4 per octet
    const float* __restrict__ y_sorted;    // [N]
    double* __restrict__ out;              // [n][ldo], indexed by output row
    int64_t n;
    int64_t ldo;
    int T;
    int Q;                                 // <= kSelect2QB
    float fx_scale;                        // fixed-point units per unit of weight (the octet headers were built with it)
    int shift1;                            // bucket = rank >> shift1
    int capw;                              // octets per warp segment of one stage
};

// shared memory: 2 histograms | 2 tables | slots | 2 stages
__host__ __device__ inline size_t select3_kernel_smem(int log2nb, int shift1, int capw, int stages) {
    const size_t nb = static_cast<size_t>(1) << log2nb;
    return 2 * nb * 4 + 2 * (nb * 2) + static_cast<size_t>(kSelect2QB) * 3 * select2_slots(shift1) * 4 +
           static_cast<size_t>(stages) * kSelect3Warps * capw * 32;
}

__device__ __forceinline__ void cp_async_16(uint32_t dst_shared, const void* src_global) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory");
}
// named barriers (PTX bar): the producer side arrives and goes on, the consumer side waits
template <int ID, int COUNT>
__device__ __forceinline__ void bar_arrive() {
    __threadfence_block();                              // the shared-memory writes / REDs before it are performed
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int ID, int COUNT>
__device__ __forceinline__ void bar_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
enum { kBarWorkers = 1, kBarMarks = 2, kBarSlots = 3 };

// what the scan hands to the select warp of a percentile (double-buffered by row parity)
struct Select3Handoff {
    uint32_t tau, bp, bq, bn, base, total, pad0, pad1;
};

// rows of a CTA: groups of four consecutive positions, groups blockIdx.x, + gridDim.x, ...
struct Select3Rows {
    int64_t g, n_groups, n;
    int l;
    __device__ __forceinline__ int64_t pos() const { return g * kSelect3RowGroup + l; }
    __device__ __forceinline__ bool next_in_group() const { return l + 1 < kSelect3RowGroup && pos() + 1 < n; }
    __device__ __forceinline__ bool has_next(unsigned stride) const { return next_in_group() || g + stride < n_groups; }
    __device__ __forceinline__ void advance(unsigned stride) {
        if (next_in_group()) { ++l; } else { g += stride; l = 0; }
    }
};

template <int LOG2NB, int IPT, int SELW, int STAGES>
__global__ void __launch_bounds__(select3_threads(SELW), select3_ctas_per_sm(STAGES))
quantile_select3_kernel(const Select3Params p, const __grid_constant__ QuantileList ql) {
    constexpr int kSelect3Threads = select3_threads(SELW);
    constexpr int NB = 1 << LOG2NB;
    constexpr int BPT = NB / kSelect3ScanThreads;       // buckets per scan thread: 16 or 32
    constexpr int QB = kSelect2QB;
    constexpr int SW = kSelect3ScanThreads / 32;        // scan warps
    constexpr uint32_t NONE = 0xffffffffu;
    static_assert(BPT == 16 || BPT == 32, "a scan thread keeps its buckets' non-empty flags in one 32-bit mask");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* h1 = reinterpret_cast<uint32_t*>(smem_raw);          // [2][NB] bucket weights (zero between uses)
    uint32_t* tbl32 = h1 + 2 * NB;                                 // [2][NB] x 16 bits: slot arrays a bucket feeds
    uint32_t* l2 = tbl32 + NB;                                     // [QB][3][SL] slots: bp | bq | bn (zero between rows)
    __shared__ uint32_t s_red[SW];                      // weight of every scan warp's buckets
    __shared__ uint32_t s_ne[SW];                       // bit l: scan thread l of the warp has a non-empty bucket
    __shared__ Select3Handoff s_hand[2][QB];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int SL = select2_slots(p.shift1);             // slots per bucket (padded to one 128-bit chunk)
    const uint32_t shift = static_cast<uint32_t>(p.shift1);
    const uint32_t h1_s = static_cast<uint32_t>(__cvta_generic_to_shared(h1));
    const uint32_t tbl_s = static_cast<uint32_t>(__cvta_generic_to_shared(tbl32));
    const uint32_t l2_s = static_cast<uint32_t>(__cvta_generic_to_shared(l2));

    // shared state starts at zero and every row leaves it at zero
    for (int i = tid; i < 2 * NB + NB + QB * 3 * SL; i += kSelect3Threads) h1[i] = 0u;
    __syncthreads();

    Select3Rows rows;
    rows.g = blockIdx.x;
    rows.n_groups = (p.n + kSelect3RowGroup - 1) / kSelect3RowGroup;
    rows.n = p.n;
    rows.l = 0;
    if (rows.g >= rows.n_groups) return;

    if (warp < kSelect3Warps) {
        // =========================================== workers ===========================================
        const uint32_t capw = static_cast<uint32_t>(p.capw);
        const uint32_t slot_mask = (1u << p.shift1) - 1u;
        const uint32_t bmask = (1u << (24 - p.shift1)) - 1u;       // (m >> shift) & bmask = bucket of a member word
        const uint32_t stage_bytes = kSelect3Warps * capw * 32u;   // one stage
        const uint32_t stage0_s = static_cast<uint32_t>(__cvta_generic_to_shared(l2 + QB * 3 * SL)) + warp * capw * 32u;

        auto load_group = [&](int64_t g, uint4 (&buf)[IPT]) {      // list words of this thread's trees, 4 rows
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const int t = tid + i * kSelect3Workers;
                buf[i] = make_uint4(kCompactEmpty, kCompactEmpty, kCompactEmpty, kCompactEmpty);
                if (g < rows.n_groups && t < p.T)
                    buf[i] = __ldg(reinterpret_cast<const uint4*>(p.seg + static_cast<int64_t>(t) * p.ld + g * kSelect3RowGroup));
            }
        };
        auto pick = [](const uint4& v, int l) -> uint32_t { return l == 0 ? v.x : (l == 1 ? v.y : (l == 2 ? v.z : v.w)); };

        // lists of one row -> this warp's segment of stage `st`.  Returns the octets of the warp's leaves
        // (warp-uniform); more than capw: nothing was copied and the row is read from global memory.
        auto issue_copies = [&](const uint32_t (&w)[IPT], uint32_t st) -> uint32_t {
            uint32_t no[IPT], tot = 0;
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                no[i] = (w[i] == kCompactEmpty) ? 0u : ((w[i] >> 28) + 1u);
                if (no[i] > static_cast<uint32_t>(kCompactMaxOctets)) no[i] = 1u << 20;   // long list: never staged
                tot += no[i];
            }
            uint32_t incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            const uint32_t wtot = __shfl_sync(0xffffffffu, incl, 31);
            if (wtot <= capw) {
                const uint32_t seg_s = stage0_s + st * stage_bytes;
                uint32_t rel = (incl - tot) * 32u;                         // byte offset of this lane's first octet
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const uint4* __restrict__ src = p.lists + 2 * static_cast<size_t>(w[i] & kCompactOffMask);
                    // octet j of the segment keeps its halves swapped when bit 2 of j is set: the 128-bit reads of
                    // eight consecutive lanes then fall into eight different 16-byte bank groups
                    for (uint32_t k = 0; k < no[i]; ++k, rel += 32u) {
                        const uint32_t sw = (rel >> 3) & 16u;              // bit 2 of the octet index
                        cp_async_16(seg_s + rel + sw, src + 2 * k);
                        cp_async_16(seg_s + rel + (sw ^ 16u), src + 2 * k + 1);
                    }
                }
            }
            return wtot;
        };

        // f(a, b) on every octet of the warp's leaves: from the stage (all lanes busy) or, for an
        // oversized row, every lane on its own lists straight from global memory
        auto for_each_octet = [&](uint32_t wtot, const uint32_t (&w)[IPT], uint32_t st, auto&& f) {
            if (wtot <= capw) {
                const uint32_t base = stage0_s + st * stage_bytes;
                const uint32_t sw = (static_cast<uint32_t>(lane) << 2) & 16u;   // bit 2 of j = bit 2 of the lane
                for (uint32_t j = lane; j < wtot; j += 32u) {
                    const uint4 a = ld_shared_v4(base + j * 32u + sw);
                    const uint4 b = ld_shared_v4(base + j * 32u + (sw ^ 16u));
                    f(a, b);
                }
            } else {
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    if (w[i] == kCompactEmpty) continue;
                    const uint4* __restrict__ src = p.lists + 2 * static_cast<size_t>(w[i] & kCompactOffMask);
                    uint32_t no = (w[i] >> 28) + 1u;
                    if (no > static_cast<uint32_t>(kCompactMaxOctets)) {   // long list: info octet first
                        no = __ldg(src).x;
                        src += 2;
                    }
                    for (uint32_t k = 0; k < no; ++k) f(__ldg(src + 2 * k), __ldg(src + 2 * k + 1));
                }
            }
        };

        uint4 cur[IPT], nxt[IPT];
        load_group(rows.g, cur);
        load_group(rows.g + gridDim.x, nxt);
        uint32_t w[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) w[i] = cur[i].x;
        uint32_t par = 0u;                              // row parity: stage / histogram / table / hand-off in use
        uint32_t wtot = issue_copies(w, 0u);
        cp_async_commit();

        for (;;) {
            uint32_t w2[IPT];
            const bool more = rows.has_next(gridDim.x);
            const bool same = rows.next_in_group();
#pragma unroll
            for (int i = 0; i < IPT; ++i) w2[i] = same ? pick(cur[i], rows.l + 1) : nxt[i].x;
            uint32_t wtot2 = 0u;
            const uint32_t st = STAGES == 2 ? par : 0u; // stage of this row
            if (STAGES == 2) {
                // the copies of the NEXT row go out first (its stage was released by this warp's pass 2 of the
                // row before this one); one commit group per row, empty when there is no next row
                if (more) wtot2 = issue_copies(w2, par ^ 1u);      // block-uniform
                cp_async_commit();
                cp_async_wait<1>();                     // everything but the group just committed: this row is in
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();

            // pass 1: bucket weights (octet header a.x = fixed-point weight of one bootstrap count).
            // w == 0 (utils.py:55-57) cannot occur for a member (counts are >= 1); padding slots carry count 0
            // and a pseudo-random rank (compact_lists.cuh): they add zero to some bucket, no predicate, no branch.
            const uint32_t h_s = h1_s + par * (NB * 4);
            for_each_octet(wtot, w, st, [&](const uint4& a, const uint4& b) {
                const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 7; ++e)
                    red_add_shared(h_s + (((m[e] >> shift) & bmask) << 2), (m[e] >> 24) * a.x);
            });
            bar_sync<kBarWorkers, kSelect3Workers>();   // B1: histogram complete

            // scan: BPT consecutive buckets per thread of the first SW warps, inclusive prefix in registers
            uint32_t c[BPT], nz = 0u, mine = 0u, incl = 0u;
            if (warp < SW) {
#pragma unroll
                for (int k = 0; k < BPT / 4; ++k) {
                    const uint4 x = ld_shared_v4(h_s + (tid * BPT + 4 * k) * 4);
                    c[4 * k] = x.x; c[4 * k + 1] = x.y; c[4 * k + 2] = x.z; c[4 * k + 3] = x.w;
                }
#pragma unroll
                for (int k = 0; k < BPT; ++k) {
                    mine += c[k];
                    nz |= (c[k] != 0u ? 1u : 0u) << k;
                }
                incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += u;
                }
                const uint32_t ne = __ballot_sync(0xffffffffu, mine != 0u);
                if (lane == 31) s_red[warp] = incl;
                if (lane == 0) s_ne[warp] = ne;
            }
            bar_sync<kBarWorkers, kSelect3Workers>();   // B2: partials visible
            if (warp < SW) {
                uint32_t run0 = incl - mine, total = 0u;            // running weight before this thread's buckets
#pragma unroll
                for (int v = 0; v < SW; ++v) {
                    const uint32_t x = s_red[v];
                    run0 += (v < warp) ? x : 0u;
                    total += x;
                }
                if (total != 0u) {
                    // target running weight of every percentile (lane q of every scan warp, then broadcast)
                    uint32_t tau_l = 0u;
                    if (lane < p.Q) {
                        const double t = ceil(ql.q[lane] * static_cast<double>(total) / 100.0);
                        tau_l = static_cast<uint32_t>(fmin(fmax(t, 1.0), static_cast<double>(total)));
                    }
                    {
                        uint32_t run = run0;
#pragma unroll
                        for (int k = 0; k < BPT; ++k) { run += c[k]; c[k] = run; }
                    }
                    for (int q = 0; q < p.Q; ++q) {
                        const uint32_t tau = __shfl_sync(0xffffffffu, tau_l, q);
                        if (run0 < tau && tau <= c[BPT - 1]) {      // exactly one thread: bq is one of its own buckets
                            int k = 0;
                            uint32_t before = run0;
#pragma unroll
                            for (int j = 0; j < BPT - 1; ++j)
                                if (c[j] < tau) { k = j + 1; before = c[j]; }
                            const uint32_t bq = static_cast<uint32_t>(tid * BPT + k);
                            // nearest non-empty buckets: among this thread's own, else in the nearest non-empty
                            // group of BPT buckets (map s_ne), read from the histogram itself
                            const uint32_t below = nz & ((1u << k) - 1u);
                            const uint32_t above = k == 31 ? 0u : (nz >> (k + 1));
                            uint32_t bp = NONE, bn = NONE;
                            if (below) {
                                bp = static_cast<uint32_t>(tid * BPT + 31 - __clz(below));
                            } else {
                                int gsel = -1;           // last non-empty group before this thread's
                                for (int v = SW - 1; v >= 0; --v) {
                                    uint32_t msk = s_ne[v];
                                    if (v > warp) msk = 0u;
                                    else if (v == warp) msk &= (1u << lane) - 1u;
                                    if (gsel < 0 && msk) gsel = v * 32 + 31 - __clz(msk);
                                }
                                if (gsel >= 0) {
                                    int best = 0;
                                    for (int j = 0; j < BPT; ++j)
                                        if (ld_shared_u32(h_s + (gsel * BPT + j) * 4)) best = j;
                                    bp = static_cast<uint32_t>(gsel * BPT + best);
                                }
                            }
                            if (above) {
                                bn = static_cast<uint32_t>(tid * BPT + k + __ffs(above));
                            } else {
                                int gsel = -1;           // first non-empty group after this thread's
                                for (int v = 0; v < SW; ++v) {
                                    uint32_t msk = s_ne[v];
                                    if (v < warp) msk = 0u;
                                    else if (v == warp) msk = lane == 31 ? 0u : (msk & ~((2u << lane) - 1u));
                                    if (gsel < 0 && msk) gsel = v * 32 + __ffs(msk) - 1;
                                }
                                if (gsel >= 0) {
                                    int best = BPT - 1;
                                    for (int j = BPT - 1; j >= 0; --j)
                                        if (ld_shared_u32(h_s + (gsel * BPT + j) * 4)) best = j;
                                    bn = static_cast<uint32_t>(gsel * BPT + best);
                                }
                            }
                            Select3Handoff hd;
                            hd.tau = tau; hd.bp = bp; hd.bq = bq; hd.bn = bn; hd.base = before; hd.total = total;
                            hd.pad0 = hd.pad1 = 0u;
                            s_hand[par][q] = hd;
                            // which slot arrays the buckets feed: 3 * q + {0, 1, 2}
                            uint32_t* tb = tbl32 + par * (NB / 2);
                            atomicOr(&tb[bq >> 1], (2u << (3 * q)) << (16 * (bq & 1)));
                            if (bp != NONE) atomicOr(&tb[bp >> 1], (1u << (3 * q)) << (16 * (bp & 1)));
                            if (bn != NONE) atomicOr(&tb[bn >> 1], (4u << (3 * q)) << (16 * (bn & 1)));
                        }
                    }
                } else if (tid < p.Q) {                 // no member with a non-zero weight: cannot happen for a fitted forest
                    s_hand[par][tid].total = 0u;
                }
            }
            // B3: marks and hand-off visible; the select warps have released the slots of the previous row
            bar_sync<kBarMarks, kSelect3Threads>();
            if (warp < SW) {                            // this histogram is used again two rows from now
#pragma unroll
                for (int k = 0; k < BPT / 4; ++k) st_shared_v4(h_s + (tid * BPT + 4 * k) * 4, make_uint4(0u, 0u, 0u, 0u));
            }

            // pass 2: members of marked buckets -> slot arrays at single-rank resolution
            const uint32_t t_s = tbl_s + par * (NB * 2);
            for_each_octet(wtot, w, st, [&](const uint4& a, const uint4& b) {
                const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                uint32_t code[7];
#pragma unroll
                for (int e = 0; e < 7; ++e) code[e] = ld_shared_u16(t_s + (((m[e] >> shift) & bmask) << 1));   // issued together
#pragma unroll
                for (int e = 0; e < 7; ++e) {
                    uint32_t cd = code[e];
                    if (cd) {                           // ~50 members of a row's 3300 (a padding slot adds zero)
                        const uint32_t wv = (m[e] >> 24) * a.x;
                        const uint32_t dst = l2_s + (m[e] & slot_mask) * 4u;
                        do {
                            const int s = __ffs(cd) - 1;
                            cd &= cd - 1u;
                            red_add_shared(dst + s * SL * 4, wv);
                        } while (cd);
                    }
                }
            });
            __syncwarp();                               // every lane of the warp is done with this stage
            if (STAGES == 1) {                          // one stage: the next row's copies fly from here on
                if (more) wtot2 = issue_copies(w2, 0u);
                cp_async_commit();
            }
            bar_arrive<kBarSlots, kSelect3Threads>();   // S: this warp's share of the slots is in
            if (!more) break;
            if (same) {
                ++rows.l;
            } else {
                rows.g += gridDim.x;
                rows.l = 0;
#pragma unroll
                for (int i = 0; i < IPT; ++i) cur[i] = nxt[i];
                load_group(rows.g + gridDim.x, nxt);
            }
            wtot = wtot2;
            par ^= 1u;
#pragma unroll
            for (int i = 0; i < IPT; ++i) w[i] = w2[i];
        }
        cp_async_wait<0>();
        return;
    }

    // ========================================= select warps =========================================
    // select warp v takes the percentiles v, v + SELW, ...; lane k keeps what the interpolation of the
    // warp's k-th percentile needs
    const int sw_id = warp - kSelect3Warps;
    const int nch = SL >> 2;                            // 128-bit chunks per slot array
    const int CH = (nch + 31) >> 5;                     // chunks per lane: 1, 2 or 4 (consecutive)
    const float inv = __fdiv_rn(1.0f, p.fx_scale);

    // lane L's share of a slot array: chunks [L * CH, L * CH + CH) -> v[], bit i of the result = slot i of the share non-empty
    auto load_share = [&](uint32_t arr_s, uint4 (&v)[4]) -> uint32_t {
        uint32_t nm = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = make_uint4(0u, 0u, 0u, 0u);
            const int ch = lane * CH + j;
            if (j < CH && ch < nch) v[j] = ld_shared_v4(arr_s + ch * 16);
            nm |= ((v[j].x ? 1u : 0u) | (v[j].y ? 2u : 0u) | (v[j].z ? 4u : 0u) | (v[j].w ? 8u : 0u)) << (4 * j);
        }
        return nm;
    };
    auto slot_of = [&](int ln, int bit) -> uint32_t { return static_cast<uint32_t>(ln * CH * 4 + bit); };

    uint32_t par = 0u;
    bar_arrive<kBarMarks, kSelect3Threads>();           // the slots are free for the first row
    for (;;) {
        const int64_t pos = rows.pos();
        const bool more = rows.has_next(gridDim.x);
        int64_t out_row = pos;
        if (p.order) out_row = static_cast<int64_t>(__ldg(p.order + pos));
        bar_sync<kBarSlots, kSelect3Threads>();         // S: the slot arrays of this row are complete
        // kept by lane k for the warp's k-th percentile
        bool k_have = false;
        uint32_t k_re = 0u, k_rs = NONE, k_rp = NONE, k_Ce = 0u, k_We = 0u, k_Ws = 0u, k_Wp = 0u, k_total = 0u;
        for (int q = sw_id, kq = 0; q < p.Q; q += SELW, ++kq) {
            const uint32_t L_s = l2_s + q * 3 * SL * 4; // [bp | bq | bn]
            uint32_t re = 0u, rs = NONE, rp = NONE, Ce = 0u, We = 0u, Ws = 0u, Wp = 0u;
            const Select3Handoff hd = s_hand[par][q];
            const uint32_t total = hd.total;
            bool have = false;
            if (total != 0u) {
                const uint32_t tau = hd.tau, bp = hd.bp, bq = hd.bq, bn = hd.bn;
                uint4 v[4];
                const uint32_t nm = load_share(L_s + SL * 4, v);
                uint32_t s16 = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) s16 += v[j].x + v[j].y + v[j].z + v[j].w;
                uint32_t in = s16;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, in, d);
                    if (lane >= d) in += u;
                }
                const uint32_t b0 = hd.base + in - s16;
                const uint32_t who = __ballot_sync(0xffffffffu, b0 < tau && tau <= b0 + s16);
                const int lc = __ffs(who) - 1;          // exactly one lane (base < tau <= weight up to bq)
                uint32_t e_bit = 0u;
                if (lane == lc) {
                    const uint32_t vv[16] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                                             v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
                    uint32_t run = b0;
                    bool got = false;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        run += vv[i];
                        if (!got && run >= tau) { got = true; e_bit = i; We = vv[i]; Ce = run; }
                    }
                }
                e_bit = __shfl_sync(0xffffffffu, e_bit, lc);
                We = __shfl_sync(0xffffffffu, We, lc);
                Ce = __shfl_sync(0xffffffffu, Ce, lc);
                const uint32_t e = slot_of(lc, e_bit);
                // neighbours of e inside bq
                const uint32_t nm_c = __shfl_sync(0xffffffffu, nm, lc);
                const uint32_t lanes_ne = __ballot_sync(0xffffffffu, nm != 0u);
                uint32_t pin = NONE, sin = NONE;        // slots
                {
                    const uint32_t below = nm_c & ((1u << e_bit) - 1u);
                    const uint32_t lower = lanes_ne & ((1u << lc) - 1u);
                    if (below) pin = slot_of(lc, 31 - __clz(below));
                    else if (lower) {
                        const int lp = 31 - __clz(lower);
                        pin = slot_of(lp, 31 - __clz(__shfl_sync(0xffffffffu, nm, lp)));
                    }
                    const uint32_t above = nm_c >> (e_bit + 1);
                    const uint32_t upper = lc == 31 ? 0u : (lanes_ne >> (lc + 1));
                    if (above) sin = slot_of(lc, e_bit + __ffs(above));
                    else if (upper) {
                        const int ls = lc + __ffs(upper);
                        sin = slot_of(ls, __ffs(__shfl_sync(0xffffffffu, nm, ls)) - 1);
                    }
                }
                // last non-empty slot of array `which` (0: bp) / first one (2: bn), and its weight
                auto extreme_slot = [&](int which, bool want_last, uint32_t& slot, uint32_t& wgt) {
                    uint4 y[4];
                    const uint32_t ym = load_share(L_s + which * SL * 4, y);
                    const uint32_t lm = __ballot_sync(0xffffffffu, ym != 0u);
                    slot = NONE;
                    wgt = 0u;
                    if (lm) {
                        const int ln = want_last ? (31 - __clz(lm)) : (__ffs(lm) - 1);
                        const uint32_t mk = __shfl_sync(0xffffffffu, ym, ln);
                        const int bit = want_last ? (31 - __clz(mk)) : (__ffs(mk) - 1);
                        slot = slot_of(ln, bit);
                        wgt = ld_shared_u32(L_s + (which * SL + slot) * 4);
                    }
                };
                if (pin != NONE) {
                    rp = (bq << shift) + pin;
                    Wp = ld_shared_u32(L_s + (SL + pin) * 4);
                } else if (bp != NONE) {                // largest rank of bp
                    uint32_t sl;
                    extreme_slot(0, true, sl, Wp);
                    if (sl != NONE) rp = (bp << shift) + sl;
                }
                if (sin != NONE) {
                    rs = (bq << shift) + sin;
                    Ws = ld_shared_u32(L_s + (SL + sin) * 4);
                } else if (bn != NONE) {                // smallest rank of bn
                    uint32_t sl;
                    extreme_slot(2, false, sl, Ws);
                    if (sl != NONE) rs = (bn << shift) + sl;
                }
                re = (bq << shift) + e;
                have = true;
                __syncwarp();
                // slots and table entries of this percentile back to zero
                for (int ch = lane; ch < 3 * nch; ch += 32) st_shared_v4(L_s + ch * 16, make_uint4(0u, 0u, 0u, 0u));
                if (lane == 0) {
                    const uint32_t t_s = tbl_s + par * (NB * 2);
                    st_shared_u16(t_s + bq * 2u, 0u);
                    if (bp != NONE) st_shared_u16(t_s + bp * 2u, 0u);
                    if (bn != NONE) st_shared_u16(t_s + bn * 2u, 0u);
                }
            }
            if (lane == kq) {
                k_have = have; k_total = total;
                k_re = re; k_rs = rs; k_rp = rp; k_Ce = Ce; k_We = We; k_Ws = Ws; k_Wp = Wp;
            }
        }
        // the slots are free again: the workers may start pass 2 of the next row
        if (more) bar_arrive<kBarMarks, kSelect3Threads>();

        // ---- utils.py:69-81 on (pred, e, succ), float32: lane k interpolates the warp's k-th percentile
        // (the y loads of the warp's percentiles overlap)
        {
            const int q = sw_id + lane * SELW;
            const bool have = k_have;
            const uint32_t re = k_re, rs = k_rs, rp = k_rp, Ce = k_Ce, We = k_We, Ws = k_Ws, Wp = k_Wp, total = k_total;
            float result = 0.0f;
            double* __restrict__ out = p.out + out_row * p.ldo;
            if (q >= p.Q) {
            } else if (!have) {
                out[q] = __longlong_as_double(0x7ff8000000000000LL);
            } else {
                const double qv = ql.q[q];
                const float s = __fdiv_rn(100.0f, __fmul_rn(static_cast<float>(total), inv));         // utils.py:69,72
                // the three candidate values are requested together (one HBM latency, not two)
                const float a_e = __ldg(p.y_sorted + re);
                const float a_s = __ldg(p.y_sorted + (rs != NONE ? rs : re));
                const float a_p = __ldg(p.y_sorted + (rp != NONE ? rp : re));
                const float p_e = plotting_position(s, __fmul_rn(static_cast<float>(Ce), inv),
                                                    __fmul_rn(static_cast<float>(We), inv));
                result = a_e;
                if (static_cast<double>(p_e) < qv) {                        // interval (e, succ)
                    if (rs != NONE) {                   // else q lies beyond the last entry: a_last (utils.py:74-75)
                        const float p_s = plotting_position(s, __fmul_rn(static_cast<float>(Ce + Ws), inv),
                                                            __fmul_rn(static_cast<float>(Ws), inv));
                        const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_e), __fsub_rn(p_s, p_e));
                        result = __fadd_rn(a_e, __fmul_rn(frac, __fsub_rn(a_s, a_e)));
                    }
                } else if (rp != NONE) {                // interval (pred, e); else before the first entry: a_0 (utils.py:76-77)
                    const float p_p = plotting_position(s, __fmul_rn(static_cast<float>(Ce - We), inv),
                                                        __fmul_rn(static_cast<float>(Wp), inv));
                    const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_p), __fsub_rn(p_e, p_p));
                    result = __fadd_rn(a_p, __fmul_rn(frac, __fsub_rn(a_e, a_p)));
                }
                out[q] = static_cast<double>(result);
            }
        }
        __syncwarp();
        if (!more) break;
        rows.advance(gridDim.x);
        par ^= 1u;
    }
}

}  // namespace qf
