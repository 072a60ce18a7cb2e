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
//   roles    a CTA is 8 WORKER warps + 1 SCANNER warp.  Workers only ever touch members; everything
//            that is serial per row (histogram scan, bucket search, selection, interpolation, output)
//            is the scanner's and runs while the workers' copies of the next row are in flight.  They
//            meet at three named barriers per row, each an arrive on one side and a sync on the
//            other (bar.arrive / bar.sync): H "histogram full", M "buckets marked", S "slots full".
//            No thread waits at a barrier for work that is not its own input.
//   copies   every worker thread owns IPT trees.  A warp scan of the octet counts places the lists of
//            the warp's 32 * IPT leaves back to back in the warp's segment of a shared-memory stage
//            and every lane moves its own lists there with 16-byte cp.async (LDGSTS): nothing waits
//            on a global load.  The copies of row r + 1 are issued as soon as the warp is done with
//            row r's second pass.  A row whose lists exceed the segment is read from global memory
//            directly (same code, synchronous loads).
//   pass 1   the warp waits for ITS OWN copies (cp.async.wait_group + __syncwarp: no block barrier)
//            and walks its segment octet by octet, all lanes busy whatever the leaf sizes: seven
//            members -> one predicated shared-memory RED each into the NB-bucket histogram.
//   scan     (scanner) every lane sums its NB / 32 consecutive buckets, one warp scan, then the group
//            of buckets that holds a percentile's crossing is scanned again across the lanes: the
//            crossing bucket bq and its nearest non-empty neighbours bp / bn come out of two ballots.
//            A 16-bit table says which slot arrays a bucket feeds.
//   pass 2   second walk over the staged octets: the seven table lookups of an octet are issued
//            together and tested with one branch; the ~50 members of marked buckets are added at
//            single-rank resolution to the slot arrays [bp | bq | bn] of their percentile(s).
//   select   (scanner) per percentile: crossing entry e in bq's slots, its neighbours inside bq or,
//            failing that, the largest rank of bp / the smallest rank of bn; the three percentiles
//            are then interpolated by three lanes side by side (their y loads overlap).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "compact_lists.cuh"
#include "quantile_kernel.cuh"
#include "select2_kernel.cuh"

namespace qf {

constexpr int kSelect3Workers = 256;                    // worker threads (8 warps)
constexpr int kSelect3Threads = kSelect3Workers + 32;   // + the scanner warp
constexpr int kSelect3Warps = kSelect3Workers / 32;
constexpr int kSelect3CtasPerSm = 3;
constexpr int kSelect3MaxIPT = 4;                       // trees per worker thread: T <= 1024
constexpr int kSelect3RowGroup = 4;                     // rows per 128-bit load of list words

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
    int capw;                              // octets per warp segment of the stage
};

__host__ __device__ inline size_t select3_kernel_smem(int log2nb, int shift1, int capw) {
    const size_t nb = static_cast<size_t>(1) << log2nb;
    return nb * 4 + (nb * 2 + 16) + static_cast<size_t>(kSelect2QB) * 3 * select2_slots(shift1) * 4 +
           static_cast<size_t>(kSelect3Warps) * capw * 32;
}

__device__ __forceinline__ void cp_async_16(uint32_t dst_shared, const void* src_global) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
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
// named barriers: the producer side arrives and goes on, the consumer side waits
template <int ID>
__device__ __forceinline__ void bar_arrive() {
    __threadfence_block();                              // the shared-memory writes / REDs before it are performed
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(kSelect3Threads) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_wait() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(kSelect3Threads) : "memory");
}
enum { kBarHist = 1, kBarMarks = 2, kBarSlots = 3 };

// rows of a CTA: groups of four consecutive positions, groups blockIdx.x, + gridDim.x, ...
struct Select3Rows {
    int64_t g, n_groups, n;
    int l;
    __device__ __forceinline__ int64_t pos() const { return g * kSelect3RowGroup + l; }
    // next row: true = same group; `more` false = this was the CTA's last row
    __device__ __forceinline__ bool next_in_group() const { return l + 1 < kSelect3RowGroup && pos() + 1 < n; }
};

template <int LOG2NB, int IPT>
__global__ void __launch_bounds__(kSelect3Threads, kSelect3CtasPerSm)
quantile_select3_kernel(const Select3Params p, const __grid_constant__ QuantileList ql) {
    constexpr int NB = 1 << LOG2NB;
    constexpr int GB = NB / 32;                         // buckets per scanner lane ("group")
    constexpr int QB = kSelect2QB;
    constexpr uint32_t NONE = 0xffffffffu;
    static_assert(GB == 64 || GB == 128, "a group of buckets is scanned as 2 or 4 buckets per lane");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* h1 = reinterpret_cast<uint32_t*>(smem_raw);          // [NB] bucket weights (zero between rows)
    uint32_t* tbl32 = h1 + NB;                                     // [NB] x 16 bits: slot arrays a bucket feeds
    uint32_t* l2 = tbl32 + NB / 2 + 4;                             // [QB][3][SL] slots: bp | bq | bn (zero between rows)

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int SL = select2_slots(p.shift1);             // slots per bucket (padded to one 128-bit chunk)
    const uint32_t shift = static_cast<uint32_t>(p.shift1);
    const uint32_t h1_s = static_cast<uint32_t>(__cvta_generic_to_shared(h1));
    const uint32_t tbl_s = static_cast<uint32_t>(__cvta_generic_to_shared(tbl32));
    const uint32_t l2_s = static_cast<uint32_t>(__cvta_generic_to_shared(l2));

    // shared state starts at zero and every row leaves it at zero
    for (int i = tid; i < NB + NB / 2 + 4 + QB * 3 * SL; i += kSelect3Threads) h1[i] = 0u;
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
        const uint32_t stage_s = static_cast<uint32_t>(__cvta_generic_to_shared(l2 + QB * 3 * SL)) + warp * capw * 32u;

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

        // lists of one row -> this warp's stage segment.  Returns the octets of the warp's leaves
        // (warp-uniform); more than capw: nothing was copied and the row is read from global memory.
        auto issue_copies = [&](const uint32_t (&w)[IPT]) -> uint32_t {
            uint32_t no[IPT], tot = 0;
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                no[i] = (w[i] == kCompactEmpty) ? 0u : ((w[i] >> 28) + 1u);
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
                uint32_t dst = stage_s + (incl - tot) * 32u;
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const uint4* __restrict__ src = p.lists + 2 * static_cast<size_t>(w[i] & kCompactOffMask);
                    for (uint32_t k = 0; k < no[i]; ++k, dst += 32u) {
                        cp_async_16(dst, src + 2 * k);
                        cp_async_16(dst + 16u, src + 2 * k + 1);
                    }
                }
            }
            cp_async_commit();
            return wtot;
        };

        // f(a, b) on every octet of the warp's leaves: from the stage (all lanes busy) or, for an
        // oversized row, every lane on its own lists straight from global memory
        auto for_each_octet = [&](uint32_t wtot, const uint32_t (&w)[IPT], auto&& f) {
            if (wtot <= capw) {
                for (uint32_t j = lane; j < wtot; j += 32u) {
                    const uint4 a = ld_shared_v4(stage_s + j * 32u);
                    const uint4 b = ld_shared_v4(stage_s + j * 32u + 16u);
                    f(a, b);
                }
            } else {
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    if (w[i] == kCompactEmpty) continue;
                    const uint4* __restrict__ src = p.lists + 2 * static_cast<size_t>(w[i] & kCompactOffMask);
                    const uint32_t no = (w[i] >> 28) + 1u;
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
        uint32_t wtot = issue_copies(w);

        for (;;) {
            // pass 1: bucket weights (octet header a.x = fixed-point weight of one bootstrap count).
            // w == 0 (utils.py:55-57) cannot occur (counts are >= 1); padding slots are predicated off.
            cp_async_wait_all();
            __syncwarp();
            for_each_octet(wtot, w, [&](const uint4& a, const uint4& b) {
                const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 7; ++e)
                    red_add_shared_unless(h1_s + (((m[e] >> shift) & bmask) << 2), (m[e] >> 24) * a.x, m[e] == 0u ? 1u : 0u);
            });
            bar_arrive<kBarHist>();                     // H: this warp's share of the histogram is in
            bar_wait<kBarMarks>();                      // M: the scanner has marked the buckets of the percentiles

            // pass 2: members of marked buckets -> slot arrays at single-rank resolution
            for_each_octet(wtot, w, [&](const uint4& a, const uint4& b) {
                const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                uint32_t code[7], any = 0u;
#pragma unroll
                for (int e = 0; e < 7; ++e) {
                    code[e] = ld_shared_u16(tbl_s + (((m[e] >> shift) & bmask) << 1));
                    any |= code[e];
                }
                if (any) {                              // rare: ~50 members of a row's 3300 (padding: weight 0)
#pragma unroll
                    for (int e = 0; e < 7; ++e) {
                        uint32_t c = code[e];
                        const uint32_t wv = (m[e] >> 24) * a.x;
                        const uint32_t dst = l2_s + (m[e] & slot_mask) * 4u;
                        while (c) {
                            const int s = __ffs(c) - 1;
                            c &= c - 1u;
                            red_add_shared(dst + s * SL * 4, wv);
                        }
                    }
                }
            });
            __syncwarp();                               // every lane of the warp is done with the stage

            // next row: of this group, or the first one of the CTA's next group
            uint32_t w2[IPT];
            bool more = true;
            const bool same = rows.next_in_group();
            if (same) {
#pragma unroll
                for (int i = 0; i < IPT; ++i) w2[i] = pick(cur[i], rows.l + 1);
            } else {
                more = rows.g + gridDim.x < rows.n_groups;
#pragma unroll
                for (int i = 0; i < IPT; ++i) w2[i] = nxt[i].x;
            }
            uint32_t wtot2 = 0u;
            if (more) wtot2 = issue_copies(w2);         // block-uniform
            bar_arrive<kBarSlots>();                    // S: this warp's share of the slots is in
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
#pragma unroll
            for (int i = 0; i < IPT; ++i) w[i] = w2[i];
        }
        return;
    }

    // ============================================= scanner =============================================
    const int nch = SL >> 2;                            // 128-bit chunks per slot array
    const int CH = (nch + 31) >> 5;                     // chunks per lane: 1, 2 or 4 (consecutive)
    const float inv = __fdiv_rn(1.0f, p.fx_scale);
    const uint32_t lt_lane = (1u << lane) - 1u;

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
    // GB buckets of group G spread over the lanes: lane holds buckets G*GB + lane + 32*j, j < GB/32
    auto load_buckets = [&](uint32_t G, uint32_t (&x)[GB / 32]) {
#pragma unroll
        for (int j = 0; j < GB / 32; ++j) x[j] = ld_shared_u32(h1_s + (G * GB + lane + 32 * j) * 4);
    };

    for (;;) {
        const int64_t pos = rows.pos();
        const int64_t out_row = p.order ? static_cast<int64_t>(__ldg(p.order + pos)) : pos;
        double* __restrict__ out = p.out + out_row * p.ldo;

        bar_wait<kBarHist>();                           // H: all eight worker warps have added their members

        // ---- scan: lane sums of GB consecutive buckets, one warp scan
        uint32_t mine = 0u;
#pragma unroll
        for (int k = 0; k < GB / 4; ++k) {
            const uint4 x = ld_shared_v4(h1_s + (lane * GB + 4 * k) * 4);
            mine += x.x + x.y + x.z + x.w;
        }
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t groups = __ballot_sync(0xffffffffu, mine != 0u);       // groups with weight

        // per percentile (kept in lane q after the loop): buckets, running weight before bq, target weight
        uint32_t k_bp = NONE, k_bq = 0u, k_bn = NONE, k_base = 0u, k_tau = 0u;
        uint32_t tau_l = 0u;
        if (total != 0u && lane < p.Q) {
            const double t = ceil(ql.q[lane] * static_cast<double>(total) / 100.0);
            tau_l = static_cast<uint32_t>(fmin(fmax(t, 1.0), static_cast<double>(total)));
        }
        if (total != 0u) {
            for (int q = 0; q < p.Q; ++q) {
                const uint32_t tau = __shfl_sync(0xffffffffu, tau_l, q);
                const uint32_t own = __ballot_sync(0xffffffffu, incl - mine < tau && tau <= incl);
                const int G = __ffs(own) - 1;           // exactly one group holds the crossing
                uint32_t run = __shfl_sync(0xffffffffu, incl - mine, G);      // weight before the group
                uint32_t x[GB / 32];
                load_buckets(G, x);
                // bucket order inside the group: j major, lane minor
                uint32_t bq = 0u, before = 0u;
                uint64_t ne = 0ull;                     // non-empty buckets of the group
#pragma unroll
                for (int j = 0; j < GB / 32; ++j) {
                    uint32_t in = x[j];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xffffffffu, in, d);
                        if (lane >= d) in += u;
                    }
                    const uint32_t hit = __ballot_sync(0xffffffffu, run + in - x[j] < tau && tau <= run + in);
                    if (hit) {                          // warp-uniform
                        const int ln = __ffs(hit) - 1;
                        bq = static_cast<uint32_t>(G * GB + 32 * j + ln);
                        before = __shfl_sync(0xffffffffu, run + in - x[j], ln);
                    }
                    const uint32_t nzj = __ballot_sync(0xffffffffu, x[j] != 0u);
                    if (j < 2) ne |= static_cast<uint64_t>(nzj) << (32 * j);
                    run += __shfl_sync(0xffffffffu, in, 31);
                }
                // nearest non-empty buckets around bq: inside the group (GB == 64: one 64-bit mask; GB == 128:
                // the second half is looked at below through the general path), else in the nearest non-empty group
                uint32_t bp = NONE, bn = NONE;
                const int kq = static_cast<int>(bq) - G * GB;
                if (GB == 64) {
                    const uint64_t below = ne & ((1ull << kq) - 1ull);
                    const uint64_t above = kq == 63 ? 0ull : (ne >> (kq + 1));
                    if (below) bp = static_cast<uint32_t>(G * GB + 63 - __clzll(static_cast<long long>(below)));
                    if (above) bn = static_cast<uint32_t>(G * GB + kq + __ffsll(static_cast<long long>(above)));
                }
                // general search: last non-empty bucket in [lo_b, bq) / first in (bq, hi_b) of one group
                auto extreme_in_group = [&](int Gx, bool want_last, int lo_k, int hi_k) -> uint32_t {   // k range [lo_k, hi_k)
                    uint32_t y[GB / 32];
                    load_buckets(Gx, y);
                    uint32_t best = NONE;
#pragma unroll
                    for (int j = 0; j < GB / 32; ++j) {
                        const int k = 32 * j + lane;
                        const uint32_t m = __ballot_sync(0xffffffffu, y[j] != 0u && k >= lo_k && k < hi_k);
                        if (m) {
                            const int ln = want_last ? (31 - __clz(m)) : (__ffs(m) - 1);
                            const uint32_t cand = static_cast<uint32_t>(Gx * GB + 32 * j + ln);
                            if (want_last) best = cand;                     // later j = larger bucket
                            else if (best == NONE) best = cand;
                        }
                    }
                    return best;
                };
                if (GB != 64) {
                    bp = extreme_in_group(G, true, 0, kq);
                    bn = extreme_in_group(G, false, kq + 1, GB);
                }
                if (bp == NONE) {
                    const uint32_t lower = groups & ((1u << G) - 1u);
                    if (lower) bp = extreme_in_group(31 - __clz(lower), true, 0, GB);
                }
                if (bn == NONE) {
                    const uint32_t upper = G == 31 ? 0u : (groups >> (G + 1));
                    if (upper) bn = extreme_in_group(G + __ffs(upper), false, 0, GB);
                }
                if (lane == q) { k_bp = bp; k_bq = bq; k_bn = bn; k_base = before; k_tau = tau; }
                if (lane == 0) {                        // which slot arrays the buckets feed: 3 * q + {0, 1, 2}
                    const uint32_t a1 = tbl_s + bq * 2u;
                    st_shared_u16(a1, ld_shared_u16(a1) | (2u << (3 * q)));
                    if (bp != NONE) {
                        const uint32_t a0 = tbl_s + bp * 2u;
                        st_shared_u16(a0, ld_shared_u16(a0) | (1u << (3 * q)));
                    }
                    if (bn != NONE) {
                        const uint32_t a2 = tbl_s + bn * 2u;
                        st_shared_u16(a2, ld_shared_u16(a2) | (4u << (3 * q)));
                    }
                }
            }
        }
        // the histogram goes back to zero
#pragma unroll
        for (int k = 0; k < GB / 4; ++k) st_shared_v4(h1_s + (lane * GB + 4 * k) * 4, make_uint4(0u, 0u, 0u, 0u));
        bar_arrive<kBarMarks>();                        // M
        bar_wait<kBarSlots>();                          // S: the slot arrays of this row are complete

        // ---- select: crossing entry and its neighbours, one percentile after the other
        uint32_t k_re = 0u, k_rs = NONE, k_rp = NONE, k_Ce = 0u, k_We = 0u, k_Ws = 0u, k_Wp = 0u;
        if (total != 0u) {
            for (int q = 0; q < p.Q; ++q) {
                const uint32_t bp = __shfl_sync(0xffffffffu, k_bp, q), bq = __shfl_sync(0xffffffffu, k_bq, q);
                const uint32_t bn = __shfl_sync(0xffffffffu, k_bn, q), tau = __shfl_sync(0xffffffffu, k_tau, q);
                const uint32_t base = __shfl_sync(0xffffffffu, k_base, q);
                const uint32_t L_s = l2_s + q * 3 * SL * 4;                    // [bp | bq | bn]
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
                const uint32_t b0 = base + in - s16;
                const uint32_t who = __ballot_sync(0xffffffffu, b0 < tau && tau <= b0 + s16);
                const int lc = __ffs(who) - 1;          // exactly one lane (base < tau <= weight up to bq)
                uint32_t e_bit = 0u, We = 0u, Ce = 0u;
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
                    const uint32_t above = e_bit == 31 ? 0u : (nm_c >> (e_bit + 1));
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
                uint32_t rp = NONE, rs = NONE, Wp = 0u, Ws = 0u;            // ranks and merged weights of pred / succ
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
                if (lane == q) {
                    k_re = (bq << shift) + e; k_rs = rs; k_rp = rp;
                    k_Ce = Ce; k_We = We; k_Ws = Ws; k_Wp = Wp;
                }
            }
            __syncwarp();
            // slots and table back to zero
            for (int ch = lane; ch < p.Q * 3 * nch; ch += 32) st_shared_v4(l2_s + ch * 16, make_uint4(0u, 0u, 0u, 0u));
            if (lane < p.Q) {
                st_shared_u16(tbl_s + k_bq * 2u, 0u);
                if (k_bp != NONE) st_shared_u16(tbl_s + k_bp * 2u, 0u);
                if (k_bn != NONE) st_shared_u16(tbl_s + k_bn * 2u, 0u);
            }
        }

        // ---- utils.py:69-81 on (pred, e, succ), float32; lane q takes percentile q
        if (lane < p.Q) {
            if (total == 0u) {                          // no member with a non-zero weight: cannot happen for a fitted forest
                out[lane] = __longlong_as_double(0x7ff8000000000000LL);
            } else {
                const double qv = ql.q[lane];
                const float s = __fdiv_rn(100.0f, __fmul_rn(static_cast<float>(total), inv));         // utils.py:69,72
                // the three candidate values are requested together (one HBM latency, not two)
                const float a_e = __ldg(p.y_sorted + k_re);
                const float a_s = __ldg(p.y_sorted + (k_rs != NONE ? k_rs : k_re));
                const float a_p = __ldg(p.y_sorted + (k_rp != NONE ? k_rp : k_re));
                const float p_e = plotting_position(s, __fmul_rn(static_cast<float>(k_Ce), inv),
                                                    __fmul_rn(static_cast<float>(k_We), inv));
                float result = a_e;
                if (static_cast<double>(p_e) < qv) {                        // interval (e, succ)
                    if (k_rs != NONE) {                 // else q lies beyond the last entry: a_last (utils.py:74-75)
                        const float p_s = plotting_position(s, __fmul_rn(static_cast<float>(k_Ce + k_Ws), inv),
                                                            __fmul_rn(static_cast<float>(k_Ws), inv));
                        const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_e), __fsub_rn(p_s, p_e));
                        result = __fadd_rn(a_e, __fmul_rn(frac, __fsub_rn(a_s, a_e)));
                    }
                } else if (k_rp != NONE) {              // interval (pred, e); else before the first entry: a_0 (utils.py:76-77)
                    const float p_p = plotting_position(s, __fmul_rn(static_cast<float>(k_Ce - k_We), inv),
                                                        __fmul_rn(static_cast<float>(k_Wp), inv));
                    const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_p), __fsub_rn(p_e, p_p));
                    result = __fadd_rn(a_p, __fmul_rn(frac, __fsub_rn(a_e, a_p)));
                }
                out[lane] = static_cast<double>(result);
            }
        }
        __syncwarp();

        // same row sequence as the workers
        if (rows.next_in_group()) {
            ++rows.l;
        } else {
            rows.g += gridDim.x;
            rows.l = 0;
            if (rows.g >= rows.n_groups) break;
        }
    }
}

}  // namespace qf
