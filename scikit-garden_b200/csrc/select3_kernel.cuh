// K2, QF_MODE_FAST on compact leaf lists (compact_lists.cuh) -- fused leaf co-membership weighting
// (ensemble.py:137-140) + weighted percentile (utils.py:46-81) by histogram selection.
//
// Same arithmetic contract as select2_kernel.cuh (32-bit fixed-point weights: exact sums whatever
// the order, duplicates of a training sample merge by landing in the same slot; utils.py:72-81 in
// float32 on the three entries around the crossing).  The fixed-point weight of a member is
// count * U with U = rn(scale / s) per leaf instead of rn(f32(count / s) * scale): both are within
// 2^-21 relative of count / s.
//
// What changes against select2 is how the bytes arrive and how many instructions touch them:
//
//   input    K1 leaves ONE 32-bit word per (tree, row): first octet | (octets - 1) << 28, tree-major,
//            so that a CTA reads the words of FOUR consecutive rows of a tree with one 128-bit load
//            (prefetched one group of rows ahead into registers).
//   copies   every thread owns IPT trees.  A warp scan of the octet counts places the lists of the
//            warp's 32 * IPT leaves back to back in the warp's segment of a shared-memory stage and
//            every lane moves its own lists there with 16-byte cp.async (LDGSTS): nothing waits on
//            a global load.  The copies of row r + 1 are issued as soon as the warp is done with row
//            r's second pass, i.e. they fly during selection / output of row r and wait / pass 1 of
//            the other CTAs on the SM.  A row whose lists exceed the segment is read from global
//            memory directly (same code, synchronous loads).
//   pass 1   the warp waits for ITS OWN copies (cp.async.wait_group + __syncwarp: no block barrier)
//            and walks its segment octet by octet, all lanes busy whatever the leaf sizes: header ->
//            U, seven members -> one predicated shared-memory RED each into the NB-bucket histogram.
//   scan     block scan of the bucket weights in registers, fused with a prefix-max / suffix-min of
//            the non-empty buckets, so that the thread that finds the crossing bucket bq of a
//            percentile also knows its non-empty neighbours bp / bn: two barriers instead of five.
//   pass 2   second walk over the staged octets: a 16-bit table says which slot arrays a bucket
//            feeds; the ~50 members of marked buckets are added at single-rank resolution.
//   select   one warp per percentile, as in select2.
//
// Four block barriers per row; 3 CTAs of 256 threads per SM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "compact_lists.cuh"
#include "quantile_kernel.cuh"
#include "select2_kernel.cuh"

namespace qf {

constexpr int kSelect3Threads = 256;
constexpr int kSelect3Warps = kSelect3Threads / 32;
constexpr int kSelect3CtasPerSm = 3;
constexpr int kSelect3MaxIPT = 4;                       // trees per thread: T <= 1024
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
    float fx_scale;                        // fixed-point units per unit of weight
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

template <int LOG2NB, int IPT>
__global__ void __launch_bounds__(kSelect3Threads, kSelect3CtasPerSm)
quantile_select3_kernel(const Select3Params p, const __grid_constant__ QuantileList ql) {
    constexpr int THREADS = kSelect3Threads;
    constexpr int NB = 1 << LOG2NB;
    constexpr int BPT = NB / THREADS;                   // buckets scanned per thread
    constexpr int NW = kSelect3Warps;
    constexpr int QB = kSelect2QB;
    constexpr uint32_t NONE = 0xffffffffu;
    static_assert(BPT % 4 == 0 && BPT <= 16 && QB <= NW && 3 * QB <= 9, "layout");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* h1 = reinterpret_cast<uint32_t*>(smem_raw);          // [NB] bucket weights (zero between rows)
    uint32_t* tbl32 = h1 + NB;                                     // [NB] x 16 bits: slot arrays a bucket feeds
    uint32_t* l2 = tbl32 + NB / 2 + 4;                             // [QB][3][SL] slots: bp | bq | bn (zero between rows)
    __shared__ uint32_t s_red[NW], s_first[NW];
    __shared__ int s_last[NW];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_tau[QB];                      // target running weight
    __shared__ uint32_t s_b[QB][3];                     // buckets bp, bq, bn (NONE: no such bucket)
    __shared__ uint32_t s_base[QB];                     // running weight before bq

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int SL = select2_slots(p.shift1);             // slots per bucket (padded to one 128-bit chunk)
    const uint32_t slot_mask = (1u << p.shift1) - 1u;
    const int nch = SL >> 2;                            // 128-bit chunks per slot array
    const uint32_t shift = static_cast<uint32_t>(p.shift1);
    const float scale = p.fx_scale;
    const uint32_t capw = static_cast<uint32_t>(p.capw);
    const uint32_t h1_s = static_cast<uint32_t>(__cvta_generic_to_shared(h1));
    const uint32_t tbl_s = static_cast<uint32_t>(__cvta_generic_to_shared(tbl32));
    const uint32_t l2_s = static_cast<uint32_t>(__cvta_generic_to_shared(l2));
    // this warp's segment of the stage
    const uint32_t stage_s = static_cast<uint32_t>(__cvta_generic_to_shared(l2 + QB * 3 * SL)) + warp * capw * 32u;

    // shared state starts at zero and every row leaves it at zero
    for (int i = tid; i < NB + NB / 2 + 4 + QB * 3 * SL; i += THREADS) h1[i] = 0u;
    __syncthreads();

    auto mark = [&](int q, int which, uint32_t b) {     // bucket b feeds slot array 3 * q + which
        atomicOr(&tbl32[b >> 1], (1u << (3 * q + which)) << (16 * (b & 1)));
    };

    const int64_t n_groups = (p.n + kSelect3RowGroup - 1) / kSelect3RowGroup;
    auto load_group = [&](int64_t g, uint4 (&buf)[IPT]) {          // list words of this thread's trees, 4 rows
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const int t = tid + i * THREADS;
            buf[i] = make_uint4(kCompactEmpty, kCompactEmpty, kCompactEmpty, kCompactEmpty);
            if (g < n_groups && t < p.T)
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
    auto unit_of = [&](uint32_t s) -> uint32_t {        // fixed-point weight of one bootstrap count of a leaf
        return __float2uint_rn(__fdiv_rn(scale, __uint2float_rn(s)));
    };

    int64_t g = blockIdx.x;
    if (g >= n_groups) return;
    uint4 cur[IPT], nxt[IPT];
    load_group(g, cur);
    load_group(g + gridDim.x, nxt);
    int l = 0;
    uint32_t w[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) w[i] = cur[i].x;
    uint32_t wtot = issue_copies(w);

    for (;;) {
        const int64_t pos = g * kSelect3RowGroup + l;

        // pass 1: bucket weights.  w == 0 (utils.py:55-57) cannot occur (counts are >= 1); padding
        // slots (m == 0) are predicated off.
        cp_async_wait_all();
        __syncwarp();
        for_each_octet(wtot, w, [&](const uint4& a, const uint4& b) {
            const uint32_t U = unit_of(a.x);
            const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 7; ++e)
                red_add_shared_unless(h1_s + (((m[e] & 0xffffffu) >> shift) << 2), (m[e] >> 24) * U, m[e] == 0u ? 1u : 0u);
        });
        __syncthreads();                                // B1

        // inclusive prefix of this thread's BPT consecutive buckets, in registers; the histogram
        // goes back to zero.  Alongside: last non-empty bucket up to here / first one from here on.
        uint32_t c[BPT];
#pragma unroll
        for (int k = 0; k < BPT / 4; ++k) {
            const uint4 x = reinterpret_cast<const uint4*>(h1)[tid * (BPT / 4) + k];
            c[4 * k] = x.x; c[4 * k + 1] = x.y; c[4 * k + 2] = x.z; c[4 * k + 3] = x.w;
            reinterpret_cast<uint4*>(h1)[tid * (BPT / 4) + k] = make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t mine = 0, nz = 0;
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            mine += c[k];
            nz |= (c[k] != 0u ? 1u : 0u) << k;
        }
        uint32_t incl = mine;
        int pm = nz ? (tid * BPT + 31 - __clz(nz)) : -1;                     // last non-empty bucket of this thread
        uint32_t sm = nz ? static_cast<uint32_t>(tid * BPT + __ffs(nz) - 1) : NONE;   // first one
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            const int v = __shfl_up_sync(0xffffffffu, pm, d);
            const uint32_t x = __shfl_down_sync(0xffffffffu, sm, d);
            if (lane >= d) { incl += u; pm = max(pm, v); }
            if (lane + d < 32) sm = min(sm, x);
        }
        if (lane == 31) { s_red[warp] = incl; s_last[warp] = pm; }
        if (lane == 0) s_first[warp] = sm;
        int last_before = __shfl_up_sync(0xffffffffu, pm, 1);               // non-empty bucket before this thread's
        uint32_t first_after = __shfl_down_sync(0xffffffffu, sm, 1);        // ... and after
        if (lane == 0) last_before = -1;
        if (lane == 31) first_after = NONE;
        __syncthreads();                                // B2
        uint32_t run0 = incl - mine;                    // running weight before this thread's buckets
        uint32_t total = 0;
#pragma unroll
        for (int v = 0; v < NW; ++v) {
            const uint32_t x = s_red[v];
            run0 += (v < warp) ? x : 0u;
            total += x;
            if (v < warp) last_before = max(last_before, s_last[v]);
            if (v > warp) first_after = min(first_after, s_first[v]);
        }
        const int64_t out_row = p.order ? static_cast<int64_t>(__ldg(p.order + pos)) : pos;
        double* __restrict__ out = p.out + out_row * p.ldo;
        if (total != 0u) {
            // target running weight of every percentile (lane q of every warp, then broadcast)
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
                if (run0 < tau && tau <= c[BPT - 1]) {  // exactly one thread: the crossing bucket bq is one of its own
                    int k = 0;
                    uint32_t before = run0;
#pragma unroll
                    for (int j = 0; j < BPT - 1; ++j)
                        if (c[j] < tau) { k = j + 1; before = c[j]; }
                    const uint32_t below = nz & ((1u << k) - 1u);
                    const uint32_t above = nz >> (k + 1);
                    const uint32_t bq = static_cast<uint32_t>(tid * BPT + k);
                    const uint32_t bp = below ? static_cast<uint32_t>(tid * BPT + 31 - __clz(below))
                                              : (last_before < 0 ? NONE : static_cast<uint32_t>(last_before));
                    const uint32_t bn = above ? static_cast<uint32_t>(tid * BPT + k + __ffs(above)) : first_after;
                    s_b[q][0] = bp;
                    s_b[q][1] = bq;
                    s_b[q][2] = bn;
                    s_base[q] = before;
                    s_tau[q] = tau;
                    mark(q, 1, bq);
                    if (bp != NONE) mark(q, 0, bp);
                    if (bn != NONE) mark(q, 2, bn);
                }
            }
            if (tid == 0) s_total = total;
            __syncthreads();                            // B3

            // pass 2: members of marked buckets -> slot arrays at single-rank resolution
            for_each_octet(wtot, w, [&](const uint4& a, const uint4& b) {
                const uint32_t U = unit_of(a.x);
                const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 7; ++e) {
                    const uint32_t rank = m[e] & 0xffffffu;
                    uint32_t code = ld_shared_u16(tbl_s + (rank >> shift) * 2u);
                    if (code) {                         // rare: ~50 members of a row's 3300
                        const uint32_t wv = (m[e] >> 24) * U;
                        const uint32_t dst = l2_s + (rank & slot_mask) * 4u;
                        while (code) {
                            const int s = __ffs(code) - 1;
                            code &= code - 1u;
                            red_add_shared(dst + s * SL * 4, wv);
                        }
                    }
                }
            });
        }

        // next row: of this group, or the first one of the CTA's next group
        __syncwarp();                                   // every lane of the warp is done with the stage
        int64_t g2 = g;
        int l2n = l + 1;
        bool more = true;
        uint32_t w2[IPT];
        if (l2n < kSelect3RowGroup && pos + 1 < p.n) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) w2[i] = pick(cur[i], l2n);
        } else {
            g2 = g + gridDim.x;
            l2n = 0;
            more = g2 < n_groups;
#pragma unroll
            for (int i = 0; i < IPT; ++i) w2[i] = nxt[i].x;
        }
        uint32_t wtot2 = 0u;
        if (more) wtot2 = issue_copies(w2);             // block-uniform

        if (total == 0u) {                              // no member with a non-zero weight: cannot happen for a fitted forest
            if (tid < p.Q) out[tid] = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            __syncthreads();                            // B4: slots complete
            if (tid < 3 * p.Q) {                        // table back to zero (next marks: after the next row's B2)
                const uint32_t b = s_b[tid / 3][tid % 3];
                if (b != NONE) atomicAnd(&tbl32[b >> 1], ~(0xffffu << (16 * (b & 1))));
            }
            // select: warp q takes percentile q; the other warps go on to the next row
            if (warp < p.Q) {
                const int q = warp;
                uint32_t* L = l2 + warp * 3 * SL;
                uint4* M4 = reinterpret_cast<uint4*>(L + SL);
                const uint32_t tau = s_tau[q];
                const uint32_t bp = s_b[q][0], bq = s_b[q][1], bn = s_b[q][2];
                uint32_t run = s_base[q];
                uint4 v[4];
                uint32_t e_loc = NONE, We = 0u, Ce = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = j * 32 + lane;
                    v[j] = make_uint4(0u, 0u, 0u, 0u);
                    if (j * 32 < nch) {                  // warp-uniform
                        if (ch < nch) v[j] = M4[ch];
                        const uint32_t s4 = v[j].x + v[j].y + v[j].z + v[j].w;
                        uint32_t in = s4;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t u = __shfl_up_sync(0xffffffffu, in, d);
                            if (lane >= d) in += u;
                        }
                        uint32_t before = run + in - s4;
                        if (before < tau && before + s4 >= tau) {           // the crossing is in this chunk
                            const uint32_t vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                            bool got = false;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t after = before + vv[i];
                                if (!got && after >= tau) { got = true; e_loc = static_cast<uint32_t>(4 * ch + i); We = vv[i]; Ce = after; }
                                before = after;
                            }
                        }
                        run += __shfl_sync(0xffffffffu, in, 31);
                    }
                }
                const uint32_t who = __ballot_sync(0xffffffffu, e_loc != NONE);
                const int src = __ffs(who) - 1;         // exactly one lane (base < tau <= weight up to bq)
                const uint32_t e = __shfl_sync(0xffffffffu, e_loc, src);
                We = __shfl_sync(0xffffffffu, We, src);
                Ce = __shfl_sync(0xffffffffu, Ce, src);
                // neighbours of e inside bq: last non-empty slot before (as slot + 1), first after
                uint32_t pin = 0u, sin = NONE;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t s0 = static_cast<uint32_t>(4 * (j * 32 + lane));
                    const uint32_t vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (vv[i] != 0u && s0 + i < e) pin = max(pin, s0 + i + 1u);
                        if (vv[i] != 0u && s0 + i > e) sin = min(sin, s0 + i);
                    }
                }
                pin = __reduce_max_sync(0xffffffffu, pin);
                sin = __reduce_min_sync(0xffffffffu, sin);
                uint32_t rp = NONE, rs = NONE, Wp = 0u, Ws = 0u;            // ranks and merged weights of pred / succ
                if (pin != 0u) {
                    rp = (bq << p.shift1) + pin - 1u;
                    Wp = L[SL + pin - 1u];
                } else if (bp != NONE) {                // largest rank of bp
                    uint32_t mx = 0u;
                    for (int ch = lane; ch < nch; ch += 32) {
                        const uint4 x = reinterpret_cast<const uint4*>(L)[ch];
                        if (x.x) mx = max(mx, 4u * ch + 1u);
                        if (x.y) mx = max(mx, 4u * ch + 2u);
                        if (x.z) mx = max(mx, 4u * ch + 3u);
                        if (x.w) mx = max(mx, 4u * ch + 4u);
                    }
                    mx = __reduce_max_sync(0xffffffffu, mx);
                    rp = (bp << p.shift1) + mx - 1u;
                    Wp = L[mx - 1u];
                }
                if (sin != NONE) {
                    rs = (bq << p.shift1) + sin;
                    Ws = L[SL + sin];
                } else if (bn != NONE) {                // smallest rank of bn
                    uint32_t mn = NONE;
                    for (int ch = lane; ch < nch; ch += 32) {
                        const uint4 x = reinterpret_cast<const uint4*>(L + 2 * SL)[ch];
                        if (x.w) mn = min(mn, 4u * ch + 3u);
                        if (x.z) mn = min(mn, 4u * ch + 2u);
                        if (x.y) mn = min(mn, 4u * ch + 1u);
                        if (x.x) mn = min(mn, 4u * ch);
                    }
                    mn = __reduce_min_sync(0xffffffffu, mn);
                    rs = (bn << p.shift1) + mn;
                    Ws = L[2 * SL + mn];
                }
                __syncwarp();
                for (int ch = lane; ch < 3 * nch; ch += 32)                  // slots back to zero
                    reinterpret_cast<uint4*>(L)[ch] = make_uint4(0u, 0u, 0u, 0u);
                if (lane == 0) {                        // utils.py:69-81 on (pred, e, succ), float32
                    const double qv = ql.q[q];
                    const uint32_t total_w = s_total;
                    const float inv = __fdiv_rn(1.0f, scale);
                    const float s = __fdiv_rn(100.0f, __fmul_rn(static_cast<float>(total_w), inv));   // utils.py:69,72
                    const uint32_t re = (bq << p.shift1) + e;
                    // the three candidate values are requested together (one HBM latency, not two)
                    const float a_e = __ldg(p.y_sorted + re);
                    const float a_s = __ldg(p.y_sorted + (rs != NONE ? rs : re));
                    const float a_p = __ldg(p.y_sorted + (rp != NONE ? rp : re));
                    const float p_e = plotting_position(s, __fmul_rn(static_cast<float>(Ce), inv),
                                                        __fmul_rn(static_cast<float>(We), inv));
                    float result = a_e;
                    if (static_cast<double>(p_e) < qv) {                    // interval (e, succ)
                        if (rs != NONE) {               // else q lies beyond the last entry: a_last (utils.py:74-75)
                            const float p_s = plotting_position(s, __fmul_rn(static_cast<float>(Ce + Ws), inv),
                                                                __fmul_rn(static_cast<float>(Ws), inv));
                            const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_e), __fsub_rn(p_s, p_e));
                            result = __fadd_rn(a_e, __fmul_rn(frac, __fsub_rn(a_s, a_e)));
                        }
                    } else if (rp != NONE) {            // interval (pred, e); else before the first entry: a_0 (utils.py:76-77)
                        const float p_p = plotting_position(s, __fmul_rn(static_cast<float>(Ce - We), inv),
                                                            __fmul_rn(static_cast<float>(Wp), inv));
                        const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_p), __fsub_rn(p_e, p_p));
                        result = __fadd_rn(a_p, __fmul_rn(frac, __fsub_rn(a_e, a_p)));
                    }
                    out[q] = static_cast<double>(result);
                }
            }
        }
        if (!more) break;
        if (g2 != g) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) cur[i] = nxt[i];
            load_group(g2 + gridDim.x, nxt);
        }
        g = g2;
        l = l2n;
        wtot = wtot2;
#pragma unroll
        for (int i = 0; i < IPT; ++i) w[i] = w2[i];
    }
}

}  // namespace qf
