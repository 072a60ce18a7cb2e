// K2, QF_MODE_FAST, single-refinement form -- weighted percentile by histogram selection for
// training sets of up to 2^(LOG2NB + 9) samples (2M with LOG2NB = 12).  Larger training sets
// take the multi-level kernel of select_kernel.cuh.
//
// Same arithmetic contract as select_kernel.cuh: weights as 32-bit fixed point (exact sums
// whatever the order; duplicates of a training sample merge by landing in the same slot),
// utils.py:72-81 in float32 on the three entries (pred, e, succ) around the crossing.
//
// What changes is the amount of work per member and the number of block barriers per row:
//
//   table    every warp owns T/8 consecutive leaves of the row.  It turns their (start, count)
//            pairs -- fetched one row ahead into registers -- into a flat table of aligned
//            128-bit member PAIRS (entry = pair index | "first slot is outside the list" |
//            "second slot is outside the list"), so that both passes below run with all 32
//            lanes busy whatever the leaf sizes (4 lanes per leaf wasted half of them).
//   pass 1   one 128-bit load per lane and pair; ONE shared-memory atomic per member: weight
//            histogram over NB buckets of the rank axis (bucket = rank >> shift1).
//   scan     block scan of the bucket weights, kept in registers (BPT per thread).  Every
//            thread then looks for the requested percentiles among its own buckets -- no
//            binary searches, no dependent shared-memory chains:
//              bq = bucket where the running weight reaches tau = ceil(q/100 * total)
//              bp = last non-empty bucket before bq, bn = first non-empty bucket after bq
//            (bp / bn hold the neighbours of the crossing entry when it is the first / last
//            entry of bq).  A 16-bit table entry per bucket says which slot arrays it feeds.
//   pass 2   runs over the bucket ids pass 1 left in shared memory (two 16-bit ids per pair): a
//            pair whose buckets feed nothing costs three shared loads and a test and no
//            global traffic.  The few others (about 50 of 1900 pairs at C3) are first compacted
//            -- in place, over summaries already read -- then loaded again (L2 hits) with all
//            lanes busy and added to the slot arrays [bp | bq | bn] of their percentile(s) at
//            single-rank resolution.
//   select   one WARP per percentile scans its bq slots (<= 512, four 128-bit loads per
//            lane), finds the crossing entry e, its neighbours inside bq or, failing that,
//            the largest rank of bp / the smallest rank of bn; lane 0 interpolates.  The other
//            warps are already in pass 1 of the CTA's next row.
//
// Six block barriers per row; histogram, table and slots are returned to zero by the
// threads that read them, so there is no clearing phase.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "quantile_kernel.cuh"
#include "select_kernel.cuh"

namespace qf {

constexpr int kSelect2QB = 3;                           // percentiles per pass 2
constexpr int kSelect2SlotBits = 9;                     // a bucket spans <= 512 ranks
constexpr int kSelect2Threads = 256;
constexpr int kSelect2Warps = kSelect2Threads / 32;
constexpr int kSelect2MaxIPT = 4;                       // leaves per lane: T <= 4 * 256
constexpr int kSelect2CtasPerSm = 4;
constexpr int kSelect2MaxPairs = 6144;                  // pair-table entries per CTA (24 KB, and as much again for the bucket summaries)

// first pair-table entry of every warp (exact upper bounds from the largest leaf of each tree)
struct Select2Layout {
    uint32_t warp_off[kSelect2Warps + 1];
};

__host__ __device__ constexpr int select2_slots(int shift1) { return (1 << shift1) < 4 ? 4 : (1 << shift1); }

__host__ __device__ inline size_t select2_kernel_smem(int log2nb, int shift1, uint32_t pairs) {
    const size_t nb = static_cast<size_t>(1) << log2nb;
    return nb * 4 + (nb * 2 + 16) + static_cast<size_t>(kSelect2QB) * 3 * select2_slots(shift1) * 4 +
           2 * ((static_cast<size_t>(pairs) + 3) & ~size_t(3)) * 4;       // pair table + bucket summaries
}

// shared-memory accesses by 32-bit shared address (no generic -> shared conversion in the loops)
__device__ __forceinline__ void red_add_shared(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// red.shared.add.u32 [addr], v  executed only when `skip` is zero (predicated, no branch)
__device__ __forceinline__ void red_add_shared_unless(uint32_t addr, uint32_t v, uint32_t skip) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                 ::"r"(addr), "r"(v), "r"(skip) : "memory");
}

// Pair-table entry: bits 0..29 = index of the aligned 128-bit member pair (member index / 2),
// bit 30 = the pair's first slot lies outside the leaf list, bit 31 = its second slot does.
constexpr uint32_t kPairSkip0 = 0x40000000u, kPairSkip1 = 0x80000000u, kPairIndex = 0x3fffffffu;

// IPT: leaves per lane (1, 2 or 4): covers T <= IPT * 256 trees; a template parameter so that the
// per-lane loops over them carry no dead iterations
template <int LOG2NB, int IPT>
__global__ void __launch_bounds__(kSelect2Threads, kSelect2CtasPerSm)
quantile_select2_kernel(const QuantileParams p, const __grid_constant__ QuantileList ql,
                        const __grid_constant__ Select2Layout lay, const float fx_scale, const int shift1) {
    constexpr int THREADS = kSelect2Threads;
    constexpr int NB = 1 << LOG2NB;
    constexpr int BPT = NB / THREADS;                   // buckets scanned per thread
    constexpr int NW = kSelect2Warps;
    constexpr int QB = kSelect2QB;
    constexpr int MAXIPT = IPT;
    constexpr uint32_t NONE = 0xffffffffu;
    static_assert(BPT % 4 == 0 && QB <= NW && 3 * QB <= 9 && kSelect2MaxPairs <= (1 << 13), "layout");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* h1 = reinterpret_cast<uint32_t*>(smem_raw);          // [NB] bucket weights (zero between rows)
    uint32_t* tbl32 = h1 + NB;                                     // [NB] x 16 bits: slot arrays a bucket feeds (zero between passes)
    uint32_t* l2 = tbl32 + NB / 2 + 4;                             // [QB][3][SL] slots: bp | bq | bn (zero between passes);
                                                                   //   table entries NB.. stay zero ("no bucket")
    __shared__ uint32_t s_red[NW];
    __shared__ uint32_t s_total;
    // per percentile of the call
    __shared__ uint32_t s_tau[kSelectMaxQ];             // target running weight
    __shared__ uint32_t s_b[kSelectMaxQ][3];            // buckets bp, bq, bn (NONE: no such bucket)
    __shared__ uint32_t s_base[kSelectMaxQ], s_upto[kSelectMaxQ];   // running weight before / after bq

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int SL = select2_slots(shift1);               // slots per bucket (padded to one 128-bit chunk)
    const uint32_t slot_mask = (1u << shift1) - 1u;
    const int nch = SL >> 2;                            // 128-bit chunks per slot array
    uint32_t* pairtab = l2 + QB * 3 * SL;                          // [lay.warp_off[NW]] pair table, one segment per warp

    // 32-bit shared addresses of the arrays, passed through shared memory once: the compiler
    // then keeps them in registers instead of re-deriving them in front of every access
    __shared__ uint32_t s_addr[4];
    __shared__ uint64_t s_members;
    __shared__ float s_scale;
    __shared__ uint32_t s_shift;
    if (tid == 0) {
        s_members = reinterpret_cast<uint64_t>(p.members);
        s_scale = fx_scale;
        s_shift = static_cast<uint32_t>(shift1);
        s_addr[0] = static_cast<uint32_t>(__cvta_generic_to_shared(h1));
        s_addr[1] = static_cast<uint32_t>(__cvta_generic_to_shared(tbl32));
        s_addr[2] = static_cast<uint32_t>(__cvta_generic_to_shared(l2));
        s_addr[3] = static_cast<uint32_t>(__cvta_generic_to_shared(pairtab));
    }
    // shared state starts at zero and every row leaves it at zero
    for (int i = tid; i < NB + NB / 2 + 4 + QB * 3 * SL; i += THREADS) h1[i] = 0u;
    __syncthreads();
    const uint32_t h1_s = s_addr[0], tbl_s = s_addr[1], l2_s = s_addr[2];
    const uint4* __restrict__ members4 = reinterpret_cast<const uint4*>(s_members);   // member pairs
    const float scale = s_scale;
    const uint32_t shift = s_shift;
    const uint32_t tab_s = s_addr[3] + lay.warp_off[warp] * 4;    // this warp's segment of the pair table
    // bucket summaries of the same pairs, written by pass 1 and read by pass 2 (second half)
    const uint32_t sum_s = tab_s + ((lay.warp_off[NW] + 3u) & ~3u) * 4;

    auto mark = [&](int q, int which, uint32_t b) {     // bucket b feeds slot array 3 * (q % QB) + which
        atomicOr(&tbl32[b >> 1], (1u << (3 * (q % QB) + which)) << (16 * (b & 1)));
    };

    // leaves of this lane: ipt consecutive ones; the warp owns 32 * ipt consecutive leaves
    constexpr int ipt = IPT;
    const int t_first = (warp * 32 + lane) * ipt;
    const int64_t n_work = p.row_list ? static_cast<int64_t>(*p.row_count) : p.n;
    auto row_of = [&](int64_t work) -> int64_t {
        return p.row_list ? static_cast<int64_t>(p.row_list[work]) : work;
    };
    auto fetch = [&](int64_t work, uint2 (&sg)[MAXIPT]) {          // (start, count) of this lane's leaves
        const uint2* __restrict__ seg = p.seg + (work < n_work ? row_of(work) : 0) * p.T;
#pragma unroll
        for (int i = 0; i < MAXIPT; ++i) {
            const int t = t_first + i;
            sg[i] = (work < n_work && i < ipt && t < p.T) ? seg[t] : make_uint2(0u, 0u);
        }
    };
    uint2 pre[MAXIPT];
    fetch(blockIdx.x, pre);

    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int64_t row = row_of(work);

        // pair table of this warp's leaves (warp-private: no block barrier), then the next
        // row's leaf lists are requested
        uint32_t n_pairs;
        {
            uint32_t np[MAXIPT], tot = 0;
#pragma unroll
            for (int i = 0; i < MAXIPT; ++i) {
                np[i] = pre[i].y ? (((pre[i].x & 1u) + pre[i].y + 1u) >> 1) : 0u;
                tot += np[i];
            }
            uint32_t incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            n_pairs = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t dst = tab_s + (incl - tot) * 4;
#pragma unroll
            for (int i = 0; i < MAXIPT; ++i) {
                if (i < ipt) {                          // block-uniform
                    const uint32_t last = (pre[i].x + pre[i].y - 1u) >> 1;   // pair holding the list's last member
                    uint32_t pair = pre[i].x >> 1;
                    if (np[i]) {                        // first pair: its first slot may precede the list
                        st_shared_u32(dst, pair | ((pre[i].x & 1u) ? kPairSkip0 : 0u) |
                                               ((pair == last && ((pre[i].x + pre[i].y) & 1u)) ? kPairSkip1 : 0u));
                        ++pair, dst += 4;
                    }
                    for (uint32_t k = 1; k < np[i]; ++k, ++pair, dst += 4)
                        st_shared_u32(dst, pair | ((pair == last && ((pre[i].x + pre[i].y) & 1u)) ? kPairSkip1 : 0u));
                }
            }
            __syncwarp();
        }
        fetch(work + gridDim.x, pre);

        // pass 1: bucket weights.  Two pairs (four members) per lane and iteration, all lanes
        // busy; w == 0 (utils.py:55-57) changes nothing.  The buckets of every pair are kept
        // (16 bits each, NB = slot outside its leaf list) so that pass 2 need not load it again.
        for (uint32_t j = lane; j < n_pairs; j += 64) {
            const bool two = j + 32 < n_pairs;
            const uint32_t e0 = ld_shared_u32(tab_s + j * 4);
            const uint32_t e1 = two ? ld_shared_u32(tab_s + (j + 32) * 4) : (kPairSkip0 | kPairSkip1);
            const uint4 a = __ldg(members4 + (e0 & kPairIndex));
            const uint4 b = __ldg(members4 + (e1 & kPairIndex));
            const uint32_t a0 = (e0 & kPairSkip0) ? NB : (a.x >> shift), a1 = (e0 & kPairSkip1) ? NB : (a.z >> shift);
            const uint32_t b0 = (e1 & kPairSkip0) ? NB : (b.x >> shift), b1 = (e1 & kPairSkip1) ? NB : (b.z >> shift);
            red_add_shared_unless(h1_s + a0 * 4, __float2uint_rn(__fmul_rn(__uint_as_float(a.y), scale)), e0 & kPairSkip0);
            red_add_shared_unless(h1_s + a1 * 4, __float2uint_rn(__fmul_rn(__uint_as_float(a.w), scale)), e0 & kPairSkip1);
            red_add_shared_unless(h1_s + b0 * 4, __float2uint_rn(__fmul_rn(__uint_as_float(b.y), scale)), e1 & kPairSkip0);
            red_add_shared_unless(h1_s + b1 * 4, __float2uint_rn(__fmul_rn(__uint_as_float(b.w), scale)), e1 & kPairSkip1);
            st_shared_u32(sum_s + j * 4, a0 | (a1 << 16));
            if (two) st_shared_u32(sum_s + (j + 32) * 4, b0 | (b1 << 16));
        }
        __syncthreads();                                // B1

        // inclusive prefix of this thread's BPT consecutive buckets, in registers; the
        // histogram goes back to zero
        uint32_t c[BPT];
#pragma unroll
        for (int k = 0; k < BPT / 4; ++k) {
            const uint4 x = reinterpret_cast<const uint4*>(h1)[tid * (BPT / 4) + k];
            c[4 * k] = x.x; c[4 * k + 1] = x.y; c[4 * k + 2] = x.z; c[4 * k + 3] = x.w;
            reinterpret_cast<uint4*>(h1)[tid * (BPT / 4) + k] = make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < BPT; ++k) mine += c[k];
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s_red[warp] = incl;
        __syncthreads();                                // B2
        uint32_t run0 = incl - mine;                    // running weight before this thread's buckets
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t v = s_red[w];
            run0 += (w < warp) ? v : 0u;
            total += v;
        }
        if (total == 0u) {                              // no member with a non-zero weight: cannot happen for a fitted forest
            double* __restrict__ out = p.out + (p.order ? static_cast<int64_t>(__ldg(p.order + row)) : row) * p.ldo;
            for (int qi = tid; qi < p.Q; qi += THREADS) out[qi] = __longlong_as_double(0x7ff8000000000000LL);
            __syncthreads();                            // s_red is rewritten by the next row
            continue;
        }
        {
            uint32_t run = run0;
#pragma unroll
            for (int k = 0; k < BPT; ++k) { run += c[k]; c[k] = run; }
        }
        if (tid < p.Q) {
            const double t = ceil(ql.q[tid] * static_cast<double>(total) / 100.0);
            s_tau[tid] = static_cast<uint32_t>(fmin(fmax(t, 1.0), static_cast<double>(total)));
            s_b[tid][0] = NONE;
            s_b[tid][2] = NONE;
        }
        if (tid == 0) s_total = total;
        __syncthreads();                                // B3

        // bq: the bucket in which the running weight reaches tau
        for (int q = 0; q < p.Q; ++q) {
            const uint32_t tau = s_tau[q];
            if (run0 < tau && tau <= c[BPT - 1]) {      // exactly one thread
                int k = 0;
                uint32_t before = run0;
#pragma unroll
                for (int j = 0; j < BPT - 1; ++j)
                    if (c[j] < tau) { k = j + 1; before = c[j]; }
                uint32_t upto = c[0];
#pragma unroll
                for (int j = 1; j < BPT; ++j)
                    if (j == k) upto = c[j];
                const uint32_t b = static_cast<uint32_t>(tid * BPT + k);
                s_b[q][1] = b;
                s_base[q] = before;
                s_upto[q] = upto;
                if (q < QB) mark(q, 1, b);              // first batch: marked here, later ones below
            }
        }
        __syncthreads();                                // B4

        // bp: last non-empty bucket before bq; bn: first non-empty bucket after bq
        for (int q = 0; q < p.Q; ++q) {
            const uint32_t base = s_base[q], upto = s_upto[q];
            if (base != 0u && run0 < base && base <= c[BPT - 1]) {      // first bucket whose prefix reaches base
                int k = 0;
#pragma unroll
                for (int j = 0; j < BPT - 1; ++j)
                    if (c[j] < base) k = j + 1;
                const uint32_t b = static_cast<uint32_t>(tid * BPT + k);
                s_b[q][0] = b;
                if (q < QB) mark(q, 0, b);
            }
            if (upto != total && run0 <= upto && upto < c[BPT - 1]) {   // first bucket whose prefix exceeds upto
                int k = 0;
#pragma unroll
                for (int j = 0; j < BPT - 1; ++j)
                    if (c[j] <= upto) k = j + 1;
                const uint32_t b = static_cast<uint32_t>(tid * BPT + k);
                s_b[q][2] = b;
                if (q < QB) mark(q, 2, b);
            }
        }
        __syncthreads();                                // B5

        for (int q0 = 0; q0 < p.Q; q0 += QB) {
            const int nqb = min(QB, p.Q - q0);
            if (q0 > 0) {                               // later batches: mark their buckets now
                __syncthreads();                        //   (the previous batch is done with table and slots)
                if (tid < 3 * nqb) {
                    const uint32_t b = s_b[q0 + tid / 3][tid % 3];
                    if (b != NONE) mark(tid / 3, tid % 3, b);
                }
                __syncthreads();
            }

            // pass 2: members of marked buckets -> slot arrays at single-rank resolution.  Only the
            // bucket summaries are read back; the few pairs that touch a marked bucket are
            // loaded again (L2 hits)
            auto feed = [&](uint32_t code, uint32_t rank, uint32_t wbits) {
                const uint32_t w = __float2uint_rn(__fmul_rn(__uint_as_float(wbits), scale));
                const uint32_t dst = l2_s + (rank & slot_mask) * 4;
                while (code) {
                    const int a = __ffs(code) - 1;
                    code &= code - 1u;
                    red_add_shared(dst + a * SL * 4, w);
                }
            };
            if (p.Q <= QB) {
                // one batch: the hits are first compacted (in place, over summaries already read)
                // and then fed with all lanes busy -- at C3 some 50 of a row's 1900 pairs hit,
                // spread over most of the 60 warp-iterations of the scan
                uint32_t n_hits = 0;                     // warp-uniform
                for (uint32_t j0 = 0; j0 < n_pairs; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    uint32_t c0 = 0u, c1 = 0u;
                    if (j < n_pairs) {
                        const uint32_t sm = ld_shared_u32(sum_s + j * 4);
                        c0 = ld_shared_u16(tbl_s + (sm & 0xffffu) * 2);
                        c1 = ld_shared_u16(tbl_s + (sm >> 16) * 2);
                    }
                    const uint32_t mask = __ballot_sync(0xffffffffu, (c0 | c1) != 0u);
                    if (mask) {                          // warp-uniform; entry = pair | codes (13 + 9 + 9 bits)
                        if (c0 | c1)
                            st_shared_u32(sum_s + (n_hits + __popc(mask & ((1u << lane) - 1u))) * 4, j | (c0 << 13) | (c1 << 22));
                        n_hits += __popc(mask);
                    }
                }
                __syncwarp();
                for (uint32_t k = lane; k < n_hits; k += 32) {
                    const uint32_t h = ld_shared_u32(sum_s + k * 4);
                    const uint4 a = __ldg(members4 + (ld_shared_u32(tab_s + (h & 0x1fffu) * 4) & kPairIndex));
                    feed((h >> 13) & 0x1ffu, a.x, a.y);
                    feed(h >> 22, a.z, a.w);
                }
            } else {
                for (uint32_t j = lane; j < n_pairs; j += 32) {
                    const uint32_t sm = ld_shared_u32(sum_s + j * 4);
                    const uint32_t c0 = ld_shared_u16(tbl_s + (sm & 0xffffu) * 2);
                    const uint32_t c1 = ld_shared_u16(tbl_s + (sm >> 16) * 2);
                    if (c0 | c1) {
                        const uint4 a = __ldg(members4 + (ld_shared_u32(tab_s + j * 4) & kPairIndex));
                        feed(c0, a.x, a.y);
                        feed(c1, a.z, a.w);
                    }
                }
            }
            __syncthreads();                            // B6
            if (tid < 3 * nqb) {                        // table back to zero (next marks: after the next row's B3)
                const uint32_t b = s_b[q0 + tid / 3][tid % 3];
                if (b != NONE) atomicAnd(&tbl32[b >> 1], ~(0xffffu << (16 * (b & 1))));
            }

            // select: warp w takes percentile q0 + w
            if (warp < nqb) {
                const int q = q0 + warp;
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
                    rp = (bq << shift1) + pin - 1u;
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
                    rp = (bp << shift1) + mx - 1u;
                    Wp = L[mx - 1u];
                }
                if (sin != NONE) {
                    rs = (bq << shift1) + sin;
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
                    rs = (bn << shift1) + mn;
                    Ws = L[2 * SL + mn];
                }
                __syncwarp();
                for (int ch = lane; ch < 3 * nch; ch += 32)                  // slots back to zero
                    reinterpret_cast<uint4*>(L)[ch] = make_uint4(0u, 0u, 0u, 0u);
                if (lane == 0) {                        // utils.py:69-81 on (pred, e, succ), float32
                    const double qv = ql.q[q];
                    const uint32_t total_w = s_total;
                    const float inv = __fdiv_rn(1.0f, fx_scale);
                    const float s = __fdiv_rn(100.0f, __fmul_rn(static_cast<float>(total_w), inv));   // utils.py:69,72
                    const uint32_t re = (bq << shift1) + e;
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
                    double* __restrict__ out = p.out + (p.order ? static_cast<int64_t>(__ldg(p.order + row)) : row) * p.ldo;
                    out[q] = static_cast<double>(result);
                }
            }
        }   // percentile batches
    }   // rows
}

// largest and smallest leaf (members) of every tree: block t reduces over the tree's packed
// nodes; out[t] = largest, out[T + t] = smallest
__global__ void __launch_bounds__(256)
tree_max_leaf_kernel(const uint2* __restrict__ nodes, const uint32_t* __restrict__ roots, int64_t n_nodes, int T,
                     uint32_t* __restrict__ out) {
    __shared__ uint32_t red[16];
    const int t = blockIdx.x;
    const int64_t n0 = roots[t], n1 = (t + 1 < T) ? static_cast<int64_t>(roots[t + 1]) : n_nodes;
    uint32_t m = 0, mn = 0xffffffffu;
    for (int64_t i = n0 + threadIdx.x; i < n1; i += blockDim.x) {
        const uint32_t hi = nodes[i].y;
        if (hi & 0x80000000u) {
            m = max(m, hi & 0x7fffffffu);
            mn = min(mn, hi & 0x7fffffffu);
        }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = m; red[8 + (threadIdx.x >> 5)] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { m = max(m, red[w]); mn = min(mn, red[8 + w]); }
        out[t] = m;
        out[T + t] = mn;
    }
}

// largest sum of member weights over the leaves of every tree (float bits, weights are > 0):
// 1.0 for forests (ensemble.py:94-99 normalises per leaf), the leaf size for unit weights
__global__ void __launch_bounds__(256)
tree_max_leaf_weight_kernel(const uint2* __restrict__ nodes, const uint32_t* __restrict__ roots, int64_t n_nodes, int T,
                            const uint2* __restrict__ members, uint32_t* __restrict__ out) {
    const int t = blockIdx.x;
    const int64_t n0 = roots[t], n1 = (t + 1 < T) ? static_cast<int64_t>(roots[t + 1]) : n_nodes;
    float mx = 0.0f;
    for (int64_t i = n0 + threadIdx.x; i < n1; i += blockDim.x) {
        const uint2 nd = nodes[i];
        if (nd.y & 0x80000000u) {
            const uint32_t cnt = nd.y & 0x7fffffffu;
            double s = 0.0;
            for (uint32_t k = 0; k < cnt; ++k) s += static_cast<double>(__uint_as_float(members[nd.x + k].y));
            mx = fmaxf(mx, static_cast<float>(s));
        }
    }
    atomicMax(out + t, __float_as_uint(mx));
}

}  // namespace qf
