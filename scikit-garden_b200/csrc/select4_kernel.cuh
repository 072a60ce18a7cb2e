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
//   state    8 KB of shared memory per row: a 1024-bucket histogram of the rank axis (4 KB), a 256-entry
//            hit list (2 KB), a list of "extra" octets (<= 2 KB); 24 rows in flight per SM, each on
//            its own warp, no block barrier anywhere (only __syncwarp).
//   members  read straight from global memory / L2 (two 128-bit loads per octet, no staging): sweep A
//            takes the first octet of each of the lane's leaves (perfectly balanced), sweep B the
//            octets beyond the first of all the warp's leaves, flattened through the small list.  The
//            second pass re-reads the same octets a few microseconds later: L2 hits.
//   pass 1   one shared-memory RED per member into the coarse histogram.
//   scan     lane totals of 32 consecutive buckets (bank-rotated reads), one warp scan; per percentile
//            the group that holds the crossing is scanned across the lanes: crossing bucket bq, nearest
//            non-empty neighbours bp / bn from ballots.  The marked buckets live in a 1024-bit map, 32
//            bits per lane, tested with one SHFL per member -- no table in shared memory.
//   pass 2   members of marked buckets (~100 of a row's 3300 at config 3) are appended to the hit list
//            (ballot + popc, no atomics).
//   select   the hits are sorted by rank in registers (bitonic network, 8 keys per lane), weights
//            prefix-summed in that order, and lane q finds percentile q's crossing entry and its
//            neighbours by binary search: because every member of bp, bq and bn is in the list, the
//            neighbours of the crossing entry are simply the adjacent rank groups of the sorted list.
//   overflow a row with more than 256 hits narrows each percentile exactly instead: sub-histograms of
//            the crossing bucket until it is one rank wide, then two more passes for the neighbours and
//            the merged weights.  Slow (a dozen passes) and rare; tests force it with `select4_hits`.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "compact_lists.cuh"
#include "quantile_kernel.cuh"
#include "select2_kernel.cuh"

namespace qf {

constexpr int kSelect4Threads = 256;
constexpr int kSelect4Warps = kSelect4Threads / 32;     // rows in flight per CTA
constexpr int kSelect4CtasPerSm = 3;
constexpr int kSelect4LogNB = 10;
constexpr int kSelect4NB = 1 << kSelect4LogNB;          // coarse buckets of the rank axis
constexpr int kSelect4Hits = 256;                       // hit list entries per row (8 keys per lane)
constexpr int kSelect4MaxLPL = 32;                      // leaves per lane: T <= 1024

struct Select4Params {
    const uint32_t* __restrict__ seg;      // [n][ldT] compact list words, ROW-major (transposed from K1's output)
    int64_t ldT;
    const int* __restrict__ order;         // [n] locality position -> output row; nullptr = identity
    const uint4* __restrict__ lists;       // two uint4 per octet
    const float* __restrict__ y_sorted;    // [N]
    double* __restrict__ out;              // [n][ldo], indexed by output row
    int64_t n;
    int64_t ldo;
    int T;
    int Q;                                 // <= kSelect2QB
    float fx_scale;
    int shift;                             // coarse bucket = rank >> shift
    int extras_cap;                        // entries of the per-warp list of octets beyond a leaf's first
    int hits_cap;                          // <= kSelect4Hits
};

__host__ __device__ inline size_t select4_warp_smem(int extras_cap) {
    return static_cast<size_t>(kSelect4NB) * 4 + static_cast<size_t>(kSelect4Hits) * 8 + static_cast<size_t>(extras_cap) * 4;
}

// K1 writes the list words tree-major ([T][ld], coalesced for its lane = row mapping); a warp that owns a
// row wants them row-major.  32 x 32 tile transpose through shared memory.
__global__ void __launch_bounds__(256)
transpose_words_kernel(const uint32_t* __restrict__ in, int64_t ld_in, int T, int64_t n, uint32_t* __restrict__ out,
                       int64_t ld_out) {
    __shared__ uint32_t tile[32][33];
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
    const int t0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {                  // rows of the tile = trees
        const int t = t0 + k;
        const int64_t r = r0 + tx;
        tile[k][tx] = (t < T && r < n) ? in[static_cast<int64_t>(t) * ld_in + r] : kCompactEmpty;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {                  // rows of the output = test rows
        const int64_t r = r0 + k;
        const int t = t0 + tx;
        if (r < n && t < ld_out) out[r * ld_out + t] = tile[tx][k];
    }
}

template <int LPL>
__global__ void __launch_bounds__(kSelect4Threads, kSelect4CtasPerSm)
quantile_select4_kernel(const Select4Params p, const __grid_constant__ QuantileList ql) {
    constexpr int NB = kSelect4NB;
    constexpr uint32_t NONE = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const size_t wbytes = select4_warp_smem(p.extras_cap);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + warp * wbytes);   // [NB], zero between rows
    uint32_t* hit_r = hist + NB;                                              // [256] ranks -> sorted ranks
    uint32_t* hit_w = hit_r + kSelect4Hits;                                   // [256] weights -> running sums
    uint32_t* extras = hit_w + kSelect4Hits;                                  // [extras_cap] octet indices
    const uint32_t hist_s = static_cast<uint32_t>(__cvta_generic_to_shared(hist));
    const uint32_t shift = static_cast<uint32_t>(p.shift);
    const uint32_t lt_lane = (1u << lane) - 1u;
    const float inv = __fdiv_rn(1.0f, p.fx_scale);

    for (int i = lane; i < NB; i += 32) hist[i] = 0u;
    __syncwarp();

    const int64_t n_warps = static_cast<int64_t>(gridDim.x) * kSelect4Warps;
    for (int64_t pos = static_cast<int64_t>(blockIdx.x) * kSelect4Warps + warp; pos < p.n; pos += n_warps) {
        // ---- the row's list words: leaf t = lane + 32 i
        uint32_t w[LPL];
#pragma unroll
        for (int i = 0; i < LPL; ++i) {
            const int t = lane + 32 * i;
            w[i] = t < p.T ? __ldg(p.seg + pos * p.ldT + t) : kCompactEmpty;
        }
        // first member octet and octet count of leaf i (a long list starts with an info octet)
        auto leaf = [&](uint32_t word, uint32_t& first, uint32_t& noct) {
            first = word & kCompactOffMask;
            noct = word == kCompactEmpty ? 0u : (word >> 28) + 1u;
            if (noct > static_cast<uint32_t>(kCompactMaxOctets)) {
                noct = __ldg(p.lists + 2 * static_cast<size_t>(first)).x;
                ++first;
            }
        };
        // ---- octets beyond the first of every leaf -> flat list (balanced second sweep)
        uint32_t n_extra;
        {
            uint32_t mine = 0u;
#pragma unroll
            for (int i = 0; i < LPL; ++i) {
                uint32_t first, noct;
                leaf(w[i], first, noct);
                mine += noct > 1u ? noct - 1u : 0u;
            }
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            n_extra = __shfl_sync(0xffffffffu, incl, 31);
            if (n_extra <= static_cast<uint32_t>(p.extras_cap)) {
                uint32_t at = incl - mine;
#pragma unroll
                for (int i = 0; i < LPL; ++i) {
                    uint32_t first, noct;
                    leaf(w[i], first, noct);
                    for (uint32_t k = 1; k < noct; ++k) extras[at++] = first + k;
                }
            }
            __syncwarp();
        }
        const bool flat = n_extra <= static_cast<uint32_t>(p.extras_cap);

        // f(ok, a, b) on every octet of the row, all 32 lanes in step (f may vote); ok = false: a lane
        // without an octet in this step (it sees octet 0 of the lists, to be ignored)
        auto each_octet = [&](auto&& f) {
#pragma unroll
            for (int i0 = 0; i0 < LPL; i0 += 4) {       // sweep A: the first octet of four leaves per lane at a time
                uint4 a[4], b[4];
                bool ok[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (i0 + k < LPL) {
                        uint32_t first, noct;
                        leaf(w[i0 + k], first, noct);
                        ok[k] = noct != 0u;
                        const uint4* __restrict__ src = p.lists + 2 * static_cast<size_t>(ok[k] ? first : 0u);
                        a[k] = __ldg(src);
                        b[k] = __ldg(src + 1);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (i0 + k < LPL) f(ok[k], a[k], b[k]);
            }
            if (flat) {                                 // sweep B: the other octets, flat over the lanes
                for (uint32_t j0 = 0; j0 < n_extra; j0 += 32u) {
                    const bool ok = j0 + lane < n_extra;
                    const uint4* __restrict__ src = p.lists + 2 * static_cast<size_t>(ok ? extras[j0 + lane] : 0u);
                    f(ok, __ldg(src), __ldg(src + 1));
                }
            } else {                                    // list too small for this row: leaf by leaf, lanes in step
#pragma unroll 1
                for (int i = 0; i < LPL; ++i) {
                    uint32_t first, noct;
                    leaf(w[i], first, noct);
                    const uint32_t most = __reduce_max_sync(0xffffffffu, noct);
                    for (uint32_t k = 1; k < most; ++k) {
                        const bool ok = k < noct;
                        const uint4* __restrict__ src = p.lists + 2 * static_cast<size_t>(ok ? first + k : 0u);
                        f(ok, __ldg(src), __ldg(src + 1));
                    }
                }
            }
        };

        // ---- pass 1: coarse bucket weights
        each_octet([&](bool ok, const uint4& a, const uint4& b) {
            const uint32_t U = ok ? a.x : 0u;
            const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 7; ++e)
                red_add_shared(hist_s + (((m[e] & 0xffffffu) >> shift) << 2), (m[e] >> 24) * U);
        });
        __syncwarp();

        // ---- scan: lane l owns buckets [32 l, 32 l + 32); reads rotated by the lane (no bank conflicts)
        uint32_t mine = 0u;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) mine += ld_shared_u32(hist_s + (lane * 32 + ((j + lane) & 31)) * 4);
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
        uint32_t k_bq = 0u, k_bp = NONE, k_bn = NONE, k_base = 0u, k_tau = 0u;   // lane q: percentile q
        uint32_t marks = 0u;                            // bit j: bucket 32 * lane + j is marked
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
            if (lane == q) { k_bq = bq; k_bp = bp; k_bn = bn; k_base = base; k_tau = tau; }
            if (static_cast<int>(bq >> 5) == lane) marks |= 1u << (bq & 31u);
            if (bp != NONE && static_cast<int>(bp >> 5) == lane) marks |= 1u << (bp & 31u);
            if (bn != NONE && static_cast<int>(bn >> 5) == lane) marks |= 1u << (bn & 31u);
        }
        // the histogram goes back to zero
#pragma unroll 8
        for (int j = 0; j < 32; ++j) st_shared_u32(hist_s + (lane * 32 + ((j + lane) & 31)) * 4, 0u);
        __syncwarp();

        // ---- pass 2: members of marked buckets -> hit list (padding slots have count 0: skipped)
        uint32_t nh = 0u;                               // warp-uniform
        each_octet([&](bool ok, const uint4& a, const uint4& b) {
            const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 7; ++e) {
                const uint32_t rank = m[e] & 0xffffffu, bk = rank >> shift;
                const uint32_t bit = (__shfl_sync(0xffffffffu, marks, bk >> 5) >> (bk & 31u)) & 1u;
                const bool hit = ok && bit != 0u && (m[e] >> 24) != 0u;
                const uint32_t votes = __ballot_sync(0xffffffffu, hit);
                if (votes) {
                    const uint32_t at = nh + __popc(votes & lt_lane);
                    if (hit && at < static_cast<uint32_t>(kSelect4Hits)) {
                        hit_r[at] = rank;
                        hit_w[at] = (m[e] >> 24) * a.x;
                    }
                    nh += __popc(votes);
                }
            }
        });
        __syncwarp();

        // per percentile (lane q): ranks of the crossing entry, its neighbours, and the weights utils.py:72 needs
        uint32_t r_e = 0u, r_s = NONE, r_p = NONE, Ce = 0u, We = 0u, Ws = 0u, Wp = 0u;
        if (nh <= static_cast<uint32_t>(p.hits_cap)) {
            // ---- sort the hits by rank (key = rank << 8 | slot), running sum of the weights in that order
            uint32_t key[8], wv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t v = static_cast<uint32_t>(lane * 8 + i);
                key[i] = v < nh ? ((hit_r[v] << 8) | v) : NONE;
            }
            warp_bitonic_sort<8>(key, lane);
            uint32_t run = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                wv[i] = key[i] != NONE ? hit_w[key[i] & 0xffu] : 0u;
                run += wv[i];
                wv[i] = run;
            }
            uint32_t lincl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, lincl, d);
                if (lane >= d) lincl += u;
            }
            __syncwarp();                               // every lane has read the unsorted list
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                hit_r[lane * 8 + i] = key[i] >> 8;      // 0xffffff in the unused slots: above every rank
                hit_w[lane * 8 + i] = wv[i] + lincl - run;
            }
            __syncwarp();
            if (lane < p.Q) {                           // scalar work of percentile `lane` on the sorted list
                const uint32_t lo_rank = k_bq << shift;
                int s = 0;                              // first slot with rank >= lo_rank
                for (int step = kSelect4Hits / 2; step > 0; step >>= 1)
                    if (s + step <= static_cast<int>(nh) && hit_r[s + step - 1] < lo_rank) s += step;
                const uint32_t cstart = s ? hit_w[s - 1] : 0u;
                const uint32_t need = k_tau - k_base + cstart;             // first slot whose running sum reaches it
                int i = 0;
                for (int step = kSelect4Hits / 2; step > 0; step >>= 1)
                    if (i + step <= static_cast<int>(nh) && hit_w[i + step - 1] < need) i += step;
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
        } else {
            // ---- too many hits for the list: narrow every percentile exactly, the whole warp on one at a time
            for (int q = 0; q < p.Q; ++q) {
                const uint32_t tau = __shfl_sync(0xffffffffu, k_tau, q);
                uint32_t lo = __shfl_sync(0xffffffffu, k_bq, q) << shift;   // ranks [lo, lo + width) hold the crossing
                uint32_t before = __shfl_sync(0xffffffffu, k_base, q);      // weight of the ranks below lo
                uint32_t wbits = shift;                                     // width = 2^wbits
                while (wbits > 0u) {
                    const uint32_t sbits = wbits > static_cast<uint32_t>(kSelect4LogNB) ? static_cast<uint32_t>(kSelect4LogNB) : wbits;
                    const uint32_t sshift = wbits - sbits;                  // sub-bucket = (rank - lo) >> sshift, 2^sbits of them
                    each_octet([&](bool ok, const uint4& a, const uint4& b) {
                        const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int e = 0; e < 7; ++e) {
                            const uint32_t d = (m[e] & 0xffffffu) - lo;
                            if (ok && (d >> wbits) == 0u) red_add_shared(hist_s + ((d >> sshift) << 2), (m[e] >> 24) * a.x);
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
                    lo += k_sub << sshift;
                    wbits = sshift;
                }
                // lo is the crossing rank: its neighbours, then the merged weights of all three
                uint32_t pr = 0u, sr = NONE, we = 0u;   // pr: rank + 1 of the largest rank below lo (0: none)
                each_octet([&](bool ok, const uint4& a, const uint4& b) {
                    const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 7; ++e) {
                        const uint32_t rank = m[e] & 0xffffffu, c = ok ? (m[e] >> 24) : 0u;
                        if (c != 0u && rank < lo) pr = max(pr, rank + 1u);
                        if (c != 0u && rank > lo) sr = min(sr, rank);
                        if (rank == lo) we += c * a.x;
                    }
                });
                pr = __reduce_max_sync(0xffffffffu, pr);
                sr = __reduce_min_sync(0xffffffffu, sr);
                we = __reduce_add_sync(0xffffffffu, we);
                uint32_t wp = 0u, ws = 0u;
                each_octet([&](bool ok, const uint4& a, const uint4& b) {
                    const uint32_t m[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 7; ++e) {
                        const uint32_t rank = m[e] & 0xffffffu, wgt = ok ? (m[e] >> 24) * a.x : 0u;
                        if (pr != 0u && rank == pr - 1u) wp += wgt;
                        if (rank == sr) ws += wgt;
                    }
                });
                wp = __reduce_add_sync(0xffffffffu, wp);
                ws = __reduce_add_sync(0xffffffffu, ws);
                if (lane == q) {
                    r_e = lo; r_p = pr ? pr - 1u : NONE; r_s = sr;
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
