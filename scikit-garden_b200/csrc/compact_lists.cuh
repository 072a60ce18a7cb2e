// Compact leaf lists for the fused weight + percentile kernel of QF_MODE_FAST (select3_kernel.cuh).
//
// The member array of the forest (include/qforest_b200.h) holds one 8-byte (rank, float32 weight)
// entry per in-bag sample, a leaf's list starting at an arbitrary 8-byte offset: a row of config 3
// reads ~3300 of them (26 KB) out of ~500 lists of 53 bytes that straddle 32-byte sectors.  The
// weight of a member is f32(count / leaf count sum) (ensemble.py:94-99), i.e. all weights of a leaf
// share ONE denominator.  The compact form stores that denominator once per 32 bytes:
//
//   octet (32 bytes, 32-byte aligned) = [ U | m0 | m1 | m2 | m3 | m4 | m5 | m6 ]
//       U  = rn(fx_scale / s), s = leaf count sum: the 32-bit fixed-point weight of ONE bootstrap count
//            of this leaf (fx_scale = 2^(31-k) units per unit of weight, 2^k > the largest total weight
//            a row can gather), so that a member weighs count * U -- one integer multiply
//       m  = rank (bits 0..23) | bootstrap count (bits 24..31); padding = count 0 with an arbitrary rank
//   a leaf of c members owns ceil(c / 7) consecutive octets, every one with the header, so that an
//   octet can be processed without its neighbours (any lane can take any octet).
//   The builder recovers (s, counts) from the stored float32 weights and verifies
//   __fdiv_rn(count, s) == weight bit for bit for every member; a forest for which that fails is
//   "not compactable" and the 8-byte kernels stay in charge.
//
// The walk kernel runs on a COPY of the node array whose leaves carry, in lo32, the compact list
// word `first octet | (octets - 1) << 28` (0xffffffff: a leaf without in-bag members); K1 emits
// that word (APPLY_EMIT_COMPACT) and K2 needs nothing else to find a leaf's members.  A leaf of more
// than 15 octets (105 members) is a LONG list: the 4-bit field holds 15 and `first octet` is an info
// octet whose first word is the number of member octets that follow it (K2 reads such a list straight
// from global memory; rare, and forests made of such leaves take the exact kernels anyway).
// A row of config 3 then reads ~650 octets = 21 KB in whole sectors instead of 26 KB + straddle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qf {

constexpr uint32_t kCompactEmpty = 0xffffffffu;
constexpr int kOctetMembers = 7;
constexpr int kCompactMaxOctets = 15;                   // (octets - 1) lives in 4 bits; 15 there = long list (below)
constexpr uint32_t kCompactLong = 15u;
constexpr uint32_t kCompactOffMask = 0x0fffffffu;       // first octet: 28 bits (8 GB of lists)
constexpr int kCompactScanTile = 4096;                  // items per scan block (256 threads x 16)

// Leaf count sum s and per-member counts from the float32 weights alone: the smallest
// s = rint(k / w_min), k = 1, 2, ..., for which every weight w of the leaf is exactly
// f32(c / s) with an integer count 1 <= c <= 255.  0 = no such s (forest not compactable).
// (A common factor of the true counts is divided out -- 2/4 and 1/2 are the same float.)
__device__ inline uint32_t leaf_count_sum(const uint2* __restrict__ m, uint32_t cnt) {
    float wmin = __uint_as_float(m[0].y);
    for (uint32_t i = 1; i < cnt; ++i) wmin = fminf(wmin, __uint_as_float(m[i].y));
    if (!(wmin > 0.0f)) return 0u;
    for (int k = 1; k <= 64; ++k) {
        const float s = rintf(__fdiv_rn(static_cast<float>(k), wmin));
        if (!(s >= 1.0f && s <= 8388608.0f)) continue;
        bool ok = true;
        for (uint32_t i = 0; i < cnt && ok; ++i) {
            const float w = __uint_as_float(m[i].y);
            const float c = rintf(__fmul_rn(w, s));
            ok = c >= 1.0f && c <= 255.0f && __fdiv_rn(c, s) == w;
        }
        if (ok) return static_cast<uint32_t>(s);
    }
    return 0u;
}

// pass 1: octets of every leaf, keyed by the leaf's first member (unique per non-empty leaf);
// flags: 2 = weights are not count / sum, 4 = rank >= 2^24
__global__ void __launch_bounds__(256)
compact_mark_kernel(const uint2* __restrict__ nodes, int64_t n_nodes, const uint2* __restrict__ members,
                    uint32_t* __restrict__ noct_at, int* __restrict__ flags, int align2) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const uint2 nd = nodes[i];
    if (!(nd.y & 0x80000000u)) return;
    const uint32_t cnt = nd.y & 0x7fffffffu;
    if (cnt == 0u) return;
    if (leaf_count_sum(members + nd.x, cnt) == 0u) { atomicOr(flags, 2); return; }
    if (members[nd.x + cnt - 1].x >= (1u << 24)) { atomicOr(flags, 4); return; }   // lists are sorted by rank
    const uint32_t noct = (cnt + kOctetMembers - 1) / kOctetMembers;
    uint32_t alloc = noct > static_cast<uint32_t>(kCompactMaxOctets) ? noct + 1u : noct;   // long list: + the info octet
    // align2: a list of two or more octets starts on an even octet (a 64-byte DRAM burst then holds both octets
    // of a two-octet leaf instead of half of it and half of a stranger); one spare octet pays for the shift
    if (align2 && alloc >= 2u) ++alloc;
    noct_at[nd.x] = alloc;
}

// in-place exclusive scan of a uint32 array in three launches: tile sums, scan of the tile
// sums (one block), tile-local scan + tile offset
__global__ void __launch_bounds__(256)
scan_tile_sums_kernel(const uint32_t* __restrict__ in, int64_t n, uint64_t* __restrict__ tile_sum) {
    __shared__ uint64_t red[8];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kCompactScanTile;
    uint64_t s = 0;
    for (int k = 0; k < kCompactScanTile / 256; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < n) s += in[i];
    }
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) s += red[w];
        tile_sum[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024)
scan_tile_offsets_kernel(uint64_t* __restrict__ tile_sum, int64_t n_tiles, uint64_t* __restrict__ total_out) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n_tiles; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_sum[i] : 0;
        uint64_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint64_t w = warp_tot[lane];
            uint64_t wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t u = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += u;
            }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const uint64_t excl = carry + warp_tot[warp] + incl - v;
        if (i < n_tiles) tile_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(256)
scan_apply_kernel(uint32_t* __restrict__ data, int64_t n, const uint64_t* __restrict__ tile_off) {
    __shared__ uint32_t warp_tot[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kCompactScanTile + threadIdx.x * 16;
    uint32_t v[16], s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0u;
        s += v[k];
    }
    uint32_t incl = s;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t run = static_cast<uint32_t>(tile_off[blockIdx.x]) + incl - s;
    for (int w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

// pass 2: the octets themselves and the compact copy of the nodes (oct_at = exclusive scan of
// pass 1's octet counts: first octet of the leaf whose first member is nd.x)
__global__ void __launch_bounds__(256)
compact_fill_kernel(const uint2* __restrict__ nodes, int64_t n_nodes, const uint2* __restrict__ members,
                    const uint32_t* __restrict__ oct_at, float fx_scale, uint32_t n_train,
                    uint2* __restrict__ nodes_c, uint4* __restrict__ lists, int align2) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint2 nd = nodes[i];
    if (nd.y & 0x80000000u) {
        const uint32_t cnt = nd.y & 0x7fffffffu;
        if (cnt == 0u) {
            nd.x = kCompactEmpty;
        } else {
            const uint2* __restrict__ m = members + nd.x;
            const uint32_t s = leaf_count_sum(m, cnt);
            const float sf = static_cast<float>(s);
            uint32_t first = oct_at[nd.x];
            const uint32_t noct = (cnt + kOctetMembers - 1) / kOctetMembers;
            const bool is_long = noct > static_cast<uint32_t>(kCompactMaxOctets);
            if (align2 && noct >= 2u) first += first & 1u;          // (compact_mark_kernel reserved the spare octet)
            const uint32_t word = first | ((is_long ? kCompactLong : noct - 1u) << 28);
            if (is_long) {                              // info octet, then the member octets
                lists[2 * static_cast<size_t>(first)] = make_uint4(noct, 0u, 0u, 0u);
                lists[2 * static_cast<size_t>(first) + 1] = make_uint4(0u, 0u, 0u, 0u);
                ++first;
            }
            for (uint32_t j = 0; j < noct; ++j) {
                uint32_t w[8];
                w[0] = __float2uint_rn(__fdiv_rn(fx_scale, sf));    // fixed-point weight of one bootstrap count
#pragma unroll
                for (int e = 0; e < kOctetMembers; ++e) {
                    const uint32_t k = j * kOctetMembers + e;
                    // padding: count 0 (weight zero) and a pseudo-random rank below n_train, so that the slots
                    // of a warp's 32 octets do not all add their zero to the same histogram bucket
                    w[e + 1] = static_cast<uint32_t>((static_cast<uint64_t>((first + j) * 7u + e) * 2654435761ull >> 8) % n_train);
                    if (k < cnt) {
                        const uint2 x = m[k];
                        const uint32_t c = static_cast<uint32_t>(rintf(__fmul_rn(__uint_as_float(x.y), sf)));
                        w[e + 1] = x.x | (c << 24);
                    }
                }
                lists[2 * static_cast<size_t>(first + j)] = make_uint4(w[0], w[1], w[2], w[3]);
                lists[2 * static_cast<size_t>(first + j) + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            nd.x = word;
        }
    }
    nodes_c[i] = nd;
}

// per tree: sum over leaves of members * octets and of members -- their ratio is the number of
// octets a row expects to read from this tree (a row lands in a leaf with probability ~ its size)
__global__ void __launch_bounds__(256)
compact_tree_stats_kernel(const uint2* __restrict__ nodes, const uint32_t* __restrict__ roots, int64_t n_nodes, int T,
                          unsigned long long* __restrict__ out) {
    const int t = blockIdx.x;
    const int64_t n0 = roots[t], n1 = (t + 1 < T) ? static_cast<int64_t>(roots[t + 1]) : n_nodes;
    unsigned long long a = 0, b = 0;
    for (int64_t i = n0 + threadIdx.x; i < n1; i += blockDim.x) {
        const uint32_t hi = nodes[i].y;
        if (hi & 0x80000000u) {
            const unsigned long long cnt = hi & 0x7fffffffu;
            a += cnt * ((cnt + kOctetMembers - 1) / kOctetMembers);
            b += cnt;
        }
    }
    atomicAdd(out + 2 * t, a);
    atomicAdd(out + 2 * t + 1, b);
}

}  // namespace qf
