// K4 -- fit-side bookkeeping on the device: the leaf -> in-bag member lists (CSR) that replace
// the reference's dense (T,N) y_train_leaves_ / y_weights_ arrays.
//
// Replaces skgarden/quantile/ensemble.py:83-101 (an O(T*L*N) Python loop in the reference) and
// tree.py:98-101 (tree_.apply(X_train) per tree on the CPU); produces bit for bit what the host
// flattener qf_pack_tree produces:
//   leaf lookup of the training rows     K1 (apply_kernel, EMIT_NODE) on X_train
//   count_kernel                         per leaf: in-bag members and sum of their multiplicities
//   leaf_order_kernel                    counts permuted into left-to-right leaf order (the
//                                        ordinal a structure-only qf_pack_tree left in lo32)
//   exclusive scan over the ordinals     leaf -> first member index (lists grouped by leaf,
//                                        leaves from left to right, trees back to back)
//   leaf_record_kernel                   leaves get (start, 0x80000000 | count)
//   scatter_kernel                       member = (rank in argsort(y_train_), f32(count / leaf sum)),
//                                        ensemble.py:94-99: int / int -> float64 -> float32
//   sort_small_kernel / sort_big_kernel  every list ordered by rank (ranks are unique inside a
//                                        list): insertion sort per thread up to 32 members,
//                                        shared-memory bitonic sort per CTA up to 4096
// Lists longer than 4096 members (very shallow trees) are left to the host flattener
// (QF_ERR_LIMIT).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qf {

constexpr int kCsrSmallLeaf = 32;
constexpr int kCsrBigLeaf = 4096;
constexpr int kCsrScanBlock = 1024;                     // elements per block of the node scan

// thread per (sample j, tree t); idx[j][T] packed leaf node, counts[j][T] multiplicities (or null)
__global__ void csr_count_kernel(const int32_t* __restrict__ idx, const uint16_t* __restrict__ counts, int64_t total,
                                 uint32_t* __restrict__ leaf_cnt, uint32_t* __restrict__ leaf_sum) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t c = counts ? counts[i] : 1u;
        if (c) {
            const int32_t node = idx[i];
            atomicAdd(leaf_cnt + node, 1u);
            atomicAdd(leaf_sum + node, c);
        }
    }
}

// exclusive scan of `in` over n elements in three passes: per-block sums, scan of the sums
// (one CTA), add back
__global__ void __launch_bounds__(256)
csr_block_sum_kernel(const uint32_t* __restrict__ in, int64_t n, uint64_t* __restrict__ block_sum) {
    __shared__ uint64_t red[8];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kCsrScanBlock;
    uint64_t s = 0;
    for (int k = threadIdx.x; k < kCsrScanBlock; k += 256) {
        const int64_t i = base + k;
        if (i < n) s += in[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) s += red[w];
        block_sum[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024)
csr_scan_sums_kernel(uint64_t* __restrict__ block_sum, int64_t n_blocks, uint64_t* __restrict__ total_out) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const uint64_t v = (i < n_blocks) ? block_sum[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint64_t before = carry_s;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (i < n_blocks) block_sum[i] = before + incl - v;           // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

// thread per packed node: a leaf drops its member count at its left-to-right ordinal (lo32
// as written by a structure-only qf_pack_tree) and records where it lives
__global__ void csr_leaf_order_kernel(const uint2* __restrict__ nodes, int64_t n, const uint32_t* __restrict__ leaf_cnt,
                                      int64_t n_leaves, uint32_t* __restrict__ cnt_ord, int32_t* __restrict__ node_of,
                                      int* __restrict__ bad_ordinal) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 nd = nodes[i];
    if (!(nd.y & 0x80000000u)) return;
    if (nd.x >= n_leaves || (nd.y & 0x7fffffffu)) { atomicExch(bad_ordinal, 1); return; }
    if (atomicCAS(node_of + nd.x, -1, static_cast<int32_t>(i)) != -1) { atomicExch(bad_ordinal, 1); return; }   // two leaves, one ordinal
    cnt_ord[nd.x] = leaf_cnt[i];
}

// over the leaf ordinals: start = exclusive prefix of cnt_ord; the leaf at every ordinal gets
// its record.  One thread walks 4 consecutive ordinals; a block covers kCsrScanBlock of them.
__global__ void __launch_bounds__(256)
csr_leaf_record_kernel(uint2* __restrict__ nodes, const uint32_t* __restrict__ cnt_ord, const int32_t* __restrict__ node_of,
                       int64_t n_leaves, const uint64_t* __restrict__ block_sum, uint32_t* __restrict__ start) {
    __shared__ uint32_t warp_tot[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * kCsrScanBlock + threadIdx.x * 4;
    uint32_t c[4], mine = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k] = (i0 + k < n_leaves) ? cnt_ord[i0 + k] : 0u;
        mine += c[k];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint64_t run = block_sum[blockIdx.x] + incl - mine;
    for (int w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n_leaves) {
            const int32_t node = node_of[i0 + k];
            if (node >= 0) {                                  // every ordinal of a well-formed forest has its leaf
                start[node] = static_cast<uint32_t>(run);
                nodes[node] = make_uint2(static_cast<uint32_t>(run), 0x80000000u | c[k]);
            }
            run += c[k];
        }
    }
}

__global__ void csr_scatter_kernel(const int32_t* __restrict__ idx, const uint16_t* __restrict__ counts, int64_t total, int T,
                                   const int32_t* __restrict__ rank, const uint32_t* __restrict__ start,
                                   uint32_t* __restrict__ leaf_cnt, const uint32_t* __restrict__ leaf_sum,
                                   uint2* __restrict__ members) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t c = counts ? counts[i] : 1u;
        if (c) {
            const int32_t node = idx[i];
            const int64_t j = i / T;
            const uint32_t slot = start[node] + atomicSub(leaf_cnt + node, 1u) - 1u;   // any order: sorted below
            // ensemble.py:98-99: int64 / int64 -> float64, stored into the float32 y_weights_
            const float w = __double2float_rn(__ddiv_rn(static_cast<double>(c), static_cast<double>(leaf_sum[node])));
            members[slot] = make_uint2(static_cast<uint32_t>(rank[j]), __float_as_uint(w));
        }
    }
}

// thread per packed node: short lists sorted in place, longer ones appended to `big`
__global__ void csr_sort_small_kernel(const uint2* __restrict__ nodes, int64_t n, uint2* __restrict__ members,
                                      int32_t* __restrict__ big, int* __restrict__ n_big, int* __restrict__ too_big) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 nd = nodes[i];
    if (!(nd.y & 0x80000000u)) return;
    const uint32_t cnt = nd.y & 0x7fffffffu;
    if (cnt < 2) return;
    if (cnt > kCsrBigLeaf) { atomicExch(too_big, 1); return; }
    if (cnt > kCsrSmallLeaf) { big[atomicAdd(n_big, 1)] = static_cast<int32_t>(i); return; }
    uint2* m = members + nd.x;
    for (uint32_t a = 1; a < cnt; ++a) {                 // insertion sort by rank
        const uint2 key = m[a];
        uint32_t b = a;
        while (b > 0 && m[b - 1].x > key.x) { m[b] = m[b - 1]; --b; }
        m[b] = key;
    }
}

// CTA per long list: bitonic sort of (rank, weight) entries in shared memory
__global__ void __launch_bounds__(256)
csr_sort_big_kernel(const uint2* __restrict__ nodes, const int32_t* __restrict__ big, uint2* __restrict__ members) {
    __shared__ uint2 s[kCsrBigLeaf];
    const uint2 nd = nodes[big[blockIdx.x]];
    const uint32_t cnt = nd.y & 0x7fffffffu;
    uint32_t cap = 64;
    while (cap < cnt) cap <<= 1;
    uint2* m = members + nd.x;
    for (uint32_t k = threadIdx.x; k < cap; k += 256) s[k] = (k < cnt) ? m[k] : make_uint2(0xffffffffu, 0u);
    __syncthreads();
    for (uint32_t size = 2; size <= cap; size <<= 1) {
        for (uint32_t str = size >> 1; str > 0; str >>= 1) {
            for (uint32_t k = threadIdx.x; k < cap / 2; k += 256) {
                const uint32_t lo = 2 * k - (k & (str - 1));     // index with bit `str` clear
                const uint32_t hi = lo + str;
                const bool up = (lo & size) == 0;
                const uint2 a = s[lo], b = s[hi];
                if ((a.x > b.x) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t k = threadIdx.x; k < cnt; k += 256) m[k] = s[k];
}

// per tree: in-bag members, largest leaf, sum over leaves of members^2 (block t)
__global__ void __launch_bounds__(256)
csr_tree_stats_kernel(const uint2* __restrict__ nodes, const int64_t* __restrict__ tree_offsets,
                      unsigned long long* __restrict__ inbag, uint32_t* __restrict__ max_leaf, double* __restrict__ sum_sq) {
    __shared__ unsigned long long r_in[8];
    __shared__ uint32_t r_mx[8];
    __shared__ double r_sq[8];
    const int t = blockIdx.x;
    unsigned long long in = 0;
    uint32_t mx = 0;
    double sq = 0.0;
    for (int64_t i = tree_offsets[t] + threadIdx.x; i < tree_offsets[t + 1]; i += 256) {
        const uint32_t hi = nodes[i].y;
        if (hi & 0x80000000u) {
            const uint32_t c = hi & 0x7fffffffu;
            in += c;
            mx = max(mx, c);
            sq += static_cast<double>(c) * static_cast<double>(c);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        in += __shfl_down_sync(0xffffffffu, in, d);
        mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
        sq += __shfl_down_sync(0xffffffffu, sq, d);
    }
    if ((threadIdx.x & 31) == 0) { r_in[threadIdx.x >> 5] = in; r_mx[threadIdx.x >> 5] = mx; r_sq[threadIdx.x >> 5] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { in += r_in[w]; mx = max(mx, r_mx[w]); sq += r_sq[w]; }
        inbag[t] = in;
        max_leaf[t] = mx;
        sum_sq[t] = sq;
    }
}

}  // namespace qf
