// K1 -- forest leaf lookup ("apply") for sm_100a.
//
// Replaces BaseForest.apply (skgarden/forest.py:583-604) -> Tree._apply_dense
// (sklearn:tree/_tree.pyx:954-996): for every (row, tree) walk root -> leaf, going left
// iff x[feature] <= threshold.
//
// What bounds this kernel is not HBM but the L1TEX tag stage: a warp-level node load costs
// one pass per DISTINCT 32-byte sector its 32 lanes touch, ~1 per cycle per SM (ncu: l1tex
// throughput 85 % of peak with sectors + shared wavefronts = 0.9 per SM cycle).  With test
// rows in arrival order every lane is in a different part of the tree after a few levels
// (32 sectors per load).  Three measures bring that down (tools/sim_wavefronts.py counts
// them on real forests):
//
//   1. locality order: rows are first bucketed by the leaf they reach in ONE tree (its
//      left-to-right leaf order is a space-filling order adapted to the data), and the
//      kernel walks the rows in that order.  Neighbouring lanes then follow the same path
//      through the upper ~10 levels of EVERY tree and stay close far below that.
//   2. lock step: all lanes of a warp walk the same tree at the same time (a chain is
//      re-armed with the warp's next tree only when every lane has reached its leaf), so
//      that the coherence of (1) is not lost to lanes drifting into different trees.
//   3. level-order nodes with sibling pairs: lanes that are close in feature space sit on
//      neighbouring nodes of a level, and a level is contiguous in memory, so k distinct
//      nodes cost about k / 4 sectors; the two children of a node share a sector, so a
//      warp that splits at a node does not pay for it on the next level (simulated on the
//      C2 forest: 5.8 -> 4.3 sectors and 5.4 -> 2.8 lines per warp load against depth-first
//      preorder).
//
// Mapping: a CTA owns ROWS = 32 * row_warps consecutive rows of the locality order, staged
// ONCE into shared memory as a transposed tile xt[f][row] (lane == row, so every
// shared-memory read of a feature is bank-conflict free whatever feature each lane
// needs).  The CTA's 8 warps are split into row_warps x tree_warps; a warp walks its 32
// rows through the trees t = tree_warp, tree_warp + tree_warps, ...  keeping ILP
// independent chains (different trees) in flight per lane so that the dependent 8-byte
// node loads overlap.
//
// Node format: include/qforest_b200.h (8 bytes: floor-f32 threshold | offset<<fb | feature;
// leaf: member start | 0x80000000|count).  Children = node + offset (left), + offset + 1 (right).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qf {

struct ApplyParams {
    const uint2* __restrict__ nodes;       // packed nodes of all trees
    const uint32_t* __restrict__ roots;    // [T] packed index of each tree's root
    const float* __restrict__ X;           // (n, ldx) float32
    const int* __restrict__ order;         // [n] position in the locality order -> row of X; nullptr = identity
    int64_t ldx;
    int64_t n;
    int T;                                 // row stride of the output (trees of the forest)
    int t_begin, t_end;                    // trees walked by this launch
    int F;
    int fbits;
    int row_warps;                         // warps along rows per CTA (block = 256 threads)
    int64_t ld_out;                        // APPLY_EMIT_COMPACT: out[tree * ld_out + pos * ld_row] -- tree-major (ld_row = 1,
    int64_t ld_row;                        //   ld_out >= rows) for the CTA-per-row kernels, row-major (ld_out = 1) for select4
};

// what a finished walk writes
//   SEGMENT: uint2 out[pos][T]   (member start, member count), pos = position in the locality order
//   NODE   : int32 out[row][T]   packed node index of the leaf, row = row of X
//   COMPACT: uint32 out[tree * ld_out + pos * ld_row]  lo32 of the leaf node (the compact list word when
//            the walk runs on the compact copy of the nodes, compact_lists.cuh).  TREE-major (ld_row = 1):
//            lane == row, so a warp writes 128 contiguous bytes per tree instead of 32 scattered sectors
enum { APPLY_EMIT_SEGMENT = 0, APPLY_EMIT_NODE = 1, APPLY_EMIT_COMPACT = 2 };

constexpr int kApplyThreads = 256;
constexpr int kApplyLevels = 2;                         // visits per chain between two warp votes (default)

template <int ILP, int EMIT, int LEVELS = kApplyLevels>
__global__ void __launch_bounds__(kApplyThreads, ILP <= 4 ? 8 : 1)
apply_kernel(const ApplyParams p, void* __restrict__ out) {
    extern __shared__ float xt[];                       // [F][ROWS]
    __shared__ uint64_t s_nodes;
    __shared__ uint32_t s_fbits;
    if (threadIdx.x == 0) {
        s_nodes = reinterpret_cast<uint64_t>(p.nodes);
        s_fbits = static_cast<uint32_t>(p.fbits);
    }
    const int rows_per_cta = 32 * p.row_warps;
    const int tree_warps = (kApplyThreads / 32) / p.row_warps;
    const int64_t pos0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;

    // stage the row tile transposed (threads sweep a row's features: coalesced per row).  Rows
    // whose length and stride are multiples of four (two) floats go through 128-bit (64-bit)
    // loads: 80-, 200- and 256-byte rows of configs 2, 3 and 5
    const bool al16 = ((p.F | p.ldx) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0;
    const bool al8 = ((p.F | p.ldx) & 1) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 7) == 0;
    const int vec = al16 ? 4 : (al8 ? 2 : 1);           // floats per load (block-uniform)
    const int fv = p.F / vec;
    const int tile = rows_per_cta * fv;
    for (int i = threadIdx.x; i < tile; i += kApplyThreads) {
        const int r = i / fv;
        const int f = (i - r * fv) * vec;
        const int64_t pos = pos0 + r;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (pos < p.n) {
            const int64_t row = p.order ? static_cast<int64_t>(__ldg(p.order + pos)) : pos;
            const float* __restrict__ src = p.X + row * p.ldx + f;
            if (vec == 4) {
                v = __ldg(reinterpret_cast<const float4*>(src));
            } else if (vec == 2) {
                const float2 u = __ldg(reinterpret_cast<const float2*>(src));
                v.x = u.x; v.y = u.y;
            } else {
                v.x = __ldg(src);
            }
        }
        xt[f * rows_per_cta + r] = v.x;
        if (vec >= 2) xt[(f + 1) * rows_per_cta + r] = v.y;
        if (vec == 4) { xt[(f + 2) * rows_per_cta + r] = v.z; xt[(f + 3) * rows_per_cta + r] = v.w; }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rw = warp % p.row_warps;
    const int tw = warp / p.row_warps;
    const int r = rw * 32 + lane;
    const int64_t pos = pos0 + r;
    const bool row_ok = pos < p.n;                      // lanes past the end walk a zero row and write nothing
    int64_t out_row = pos;
    if (EMIT == APPLY_EMIT_NODE && row_ok && p.order) out_row = __ldg(p.order + pos);
    // shared-memory byte address of this lane's column of the tile; feature f is at
    // xcol_s + f * 4 * rows_per_cta (32-bit address arithmetic, ld.shared)
    const uint32_t xcol_s = static_cast<uint32_t>(__cvta_generic_to_shared(xt + r));
    const uint32_t fstride = static_cast<uint32_t>(rows_per_cta) * 4u;
    const uint32_t fmask = (1u << p.fbits) - 1u;

    // chain k of every lane walks the same tree (warp-uniform tree[k]).  Per visit and chain:
    // one predicated 8-byte node load, one predicated shared-memory load, compare, select, add.
    // The sign bit of nd[k].y is the chain's state: a lane that reached its leaf keeps the leaf
    // node in nd[k] and its loads are predicated off.  Every LEVELS visits the sign bits of
    // all chains are AND-reduced over the warp (one REDUX); a chain whose 32 lanes all sit on a
    // leaf writes its results and is re-armed with the warp's next tree.
    // fbits and the node base take a round trip through shared memory: that makes them opaque
    // to the compiler, which keeps them in registers instead of re-reading the constant bank
    // in front of every shift / address computation of the loop below.
    const uint32_t fbits = s_fbits;
    const uint2* __restrict__ nodes = reinterpret_cast<const uint2*>(s_nodes);
    uint32_t node[ILP];
    uint2 nd[ILP];
    int tree[ILP];
    uint32_t armed = 0;                                 // bit k: chain k holds a tree (warp-uniform)
    int next_tree = p.t_begin + tw;
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
        tree[k] = next_tree;
        node[k] = 0u;
        nd[k] = make_uint2(0u, 0x80000000u);
        if (next_tree < p.t_end) {
            armed |= 1u << k;
            node[k] = __ldg(p.roots + next_tree);
            nd[k].y = 0u;
        }
        next_tree += tree_warps;
    }

    while (armed) {
#pragma unroll
        for (int lvl = 0; lvl < LEVELS; ++lvl) {
#pragma unroll
            for (int k = 0; k < ILP; ++k)
                if (static_cast<int32_t>(nd[k].y) >= 0) nd[k] = __ldg(nodes + node[k]);
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (static_cast<int32_t>(nd[k].y) >= 0) {
                    float x;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(xcol_s + (nd[k].y & fmask) * fstride));
                    // sklearn:tree/_tree.pyx:989  `X[i, feature] <= threshold` -> left, else the
                    // sibling right behind it.  set.gtu gives 0xffffffff for "not <=": one
                    // three-input add takes the step (node + offset - (-1))
                    uint32_t right;
                    asm("set.gtu.u32.f32 %0, %1, %2;" : "=r"(right) : "f"(x), "f"(__uint_as_float(nd[k].x)));
                    node[k] = node[k] + (nd[k].y >> fbits) - right;
                }
            }
        }
        uint32_t signs = 0;                             // bit k: this lane's chain k sits on a leaf
#pragma unroll
        for (int k = ILP - 1; k >= 0; --k) signs = __funnelshift_l(nd[k].y, signs, 1);
        const uint32_t done = __reduce_and_sync(0xffffffffu, signs) & armed;
        if (done) {                                     // warp-uniform
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (!(done & (1u << k))) continue;
                if (row_ok) {
                    if (EMIT == APPLY_EMIT_SEGMENT)
                        static_cast<uint2*>(out)[out_row * p.T + tree[k]] =
                            make_uint2(nd[k].x, nd[k].y & 0x7fffffffu);
                    else if (EMIT == APPLY_EMIT_COMPACT)
                        static_cast<uint32_t*>(out)[static_cast<int64_t>(tree[k]) * p.ld_out + out_row * p.ld_row] = nd[k].x;
                    else
                        static_cast<int32_t*>(out)[out_row * p.T + tree[k]] = static_cast<int32_t>(node[k]);
                }
                if (next_tree < p.t_end) {
                    tree[k] = next_tree;
                    node[k] = __ldg(p.roots + next_tree);
                    nd[k].y = 0u;
                } else {
                    armed &= ~(1u << k);
                }
                next_tree += tree_warps;
            }
        }
    }
}

// Leaf lookup for rows too wide for the shared-memory tile (more than ~1600 features): one
// thread per row, features read from global memory (L1/L2-resident after the first touch of a
// line), trees one after the other.  Same outputs as apply_kernel; a correctness path, not a
// tuned one.
template <int EMIT>
__global__ void __launch_bounds__(128)
apply_wide_rows_kernel(const ApplyParams p, void* __restrict__ out) {
    const int64_t pos = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pos >= p.n) return;
    const int64_t row = p.order ? static_cast<int64_t>(__ldg(p.order + pos)) : pos;
    const float* __restrict__ x = p.X + row * p.ldx;
    const uint32_t fmask = (1u << p.fbits) - 1u;
    for (int t = p.t_begin; t < p.t_end; ++t) {
        uint32_t node = __ldg(p.roots + t);
        uint2 nd;
        for (;;) {
            nd = __ldg(p.nodes + node);
            if (static_cast<int32_t>(nd.y) < 0) break;
            node += (nd.y >> p.fbits) + ((__ldg(x + (nd.y & fmask)) <= __uint_as_float(nd.x)) ? 0u : 1u);   // _tree.pyx:989
        }
        if (EMIT == APPLY_EMIT_SEGMENT)
            static_cast<uint2*>(out)[pos * p.T + t] = make_uint2(nd.x, nd.y & 0x7fffffffu);
        else if (EMIT == APPLY_EMIT_COMPACT)
            static_cast<uint32_t*>(out)[static_cast<int64_t>(t) * p.ld_out + pos * p.ld_row] = nd.x;
        else
            static_cast<int32_t*>(out)[row * p.T + t] = static_cast<int32_t>(node);
    }
}

// ---------------------------------------------------------------------------------------
// Locality order: counting sort of the rows by the leaf they reach in ONE tree.
//   locality_key_kernel   : bucket[row] = (first member of the leaf) >> key_shift, histogram
//   exclusive_scan_kernel : single CTA, in-place exclusive prefix sum of the kKeyBins counters
//   order_scatter_kernel  : order[atomicAdd(&cursor[bucket[row]], 1)] = row
// The order inside a bucket is whatever the atomics give; results do not depend on it
// (every row is computed independently and written to its own output row).
// ---------------------------------------------------------------------------------------
constexpr int kKeyBins = 8192;                         // default number of histogram counters
constexpr int kKeyBinsMax = 1 << 18;                   // option "key_bins": up to this many (multiples of 8192)
constexpr int kScanThreads = 1024;

// lane == row; x is read straight from global memory (a row is one or two lines, L1-resident
// after the first touch), so there is no tile staging and no barrier for a single tree
__global__ void __launch_bounds__(128)
locality_key_kernel(const uint2* __restrict__ nodes, uint32_t root, const float* __restrict__ X,
                    int64_t ldx, int64_t n, int fbits, int key_shift, int key_bins, int* __restrict__ bucket,
                    int* __restrict__ hist) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const float* __restrict__ x = X + row * ldx;
    const uint32_t fmask = (1u << fbits) - 1u;
    uint32_t node = root;
    uint2 nd;
    for (;;) {
        nd = __ldg(nodes + node);
        if (static_cast<int32_t>(nd.y) < 0) break;
        node += (nd.y >> fbits) + ((__ldg(x + (nd.y & fmask)) <= __uint_as_float(nd.x)) ? 0u : 1u);   // _tree.pyx:989
    }
    // lo32 of a leaf of the FIRST tree: index of its first member, which grows from left to
    // right over the leaves (before qf_build_csr: the leaf's left-to-right ordinal)
    const int b = static_cast<int>(min(nd.x >> key_shift, static_cast<uint32_t>(key_bins - 1)));
    bucket[row] = b;
    atomicAdd(hist + b, 1);
}

// single CTA: in-place exclusive prefix sum of `nbins` counters (a multiple of 4 * kScanThreads),
// nbins / kScanThreads consecutive counters per thread
__global__ void __launch_bounds__(kScanThreads)
exclusive_scan_kernel(int* __restrict__ hist, int nbins) {
    __shared__ int warp_tot[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = nbins / kScanThreads;               // multiple of 4
    int4* __restrict__ mine = reinterpret_cast<int4*>(hist + tid * per);
    int sum = 0;
    for (int k = 0; k < per / 4; ++k) {
        const int4 a = mine[k];
        sum += a.x + a.y + a.z + a.w;
    }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int v = warp_tot[lane];
        int iv = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, iv, d);
            if (lane >= d) iv += u;
        }
        warp_tot[lane] = iv - v;                       // exclusive offset of each warp
    }
    __syncthreads();
    int run = warp_tot[warp] + incl - sum;
    for (int k = 0; k < per / 4; ++k) {
        const int4 a = mine[k];
        int4 o;
        o.x = run; run += a.x; o.y = run; run += a.y; o.z = run; run += a.z; o.w = run; run += a.w;
        mine[k] = o;
    }
}

__global__ void order_scatter_kernel(const int* __restrict__ bucket, int64_t n, int* __restrict__ cursor,
                                     int* __restrict__ order) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= n) return;
    order[atomicAdd(cursor + bucket[row], 1)] = static_cast<int>(row);
}

// EMIT_NODE output -> scikit-learn node ids (int64), what BaseForest.apply returns
// (forest.py:604).  node_ids == nullptr means packed position - root == sklearn id.
__global__ void node_to_sklearn_id_kernel(const int32_t* __restrict__ node_idx, int64_t total,
                                          int T, const uint32_t* __restrict__ roots,
                                          const int32_t* __restrict__ node_ids,
                                          int64_t* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += stride) {
        const int32_t idx = node_idx[i];
        const int t = static_cast<int>(i % T);
        out[i] = node_ids ? static_cast<int64_t>(__ldg(node_ids + idx))
                          : static_cast<int64_t>(idx) - static_cast<int64_t>(__ldg(roots + t));
    }
}

// K3 -- mean path.  ForestRegressor.predict (forest.py:1093-1131): float64 sum of
// tree_.value[leaf] in increasing tree order (n_jobs=1 accumulation), then / T.
__global__ void mean_kernel(const int32_t* __restrict__ node_idx, int64_t n, int T,
                            const double* __restrict__ values, double* __restrict__ out) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int32_t* __restrict__ idx = node_idx + row * T;
    double acc = 0.0;                                   // forest.py:1120
    for (int t = 0; t < T; ++t) acc = __dadd_rn(acc, __ldg(values + idx[t]));   // :806-818
    out[row] = __ddiv_rn(acc, static_cast<double>(T));  // forest.py:1129
}

// Mean and standard deviation of y | x -- skgarden/forest.py:133-178 (`_return_std`) on top of
// ForestRegressor.predict (forest.py:332-362, :516-548).  Per row, float64, trees in
// increasing order like the reference's loop (forest.py:165-174):
//   s   += max(impurity[leaf], min_variance) + value[leaf]^2
//   std  = sqrt(max(0, s / T - mean^2)),   mean = (sum of value[leaf]) / T
__global__ void mean_std_kernel(const int32_t* __restrict__ node_idx, int64_t n, int T,
                                const double* __restrict__ values, const double* __restrict__ impurity,
                                double min_variance, double* __restrict__ mean_out, double* __restrict__ std_out) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int32_t* __restrict__ idx = node_idx + row * T;
    double acc = 0.0, s = 0.0;
    for (int t = 0; t < T; ++t) {
        const double m = __ldg(values + idx[t]);
        double v = __ldg(impurity + idx[t]);
        if (v < min_variance) v = min_variance;          // forest.py:170
        acc = __dadd_rn(acc, m);
        s = __dadd_rn(s, __dadd_rn(v, __dmul_rn(m, m)));   // forest.py:172
    }
    const double mean = __ddiv_rn(acc, static_cast<double>(T));
    s = __ddiv_rn(s, static_cast<double>(T));           // forest.py:174
    s = __dsub_rn(s, __dmul_rn(mean, mean));             // forest.py:175
    if (s < 0.0) s = 0.0;                                // forest.py:176
    mean_out[row] = mean;
    std_out[row] = __dsqrt_rn(s);                        // forest.py:177
}

}  // namespace qf
