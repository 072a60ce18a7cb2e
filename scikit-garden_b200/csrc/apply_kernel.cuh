// K1 -- forest leaf lookup ("apply") for sm_100a.
//
// Replaces BaseForest.apply (skgarden/forest.py:583-604) -> Tree._apply_dense
// (sklearn:tree/_tree.pyx:954-996): for every (row, tree) walk root -> leaf, going left
// iff x[feature] <= threshold.
//
// Mapping: a CTA owns ROWS = 32 * row_warps test rows, staged ONCE into shared memory
// as a transposed tile xt[f][row] (coalesced global reads; lane == row, so every
// shared-memory read of a feature is bank-conflict free whatever feature each lane
// needs).  The CTA's warps are split into row_warps x tree_warps; a warp walks its 32
// rows through the trees t = tree_warp, tree_warp + tree_warps, ...  Each lane keeps
// ILP independent root->leaf chains in flight (different trees) so the dependent 8-byte
// node loads (L1/L2 hits: one tree is a few MB, co-resident CTAs sweep the trees in the
// same order) overlap.  When a chain reaches a leaf the lane emits the result and
// starts its next tree at once, so lanes never wait for the deepest path of the warp.
//
// Node format: include/qforest_b200.h (8 bytes: floor-f32 threshold | offset<<fb | feature;
// leaf: member start | 0x80000000|count).  Left child = node + 1 (preorder).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qf {

struct ApplyParams {
    const uint2* __restrict__ nodes;       // packed nodes of all trees
    const uint32_t* __restrict__ roots;    // [T] packed index of each tree's root
    const float* __restrict__ X;           // (n, ldx) float32
    int64_t ldx;
    int64_t n;
    int T;
    int F;
    int fbits;
    int row_warps;                         // warps along rows per CTA (block = 256 threads)
};

enum { APPLY_EMIT_SEGMENT = 0, APPLY_EMIT_NODE = 1 };

constexpr int kApplyThreads = 256;

// out: EMIT_SEGMENT -> uint2 [n][T] (member start, member count)
//      EMIT_NODE    -> int32 [n][T] packed node index of the leaf
template <int ILP, int EMIT>
__global__ void __launch_bounds__(kApplyThreads)
apply_kernel(const ApplyParams p, void* __restrict__ out) {
    extern __shared__ float xt[];                       // [F][ROWS]
    const int rows_per_cta = 32 * p.row_warps;
    const int tree_warps = (kApplyThreads / 32) / p.row_warps;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;

    // stage the row tile transposed
    const int tile = rows_per_cta * p.F;
    for (int i = threadIdx.x; i < tile; i += kApplyThreads) {
        const int r = i / p.F;
        const int f = i - r * p.F;
        const int64_t row = row0 + r;
        xt[f * rows_per_cta + r] = (row < p.n) ? __ldg(p.X + row * p.ldx + f) : 0.0f;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rw = warp % p.row_warps;
    const int tw = warp / p.row_warps;
    const int r = rw * 32 + lane;
    const int64_t row = row0 + r;
    if (row >= p.n) return;                             // no further barriers below
    const float* __restrict__ xcol = xt + r;
    const uint32_t fmask = (1u << p.fbits) - 1u;

    uint32_t node[ILP];
    int tree[ILP];
    bool live[ILP];
    int next_tree = tw;
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
        tree[k] = next_tree;
        live[k] = next_tree < p.T;
        node[k] = live[k] ? __ldg(p.roots + next_tree) : 0u;
        next_tree += tree_warps;
    }

    bool any = false;
#pragma unroll
    for (int k = 0; k < ILP; ++k) any |= live[k];
    while (any) {
        uint2 nd[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k)
            if (live[k]) nd[k] = __ldg(p.nodes + node[k]);
        any = false;
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (!live[k]) continue;
            if (static_cast<int32_t>(nd[k].y) < 0) {    // leaf
                const int64_t o = row * p.T + tree[k];
                if (EMIT == APPLY_EMIT_SEGMENT)
                    static_cast<uint2*>(out)[o] = make_uint2(nd[k].x, nd[k].y & 0x7fffffffu);
                else
                    static_cast<int32_t*>(out)[o] = static_cast<int32_t>(node[k]);
                tree[k] = next_tree;
                live[k] = next_tree < p.T;
                if (live[k]) node[k] = __ldg(p.roots + next_tree);
                next_tree += tree_warps;
            } else {
                const uint32_t f = nd[k].y & fmask;
                const uint32_t off = nd[k].y >> p.fbits;
                const float x = xcol[f * rows_per_cta];
                // sklearn:tree/_tree.pyx:989  `X[i, feature] <= threshold` -> left
                node[k] += (x <= __uint_as_float(nd[k].x)) ? 1u : off;
            }
            any |= live[k];
        }
    }
}

// ---------------------------------------------------------------------------------------
// K1, shared-memory top-block variant (the default).
//
// ncu on the kernel above shows it bound by L1TEX wavefronts: every lane of a node load
// touches its own 128-byte line, so a warp-level load costs ~32 tag lookups and the SM
// retires about one (row, tree, level) visit per cycle.  Here the first `top_levels`
// levels of each tree come from a breadth-first "top block" (qf_build_top_blocks) that
// the CTA stages in shared memory with coalesced 16-byte loads, `trees_per_chunk` trees
// at a time: those levels cost shared-memory reads (no tags), all lanes in lock step, no
// divergence.  Only the deeper levels walk the depth-first array in global memory.
// Leaf results are staged in shared memory and written out as contiguous runs of
// trees_per_chunk entries per row (4 wavefronts per warp store instead of 32).
//
// CTA = `rows` threads, thread == row.  smem: xt[F][rows] | top[TC][2^L] | res[TC][rows+1]
// ---------------------------------------------------------------------------------------
struct ApplyTopParams {
    const uint2* __restrict__ nodes;
    const uint2* __restrict__ top;         // [T][2^top_levels]
    const float* __restrict__ X;
    int64_t ldx;
    int64_t n;
    int T;
    int F;
    int fbits;
    int top_levels;                        // L
    int chunk_log2;                        // trees per chunk TC = 1 << chunk_log2
};

template <int ILP, int EMIT>
__global__ void __launch_bounds__(256)
apply_top_kernel(const ApplyTopParams p, void* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_apply[];
    const int rows = blockDim.x;
    const int TC = 1 << p.chunk_log2;
    const int S = 1 << p.top_levels;
    float* xt = reinterpret_cast<float*>(smem_apply);                           // [F][rows]
    uint2* stop = reinterpret_cast<uint2*>(xt + static_cast<size_t>(p.F) * rows);  // [TC][S]
    uint2* sres = stop + static_cast<size_t>(TC) * S;                            // [TC][rows+1]
    const int tid = threadIdx.x;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * rows;
    const int64_t row = row0 + tid;
    const bool row_ok = row < p.n;

    const int tile = rows * p.F;
    for (int i = tid; i < tile; i += rows) {
        const int r = i / p.F;
        const int f = i - r * p.F;
        const int64_t rr = row0 + r;
        xt[f * rows + r] = (rr < p.n) ? __ldg(p.X + rr * p.ldx + f) : 0.0f;
    }
    const float* __restrict__ xcol = xt + tid;
    const uint32_t fmask = (1u << p.fbits) - 1u;

    for (int c0 = 0; c0 < p.T; c0 += TC) {
        const int tc = min(TC, p.T - c0);
        __syncthreads();                                // previous chunk fully written out / xt staged
        {   // stage the chunk's top blocks: tc * S contiguous 8-byte slots
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>(p.top + static_cast<size_t>(c0) * S);
            uint4* dst = reinterpret_cast<uint4*>(stop);
            const int n16 = tc * (S >> 1);
            for (int i = tid; i < n16; i += rows) dst[i] = __ldg(src + i);
        }
        __syncthreads();

        for (int g = 0; g < tc; g += ILP) {
            uint32_t node[ILP];
            bool live[ILP];
            {   // upper levels from shared memory, lock step
                uint32_t h[ILP];
#pragma unroll
                for (int k = 0; k < ILP; ++k) h[k] = 1u;
                for (int l = 0; l + 1 < p.top_levels; ++l) {
#pragma unroll
                    for (int k = 0; k < ILP; ++k) {
                        if (g + k < tc) {
                            const uint2 s = stop[(g + k) * S + h[k]];
                            if (static_cast<int32_t>(s.y) >= 0) {
                                const float x = xcol[s.y * rows];
                                h[k] = 2u * h[k] + ((x <= __uint_as_float(s.x)) ? 0u : 1u);
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < ILP; ++k) {
                    live[k] = row_ok && (g + k < tc);
                    node[k] = live[k] ? stop[(g + k) * S + h[k]].x : 0u;
                }
            }
            // deeper levels from the depth-first array
            bool any = false;
#pragma unroll
            for (int k = 0; k < ILP; ++k) any |= live[k];
            while (any) {
                uint2 nd[ILP];
#pragma unroll
                for (int k = 0; k < ILP; ++k)
                    if (live[k]) nd[k] = __ldg(p.nodes + node[k]);
                any = false;
#pragma unroll
                for (int k = 0; k < ILP; ++k) {
                    if (!live[k]) continue;
                    if (static_cast<int32_t>(nd[k].y) < 0) {
                        sres[(g + k) * (rows + 1) + tid] =
                            (EMIT == APPLY_EMIT_SEGMENT) ? make_uint2(nd[k].x, nd[k].y & 0x7fffffffu)
                                                         : make_uint2(node[k], 0u);
                        live[k] = false;
                    } else {
                        const uint32_t f = nd[k].y & fmask;
                        const uint32_t off = nd[k].y >> p.fbits;
                        const float x = xcol[f * rows];
                        node[k] += (x <= __uint_as_float(nd[k].x)) ? 1u : off;   // _tree.pyx:989
                        any = true;
                    }
                }
            }
        }
        __syncthreads();
        // write the chunk: for every row a contiguous run of tc entries
        const int total = rows << p.chunk_log2;
        for (int i = tid; i < total; i += rows) {
            const int r = i >> p.chunk_log2;
            const int j = i & (TC - 1);
            const int64_t rr = row0 + r;
            if (j < tc && rr < p.n) {
                const uint2 v = sres[j * (rows + 1) + r];
                const int64_t o = rr * p.T + c0 + j;
                if (EMIT == APPLY_EMIT_SEGMENT) static_cast<uint2*>(out)[o] = v;
                else static_cast<int32_t*>(out)[o] = static_cast<int32_t>(v.x);
            }
        }
    }
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

}  // namespace qf
