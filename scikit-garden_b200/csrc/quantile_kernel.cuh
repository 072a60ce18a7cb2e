// K2 -- fused leaf co-membership weighting + weighted percentile for sm_100a.
//
// Replaces, per test row, the dense (T,N) mask + masked sum of
// skgarden/quantile/ensemble.py:137-140 and weighted_percentile
// (skgarden/quantile/utils.py:46-81) by a pass over only the non-zero weights:
//   gather the (rank, weight) members of the row's T leaves  ->  order by rank, merging
//   the contributions of a sample that several trees share  ->  running sum  ->  select /
//   interpolate every requested percentile.
//
// Arithmetic contract (QF_MODE_EXACT, bit-identical to the reference under numpy >= 2):
//   W_j  = ((0 + w_t1[j]) + w_t2[j]) + ...   float32, trees in increasing order  (ensemble.py:140)
//   C_k  = float32 running sum of W in rank order                                (utils.py:68)
//   p_k  = (100f / C_last) * (C_k - W_k / 2)                                     (utils.py:72)
//   i    = numpy searchsorted(p, q, 'left') with the compare done in float64     (utils.py:73)
//   out  = a_first / a_last at the ends, else a_s + ((float)q - p_s)/(p_{s+1} - p_s) * (a_{s+1} - a_s)
//          with separately rounded float32 operations                           (utils.py:74-81)
// All float32 operations use the _rn intrinsics so that nvcc never contracts them to FMA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qf {

constexpr int kMaxQuantilesPerLaunch = 128;

struct QuantileList {
    double q[kMaxQuantilesPerLaunch];
};

struct QuantileParams {
    const uint2* __restrict__ seg;        // [n][T] (member start, member count) from K1, rows in
                                          //   the locality order of K1
    const int* __restrict__ order;        // [n] locality position -> output row; nullptr = identity
    const uint2* __restrict__ members;    // (rank, float bits of weight)
    const float* __restrict__ y_sorted;   // [N]  y_train_.astype(f32)[sorter]
    double* __restrict__ out;             // [n][ldo], indexed by output row
    int64_t n;
    int64_t ldo;                          // row stride of out (total number of quantiles)
    int T;
    int Q;                                // quantiles in this launch (<= kMaxQuantilesPerLaunch)
    int cap;                              // smem capacity in members (power of two)
    const int* row_list;                  // nullptr: rows 0..n-1; else rows row_list[0..*row_count)
    const int* row_count;
    int* overflow_count;                  // rows whose gather exceeds this kernel's capacity
    int* overflow_rows;                   //   are appended here for the next, larger kernel
};

// p_k of utils.py:72 from (C_k, W_k); s = 100f / total.
__device__ __forceinline__ float plotting_position(float s, float C, float W) {
    return __fmul_rn(s, __fsub_rn(C, __fmul_rn(W, 0.5f)));
}

// utils.py:73-81 for one q.  rankOf(k), weightOf(k), cumOf(k) give entry k of the merged,
// rank-ordered list; m >= 1 entries.
template <typename RankOf, typename WeightOf, typename CumOf>
__device__ __forceinline__ double select_percentile(double q, int m, float s, RankOf rankOf,
                                                    WeightOf weightOf, CumOf cumOf,
                                                    const float* __restrict__ y_sorted) {
    // numpy's binary search, side='left', float32 element promoted to float64 for the compare.
    // p < q in float64  <=>  p < (smallest float32 >= q) in float32: no float64 instruction in
    // the loop (the FP64 pipe of sm_100 is narrow)
    const float qc = __double2float_ru(q);
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        const float pm = plotting_position(s, cumOf(mid), weightOf(mid));
        if (pm < qc) lo = mid + 1; else hi = mid;
    }
    const int start = lo - 1;
    if (start == m - 1) return static_cast<double>(__ldg(y_sorted + rankOf(m - 1)));   // utils.py:74-75
    if (start == -1) return static_cast<double>(__ldg(y_sorted + rankOf(0)));          // utils.py:76-77
    const float p0 = plotting_position(s, cumOf(start), weightOf(start));
    const float p1 = plotting_position(s, cumOf(start + 1), weightOf(start + 1));
    const float a0 = __ldg(y_sorted + rankOf(start));
    const float a1 = __ldg(y_sorted + rankOf(start + 1));
    const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(q), p0), __fsub_rn(p1, p0));  // utils.py:80
    return static_cast<double>(__fadd_rn(a0, __fmul_rn(frac, __fsub_rn(a1, a0))));          // utils.py:81
}

// ---------------------------------------------------------------------------------------
// Warp-per-row kernel for small member sets (P <= 32*E): the sort runs in registers.
//
//   gather (rank,weight) pairs in tree order into a per-warp shared-memory strip
//   -> 32-bit keys (rank << log2(CAP)) | position, E per lane, blocked layout
//   -> bitonic network in its all-ascending form (mirror step + xor steps): in-register
//      compare-exchanges for strides < E, one SHFL + one predicated min/max per key for
//      strides >= E
//   -> group heads add their duplicates in tree order (float32, exact), compaction by a
//      warp scan, lane 0 runs the float32 running sum over the compacted weights with
//      128-bit shared loads/stores, then one lane per requested percentile selects and
//      interpolates.
// Rows that do not fit are appended to the overflow list for the CTA kernel.
// ---------------------------------------------------------------------------------------
constexpr int kWarpKernelWarps = 8;

__device__ __forceinline__ void cmpswap(uint32_t& a, uint32_t& b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

// sorts the 32*E keys of a warp ascending; key v lives in lane v / E, register v % E
template <int E>
__device__ __forceinline__ void warp_bitonic_sort(uint32_t (&key)[E], int lane) {
#pragma unroll
    for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
        for (int i = 0; i < E; ++i)
            if ((i ^ (k - 1)) > i) cmpswap(key[i], key[i ^ (k - 1)]);
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < E; ++i)
                if ((i ^ j) > i) cmpswap(key[i], key[i ^ j]);
    }
#pragma unroll
    for (int k = 2 * E; k <= 32 * E; k <<= 1) {
        {   // mirror step: v <-> v ^ (k-1)
            const int lm = k / E - 1;
            const bool lower = (lane & ((k / E) >> 1)) == 0;
            uint32_t other[E];
#pragma unroll
            for (int i = 0; i < E; ++i) other[i] = __shfl_xor_sync(0xffffffffu, key[E - 1 - i], lm);
#pragma unroll
            for (int i = 0; i < E; ++i) key[i] = lower ? min(key[i], other[i]) : max(key[i], other[i]);
        }
#pragma unroll
        for (int j = k >> 2; j >= E; j >>= 1) {
            const int lx = j / E;
            const bool lower = (lane & lx) == 0;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const uint32_t o = __shfl_xor_sync(0xffffffffu, key[i], lx);
                key[i] = lower ? min(key[i], o) : max(key[i], o);
            }
        }
#pragma unroll
        for (int j = E >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < E; ++i)
                if ((i ^ j) > i) cmpswap(key[i], key[i ^ j]);
    }
}

template <int E, bool UNIT>
__global__ void __launch_bounds__(kWarpKernelWarps * 32)
quantile_warp_kernel(const QuantileParams p, const __grid_constant__ QuantileList ql) {
    constexpr int CAP = 32 * E;
    constexpr int POSBITS = (E == 4) ? 7 : ((E == 8) ? 8 : 6);
    constexpr uint32_t POSMASK = CAP - 1;
    static_assert(E == 2 || E == 4 || E == 8, "E must be 2, 4 or 8");
    __shared__ __align__(16) uint32_t sA[kWarpKernelWarps][CAP];   // keys -> sorted ranks -> compact ranks
    __shared__ __align__(16) float sB[kWarpKernelWarps][CAP];      // weights by position -> sorted -> compact
    __shared__ __align__(16) float sC[kWarpKernelWarps][CAP];      // running sum

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * kWarpKernelWarps + warp;
    if (row >= p.n) return;                              // warps are independent: no block barriers
    uint32_t* __restrict__ A = sA[warp];
    float* __restrict__ B = sB[warp];
    float* __restrict__ Cc = sC[warp];
    const uint2* __restrict__ seg = p.seg + row * p.T;

    // 1. gather in tree order
    uint32_t P = 0;
    if (UNIT) {                                          // every leaf holds exactly one member: position = tree
        P = static_cast<uint32_t>(p.T);
        if (P <= CAP) {
            for (int t = lane; t < p.T; t += 32) {
                const uint2 e = __ldg(p.members + seg[t].x);
                A[t] = (e.x << POSBITS) | static_cast<uint32_t>(t);
                B[t] = __uint_as_float(e.y);
            }
        }
    } else {
        for (int base = 0; base < p.T; base += 32) {
            const int t = base + lane;
            const uint2 s = (t < p.T) ? seg[t] : make_uint2(0u, 0u);
            uint32_t incl = s.y;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            const uint32_t excl = P + incl - s.y;
            if (excl + s.y <= CAP) {
                const uint2* __restrict__ src = p.members + s.x;
                for (uint32_t i = 0; i < s.y; ++i) {
                    const uint2 e = __ldg(src + i);
                    A[excl + i] = (e.x << POSBITS) | (excl + i);
                    B[excl + i] = __uint_as_float(e.y);
                }
            }
            P += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    if (P > CAP) {
        if (lane == 0) p.overflow_rows[atomicAdd(p.overflow_count, 1)] = static_cast<int>(row);
        return;
    }
    double* __restrict__ out = p.out + (p.order ? static_cast<int64_t>(__ldg(p.order + row)) : row) * p.ldo;
    __syncwarp();

    // 2. keys into registers (blocked: v = lane*E + i), padding sorts last
    uint32_t key[E];
    if constexpr (E >= 4) {
#pragma unroll
        for (int c = 0; c < E / 4; ++c) {
            const uint4 v = reinterpret_cast<const uint4*>(A)[lane * (E / 4) + c];
            key[4 * c + 0] = v.x; key[4 * c + 1] = v.y; key[4 * c + 2] = v.z; key[4 * c + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < E; ++i) key[i] = A[lane * E + i];
    }
#pragma unroll
    for (int i = 0; i < E; ++i)
        if (static_cast<uint32_t>(lane * E + i) >= P) key[i] = 0xffffffffu;

    // 3. sort by (rank, position): equal ranks stay in tree order
    warp_bitonic_sort<E>(key, lane);

    // 4. sorted (rank, weight) back to shared memory
    uint32_t rk[E];
    float w[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const bool valid = static_cast<uint32_t>(lane * E + i) < P;
        w[i] = valid ? B[key[i] & POSMASK] : 0.0f;
        rk[i] = valid ? (key[i] >> POSBITS) : 0xffffffffu;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < E; ++i) {
        A[lane * E + i] = rk[i];
        B[lane * E + i] = w[i];
    }
    __syncwarp();

    // 5. group heads sum their duplicates in tree order (ensemble.py:140), then compact
    uint32_t prev = __shfl_up_sync(0xffffffffu, rk[E - 1], 1);
    if (lane == 0) prev = 0xfffffffeu;
    float Wg[E];
    bool head[E];
    uint32_t n_heads = 0;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const uint32_t v = lane * E + i;
        const uint32_t before = (i == 0) ? prev : rk[i - 1];
        head[i] = (v < P) && (rk[i] != before);
        Wg[i] = 0.0f;
        if (head[i]) {
            float W = w[i];
            for (uint32_t u = v + 1; u < P && A[u] == rk[i]; ++u) W = __fadd_rn(W, B[u]);
            Wg[i] = W;
            head[i] = W != 0.0f;                         // utils.py:55-57
        }
        n_heads += head[i] ? 1u : 0u;
    }
    uint32_t incl = n_heads;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    const int m = static_cast<int>(__shfl_sync(0xffffffffu, incl, 31));
    __syncwarp();
    {
        uint32_t c = incl - n_heads;
#pragma unroll
        for (int i = 0; i < E; ++i)
            if (head[i]) { A[c] = rk[i]; B[c] = Wg[i]; ++c; }
    }
    __syncwarp();
    if (m == 0) {
        for (int qi = lane; qi < p.Q; qi += 32) out[qi] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }

    // 6. float32 running sum in rank order (utils.py:68).
    //    UNIT: every member weight is f32(c / c) = 1.0, so every merged weight and every partial
    //    sum is an integer below 2^24 -- float32 addition is exact on those and the sequential
    //    sum equals a warp scan, bit for bit.  Otherwise lane 0 adds in sequence, 4 entries per
    //    shared access.
    if (UNIT) {
        float wv[E], run = 0.0f;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            wv[i] = (lane * E + i < m) ? B[lane * E + i] : 0.0f;
            run = __fadd_rn(run, wv[i]);
        }
        float incl_f = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float u = __shfl_up_sync(0xffffffffu, incl_f, d);
            if (lane >= d) incl_f = __fadd_rn(incl_f, u);
        }
        float C = __fsub_rn(incl_f, run);
#pragma unroll
        for (int i = 0; i < E; ++i) {
            C = __fadd_rn(C, wv[i]);
            if (lane * E + i < m) Cc[lane * E + i] = C;
        }
    } else if (lane == 0) {
        float C = 0.0f;
        for (int c = 0; c < m; c += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(B + c);
            float4 o;
            C = __fadd_rn(C, w4.x); o.x = C;
            C = __fadd_rn(C, w4.y); o.y = C;
            C = __fadd_rn(C, w4.z); o.z = C;
            C = __fadd_rn(C, w4.w); o.w = C;
            *reinterpret_cast<float4*>(Cc + c) = o;
        }
    }
    __syncwarp();

    // 7. one lane per requested percentile
    const float s = __fdiv_rn(100.0f, Cc[m - 1]);       // utils.py:69,72
    auto rankOf = [&](int k) { return A[k]; };
    auto weightOf = [&](int k) { return B[k]; };
    auto cumOf = [&](int k) { return Cc[k]; };
    for (int qi = lane; qi < p.Q; qi += 32)
        out[qi] = select_percentile(ql.q[qi], m, s, rankOf, weightOf, cumOf, p.y_sorted);
}

// ---------------------------------------------------------------------------------------
// Shared by the CTA-per-row kernels: exclusive offsets of the row's T leaf lists and the
// gather of their members into shared memory in tree order.
// ---------------------------------------------------------------------------------------
// offs[0..T] <- exclusive prefix of the member counts, starts[t] <- first member of leaf t.
// Contains block barriers; returns the number of members of the row.
template <int THREADS>
__device__ __forceinline__ uint32_t row_leaf_offsets(const uint2* __restrict__ seg, int T, uint32_t* offs,
                                                     uint32_t* starts) {
    const int tid = threadIdx.x, lane = tid & 31;
    for (int t = tid; t < T; t += THREADS) {
        const uint2 s = seg[t];
        starts[t] = s.x;
        offs[t + 1] = s.y;
    }
    if (tid == 0) offs[0] = 0;
    __syncthreads();
    if (tid < 32) {
        uint32_t carry = 0;
        for (int base = 0; base < T; base += 32) {
            const int t = base + lane;
            uint32_t v = (t < T) ? offs[t + 1] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += u;
            }
            v += carry;
            if (t < T) offs[t + 1] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    return offs[T];
}

// 4 lanes per leaf, two 8-byte entries per lane and two leaves in flight (4 loads
// outstanding per lane); store(pos, rank, weight bits) is called once per member, pos =
// position in tree order.  lo / hi: running min / max of the ranks this thread saw.
template <int THREADS, typename Store>
__device__ __forceinline__ void gather_members(const uint2* __restrict__ members, int T,
                                               const uint32_t* __restrict__ offs,
                                               const uint32_t* __restrict__ starts, uint32_t& lo,
                                               uint32_t& hi, Store store) {
    constexpr int GROUPS = THREADS / 4;
    const int group = threadIdx.x >> 2;
    const uint32_t sub = threadIdx.x & 3;
    for (int t0 = group; t0 < T; t0 += 2 * GROUPS) {
        uint32_t o[2], len[2];
        uint2 e0[2], e1[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int t = t0 + k * GROUPS;
            o[k] = 0u;
            len[k] = 0u;
            if (t < T) {
                o[k] = offs[t];
                len[k] = offs[t + 1] - o[k];
                const uint2* __restrict__ src = members + starts[t];
                if (sub < len[k]) e0[k] = __ldg(src + sub);
                if (sub + 4 < len[k]) e1[k] = __ldg(src + sub + 4);
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (sub < len[k]) {
                store(o[k] + sub, e0[k].x, e0[k].y);
                lo = min(lo, e0[k].x);
                hi = max(hi, e0[k].x);
            }
            if (sub + 4 < len[k]) {
                store(o[k] + sub + 4, e1[k].x, e1[k].y);
                lo = min(lo, e1[k].x);
                hi = max(hi, e1[k].x);
            }
            if (len[k] > 8) {                           // long leaf lists: the rest, 4 entries per step
                const uint2* __restrict__ src = members + starts[t0 + k * GROUPS];
                for (uint32_t i = sub + 8; i < len[k]; i += 4) {
                    const uint2 x = __ldg(src + i);
                    store(o[k] + i, x.x, x.y);
                    lo = min(lo, x.x);
                    hi = max(hi, x.x);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// CTA-per-row kernel for larger member sets (P <= 16 * THREADS): bucket sort in shared
// memory, no O(P log^2 P) network.
//
//   1. gather the (rank, weight) pairs in tree order: rk[pos], wt[pos]; block min / max of
//      the ranks gives the row's rank window [rmin, rmin + range]
//   2. counting sort by bucket = (rank - rmin) >> shift with B = 8 * THREADS buckets
//      (about two members per bucket): integer shared-memory atomics for the histogram
//      and for the slots (ATOMS ~16 lanes/clk/SM measured, tools/microbench), one block
//      scan in between.  An entry is the 32-bit word (low rank bits << POSBITS) | pos.
//   3. order inside a bucket: every entry counts the smaller entries of its bucket
//      (entries are distinct, so this is a permutation) and moves to base + count.  Equal
//      ranks end up adjacent and in tree order (pos ascending).
//   4. group heads add their duplicates in tree order (float32, ensemble.py:140); block
//      compaction of the heads by ballots + one small scan
//   5. lane 0: float32 running sum in rank order (utils.py:68) with 128-bit shared accesses
//   6. one thread per requested percentile (utils.py:72-81)
//
// smem: u32 rk[cap] | f32 wt[cap] | u32 srt[cap] | u32 cnt[B + 8] | u32 offs[T + 1] | u32 starts[T]
// ---------------------------------------------------------------------------------------
constexpr int kBucketEPT = 16;                          // entries per thread at full capacity

__host__ __device__ constexpr int bucket_kernel_capacity(int threads) { return threads * kBucketEPT; }

__host__ __device__ inline size_t bucket_kernel_smem(int threads, int T) {
    const size_t cap = static_cast<size_t>(bucket_kernel_capacity(threads));
    return cap * 12 + (cap / 2 + 8) * 4 + static_cast<size_t>(2 * T + 1) * 4 + 16;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256) ? 768 / THREADS : 1)
quantile_bucket_kernel(const QuantileParams p, const __grid_constant__ QuantileList ql) {
    constexpr int CAP = THREADS * kBucketEPT;
    // CAP / 2 buckets.  (One bucket per capacity slot was measured in round 2: the in-bucket ordering of
    // step 3 gets cheaper, the scan twice as long, config 3 exact 65.3 -> 68.1 ms: not kept.)
    constexpr int NB = CAP / 2;                         // buckets
    constexpr int NW = THREADS / 32;
    constexpr int POSBITS = (THREADS == 64) ? 10 : (THREADS == 128) ? 11 : (THREADS == 256) ? 12 : 13;
    constexpr int LOG2NB = POSBITS - 1;
    constexpr uint32_t POSMASK = CAP - 1;
    static_assert(THREADS == 64 || THREADS == 128 || THREADS == 256 || THREADS == 512, "unsupported block size");
    static_assert((1 << POSBITS) == CAP, "POSBITS must be log2 of the capacity");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* rk = reinterpret_cast<uint32_t*>(smem_raw);          // rank by pos -> compact ranks
    float* wt = reinterpret_cast<float*>(rk + CAP);                // weight by pos -> compact weights
    uint32_t* srt = reinterpret_cast<uint32_t*>(wt + CAP);         // entries -> running sum
    uint32_t* cnt = srt + CAP;                                     // [NB + 8]
    uint32_t* offs = cnt + NB + 8;                                 // [T + 1]
    uint32_t* starts = offs + (p.T + 1);                           // [T]
    __shared__ uint32_t s_red[2 * NW];
    __shared__ int s_m;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t n_work = p.row_list ? static_cast<int64_t>(*p.row_count) : p.n;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int64_t row = p.row_list ? static_cast<int64_t>(p.row_list[work]) : work;
    const uint2* __restrict__ seg = p.seg + row * p.T;
    __syncthreads();                                    // shared memory reused from the previous row

    // 0. member counts of the row's leaves -> exclusive offsets (tree order)
    for (int i = tid; i < NB + 8; i += THREADS) cnt[i] = 0u;
    const uint32_t P = row_leaf_offsets<THREADS>(seg, p.T, offs, starts);
    if (P > static_cast<uint32_t>(min(p.cap, CAP))) {   // too many members for this kernel
        if (tid == 0) p.overflow_rows[atomicAdd(p.overflow_count, 1)] = static_cast<int>(row);
        continue;
    }
    double* __restrict__ out = p.out + (p.order ? static_cast<int64_t>(__ldg(p.order + row)) : row) * p.ldo;
    if (P == 0) {                                       // cannot happen for a fitted forest
        for (int qi = tid; qi < p.Q; qi += THREADS) out[qi] = __longlong_as_double(0x7ff8000000000000LL);
        continue;
    }

    // 1. gather (tree order); rank window of the row
    uint32_t rmin, shift;
    bool full;                                          // an entry holds the whole (rank - rmin)
    {
        uint32_t lo = 0xffffffffu, hi = 0u;
        gather_members<THREADS>(p.members, p.T, offs, starts, lo, hi,
                                [&](uint32_t pos, uint32_t rank, uint32_t wbits) {
                                    rk[pos] = rank;
                                    wt[pos] = __uint_as_float(wbits);
                                });
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { s_red[warp] = lo; s_red[NW + warp] = hi; }
        if (tid < 4 && P + tid < CAP) srt[P + tid] = 0xffffffffu;   // pad for the 128-bit reads of step 3
        __syncthreads();
        lo = 0xffffffffu;
        hi = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) { lo = min(lo, s_red[w]); hi = max(hi, s_red[NW + w]); }
        rmin = lo;
        const uint32_t range = hi - lo;
        const int bits = 32 - __clz(range);             // range < 2^bits
        shift = static_cast<uint32_t>(max(0, bits - LOG2NB));      // (range >> shift) < NB
        full = bits + POSBITS <= 32;
    }
    const uint32_t lowmask = full ? 0xffffffffu : ((1u << shift) - 1u);

    // 2a. bucket populations
    for (uint32_t i = tid; i < P; i += THREADS) atomicAdd(&cnt[(rk[i] - rmin) >> shift], 1u);
    __syncthreads();
    // 2b. exclusive scan of the NB = 8 * THREADS populations (8 consecutive per thread)
    {
        constexpr int C4 = NB / THREADS / 4;            // 128-bit chunks per thread
        uint4 c[C4];
        uint32_t total = 0;
#pragma unroll
        for (int k = 0; k < C4; ++k) {
            c[k] = reinterpret_cast<uint4*>(cnt)[C4 * tid + k];
            total += c[k].x + c[k].y + c[k].z + c[k].w;
        }
        uint32_t incl = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s_red[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) before += (w < warp) ? s_red[w] : 0u;
        uint32_t run = before + incl - total;
#pragma unroll
        for (int k = 0; k < C4; ++k) {
            uint4 o;
            o.x = run; run += c[k].x; o.y = run; run += c[k].y; o.z = run; run += c[k].z; o.w = run; run += c[k].w;
            reinterpret_cast<uint4*>(cnt)[C4 * tid + k] = o;
        }
    }
    __syncthreads();
    // 2c. slots: afterwards cnt[b] is the END of bucket b (= start of bucket b + 1).
    //     entry = ((rank - rmin) << POSBITS) | pos when that fits 32 bits (`full`: entries of
    //     different buckets compare like their buckets), else only the bits below the bucket
    for (uint32_t i = tid; i < P; i += THREADS) {
        const uint32_t d = rk[i] - rmin;
        const uint32_t slot = atomicAdd(&cnt[d >> shift], 1u);
        srt[slot] = ((d & lowmask) << POSBITS) | i;
    }
    __syncthreads();

    // 3. order inside the buckets: count the smaller entries, 4 per shared-memory access
    {
        uint32_t key[kBucketEPT], dst[kBucketEPT];
#pragma unroll
        for (int j = 0; j < kBucketEPT; ++j) {
            const uint32_t i = j * THREADS + tid;
            key[j] = 0u;
            dst[j] = 0xffffffffu;
            if (i < P) {
                const uint32_t k = srt[i];
                const uint32_t b = full ? ((k >> POSBITS) >> shift) : ((rk[k & POSMASK] - rmin) >> shift);
                const uint32_t lo = b ? cnt[b - 1] : 0u, hi = cnt[b];
                const uint32_t lo4 = lo & ~3u;
                uint32_t smaller = 0;
                if (full) {                             // entries before lo are smaller, entries from hi on larger
                    for (uint32_t u = lo4; u < hi; u += 4) {
                        const uint4 v = *reinterpret_cast<const uint4*>(srt + u);
                        smaller += (v.x < k) + (v.y < k) + (v.z < k) + (v.w < k);
                    }
                    dst[j] = lo4 + smaller;
                } else {
                    for (uint32_t u = lo4; u < hi; u += 4) {
                        const uint4 v = *reinterpret_cast<const uint4*>(srt + u);
                        smaller += (u >= lo && u < hi && v.x < k) + (u + 1 >= lo && u + 1 < hi && v.y < k) +
                                   (u + 2 >= lo && u + 2 < hi && v.z < k) + (u + 3 >= lo && u + 3 < hi && v.w < k);
                    }
                    dst[j] = lo + smaller;
                }
                key[j] = k;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kBucketEPT; ++j)
            if (dst[j] != 0xffffffffu) srt[dst[j]] = key[j];
    }
    __syncthreads();

    // 4. merge equal ranks in tree order (ensemble.py:140).  Thread t walks the entries
    //    [16 t, 16 t + 16) of the sorted array; a group belongs to the thread that holds its
    //    first entry, which follows it past the end of its chunk if need be, so that the
    //    float32 additions happen in exactly the reference's order.
    {
        auto rank_of = [&](uint32_t entry) -> uint32_t {
            return full ? (rmin + (entry >> POSBITS)) : rk[entry & POSMASK];
        };
        const uint32_t base = static_cast<uint32_t>(tid) * kBucketEPT;
        uint32_t ent[kBucketEPT];
#pragma unroll
        for (int c = 0; c < kBucketEPT / 4; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(srt + base + 4 * c);
            ent[4 * c + 0] = v.x; ent[4 * c + 1] = v.y; ent[4 * c + 2] = v.z; ent[4 * c + 3] = v.w;
        }
        uint32_t hr[kBucketEPT + 1];
        float hw[kBucketEPT + 1];
        uint32_t flags = 0;                             // bit j: slot j holds a finished group
        uint32_t cur = (base > 0 && base < P) ? rank_of(srt[base - 1]) : 0xffffffffu;
        float W = 0.0f;
        bool open = false;                              // the group being accumulated is this thread's
#pragma unroll
        for (int j = 0; j < kBucketEPT; ++j) {
            hr[j] = 0u;
            hw[j] = 0.0f;
            if (base + j < P) {
                const uint32_t r = rank_of(ent[j]);
                const float w = wt[ent[j] & POSMASK];
                if (r != cur) {
                    if (open && W != 0.0f) { hr[j] = cur; hw[j] = W; flags |= 1u << j; }   // utils.py:55-57
                    cur = r;
                    W = w;
                    open = true;
                } else if (open) {
                    W = __fadd_rn(W, w);
                }
            }
        }
        hr[kBucketEPT] = 0u;
        hw[kBucketEPT] = 0.0f;
        if (open) {
            for (uint32_t u = base + kBucketEPT; u < P; ++u) {
                const uint32_t en = srt[u];
                if (rank_of(en) != cur) break;
                W = __fadd_rn(W, wt[en & POSMASK]);
            }
            if (W != 0.0f) { hr[kBucketEPT] = cur; hw[kBucketEPT] = W; flags |= 1u << kBucketEPT; }
        }
        // compaction: exclusive scan of the per-thread group counts (blocked order = rank order)
        const uint32_t mine = __popc(flags);
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s_red[warp] = incl;
        __syncthreads();                                // every read of rk / wt / srt is done
        uint32_t c = incl - mine;
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t v = s_red[w];
            c += (w < warp) ? v : 0u;
            total += v;
        }
        if (tid == 0) s_m = static_cast<int>(total);
#pragma unroll
        for (int j = 0; j <= kBucketEPT; ++j) {
            if (flags & (1u << j)) {
                rk[c] = hr[j];
                wt[c] = hw[j];
                ++c;
            }
        }
    }
    __syncthreads();
    const int m = s_m;
    if (m == 0) {
        for (int qi = tid; qi < p.Q; qi += THREADS) out[qi] = __longlong_as_double(0x7ff8000000000000LL);
        continue;
    }

    // 5. lane 0: float32 running sum in rank order (utils.py:68), 4 entries per shared access
    float* cum = reinterpret_cast<float*>(srt);
    if (tid == 0) {
        // the chain of dependent float32 adds is the floor (4 cycles each); the loads run two chunks ahead
        // of it so that no add ever waits for shared memory (wt is padded: reads past m stay inside it)
        float C = 0.0f;
        float4 w4 = *reinterpret_cast<const float4*>(wt);
        float4 n4 = *reinterpret_cast<const float4*>(wt + 4);
        for (int c = 0; c < m; c += 4) {
            const float4 f4 = *reinterpret_cast<const float4*>(wt + min(c + 8, CAP - 4));
            float4 o;
            C = __fadd_rn(C, w4.x); o.x = C;
            C = __fadd_rn(C, w4.y); o.y = C;
            C = __fadd_rn(C, w4.z); o.z = C;
            C = __fadd_rn(C, w4.w); o.w = C;
            *reinterpret_cast<float4*>(cum + c) = o;
            w4 = n4;
            n4 = f4;
        }
    }
    __syncthreads();

    // 6. every requested percentile
    const float s = __fdiv_rn(100.0f, cum[m - 1]);      // utils.py:69,72
    auto rankOf = [&](int k) { return rk[k]; };
    auto weightOf = [&](int k) { return wt[k]; };
    auto cumOf = [&](int k) { return cum[k]; };
    for (int qi = tid; qi < p.Q; qi += THREADS)
        out[qi] = select_percentile(ql.q[qi], m, s, rankOf, weightOf, cumOf, p.y_sorted);
    }   // rows
}

// ---------------------------------------------------------------------------------------
// Dense fallback for rows whose member set does not fit shared memory (shallow trees /
// large min_samples_leaf): accumulate the weights into a rank-indexed float32 vector in
// global scratch, exactly like the reference's masked sum but touching only members.
// Per CTA scratch: f32 dense[N] (kept zero between rows) | u32 bitmap[ceil(N/32)] (kept
// zero) | u32 c_rank[N] | f32 c_w[N] | f32 c_cum[N].
// ---------------------------------------------------------------------------------------
struct DenseScratch {
    float* dense;
    uint32_t* bitmap;
    uint32_t* c_rank;
    float* c_w;
    float* c_cum;
};

__host__ __device__ inline int64_t dense_scratch_bytes(int64_t N) {
    const int64_t words = (N + 31) / 32;
    int64_t b = (4 * N + 4 * words + 12 * N);
    return (b + 255) / 256 * 256;
}

__device__ inline DenseScratch dense_scratch_at(unsigned char* base, int64_t N, int cta) {
    unsigned char* ptr = base + dense_scratch_bytes(N) * cta;
    const int64_t words = (N + 31) / 32;
    DenseScratch s;
    s.dense = reinterpret_cast<float*>(ptr);
    s.bitmap = reinterpret_cast<uint32_t*>(s.dense + N);
    s.c_rank = s.bitmap + words;
    s.c_w = reinterpret_cast<float*>(s.c_rank + N);
    s.c_cum = s.c_w + N;
    return s;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
quantile_dense_kernel(const QuantileParams p, const __grid_constant__ QuantileList ql,
                      unsigned char* scratch_base, int64_t N) {
    __shared__ uint32_t warp_sums[THREADS / 32];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const DenseScratch sc = dense_scratch_at(scratch_base, N, blockIdx.x);
    const int64_t words = (N + 31) / 32;
    const int n_over = *p.row_count;                    // this kernel always runs off a row list

    for (int oi = blockIdx.x; oi < n_over; oi += gridDim.x) {
        const int64_t row = p.row_list[oi];
        const uint2* __restrict__ seg = p.seg + row * p.T;

        // A. trees in order; inside one leaf every sample appears once, so the plain
        //    read-modify-write below is race free and the per-sample sum is in tree order.
        for (int t = 0; t < p.T; ++t) {
            const uint2 s = seg[t];
            const uint2* __restrict__ src = p.members + s.x;
            for (uint32_t i = tid; i < s.y; i += THREADS) {
                const uint2 e = __ldg(src + i);
                sc.dense[e.x] = __fadd_rn(sc.dense[e.x], __uint_as_float(e.y));   // ensemble.py:140
                atomicOr(sc.bitmap + (e.x >> 5), 1u << (e.x & 31));
            }
            __syncthreads();
        }

        // B. compact the touched ranks in rank order (and restore the scratch to zero)
        if (tid == 0) s_base = 0;
        __syncthreads();
        for (int64_t w0 = 0; w0 < words; w0 += THREADS) {
            const int64_t w = w0 + tid;
            uint32_t bits = (w < words) ? sc.bitmap[w] : 0u;
            const uint32_t c = __popc(bits);
            uint32_t v = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += u;
            }
            if (lane == 31) warp_sums[warp] = v;
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int i = 0; i < THREADS / 32; ++i) {
                const uint32_t ws = warp_sums[i];
                if (i < warp) before += ws;
                total += ws;
            }
            uint32_t o = s_base + before + v - c;
            if (bits) {
                sc.bitmap[w] = 0u;
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const uint32_t r = static_cast<uint32_t>(w * 32 + b);
                    sc.c_rank[o] = r;
                    sc.c_w[o] = sc.dense[r];
                    sc.dense[r] = 0.0f;
                    ++o;
                }
            }
            __syncthreads();
            if (tid == 0) s_base += total;
            __syncthreads();
        }
        const int m_all = static_cast<int>(s_base);

        // C. warp 0: float32 running sum in rank order (utils.py:68), dropping zero weights
        //    (utils.py:55-57); sequential arithmetic replayed by all lanes on shuffled values.
        __shared__ int s_m;
        if (tid < 32) {
            int m = 0;
            float C = 0.0f;
            for (int base = 0; base < m_all; base += 32) {
                const int k = base + lane;
                const float wk = (k < m_all) ? sc.c_w[k] : 0.0f;
                const uint32_t rk = (k < m_all) ? sc.c_rank[k] : 0u;
                const int cnt = min(32, m_all - base);
                float my_c = 0.0f;
                int my_slot = -1;
                for (int i = 0; i < cnt; ++i) {
                    const float wi = __shfl_sync(0xffffffffu, wk, i);
                    if (wi != 0.0f) {
                        C = __fadd_rn(C, wi);
                        if (lane == i) { my_c = C; my_slot = m; }
                        ++m;
                    }
                }
                __syncwarp();
                // slots are <= k, and every read of this batch happened above
                if (my_slot >= 0) { sc.c_rank[my_slot] = rk; sc.c_w[my_slot] = wk; sc.c_cum[my_slot] = my_c; }
                __syncwarp();
            }
            if (lane == 0) s_m = m;
        }
        __syncthreads();

        // D. percentiles
        const int m = s_m;
        double* __restrict__ out = p.out + (p.order ? static_cast<int64_t>(__ldg(p.order + row)) : row) * p.ldo;
        if (m == 0) {
            for (int qi = tid; qi < p.Q; qi += THREADS) out[qi] = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const float s = __fdiv_rn(100.0f, sc.c_cum[m - 1]);
            auto rankOf = [&](int k) { return sc.c_rank[k]; };
            auto weightOf = [&](int k) { return sc.c_w[k]; };
            auto cumOf = [&](int k) { return sc.c_cum[k]; };
            for (int qi = tid; qi < p.Q; qi += THREADS)
                out[qi] = select_percentile(ql.q[qi], m, s, rankOf, weightOf, cumOf, p.y_sorted);
        }
        __syncthreads();                                 // scratch reused by the next row
    }
}

}  // namespace qf
