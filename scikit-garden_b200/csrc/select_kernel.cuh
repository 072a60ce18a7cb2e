// K2, QF_MODE_FAST -- weighted percentile by hierarchical histogram selection: no sort, no
// staging of the members, no limit on the number of members per row.
//
// Same result as quantile_kernel.cuh up to the summation order of the weights (the
// reference's own float32 running sum carries ~1e-5 relative error; this path is within
// 1e-4 relative of it, the tolerance BASELINE.json states for float32 arithmetic; order
// statistics -- q = 0, q = 100, pure leaves -- are bit-exact).
//
// For a requested percentile q the reference (skgarden/quantile/utils.py:68-81) needs only
// three entries of the merged, rank-ordered member list: with tau = q/100 * total weight and
// e = the first entry whose running weight C_e reaches tau, the interpolation interval is
// (e-1, e) or (e, e+1) depending on the sign of p_e - q.  So instead of sorting P members the
// CTA streams the row's leaf lists from global memory twice (the second time out of L2):
//
//   pass 1: per bucket of the rank axis (2048 buckets over [0, N)) the weight, the largest and
//           the smallest rank -- three integer shared-memory atomics per member.  Weights are
//           32-bit fixed point (scale 2^31 / next_pow2(T + 1): the weights of one leaf sum to
//           1, so a row's total is <= T), which makes every sum exact whatever the order;
//           duplicates of a training sample merge by landing in the same slot.
//           Block scan; per q a binary search for the bucket that holds e, and the nearest
//           rank in the non-empty buckets before and after it.
//   pass 2: weight histogram of that bucket alone at single-rank resolution (one bucket spans
//           N / 2048 <= 1024 ranks for N <= 2M), plus the merged weights of the two outside
//           neighbours.  Scan, locate e and its neighbours inside the bucket.
//           (Larger N: the bucket is refined by further passes at 2^sel_bits slots per level;
//           the neighbours are then tracked as running max / min outside the final window
//           and their weights collected by one more pass.)
//   then utils.py:72-81 in float32 on (pred, e, succ).
//
// Shared memory: 8 KB level-1 weights + 16 KB (refinement slots, doubling as the per-bucket
// max / min during pass 1) -- eight 256-thread CTAs per SM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "quantile_kernel.cuh"

namespace qf {

constexpr int kSelectQB = 4;                            // percentiles per refinement pass
constexpr int kSelectL2Bits = 10;
constexpr int kSelectL2Slots = 1 << kSelectL2Bits;      // refinement slots per percentile
constexpr int kSelectMaxQ = 16;                         // above this the sorting kernel is cheaper
constexpr int kSelectLog2Buckets = 11;
constexpr int kSelectBuckets = 1 << kSelectLog2Buckets; // level-1 buckets over the rank axis

__host__ __device__ inline size_t select_kernel_smem() {
    return static_cast<size_t>(kSelectBuckets + 8) * 4 + static_cast<size_t>(kSelectQB) * kSelectL2Slots * 4;
}

// f(rank, weight bits) for every member of the row: 4 lanes per leaf list, two lists and up
// to four 8-byte loads in flight per lane.  No barriers; every thread of the CTA calls it.
template <int THREADS, typename F>
__device__ __forceinline__ void for_each_member(const uint2* __restrict__ seg, const uint2* __restrict__ members,
                                                int T, F f) {
    constexpr int GROUPS = THREADS / 4;
    const int group = threadIdx.x >> 2;
    const uint32_t sub = threadIdx.x & 3;
    for (int t0 = group; t0 < T; t0 += 2 * GROUPS) {
        uint2 s[2], e0[2], e1[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int t = t0 + k * GROUPS;
            s[k] = (t < T) ? seg[t] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint2* __restrict__ src = members + s[k].x;
            if (sub < s[k].y) e0[k] = __ldg(src + sub);
            if (sub + 4 < s[k].y) e1[k] = __ldg(src + sub + 4);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (sub < s[k].y) f(e0[k].x, e0[k].y);
            if (sub + 4 < s[k].y) f(e1[k].x, e1[k].y);
            if (s[k].y > 8) {                           // long leaf lists: the rest, 4 entries per step
                const uint2* __restrict__ src = members + s[k].x;
                for (uint32_t i = sub + 8; i < s[k].y; i += 4) {
                    const uint2 x = __ldg(src + i);
                    f(x.x, x.y);
                }
            }
        }
    }
}

// n_train_bits: bits of N - 1 (ranks are < 2^n_train_bits).  sel_bits (<= kSelectL2Bits):
// log2 of the refinement slots actually used; small values force several refinement levels
// on small forests (tests).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
quantile_select_kernel(const QuantileParams p, const __grid_constant__ QuantileList ql, const float fx_scale,
                       const int n_train_bits, const int sel_bits) {
    constexpr int NB = kSelectBuckets;
    constexpr int BPT = NB / THREADS;                   // buckets scanned per thread
    constexpr int NW = THREADS / 32;
    constexpr int QB = kSelectQB;
    constexpr int L2S = kSelectL2Slots;
    constexpr int SPT = QB * L2S / THREADS;             // refinement slots scanned per thread
    constexpr uint32_t NONE = 0xffffffffu;
    static_assert(THREADS == 128 || THREADS == 256, "unsupported block size");
    static_assert(L2S % SPT == 0 && SPT % 4 == 0 && BPT % 4 == 0 && 2 * NB <= QB * L2S, "layout");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* h1 = reinterpret_cast<uint32_t*>(smem_raw);          // [NB + 8] level-1 weights -> inclusive prefix
    uint32_t* l2 = h1 + NB + 8;                                    // [QB][L2S]; during pass 1: bmx[NB] | bmn[NB]
    uint32_t* bmx = l2;                                            // largest rank + 1 per bucket
    uint32_t* bmn = l2 + NB;                                       // smallest rank per bucket
    __shared__ uint32_t s_red[NW];
    // per percentile of the call: target weight, window start (rank), weight before the window,
    // nearest rank + 1 below / rank above the window
    __shared__ uint32_t s_tau[kSelectMaxQ], s_wlo[kSelectMaxQ], s_base[kSelectMaxQ];
    __shared__ uint32_t s_pout[kSelectMaxQ], s_sout[kSelectMaxQ];
    // per percentile of the batch: final slot of e, its merged weight, running weight;
    // neighbours of e inside the window (slot + 1 / slot); merged weights of the outside ones
    __shared__ uint32_t s_e[QB], s_We[QB], s_Ce[QB], s_pin[QB], s_sin[QB], s_Wp[QB], s_Ws[QB], s_seg[QB];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int shift1 = max(0, n_train_bits - kSelectLog2Buckets);  // (rank >> shift1) < NB
    const bool single = shift1 <= sel_bits;             // one refinement level reaches single ranks
    const int64_t n_work = p.row_list ? static_cast<int64_t>(*p.row_count) : p.n;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int64_t row = p.row_list ? static_cast<int64_t>(p.row_list[work]) : work;
    const uint2* __restrict__ seg = p.seg + row * p.T;
    double* __restrict__ out = p.out + (p.order ? static_cast<int64_t>(__ldg(p.order + row)) : row) * p.ldo;
    __syncthreads();                                    // shared memory reused from the previous row
    for (int i = tid; i < NB + 8; i += THREADS) h1[i] = 0u;
    for (int i = tid; i < NB; i += THREADS) { bmx[i] = 0u; bmn[i] = NONE; }
    __syncthreads();

    // pass 1: weight, largest and smallest rank per bucket
    for_each_member<THREADS>(seg, p.members, p.T, [&](uint32_t rank, uint32_t wbits) {
        const uint32_t w = __float2uint_rn(__fmul_rn(__uint_as_float(wbits), fx_scale));
        if (w != 0u) {                                  // utils.py:55-57
            const uint32_t b = rank >> shift1;
            atomicAdd(&h1[b], w);
            atomicMax(&bmx[b], rank + 1u);
            atomicMin(&bmn[b], rank);
        }
    });
    __syncthreads();
    {   // inclusive prefix of the bucket weights, BPT consecutive buckets per thread
        uint32_t c[BPT];
#pragma unroll
        for (int k = 0; k < BPT / 4; ++k) {
            const uint4 x = reinterpret_cast<const uint4*>(h1)[tid * (BPT / 4) + k];
            c[4 * k] = x.x; c[4 * k + 1] = x.y; c[4 * k + 2] = x.z; c[4 * k + 3] = x.w;
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
        __syncthreads();
        uint32_t run = incl - mine;
#pragma unroll
        for (int w = 0; w < NW; ++w) run += (w < warp) ? s_red[w] : 0u;
#pragma unroll
        for (int k = 0; k < BPT; ++k) { run += c[k]; c[k] = run; }
#pragma unroll
        for (int k = 0; k < BPT / 4; ++k)
            reinterpret_cast<uint4*>(h1)[tid * (BPT / 4) + k] = make_uint4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
    }
    __syncthreads();
    const uint32_t total = h1[NB - 1];
    if (total == 0u) {                                  // no member with a non-zero weight: cannot happen for a fitted forest
        for (int qi = tid; qi < p.Q; qi += THREADS) out[qi] = __longlong_as_double(0x7ff8000000000000LL);
        continue;
    }

    // every percentile of the call: target weight, level-1 bucket, outside neighbours
    if (tid < p.Q) {
        const double t = ceil(ql.q[tid] * static_cast<double>(total) / 100.0);
        const uint32_t tau = static_cast<uint32_t>(fmin(fmax(t, 1.0), static_cast<double>(total)));
        int lo = 0, hi = NB - 1;                        // first bucket whose inclusive prefix reaches tau
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (h1[mid] >= tau) hi = mid; else lo = mid + 1;
        }
        const int bq = lo;
        const uint32_t base = bq ? h1[bq - 1] : 0u, upto = h1[bq];
        uint32_t pout = 0u, sout = NONE;
        if (single) {
            if (base != 0u) {                           // last non-empty bucket before bq
                lo = 0; hi = bq - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (h1[mid] >= base) hi = mid; else lo = mid + 1;
                }
                pout = bmx[lo];
            }
            if (upto != total) {                        // first non-empty bucket after bq
                lo = bq + 1; hi = NB - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (h1[mid] > upto) hi = mid; else lo = mid + 1;
                }
                sout = bmn[lo];
            }
        }
        s_tau[tid] = tau;
        s_wlo[tid] = static_cast<uint32_t>(bq) << shift1;
        s_base[tid] = base;
        s_pout[tid] = pout;
        s_sout[tid] = sout;
    }
    __syncthreads();

    for (int q0 = 0; q0 < p.Q; q0 += QB) {
        const int nqb = min(QB, p.Q - q0);
        if (tid < QB) {
            s_pin[tid] = 0u; s_sin[tid] = NONE; s_Wp[tid] = 0u; s_Ws[tid] = 0u;
            s_e[tid] = 0u; s_We[tid] = 0u; s_Ce[tid] = 0u;
        }
        // refine the window of every percentile of the batch down to single ranks
        int wb = shift1;                                // the window of a percentile spans 2^wb ranks
        for (;;) {
            const int sh = max(0, wb - sel_bits);       // ranks per slot = 2^sh
            const bool final_level = sh == 0;
            for (int i = tid; i < QB * L2S; i += THREADS) l2[i] = 0u;
            __syncthreads();
            {
                uint32_t wlo[QB], po[QB], so[QB], pm[QB], sn[QB];
#pragma unroll
                for (int q = 0; q < QB; ++q) {
                    const bool on = q < nqb;
                    wlo[q] = on ? s_wlo[q0 + q] : 0u;
                    po[q] = on ? s_pout[q0 + q] : 0u;
                    so[q] = on ? s_sout[q0 + q] : NONE;
                    pm[q] = 0u;
                    sn[q] = NONE;
                }
                const uint32_t wsize = 1u << wb;
                for_each_member<THREADS>(seg, p.members, p.T, [&](uint32_t rank, uint32_t wbits) {
                    const uint32_t w = __float2uint_rn(__fmul_rn(__uint_as_float(wbits), fx_scale));
                    if (w == 0u) return;                // utils.py:55-57
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        if (q < nqb) {
                            const uint32_t off = rank - wlo[q];
                            if (off < wsize) atomicAdd(&l2[q * L2S + (off >> sh)], w);
                            if (single) {               // weights of the outside neighbours, known up front
                                if (rank + 1u == po[q]) atomicAdd(&s_Wp[q], w);
                                if (rank == so[q]) atomicAdd(&s_Ws[q], w);
                            } else {                    // nearest ranks outside the window
                                pm[q] = max(pm[q], (rank < wlo[q]) ? rank + 1u : 0u);
                                sn[q] = min(sn[q], (off >= wsize && rank > wlo[q]) ? rank : NONE);
                            }
                        }
                    }
                });
                if (!single && final_level) {
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        if (q < nqb) {
                            const uint32_t a = __reduce_max_sync(0xffffffffu, pm[q]);
                            const uint32_t b = __reduce_min_sync(0xffffffffu, sn[q]);
                            if (lane == 0) {
                                if (a) atomicMax(&s_pout[q0 + q], a);
                                if (b != NONE) atomicMin(&s_sout[q0 + q], b);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // scan the QB * L2S slots (SPT consecutive slots per thread), find the slot of e
            {
                uint32_t v[SPT];
#pragma unroll
                for (int c = 0; c < SPT / 4; ++c) {
                    const uint4 x = reinterpret_cast<const uint4*>(l2)[tid * (SPT / 4) + c];
                    v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
                }
                uint32_t mine = 0;
#pragma unroll
                for (int c = 0; c < SPT; ++c) mine += v[c];
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += u;
                }
                if (lane == 31) s_red[warp] = incl;
                __syncthreads();
                uint32_t excl = incl - mine;
#pragma unroll
                for (int w = 0; w < NW; ++w) excl += (w < warp) ? s_red[w] : 0u;
                const int q = (tid * SPT) / L2S;
                const int slot0 = (tid * SPT) % L2S;
                if (slot0 == 0) s_seg[q] = excl;
                __syncthreads();
                uint32_t found = NONE, fw = 0u, fc = 0u;
                if (q < nqb) {
                    const uint32_t tau = s_tau[q0 + q];
                    uint32_t run = s_base[q0 + q] + (excl - s_seg[q]);
#pragma unroll
                    for (int c = 0; c < SPT; ++c) {
                        const uint32_t before = run;
                        run += v[c];
                        if (v[c] != 0u && before < tau && run >= tau) { found = slot0 + c; fw = v[c]; fc = run; }
                    }
                }
                __syncthreads();                        // every thread has read s_base before it moves
                if (found != NONE) {                    // exactly one thread per percentile gets here
                    s_e[q] = found;
                    s_We[q] = fw;
                    s_Ce[q] = fc;
                    if (!final_level) {
                        s_wlo[q0 + q] += found << sh;
                        s_base[q0 + q] = fc - fw;
                    }
                }
                __syncthreads();
                if (final_level && q < nqb) {           // neighbours of e inside the window
                    const uint32_t e = s_e[q];
                    uint32_t pin = 0u, sin = NONE;
#pragma unroll
                    for (int c = 0; c < SPT; ++c) {
                        const uint32_t s = slot0 + c;
                        if (v[c] != 0u && s < e) pin = s + 1u;
                        if (v[c] != 0u && s > e && sin == NONE) sin = s;
                    }
                    if (pin) atomicMax(&s_pin[q], pin);
                    if (sin != NONE) atomicMin(&s_sin[q], sin);
                }
            }
            __syncthreads();
            if (final_level) break;
            wb = sh;
        }
        // several levels only: merged weights of the outside neighbours that are needed
        if (!single) {
            bool need = false;
            uint32_t pd[QB], sd[QB];
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                pd[q] = NONE;
                sd[q] = NONE;
                if (q < nqb) {
                    if (s_pin[q] == 0u && s_pout[q0 + q] != 0u) { pd[q] = s_pout[q0 + q] - 1u; need = true; }
                    if (s_sin[q] == NONE && s_sout[q0 + q] != NONE) { sd[q] = s_sout[q0 + q]; need = true; }
                }
            }
            if (need) {                                 // block-uniform: computed from shared state
                for_each_member<THREADS>(seg, p.members, p.T, [&](uint32_t rank, uint32_t wbits) {
                    const uint32_t w = __float2uint_rn(__fmul_rn(__uint_as_float(wbits), fx_scale));
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        if (rank == pd[q]) atomicAdd(&s_Wp[q], w);
                        if (rank == sd[q]) atomicAdd(&s_Ws[q], w);
                    }
                });
            }
            __syncthreads();
        }
        // utils.py:69-81 on (pred, e, succ), float32
        if (tid < nqb) {
            const int q = tid;
            const double qv = ql.q[q0 + q];
            const float inv = __fdiv_rn(1.0f, fx_scale);
            const float s = __fdiv_rn(100.0f, __fmul_rn(static_cast<float>(total), inv));     // utils.py:69,72
            const uint32_t wlo = s_wlo[q0 + q];         // window of the final level: slot == rank offset
            const uint32_t re = wlo + s_e[q];
            const uint32_t We = s_We[q], Ce = s_Ce[q];
            const float a_e = __ldg(p.y_sorted + re);
            const float p_e = plotting_position(s, __fmul_rn(static_cast<float>(Ce), inv),
                                                __fmul_rn(static_cast<float>(We), inv));
            float result = a_e;
            if (static_cast<double>(p_e) < qv) {        // interval (e, succ)
                uint32_t rs = NONE, Ws = 0u;
                if (s_sin[q] != NONE) { rs = wlo + s_sin[q]; Ws = l2[q * L2S + s_sin[q]]; }
                else if (s_sout[q0 + q] != NONE) { rs = s_sout[q0 + q]; Ws = s_Ws[q]; }
                if (rs != NONE) {                       // else q lies beyond the last entry: a_last (utils.py:74-75)
                    const float a_s = __ldg(p.y_sorted + rs);
                    const float p_s = plotting_position(s, __fmul_rn(static_cast<float>(Ce + Ws), inv),
                                                        __fmul_rn(static_cast<float>(Ws), inv));
                    const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_e), __fsub_rn(p_s, p_e));
                    result = __fadd_rn(a_e, __fmul_rn(frac, __fsub_rn(a_s, a_e)));
                }
            } else {                                    // interval (pred, e)
                uint32_t rp = NONE, Wp = 0u;
                if (s_pin[q] != 0u) { rp = wlo + s_pin[q] - 1u; Wp = l2[q * L2S + s_pin[q] - 1u]; }
                else if (s_pout[q0 + q] != 0u) { rp = s_pout[q0 + q] - 1u; Wp = s_Wp[q]; }
                if (rp != NONE) {                       // else q lies before the first entry: a_0 (utils.py:76-77)
                    const float a_p = __ldg(p.y_sorted + rp);
                    const float p_p = plotting_position(s, __fmul_rn(static_cast<float>(Ce - We), inv),
                                                        __fmul_rn(static_cast<float>(Wp), inv));
                    const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_p), __fsub_rn(p_e, p_p));
                    result = __fadd_rn(a_p, __fmul_rn(frac, __fsub_rn(a_e, a_p)));
                }
            }
            out[q0 + q] = static_cast<double>(result);
        }
        __syncthreads();
    }   // percentile batches
    }   // rows
}

}  // namespace qf
