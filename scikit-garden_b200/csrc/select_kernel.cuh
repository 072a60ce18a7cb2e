// K2, QF_MODE_FAST -- weighted percentile by hierarchical histogram selection (no sort).
//
// Same result as quantile_kernel.cuh up to the summation order of the weights (the
// reference's own float32 running sum carries ~1e-5 relative error; this path is within
// 1e-4 relative of it, the tolerance BASELINE.json states for float32 arithmetic; order
// statistics -- q = 0, q = 100, pure leaves -- are bit-exact).
//
// For a requested percentile q the reference (skgarden/quantile/utils.py:68-81) needs only
// three entries of the merged, rank-ordered member list: with tau = q/100 * total weight and
// e = the first entry whose running weight C_e reaches tau, the interpolation interval is
// (e-1, e) or (e, e+1) depending on the sign of p_e - q.  So instead of sorting P members:
//
//   1. gather (rank, weight) in tree order into shared memory; weights become 32-bit fixed
//      point (scale 2^31 / next_pow2(T + 1): the weights of one leaf sum to 1, so a row's
//      total is <= T), which makes every accumulation an exact integer atomic whatever the
//      order -- duplicates of a training sample merge by landing in the same slot
//   2. level 1: per bucket of the row's rank window (NB <= 2048 buckets) the weight, the
//      largest and the smallest rank (three integer atomics per member); block scan; per q a
//      binary search for the bucket that holds e, and the nearest rank in the non-empty
//      buckets before and after it
//   3. refinement: weight histogram of that bucket alone at 2^sel_bits slots per q, repeated
//      until a slot is a single rank (one level when the bucket spans <= 1024 ranks, i.e. for
//      rank windows up to 2M); the same pass collects the merged weights of the two outside
//      neighbours.  (More levels: the neighbours are tracked as running max / min of the
//      ranks outside the final window and their weights collected by one more pass.)
//   4. e, its neighbours inside the window (else the outside ones) -> utils.py:72-81 in float32
//
// Handles up to kSelectQB percentiles per pass over the shared-memory resident members.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "quantile_kernel.cuh"

namespace qf {

constexpr int kSelectQB = 4;                            // percentiles per refinement pass
constexpr int kSelectL2Bits = 10;
constexpr int kSelectL2Slots = 1 << kSelectL2Bits;      // refinement slots per percentile
constexpr int kSelectMaxQ = 16;                         // above this the sorting kernel is cheaper

__host__ __device__ constexpr int select_kernel_buckets(int threads) {
    return threads * 8 < 2048 ? threads * 8 : 2048;
}

__host__ __device__ inline size_t select_kernel_smem(int threads, int T) {
    const size_t cap = static_cast<size_t>(bucket_kernel_capacity(threads));
    return cap * 8 + (static_cast<size_t>(select_kernel_buckets(threads)) + 8) * 4 +
           static_cast<size_t>(kSelectQB) * kSelectL2Slots * 4 + static_cast<size_t>(2 * T + 1) * 4 + 16;
}

// sel_bits (<= kSelectL2Bits): log2 of the refinement slots actually used; smaller values
// force several refinement levels on small forests (tests)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 64) ? 8 : (THREADS <= 256) ? 768 / THREADS : 1)
quantile_select_kernel(const QuantileParams p, const __grid_constant__ QuantileList ql, const float fx_scale,
                       const int sel_bits) {
    constexpr int CAP = THREADS * kBucketEPT;
    constexpr int NB = select_kernel_buckets(THREADS);  // level-1 buckets
    constexpr int BPT = NB / THREADS;                   // buckets scanned per thread (8 or 4)
    constexpr int NW = THREADS / 32;
    constexpr int LOG2NB = (THREADS == 64) ? 9 : (THREADS == 128) ? 10 : 11;
    constexpr int QB = kSelectQB;
    constexpr int L2S = kSelectL2Slots;
    constexpr int SPT = QB * L2S / THREADS;             // refinement slots scanned per thread
    constexpr uint32_t NONE = 0xffffffffu;
    static_assert(THREADS == 64 || THREADS == 128 || THREADS == 256 || THREADS == 512, "unsupported block size");
    static_assert((1 << LOG2NB) == NB && L2S % SPT == 0 && SPT % 4 == 0 && BPT % 4 == 0 && 2 * NB <= QB * L2S, "layout");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* rk = reinterpret_cast<uint32_t*>(smem_raw);          // rank by pos
    uint32_t* wfx = rk + CAP;                                      // fixed-point weight by pos
    uint32_t* h1 = wfx + CAP;                                      // [NB + 8] level-1 weights -> inclusive prefix
    uint32_t* l2 = h1 + NB + 8;                                    // [QB][L2S]; during level 1: bmx[NB] | bmn[NB]
    uint32_t* offs = l2 + QB * L2S;                                // [T + 1]
    uint32_t* starts = offs + (p.T + 1);                           // [T]
    uint32_t* bmx = l2;                                            // largest (rank - rmin) + 1 per bucket
    uint32_t* bmn = l2 + NB;                                       // smallest rank - rmin per bucket
    __shared__ uint32_t s_red[2 * NW];
    // per percentile of the call: target weight, window start (rank - rmin), weight before the
    // window, nearest (rank - rmin) + 1 below / rank - rmin above the window
    __shared__ uint32_t s_tau[kSelectMaxQ], s_wlo[kSelectMaxQ], s_base[kSelectMaxQ];
    __shared__ uint32_t s_pout[kSelectMaxQ], s_sout[kSelectMaxQ];
    // per percentile of the batch: final slot of e, its merged weight, running weight;
    // neighbours of e inside the window (slot + 1 / slot); merged weights of the outside ones
    __shared__ uint32_t s_e[QB], s_We[QB], s_Ce[QB], s_pin[QB], s_sin[QB], s_Wp[QB], s_Ws[QB], s_seg[QB];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t n_work = p.row_list ? static_cast<int64_t>(*p.row_count) : p.n;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int64_t row = p.row_list ? static_cast<int64_t>(p.row_list[work]) : work;
    const uint2* __restrict__ seg = p.seg + row * p.T;
    __syncthreads();                                    // shared memory reused from the previous row

    // 0. member counts of the row's leaves -> exclusive offsets (tree order)
    for (int i = tid; i < NB + 8; i += THREADS) h1[i] = 0u;
    for (int i = tid; i < NB; i += THREADS) { bmx[i] = 0u; bmn[i] = NONE; }
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

    // 1. gather (tree order), weights to fixed point; rank window of the row
    uint32_t rmin;
    int shift1;
    {
        uint32_t lo = NONE, hi = 0u;
        gather_members<THREADS>(p.members, p.T, offs, starts, lo, hi,
                                [&](uint32_t pos, uint32_t rank, uint32_t wbits) {
                                    rk[pos] = rank;
                                    wfx[pos] = __float2uint_rn(__fmul_rn(__uint_as_float(wbits), fx_scale));
                                });
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { s_red[warp] = lo; s_red[NW + warp] = hi; }
        __syncthreads();
        lo = NONE;
        hi = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) { lo = min(lo, s_red[w]); hi = max(hi, s_red[NW + w]); }
        rmin = lo;
        const int bits = 32 - __clz(hi - lo);           // rank window < 2^bits
        shift1 = max(0, bits - LOG2NB);                 // (window >> shift1) < NB
    }
    const bool single = shift1 <= sel_bits;             // one refinement level reaches single ranks

    // 2. level 1: weight, largest and smallest rank per bucket; inclusive prefix of the weights
    for (uint32_t i = tid; i < P; i += THREADS) {
        const uint32_t w = wfx[i];
        if (w == 0u) continue;                          // utils.py:55-57
        const uint32_t d = rk[i] - rmin, b = d >> shift1;
        atomicAdd(&h1[b], w);
        atomicMax(&bmx[b], d + 1u);
        atomicMin(&bmn[b], d);
    }
    __syncthreads();
    {
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

    // 3a. every percentile of the call: target weight, level-1 bucket, outside neighbours
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
        // 3b. refine the window of every percentile of the batch down to single ranks
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
                for (uint32_t i = tid; i < P; i += THREADS) {
                    const uint32_t w = wfx[i];
                    if (w == 0u) continue;              // utils.py:55-57
                    const uint32_t d = rk[i] - rmin;
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        if (q < nqb) {
                            const uint32_t off = d - wlo[q];
                            if (off < wsize) atomicAdd(&l2[q * L2S + (off >> sh)], w);
                            if (single) {               // weights of the outside neighbours, known up front
                                if (d + 1u == po[q]) atomicAdd(&s_Wp[q], w);
                                if (d == so[q]) atomicAdd(&s_Ws[q], w);
                            } else {                    // nearest ranks outside the window
                                pm[q] = max(pm[q], (d < wlo[q]) ? d + 1u : 0u);
                                sn[q] = min(sn[q], (off >= wsize && d > wlo[q]) ? d : NONE);
                            }
                        }
                    }
                }
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
        // 3c. several levels only: merged weights of the outside neighbours that are needed
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
                for (uint32_t i = tid; i < P; i += THREADS) {
                    const uint32_t d = rk[i] - rmin;
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        if (d == pd[q]) atomicAdd(&s_Wp[q], wfx[i]);
                        if (d == sd[q]) atomicAdd(&s_Ws[q], wfx[i]);
                    }
                }
            }
            __syncthreads();
        }
        // 4. utils.py:69-81 on (pred, e, succ), float32
        if (tid < nqb) {
            const int q = tid;
            const double qv = ql.q[q0 + q];
            const float inv = __fdiv_rn(1.0f, fx_scale);
            const float s = __fdiv_rn(100.0f, __fmul_rn(static_cast<float>(total), inv));     // utils.py:69,72
            const uint32_t wlo = s_wlo[q0 + q];         // window of the final level: slot == rank offset
            const uint32_t de = wlo + s_e[q];
            const uint32_t We = s_We[q], Ce = s_Ce[q];
            const float a_e = __ldg(p.y_sorted + (rmin + de));
            const float p_e = plotting_position(s, __fmul_rn(static_cast<float>(Ce), inv),
                                                __fmul_rn(static_cast<float>(We), inv));
            float result = a_e;
            if (static_cast<double>(p_e) < qv) {        // interval (e, succ)
                uint32_t ds = NONE, Ws = 0u;
                if (s_sin[q] != NONE) { ds = wlo + s_sin[q]; Ws = l2[q * L2S + s_sin[q]]; }
                else if (s_sout[q0 + q] != NONE) { ds = s_sout[q0 + q]; Ws = s_Ws[q]; }
                if (ds != NONE) {                       // else q lies beyond the last entry: a_last (utils.py:74-75)
                    const float a_s = __ldg(p.y_sorted + (rmin + ds));
                    const float p_s = plotting_position(s, __fmul_rn(static_cast<float>(Ce + Ws), inv),
                                                        __fmul_rn(static_cast<float>(Ws), inv));
                    const float frac = __fdiv_rn(__fsub_rn(static_cast<float>(qv), p_e), __fsub_rn(p_s, p_e));
                    result = __fadd_rn(a_e, __fmul_rn(frac, __fsub_rn(a_s, a_e)));
                }
            } else {                                    // interval (pred, e)
                uint32_t dp = NONE, Wp = 0u;
                if (s_pin[q] != 0u) { dp = wlo + s_pin[q] - 1u; Wp = l2[q * L2S + s_pin[q] - 1u]; }
                else if (s_pout[q0 + q] != 0u) { dp = s_pout[q0 + q] - 1u; Wp = s_Wp[q]; }
                if (dp != NONE) {                       // else q lies before the first entry: a_0 (utils.py:76-77)
                    const float a_p = __ldg(p.y_sorted + (rmin + dp));
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
