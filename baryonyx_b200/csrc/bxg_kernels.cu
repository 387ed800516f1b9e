// bxg_kernels.cu -- sm_100a kernels of the constraint sweep.
//
// One CTA per lane (= one independent Solver instance of the reference with
// its own P, pi, x, random engine and order list).  Rows of one lane are
// processed one after the other in the order list (Gauss-Seidel, like
// compute_update_row, lib/src/itm-solver-inequalities-01.cpp:258-296); the
// threads of the CTA cooperate inside a row.
//
// Layout decisions:
//  * P is stored in CSC order, so the reduced-cost gather for column j reads
//    one contiguous run P[col_ptr[j]..col_ptr[j+1]) and the row's own entry
//    (k, j) is part of that run: theta-decay, gather and the +-d update touch
//    the same sectors.
//  * pi, x and the reduced-cost scratch R of the lane live in shared memory
//    for the duration of a kernel when they fit (generic pointers otherwise).
//  * R is persistent per lane, like the reference's R member: the slot one
//    past the current row keeps whatever a previous longer row left there
//    (the reference reads it, itm-solver-inequalities-01.cpp:198-203).
#include "bxg_device.cuh"
#include "bxg_kernels.cuh"

#include <climits>
#include <math_constants.h>
#include <cstdio>

namespace bxg {

namespace {

struct Shared
{
    uint32_t rng;
    int selected;
    int pi_changed;
    int remaining;
    int list_len;
    int order_state;
    int tmp_i;
    int slow; // the current row takes the single-thread path
    int n_pre;
    double d;
    double tmp_d;
    double red[32];
    int redi[32];
};

struct Lane
{
    const DevProblem* pb;
    const DevParams* pr;
    double* P;
    double* pi;
    uint32_t* x;
    // reduced-cost scratch of the current row
    unsigned long long* key0; // [max_row]   sort keys in row order
    unsigned long long* sk;   // [sort_cap]  keys being sorted
    int* sid;                 // [sort_cap]  position in the row, moved with sk
    double* Rv;               // [max_row+2] value left at each slot of the reference's R
    double* pown;
    int* opos;
    int* viol;
    int* oR;
    int2* mo;
    LaneCtl* ctl;
    Shared* sh;
    int sort_cap; // the CTA-wide sort handles rows up to this length
    unsigned long long updates;
    unsigned long long sweeps;
    double alg_bytes;
};

__device__ __forceinline__ int
xbit(const uint32_t* x, int i)
{
    return (int)((x[i >> 5] >> (i & 31)) & 1u);
}

// ---- sort keys ------------------------------------------------------------
// The reduced costs of a row are ordered through 64-bit integer keys: the
// usual order-preserving map of a double (sign bit flipped for positives, all
// bits flipped for negatives), complemented when the objective is maximised
// so that ascending keys always give the reference's order
// (calculator_sort, itm-common.hpp:134-148).  Rows holding a NaN or a -0.0
// keep the raw bits instead and are compared as doubles by one thread
// ("raw" mode): there the reference's comparator is not a total order and
// only its own algorithm defines the result.
__device__ __forceinline__ unsigned long long
key_of(double v, bool maximize)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    return maximize ? ~b : b;
}

__device__ __forceinline__ double
value_of_key(unsigned long long k, bool maximize)
{
    if (maximize)
        k = ~k;
    k = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)k);
}

struct LessKey
{
    __device__ bool operator()(const KeyId& a, const KeyId& b) const { return a.k < b.k; }
};

struct LessRawMin
{
    __device__ bool operator()(const KeyId& a, const KeyId& b) const
    {
        return __longlong_as_double((long long)a.k) < __longlong_as_double((long long)b.k);
    }
};

struct LessRawMax
{
    __device__ bool operator()(const KeyId& a, const KeyId& b) const
    {
        return __longlong_as_double((long long)b.k) < __longlong_as_double((long long)a.k);
    }
};

struct PairLessDecr
{
    __device__ bool operator()(const int2& a, const int2& b) const { return b.y < a.y; }
};

struct PairLessIncr
{
    __device__ bool operator()(const int2& a, const int2& b) const { return a.y < b.y; }
};

struct PiLessDecr
{
    const double* pi;
    __device__ bool operator()(const int& a, const int& b) const { return pi[b] < pi[a]; }
};

struct PiLessIncr
{
    const double* pi;
    __device__ bool operator()(const int& a, const int& b) const { return pi[a] < pi[b]; }
};

// stop_iterating with the engine (itm-common.hpp:150-163)
__device__ __forceinline__ bool
stop_iterating_rng(double value, bool maximize, uint32_t& rng)
{
    if (value == 0)
        return bernoulli(rng, 0.5);
    return maximize ? value < 0 : value > 0;
}

// Bounds of the scan of select_variables (itm-solver-inequalities-01.cpp:
// 193-206, -101.cpp:235-248) for an inequality row.
__device__ __forceinline__ void
select_window(int family, int r_size, int c_size, int bkmin, int bkmax, int& lo, int& hi)
{
    if (family == FAM_01) {
        lo = bkmin < r_size ? bkmin : r_size;
        hi = bkmax < r_size ? bkmax : r_size;
    } else {
        lo = bkmin + c_size;
        hi = (bkmax + c_size) < r_size ? (bkmax + c_size) : r_size;
    }
}

// Scalar part of affect (itm-common.hpp:382-467): pi_k and the preference
// step d from the values around the selection boundary.  One thread.
__device__ __forceinline__ void
affect_scalars(Lane& L, int k, int selected, int r_size, double kappa, double v0, double vsel,
               double vnext)
{
    const double old_pi = L.pi[k];
    const double kk = __ddiv_rn(kappa, __dsub_rn(1.0, kappa));
    double d = L.pr->delta;
    if (selected < 0) {
        d = __dadd_rn(d, __dmul_rn(kk, __ddiv_rn(v0, 2.0)));
    } else if (selected + 1 >= r_size) {
        d = __dadd_rn(d, __dmul_rn(kk, __dmul_rn(vsel, 1.5)));
    } else {
        const double npi = __dadd_rn(old_pi, __ddiv_rn(__dadd_rn(vsel, vnext), 2.0));
        L.pi[k] = npi;
        d = __dadd_rn(d, __dmul_rn(kk, __dsub_rn(vnext, vsel)));
        if ((__double_as_longlong(old_pi) < 0) != (__double_as_longlong(npi) < 0))
            L.sh->pi_changed = 1;
    }
    L.sh->selected = selected;
    L.sh->d = d;
}

// Single-thread statement of calculator_sort + select_variables on the row:
// the exact libstdc++ sort and shuffle permutations.  Used when the parallel
// path cannot decide the row (ties across the selection boundary, NaN, -0.0).
template<typename Less>
__device__ void
row_select_exact(Lane& L, int k, int r_size, double kappa, bool raw, Less less)
{
    const DevProblem& pb = *L.pb;
    const bool maximize = pb.maximize != 0;
    uint32_t rng = L.sh->rng;
    Pairs a{ L.sk, L.sid };

    for (int i = 0; i != r_size; ++i) {
        L.sk[i] = L.key0[i];
        L.sid[i] = i;
    }

    auto value_at = [&](int p) {
        return raw ? __longlong_as_double((long long)L.sk[p]) : value_of_key(L.sk[p], maximize);
    };

    if (r_size > 1) {
        seq_std_sort(a, r_size, less);
        // random_shuffle_unique, itm-common.hpp:117-132
        int ret = 0;
        int i = 1;
        for (; i != r_size; ++i) {
            const bool differ = raw ? (value_at(ret) != value_at(i)) : (L.sk[ret] != L.sk[i]);
            if (differ) {
                seq_shuffle(a, ret, i - ret, rng);
                ret = i;
            }
        }
        seq_shuffle(a, ret, i - ret, rng);
    }

    const int c_size = pb.family == FAM_101 ? pb.row_neg[k] : 0;
    const int bkmin = pb.bmin[k], bkmax = pb.bmax[k];
    int selected;
    if (bkmin == bkmax) {
        const int b = bkmin + c_size;
        selected = (b < r_size ? b : r_size) - 1;
    } else {
        int lo, hi;
        select_window(pb.family, r_size, c_size, bkmin, bkmax, lo, hi);
        selected = hi - 1;
        for (int i = lo; i <= hi; ++i) {
            const double v = i < r_size ? value_at(i) : L.Rv[r_size];
            if (stop_iterating_rng(v, maximize, rng)) {
                selected = i - 1;
                break;
            }
        }
    }

    L.sh->rng = rng;
    const double v0 = value_at(0);
    const double vsel = selected >= 0 ? value_at(selected) : 0.0;
    const double vnext = (selected >= 0 && selected + 1 < r_size) ? value_at(selected + 1) : 0.0;
    affect_scalars(L, k, selected, r_size, kappa, v0, vsel, vnext);
}

// Bitonic sort of sk[0..n2) (ascending) carrying sid, by the whole CTA.  n2
// is a power of two.  Strides below 64 stay inside the 64 elements one warp
// owns, so only the wide strides need a CTA barrier.
__device__ void
cta_bitonic_sort(unsigned long long* sk, int* sid, int n2)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int half = n2 >> 1;
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < half; t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const unsigned long long a = sk[i], b = sk[l];
                const bool up = (i & kk) == 0;
                if ((a > b) == up && a != b) {
                    sk[i] = b;
                    sk[l] = a;
                    const int ia = sid[i];
                    sid[i] = sid[l];
                    sid[l] = ia;
                }
            }
            // the next stage works with stride j/2, or kk when the next merge
            // size starts
            const int nj = j > 1 ? (j >> 1) : kk;
            if (j > 32 || nj > 32)
                __syncthreads();
            else
                __syncwarp();
        }
    }
    __syncthreads();
}

// One constraint update (the loop body of compute_update_row /
// push_and_compute_update_row).  Called by all threads of the CTA.
__device__ void
update_row(Lane& L, int k, double kappa, int push, double obj_amp)
{
    const DevProblem& pb = *L.pb;
    const DevParams& pr = *L.pr;
    const int tid = threadIdx.x, T = blockDim.x;
    const int r0 = pb.row_ptr[k];
    const int r_size = pb.row_ptr[k + 1] - r0;
    const double theta = pr.theta;
    const double* __restrict__ cost = pr.cost;
    const bool maximize = pb.maximize != 0;

    if (tid == 0) {
        L.sh->slow = 0;
        L.alg_bytes += pb.row_bytes[k];
    }
    __syncthreads();

    // ---- A: theta decay + reduced costs (compute_reduced_costs) ----------
    int odd = 0;   // NaN or -0.0 seen by this thread
    int n_pre = 0; // values that neither stop the scan nor are zero
    for (int i = tid; i < r_size; i += T) {
        const int e = r0 + i;
        const int j = pb.row_col[e];
        const int c0 = pb.col_ptr[j], c1 = pb.col_ptr[j + 1];
        double pown = 0.0;
        int opos = c0;
        double val;

        if (pb.family == FAM_01) {
            double sum_pi = 0.0, sum_p = 0.0;
            for (int t = c0; t < c1; ++t) {
                const int rw = pb.csc_row[t];
                double p = L.P[t];
                if (rw == k) {
                    p = __dmul_rn(p, theta);
                    pown = p;
                    opos = t;
                }
                sum_pi = __dadd_rn(sum_pi, L.pi[rw]);
                sum_p = __dadd_rn(sum_p, p);
            }
            val = __dsub_rn(__dsub_rn(cost[j], sum_pi), sum_p);
            if (push)
                val = __dadd_rn(val, __dmul_rn(obj_amp, cost[i]));
        } else {
            double sum_pi = 0.0, sum_p = 0.0;
            for (int t = c0; t < c1; ++t) {
                const int rw = pb.csc_row[t];
                const double a = (double)pb.csc_a[t];
                double p = L.P[t];
                if (rw == k) {
                    p = __dmul_rn(p, theta);
                    pown = p;
                    opos = t;
                }
                sum_pi = __dadd_rn(sum_pi, __dmul_rn(a, L.pi[rw]));
                sum_p = __dadd_rn(sum_p, __dmul_rn(a, p));
            }
            val = __dsub_rn(__dsub_rn(cost[j], sum_pi), sum_p);
            if (pb.row_a[e] < 0)
                val = -val;
            if (push)
                val = __dadd_rn(val, __dmul_rn(obj_amp, cost[j]));
        }

        if (val != val || (val == 0.0 && __double_as_longlong(val) < 0))
            odd = 1;
        if (val != 0.0 && !(maximize ? val < 0 : val > 0))
            ++n_pre;
        // raw bits here; rewritten as keys below once the mode is known
        L.key0[i] = (unsigned long long)__double_as_longlong(val);
        L.pown[i] = pown;
        L.opos[i] = opos;
    }
    const bool raw = __syncthreads_or(odd) != 0;

    int n2 = 2;
    while (n2 < r_size)
        n2 <<= 1;

    if (raw || r_size < 2 || n2 > L.sort_cap) {
        if (!raw)
            for (int i = tid; i < r_size; i += T)
                L.key0[i] = key_of(__longlong_as_double((long long)L.key0[i]), maximize);
        __syncthreads();
        if (tid == 0)
            L.sh->slow = 1;
    } else {
        // ---- B: parallel order + selection ---------------------------------
        for (int i = tid; i < n2; i += T) {
            unsigned long long key = ~0ull;
            if (i < r_size) {
                key = key_of(__longlong_as_double((long long)L.key0[i]), maximize);
                L.key0[i] = key;
            }
            L.sk[i] = key;
            L.sid[i] = i;
        }
        // block count of n_pre
#pragma unroll
        for (int o = 16; o; o >>= 1)
            n_pre += __shfl_xor_sync(0xffffffffu, n_pre, o);
        if ((tid & 31) == 0)
            L.sh->redi[tid >> 5] = n_pre;
        __syncthreads();
        cta_bitonic_sort(L.sk, L.sid, n2);

        // equal keys next to each other = a run random_shuffle_unique shuffles
        int tie = 0;
        for (int p = tid + 1; p < r_size; p += T)
            tie |= (L.sk[p] == L.sk[p - 1]);
        const bool has_tie = __syncthreads_or(tie) != 0;
        const uint32_t rng0 = L.sh->rng;
        int total_draws = 0;
        int reject = 0;

        if (has_tie) {
            // Draws of all the std::shuffle calls, in sorted order.  A run of
            // n equal keys starting at a draws floor(n/2) numbers: the element
            // at odd offset o owns draw number (o-1)/2 (n odd) or o/2 (n even,
            // where offset 1 owns the single-swap draw).  Its index in the
            // row's draw sequence is the count of odd offsets before it.
            const int C = (r_size + T - 1) / T;
            const int p0 = min(r_size, tid * C), p1 = min(r_size, p0 + C);
            int local = 0;
            for (int p = p0; p < p1; ++p) {
                const unsigned long long key = L.sk[p];
                if (p > 0 && L.sk[p - 1] == key) {
                    int lo = 0, hi = p; // first index holding key
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (L.sk[mid] < key)
                            lo = mid + 1;
                        else
                            hi = mid;
                    }
                    local += (p - lo) & 1;
                }
            }
            // exclusive scan of `local` over the threads
            int incl = local;
            const int ln = tid & 31, wp = tid >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (ln >= o)
                    incl += v;
            }
            if (ln == 31)
                L.sh->red[wp] = (double)incl;
            __syncthreads();
            int before = 0;
            const int nw = T >> 5;
            for (int w = 0; w < nw; ++w) {
                const int v = (int)L.sh->red[w];
                if (w < wp)
                    before += v;
                total_draws += v;
            }
            int t = before + incl - local;

            for (int p = p0; p < p1; ++p) {
                const unsigned long long key = L.sk[p];
                const bool eq_prev = p > 0 && L.sk[p - 1] == key;
                const bool eq_next = p + 1 < r_size && L.sk[p + 1] == key;
                if (!eq_prev && !eq_next)
                    continue;
                int lo = 0, hi = p;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (L.sk[mid] < key)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                const int a = lo;
                lo = p;
                hi = r_size; // first index above key
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (L.sk[mid] <= key)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                const int n = lo - a;
                const int o = p - a;
                uint32_t uerange = 0;
                if (n & 1) {
                    if (o & 1)
                        uerange = (uint32_t)(o + 1) * (uint32_t)(o + 2);
                } else if (o == 1)
                    uerange = 2u;
                else if (o >= 2 && !(o & 1))
                    uerange = (uint32_t)(o + 1) * (uint32_t)(o + 2);
                if (uerange) {
                    // this draw is number t' = (count of odd offsets before p)
                    // for o odd that is t; for o even (n even, o >= 2) the
                    // odd offset o-1 < p is already counted in t as well
                    const uint32_t g = lcg_mul(rng0, lcg_pow((uint64_t)t + 1u));
                    const uint32_t scaling = 2147483645u / uerange;
                    if (g - 1u >= uerange * scaling)
                        reject = 1;
                }
                t += o & 1;
            }
        }
        const bool any_reject = has_tie && __syncthreads_or(reject) != 0;

        if (tid == 0) {
            uint32_t rng = rng0;
            if (any_reject) {
                // a uniform draw was rejected somewhere: replay the draws
                int ret = 0;
                int i = 1;
                for (; i != r_size; ++i) {
                    if (L.sk[ret] != L.sk[i]) {
                        seq_shuffle_draws(i - ret, rng);
                        ret = i;
                    }
                }
                seq_shuffle_draws(i - ret, rng);
            } else if (total_draws)
                rng = lcg_mul(rng0, lcg_pow((uint64_t)total_draws));

            int npre = 0;
            const int nw = T >> 5;
            for (int w = 0; w < nw; ++w)
                npre += L.sh->redi[w];

            const int c_size = pb.family == FAM_101 ? pb.row_neg[k] : 0;
            const int bkmin = pb.bmin[k], bkmax = pb.bmax[k];
            int selected;
            if (bkmin == bkmax) {
                const int b = bkmin + c_size;
                selected = (b < r_size ? b : r_size) - 1;
            } else {
                int lo, hi;
                select_window(pb.family, r_size, c_size, bkmin, bkmax, lo, hi);
                selected = hi - 1;
                // the first npre sorted values neither stop nor draw
                int i = lo;
                if (i < npre)
                    i = npre;
                for (; i <= hi; ++i) {
                    const double v = i < r_size ? value_of_key(L.sk[i], maximize) : L.Rv[r_size];
                    if (stop_iterating_rng(v, maximize, rng)) {
                        selected = i - 1;
                        break;
                    }
                }
            }

            if (selected >= 0 && selected + 1 < r_size && L.sk[selected] == L.sk[selected + 1]) {
                // equal values on both sides of the boundary: which of them
                // are selected is decided by the exact permutations
                L.sh->slow = 1;
            } else {
                L.sh->rng = rng;
                const double v0 = value_of_key(L.sk[0], maximize);
                const double vsel = selected >= 0 ? value_of_key(L.sk[selected], maximize) : 0.0;
                const double vnext =
                  (selected >= 0 && selected + 1 < r_size) ? value_of_key(L.sk[selected + 1], maximize) : 0.0;
                affect_scalars(L, k, selected, r_size, kappa, v0, vsel, vnext);
            }
        }
    }
    __syncthreads();

    if (L.sh->slow) {
        if (tid == 0) {
            if (!raw)
                row_select_exact(L, k, r_size, kappa, false, LessKey());
            else if (maximize)
                row_select_exact(L, k, r_size, kappa, true, LessRawMax());
            else
                row_select_exact(L, k, r_size, kappa, true, LessRawMin());
        }
        __syncthreads();
    }

    // ---- D: x / P part of affect ------------------------------------------
    {
        const int selected = L.sh->selected;
        const double d = L.sh->d;
        const bool all_out = selected < 0;
        const bool all_in = !all_out && (selected + 1 >= r_size);
        const bool keep_tail = pb.tail_rows != 0;

        for (int p = tid; p < r_size; p += T) {
            const int id = L.sid[p];
            const bool in = all_in || (!all_out && p <= selected);
            const bool neg = pb.family != FAM_01 && pb.row_a[r0 + id] < 0;
            const bool set = in != neg;
            const int j = pb.row_col[r0 + id];
            const double pv = L.pown[id];
            L.P[L.opos[id]] = set ? __dadd_rn(pv, d) : __dsub_rn(pv, d);
            if (set)
                atomicOr(&L.x[j >> 5], 1u << (j & 31));
            else
                atomicAnd(&L.x[j >> 5], ~(1u << (j & 31)));
            if (keep_tail)
                L.Rv[p] = raw ? __longlong_as_double((long long)L.sk[p]) : value_of_key(L.sk[p], maximize);
        }
    }
    __syncthreads();
}

// Row activities and the violated list (compute_violated_constraints,
// itm-common.hpp:102-115; infeasibility_local_compute_violated_constraints,
// :893-914; local_compute_violated_constraints, :882-891).  mode: 0 = list of
// violated rows into oR, 1 = (k, amount) pairs into mo, 2 = count only.
__device__ int
violated_scan(Lane& L, int mode)
{
    const DevProblem& pb = *L.pb;
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

    for (int k = warp; k < pb.m; k += nwarps) {
        const int e0 = pb.row_ptr[k], e1 = pb.row_ptr[k + 1];
        int v = 0;
        if (pb.family == FAM_01) {
            for (int e = e0 + lane; e < e1; e += 32)
                v += xbit(L.x, pb.row_col[e]);
        } else {
            for (int e = e0 + lane; e < e1; e += 32)
                v += pb.row_a[e] * xbit(L.x, pb.row_col[e]);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            const int lo = pb.bmin[k], hi = pb.bmax[k];
            L.viol[k] = lo > v ? lo - v : (hi < v ? v - hi : 0);
        }
    }
    __syncthreads();

    // ordered compaction by warp 0
    if (warp == 0) {
        int count = 0;
        for (int base = 0; base < pb.m; base += 32) {
            const int k = base + lane;
            const int amount = k < pb.m ? L.viol[k] : 0;
            const unsigned mask = __ballot_sync(0xffffffffu, amount != 0);
            if (amount != 0) {
                const int pos = count + __popc(mask & ((1u << lane) - 1u));
                if (mode == 0)
                    L.oR[pos] = k;
                else if (mode == 1)
                    L.mo[pos] = make_int2(k, amount);
            }
            count += __popc(mask);
        }
        if (lane == 0) {
            L.sh->remaining = count;
            if (mode == 0)
                L.sh->list_len = count;
            else if (mode == 1)
                L.sh->tmp_i = count;
        }
    }
    __syncthreads();
    if (tid == 0)
        L.alg_bytes += 4.0 * pb.nnz * (pb.family == FAM_01 ? 1.0 : 2.0) + 8.0 * pb.m + pb.n / 8.0;
    return L.sh->remaining;
}

// Lane-local list lengths live in registers of every thread (uniform).
struct Lists
{
    int oR_len;
    int mo_len;
};

// compute_order::init, itm-common.hpp:660-680
__device__ void
order_init(Lane& L, Lists& ls)
{
    const int order = L.sh->order_state;
    if (order == ORDER_INFEASIBILITY_DECR || order == ORDER_INFEASIBILITY_INCR) {
        violated_scan(L, 1);
        ls.mo_len = L.sh->tmp_i;
    } else if (order == ORDER_PI_SIGN_CHANGE) {
        for (int i = threadIdx.x; i < ls.oR_len; i += blockDim.x)
            L.oR[i] = i;
        __syncthreads();
    } else {
        violated_scan(L, 0);
        ls.oR_len = L.sh->list_len;
    }
}

// compute_order::run / push_and_run, itm-common.hpp:682-880
__device__ int
order_run(Lane& L, Lists& ls, double kappa, int push, double obj_amp)
{
    const int tid = threadIdx.x;

    if (push && L.pr->order == ORDER_CYCLE) {
        if (tid == 0) {
            const int next = L.sh->order_state + 1;
            L.sh->order_state = next >= ORDER_CYCLE ? 0 : next;
        }
        __syncthreads();
    }
    const int order = L.sh->order_state;
    L.sweeps += 1;

    switch (order) {
    case ORDER_REVERSING:
        L.updates += ls.oR_len;
        for (int i = ls.oR_len - 1; i >= 0; --i)
            update_row(L, L.oR[i], kappa, push, obj_amp);
        violated_scan(L, 0);
        ls.oR_len = L.sh->list_len;
        return L.sh->remaining;
    case ORDER_RANDOM_SORTING:
        if (tid == 0) {
            uint32_t rng = L.sh->rng;
            seq_shuffle(Plain<int>{ L.oR }, 0, ls.oR_len, rng);
            L.sh->rng = rng;
        }
        __syncthreads();
        L.updates += ls.oR_len;
        for (int i = 0; i < ls.oR_len; ++i)
            update_row(L, L.oR[i], kappa, push, obj_amp);
        violated_scan(L, 0);
        ls.oR_len = L.sh->list_len;
        return L.sh->remaining;
    case ORDER_INFEASIBILITY_DECR:
    case ORDER_INFEASIBILITY_INCR:
        if (tid == 0) {
            if (order == ORDER_INFEASIBILITY_DECR)
                seq_std_sort(Plain<int2>{ L.mo }, ls.mo_len, PairLessDecr());
            else
                seq_std_sort(Plain<int2>{ L.mo }, ls.mo_len, PairLessIncr());
        }
        __syncthreads();
        L.updates += ls.mo_len;
        for (int i = 0; i < ls.mo_len; ++i)
            update_row(L, L.mo[i].x, kappa, push, obj_amp);
        violated_scan(L, 1);
        ls.mo_len = L.sh->tmp_i;
        return L.sh->remaining;
    case ORDER_LAGRANGIAN_DECR:
    case ORDER_LAGRANGIAN_INCR:
        if (tid == 0) {
            if (order == ORDER_LAGRANGIAN_DECR)
                seq_std_sort(Plain<int>{ L.oR }, ls.oR_len, PiLessDecr{ L.pi });
            else
                seq_std_sort(Plain<int>{ L.oR }, ls.oR_len, PiLessIncr{ L.pi });
        }
        __syncthreads();
        L.updates += ls.oR_len;
        for (int i = 0; i < ls.oR_len; ++i)
            update_row(L, L.oR[i], kappa, push, obj_amp);
        violated_scan(L, 0);
        ls.oR_len = L.sh->list_len;
        return L.sh->remaining;
    case ORDER_PI_SIGN_CHANGE: {
        if (tid == 0) {
            uint32_t rng = L.sh->rng;
            seq_shuffle(Plain<int>{ L.oR }, 0, ls.oR_len, rng);
            L.sh->rng = rng;
            L.sh->pi_changed = 0;
        }
        __syncthreads();
        L.updates += ls.oR_len;
        for (int i = 0; i < ls.oR_len; ++i)
            update_row(L, L.oR[i], kappa, push, obj_amp);
        const int remaining = violated_scan(L, 2);
        // "if (!pi_changed && remaining == 0) return 0; return remaining;"
        return remaining;
    }
    case ORDER_NONE:
    default:
        L.updates += ls.oR_len;
        for (int i = 0; i < ls.oR_len; ++i)
            update_row(L, L.oR[i], kappa, push, obj_amp);
        violated_scan(L, 0);
        ls.oR_len = L.sh->list_len;
        return L.sh->remaining;
    }
}

// default_cost_type::results, itm-common.hpp:1140-1147.  Result in
// sh->tmp_d, valid for all threads after the call.
__device__ double
objective(Lane& L)
{
    const DevProblem& pb = *L.pb;
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

    if (pb.obj_any_order) {
        double s = 0.0;
        for (int i = tid; i < pb.n_obj; i += T)
            if (xbit(L.x, pb.obj_var[i]))
                s += pb.obj_factor[i];
#pragma unroll
        for (int o = 16; o; o >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0)
            L.sh->red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            double t = pb.cost_constant;
            for (int w = 0; w < nwarps; ++w)
                t += L.sh->red[w];
            L.sh->tmp_d = t;
        }
        __syncthreads();
        return L.sh->tmp_d;
    }

    // general costs: the reference adds in objective-element order
    if (tid == 0) {
        double t = pb.cost_constant;
        for (int i = 0; i < pb.n_obj; ++i)
            if (xbit(L.x, pb.obj_var[i]))
                t = __dadd_rn(t, pb.obj_factor[i]);
        L.sh->tmp_d = t;
    }
    __syncthreads();
    return L.sh->tmp_d;
}

// ---------------------------------------------------------------------------
// lane set-up / tear-down inside a kernel
// ---------------------------------------------------------------------------
struct SmemPlan
{
    int pi_in_smem, x_in_smem, r_in_smem;
    int sort_cap;
};

// Bytes of the per-lane row scratch (key0, sk, sid, pown, opos), laid out the
// same way in shared memory or in the batch's global scratch.
__host__ __device__ inline size_t
row_scratch_bytes(int max_row, int sort_cap)
{
    size_t n = 0;
    n += sizeof(unsigned long long) * (size_t)max_row;  // key0
    n += sizeof(unsigned long long) * (size_t)sort_cap; // sk
    n += sizeof(double) * (size_t)(max_row + 1);        // pown
    n += sizeof(int) * (size_t)sort_cap;                // sid
    n += sizeof(int) * (size_t)(max_row + 1);           // opos
    return (n + 15) & ~size_t(15);
}

__device__ void
lane_open(Lane& L, Lists& ls, const DevProblem& pb, const DevParams& pr,
          const DevBatch& b, int lane_id, SmemPlan plan, unsigned char* smem)
{
    const int tid = threadIdx.x, T = blockDim.x;
    size_t off = 0;

    L.pb = &pb;
    L.pr = &pr;
    L.sh = reinterpret_cast<Shared*>(smem);
    off += (sizeof(Shared) + 15) & ~size_t(15);

    L.P = b.P + (size_t)lane_id * pb.nnz;
    L.ctl = b.ctl + lane_id;
    L.oR = b.oR + (size_t)lane_id * b.order_cap;
    L.mo = b.mo + (size_t)lane_id * b.order_cap;
    L.sort_cap = plan.r_in_smem ? plan.sort_cap : 0;

    double* g_pi = b.pi + (size_t)lane_id * pb.m;
    uint32_t* g_x = b.x + (size_t)lane_id * pb.xw32;
    double* g_Rv = b.Rv + (size_t)lane_id * (pb.max_row + 2);

    {
        unsigned char* base;
        if (plan.r_in_smem) {
            base = smem + off;
            off += row_scratch_bytes(pb.max_row, plan.sort_cap);
        } else
            base = b.g_scratch + (size_t)lane_id * row_scratch_bytes(pb.max_row, plan.sort_cap);
        size_t o = 0;
        L.key0 = reinterpret_cast<unsigned long long*>(base + o);
        o += sizeof(unsigned long long) * (size_t)pb.max_row;
        L.sk = reinterpret_cast<unsigned long long*>(base + o);
        o += sizeof(unsigned long long) * (size_t)plan.sort_cap;
        L.pown = reinterpret_cast<double*>(base + o);
        o += sizeof(double) * (size_t)(pb.max_row + 1);
        L.sid = reinterpret_cast<int*>(base + o);
        o += sizeof(int) * (size_t)plan.sort_cap;
        L.opos = reinterpret_cast<int*>(base + o);
    }

    if (plan.r_in_smem) {
        L.Rv = reinterpret_cast<double*>(smem + off);
        off += sizeof(double) * (size_t)(pb.max_row + 2);
        for (int i = tid; i < pb.max_row + 2; i += T)
            L.Rv[i] = g_Rv[i];
    } else
        L.Rv = g_Rv;

    if (plan.pi_in_smem) {
        L.pi = reinterpret_cast<double*>(smem + off);
        off += sizeof(double) * (size_t)pb.m;
        for (int i = tid; i < pb.m; i += T)
            L.pi[i] = g_pi[i];
    } else
        L.pi = g_pi;

    L.viol = reinterpret_cast<int*>(smem + off);
    off += sizeof(int) * (size_t)pb.m;

    if (plan.x_in_smem) {
        L.x = reinterpret_cast<uint32_t*>(smem + off);
        off += sizeof(uint32_t) * (size_t)pb.xw32;
        for (int i = tid; i < pb.xw32; i += T)
            L.x[i] = g_x[i];
    } else
        L.x = g_x;

    if (tid == 0) {
        L.sh->rng = L.ctl->rng;
        L.sh->order_state = L.ctl->order_state;
        L.sh->pi_changed = 0;
        L.sh->remaining = L.ctl->remaining;
    }
    ls.oR_len = L.ctl->oR_len;
    ls.mo_len = L.ctl->mo_len;
    L.updates = 0;
    L.sweeps = 0;
    L.alg_bytes = 0.0;
    __syncthreads();
}

__device__ void
lane_close(Lane& L, Lists& ls, const DevProblem& pb, const DevBatch& b,
           int lane_id, SmemPlan plan)
{
    const int tid = threadIdx.x, T = blockDim.x;
    __syncthreads();

    if (plan.r_in_smem) {
        double* g_Rv = b.Rv + (size_t)lane_id * (pb.max_row + 2);
        for (int i = tid; i < pb.max_row + 2; i += T)
            g_Rv[i] = L.Rv[i];
    }
    if (plan.pi_in_smem) {
        double* g_pi = b.pi + (size_t)lane_id * pb.m;
        for (int i = tid; i < pb.m; i += T)
            g_pi[i] = L.pi[i];
    }
    if (plan.x_in_smem) {
        uint32_t* g_x = b.x + (size_t)lane_id * pb.xw32;
        for (int i = tid; i < pb.xw32; i += T)
            g_x[i] = L.x[i];
    }
    if (tid == 0) {
        L.ctl->rng = L.sh->rng;
        L.ctl->order_state = L.sh->order_state;
        L.ctl->oR_len = ls.oR_len;
        L.ctl->mo_len = ls.mo_len;
        L.ctl->remaining = L.sh->remaining;
        L.ctl->updates += L.updates;
        L.ctl->sweeps += L.sweeps;
        L.ctl->alg_bytes += L.alg_bytes;
    }
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char g_smem[];

// Start of a run: engine, solver state, x (init policy bastert + Bernoulli
// flips on the device, or a caller-provided x0), compute_order::init.
// solver_functor::operator(), itm-solver-common.hpp:94-133.
__global__ void
k_begin(DevProblem pb, DevParams pr, DevBatch b, const uint32_t* seeds,
        const uint32_t* x0, const uint32_t* rng0, int reset_solver,
        int init_on_device, double init_random, const double* cost_orig,
        SmemPlan plan)
{
    const int lane_id = blockIdx.x;
    const int tid = threadIdx.x, T = blockDim.x;

    double* g_pi = b.pi + (size_t)lane_id * pb.m;
    uint32_t* g_x = b.x + (size_t)lane_id * pb.xw32;
    double* g_Rv = b.Rv + (size_t)lane_id * (pb.max_row + 2);
    double* g_P = b.P + (size_t)lane_id * pb.nnz;
    LaneCtl* ctl = b.ctl + lane_id;

    if (reset_solver) {
        for (int i = tid; i < pb.nnz; i += T)
            g_P[i] = 0.0;
        for (int i = tid; i < pb.m; i += T)
            g_pi[i] = 0.0;
        for (int i = tid; i < pb.max_row + 2; i += T)
            g_Rv[i] = (i == pb.max_row) ? pr.tail_value : 0.0;
    }

    uint32_t rng;
    if (init_on_device) {
        // linear_congruential_engine::seed: 0 (mod m) becomes 1
        uint32_t s0 = seeds[lane_id] % 2147483647u;
        if (s0 == 0)
            s0 = 1;
        // init_with_bastert on the ORIGINAL costs (itm-common.hpp:255-267)
        // then one Bernoulli(init_random) per variable (itm-solver-common.hpp
        // :120-122).  Variable i consumes engine steps 2i+1 and 2i+2, so the
        // stream is split over threads with the LCG jump-ahead.
        for (int w = tid; w < pb.xw32; w += T) {
            uint32_t word = 0;
            const int i0 = w * 32;
            if (i0 < pb.n) {
                uint32_t cur = lcg_mul(s0, lcg_pow(2ull * (uint64_t)i0));
                const int i1 = min(pb.n, i0 + 32);
                for (int i = i0; i < i1; ++i) {
                    const double c = cost_orig[i];
                    // init_x<Mode>(cost, value_if_cost_0 = 0), itm-common.hpp:202-221
                    bool bit = pb.maximize ? (c > 0) : (c < 0);
                    const uint32_t g1 = lcg_next(cur);
                    const uint32_t g2 = lcg_next(cur);
                    if (canonical_from(g1, g2) < init_random)
                        bit = !bit;
                    word |= (uint32_t)bit << (i - i0);
                }
            }
            g_x[w] = word;
        }
        rng = lcg_mul(s0, lcg_pow(2ull * (uint64_t)pb.n));
    } else {
        const uint32_t* src = x0 + (size_t)lane_id * pb.xw32;
        for (int w = tid; w < pb.xw32; w += T)
            g_x[w] = src[w];
        rng = rng0[lane_id];
    }

    if (tid == 0) {
        ctl->rng = rng;
        ctl->order_state = pr.order == ORDER_CYCLE ? ORDER_NONE : pr.order;
        // compute_order ctor: R and m_order start with `variables` zero
        // entries (itm-common.hpp:634-642); only pi_sign_change and cycle
        // ever look at them before the first scan.
        ctl->oR_len = (pr.order == ORDER_PI_SIGN_CHANGE || pr.order == ORDER_CYCLE) ? pb.n : 0;
        ctl->mo_len = (pr.order == ORDER_CYCLE) ? pb.n : 0;
        ctl->remaining = 0;
        ctl->status = 2; // uninitialized
        ctl->best_remaining = INT_MAX;
        ctl->best_has_x = 0;
        ctl->best_loop = 0;
        ctl->best_value = pb.maximize ? -CUDART_INF : CUDART_INF;
    }
    if (pr.order == ORDER_CYCLE) {
        int* oR = b.oR + (size_t)lane_id * b.order_cap;
        int2* mo = b.mo + (size_t)lane_id * b.order_cap;
        for (int i = tid; i < pb.n; i += T) {
            oR[i] = 0;
            mo[i] = make_int2(0, 0);
        }
    }
    __syncthreads();

    Lane L;
    Lists ls;
    lane_open(L, ls, pb, pr, b, lane_id, plan, g_smem);
    order_init(L, ls);
    lane_close(L, ls, pb, b, lane_id, plan);
}

// compute_order::init on a range of lanes (after bxg_batch_set_state).
__global__ void
k_order_init(DevProblem pb, DevParams pr, DevBatch b, int lane_first,
             SmemPlan plan)
{
    const int lane_id = lane_first + blockIdx.x;
    Lane L;
    Lists ls;
    lane_open(L, ls, pb, pr, b, lane_id, plan, g_smem);
    order_init(L, ls);
    lane_close(L, ls, pb, b, lane_id, plan);
}

// One compute_order::run / push_and_run per lane.
__global__ void
k_sweep(DevProblem pb, DevParams pr, DevBatch b, int push, double obj_amp,
        SmemPlan plan)
{
    const int lane_id = blockIdx.x;
    Lane L;
    Lists ls;
    lane_open(L, ls, pb, pr, b, lane_id, plan, g_smem);
    const double kappa = b.kappa[lane_id];
    order_run(L, ls, kappa, push, obj_amp);
    lane_close(L, ls, pb, b, lane_id, plan);
}

__global__ void
k_objective(DevProblem pb, DevParams pr, DevBatch b, SmemPlan plan)
{
    const int lane_id = blockIdx.x;
    Lane L;
    Lists ls;
    lane_open(L, ls, pb, pr, b, lane_id, plan, g_smem);
    const double v = objective(L);
    if (threadIdx.x == 0)
        L.ctl->objective = v;
    lane_close(L, ls, pb, b, lane_id, plan);
}

// store_if_better(x, int remaining, i) / (x, double value, i),
// itm-solver-common.hpp:242-261.  Uniform branch; all threads copy.
__device__ void
store_best(Lane& L, const DevBatch& b, int lane_id, bool by_value, double value,
           int remaining, long long loop, int& best_remaining, double& best_value,
           long long& best_loop, int& best_has_x)
{
    const DevProblem& pb = *L.pb;
    bool take;
    if (by_value)
        take = pb.maximize ? value > best_value : value < best_value;
    else
        take = best_remaining > remaining;
    if (!take)
        return;

    uint32_t* g_best = b.bestx + (size_t)lane_id * pb.xw32;
    for (int i = threadIdx.x; i < pb.xw32; i += blockDim.x)
        g_best[i] = L.x[i];
    best_has_x = 1;
    best_loop = loop;
    if (by_value) {
        best_remaining = 0;
        best_value = value;
    } else
        best_remaining = remaining;
}

// Whole solver_functor::operator() loop (itm-solver-common.hpp:135-225) for
// one lane; the wall-clock time limit is handled by the host (limit sweeps).
__global__ void
k_solve(DevProblem pb, DevParams pr, DevBatch b, SmemPlan plan)
{
    const int lane_id = blockIdx.x;
    Lane L;
    Lists ls;
    lane_open(L, ls, pb, pr, b, lane_id, plan, g_smem);

    int status = 2;
    int m_best_remaining = INT_MAX; // m_best.remaining_constraints
    double m_best_value = pb.maximize ? -CUDART_INF : CUDART_INF;
    long long m_best_loop = 0;
    int m_best_has_x = 0;
    int best_remaining = INT_MAX;

    const long w_limit = (long)pr.w;
    bool start_push = false;
    double kappa = pr.kappa_min;

    for (long long i = 0; i != pr.limit; ++i) {
        const int remaining = order_run(L, ls, kappa, 0, 0.0);

        if (remaining == 0) {
            const double v = objective(L);
            store_best(L, b, lane_id, true, v, 0, i, m_best_remaining, m_best_value,
                       m_best_loop, m_best_has_x);
            best_remaining = remaining;
            start_push = true;
            break;
        }

        if (remaining < best_remaining) {
            store_best(L, b, lane_id, false, 0.0, remaining, i, m_best_remaining,
                       m_best_value, m_best_loop, m_best_has_x);
            best_remaining = remaining;
        }

        if (i > w_limit)
            kappa = __dadd_rn(kappa, __dmul_rn(pr.kappa_step, pr.powtab[remaining]));

        if (kappa > pr.kappa_max) {
            status = 4; // kappa_max_reached
            break;
        }
    }

    if (!start_push) {
        status = 5; // limit_reached
    } else {
        for (int push = 0; push < pr.pushes_limit; ++push) {
            int remaining =
              order_run(L, ls, __dmul_rn(pr.pushing_k_factor, kappa), 1,
                        pr.pushing_objective_amplifier);

            if (remaining == 0) {
                const double v = objective(L);
                store_best(L, b, lane_id, true, v, 0,
                           -(long long)push * pr.pushing_iteration_limit - 1,
                           m_best_remaining, m_best_value, m_best_loop, m_best_has_x);
            }

            for (int iter = 0; iter < pr.pushing_iteration_limit; ++iter) {
                remaining = order_run(L, ls, kappa, 0, 0.0);

                if (remaining == 0) {
                    const double v = objective(L);
                    store_best(L, b, lane_id, true, v, 0,
                               -(long long)push * pr.pushing_iteration_limit - iter - 1,
                               m_best_remaining, m_best_value, m_best_loop, m_best_has_x);
                    break;
                }

                if ((double)iter > pr.w)
                    kappa = __dadd_rn(kappa, __dmul_rn(pr.kappa_step, pr.powtab[remaining]));

                if (kappa > pr.kappa_max)
                    break;
            }
        }
    }

    if (m_best_remaining == 0)
        status = 0;

    if (threadIdx.x == 0) {
        L.ctl->status = status;
        L.ctl->best_remaining = m_best_remaining;
        L.ctl->best_has_x = m_best_has_x;
        L.ctl->best_loop = m_best_loop;
        L.ctl->best_value = m_best_value;
    }
    lane_close(L, ls, pb, b, lane_id, plan);
}

} // namespace

// ---------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------
static int
pow2_at_least(int v)
{
    int p = 2;
    while (p < v)
        p <<= 1;
    return p;
}

static size_t
smem_bytes(const DevProblem& pb, int pi_in, int x_in, int r_in, int sort_cap)
{
    size_t off = (sizeof(Shared) + 15) & ~size_t(15);
    if (r_in) {
        off += row_scratch_bytes(pb.max_row, sort_cap);
        off += sizeof(double) * (size_t)(pb.max_row + 2);
    }
    if (pi_in)
        off += sizeof(double) * (size_t)pb.m;
    off += sizeof(int) * (size_t)pb.m;
    if (x_in)
        off += sizeof(uint32_t) * (size_t)pb.xw32;
    return off + 16;
}

size_t
lane_scratch_bytes(const DevProblem& pb, const LaunchCfg& cfg)
{
    return row_scratch_bytes(pb.max_row, cfg.sort_cap);
}

LaunchCfg
plan_launch(const DevProblem& pb, int device)
{
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (max_optin <= 0)
        max_optin = 48 * 1024;
    const size_t budget = (size_t)max_optin - 1024;

    LaunchCfg cfg;
    cfg.r_in_smem = 1;
    cfg.pi_in_smem = 1;
    cfg.x_in_smem = 1;
    cfg.sort_cap = pow2_at_least(pb.max_row);
    // drop the largest consumers first until the plan fits
    auto fits = [&]() {
        return smem_bytes(pb, cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap) <= budget;
    };
    if (!fits())
        cfg.x_in_smem = 0;
    if (!fits()) {
        // rows too long for the shared-memory sort: scratch in global memory,
        // single-thread ordering only
        cfg.r_in_smem = 0;
        cfg.sort_cap = pb.max_row;
    }
    if (!fits())
        cfg.pi_in_smem = 0;
    cfg.smem = smem_bytes(pb, cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap);

    int t = 64;
    while (t < pb.max_row && t < 1024)
        t *= 2;
    if (t < 128)
        t = 128;
    cfg.threads = t;
    return cfg;
}

template<typename K>
static void
set_smem(K kernel, size_t smem)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

void
launch_begin(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
             const uint32_t* d_seeds, const uint32_t* d_x0, const uint32_t* d_rng,
             int reset_solver, int init_on_device, double init_random,
             const double* d_cost_orig, const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap };
    set_smem(k_begin, cfg.smem);
    k_begin<<<b.n_lanes, cfg.threads, cfg.smem, s>>>(pb, pr, b, d_seeds, d_x0, d_rng,
                                                      reset_solver, init_on_device,
                                                      init_random, d_cost_orig, plan);
}

void
launch_order_init(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
                  int lane_first, int lane_count, const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap };
    set_smem(k_order_init, cfg.smem);
    k_order_init<<<lane_count, cfg.threads, cfg.smem, s>>>(pb, pr, b, lane_first, plan);
}

void
launch_sweep(const DevProblem& pb, const DevParams& pr, const DevBatch& b, int push,
             double obj_amp, const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap };
    set_smem(k_sweep, cfg.smem);
    k_sweep<<<b.n_lanes, cfg.threads, cfg.smem, s>>>(pb, pr, b, push, obj_amp, plan);
}

void
launch_solve(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
             const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap };
    set_smem(k_solve, cfg.smem);
    k_solve<<<b.n_lanes, cfg.threads, cfg.smem, s>>>(pb, pr, b, plan);
}

void
launch_objective(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
                 const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap };
    set_smem(k_objective, cfg.smem);
    k_objective<<<b.n_lanes, cfg.threads, cfg.smem, s>>>(pb, pr, b, plan);
}

} // namespace bxg
