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
//
// This file is compiled once per solver family (-DBXG_FAM=0|1|2: 01, 101, Z
// coefficients): every family gets its own kernels, so the register budget of
// the 01 sweep is not set by the branch-and-bound code of the Z rows, and the
// three translation units build in parallel.
#ifndef BXG_FAM
#error "compile with -DBXG_FAM=0, 1 or 2 (see baryonyx_b200/build.py)"
#endif

#include "bxg_device.cuh"
#include "bxg_kernels.cuh"
#if BXG_FAM == 2
#include "bxg_zsolver.cuh"
#endif

#include <climits>
#include <math_constants.h>
#include <cstdio>
#include <cstdlib>

namespace bxg {

namespace {

constexpr int kFam = BXG_FAM;

#ifndef BXG_GATHER_UNROLL
#define BXG_GATHER_UNROLL 4
#endif
#ifndef BXG_CLASS_160
#define BXG_CLASS_160 0
#endif
#ifndef BXG_TIMERS
#define BXG_TIMERS 0 // -DBXG_TIMERS=1: per-path / per-phase cycle counters (profiling builds)
#endif

// The CTA's dynamic shared memory: [Lane][Shared][row scratch][R values][pi]
// [activities][x].  Views built from this symbol keep their address space, so
// the hot path compiles to LDS/STS/ATOMS instead of generic accesses.
extern __shared__ __align__(16) unsigned char g_smem[];

struct Shared
{
    SortQueue sq; // task stack of cta_std_sort
    uint32_t rng;
    int selected;
    int pi_changed;
    int remaining;
    int list_len;
    int order_state;
    int tmp_i;
    int n_dup; // equal keys listed by the tie table of the current row (update_row, mode 4)
    int slow; // the current row goes to the single-thread path
    int stop;    // the deadline has passed (time_limit / stop_task)
    int x_dirty; // a bit of x changed since x_value / x_hash were computed
    double x_value;            // default_cost_type::results of x
    unsigned long long x_hash; // bit_array_hash of x
    double d;
    double kk; // kappa / (1 - kappa) of the current sweep
    double tmp_d;
    double red[32];
    int redi[32];
    unsigned long long redk[32][2];
};

struct Lane;
__host__ __device__ constexpr size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

struct Lane
{
    const DevProblem* pb;
    const DevParams* pr;
    // what a constraint update reads of pb / pr, copied here so the hot loop
    // finds it with one shared-memory load
    const RowHdr* row_hdr;
    const int4* row_info;
    const int* row_eptr;
    const int2* gp;
    const int* row_a;
    const int* csc_rs;
    const int* csc_a;
    const double* cost;
    double theta, delta;          // of this lane (bxg_batch_set_lane_params)
    double kappa_min, kappa_step; // idem
    int maximize, tail_rows;
    double* P;
    double* pi;
    uint32_t* x;
    int* act; // [m] row activities sum a_kj x_j, kept up to date by the updates
    // reduced-cost scratch of the current row
    unsigned long long* key0; // [max_row]   value bits / sort keys in row order
    unsigned long long* sk;   // [sort_cap]  keys being sorted; doubles as the tie table
    int* sid;                 // [sort_cap]  position in the row, moved with sk
    double* Rv;               // [max_row+2] value left at each slot of the reference's R
    double* pown;             // [max_row+1] theta-decayed P of the row's own entries
    unsigned long long* xk;   // [max_row] workspace of the exact ordering in shared memory, or nullptr
    int* xid;                 // [max_row]
    int xcap;                 // elements xk / xid hold
    int xchunk;               // the largest power of two <= xcap: chunk of cta_bitonic_chunked
    int tie_streak;           // consecutive tie-table passes over long rows that found equal values
    double* stage;            // [stage_cap] {P term, pi term} of the entries being summed
    int stage_cap;
    int* oR;
    int2* mo;
    // the (row, amount) list of the infeasibility orders packed amount << 16 | row
    // in shared memory when the plan says so (one thread sorts it every sweep)
    uint32_t* mop;
    LaneCtl* ctl;
    Shared* sh;
    unsigned char* zs; // (Z) scratch of the sub-solvers, global memory
    int all_smem;  // pi, x and the row scratch are all in shared memory
    // byte offsets into g_smem of the views that live there (all_smem)
    int o_stage, o_key0, o_sk, o_pown, o_pi, o_act, o_x;
    int sort_cap;  // the CTA-wide sort handles rows up to this length (0 = none)
    int tab_bits;  // log2 of the tie-table size (32-bit entries laid over sk)
    int tab_gen;   // generation stamp of the tie table (uniform over the CTA)
    int tab_dirty; // the table must be cleared before its next use
    unsigned long long updates;
    unsigned long long sweeps;
    unsigned long long offers, inserts;
    double alg_bytes;
    // updates by selection path: 0 exact single thread, 1 CTA sort, 2 two
    // smallest, 3 boundary by sign; and the cycles thread 0 spent in each
    unsigned int n_path[4];
    unsigned long long cyc_path[4];
    // cycles of thread 0 by phase of the unordered paths: gather, sums,
    // ties + candidates, affect (each up to its closing barrier)
    unsigned long long cyc_phase[4];
};

constexpr size_t kLaneBytes = align16(sizeof(Lane));

// A lane view: the shared-memory window when the plan put it there (SM), the
// generic pointer otherwise.
template<bool SM, typename T>
__device__ __forceinline__ T*
view(T* generic, int offset)
{
    return SM ? reinterpret_cast<T*>(g_smem + offset) : generic;
}

__device__ __forceinline__ Lane&
the_lane()
{
    return *reinterpret_cast<Lane*>(g_smem);
}

__device__ __forceinline__ Shared*
the_shared()
{
    return reinterpret_cast<Shared*>(g_smem + kLaneBytes);
}

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

struct PackedLessDecr
{
    __device__ bool operator()(const uint32_t& a, const uint32_t& b) const { return (b >> 16) < (a >> 16); }
};

struct PackedLessIncr
{
    __device__ bool operator()(const uint32_t& a, const uint32_t& b) const { return (a >> 16) < (b >> 16); }
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

// Scalar part of affect (itm-common.hpp:382-467): the preference step d and
// the new pi_k from the values around the selection boundary.  Pure.
__device__ __forceinline__ double
affect_step(double delta, double kk, double old_pi, int selected, int r_size, double v0,
            double vsel, double vnext, double& new_pi, bool& pi_moves)
{
    // kk = kappa / (1 - kappa), once per sweep; x / 2 is written x * 0.5
    // (the same correctly rounded value)
    double d = delta;
    pi_moves = false;
    new_pi = old_pi;
    if (selected < 0) {
        d = __dadd_rn(d, __dmul_rn(kk, __dmul_rn(v0, 0.5)));
    } else if (selected + 1 >= r_size) {
        d = __dadd_rn(d, __dmul_rn(kk, __dmul_rn(vsel, 1.5)));
    } else {
        new_pi = __dadd_rn(old_pi, __dmul_rn(__dadd_rn(vsel, vnext), 0.5));
        pi_moves = true;
        d = __dadd_rn(d, __dmul_rn(kk, __dsub_rn(vnext, vsel)));
    }
    return d;
}

// One thread: applies affect's scalars to the lane and publishes selected / d.
__device__ __forceinline__ void
affect_scalars(Lane& L, int k, int selected, int r_size, double kappa, double v0, double vsel,
               double vnext)
{
    const double old_pi = L.pi[k];
    double npi;
    bool moves;
    const double d =
      affect_step(L.delta, L.sh->kk, old_pi, selected, r_size, v0, vsel, vnext, npi, moves);
    if (moves) {
        L.pi[k] = npi;
        if ((__double_as_longlong(old_pi) < 0) != (__double_as_longlong(npi) < 0))
            L.sh->pi_changed = 1;
    }
    L.sh->selected = selected;
    L.sh->d = d;
}

// Single-thread statement of calculator_sort + select_variables on the row:
// the exact libstdc++ sort and shuffle permutations.  Used when the parallel
// paths cannot decide the row (ties across the selection boundary, NaN, -0.0,
// rows longer than the shared-memory sort).
// @region row_select_exact
template<typename Less>
__device__ __noinline__ void
row_select_exact(Lane& L, int k, int r_size, double kappa, bool raw, Less less,
                 unsigned long long* wk, int* wid)
{
    // wk / wid: the row's keys and positions in row order (filled by the CTA),
    // in shared memory when the plan found room for them: this is one thread
    // doing some n log n dependent accesses
    const DevProblem& pb = *L.pb;
    const bool maximize = pb.maximize != 0;
    uint32_t rng = L.sh->rng;
    Pairs a{ wk, wid };

    auto value_at = [&](int p) {
        return raw ? __longlong_as_double((long long)wk[p]) : value_of_key(wk[p], maximize);
    };

    (void)less; // the CTA sorted wk / wid already (cta_std_sort: libstdc++'s std::sort, exactly)
    if (r_size > 1) {
        // random_shuffle_unique, itm-common.hpp:117-132
        int ret = 0;
        int i = 1;
        for (; i != r_size; ++i) {
            const bool differ = raw ? (value_at(ret) != value_at(i)) : (wk[ret] != wk[i]);
            if (differ) {
                seq_shuffle(a, ret, i - ret, rng);
                ret = i;
            }
        }
        seq_shuffle(a, ret, i - ret, rng);
    }

    const int c_size = kFam == FAM_101 ? pb.row_hdr[k].c_size : 0;
    const int bkmin = pb.row_hdr[k].bmin, bkmax = pb.row_hdr[k].bmax;
    int selected;
    if (bkmin == bkmax) {
        const int b = bkmin + c_size;
        selected = (b < r_size ? b : r_size) - 1;
    } else {
        int lo, hi;
        select_window(kFam, r_size, c_size, bkmin, bkmax, lo, hi);
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

// calculator_sort's std::sort (itm-common.hpp:134-148) of the row's keys by the
// whole CTA, with libstdc++'s exact result (cta_std_sort).  Out of line: its
// stack must not weigh on the update's registers.
__device__ __noinline__ void
row_order_exact(Lane& L, unsigned long long* wk, int* wid, int r_size, bool raw, bool maximize)
{
    SortQueue* const q = &L.sh->sq;
    if (threadIdx.x == 0) {
        q->tasks = reinterpret_cast<int*>(L.stage);
        q->cap = (int)(((size_t)L.stage_cap * sizeof(double2)) / (3 * sizeof(int)));
    }
    const Pairs a{ wk, wid };
    if (!raw)
        cta_std_sort(a, r_size, LessKey(), q);
    else if (maximize)
        cta_std_sort(a, r_size, LessRawMax(), q);
    else
        cta_std_sort(a, r_size, LessRawMin(), q);
}

// @region cta_bitonic_sort
// Bitonic sort of sk[0..n2) (ascending) carrying sid, by the whole CTA.  n2
// is a power of two.  Strides up to 32 stay inside the 64 elements one warp
// owns, so only the wide strides need a CTA barrier.
__device__ __noinline__ void
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

// @region cta_bitonic_chunked
// The same order for rows whose scratch is in global memory, with the
// exchanges done in shared memory a chunk at a time.  The network is the
// "flip" form of the bitonic sort, every exchange ascending: positions past
// n stand for +infinity and never move, so any n is sorted without padding
// and an exchange that reaches past n is skipped.  Blocks up to CH elements
// (a power of two) are exchanged in the chunk buffer wk / wid: one visit per
// chunk sorts it completely (block sizes 2 .. CH), and for each larger block
// size only the strides >= CH go through global memory (csplib022-shaped rows
// of 16.7 k elements with CH = 16 384: two visits and one global step instead
// of 120 passes over an L2-resident scratch).  Keys come from src in row
// order; ids are generated.
__device__ __forceinline__ void
bitonic_exchange(unsigned long long* k, int* id, int i, int l)
{
    const unsigned long long a = k[i], b = k[l];
    if (a > b) {
        k[i] = b;
        k[l] = a;
        const int t = id[i];
        id[i] = id[l];
        id[l] = t;
    }
}

// pairs (o, kk - 1 - o) of every block of kk elements
__device__ __forceinline__ void
bitonic_flip(unsigned long long* k, int* id, int len, int kk)
{
    const int hb = kk >> 1;
    const int pairs = ((len + kk - 1) / kk) * hb;
    for (int t = threadIdx.x; t < pairs; t += blockDim.x) {
        const int b = t >> (31 - __clz(hb)), o = t & (hb - 1);
        const int i = b * kk + o, l = b * kk + (kk - 1 - o);
        if (l < len)
            bitonic_exchange(k, id, i, l);
    }
}

// pairs (i, i + j) with bit j of i clear
__device__ __forceinline__ void
bitonic_stride(unsigned long long* k, int* id, int len, int j)
{
    const int pairs = ((len + 2 * j - 1) / (2 * j)) * j;
    for (int t = threadIdx.x; t < pairs; t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        if (l < len)
            bitonic_exchange(k, id, i, l);
    }
}

// barrier between two steps: exchanges that stay inside 64 elements are done
// by the warp that owns them in both steps
__device__ __forceinline__ void
bitonic_sync(int reach, int next_reach)
{
    if (reach > 32 || next_reach > 32)
        __syncthreads();
    else
        __syncwarp();
}

__device__ __noinline__ void
cta_bitonic_chunked(const unsigned long long* src, unsigned long long* gk, int* gid, int n,
                    unsigned long long* wk, int* wid, int CH)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int nch = (n + CH - 1) / CH;
    for (int c = 0; c < nch; ++c) {
        const int base = c * CH;
        const int len = n - base < CH ? n - base : CH;
        for (int i = tid; i < len; i += T) {
            wk[i] = src[base + i];
            wid[i] = base + i;
        }
        __syncthreads();
        for (int kk = 2; kk < 2 * len; kk <<= 1) {
            bitonic_flip(wk, wid, len, kk);
            bitonic_sync(kk >> 1, kk > 2 ? kk >> 2 : kk);
            for (int j = kk >> 2; j > 0; j >>= 1) {
                bitonic_stride(wk, wid, len, j);
                bitonic_sync(j, j > 1 ? j >> 1 : kk);
            }
        }
        __syncthreads();
        for (int i = tid; i < len; i += T) {
            gk[base + i] = wk[i];
            gid[base + i] = wid[i];
        }
        __syncthreads();
    }
    for (int kk = 2 * CH; kk < 2 * n; kk <<= 1) {
        bitonic_flip(gk, gid, n, kk);
        __syncthreads();
        for (int j = kk >> 2; j >= CH; j >>= 1) {
            bitonic_stride(gk, gid, n, j);
            __syncthreads();
        }
        for (int c = 0; c < nch; ++c) {
            const int base = c * CH;
            const int len = n - base < CH ? n - base : CH;
            for (int i = tid; i < len; i += T) {
                wk[i] = gk[base + i];
                wid[i] = gid[base + i];
            }
            __syncthreads();
            for (int j = CH >> 1; j > 0; j >>= 1) {
                bitonic_stride(wk, wid, len, j);
                bitonic_sync(j, j > 1 ? j >> 1 : 64);
            }
            for (int i = tid; i < len; i += T) {
                gk[base + i] = wk[i];
                gid[base + i] = wid[i];
            }
            __syncthreads();
        }
    }
}

// @region row_select_sorted
// Parallel statement of calculator_sort + select_variables: the whole CTA
// sorts the keys, the shuffle draws are accounted with the LCG jump-ahead
// (they do not depend on the elements), one thread walks the window of
// select_variables.  Falls back to row_select_exact (sh->slow) when equal
// values straddle the selection boundary.  key0 must hold keys.
__device__ __noinline__ void
row_select_sorted(Lane& L, int k, int r_size, int n2, int n_pre, double kappa)
{
    const DevProblem& pb = *L.pb;
    const int tid = threadIdx.x, T = blockDim.x;
    const bool maximize = pb.maximize != 0;

#if BXG_TIMERS
    const long long t_sort0 = tid == 0 ? clock64() : 0;
#endif
    if (L.xk && L.xchunk >= 64) {
        // scratch in global memory, a chunk buffer in shared memory
        cta_bitonic_chunked(L.key0, L.sk, L.sid, r_size, L.xk, L.xid, L.xchunk);
    } else {
        for (int i = tid; i < n2; i += T) {
            L.sk[i] = i < r_size ? L.key0[i] : ~0ull;
            L.sid[i] = i;
        }
        __syncthreads();
        cta_bitonic_sort(L.sk, L.sid, n2);
    }
#if BXG_TIMERS
    // (ordered rows: phase 2 = the CTA sort, phase 3 = runs, draws and the window)
    const long long t_sort1 = tid == 0 ? clock64() : 0;
    if (tid == 0)
        L.cyc_phase[2] += (unsigned long long)(t_sort1 - t_sort0);
#endif

    // equal keys next to each other = a run random_shuffle_unique shuffles
    int tie = 0;
    for (int p = tid + 1; p < r_size; p += T)
        tie |= (L.sk[p] == L.sk[p - 1]);
    const bool has_tie = __syncthreads_or(tie) != 0;
    const uint32_t rng0 = L.sh->rng;
    int total_draws = 0;
    int reject = 0;

    if (has_tie) {
        // Draws of all the std::shuffle calls, in sorted order.  A run of n
        // equal keys starting at a draws floor(n/2) numbers: with n odd the
        // element at odd offset o owns draw (o-1)/2; with n even offset 1
        // owns the single-swap draw 0 and even offsets o >= 2 own draw o/2.
        // Either way the index of a draw in the row's sequence is the count
        // of odd offsets before its owner.
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
            L.sh->redi[wp] = incl;
        __syncthreads();
        int before = 0;
        const int nw = T >> 5;
        for (int w = 0; w < nw; ++w) {
            const int v = L.sh->redi[w];
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
                const uint32_t g = lcg_mul(rng0, lcg_pow((uint64_t)t + 1u));
                const uint32_t scaling = 2147483645u / uerange;
                if (g - 1u >= uerange * scaling)
                    reject = 1;
            }
            t += o & 1;
        }
    }
    const bool any_reject = has_tie && __syncthreads_or(reject) != 0;

    // A uniform draw was rejected somewhere (with runs of thousands of equal
    // values the ranges reach n^2 and one draw in ten is): the draws are
    // replayed one after the other.  The replay only needs the LENGTHS of the
    // runs: the CTA lists the run starts (ordered compaction into the staging
    // buffer, idle here), so the one thread walks a few runs instead of
    // re-reading every key.
    int* const runs = reinterpret_cast<int*>(L.stage);
    const bool listed = any_reject && (size_t)(r_size + 2) * sizeof(int) <= (size_t)L.stage_cap * sizeof(double2);
    int n_runs = 0;
    if (listed) {
        const int C = (r_size + T - 1) / T;
        const int p0 = min(r_size, tid * C), p1 = min(r_size, p0 + C);
        int mine = 0;
        for (int p = p0; p < p1; ++p)
            mine += p == 0 || L.sk[p] != L.sk[p - 1];
        int incl = mine;
        const int ln = tid & 31, wp = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (ln >= o)
                incl += v;
        }
        __syncthreads(); // redi: the scan above is done with it
        if (ln == 31)
            L.sh->redi[wp] = incl;
        __syncthreads();
        int before = 0;
        const int nw = T >> 5;
        for (int w = 0; w < nw; ++w) {
            const int v = L.sh->redi[w];
            if (w < wp)
                before += v;
            n_runs += v;
        }
        int at = before + incl - mine;
        for (int p = p0; p < p1; ++p)
            if (p == 0 || L.sk[p] != L.sk[p - 1])
                runs[at++] = p;
        if (tid == 0)
            runs[n_runs] = r_size;
        __syncthreads();
    }

    if (tid == 0) {
        uint32_t rng = rng0;
        if (listed) {
            for (int j = 0; j != n_runs; ++j)
                seq_shuffle_draws(runs[j + 1] - runs[j], rng);
        } else if (any_reject) {
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

        const int c_size = kFam == FAM_101 ? pb.row_hdr[k].c_size : 0;
        const int bkmin = pb.row_hdr[k].bmin, bkmax = pb.row_hdr[k].bmax;
        int selected;
        if (bkmin == bkmax) {
            const int b = bkmin + c_size;
            selected = (b < r_size ? b : r_size) - 1;
        } else {
            int lo, hi;
            select_window(kFam, r_size, c_size, bkmin, bkmax, lo, hi);
            selected = hi - 1;
            // the first n_pre sorted values neither stop the scan nor draw
            int i = lo;
            if (i < n_pre)
                i = n_pre;
            for (; i <= hi; ++i) {
                const double v = i < r_size ? value_of_key(L.sk[i], maximize) : L.Rv[r_size];
                if (stop_iterating_rng(v, maximize, rng)) {
                    selected = i - 1;
                    break;
                }
            }
        }

        if (selected >= 0 && selected + 1 < r_size && L.sk[selected] == L.sk[selected + 1]) {
            // equal values on both sides of the boundary: which of them are
            // selected is decided by the exact permutations
            L.sh->slow = 1;
        } else {
            L.sh->slow = 0;
            L.sh->rng = rng;
            const double v0 = value_of_key(L.sk[0], maximize);
            const double vsel = selected >= 0 ? value_of_key(L.sk[selected], maximize) : 0.0;
            const double vnext =
              (selected >= 0 && selected + 1 < r_size) ? value_of_key(L.sk[selected + 1], maximize) : 0.0;
            affect_scalars(L, k, selected, r_size, kappa, v0, vsel, vnext);
        }
#if BXG_TIMERS
        L.cyc_phase[3] += (unsigned long long)(clock64() - t_sort1);
#endif
    }
    __syncthreads();
}

// @region affect_element
// x / P part of affect (itm-common.hpp:404-463) for the element `id` of row k,
// plus the row activities of the column when the bit flips.
template<int FAM, bool SM>
__device__ __forceinline__ void
affect_element(Lane& L, int r0, int id, const int4 info, double pv, bool in, double d)
{
    // info = row_info[r0 + id]: c0, c1, own CSC position, column
    const bool neg = FAM != FAM_01 && L.row_a[r0 + id] < 0;
    const bool set = in != neg;
    double* const gP = L.P;
    gP[info.z] = set ? __dadd_rn(pv, d) : __dsub_rn(pv, d);

    uint32_t* const x = view<SM>(L.x, L.o_x);
    int* const act = view<true>(L.act, L.o_act);
    const int j = info.w;
    const uint32_t bit = 1u << (j & 31);
    const bool was = (x[j >> 5] & bit) != 0;
    if (was != set) {
        atomicXor(&x[j >> 5], bit);
        the_shared()->x_dirty = 1;
        const int sgn = set ? 1 : -1;
        // the rows of the column, eight independent loads at a time (the CTA
        // waits for the few threads whose bit flips: one memory round trip
        // per eight entries instead of one per entry)
        const int* __restrict__ csc_rs = L.csc_rs;
        const int t_last = info.y - 1;
        for (int t0 = info.x; t0 < info.y; t0 += 8) {
            int rs[8], a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u < t_last ? t0 + u : t_last;
                rs[u] = csc_rs[t];
                a[u] = FAM == FAM_Z ? L.csc_a[t] : 1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (t0 + u < info.y) {
                    if (FAM == FAM_01)
                        atomicAdd(&act[rs[u] >> 1], sgn);
                    else if (FAM == FAM_101)
                        atomicAdd(&act[rs[u] >> 1], (rs[u] & 1) ? -sgn : sgn);
                    else
                        atomicAdd(&act[rs[u] >> 1], sgn * a[u]);
                }
            }
        }
    }
}

// @region helpers
// 64-bit minimum over a warp with two 32-bit REDUX steps.
__device__ __forceinline__ unsigned long long
warp_min_u64(unsigned long long v)
{
    const uint32_t hi = (uint32_t)(v >> 32);
    const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t lo = hi == mhi ? (uint32_t)v : 0xffffffffu;
    const uint32_t mlo = __reduce_min_sync(0xffffffffu, lo);
    return ((unsigned long long)mhi << 32) | mlo;
}

// Top-2 merge of ascending pairs.
__device__ __forceinline__ void
merge_min2(unsigned long long& m1, unsigned long long& m2, unsigned long long o1,
           unsigned long long o2)
{
    const unsigned long long lo = m1 < o1 ? m1 : o1;
    const unsigned long long hi = m1 < o1 ? o1 : m1;
    const unsigned long long s2 = m2 < o2 ? m2 : o2;
    m1 = lo;
    m2 = hi < s2 ? hi : s2;
}

// @region cta_rank_select
// The key of rank `rank` (0-based, ascending) among keys[0..n) by the whole
// CTA, without ordering them: most-significant-digit-first radix select, eight
// bits a pass, a 256-bin histogram in shared memory (equal digits of a warp
// are counted with one atomic).  Returns the key; `mult` = how many elements
// hold it, `pos` = the rank's offset inside that group of equal keys; `next`
// = the smallest key above it (~0 when none).  What nth_element would give
// the reference if calculator_sort were not a full sort: enough when the
// equal-value runs of the row are known separately (update_row, mode 4).
__device__ __noinline__ unsigned long long
cta_rank_select(const unsigned long long* keys, int n, int rank, uint32_t* hist, Shared* sh, int& mult,
                int& pos, unsigned long long& next)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int ln = tid & 31, wp = tid >> 5, nw = T >> 5;
    unsigned long long prefix = 0, decided = 0;
    int r = rank, cnt = n;
    for (int shift = 56; shift >= 0 && cnt > 1; shift -= 8) {
        for (int i = tid; i < 256; i += T)
            hist[i] = 0u;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += T) {
            const int i = i0 + tid;
            unsigned d = 256u;
            if (i < n) {
                const unsigned long long key = keys[i];
                if ((key & decided) == prefix)
                    d = (unsigned)(key >> shift) & 255u;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (d < 256u && ln == __ffs(peers) - 1)
                atomicAdd(&hist[d], (uint32_t)__popc(peers));
        }
        __syncthreads();
        if (wp == 0) {
            uint32_t c[8];
            int s = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                c[u] = hist[ln * 8 + u];
                s += (int)c[u];
            }
            int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (ln >= o)
                    incl += v;
            }
            int acc = incl - s;
            if (acc <= r && r < incl) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (r >= acc && r < acc + (int)c[u]) {
                        sh->redi[0] = ln * 8 + u;
                        sh->redi[1] = r - acc;
                        sh->redi[2] = (int)c[u];
                    }
                    acc += (int)c[u];
                }
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)(unsigned)sh->redi[0] << shift;
        decided |= 255ull << shift;
        r = sh->redi[1];
        cnt = sh->redi[2];
        __syncthreads(); // redi and hist are rewritten by the next pass
    }
    // the (first) element under the prefix is the key; the smallest key above it
    unsigned long long mine = ~0ull, above = ~0ull;
    for (int i = tid; i < n; i += T) {
        const unsigned long long key = keys[i];
        if ((key & decided) == prefix)
            mine = key < mine ? key : mine;
    }
    mine = warp_min_u64(mine);
    if (ln == 0)
        sh->redk[wp][0] = mine;
    __syncthreads();
    unsigned long long sel = sh->redk[0][0];
    for (int w = 1; w < nw; ++w) {
        const unsigned long long o = sh->redk[w][0];
        sel = o < sel ? o : sel;
    }
    for (int i = tid; i < n; i += T) {
        const unsigned long long key = keys[i];
        if (key > sel)
            above = key < above ? key : above;
    }
    above = warp_min_u64(above);
    if (ln == 0)
        sh->redk[wp][1] = above;
    __syncthreads();
    above = sh->redk[0][1];
    for (int w = 1; w < nw; ++w) {
        const unsigned long long o = sh->redk[w][1];
        above = o < above ? o : above;
    }
    __syncthreads(); // redk is reused by the caller
    mult = cnt;
    pos = r;
    next = above;
    return sel;
}

// @region update_row.head
// One constraint update (the loop body of compute_update_row /
// push_and_compute_update_row).  Called by all threads of the CTA.
//
// A. compute_reduced_costs (itm-solver-inequalities-01.cpp:166-191,
//    -101.cpp:199-233) in two stages.  Stage 1 is parallel over the column
//    ENTRIES of the row (the expanded list gp, built once per problem): every
//    thread loads P and pi for a strided share of the entries -- independent,
//    coalesced loads -- applies the theta decay to the row's own entry and the
//    +-1 factor, and parks the terms in shared memory.  Stage 2 is parallel
//    over the row ELEMENTS: each adds its column's terms left to right in CSC
//    order with the reference's two accumulators, so the sums are bit-equal.
// B. selection without ordering when all values differ (the usual case once
//    pi and P have moved): two smallest keys, or the boundary by sign.
template<int FAM, bool SM>
__device__ __noinline__ void
update_row(int k, int k_next, double kappa, int push, double obj_amp)
{
    Lane& L = the_lane();
    const int tid = threadIdx.x, T = blockDim.x;
    const int ln = tid & 31, wp = tid >> 5, nw = T >> 5;
    Shared* const sh = the_shared();
    const RowHdr hdr = L.row_hdr[k];
    const int r0 = hdr.r0;
    const int r_size = hdr.r_size;
    const double theta = L.theta;
    const bool maximize = L.maximize != 0;
    const double* __restrict__ cost = L.cost;
    const int4* __restrict__ row_info = L.row_info;
    const int* __restrict__ row_eptr = L.row_eptr;
    // shared-memory views (the Lane block itself always lives there)
    // [stage_cap] {P term, pi term} of the column entries being summed.  Rows
    // whose scratch is in global memory stage them in the shared-memory
    // workspace of the exact ordering (xk / xid, one block, idle during the
    // gather), a chunk of elements at a time.
    const int x_entries = SM || !L.xk ? 0 : (int)(((size_t)L.xcap * 12u) / sizeof(double2));
    const bool staged = SM || (x_entries >= 256 && x_entries >= L.pb->max_col);
    const int stage_cap = SM ? L.stage_cap : x_entries;
    double2* const st =
      SM ? reinterpret_cast<double2*>(view<SM>(L.stage, L.o_stage)) : reinterpret_cast<double2*>(L.xk);
    double* const s_pi = view<SM>(L.pi, L.o_pi);
    unsigned long long* const key0 = view<SM>(L.key0, L.o_key0);
    double* const pown = view<SM>(L.pown, L.o_pown);
    double* const gP = L.P;
    const int2* __restrict__ gp = L.gp + hdr.g0;
    const int S = hdr.S;
    const int bkmin = hdr.bmin, bkmax = hdr.bmax;
    const bool is_eq = bkmin == bkmax;

    if (tid == 0)
        L.alg_bytes += hdr.bytes;
#if BXG_TIMERS
    // cycle stamps of thread 0 (bxg_batch_path_counters); a build option: the
    // four 64-bit stamps cost every thread eight registers
    const long long t_start = tid == 0 ? clock64() : 0;
    long long t_gather = t_start, t_sums = 0, t_ties = 0;
#endif
    // {r0, r_size} of the row the sweep visits next: its P lines are sent
    // for while this row is being processed (see "prefetch" below)
    int2 nh = make_int2(0, 0);
    if (k_next >= 0)
        nh = *reinterpret_cast<const int2*>(L.row_hdr + k_next);
    int4 ni = make_int4(0, 0, 0, 0);
    int odd = 0;    // NaN or -0.0 seen by this thread
    int counts = 0; // low half: values that neither stop the scan nor are zero; high half: zeros
    // candidates of this thread: the two smallest keys (ca <= cb); for an
    // inequality row also the largest non-stopping key (complemented, sa) and
    // the smallest stopping key (sb)
    unsigned long long ca = ~0ull, cb = ~0ull, sa = ~0ull, sb = ~0ull;

// @region update_row.finish
    // What both gathers end with for element i of the row: the reduced cost
    // from the two column sums (compute_reduced_costs' last line), the push
    // term, and this thread's share of the selection reductions.
    auto finish = [&](int i, double cj, double v, double own) {
        // v = (c_j - sum a pi) - sum a P on entry
        if (FAM != FAM_01 && L.row_a[r0 + i] < 0) {
            v = -v;
            own = -own;
        }
        if (push)
            v = __dadd_rn(v, __dmul_rn(obj_amp, FAM == FAM_01 ? cost[i] : cj));

        if (v != v || (v == 0.0 && __double_as_longlong(v) < 0))
            odd = 1;
        const bool stop = maximize ? v < 0 : v > 0;
        const unsigned long long key = key_of(v, maximize);
        merge_min2(ca, cb, key, ~0ull);
        if (!is_eq) {
            if (v == 0.0)
                counts += 1 << 16;
            else if (!stop)
                counts += 1;
            if (stop)
                sb = key < sb ? key : sb;
            else
                sa = ~key < sa ? ~key : sa;
        }
        key0[i] = (unsigned long long)__double_as_longlong(v);
        pown[i] = own;
    };

// @region update_row.direct
    // Rows whose scratch is in global memory and that have no shared-memory
    // workspace to stage in: staging the terms in global memory would write
    // and re-read 16 bytes per entry on top of the gather itself.  One thread
    // per element instead, walking its column in CSC order -- P is
    // CSC-ordered, so the column's P values and (row, sign) words are two
    // contiguous runs -- with the same two accumulators, left to right: the
    // same sums bit for bit, 12 bytes per entry.  (Every load instruction
    // touches 32 different sectors: bound by the L1 wavefronts, 830 k cycles
    // for a 16.7 k-element row -- the staged gather below is preferred.)
    if (!staged) {
        const int* __restrict__ csc_rs = L.csc_rs;
        constexpr int V = 4; // independent (word, P) load pairs per trip
        for (int i = tid; i < r_size; i += T) {
            const int4 info = row_info[r0 + i]; // c0, c1, own CSC position, column
            const double cj = cost[info.w];
            const int t_last = info.y - 1;
            double sum_pi = 0.0, sum_p = 0.0, own = 0.0;
            for (int t0 = info.x; t0 < info.y; t0 += V) {
                int rs[V];
                double p[V];
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const int t = t0 + u < t_last ? t0 + u : t_last;
                    rs[u] = csc_rs[t];
                    p[u] = gP[t];
                }
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    if (t0 + u < info.y) {
                        double a_pi = s_pi[rs[u] >> 1];
                        double pv = p[u];
                        if ((rs[u] >> 1) == k)
                            pv = __dmul_rn(pv, theta);
                        if (FAM != FAM_01 && (rs[u] & 1)) {
                            pv = -pv;
                            a_pi = -a_pi;
                        }
                        sum_pi = __dadd_rn(sum_pi, a_pi);
                        sum_p = __dadd_rn(sum_p, pv);
                        if (t0 + u == info.z)
                            own = pv;
                    }
                }
            }
            finish(i, cj, __dsub_rn(__dsub_rn(cj, sum_pi), sum_p), own);
        }
        if (tid < nh.y)
            ni = row_info[nh.x + tid];
        __syncthreads();
#if BXG_TIMERS
        if (tid == 0)
            t_gather = clock64();
#endif
    }

// @region update_row.chunk
    for (int ia = 0; staged && ia < r_size;) {
        // elements [ia, ib) whose entries fit the staging buffer
        int ib = r_size;
        const int fa = ia == 0 ? 0 : row_eptr[r0 + ia];
        int fb = S;
        if (S - fa > stage_cap) {
            // last ib with eptr[ib] - fa <= cap (row_eptr has one entry per
            // element: the end of the last element is S, not the entry after
            // it).  The ends grow with ib, so the CTA probes a spread of
            // candidates at once and counts those that fit: two rounds of one
            // load each instead of a bisection of dependent loads (a
            // 16.7 k-element row is staged in 14 chunks).  ia + 1 always fits
            // (a column is never longer than the buffer) and every element
            // has an entry, so ib <= ia + cap.
            int lo = ia + 1;
            int hi = r_size - ia > stage_cap ? ia + stage_cap : r_size;
            while (lo < hi) {
                const int step = (hi - lo) / T + 1;
                const int c = lo + tid * step;
                bool fits = false;
                if (c <= hi) {
                    const int end = c < r_size ? row_eptr[r0 + c] : S;
                    fits = end - fa <= stage_cap;
                }
                const int n_fit = __syncthreads_count(fits); // >= 1: lo fits
                lo += (n_fit - 1) * step;
                if (n_fit < T) // the candidate after the last that fits does not, or lies past hi
                    hi = lo + step - 1 < hi ? lo + step - 1 : hi;
            }
            ib = lo;
            fb = ib < r_size ? row_eptr[r0 + ib] : S;
        }

// @region update_row.stage1
        // stage 1: entries.  The index is clamped rather than predicated, so
        // the loads of a trip are unconditional and independent.
        const int g_last = fb - 1;
        constexpr int U = BXG_GATHER_UNROLL; // independent (entry, P) load pairs per thread and trip
        for (int g = fa + tid; g < fb; g += U * T) {
            int2 e[U];
            double p[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int gg = g + u * T;
                e[u] = gp[gg < g_last ? gg : g_last];
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                p[u] = gP[e[u].x];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int gg = g + u * T;
                const int rs = e[u].y;
                double a_pi = s_pi[rs >> 1];
                double pv = p[u];
                if ((rs >> 1) == k)
                    pv = __dmul_rn(pv, theta);
                if (FAM != FAM_01 && (rs & 1)) {
                    pv = -pv;
                    a_pi = -a_pi;
                }
                if (gg < fb)
                    st[gg - fa] = make_double2(pv, a_pi);
            }
        }
// @region update_row.prefetch
        // prefetch, first half: where the next row's columns lie in P
        if (ia == 0 && tid < nh.y)
            ni = row_info[nh.x + tid];
        __syncthreads();
#if BXG_TIMERS
        if (ia == 0 && tid == 0)
            t_gather = clock64();
#endif

// @region update_row.stage2
        // stage 2: elements, each adding its column's terms left to right
        for (int i = ia + tid; i < ib; i += T) {
            const int4 info = row_info[r0 + i]; // c0, c1, own CSC position, column
            const double2* const c = st + (row_eptr[r0 + i] - fa);
            const int len = info.y - info.x;
            const double cj = cost[info.w];
            double sum_pi = 0.0, sum_p = 0.0;
            for (int t = 0; t < len; ++t) {
                const double2 term = c[t];
                sum_pi = __dadd_rn(sum_pi, term.y);
                sum_p = __dadd_rn(sum_p, term.x);
            }
            finish(i, cj, __dsub_rn(__dsub_rn(cj, sum_pi), sum_p), c[info.z - info.x].x);
        }
        ia = ib;
        if (ia < r_size)
            __syncthreads(); // the staging buffer is reused
    }

// @region update_row.prefetch2
    // prefetch, second half: every 32-byte sector of the next row's columns
    // goes to L2 now; its gather starts after this row's selection and affect.
    if (k_next >= 0) {
        // (a row a few elements longer than the CTA: the tail is not worth a
        // second trip with its own row_info round trip)
        const int n_pref = nh.y <= T + 32 && nh.y > T ? T : nh.y;
        for (int i = tid; i < n_pref; i += T) {
            if (i >= T)
                ni = row_info[nh.x + i];
            const char* a = reinterpret_cast<const char*>(gP + ni.x);
            const char* const end = reinterpret_cast<const char*>(gP + ni.y);
            a = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(a) & ~uintptr_t(31));
            for (; a < end; a += 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
    }

// @region update_row.reduceA
    if (!is_eq) {
        counts = __reduce_add_sync(0xffffffffu, counts);
        if (ln == 0)
            sh->redi[wp] = counts;
    }
    const double old_pi = s_pi[k];
    const int tab_gen0 = L.tab_gen;
    const int tab_dirty0 = L.tab_dirty;
    // (read before the barrier: thread 0 writes them later in this update)
    const int tie_streak0 = L.tie_streak;
    const unsigned ordered0 = L.n_path[0] + L.n_path[1];
    const bool raw = __syncthreads_or(odd) != 0;
#if BXG_TIMERS
    if (tid == 0)
        t_sums = clock64();
#endif
    int n_pre = 0, n_zero = 0;
    if (!is_eq) {
        counts = 0;
        for (int w = 0; w < nw; ++w)
            counts += sh->redi[w];
        n_pre = counts & 0xffff;
        n_zero = counts >> 16;
    }

    const int n2 = r_size <= 2 ? 2 : 1 << (32 - __clz(r_size - 1));

// @region update_row.mode
    // ---- B: which elements are selected ------------------------------------
    // mode 0: exact single thread; 1: CTA sort; 2: two smallest; 3: boundary by sign
    int mode;
    int selected = -1;
    if (raw || r_size < 2 || r_size > 0xffff || n2 > L.sort_cap) {
        mode = 0;
    } else {
        mode = 1;
        if (L.tail_rows == 0) {
            const int c_size = FAM == FAM_101 ? hdr.c_size : 0;
            if (is_eq) {
                const int b = bkmin + c_size;
                selected = (b < r_size ? b : r_size) - 1;
                if (selected <= 0)
                    mode = 2;
            } else if (n_zero == 0) {
                int lo, hi;
                select_window(kFam, r_size, c_size, bkmin, bkmax, lo, hi);
                const int i0 = lo > n_pre ? lo : n_pre; // first index holding a stopping value
                selected = (i0 <= hi ? i0 : hi) - 1;
                if (selected + 1 == n_pre)
                    mode = 3;
                else if (selected <= 0)
                    mode = 2;
            }
        }
    }

    // Long rows that keep showing equal values (integer costs, csplib022-shaped)
    // go to the ordered path without asking the table first -- it handles rows
    // without ties as well, so this only chooses the cheaper route; one row in
    // sixteen still asks, and a pass that finds no equal values resets the streak.
    if (mode >= 2 && !SM && tie_streak0 >= 4 && (ordered0 & 15u) != 0)
        mode = 1;
    // mode 4: the boundary lies inside the row, but `selected` is known before
    // any ordering (an equality row, or an inequality row without zeros).  When
    // all values differ -- or only a few are equal, kRankDups at most, and
    // their run does not straddle the boundary -- affect needs the key of rank
    // `selected`, the next key above it and the smallest one, and the engine
    // needs the lengths of the few equal-value runs in ascending order: a
    // radix select (cta_rank_select) and a list of the equal keys the tie
    // table meets, instead of ordering thousands of elements.  Rows whose
    // scratch is in global memory only (the short ones sort in a few
    // thousand cycles); keys in the shared-memory workspace of the exact
    // ordering, idle here.
    constexpr int kRankDups = 48;
    unsigned long long* dup_keys = nullptr;
    if (!SM && mode == 1 && selected > 0 && L.xk && r_size <= L.xcap && L.xcap >= 1024 &&
        !(tie_streak0 >= 4 && (ordered0 & 15u) != 0)) {
        mode = 4;
        // (xid: 256 histogram bins, then the list of equal keys)
        dup_keys = reinterpret_cast<unsigned long long*>(L.xid + 256);
    }

    if (mode >= 2) {
// @region update_row.ties
        // No order is needed as long as all values differ: look for equal
        // values with a table of row positions hashed by key (32-bit entries
        // laid over sk, stamped with a generation so it is cleared rarely).
        uint32_t* tab = reinterpret_cast<uint32_t*>(view<SM>(L.sk, L.o_sk));
        const int tab_bits = L.tab_bits;
        int tab_gen = tab_gen0;
        if (tab_dirty0 || tab_gen >= 0xffff) {
            for (int i = tid; i < (1 << tab_bits); i += T)
                tab[i] = 0u;
            tab_gen = 0;
            __syncthreads();
            if (tid == 0)
                L.tab_dirty = 0;
        }
        const uint32_t gen = (uint32_t)++tab_gen;
        const uint32_t mask = (1u << tab_bits) - 1u;
        int dup = 0;
        for (int i = tid; i < r_size; i += T) {
            const unsigned long long bits = key0[i];
            if (!SM && mode == 4)
                L.xk[i] = key_of(__longlong_as_double((long long)bits), maximize);
            uint32_t h = ((uint32_t)bits ^ (uint32_t)(bits >> 32)) * 0x9E3779B1u;
            h = (h >> (32 - tab_bits)) & mask;
            const uint32_t mine = (gen << 16) | (uint32_t)(i + 1);
            for (;;) {
                uint32_t cur = tab[h];
                if ((cur >> 16) != gen) {
                    const uint32_t prev = atomicCAS(&tab[h], cur, mine);
                    if (prev == cur)
                        break;
                    cur = prev;
                    if ((cur >> 16) != gen)
                        continue; // lost a race against a stale entry: look again
                }
                const int other = (int)(cur & 0xffffu) - 1;
                if (key0[other] == bits) {
                    dup = 1;
                    if (!SM && mode == 4) {
                        // one entry per element past the first of its run
                        const int slot = atomicAdd(&sh->n_dup, 1);
                        if (slot < kRankDups)
                            dup_keys[slot] = key_of(__longlong_as_double((long long)bits), maximize);
                    }
                    break;
                }
                h = (h + 1u) & mask;
            }
        }

// @region update_row.candidates
        // candidates over the warp: mode 2 = the two smallest keys, mode 3 =
        // (largest non-stopping key, complemented; smallest stopping key)
        unsigned long long a, b;
        if (mode != 3) {
            a = warp_min_u64(ca);
            // one holder of the minimum gives it up, everybody else keeps theirs
            const unsigned holders = __ballot_sync(0xffffffffu, ca == a);
            const bool gives = ca == a && (holders & ((1u << ln) - 1u)) == 0;
            b = warp_min_u64(gives ? cb : ca);
        } else {
            a = warp_min_u64(sa);
            b = warp_min_u64(sb);
        }
        if (ln == 0) {
            sh->redk[wp][0] = a;
            sh->redk[wp][1] = b;
        }
        const bool any_dup = __syncthreads_or(dup) != 0;
        bool ordered = any_dup; // equal values: their runs draw from the engine
        if (any_dup)
            mode = mode == 4 ? 4 : 1;
        else {
            a = sh->redk[0][0];
            b = sh->redk[0][1];
            for (int w = 1; w < nw; ++w) {
                const unsigned long long oa = sh->redk[w][0], ob = sh->redk[w][1];
                if (mode != 3)
                    merge_min2(a, b, oa, ob);
                else {
                    a = oa < a ? oa : a;
                    b = ob < b ? ob : b;
                }
            }
        }
        unsigned long long key_sel = 0;
        double v0 = 0.0, vsel = 0.0, vnext = 0.0;
        if (!SM && mode == 4) {
// @region update_row.rank
            const int n_dup = any_dup ? sh->n_dup : 0;
            if (any_dup) {
                // (the candidates of the warps were not merged above)
                a = sh->redk[0][0];
                for (int w = 1; w < nw; ++w)
                    a = sh->redk[w][0] < a ? sh->redk[w][0] : a;
            }
            __syncthreads(); // redk, n_dup read by everybody
            if (n_dup > kRankDups)
                mode = 1; // many equal values (integer costs, first sweeps): the ordered path
            else {
                int mult, pos;
                unsigned long long above;
                key_sel = cta_rank_select(L.xk, r_size, selected, reinterpret_cast<uint32_t*>(L.xid), sh,
                                          mult, pos, above);
                if (selected + 1 < r_size && pos + 1 < mult)
                    mode = 0; // equal values on both sides of the boundary: the exact permutations decide
                else {
                    ordered = false;
                    v0 = value_of_key(a, maximize);
                    vsel = value_of_key(key_sel, maximize);
                    vnext = selected + 1 < r_size ? value_of_key(above, maximize) : 0.0;
                    if (tid == 0 && n_dup) {
                        // random_shuffle_unique's draws: the runs of equal values in
                        // ascending order, each one element longer than its entries here
                        for (int i = 1; i < n_dup; ++i) {
                            const unsigned long long key = dup_keys[i];
                            int j = i;
                            for (; j > 0 && dup_keys[j - 1] > key; --j)
                                dup_keys[j] = dup_keys[j - 1];
                            dup_keys[j] = key;
                        }
                        uint32_t rng = sh->rng;
                        for (int i = 0; i < n_dup;) {
                            int j = i + 1;
                            while (j < n_dup && dup_keys[j] == dup_keys[i])
                                ++j;
                            seq_shuffle_draws(j - i + 1, rng);
                            i = j;
                        }
                        sh->rng = rng;
                    }
                }
            }
            if (tid == 0)
                sh->n_dup = 0;
        }
        if (tid == 0) {
            L.tab_gen = tab_gen;
            if (!SM)
                L.tie_streak = ordered ? tie_streak0 + 1 : 0;
#if BXG_TIMERS
            t_ties = clock64();
#endif
        }
        if (!ordered) {
// @region update_row.scalars
            // every thread derives the same scalars; thread 0 publishes pi
            if (mode == 4) {
                // (cta_rank_select gave them)
            } else if (mode == 2) {
                v0 = value_of_key(a, maximize);
                vsel = v0;
                vnext = value_of_key(b, maximize);
                key_sel = a;
            } else {
                key_sel = ~a;
                vsel = value_of_key(~a, maximize);
                vnext = value_of_key(b, maximize);
                v0 = selected >= 0 ? vsel : vnext; // only read when nothing is selected
            }
            double npi;
            bool moves;
            const double d =
              affect_step(L.delta, sh->kk, old_pi, selected, r_size, v0, vsel, vnext, npi, moves);
            if (tid == 0 && moves) {
                s_pi[k] = npi; // all threads read pi[k] before the barrier above
                if ((__double_as_longlong(old_pi) < 0) != (__double_as_longlong(npi) < 0))
                    sh->pi_changed = 1;
            }
// @region update_row.fastD
            for (int i = tid; i < r_size; i += T) {
                const int4 info = row_info[r0 + i];
                const unsigned long long key =
                  key_of(__longlong_as_double((long long)key0[i]), maximize);
                affect_element<FAM, SM>(L, r0, i, info, pown[i], selected >= 0 && key <= key_sel, d);
            }
            __syncthreads();
            if (tid == 0) {
                const int path = mode == 4 ? 2 : mode; // (counted with the other unordered rows)
                L.n_path[path] += 1;
#if BXG_TIMERS
                const long long t_end = clock64();
                L.cyc_path[path] += (unsigned long long)(t_end - t_start);
                L.cyc_phase[0] += (unsigned long long)(t_gather - t_start);
                L.cyc_phase[1] += (unsigned long long)(t_sums - t_gather);
                L.cyc_phase[2] += (unsigned long long)(t_ties - t_sums);
                L.cyc_phase[3] += (unsigned long long)(t_end - t_ties);
#endif
            }
            return;
        }
    }

// @region update_row.ordered
    // keys for the ordered paths (raw rows keep the value bits)
    if (!raw) {
        for (int i = tid; i < r_size; i += T)
            L.key0[i] = key_of(__longlong_as_double((long long)L.key0[i]), maximize);
    }
    if (tid == 0) {
        L.tab_dirty = 1;
#if BXG_TIMERS
        // ordered rows: phase 0 = gather, phase 1 = column sums (+ a tie table that found equal values)
        L.cyc_phase[0] += (unsigned long long)(t_gather - t_start);
        L.cyc_phase[1] += (unsigned long long)(clock64() - t_gather);
#endif
    }
    __syncthreads();

    bool slow = mode == 0;
    if (!slow) {
        row_select_sorted(L, k, r_size, n2, n_pre, kappa);
        slow = L.sh->slow != 0;
    }
    if (slow) {
        // the exact ordering by one thread: on a copy in shared memory when there is room
        const bool in_smem = L.xk && r_size <= L.xcap;
        unsigned long long* const wk = in_smem ? L.xk : L.sk;
        int* const wid = in_smem ? L.xid : L.sid;
        for (int i = tid; i < r_size; i += T) {
            wk[i] = L.key0[i];
            wid[i] = i;
        }
        __syncthreads();
        row_order_exact(L, wk, wid, r_size, raw, maximize);
        if (tid == 0) {
            if (!raw)
                row_select_exact(L, k, r_size, kappa, false, LessKey(), wk, wid);
            else if (maximize)
                row_select_exact(L, k, r_size, kappa, true, LessRawMax(), wk, wid);
            else
                row_select_exact(L, k, r_size, kappa, true, LessRawMin(), wk, wid);
        }
        __syncthreads();
        if (in_smem) {
            for (int i = tid; i < r_size; i += T) {
                L.sk[i] = wk[i];
                L.sid[i] = wid[i];
            }
            __syncthreads();
        }
    }

    // ---- D: x / P part of affect, by sorted position ------------------------
    {
        const int sel = L.sh->selected;
        const double d = L.sh->d;
        const bool keep_tail = L.tail_rows != 0;
        for (int p = tid; p < r_size; p += T) {
            const int id = L.sid[p];
            affect_element<FAM, SM>(L, r0, id, L.row_info[r0 + id], pown[id], p <= sel, d);
            if (keep_tail)
                L.Rv[p] = raw ? __longlong_as_double((long long)L.sk[p]) : value_of_key(L.sk[p], maximize);
        }
    }
    __syncthreads();
    if (tid == 0) {
        L.n_path[slow ? 0 : 1] += 1;
#if BXG_TIMERS
        L.cyc_path[slow ? 0 : 1] += (unsigned long long)(clock64() - t_start);
#endif
    }
}

#if BXG_FAM == 2
// @region update_row_z
// select_variables / select_variables_101 of the Z solver (identical bodies,
// itm-solver-inequalities-Z.cpp:308-344): the longest prefix of the ordered
// row whose factor sum is inside the bounds, stopping at the first value of
// the wrong sign once one prefix fits.  -2 when no prefix fits.  (The
// reference's last trip reads R[r_size]; it cannot change the result.)
__device__ inline int
z_select_prefix(const ZRow& R, int r_size, int bkmin, int bkmax, bool maximize)
{
    int sum = 0;
    int best = -2;
    for (int i = -1; i < r_size; ++i) {
        if (bkmin <= sum && sum <= bkmax)
            best = i;
        if (i + 1 == r_size)
            break;
        if (best != -2) {
            const double v = R.value(i + 1);
            if (maximize ? v < 0 : v > 0)
                break;
        }
        sum += R.factor(i + 1);
    }
    return best;
}

// One constraint update of the Z solver (compute_update_row /
// push_and_compute_update_row, itm-solver-inequalities-Z.cpp:441-564).
//  A. decrease_preference + compute_reduced_costs (:242-293): one thread per
//     row element walks its column in CSC order with the reference's single
//     accumulator sum |a| (pi + P).
//  B. calculator_sort where the row type asks for it: the CTA's bitonic
//     network when all values differ (the order is then unique), else the
//     exact libstdc++ sort + shuffles by one thread; then the row's selection
//     rule (prefix, exhaustive, branch and bound) by one thread.
//  C. local_affect (:346-439), one thread per element: P +- d, the x bit, and
//     the "repair" step that re-reads the element's column and pushes P past
//     zero.  Elements of a row sit in different columns, so they are
//     independent once pi_k is known.
template<bool SM>
__device__ __noinline__ void
update_row_z(int k, double kappa, int push, double obj_amp)
{
    Lane& L = the_lane();
    const DevProblem& pb = *L.pb;
    const DevParams& pr = *L.pr;
    const int tid = threadIdx.x, T = blockDim.x;
    const RowHdr hdr = pb.row_hdr[k];
    const int r0 = hdr.r0;
    const int r_size = hdr.r_size;
    const double theta = L.theta;
    const bool maximize = pb.maximize != 0;
    const double* __restrict__ cost = pr.cost;
    double* const s_pi = view<SM>(L.pi, L.o_pi);
    unsigned long long* const key0 = view<SM>(L.key0, L.o_key0);
    unsigned long long* const sk = view<SM>(L.sk, L.o_sk);
    double* const pown = view<SM>(L.pown, L.o_pown);
    int* const sid = L.sid;
    double* const gP = L.P;
    Shared* const sh = the_shared();
    const int zt = pb.ztype[k];

    if (tid == 0)
        L.alg_bytes += hdr.bytes;

    // ---- A ---------------------------------------------------------------
    int odd = 0;
    for (int i = tid; i < r_size; i += T) {
        const int4 info = pb.row_info[r0 + i]; // c0, c1, own CSC position, column
        double sum = 0.0, own = 0.0;
        for (int t = info.x; t < info.y; ++t) {
            const int a = pb.csc_a[t];
            double p = gP[t];
            if (t == info.z) {
                p = __dmul_rn(p, theta);
                own = p;
            }
            sum = __dadd_rn(sum, __dmul_rn((double)(a < 0 ? -a : a),
                                           __dadd_rn(s_pi[pb.csc_rs[t] >> 1], p)));
        }
        const double cj = cost[info.w];
        double v = __dsub_rn(cj, sum);
        if (push)
            v = __dadd_rn(v, __dmul_rn(obj_amp, cj));
        if (v != v || (v == 0.0 && __double_as_longlong(v) < 0))
            odd = 1;
        key0[i] = (unsigned long long)__double_as_longlong(v);
        pown[i] = own;
    }
    const bool raw = __syncthreads_or(odd) != 0;

    // ---- B ---------------------------------------------------------------
    const bool ordered = zt != Z_EXHAUSTIVE && r_size > 1;
    bool sorted = false;
    if (ordered && !raw && r_size > 32) {
        const int n2 = 1 << (32 - __clz(r_size - 1));
        if (n2 <= L.sort_cap) {
            for (int i = tid; i < n2; i += T) {
                sk[i] = i < r_size ? key_of(__longlong_as_double((long long)key0[i]), maximize)
                                   : ~0ull;
                sid[i] = i;
            }
            __syncthreads();
            cta_bitonic_sort(sk, sid, n2);
            int tie = 0;
            for (int p = tid + 1; p < r_size; p += T)
                tie |= (sk[p] == sk[p - 1]);
            sorted = __syncthreads_or(tie) == 0;
            if (sorted) {
                for (int p = tid; p < r_size; p += T)
                    sk[p] = (unsigned long long)__double_as_longlong(value_of_key(sk[p], maximize));
            }
            __syncthreads();
        }
    }
    if (!sorted) {
        for (int i = tid; i < r_size; i += T) {
            sk[i] = key0[i];
            sid[i] = i;
        }
        __syncthreads();
    }

    if (tid == 0) {
        const ZRow R{ L.sk, L.sid, pb.row_a + r0 };
        uint32_t rng = sh->rng;
        if (ordered && !sorted) {
            // calculator_sort, itm-common.hpp:134-148
            const Pairs a{ L.sk, L.sid };
            if (maximize)
                seq_std_sort(a, r_size, LessRawMax());
            else
                seq_std_sort(a, r_size, LessRawMin());
            int ret = 0;
            int i = 1;
            for (; i != r_size; ++i) {
                if (R.value(ret) != R.value(i)) {
                    seq_shuffle(a, ret, i - ret, rng);
                    ret = i;
                }
            }
            seq_shuffle(a, ret, i - ret, rng);
            sh->rng = rng;
        }

        const int bkmin = hdr.bmin, bkmax = hdr.bmax;
        int selected;
        int need_cta = 0; // the branch-and-bound enumeration runs on the whole CTA
        if (zt == Z_LINEAR)
            selected = z_select_prefix(R, r_size, bkmin, bkmax, maximize);
        else {
            const ZWork w(L.zs, pb.max_zrow);
            selected = -2;
            if (zt == Z_UNCONSTRAINED_BRANCH_AND_BOUND || zt == Z_UNCONSTRAINED_EXHAUSTIVE)
                selected = z_select_prefix(R, r_size, bkmin, bkmax, maximize);
            if (selected == -2) {
                if (zt == Z_BRANCH_AND_BOUND || zt == Z_UNCONSTRAINED_BRANCH_AND_BOUND) {
                    if (r_size <= kZParallelMax) {
                        int b0 = bkmin, b1 = bkmax;
                        z_bb_prepare(w, R, r_size, maximize, b0, b1);
                        sh->redi[0] = b0;
                        sh->redi[1] = b1;
                        need_cta = 1;
                    } else
                        selected = z_bb_solve(w, R, r_size, maximize, bkmin, bkmax);
                } else if (pb.ex_count[k] >= 64) {
                    z_ex_prepare(w, R, r_size);
                    need_cta = 2;
                } else
                    selected = z_ex_solve(w, R, r_size, maximize, pb.ex_flat + pb.ex_start[k],
                                          pb.ex_count[k]);
            }
        }
        sh->selected = selected;
        sh->tmp_i = need_cta;
    }
    __syncthreads();
    if (sh->tmp_i == 2) {
        // (uniform) the feasible patterns of an exhaustive row, a share per thread
        const ZWork w(L.zs, pb.max_zrow);
        const int* masks = pb.ex_flat + pb.ex_start[k];
        const int best = z_ex_enumerate(w, r_size, maximize, masks, pb.ex_count[k], sh->red, sh->redi);
        if (tid == 0) {
            const ZRow R{ L.sk, L.sid, pb.row_a + r0 };
            sh->selected = best >= 0 ? z_ex_finish(w, R, r_size, maximize, masks[best])
                                     : z_ex_solve(w, R, r_size, maximize, masks, pb.ex_count[k]);
        }
    } else if (sh->tmp_i) {
        // (uniform) every thread enumerates its share of the subsets
        const ZWork w(L.zs, pb.max_zrow);
        const int b0 = sh->redi[0], b1 = sh->redi[1];
        __syncthreads(); // redi is the reduction's scratch next
        const unsigned chosen = z_bb_enumerate(w, r_size, maximize, b0, b1, sh->red, sh->redi);
        if (tid == 0) {
            const ZRow R{ L.sk, L.sid, pb.row_a + r0 };
            sh->selected = z_bb_finish(w, R, r_size, maximize, chosen);
        }
    }
    if (tid == 0) {
        const ZRow R{ L.sk, L.sid, pb.row_a + r0 };
        const int selected = sh->selected;

        // scalars of local_affect: pi_k moves in all three cases
        const double old_pi = L.pi[k];
        double npi;
        if (selected < 0)
            npi = __dadd_rn(old_pi, __dmul_rn(R.value(0), 0.5));
        else if (selected + 1 >= r_size) {
            // a sub-solver may answer r_size: the reference then reads the
            // slot one past the row (what an earlier, longer row left there)
            const double vs = selected < r_size ? R.value(selected) : L.Rv[selected];
            npi = __dadd_rn(old_pi, __dmul_rn(vs, 1.5));
        } else
            npi = __dadd_rn(old_pi, __dmul_rn(__dadd_rn(R.value(selected), R.value(selected + 1)), 0.5));
        L.pi[k] = npi;
        if ((__double_as_longlong(old_pi) < 0) != (__double_as_longlong(npi) < 0))
            sh->pi_changed = 1;
        sh->selected = selected;
        sh->d = __dadd_rn(sh->kk, L.delta);
    }
    __syncthreads();

    // ---- C ---------------------------------------------------------------
    {
        const int sel = sh->selected;
        const double d = sh->d;
        const bool keep_tail = pb.tail_rows != 0;
        uint32_t* const x = view<SM>(L.x, L.o_x);
        int* const act = view<true>(L.act, L.o_act);
        for (int p = tid; p < r_size; p += T) {
            const int id = sid[p];
            const int4 info = pb.row_info[r0 + id];
            const bool set = sel >= 0 && (sel + 1 >= r_size || p <= sel);
            double pv = pown[id];
            pv = set ? __dadd_rn(pv, d) : __dsub_rn(pv, d);

            const int j = info.w;
            const uint32_t bit = 1u << (j & 31);
            const bool flips = ((x[j >> 5] & bit) != 0) != set;
            if (flips) {
                atomicXor(&x[j >> 5], bit);
                sh->x_dirty = 1;
            }

            // local_compute_reduced_cost, :295-306, with the new pi_k and P
            double sum = 0.0;
            for (int t = info.x; t < info.y; ++t) {
                const int a = pb.csc_a[t];
                const double pt = t == info.z ? pv : gP[t];
                sum = __dadd_rn(sum, __dmul_rn((double)(a < 0 ? -a : a),
                                               __dadd_rn(s_pi[pb.csc_rs[t] >> 1], pt)));
                if (flips)
                    atomicAdd(&act[pb.csc_rs[t] >> 1], set ? a : -a);
            }
            const double repair = __dsub_rn(cost[j], sum);
            if (set) {
                if (repair >= 0)
                    pv = __dadd_rn(__dsub_rn(pv, repair), d);
            } else if (repair <= 0)
                pv = __dsub_rn(__dadd_rn(pv, repair), d);
            gP[info.z] = pv;

            if (keep_tail)
                L.Rv[p] = __longlong_as_double((long long)sk[p]);
        }
    }
    __syncthreads();
}


#endif // BXG_FAM == 2

// @region activities
// Row activities from scratch (is_valid_constraint's sum, itm-common.hpp:76-89);
// afterwards the updates keep them current.
__device__ void
activities_recompute(Lane& L)
{
    const DevProblem& pb = *L.pb;
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

    for (int k = warp; k < pb.m; k += nwarps) {
        const int e0 = pb.row_ptr[k], e1 = pb.row_ptr[k + 1];
        int v = 0;
        if (kFam == FAM_01) {
            for (int e = e0 + lane; e < e1; e += 32)
                v += xbit(L.x, pb.row_col[e]);
        } else {
            for (int e = e0 + lane; e < e1; e += 32)
                v += pb.row_a[e] * xbit(L.x, pb.row_col[e]);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0)
            L.act[k] = v;
    }
    __syncthreads();
    if (tid == 0)
        L.alg_bytes += 4.0 * pb.nnz * (kFam == FAM_01 ? 1.0 : 2.0) + pb.n / 8.0;
}

// @region violated_scan
// The violated list (compute_violated_constraints, itm-common.hpp:102-115;
// infeasibility_local_compute_violated_constraints, :893-914;
// local_compute_violated_constraints, :882-891) from the row activities.
// mode: 0 = list of violated rows into oR, 1 = (k, amount) pairs into mo,
// 2 = count only.
__device__ int
violated_scan(Lane& L, int mode)
{
    const DevProblem& pb = *L.pb;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int m = pb.m;

    // ordered compaction: every warp takes a contiguous share of the rows
    // (whole groups of 32), counts its violated rows, and after the prefix over
    // the warps writes them in place
    const int groups = (m + 31) >> 5;
    const int per_warp = (groups + nw - 1) / nw;
    const int g0 = warp * per_warp;
    const int g1 = min(groups, g0 + per_warp);
    const int* __restrict__ bmin = pb.bmin;
    const int* __restrict__ bmax = pb.bmax;

    int mine = 0;
    for (int g = g0; g < g1; ++g) {
        const int k = (g << 5) + lane;
        int amount = 0;
        if (k < m) {
            const int v = L.act[k], lo = bmin[k], hi = bmax[k];
            amount = lo > v ? lo - v : (hi < v ? v - hi : 0);
        }
        mine += __popc(__ballot_sync(0xffffffffu, amount != 0));
    }
    if (lane == 0)
        L.sh->redi[warp] = mine;
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nw; ++w) {
        const int v = L.sh->redi[w];
        if (w < warp)
            before += v;
        total += v;
    }
    if (mode != 2 && mine) {
        int count = before;
        for (int g = g0; g < g1; ++g) {
            const int k = (g << 5) + lane;
            int amount = 0;
            if (k < m) {
                const int v = L.act[k], lo = bmin[k], hi = bmax[k];
                amount = lo > v ? lo - v : (hi < v ? v - hi : 0);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, amount != 0);
            if (amount != 0) {
                const int pos = count + __popc(mask & ((1u << lane) - 1u));
                if (mode == 0)
                    L.oR[pos] = k;
                else if (L.mop)
                    L.mop[pos] = ((uint32_t)amount << 16) | (uint32_t)k;
                else
                    L.mo[pos] = make_int2(k, amount);
            }
            count += __popc(mask);
        }
    }
    if (tid == 0) {
        L.sh->remaining = total;
        if (mode == 0)
            L.sh->list_len = total;
        else if (mode == 1)
            L.sh->tmp_i = total;
        L.alg_bytes += 12.0 * m;
    }
    __syncthreads();
    return total;
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
    activities_recompute(L);
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

// @region order
// Single-thread preparation of the order list (the std::sort / std::shuffle
// calls of compute_order::run, itm-common.hpp:700-780).  Kept out of line: its
// stack and registers must not weigh on the update loop.
__device__ __noinline__ void
order_prepare(Lane& L, int order, int oR_len, int mo_len)
{
    uint32_t rng = L.sh->rng;
    switch (order) {
    case ORDER_RANDOM_SORTING:
        seq_shuffle(Plain<int>{ L.oR }, 0, oR_len, rng);
        break;
    case ORDER_INFEASIBILITY_DECR:
        if (L.mop)
            seq_std_sort(Plain<uint32_t>{ L.mop }, mo_len, PackedLessDecr());
        else
            seq_std_sort(Plain<int2>{ L.mo }, mo_len, PairLessDecr());
        break;
    case ORDER_INFEASIBILITY_INCR:
        if (L.mop)
            seq_std_sort(Plain<uint32_t>{ L.mop }, mo_len, PackedLessIncr());
        else
            seq_std_sort(Plain<int2>{ L.mo }, mo_len, PairLessIncr());
        break;
    case ORDER_LAGRANGIAN_DECR:
        seq_std_sort(Plain<int>{ L.oR }, oR_len, PiLessDecr{ L.pi });
        break;
    case ORDER_LAGRANGIAN_INCR:
        seq_std_sort(Plain<int>{ L.oR }, oR_len, PiLessIncr{ L.pi });
        break;
    case ORDER_PI_SIGN_CHANGE:
        seq_shuffle(Plain<int>{ L.oR }, 0, oR_len, rng);
        L.sh->pi_changed = 0;
        break;
    default:
        break;
    }
    L.sh->rng = rng;
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
    if (tid == 0) {
        L.sweeps += 1;
        L.sh->kk = __ddiv_rn(kappa, __dsub_rn(1.0, kappa));
        if (L.pr->deadline_ns) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now > L.pr->deadline_ns)
                L.sh->stop = 1;
        }
    }
    __syncthreads();

    if (order >= ORDER_RANDOM_SORTING && order <= ORDER_PI_SIGN_CHANGE) {
        if (tid == 0)
            order_prepare(L, order, ls.oR_len, ls.mo_len);
        __syncthreads();
    }

    const bool pairs = order == ORDER_INFEASIBILITY_DECR || order == ORDER_INFEASIBILITY_INCR;
    const int len = pairs ? ls.mo_len : ls.oR_len;
    if (tid == 0)
        L.updates += len;
    // An empty list is the list of violated rows of an x that has none: the
    // sweep changes nothing and the scan would find what the last one found
    // (the skipped scan is not counted as algorithmic bytes).
    if (len == 0 && order != ORDER_PI_SIGN_CHANGE)
        return 0;
    // one way to read row i of the list, whichever form it has: plain rows,
    // (row, amount) pairs, or pairs packed amount << 16 | row
    const int* list = pairs ? (L.mop ? reinterpret_cast<const int*>(L.mop) : reinterpret_cast<const int*>(L.mo))
                            : L.oR;
    const int stride = pairs && !L.mop ? 2 : 1;
    const int mask = pairs && L.mop ? 0xffff : 0x7fffffff;
    const bool reversed = order == ORDER_REVERSING;
    for (int i = 0; i < len; ++i) {
        const int k = list[(reversed ? len - 1 - i : i) * stride] & mask;
        const int k_next = i + 1 < len ? (list[(reversed ? len - 2 - i : i + 1) * stride] & mask) : -1;
#if BXG_FAM == 2
        (void)k_next;
        if (L.all_smem)
            update_row_z<true>(k, kappa, push, obj_amp);
        else
            update_row_z<false>(k, kappa, push, obj_amp);
#else
        if (L.all_smem)
            update_row<kFam, true>(k, k_next, kappa, push, obj_amp);
        else
            update_row<kFam, false>(k, k_next, kappa, push, obj_amp);
#endif
    }

    if (pairs) {
        violated_scan(L, 1);
        ls.mo_len = L.sh->tmp_i;
    } else if (order == ORDER_PI_SIGN_CHANGE) {
        // "if (!pi_changed && remaining == 0) return 0; return remaining;"
        violated_scan(L, 2);
    } else {
        violated_scan(L, 0);
        ls.oR_len = L.sh->list_len;
    }
    return L.sh->remaining;
}

// @region objective
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
    int pi_in_smem, x_in_smem, r_in_smem, mo_in_smem, z_in_smem, exact_in_smem;
    int sort_cap;
    int stage_cap;
};

// Bytes of the per-lane row scratch (key0, sk, pown, sid), laid out the
// same way in shared memory or in the batch's global scratch.
__host__ __device__ inline size_t
row_scratch_bytes(int max_row, int sort_cap, int stage_cap)
{
    size_t n = 0;
    n += sizeof(double2) * (size_t)stage_cap; // stage: {P term, pi term}
    n = (n + 15) & ~size_t(15);
    n += sizeof(unsigned long long) * (size_t)max_row;  // key0
    n += sizeof(unsigned long long) * (size_t)sort_cap; // sk
    n += sizeof(double) * (size_t)(max_row + 1);        // pown
    n += sizeof(int) * (size_t)sort_cap;                // sid
    return (n + 15) & ~size_t(15);
}

// @region lane_open_close
// Builds the lane's control block in shared memory (its fields are uniform
// over the CTA) and stages pi, x, the row activities and R's leftover values.
__device__ Lane&
lane_open(Lists& ls, const DevProblem& pb, const DevParams& pr, const DevBatch& b, int lane_id,
          SmemPlan plan, unsigned char* smem)
{
    const int tid = threadIdx.x, T = blockDim.x;
    size_t off = 0;

    Lane& L = *reinterpret_cast<Lane*>(smem);
    off += kLaneBytes;
    Shared* sh = reinterpret_cast<Shared*>(smem + off);
    off += align16(sizeof(Shared));

    double* g_pi = b.pi + (size_t)lane_id * pb.m;
    uint32_t* g_x = b.x + (size_t)lane_id * pb.xw32;
    double* g_Rv = b.Rv + (size_t)lane_id * (pb.max_row + 2);
    const int* g_act = b.act + (size_t)lane_id * pb.m;
    LaneCtl* ctl = b.ctl + lane_id;

    unsigned char* scratch;
    if (plan.r_in_smem) {
        scratch = smem + off;
        off += row_scratch_bytes(pb.max_row, plan.sort_cap, plan.stage_cap);
    } else
        scratch = b.g_scratch + (size_t)lane_id * row_scratch_bytes(pb.max_row, plan.sort_cap, plan.stage_cap);

    double* s_Rv = g_Rv;
    if (plan.r_in_smem && pb.tail_rows) {
        s_Rv = reinterpret_cast<double*>(smem + off);
        off += sizeof(double) * (size_t)(pb.max_row + 2);
        for (int i = tid; i < pb.max_row + 2; i += T)
            s_Rv[i] = g_Rv[i];
    }

    double* s_pi = g_pi;
    if (plan.pi_in_smem) {
        s_pi = reinterpret_cast<double*>(smem + off);
        off += sizeof(double) * (size_t)pb.m;
        for (int i = tid; i < pb.m; i += T)
            s_pi[i] = g_pi[i];
    }

    int* s_act = reinterpret_cast<int*>(smem + off);
    off += sizeof(int) * (size_t)pb.m;
    for (int i = tid; i < pb.m; i += T)
        s_act[i] = g_act[i];

    uint32_t* s_x = g_x;
    if (plan.x_in_smem) {
        s_x = reinterpret_cast<uint32_t*>(smem + off);
        off += sizeof(uint32_t) * (size_t)pb.xw32;
        for (int i = tid; i < pb.xw32; i += T)
            s_x[i] = g_x[i];
    }

    // workspace of the single-thread exact ordering (rows whose scratch is in
    // global memory: every step of that thread would be an L2 round trip)
    unsigned long long* s_xk = nullptr;
    int* s_xid = nullptr;
    if (plan.exact_in_smem) {
        off = (off + 15) & ~size_t(15);
        s_xk = reinterpret_cast<unsigned long long*>(smem + off);
        off += sizeof(unsigned long long) * (size_t)plan.exact_in_smem;
        s_xid = reinterpret_cast<int*>(smem + off);
        off += sizeof(int) * (size_t)plan.exact_in_smem;
    }

    // (Z) items / node stack / subvector pool of the sub-solvers: one thread
    // walks them node by node, so they belong in shared memory when they fit
    unsigned char* s_zs = nullptr;
    if (plan.z_in_smem) {
        off = (off + 15) & ~size_t(15);
        s_zs = smem + off;
        off += b.z_bytes;
    }

    uint32_t* s_mop = nullptr;
    if (plan.mo_in_smem) {
        s_mop = reinterpret_cast<uint32_t*>(smem + off);
        off += sizeof(uint32_t) * (size_t)(pb.m + 1);
        const int2* g_mo = b.mo + (size_t)lane_id * b.order_cap;
        const int n_mo = min(ctl->mo_len, pb.m);
        for (int i = tid; i < n_mo; i += T)
            s_mop[i] = ((uint32_t)g_mo[i].y << 16) | (uint32_t)g_mo[i].x;
    }

    if (tid == 0) {
        L.pb = &pb;
        L.pr = &pr;
        L.mop = s_mop;
        L.xk = s_xk;
        L.xid = s_xid;
        L.xcap = s_xk ? plan.exact_in_smem : 0;
        L.xchunk = 0;
        while (L.xcap >= 2 && (2 << L.xchunk) <= L.xcap)
            ++L.xchunk;
        L.xchunk = L.xcap >= 2 ? 1 << L.xchunk : 0;
        L.tie_streak = 0;
        L.row_hdr = pb.row_hdr;
        L.row_info = pb.row_info;
        L.row_eptr = pb.row_eptr;
        L.gp = pb.gp;
        L.row_a = pb.row_a;
        L.csc_rs = pb.csc_rs;
        L.csc_a = pb.csc_a;
        L.cost = pr.cost;
        const double* par = b.lane_par ? b.lane_par + 5 * (size_t)lane_id : nullptr;
        L.theta = par ? par[0] : pr.theta;
        L.delta = par ? par[1] : pr.delta;
        L.kappa_min = par ? par[2] : pr.kappa_min;
        L.kappa_step = par ? par[3] : pr.kappa_step;
        L.maximize = pb.maximize;
        L.tail_rows = pb.tail_rows;
        L.sh = sh;
        L.P = b.P + (size_t)lane_id * pb.nnz;
        L.ctl = ctl;
        L.zs = s_zs ? s_zs : (b.z_scratch ? b.z_scratch + (size_t)lane_id * b.z_bytes : nullptr);
        L.oR = b.oR + (size_t)lane_id * b.order_cap;
        L.mo = b.mo + (size_t)lane_id * b.order_cap;
        L.sort_cap = plan.sort_cap; // (the CTA-wide sort also runs on a scratch in global memory)
        L.all_smem = plan.r_in_smem && plan.pi_in_smem && plan.x_in_smem;
        int bits = 1;
        while ((1 << bits) < 2 * plan.sort_cap)
            ++bits;
        L.tab_bits = bits;
        L.tab_gen = 0;
        L.tab_dirty = 1;

        size_t o = 0;
        L.stage = reinterpret_cast<double*>(scratch + o);
        L.stage_cap = plan.stage_cap;
        o += sizeof(double2) * (size_t)plan.stage_cap;
        o = (o + 15) & ~size_t(15);
        L.key0 = reinterpret_cast<unsigned long long*>(scratch + o);
        o += sizeof(unsigned long long) * (size_t)pb.max_row;
        L.sk = reinterpret_cast<unsigned long long*>(scratch + o);
        o += sizeof(unsigned long long) * (size_t)plan.sort_cap;
        L.pown = reinterpret_cast<double*>(scratch + o);
        o += sizeof(double) * (size_t)(pb.max_row + 1);
        L.sid = reinterpret_cast<int*>(scratch + o);

        L.Rv = s_Rv;
        L.pi = s_pi;
        L.act = s_act;
        L.x = s_x;
        // offsets of the shared-memory views (meaningful when all_smem)
        L.o_stage = (int)(reinterpret_cast<unsigned char*>(L.stage) - smem);
        L.o_key0 = (int)(reinterpret_cast<unsigned char*>(L.key0) - smem);
        L.o_sk = (int)(reinterpret_cast<unsigned char*>(L.sk) - smem);
        L.o_pown = (int)(reinterpret_cast<unsigned char*>(L.pown) - smem);
        L.o_pi = (int)(reinterpret_cast<unsigned char*>(s_pi) - smem);
        L.o_act = (int)(reinterpret_cast<unsigned char*>(s_act) - smem);
        L.o_x = (int)(reinterpret_cast<unsigned char*>(s_x) - smem);

        sh->rng = ctl->rng;
        sh->order_state = ctl->order_state;
        sh->pi_changed = 0;
        sh->remaining = ctl->remaining;
        sh->stop = 0;
        sh->n_dup = 0;
        sh->x_dirty = 1;
        L.updates = 0;
        L.sweeps = 0;
        L.offers = 0;
        L.inserts = 0;
        L.alg_bytes = 0.0;
        for (int i = 0; i < 4; ++i) {
            L.n_path[i] = 0;
            L.cyc_path[i] = 0;
            L.cyc_phase[i] = 0;
        }
    }
    ls.oR_len = ctl->oR_len;
    ls.mo_len = ctl->mo_len;
    __syncthreads();
    return L;
}

__device__ void
lane_close(Lane& L, Lists& ls, const DevProblem& pb, const DevBatch& b,
           int lane_id, SmemPlan plan)
{
    const int tid = threadIdx.x, T = blockDim.x;
    __syncthreads();

    if (plan.r_in_smem && pb.tail_rows) {
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
    {
        int* g_act = b.act + (size_t)lane_id * pb.m;
        for (int i = tid; i < pb.m; i += T)
            g_act[i] = L.act[i];
    }
    if (plan.mo_in_smem) {
        int2* g_mo = b.mo + (size_t)lane_id * b.order_cap;
        const int n_mo = min(ls.mo_len, pb.m);
        for (int i = tid; i < n_mo; i += T)
            g_mo[i] = make_int2((int)(L.mop[i] & 0xffffu), (int)(L.mop[i] >> 16));
    }
    if (tid == 0) {
        L.ctl->rng = L.sh->rng;
        L.ctl->order_state = L.sh->order_state;
        L.ctl->oR_len = ls.oR_len;
        L.ctl->mo_len = ls.mo_len;
        L.ctl->remaining = L.sh->remaining;
        L.ctl->updates += L.updates;
        L.ctl->sweeps += L.sweeps;
        L.ctl->offers += L.offers;
        L.ctl->inserts += L.inserts;
        L.ctl->alg_bytes += L.alg_bytes;
        for (int i = 0; i < 4; ++i) {
            L.ctl->n_path[i] += L.n_path[i];
            L.ctl->cyc_path[i] += L.cyc_path[i];
            L.ctl->cyc_phase[i] += L.cyc_phase[i];
        }
    }
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

// @region kernels
// Start of a run: engine, solver state, x (init policy bastert + Bernoulli
// flips on the device, or a caller-provided x0), compute_order::init.
// solver_functor::operator(), itm-solver-common.hpp:94-133.
template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_begin(const __grid_constant__ DevProblem pb, const __grid_constant__ DevParams pr,
        const __grid_constant__ DevBatch b, const uint32_t* seeds,
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

    const double lane_init_random = b.lane_par ? b.lane_par[5 * (size_t)lane_id + 4] : init_random;
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
                    if (canonical_from(g1, g2) < lane_init_random)
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

    Lists ls;
    Lane& L = lane_open(ls, pb, pr, b, lane_id, plan, g_smem);
    order_init(L, ls);
    lane_close(L, ls, pb, b, lane_id, plan);
}

// compute_order::init on a range of lanes (after bxg_batch_set_state).
template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_order_init(const __grid_constant__ DevProblem pb, const __grid_constant__ DevParams pr,
        const __grid_constant__ DevBatch b, int lane_first,
             SmemPlan plan)
{
    const int lane_id = lane_first + blockIdx.x;
    Lists ls;
    Lane& L = lane_open(ls, pb, pr, b, lane_id, plan, g_smem);
    order_init(L, ls);
    lane_close(L, ls, pb, b, lane_id, plan);
}

// One compute_order::run / push_and_run per lane.
template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_sweep(const __grid_constant__ DevProblem pb, const __grid_constant__ DevParams pr,
        const __grid_constant__ DevBatch b, int push, double obj_amp,
        SmemPlan plan)
{
    const int lane_id = blockIdx.x;
    Lists ls;
    Lane& L = lane_open(ls, pb, pr, b, lane_id, plan, g_smem);
    const double kappa = b.kappa[lane_id];
    order_run(L, ls, kappa, push, obj_amp);
    lane_close(L, ls, pb, b, lane_id, plan);
}

template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_objective(const __grid_constant__ DevProblem pb, const __grid_constant__ DevParams pr,
        const __grid_constant__ DevBatch b, SmemPlan plan)
{
    const int lane_id = blockIdx.x;
    Lists ls;
    Lane& L = lane_open(ls, pb, pr, b, lane_id, plan, g_smem);
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

// ---------------------------------------------------------------------------
// optimizer: offers to the population and one call of a worker's loop body
// ---------------------------------------------------------------------------

// storage::sort's comparator, itm-optimizer-common.hpp:427-441
struct PopLess
{
    const int* rem;
    const double* val;
    bool maximize;
    __device__ bool operator()(const int& a, const int& b) const
    {
        const int ra = rem[a], rb = rem[b];
        if (ra < rb)
            return true;
        if (ra == rb)
            return maximize ? val[a] > val[b] : val[a] < val[b];
        return false;
    }
};

template<typename T>
__device__ __forceinline__ T
ld_vol(const T* p)
{
    return *reinterpret_cast<const volatile T*>(p);
}

// @region pop_offer
// best_solution_recorder::try_update (solution) / try_advance (:556-584) with
// the lane's current x: bit_array_hash + objective (kept while x does not
// change), can_be_inserted (:302-326), and when the population takes it
// storage::insert (:228-300): the victim drawn from the lane's engine --
// the same engine the sweeps draw from, as in the reference, where
// local_context::rng is the engine the Solver holds -- replace_result, sort.
// All threads call; the population is locked for the duration.
__device__ __noinline__ void
pop_offer(Lane& L, const DevBatch& b, bool solution, int remaining, long long loop)
{
    const DevProblem& pb = *L.pb;
    const DevPop& pop = b.pop;
    Shared* const sh = L.sh;
    const int tid = threadIdx.x, T = blockDim.x;

    if (sh->x_dirty) { // uniform: written before the barrier that ended the sweep
        const double v = objective(L);
        if (tid < 32) {
            // bit_array_hash, bit-array.hpp:410-423: a sequential fold over the
            // n/64+1 blocks; the warp fetches 32 blocks at a time
            const int blocks = pb.xw32 >> 1;
            unsigned long long h = 0;
            for (int base = 0; base < blocks; base += 32) {
                const int i = base + tid;
                unsigned long long w = 0;
                if (i < blocks)
                    w = (unsigned long long)L.x[2 * i] | ((unsigned long long)L.x[2 * i + 1] << 32);
                const int cnt = min(32, blocks - base);
                for (int j = 0; j < cnt; ++j) {
                    const unsigned long long bj = __shfl_sync(0xffffffffu, w, j);
                    h ^= bj + 0x9e3779b9ull + (h << 6) + (h >> 2);
                }
            }
            if (tid == 0) {
                sh->x_hash = h;
                sh->x_value = v;
                sh->x_dirty = 0;
            }
        }
        __syncthreads();
    }
    const double value = sh->x_value;
    const unsigned long long hash = sh->x_hash;

    // can_be_inserted (:302-326) reads the scores without the writers' lock,
    // like the reference (a shared lock there): a lane in the feasible regime
    // offers the same x sweep after sweep and must not queue for that
    int found = 0;
    for (int i = tid; i < pop.size; i += T) {
        const int r = ld_vol(pop.remaining + i);
        const unsigned long long h = ld_vol(pop.hash + i);
        if (solution)
            found |= (r == 0 && ld_vol(pop.value + i) == value && h == hash);
        else
            found |= (r == remaining && h == hash);
    }
    found = __syncthreads_or(found);
    if (tid == 0)
        L.offers += 1;
    if (found)
        return;

    if (tid == 0) {
        while (atomicCAS(pop.lock, 0, 1) != 0)
            __nanosleep(64);
        __threadfence();
    }
    __syncthreads();

    {
        if (tid == 0) {
            // choose_a_bad_solution: uniform_int_distribution<int>(size / 5, size - 1)
            uint32_t rng = sh->rng;
            const int pos = pop.victim_lo + (int)uniform_below(rng, (uint32_t)(pop.victim_hi - pop.victim_lo));
            sh->rng = rng;
            const int id = ld_vol(pop.order + pos);
            sh->tmp_i = id;
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            // replace_result, :127-144 (a solution keeps the value it was offered with;
            // an advance stores costs.results(x))
            pop.value[id] = value;
            pop.remaining[id] = solution ? 0 : remaining;
            pop.hash[id] = hash;
            pop.loop[id] = loop;
            pop.duration[id] = (double)(now - pop.t0_ns) * 1e-9;
            if (solution)
                atomicMin(pop.first_ns, now - pop.t0_ns);
            L.inserts += 1;
        }
        __syncthreads();
        uint32_t* dst = pop.x + (size_t)sh->tmp_i * pb.xw32;
        for (int w = tid; w < pb.xw32; w += T)
            dst[w] = L.x[w];

        // sort(): keys staged in the lane's staging buffer (idle between sweeps)
        double* s_val = L.stage;
        int* s_rem = reinterpret_cast<int*>(s_val + pop.size);
        int* s_ord = s_rem + pop.size;
        for (int i = tid; i < pop.size; i += T) {
            s_val[i] = ld_vol(pop.value + i);
            s_rem[i] = ld_vol(pop.remaining + i);
            s_ord[i] = ld_vol(pop.order + i);
        }
        __syncthreads();
        if (tid == 0)
            seq_std_sort(Plain<int>{ s_ord }, pop.size, PopLess{ s_rem, s_val, pb.maximize != 0 });
        __syncthreads();
        for (int i = tid; i < pop.size; i += T)
            pop.order[i] = s_ord[i];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0)
        atomicExch(pop.lock, 0);
}

// @region k_optimize
// One call of the loop body of optimize_functor::operator()
// (itm-optimizer-common.hpp:658-750) for one lane = one worker: compute.init,
// the sweeps with the kappa schedule from the kappa reinit returned, then the
// push phase; every solution met is offered to the population where the
// reference offers it, the run's last x with the best count otherwise.  x was
// prepared by k_reinit; the Solver state (P, pi) carries over from the
// previous call.  The deadline plays stop_task.
template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_optimize(const __grid_constant__ DevProblem pb, const __grid_constant__ DevParams pr,
           const __grid_constant__ DevBatch b, SmemPlan plan)
{
    const int lane_id = blockIdx.x;
    Lists ls;
    Lane& L = lane_open(ls, pb, pr, b, lane_id, plan, g_smem);
    Shared* const sh = L.sh;
    order_init(L, ls);

    const long w_limit = (long)pr.w;
    const double kappa_start = b.plan[lane_id].kappa;
    const double kappa_step = L.kappa_step;
    double kappa = kappa_start;
    int best_remaining = INT_MAX;
    int is_solution = 0;
    double last_value = 0.0;

    for (long long i = 0; !sh->stop && i != pr.limit; ++i) {
        const int remaining = order_run(L, ls, kappa, 0, 0.0);
        if (remaining == 0) {
            pop_offer(L, b, true, 0, i);
            best_remaining = 0;
            is_solution = 1;
            break;
        }
        best_remaining = remaining < best_remaining ? remaining : best_remaining;
        if (i > w_limit)
            kappa = __dadd_rn(kappa, __dmul_rn(kappa_step, pr.powtab[remaining]));
        if (kappa > pr.kappa_max)
            break;
    }

    if (best_remaining > 0) {
        pop_offer(L, b, false, best_remaining, pr.limit);
    } else {
        for (int push = 0; !sh->stop && push < pr.pushes_limit; ++push) {
            int remaining =
              order_run(L, ls, pr.pushing_k_factor, 1, pr.pushing_objective_amplifier);
            if (remaining == 0)
                pop_offer(L, b, true, 0, -(long long)push * pr.pushing_iteration_limit - 1);

            kappa = kappa_start;
            for (int iter = 0; !sh->stop && iter < pr.pushing_iteration_limit; ++iter) {
                remaining = order_run(L, ls, kappa, 0, 0.0);
                if (remaining == 0) {
                    pop_offer(L, b, true, 0,
                              -(long long)push * pr.pushing_iteration_limit - iter - 1);
                    break;
                }
                if ((double)iter > pr.w)
                    kappa = __dadd_rn(kappa, __dmul_rn(kappa_step, pr.powtab[remaining]));
                if (kappa > pr.kappa_max)
                    break;
            }
        }
    }
    if (is_solution)
        last_value = sh->x_dirty ? 0.0 : sh->x_value;

    if (threadIdx.x == 0) {
        L.ctl->status = sh->stop ? 3 : (is_solution ? 0 : 5);
        L.ctl->best_remaining = best_remaining;
        L.ctl->best_has_x = is_solution;
        L.ctl->best_loop = 0;
        L.ctl->best_value = last_value;
    }
    lane_close(L, ls, pb, b, lane_id, plan);
}

// Whole solver_functor::operator() loop (itm-solver-common.hpp:135-225) for
// one lane; the wall-clock time limit is handled by the host (limit sweeps).
template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_solve(const __grid_constant__ DevProblem pb, const __grid_constant__ DevParams pr,
        const __grid_constant__ DevBatch b, SmemPlan plan, int optimizer)
{
    const int lane_id = blockIdx.x;
    Lists ls;
    Lane& L = lane_open(ls, pb, pr, b, lane_id, plan, g_smem);

    int status = 2;
    int m_best_remaining = INT_MAX; // m_best.remaining_constraints
    double m_best_value = pb.maximize ? -CUDART_INF : CUDART_INF;
    long long m_best_loop = 0;
    int m_best_has_x = 0;
    int best_remaining = INT_MAX;

    const long w_limit = (long)pr.w;
    bool start_push = false;
    // optimizer = 1: one call of a worker of optimize_functor::operator()
    // (itm-optimizer-common.hpp:658-750) instead of solver_functor's run: the
    // restart begins at the kappa best_solution_recorder::reinit returned, the
    // push sweep runs at kappa = pushing_k_factor and every push restarts the
    // schedule from that kappa
    const double kappa_start = optimizer ? b.kappa[lane_id] : L.kappa_min;
    const double kappa_step = L.kappa_step;
    double kappa = kappa_start;

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
            kappa = __dadd_rn(kappa, __dmul_rn(kappa_step, pr.powtab[remaining]));

        if (kappa > pr.kappa_max) {
            status = 4; // kappa_max_reached
            break;
        }
        if (L.sh->stop) { // is_timelimit_reached(), itm-solver-common.hpp:162-165
            status = 3;
            break;
        }
    }

    if (!start_push) {
        // (the reference overwrites kappa_max_reached / time_limit_reached here
        // too: itm-solver-common.hpp:169-170)
        status = 5; // limit_reached
    } else {
        for (int push = 0; push < pr.pushes_limit; ++push) {
            int remaining =
              order_run(L, ls, optimizer ? pr.pushing_k_factor : __dmul_rn(pr.pushing_k_factor, kappa),
                        1, pr.pushing_objective_amplifier);

            if (remaining == 0) {
                const double v = objective(L);
                store_best(L, b, lane_id, true, v, 0,
                           -(long long)push * pr.pushing_iteration_limit - 1,
                           m_best_remaining, m_best_value, m_best_loop, m_best_has_x);
            }
            if (L.sh->stop)
                break;

            if (optimizer)
                kappa = kappa_start;
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
                    kappa = __dadd_rn(kappa, __dmul_rn(kappa_step, pr.powtab[remaining]));

                if (kappa > pr.kappa_max)
                    break;
                if (L.sh->stop)
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
#if BXG_FAM == 2
static int
pow2_at_least(int v)
{
    int p = 2;
    while (p < v)
        p <<= 1;
    return p;
}

size_t z_scratch_bytes(int max_zrow);

static size_t
smem_bytes(const DevProblem& pb, int pi_in, int x_in, int r_in, int sort_cap, int stage_cap, int mo_in = 0,
           int z_in = 0, int exact_in = 0)
{
    size_t off = kLaneBytes + align16(sizeof(Shared));
    if (r_in) {
        off += row_scratch_bytes(pb.max_row, sort_cap, stage_cap);
        if (pb.tail_rows)
            off += sizeof(double) * (size_t)(pb.max_row + 2);
    }
    if (pi_in)
        off += sizeof(double) * (size_t)pb.m;
    off += sizeof(int) * (size_t)pb.m;
    if (x_in)
        off += sizeof(uint32_t) * (size_t)pb.xw32;
    if (exact_in) // (a capacity in elements)
        off += 16 + 12 * (size_t)exact_in;
    if (z_in)
        off += 16 + z_scratch_bytes(pb.max_zrow);
    if (mo_in)
        off += sizeof(uint32_t) * (size_t)(pb.m + 1);
    return off + 16;
}

size_t
z_scratch_bytes(int max_zrow)
{
    if (max_zrow <= 0)
        return 0;
    const size_t n = z_items_bytes(max_zrow) + z_nodes_bytes(max_zrow) + z_pool_bytes(max_zrow);
    return (n + 15) & ~size_t(15);
}

size_t
lane_scratch_bytes(const DevProblem& pb, const LaunchCfg& cfg)
{
    return row_scratch_bytes(pb.max_row, cfg.sort_cap, cfg.stage_cap);
}

// CTAs per SM the register budget of a CTA size class allows (the MINB of
// its __launch_bounds__, see BXG_LAUNCH_HOT).
static int
cta_class_blocks(int threads)
{
    return threads <= 96 ? 8 : threads <= 128 ? 7 : threads <= 192 ? 5 : threads <= 256 ? 3 : 1;
}

LaunchCfg
plan_launch(const DevProblem& pb, int device, int n_lanes, int pairs_order)
{
    int max_optin = 0, per_sm = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    if (max_optin <= 0)
        max_optin = 48 * 1024;
    if (per_sm <= 0)
        per_sm = max_optin;
    const size_t budget = (size_t)max_optin - 1024;

    LaunchCfg cfg;
    cfg.r_in_smem = 1;
    cfg.pi_in_smem = 1;
    cfg.x_in_smem = 1;
    // the sorted (row, amount) list of the infeasibility orders: packed in shared
    // memory when rows and amounts fit 16 bits each
    cfg.mo_in_smem = pairs_order && pb.m <= 0xffff && pb.max_amount <= 0xffff;
    cfg.z_in_smem = pb.max_zrow > 0 && z_scratch_bytes(pb.max_zrow) <= 8192;
    cfg.exact_in_smem = 0;
    cfg.sort_cap = pow2_at_least(pb.max_row);
    // staging buffer of the reduced-cost gather: at least one column, at most
    // the longest row's entries
    const int stage_min = pb.max_col > 256 ? pb.max_col : 256;
    const int stage_max = pb.max_S < stage_min ? stage_min : pb.max_S;
    cfg.stage_cap = stage_min;
    auto bytes = [&]() {
        return smem_bytes(pb, cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.sort_cap,
                          cfg.stage_cap, cfg.mo_in_smem, cfg.z_in_smem, cfg.exact_in_smem);
    };
    // with the smallest staging buffer, drop the largest consumers first until
    // the plan fits
    if (bytes() > budget)
        cfg.mo_in_smem = 0;
    if (bytes() > budget)
        cfg.x_in_smem = 0;
    if (bytes() > budget) {
        // rows too long for a shared-memory scratch: it goes to global memory
        // (L2-resident, a few hundred KB per lane); the CTA-wide sort runs there
        cfg.r_in_smem = 0;
    }
    if (bytes() > budget)
        cfg.pi_in_smem = 0;
    if (!cfg.r_in_smem) {
        cfg.stage_cap = stage_max; // the scratch lives in global memory
        // ... except a copy of the keys for the single-thread exact ordering, when it fits
        // (a capacity in elements: the whole row when it fits, else what is left
        // -- the chunk buffer of the CTA-wide sort, cta_bitonic_chunked)
        const size_t fixed = bytes();
        const size_t room = fixed + 64 < budget ? (budget - fixed - 64) / 12 : 0;
        cfg.exact_in_smem = room >= (size_t)pb.max_row ? pb.max_row : (room >= 1024 ? (int)room : 0);
    }
    else {
        // then let the staging buffer grow (rows are summed in chunks when it
        // is shorter than the row's expansion): up to what lets seven lanes
        // share an SM -- the registers of the 128-thread class allow as many --
        // or, when the rest alone is past that, up to the CTA's limit
        const size_t target = (size_t)per_sm / 7 - 1024;
        const size_t fixed = bytes();
        const size_t limit = fixed < target ? target : budget;
        const size_t room = fixed < limit ? limit - fixed : 0;
        const size_t grow = room / sizeof(double2);
        cfg.stage_cap = (size_t)stage_min + grow > (size_t)stage_max ? stage_max
                                                                     : stage_min + (int)grow;
    }
    cfg.smem = bytes();

    // One thread per row element in the second gather stage (whole warps),
    // narrowed while that lets every lane of the batch be resident at once:
    // a second wave of CTAs costs more than a second trip over a long row.
    // BXG_THREADS overrides (experiments).
    int t = ((pb.p95_row + 31) / 32) * 32;
    if (t < 64)
        t = 64;
    // branch-and-bound rows are enumerated by the whole CTA (z_bb_enumerate)
    if (pb.max_zrow >= 10 && t < 128)
        t = 128;
    if (t > 1024)
        t = 1024;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (sms <= 0)
        sms = 148;
    auto resident = [&](int threads) {
        const int by_regs = cta_class_blocks(threads);
        const int by_smem = (int)((size_t)per_sm / (cfg.smem + 1024));
        return (long)sms * (by_regs < by_smem ? by_regs : (by_smem < 1 ? 1 : by_smem));
    };
    while (t > 128 && n_lanes > resident(t) && resident(t - 32) > resident(t))
        t -= 32;
    if (const char* env = getenv("BXG_THREADS"))
        t = (atoi(env) + 31) / 32 * 32;
    if (t < 32)
        t = 32;
    if (t > 1024)
        t = 1024;
    cfg.threads = t;

    return cfg;
}

#endif // BXG_FAM == 2 (the shared host helpers live in one translation unit)

#define BXG_CAT2(a, b) a##_f##b
#define BXG_CAT(a, b) BXG_CAT2(a, b)
#define BXG_FAM_FN(name) BXG_CAT(name, BXG_FAM)

template<typename K>
static void
set_smem(K kernel, size_t smem)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
}

// CTA size classes: the sweep kernels are compiled once per class so that the
// register budget follows the CTA size (MINB CTAs of MAXT threads share the
// 64 K registers of an SM); the kernels outside the sweep loop are compiled
// once, for the largest CTA.
#define BXG_LAUNCH_CLASS(KERNEL, MAXT, MINB, GRID, CFG, STREAM, ...)                         \
    do {                                                                                    \
        set_smem(KERNEL<MAXT, MINB>, (CFG).smem);                                           \
        KERNEL<MAXT, MINB><<<(GRID), (CFG).threads, (CFG).smem, (STREAM)>>>(__VA_ARGS__);   \
    } while (0)

#define BXG_LAUNCH_HOT(KERNEL, GRID, CFG, STREAM, ...)                                      \
    do {                                                                                    \
        if ((CFG).threads <= 96)                                                            \
            BXG_LAUNCH_CLASS(KERNEL, 96, 8, GRID, CFG, STREAM, __VA_ARGS__);                \
        else if ((CFG).threads <= 128)                                                      \
            BXG_LAUNCH_CLASS(KERNEL, 128, 7, GRID, CFG, STREAM, __VA_ARGS__);               \
        else if ((CFG).threads <= 160 && BXG_CLASS_160)                                     \
            BXG_LAUNCH_CLASS(KERNEL, 160, 7, GRID, CFG, STREAM, __VA_ARGS__);               \
        else if ((CFG).threads <= 192)                                                      \
            BXG_LAUNCH_CLASS(KERNEL, 192, 5, GRID, CFG, STREAM, __VA_ARGS__);               \
        else if ((CFG).threads <= 256)                                                      \
            BXG_LAUNCH_CLASS(KERNEL, 256, 3, GRID, CFG, STREAM, __VA_ARGS__);               \
        else if ((CFG).threads <= 512)                                                      \
            BXG_LAUNCH_CLASS(KERNEL, 512, 1, GRID, CFG, STREAM, __VA_ARGS__);               \
        else                                                                                \
            BXG_LAUNCH_CLASS(KERNEL, 1024, 1, GRID, CFG, STREAM, __VA_ARGS__);              \
    } while (0)

#define BXG_LAUNCH_COLD(KERNEL, GRID, CFG, STREAM, ...)                                     \
    BXG_LAUNCH_CLASS(KERNEL, 1024, 1, GRID, CFG, STREAM, __VA_ARGS__)

void
BXG_FAM_FN(launch_begin)(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
             const uint32_t* d_seeds, const uint32_t* d_x0, const uint32_t* d_rng,
             int reset_solver, int init_on_device, double init_random,
             const double* d_cost_orig, const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.mo_in_smem, cfg.z_in_smem, cfg.exact_in_smem, cfg.sort_cap, cfg.stage_cap };
    BXG_LAUNCH_COLD(k_begin, b.n_lanes, cfg, s, pb, pr, b, d_seeds, d_x0, d_rng, reset_solver,
               init_on_device, init_random, d_cost_orig, plan);
}

void
BXG_FAM_FN(launch_order_init)(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
                  int lane_first, int lane_count, const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.mo_in_smem, cfg.z_in_smem, cfg.exact_in_smem, cfg.sort_cap, cfg.stage_cap };
    BXG_LAUNCH_COLD(k_order_init, lane_count, cfg, s, pb, pr, b, lane_first, plan);
}

void
BXG_FAM_FN(launch_sweep)(const DevProblem& pb, const DevParams& pr, const DevBatch& b, int push,
             double obj_amp, const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.mo_in_smem, cfg.z_in_smem, cfg.exact_in_smem, cfg.sort_cap, cfg.stage_cap };
    BXG_LAUNCH_HOT(k_sweep, b.n_lanes, cfg, s, pb, pr, b, push, obj_amp, plan);
}

void
BXG_FAM_FN(launch_solve)(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
             const LaunchCfg& cfg, cudaStream_t s, int optimizer)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.mo_in_smem, cfg.z_in_smem, cfg.exact_in_smem, cfg.sort_cap, cfg.stage_cap };
    BXG_LAUNCH_HOT(k_solve, b.n_lanes, cfg, s, pb, pr, b, plan, optimizer);
}

void
BXG_FAM_FN(launch_optimize)(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
                const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.mo_in_smem, cfg.z_in_smem, cfg.exact_in_smem, cfg.sort_cap, cfg.stage_cap };
    BXG_LAUNCH_HOT(k_optimize, b.n_lanes, cfg, s, pb, pr, b, plan);
}

void
BXG_FAM_FN(launch_objective)(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
                 const LaunchCfg& cfg, cudaStream_t s)
{
    SmemPlan plan{ cfg.pi_in_smem, cfg.x_in_smem, cfg.r_in_smem, cfg.mo_in_smem, cfg.z_in_smem, cfg.exact_in_smem, cfg.sort_cap, cfg.stage_cap };
    BXG_LAUNCH_COLD(k_objective, b.n_lanes, cfg, s, pb, pr, b, plan);
}

} // namespace bxg
