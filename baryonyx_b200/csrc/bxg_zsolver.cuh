// bxg_zsolver.cuh -- the sub-solvers of the Z-coefficient rows, ONE thread.
//
// Rows of the Z solver that carry a factor outside {-1, 0, 1} pick their
// variables with an exact sub-solver instead of the prefix rule
// (itm-solver-inequalities-Z.cpp:494-546):
//   * exhaustive_solver (lib/src/exhaustive-solver.hpp:40-253) for rows with
//     fewer than 10 elements: the feasible 0/1 patterns of the row are listed
//     once per problem (on the host, bxg_api.cu) and every update takes the
//     pattern with the best sum of reduced costs;
//   * branch_and_bound_solver (lib/src/branch-and-bound-solver.hpp:192-534)
//     for longer rows: a depth-first enumeration over the items in ratio
//     order, cut only by the factor sum.
// Both re-order the row's scratch so that the chosen elements come first and
// return the split index.  The enumeration order, the comparators and the
// std::sort calls decide ties, so they are restated exactly; what the
// reference computes but never reads (node bounds z, upper_bound) is left out.
//
// The subvectors of the reference's node stack (one int8 per item) are bit
// sets here, one slot per stack position: a node owns exactly one subvector
// and nodes are only ever removed from the top.
#ifndef BXG_ZSOLVER_CUH
#define BXG_ZSOLVER_CUH

#include "bxg_device.cuh"

#include <math_constants.h>

namespace bxg {

struct ZItem
{
    double r;
    double sum_z;
    int factor;
    int original_factor;
    int variable;
    int result;
};

struct ZNode
{
    double sumr;
    int sumfactor;
    int level;
};

__host__ __device__ inline int
z_words(int max_zrow)
{
    return (max_zrow + 63) / 64;
}

__host__ __device__ inline size_t
z_items_bytes(int max_zrow)
{
    return sizeof(ZItem) * (size_t)max_zrow;
}

__host__ __device__ inline size_t
z_nodes_bytes(int max_zrow)
{
    return sizeof(ZNode) * (size_t)(max_zrow + 2);
}

__host__ __device__ inline size_t
z_pool_bytes(int max_zrow)
{
    // one slot per stack position + the incumbent
    return sizeof(unsigned long long) * (size_t)(max_zrow + 3) * (size_t)z_words(max_zrow);
}

struct ZWork
{
    ZItem* items;
    ZNode* nodes;
    unsigned long long* pool;
    int words;     // 64-bit words per subvector
    int solution;  // slot of the incumbent

    __device__ ZWork(unsigned char* base, int max_zrow)
    {
        items = reinterpret_cast<ZItem*>(base);
        nodes = reinterpret_cast<ZNode*>(base + z_items_bytes(max_zrow));
        pool = reinterpret_cast<unsigned long long*>(base + z_items_bytes(max_zrow) +
                                                     z_nodes_bytes(max_zrow));
        words = z_words(max_zrow);
        solution = max_zrow + 2;
    }

    __device__ __forceinline__ unsigned long long* slot(int s) const
    {
        return pool + (size_t)s * words;
    }
    __device__ __forceinline__ void clear(int s) const
    {
        for (int w = 0; w < words; ++w)
            slot(s)[w] = 0ull;
    }
    __device__ __forceinline__ void copy(int src, int dst) const
    {
        for (int w = 0; w < words; ++w)
            slot(dst)[w] = slot(src)[w];
    }
    __device__ __forceinline__ void set(int s, int i) const
    {
        slot(s)[i >> 6] |= 1ull << (i & 63);
    }
    __device__ __forceinline__ bool test(int s, int i) const
    {
        return (slot(s)[i >> 6] >> (i & 63)) & 1ull;
    }
};

// operator< of the items, branch-and-bound-solver.hpp:212-220: both ratios
// use the LEFT operand's factor.
struct ZRatioLess
{
    bool maximize;
    __device__ bool operator()(const ZItem& l, const ZItem& r) const
    {
        const double f = (double)l.factor;
        const double a = __ddiv_rn(l.r, f), b = __ddiv_rn(r.r, f);
        return maximize ? b < a : a < b;
    }
};

// chosen items first, then by reduced cost (branch-and-bound-solver.hpp:
// 500-515, exhaustive-solver.hpp:222-237)
struct ZResultLess
{
    bool maximize;
    __device__ bool operator()(const ZItem& l, const ZItem& r) const
    {
        if (l.result == r.result)
            return maximize ? l.r > r.r : l.r < r.r;
        return l.result > r.result;
    }
};

// The row's scratch as the sub-solvers see it: value bits + position in the
// row, the factor looked up through the row's factors.
struct ZRow
{
    unsigned long long* sk;
    int* sid;
    const int* a; // factors of the row, by position
    __device__ __forceinline__ double value(int p) const
    {
        return __longlong_as_double((long long)sk[p]);
    }
    __device__ __forceinline__ int factor(int p) const { return a[sid[p]]; }
};

// Writes the items back in their final order and returns the split index
// (tail of both solve functions).
__device__ inline int
z_finish(const ZWork& w, const ZRow& R, int r_size, bool maximize)
{
    seq_std_sort(Plain<ZItem>{ w.items }, r_size, ZResultLess{ maximize });

    int middle = 0;
    while (middle != r_size && w.items[middle].result != 0)
        ++middle;

    for (int i = 0; i != r_size; ++i) {
        R.sk[i] = (unsigned long long)__double_as_longlong(w.items[i].r);
        R.sid[i] = w.items[i].variable;
    }

    if (middle == r_size)
        return w.items[0].result == 0 ? -1 : r_size;
    return middle - 1;
}

// exhaustive_solver::solve, exhaustive-solver.hpp:185-252.  masks[i] bit j =
// element j is taken by pattern i (positional, whatever order R is in).
__device__ inline int
z_ex_solve(const ZWork& w, const ZRow& R, int r_size, bool maximize, const int* masks, int solutions)
{
    for (int i = 0; i != r_size; ++i) {
        ZItem it;
        it.r = R.value(i);
        it.sum_z = 0.0;
        it.variable = R.sid[i];
        it.factor = it.original_factor = R.factor(i);
        it.result = 0;
        w.items[i] = it;
    }

    double z_best = 0.0;
    int best = 0;
    {
        const int mk = masks[0];
        for (int j = 0; j != r_size; ++j)
            if ((mk >> j) & 1)
                z_best = __dadd_rn(z_best, w.items[j].r);
    }
    for (int i = 1; i < solutions; ++i) {
        const int mk = masks[i];
        double z = 0.0;
        for (int j = 0; j != r_size; ++j)
            if ((mk >> j) & 1)
                z = __dadd_rn(z, w.items[j].r);
        if (maximize ? (z > z_best) : (z < z_best)) {
            z_best = z;
            best = i;
        }
    }

    const int mk = masks[best];
    for (int i = 0; i != r_size; ++i)
        w.items[i].result = (mk >> i) & 1;

    return z_finish(w, R, r_size, maximize);
}

// branch_and_bound_solver::solve, branch-and-bound-solver.hpp:450-533, with
// make_init_node_0/1 (:296-313), make_next_*_node (:315-394) and
// solve_equality / solve_inequality (:396-448).
__device__ inline int
z_bb_solve(const ZWork& w, const ZRow& R, int r_size, bool maximize, int b_min, int b_max)
{
    for (int i = 0; i != r_size; ++i) {
        ZItem it;
        it.r = R.value(i);
        it.sum_z = 0.0;
        it.variable = R.sid[i];
        it.original_factor = R.factor(i);
        it.factor = it.original_factor < 0 ? -it.original_factor : it.original_factor;
        it.result = 0;
        w.items[i] = it;
    }
    w.clear(w.solution);

    seq_std_sort(Plain<ZItem>{ w.items }, r_size, ZRatioLess{ maximize });

    for (int i = 0; i != r_size; ++i) {
        if (w.items[i].original_factor < 0) {
            b_min += w.items[i].factor;
            b_max += w.items[i].factor;
        }
    }

    const bool equality = b_max == b_min;
    double lower_bound = maximize ? -CUDART_INF : CUDART_INF;
    auto is_best = [&](double z) { return maximize ? z > lower_bound : z < lower_bound; };
    auto inside = [&](int s) { return equality ? s == b_min : (b_min <= s && s <= b_max); };

    int n = 0; // nodes on the stack
    bool done = false;

    // node 0: nothing taken
    w.nodes[0] = ZNode{ 0.0, 0, 0 };
    w.clear(0);
    n = 1;
    if (inside(0)) {
        w.copy(0, w.solution);
        if (equality)
            done = true; // solve_equality returns at once
    } else if (!equality && 0 > b_max)
        n = 0;

    if (!done) {
        // node 1: the first item taken
        w.nodes[n] = ZNode{ w.items[0].r, w.items[0].factor, 0 };
        w.clear(n);
        w.set(n, 0);
        ++n;
        const int sf = w.nodes[n - 1].sumfactor;
        if (inside(sf)) {
            w.copy(n - 1, w.solution);
            if (equality)
                --n;
        } else if (!equality && sf > b_max)
            --n;

        while (n) {
            const int top = n - 1;
            const int level = w.nodes[top].level + 1;
            if (level >= r_size) {
                --n;
                continue;
            }
            w.nodes[top].level = level;
            const double sumr = w.nodes[top].sumr;
            const int sumfactor = w.nodes[top].sumfactor;

            const double sumr1 = __dadd_rn(sumr, w.items[level].r);
            const int sf1 = sumfactor + w.items[level].factor;
            w.nodes[n] = ZNode{ sumr1, sf1, level };
            w.copy(top, n);
            w.set(n, level);
            ++n;

            if (inside(sf1)) {
                if (is_best(sumr1)) {
                    w.copy(n - 1, w.solution);
                    lower_bound = sumr1;
                }
                if (equality)
                    --n;
            } else if (!equality && sf1 > b_max)
                --n;
        }
    }

    for (int i = 0; i != r_size; ++i) {
        const int bit = w.test(w.solution, i) ? 1 : 0;
        w.items[i].result = w.items[i].original_factor > 0 ? bit : 1 - bit;
    }

    return z_finish(w, R, r_size, maximize);
}

// ---------------------------------------------------------------------------
// branch_and_bound_solver::solve by the whole CTA.
//
// The reference's search (branch-and-bound-solver.hpp:396-448) is a depth-first
// enumeration cut only by the factor sum, so what it returns is a closed form:
// let C be the subsets other than {} and {first item} whose factor sum is inside
// the bounds and whose reduced-cost sum z strictly beats the initial bound
// (+-infinity).  If C is not empty the answer is the member of C with the best
// z, and among equal z the one the search meets first: subsets are met in
// lexicographic order of their increasing index sequences, a prefix before its
// extensions.  If C is empty the answer is {first item} when that is inside,
// else {} (the two nodes the search starts from are stored without being
// compared).  An equality row whose adjusted bound is 0 returns {} at once.
// z is accumulated over the chosen items in increasing index, as the search
// accumulates it down a branch, so every sum has the reference's bits.
//
// Here every thread owns one pattern of the first q items and walks all
// patterns of the others; a block reduction picks the winner.
// ---------------------------------------------------------------------------
constexpr int kZParallelMax = 20; // rows up to this length (2^20 subsets); longer ones stay sequential

// sorted items, positive factors, adjusted bounds: z_bb_prepare
__device__ inline void
z_bb_prepare(const ZWork& w, const ZRow& R, int r_size, bool maximize, int& b_min, int& b_max)
{
    for (int i = 0; i != r_size; ++i) {
        ZItem it;
        it.r = R.value(i);
        it.sum_z = 0.0;
        it.variable = R.sid[i];
        it.original_factor = R.factor(i);
        it.factor = it.original_factor < 0 ? -it.original_factor : it.original_factor;
        it.result = 0;
        w.items[i] = it;
    }
    seq_std_sort(Plain<ZItem>{ w.items }, r_size, ZRatioLess{ maximize });
    for (int i = 0; i != r_size; ++i) {
        if (w.items[i].original_factor < 0) {
            b_min += w.items[i].factor;
            b_max += w.items[i].factor;
        }
    }
}

// A is met before B by the depth-first search (A != B)
__device__ __forceinline__ bool
z_met_first(unsigned A, unsigned B)
{
    const int j = __ffs((int)(A ^ B)) - 1;
    return ((A >> j) & 1u) ? ((B >> (j + 1)) != 0u) : ((A >> (j + 1)) == 0u);
}

// All threads.  red_z / red_m: shared scratch of 32 entries each.  Returns the
// chosen subset as a bit mask over the sorted items (uniform).
__device__ inline unsigned
z_bb_enumerate(const ZWork& w, int r_size, bool maximize, int b_min, int b_max, double* red_z,
               int* red_m)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int ln = tid & 31, wp = tid >> 5, nw = T >> 5;
    const bool equality = b_min == b_max;
    const double bad = maximize ? -CUDART_INF : CUDART_INF;
    auto inside = [&](int s) { return equality ? s == b_min : (b_min <= s && s <= b_max); };
    auto better = [&](double z, unsigned m, double bz, unsigned bm) {
        if (bm == 0xffffffffu)
            return true;
        if (z != bz)
            return maximize ? z > bz : z < bz;
        return z_met_first(m, bm);
    };

    double bz = bad;
    unsigned bm = 0xffffffffu; // none
    if (!(equality && b_min == 0)) {
        int q = 0;
        while ((2 << q) <= T && q + 1 <= r_size)
            ++q;
        const unsigned lo = (unsigned)tid;
        if (lo < (1u << q)) {
            int sf0 = 0;
            double z0 = 0.0;
            for (int i = 0; i < q; ++i)
                if ((lo >> i) & 1u) {
                    sf0 += w.items[i].factor;
                    z0 = __dadd_rn(z0, w.items[i].r);
                }
            // the other items in three groups, lowest indices outermost: a group's
            // partial sums are shared by everything nested below it, and an
            // inequality row drops a whole nest once the factor sum is past b_max
            const int nb = r_size - q;
            const int n1 = nb / 3, n2 = nb / 3, n3 = nb - n1 - n2;
            const int i1 = q, i2 = q + n1, i3 = q + n1 + n2;
            auto add_group = [&](unsigned pat, int first, int& sf, double& z) {
                for (unsigned rest = pat, i = (unsigned)first; rest; rest >>= 1, ++i)
                    if (rest & 1u) {
                        sf += w.items[i].factor;
                        z = __dadd_rn(z, w.items[i].r);
                    }
            };
            // (factors are positive here: a sum past b_max never comes back inside,
            // for an equality row as well)
            if (sf0 <= b_max) {
                for (unsigned p1 = 0; p1 < (1u << n1); ++p1) {
                    int sf1 = sf0;
                    double z1 = z0;
                    add_group(p1, i1, sf1, z1);
                    if (sf1 > b_max)
                        continue;
                    for (unsigned p2 = 0; p2 < (1u << n2); ++p2) {
                        int sf2 = sf1;
                        double z2 = z1;
                        add_group(p2, i2, sf2, z2);
                        if (sf2 > b_max)
                            continue;
                        for (unsigned p3 = 0; p3 < (1u << n3); ++p3) {
                            const unsigned m = lo | (p1 << i1) | (p2 << i2) | (p3 << i3);
                            if (m < 2u)
                                continue; // {} and {first item}: the fallbacks
                            int sf = sf2;
                            double z = z2;
                            add_group(p3, i3, sf, z);
                            if (inside(sf) && (maximize ? z > bad : z < bad) && better(z, m, bz, bm)) {
                                bz = z;
                                bm = m;
                            }
                        }
                    }
                }
            }
        }
    }
    // block reduction with the same order
    for (int o = 16; o; o >>= 1) {
        const double oz = __shfl_xor_sync(0xffffffffu, bz, o);
        const unsigned om = __shfl_xor_sync(0xffffffffu, bm, o);
        if (om != 0xffffffffu && better(oz, om, bz, bm)) {
            bz = oz;
            bm = om;
        }
    }
    if (ln == 0) {
        red_z[wp] = bz;
        red_m[wp] = (int)bm;
    }
    __syncthreads();
    bz = red_z[0];
    bm = (unsigned)red_m[0];
    for (int v = 1; v < nw; ++v) {
        const double oz = red_z[v];
        const unsigned om = (unsigned)red_m[v];
        if (om != 0xffffffffu && better(oz, om, bz, bm)) {
            bz = oz;
            bm = om;
        }
    }
    if (bm == 0xffffffffu) // nothing met in the loop: the two start nodes
        bm = (!(equality && b_min == 0) && inside(w.items[0].factor)) ? 1u : 0u;
    return bm;
}

// tail of branch_and_bound_solver::solve (:489-533) from the chosen subset
__device__ inline int
z_bb_finish(const ZWork& w, const ZRow& R, int r_size, bool maximize, unsigned chosen)
{
    for (int i = 0; i != r_size; ++i) {
        const int bit = (chosen >> i) & 1u;
        w.items[i].result = w.items[i].original_factor > 0 ? bit : 1 - bit;
    }
    return z_finish(w, R, r_size, maximize);
}

// exhaustive_solver::solve (exhaustive-solver.hpp:185-252) by the whole CTA:
// the pattern with the best sum, the first of them when sums are equal (the
// reference keeps the incumbent unless strictly beaten).  items[] set up by
// z_ex_prepare.  Returns the pattern's index, or -1 when a sum is NaN (the
// sequential statement decides then).
__device__ inline void
z_ex_prepare(const ZWork& w, const ZRow& R, int r_size)
{
    for (int i = 0; i != r_size; ++i) {
        ZItem it;
        it.r = R.value(i);
        it.sum_z = 0.0;
        it.variable = R.sid[i];
        it.factor = it.original_factor = R.factor(i);
        it.result = 0;
        w.items[i] = it;
    }
}

__device__ inline int
z_ex_enumerate(const ZWork& w, int r_size, bool maximize, const int* masks, int solutions, double* red_z,
               int* red_m)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int ln = tid & 31, wp = tid >> 5, nw = T >> 5;
    double bz = 0.0;
    int bi = -1, nan = 0;
    auto better = [&](double z, int i, double oz, int oi) {
        if (oi < 0)
            return true;
        if (z != oz)
            return maximize ? z > oz : z < oz;
        return i < oi;
    };
    for (int i = tid; i < solutions; i += T) {
        const int mk = masks[i];
        double z = 0.0;
        for (int j = 0; j != r_size; ++j)
            if ((mk >> j) & 1)
                z = __dadd_rn(z, w.items[j].r);
        nan |= z != z;
        if (better(z, i, bz, bi)) {
            bz = z;
            bi = i;
        }
    }
    for (int o = 16; o; o >>= 1) {
        const double oz = __shfl_xor_sync(0xffffffffu, bz, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
        if (oi >= 0 && better(oz, oi, bz, bi)) {
            bz = oz;
            bi = oi;
        }
    }
    if (ln == 0) {
        red_z[wp] = bz;
        red_m[wp] = nan ? -2 : bi;
    }
    __syncthreads();
    bz = red_z[0];
    bi = red_m[0];
    bool any_nan = bi == -2;
    for (int v = 1; v < nw; ++v) {
        const double oz = red_z[v];
        const int oi = red_m[v];
        if (oi == -2)
            any_nan = true;
        else if (oi >= 0 && (bi < 0 || better(oz, oi, bz, bi))) {
            bz = oz;
            bi = oi;
        }
    }
    return any_nan ? -1 : bi;
}

__device__ inline int
z_ex_finish(const ZWork& w, const ZRow& R, int r_size, bool maximize, int mask)
{
    for (int i = 0; i != r_size; ++i)
        w.items[i].result = (mask >> i) & 1;
    return z_finish(w, R, r_size, maximize);
}

} // namespace bxg



#endif
