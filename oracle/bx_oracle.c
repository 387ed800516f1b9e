/* bx_oracle.c -- CPU restatement of the reference's constraint-sweep path.
 *
 * TEST INFRASTRUCTURE ONLY (see bx_oracle.h).  Build with
 *   gcc -O2 -ffp-contract=off -std=c11 -fPIC -shared   (oracle/Makefile)
 * -ffp-contract=off matters: the reference oracle build uses no FMA either.
 *
 * Reference = quesnel/baryonyx v0.5.0 mounted at /root/reference; citations
 * are relative to it.  libstdc++ citations are /usr/include/c++/13/bits/.
 */
#include "bx_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/* Parameters                                                               */
/* ======================================================================== */

/* Defaults: lib/include/baryonyx/core:410-461 */
void
ox_default_params(ox_params* p)
{
    memset(p, 0, sizeof(*p));
    p->time_limit = -1;
    p->theta = 0.5;
    p->delta = -1;
    p->kappa_min = 0;
    p->kappa_step = 1.e-3;
    p->kappa_max = 0.6;
    p->alpha = 1;
    p->pushing_k_factor = 0.9;
    p->pushing_objective_amplifier = 5;
    p->init_policy_random = 0.5;
    p->init_crossover_bastert_insertion = 0.01;
    p->init_crossover_solution_selection_mean = 0.0;
    p->init_crossover_solution_selection_stddev = 0.3;
    p->init_mutation_variable_mean = 0.0001;
    p->init_mutation_variable_stddev = 0.001;
    p->init_mutation_value_mean = 0.5;
    p->init_mutation_value_stddev = 0.2;
    p->init_kappa_improve_start = 0;
    p->init_kappa_improve_increase = 0.02;
    p->init_kappa_improve_stop = 0.2;
    p->w = 0.05;
    p->seed = 0;
    p->thread = 0;
    p->limit = 1000;
    p->print_level = 0;
    p->pushes_limit = 100;
    p->pushing_iteration_limit = 50;
    p->init_population_size = 100;
    p->pre_order = 0;
    p->order = 0;
    p->float_type = 1;
    p->init_policy = 0;
    p->cost_norm = 4;
    p->mode = 0;
    p->preprocessor = 1;
    p->observer = 0;
    p->storage = 0;
    p->solver = 0;
    p->debug = 0;
}

static double
clampd(double v, double lo, double hi)
{
    /* std::clamp: (v < lo) ? lo : (hi < v) ? hi : v */
    return (v < lo) ? lo : (hi < v) ? hi : v;
}

/* lib/src/private.cpp:51-172.  The context starts from the defaults, then
 * every field is copied only if it passes the reference's test. */
void
ox_clamp_params(const ox_params* in, ox_params* out)
{
    ox_params c;
    ox_default_params(&c);

    c.time_limit = in->time_limit <= 0 ? INFINITY : in->time_limit;

    if (in->theta >= 0 && in->theta <= 1)
        c.theta = in->theta;

    if (in->delta == -1 || in->delta >= 0)
        c.delta = in->delta;

    if (in->kappa_min < in->kappa_max && in->kappa_min >= 0 &&
        in->kappa_min < 1 && in->kappa_max <= 1 && in->kappa_step >= 0 &&
        in->kappa_step < 1) {
        c.kappa_min = in->kappa_min;
        c.kappa_step = in->kappa_step;
        c.kappa_max = in->kappa_max;
    }

    if (in->alpha >= 0)
        c.alpha = in->alpha;

    if (isfinite(in->pushing_k_factor))
        c.pushing_k_factor = in->pushing_k_factor;

    if (isfinite(in->pushing_objective_amplifier))
        c.pushing_objective_amplifier = in->pushing_objective_amplifier;

    if (isfinite(in->init_crossover_bastert_insertion))
        c.init_crossover_bastert_insertion =
          clampd(in->init_crossover_bastert_insertion, 0.0, 1.0);

    if (isfinite(in->init_crossover_solution_selection_mean))
        c.init_crossover_solution_selection_mean =
          clampd(in->init_crossover_solution_selection_mean, 0.0, 1.0);

    if (isfinite(in->init_crossover_solution_selection_stddev))
        c.init_crossover_solution_selection_stddev =
          clampd(in->init_crossover_solution_selection_stddev, 0.0, 1.0);

    if (isfinite(in->init_mutation_variable_mean))
        c.init_mutation_variable_mean =
          clampd(in->init_mutation_variable_mean, 0.0, 1.0);

    if (isfinite(in->init_mutation_variable_stddev))
        c.init_mutation_variable_stddev =
          clampd(in->init_mutation_variable_stddev, 0.0, 1.0);

    if (isfinite(in->init_mutation_value_mean))
        c.init_mutation_value_mean =
          clampd(in->init_mutation_value_mean, 0.0, 1.0);

    if (isfinite(in->init_mutation_value_stddev))
        c.init_mutation_value_stddev =
          clampd(in->init_mutation_value_stddev, 0.0, 1.0);

    if (in->init_policy_random >= 0 && in->init_policy_random <= 1)
        c.init_policy_random = in->init_policy_random;

    c.seed = in->seed;
    c.thread = in->thread;

    c.limit = (in->limit <= 0) ? INT64_MAX : in->limit;

    if (in->print_level >= 0)
        c.print_level = in->print_level;

    if (in->w < 0) {
        c.w = 0;
    } else {
        if (in->limit > 0) {
            double limit = (double)in->limit;

            if (in->w < 1)
                c.w = limit * in->w;
            else
                c.w = limit < in->w ? limit : in->w; /* std::min */
        } else {
            c.w = floor(in->w);
        }
    }

    c.pushes_limit = in->pushes_limit <= 0 ? 0 : in->pushes_limit;
    c.pushing_iteration_limit =
      in->pushing_iteration_limit <= 0 ? 0 : in->pushing_iteration_limit;

    if (in->init_population_size >= 5)
        c.init_population_size = in->init_population_size;

    c.init_kappa_improve_start =
      clampd(in->init_kappa_improve_start, 0.0, 1.0);
    c.init_kappa_improve_increase =
      clampd(in->init_kappa_improve_increase, 0.0, 1.0);
    c.init_kappa_improve_stop =
      clampd(in->init_kappa_improve_stop, c.init_kappa_improve_start, 1.0);

    c.pre_order = in->pre_order;
    c.order = in->order;
    c.float_type = in->float_type;
    c.init_policy = in->init_policy;
    c.cost_norm = in->cost_norm;
    c.mode = in->mode;
    c.preprocessor = in->preprocessor;
    c.observer = in->observer;
    c.storage = in->storage;
    c.solver = in->solver;
    c.debug = in->debug;

    *out = c;
}

/* lib/src/manual-optimizer.cpp:31-111 (manual_course: the value tables, filled
 * in the reference's statement order; `next()`: the odometer) and :113-160
 * (optimize(): which table entry each parameter of a pass takes). */
int
ox_manual_grid(const ox_params* p, double* theta, double* delta, double* kappa_min,
               double* kappa_step, double* init_policy_random, double* tables)
{
    enum { LENGTH = 5 };
    double t_theta[LENGTH], t_delta[LENGTH], t_kmin[LENGTH], t_kstep[LENGTH], t_ipr[LENGTH];
    int it[LENGTH];
    int i, count = 0;
    const double length = (double)LENGTH;

    t_theta[0] = p->theta; /* :60-61 */
    t_theta[1] = 1.0 / length;
    t_delta[0] = p->delta; /* :63-64 */
    t_delta[1] = 0.1 / length;
    t_kmin[0] = p->kappa_min; /* :66-67 */
    t_kmin[1] = 1e-2 / length;
    t_kstep[0] = p->kappa_step; /* :69-70 */
    t_kstep[1] = 1e-7 / length;
    t_ipr[0] = p->init_policy_random; /* :72-73 */
    t_ipr[1] = 0.9 / length;
    for (i = 2; i != LENGTH; ++i) /* :75-76 */
        t_theta[i] = t_theta[i - 1] + 1.0 / length;
    for (i = 2; i != LENGTH; ++i) /* :78-80: built on theta's table */
        t_delta[i] = t_theta[i - 1] + 0.1 / (5.0 * length);
    for (i = 2; i != LENGTH; ++i) /* :82-83 */
        t_kmin[i] = t_theta[i - 1] + 1e-2 / length;
    for (i = 2; i != LENGTH; ++i) /* :85-87 */
        t_kstep[i] = t_theta[i - 1] + 1e-7 / (5.0 * length);
    for (i = 2; i != LENGTH; ++i) /* :89-91 */
        t_ipr[i] = t_theta[i - 1] + 0.9 / length;
    if (tables)
        for (i = 0; i != LENGTH; ++i) {
            tables[0 * LENGTH + i] = t_theta[i];
            tables[1 * LENGTH + i] = t_delta[i];
            tables[2 * LENGTH + i] = t_kmin[i];
            tables[3 * LENGTH + i] = t_kstep[i];
            tables[4 * LENGTH + i] = t_ipr[i];
        }

    for (i = 0; i != LENGTH; ++i) /* reset(), :96 */
        it[i] = 0;
    for (;;) {
        int moved = 0;
        theta[count] = t_theta[it[0]]; /* :130-137 */
        delta[count] = t_delta[it[1]];
        kappa_min[count] = t_kmin[it[2]];
        kappa_step[count] = t_kstep[it[3]];
        init_policy_random[count] = t_ipr[it[4]];
        ++count;
        for (i = LENGTH - 1; i >= 0; --i) { /* next(), :98-111 */
            if (it[i] + 1 >= LENGTH)
                it[i] = 0;
            else {
                ++it[i];
                moved = 1;
                break;
            }
        }
        if (!moved)
            break;
    }
    return count;
}

/* ======================================================================== */
/* libstdc++ 13: minstd_rand0 and the distributions                         */
/* ======================================================================== */

#define OX_RNG_M 2147483647u /* 2^31 - 1 */
#define OX_RNG_A 16807u
#define OX_RNG_MIN 1u
#define OX_RNG_MAX 2147483646u

/* std::default_random_engine = minstd_rand0, random.h:1592-1641 */
uint32_t
ox_minstd_next(uint32_t* state)
{
    *state = (uint32_t)(((uint64_t)*state * OX_RNG_A) % OX_RNG_M);
    return *state;
}

/* generate_canonical<double,53>, random.tcc:3349-3381: r = 2147483646,
 * log2r = 30, m = 2 calls; sum = (g1-min)*1 + (g2-min)*r; ret = sum/(r*r). */
static double
ox_canonical(uint32_t* state)
{
    const double r = 2147483646.0;
    double sum = 0.0;
    double tmp = 1.0;

    sum += (double)(ox_minstd_next(state) - OX_RNG_MIN) * tmp;
    tmp *= r;
    sum += (double)(ox_minstd_next(state) - OX_RNG_MIN) * tmp;
    tmp *= r;

    double ret = sum / tmp;
    if (ret >= 1.0)
        ret = nextafter(1.0, 0.0);
    return ret;
}

/* bernoulli_distribution::operator(), random.h:3739-3750 */
int
ox_bernoulli(double p, uint32_t* state)
{
    return (ox_canonical(state) - 0.0) < p * (1.0 - 0.0);
}

/* uniform_int_distribution<unsigned long>::operator(),
 * uniform_int_dist.h:284-373, for a generator with range 2147483645. */
uint64_t
ox_uniform_u64(uint64_t a, uint64_t b, uint32_t* state)
{
    const uint64_t urngrange = (uint64_t)OX_RNG_MAX - OX_RNG_MIN;
    const uint64_t urange = b - a;
    uint64_t ret;

    if (urngrange > urange) {
        const uint64_t uerange = urange + 1;
        const uint64_t scaling = urngrange / uerange;
        const uint64_t past = uerange * scaling;
        do
            ret = (uint64_t)ox_minstd_next(state) - OX_RNG_MIN;
        while (ret >= past);
        ret /= scaling;
    } else if (urngrange < urange) {
        uint64_t tmp;
        do {
            const uint64_t uerngrange = urngrange + 1;
            tmp = uerngrange * ox_uniform_u64(0, urange / uerngrange, state);
            ret = tmp + ((uint64_t)ox_minstd_next(state) - OX_RNG_MIN);
        } while (ret > urange || ret < tmp);
    } else {
        ret = (uint64_t)ox_minstd_next(state) - OX_RNG_MIN;
    }

    return ret + a;
}

/* normal_distribution<double>::operator() (Marsaglia polar, random.tcc),
 * stateful: the second value of a pair is cached. */
typedef struct ox_normal_dist
{
    double mean, stddev, saved;
    int saved_available;
} ox_normal_dist;

static double
ox_normal_draw(ox_normal_dist* d, uint32_t* state)
{
    double ret;

    if (d->saved_available) {
        d->saved_available = 0;
        ret = d->saved;
    } else {
        double x, y, r2;
        do {
            x = 2.0 * ox_canonical(state) - 1.0;
            y = 2.0 * ox_canonical(state) - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);

        const double mult = sqrt(-2 * log(r2) / r2);
        d->saved = x * mult;
        d->saved_available = 1;
        ret = y * mult;
    }

    return ret * d->stddev + d->mean;
}

void
ox_normal(double mean, double stddev, int count, double* out, uint32_t* state)
{
    ox_normal_dist d = { mean, stddev, 0.0, 0 };
    for (int i = 0; i != count; ++i)
        out[i] = ox_normal_draw(&d, state);
}

/* ======================================================================== */
/* libstdc++ 13: std::sort and std::shuffle on generic elements             */
/* ======================================================================== */

typedef int (*ox_less_fn)(const void* a, const void* b, void* ctx);

#define OX_MAX_ELEM 64

typedef struct ox_seq
{
    char* base;
    size_t size;
    ox_less_fn less;
    void* ctx;
} ox_seq;

#define AT(s, i) ((s)->base + (size_t)(i) * (s)->size)

static void
seq_swap(ox_seq* s, long i, long j)
{
    char tmp[OX_MAX_ELEM];
    if (i == j)
        return;
    memcpy(tmp, AT(s, i), s->size);
    memcpy(AT(s, i), AT(s, j), s->size);
    memcpy(AT(s, j), tmp, s->size);
}

/* __move_median_to_first, stl_algo.h:82-104 */
static void
move_median_to_first(ox_seq* s, long result, long a, long b, long c)
{
    if (s->less(AT(s, a), AT(s, b), s->ctx)) {
        if (s->less(AT(s, b), AT(s, c), s->ctx))
            seq_swap(s, result, b);
        else if (s->less(AT(s, a), AT(s, c), s->ctx))
            seq_swap(s, result, c);
        else
            seq_swap(s, result, a);
    } else if (s->less(AT(s, a), AT(s, c), s->ctx))
        seq_swap(s, result, a);
    else if (s->less(AT(s, b), AT(s, c), s->ctx))
        seq_swap(s, result, c);
    else
        seq_swap(s, result, b);
}

/* __unguarded_partition, stl_algo.h:1873-1892 */
static long
unguarded_partition(ox_seq* s, long first, long last, long pivot)
{
    for (;;) {
        while (s->less(AT(s, first), AT(s, pivot), s->ctx))
            ++first;
        --last;
        while (s->less(AT(s, pivot), AT(s, last), s->ctx))
            --last;
        if (!(first < last))
            return first;
        seq_swap(s, first, last);
        ++first;
    }
}

/* __push_heap, stl_heap.h:135-148 (value compared as iter_comp_val) */
static void
push_heap_(ox_seq* s, long first, long hole, long top, const char* value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && s->less(AT(s, first + parent), value, s->ctx)) {
        memcpy(AT(s, first + hole), AT(s, first + parent), s->size);
        hole = parent;
        parent = (hole - 1) / 2;
    }
    memcpy(AT(s, first + hole), value, s->size);
}

/* __adjust_heap, stl_heap.h:224-251 */
static void
adjust_heap(ox_seq* s, long first, long hole, long len, const char* value)
{
    const long top = hole;
    long second = hole;

    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (s->less(AT(s, first + second), AT(s, first + (second - 1)), s->ctx))
            second--;
        memcpy(AT(s, first + hole), AT(s, first + second), s->size);
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        memcpy(AT(s, first + hole), AT(s, first + (second - 1)), s->size);
        hole = second - 1;
    }
    push_heap_(s, first, hole, top, value);
}

/* __partial_sort(first, last, last) = __heap_select (make_heap only, the
 * [middle,last) loop is empty) + __sort_heap; stl_algo.h:1906-1917,
 * stl_heap.h:340-361, 419-428, 254-268 */
static void
heap_sort(ox_seq* s, long first, long last)
{
    char value[OX_MAX_ELEM];
    const long len = last - first;

    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            memcpy(value, AT(s, first + parent), s->size);
            adjust_heap(s, first, parent, len, value);
            if (parent == 0)
                break;
            parent--;
        }
    }

    while (last - first > 1) {
        --last;
        /* __pop_heap(first, last, last) */
        memcpy(value, AT(s, last), s->size);
        memcpy(AT(s, last), AT(s, first), s->size);
        adjust_heap(s, first, 0, last - first, value);
    }
}

/* __introsort_loop, stl_algo.h:1920-1941 */
static void
introsort_loop(ox_seq* s, long first, long last, long depth_limit)
{
    while (last - first > 16) {
        if (depth_limit == 0) {
            heap_sort(s, first, last);
            return;
        }
        --depth_limit;
        /* __unguarded_partition_pivot, stl_algo.h:1895-1904 */
        long mid = first + (last - first) / 2;
        move_median_to_first(s, first, first + 1, mid, last - 1);
        long cut = unguarded_partition(s, first + 1, last, first);
        introsort_loop(s, cut, last, depth_limit);
        last = cut;
    }
}

/* __unguarded_linear_insert, stl_algo.h:1792-1808 */
static void
unguarded_linear_insert(ox_seq* s, long last)
{
    char val[OX_MAX_ELEM];
    memcpy(val, AT(s, last), s->size);
    long next = last - 1;
    while (s->less(val, AT(s, next), s->ctx)) {
        memcpy(AT(s, last), AT(s, next), s->size);
        last = next;
        --next;
    }
    memcpy(AT(s, last), val, s->size);
}

/* __insertion_sort, stl_algo.h:1811-1832 */
static void
insertion_sort(ox_seq* s, long first, long last)
{
    char val[OX_MAX_ELEM];
    if (first == last)
        return;

    for (long i = first + 1; i != last; ++i) {
        if (s->less(AT(s, i), AT(s, first), s->ctx)) {
            memcpy(val, AT(s, i), s->size);
            memmove(AT(s, first + 1), AT(s, first), (size_t)(i - first) * s->size);
            memcpy(AT(s, first), val, s->size);
        } else
            unguarded_linear_insert(s, i);
    }
}

/* std::sort = __sort, stl_algo.h:1945-1957 with __final_insertion_sort
 * :1857-1870 */
static void
ox_std_sort(void* base, long n, size_t size, ox_less_fn less, void* ctx)
{
    ox_seq s = { (char*)base, size, less, ctx };

    if (n == 0)
        return;

    long lg = 0;
    for (long t = n; t > 1; t >>= 1)
        ++lg;

    introsort_loop(&s, 0, n, lg * 2);

    if (n > 16) {
        insertion_sort(&s, 0, 16);
        for (long i = 16; i != n; ++i)
            unguarded_linear_insert(&s, i);
    } else
        insertion_sort(&s, 0, n);
}

/* std::shuffle, stl_algo.h:3738-3803 */
static void
ox_std_shuffle(void* base, long n, size_t size, uint32_t* state)
{
    ox_seq s = { (char*)base, size, NULL, NULL };

    if (n == 0)
        return;

    const uint64_t urngrange = (uint64_t)OX_RNG_MAX - OX_RNG_MIN;
    const uint64_t urange = (uint64_t)n;

    if (urngrange / urange >= urange) {
        long i = 1;

        if ((urange % 2) == 0) {
            seq_swap(&s, i, (long)ox_uniform_u64(0, 1, state));
            ++i;
        }

        while (i != n) {
            const uint64_t swap_range = (uint64_t)i + 1;
            /* __gen_two_uniform_ints(swap_range, swap_range + 1) */
            const uint64_t b1 = swap_range + 1;
            const uint64_t x =
              ox_uniform_u64(0, (swap_range * b1) - 1, state);
            seq_swap(&s, i, (long)(x / b1));
            ++i;
            seq_swap(&s, i, (long)(x % b1));
            ++i;
        }
        return;
    }

    for (long i = 1; i != n; ++i)
        seq_swap(&s, i, (long)ox_uniform_u64(0, (uint64_t)i, state));
}

void
ox_shuffle_int(int* a, int n, uint32_t* state)
{
    ox_std_shuffle(a, n, sizeof(int), state);
}

/* ---- reduced-cost element (union of the three solvers' rc_data) -------- */
typedef struct ox_rc
{
    double value;
    int id;
    int f;
} ox_rc;

static int
rc_less_min(const void* a, const void* b, void* ctx)
{
    (void)ctx;
    return ((const ox_rc*)a)->value < ((const ox_rc*)b)->value;
}

static int
rc_less_max(const void* a, const void* b, void* ctx)
{
    (void)ctx;
    return ((const ox_rc*)b)->value < ((const ox_rc*)a)->value;
}

void
ox_sort_rc(double* value, int* id, int n, int descending)
{
    ox_rc* v = (ox_rc*)malloc(sizeof(ox_rc) * (size_t)(n ? n : 1));
    for (int i = 0; i != n; ++i) {
        v[i].value = value[i];
        v[i].id = id[i];
        v[i].f = 0;
    }
    ox_std_sort(v, n, sizeof(ox_rc), descending ? rc_less_max : rc_less_min,
                NULL);
    for (int i = 0; i != n; ++i) {
        value[i] = v[i].value;
        id[i] = v[i].id;
    }
    free(v);
}

typedef struct ox_pair
{
    int first;
    int second;
} ox_pair;

static int
pair_less_incr(const void* a, const void* b, void* ctx)
{
    (void)ctx;
    return ((const ox_pair*)a)->second < ((const ox_pair*)b)->second;
}

static int
pair_less_decr(const void* a, const void* b, void* ctx)
{
    (void)ctx;
    return ((const ox_pair*)b)->second < ((const ox_pair*)a)->second;
}

void
ox_sort_pairs(int* first, int* second, int n, int descending)
{
    ox_pair* v = (ox_pair*)malloc(sizeof(ox_pair) * (size_t)(n ? n : 1));
    for (int i = 0; i != n; ++i) {
        v[i].first = first[i];
        v[i].second = second[i];
    }
    ox_std_sort(v, n, sizeof(ox_pair),
                descending ? pair_less_decr : pair_less_incr, NULL);
    for (int i = 0; i != n; ++i) {
        first[i] = v[i].first;
        second[i] = v[i].second;
    }
    free(v);
}

/* ======================================================================== */
/* Z sub-solvers                                                            */
/* ======================================================================== */

/* ---- exhaustive_solver, lib/src/exhaustive-solver.hpp:40-253 ----------- */

typedef struct ox_ex_item
{
    double r;
    int factor;
    int variable;
    int result;
} ox_ex_item;

static int
ex_item_less_min(const void* a, const void* b, void* ctx)
{
    const ox_ex_item* l = (const ox_ex_item*)a;
    const ox_ex_item* r = (const ox_ex_item*)b;
    (void)ctx;
    if (l->result == r->result)
        return l->r < r->r;
    return l->result > r->result;
}

static int
ex_item_less_max(const void* a, const void* b, void* ctx)
{
    const ox_ex_item* l = (const ox_ex_item*)a;
    const ox_ex_item* r = (const ox_ex_item*)b;
    (void)ctx;
    if (l->result == r->result)
        return l->r > r->r;
    return l->result > r->result;
}

/* build_constraints, exhaustive-solver.hpp:111-167.  Returns the number of
 * feasible patterns; flat receives nb*size ints (factor or 0). */
static int
ex_build(const int* factor, int size, int bk_min, int bk_max, int** flat_out)
{
    int cap = 16 * size, len = 0, nb = 0;
    int* flat = (int*)malloc(sizeof(int) * (size_t)cap);
    signed char* walkers = (signed char*)calloc((size_t)size, 1);
    int end = 0;

    walkers[size - 1] = 1;

    do {
        int i = size - 1;

        do {
            int sum = 0;
            for (int j = 0; j != size; ++j)
                if (walkers[j])
                    sum += factor[j];

            if (bk_min <= sum && sum <= bk_max) {
                ++nb;
                if (len + size > cap) {
                    cap *= 2;
                    flat = (int*)realloc(flat, sizeof(int) * (size_t)cap);
                }
                for (int j = 0; j != size; ++j)
                    flat[len++] = walkers[j] ? factor[j] : 0;
            }

            ++walkers[i];

            if (walkers[i] > 1) {
                walkers[i] = 0;

                if (i == 0) {
                    end = 1;
                    break;
                } else {
                    --i;
                }
            } else {
                break;
            }
        } while (!end);
    } while (!end);

    free(walkers);
    *flat_out = flat;
    return nb;
}

/* solve, exhaustive-solver.hpp:185-252 */
static int
ex_solve(int maximize, const int* flat, int solutions, ox_rc* R, int r_size)
{
    ox_ex_item* items =
      (ox_ex_item*)malloc(sizeof(ox_ex_item) * (size_t)r_size);

    for (int i = 0; i != r_size; ++i) {
        items[i].r = R[i].value;
        items[i].variable = R[i].id;
        items[i].factor = R[i].f;
        items[i].result = 0;
    }

    double z_best = 0;
    int best = 0;

    for (int j = 0; j != r_size; ++j)
        if (flat[j])
            z_best += R[j].value;

    for (int i = 1; i != solutions; ++i) {
        const int start = i * r_size;
        double z = 0;
        for (int j = 0; j != r_size; ++j)
            if (flat[start + j])
                z += R[j].value;

        if (maximize ? (z > z_best) : (z < z_best)) {
            z_best = z;
            best = i;
        }
    }

    for (int i = 0; i != r_size; ++i)
        items[i].result = flat[best * r_size + i] ? 1 : 0;

    ox_std_sort(items, r_size, sizeof(ox_ex_item),
                maximize ? ex_item_less_max : ex_item_less_min, NULL);

    int middle = 0;
    while (middle != r_size && items[middle].result != 0)
        ++middle;

    for (int i = 0; i != r_size; ++i) {
        R[i].value = items[i].r;
        R[i].id = items[i].variable;
        R[i].f = items[i].factor;
    }

    int ret;
    if (middle == r_size)
        ret = items[0].result == 0 ? -1 : r_size;
    else
        ret = middle - 1;

    free(items);
    return ret;
}

int
ox_exhaustive_solve(int maximize, const int* cst_factor, double* value,
                    int* factor, int* id, int n, int bmin, int bmax)
{
    int* flat = NULL;
    int nb = ex_build(cst_factor, n, bmin, bmax, &flat);
    ox_rc* R = (ox_rc*)malloc(sizeof(ox_rc) * (size_t)n);
    for (int i = 0; i != n; ++i) {
        R[i].value = value[i];
        R[i].id = id[i];
        R[i].f = factor[i];
    }
    int ret = ex_solve(maximize, flat, nb, R, n);
    for (int i = 0; i != n; ++i) {
        value[i] = R[i].value;
        id[i] = R[i].id;
        factor[i] = R[i].f;
    }
    free(R);
    free(flat);
    return ret;
}

/* ---- branch_and_bound_solver, lib/src/branch-and-bound-solver.hpp:192-534 */

typedef struct ox_bb_item
{
    double r;
    double sum_z;
    int factor;
    int original_factor;
    int variable;
    int result;
} ox_bb_item;

typedef struct ox_bb_node
{
    unsigned variables;
    double z;
    double sumr;
    int sumfactor;
    unsigned level;
} ox_bb_node;

typedef struct ox_bb
{
    int maximize;
    /* shared_subvector, branch-and-bound-solver.hpp:48-190 */
    signed char* pool;
    unsigned pool_size, pool_cap, elem, free_head;
    ox_bb_item* items;
    int n_items;
    ox_bb_node* nodes;
    int n_nodes, cap_nodes;
    double upper_bound, lower_bound;
    unsigned solution;
    int b_min, b_max;
} ox_bb;

#define OX_NPOS 0xffffffffu

static unsigned
sv_emplace(ox_bb* b)
{
    unsigned pos;
    if (b->free_head != OX_NPOS) {
        pos = b->free_head;
        memcpy(&b->free_head, &b->pool[pos], sizeof(unsigned));
    } else {
        pos = b->pool_size;
        if (b->pool_size + b->elem > b->pool_cap) {
            while (b->pool_size + b->elem > b->pool_cap)
                b->pool_cap *= 2;
            b->pool = (signed char*)realloc(b->pool, b->pool_cap);
        }
        memset(&b->pool[pos], 0, b->elem);
        b->pool_size += b->elem;
    }
    return pos;
}

static void
sv_remove(ox_bb* b, unsigned i)
{
    memcpy(&b->pool[i], &b->free_head, sizeof(unsigned));
    b->free_head = i;
}

static void
sv_copy(ox_bb* b, unsigned src, unsigned dst)
{
    memmove(&b->pool[dst], &b->pool[src], b->elem);
}

static void
sv_invert(ox_bb* b, unsigned i, unsigned element)
{
    b->pool[i + element] = b->pool[i + element] ? 0 : 1;
}

static ox_bb_node*
bb_push_node(ox_bb* b, double sumr, int sumfactor, unsigned level)
{
    if (b->n_nodes == b->cap_nodes) {
        b->cap_nodes *= 2;
        b->nodes = (ox_bb_node*)realloc(
          b->nodes, sizeof(ox_bb_node) * (size_t)b->cap_nodes);
    }
    ox_bb_node* nd = &b->nodes[b->n_nodes++];
    nd->variables = 0;
    nd->z = b->maximize ? -INFINITY : INFINITY; /* init_z() */
    nd->sumr = sumr;
    nd->sumfactor = sumfactor;
    nd->level = level;
    return nd;
}

static int
bb_is_best(const ox_bb* b, double node_z)
{
    return b->maximize ? node_z > b->lower_bound : node_z < b->lower_bound;
}

/* operator<, branch-and-bound-solver.hpp:212-220: both sides are divided by
 * the LEFT operand's factor (reference quirk, kept). */
static int
bb_item_less(const void* a, const void* b, void* ctx)
{
    const ox_bb_item* l = (const ox_bb_item*)a;
    const ox_bb_item* r = (const ox_bb_item*)b;
    const int maximize = *(const int*)ctx;
    if (!maximize)
        return (l->r / (double)l->factor) < (r->r / (double)l->factor);
    return (r->r / (double)l->factor) < (l->r / (double)l->factor);
}

static int
bb_result_less_min(const void* a, const void* b, void* ctx)
{
    const ox_bb_item* l = (const ox_bb_item*)a;
    const ox_bb_item* r = (const ox_bb_item*)b;
    (void)ctx;
    if (l->result == r->result)
        return l->r < r->r;
    return l->result > r->result;
}

static int
bb_result_less_max(const void* a, const void* b, void* ctx)
{
    const ox_bb_item* l = (const ox_bb_item*)a;
    const ox_bb_item* r = (const ox_bb_item*)b;
    (void)ctx;
    if (l->result == r->result)
        return l->r > r->r;
    return l->result > r->result;
}

/* make_init_node_0 / _1, :296-313 */
static void
bb_init_node_0(ox_bb* b)
{
    ox_bb_node* nd = bb_push_node(b, 0.0, 0, 0u);
    nd->variables = sv_emplace(b);
    nd = &b->nodes[b->n_nodes - 1];
    nd->z = b->items[1].sum_z;
    memset(&b->pool[nd->variables], 0, b->elem);
}

static void
bb_init_node_1(ox_bb* b)
{
    ox_bb_node* nd = bb_push_node(b, b->items[0].r, b->items[0].factor, 0u);
    nd->variables = sv_emplace(b);
    nd = &b->nodes[b->n_nodes - 1];
    nd->z = b->items[0].sum_z;
    memset(&b->pool[nd->variables], 0, b->elem);
    b->pool[nd->variables] = 1;
}

/* make_next_equality_node / make_next_inequality_node, :315-394 */
static void
bb_next_node(ox_bb* b, int equality)
{
    ox_bb_node* node_0 = &b->nodes[b->n_nodes - 1];

    const unsigned level = node_0->level + 1;
    if (level >= b->elem) {
        sv_remove(b, node_0->variables);
        b->n_nodes--;
        return;
    }

    node_0->level = level;
    node_0->z = node_0->sumr;

    if (level + 1 < b->elem)
        node_0->z += b->items[level + 1].sum_z;

    const double sumr = node_0->sumr;
    const int sumfactor = node_0->sumfactor;
    const unsigned variables = node_0->variables;

    ox_bb_node* node_1 = bb_push_node(b, sumr + b->items[level].r,
                                      sumfactor + b->items[level].factor,
                                      level);
    unsigned v1 = sv_emplace(b);
    node_1 = &b->nodes[b->n_nodes - 1];
    node_1->variables = v1;
    sv_copy(b, variables, node_1->variables);
    b->pool[node_1->variables + level] = 1;
    if (level + 1 < b->elem)
        node_1->z = node_1->sumr + b->items[level + 1].sum_z;

    if (equality) {
        if (node_1->sumfactor == b->b_min) {
            if (bb_is_best(b, node_1->sumr)) {
                sv_copy(b, variables, b->solution);
                sv_invert(b, b->solution, level);
                b->lower_bound = b->nodes[b->n_nodes - 1].sumr;
            }
            sv_remove(b, node_1->variables);
            b->n_nodes--;
        }
    } else {
        if (b->b_min <= node_1->sumfactor && node_1->sumfactor <= b->b_max) {
            if (bb_is_best(b, node_1->sumr)) {
                sv_copy(b, variables, b->solution);
                sv_invert(b, b->solution, level);
                b->lower_bound = b->nodes[b->n_nodes - 1].sumr;
            }
        } else if (node_1->sumfactor > b->b_max) {
            sv_remove(b, node_1->variables);
            b->n_nodes--;
        }
    }
}

static double
bb_upper_bound(const ox_bb* b)
{
    double ret = 0;
    for (int i = 0; i != b->n_items; ++i) {
        if (!b->maximize) {
            if (b->items[i].r < 0)
                ret += b->items[i].r;
        } else {
            if (b->items[i].r > 0)
                ret += b->items[i].r;
        }
    }
    return ret;
}

/* solve_equality / solve_inequality, :396-448 */
static void
bb_solve_equality(ox_bb* b)
{
    b->lower_bound = b->maximize ? -INFINITY : INFINITY;
    b->upper_bound = bb_upper_bound(b);

    bb_init_node_0(b);
    if (b->nodes[b->n_nodes - 1].sumfactor == b->b_min) {
        b->upper_bound = b->nodes[b->n_nodes - 1].sumr;
        sv_copy(b, b->nodes[b->n_nodes - 1].variables, b->solution);
        sv_remove(b, b->nodes[b->n_nodes - 1].variables);
        return;
    }

    bb_init_node_1(b);
    if (b->nodes[b->n_nodes - 1].sumfactor == b->b_min) {
        b->upper_bound = b->nodes[b->n_nodes - 1].sumr;
        sv_copy(b, b->nodes[b->n_nodes - 1].variables, b->solution);
        sv_remove(b, b->nodes[b->n_nodes - 1].variables);
        b->n_nodes--;
    }

    while (b->n_nodes)
        bb_next_node(b, 1);
}

static void
bb_solve_inequality(ox_bb* b)
{
    b->lower_bound = b->maximize ? -INFINITY : INFINITY;
    b->upper_bound = bb_upper_bound(b);

    bb_init_node_0(b);
    if (b->b_min <= b->nodes[b->n_nodes - 1].sumfactor &&
        b->nodes[b->n_nodes - 1].sumfactor <= b->b_max) {
        b->upper_bound = b->nodes[b->n_nodes - 1].sumr;
        sv_copy(b, b->nodes[b->n_nodes - 1].variables, b->solution);
    } else if (b->nodes[b->n_nodes - 1].sumfactor > b->b_max) {
        sv_remove(b, b->nodes[b->n_nodes - 1].variables);
        b->n_nodes--;
    }

    bb_init_node_1(b);
    if (b->b_min <= b->nodes[b->n_nodes - 1].sumfactor &&
        b->nodes[b->n_nodes - 1].sumfactor <= b->b_max) {
        b->upper_bound = b->nodes[b->n_nodes - 1].sumr;
        sv_copy(b, b->nodes[b->n_nodes - 1].variables, b->solution);
    } else if (b->nodes[b->n_nodes - 1].sumfactor > b->b_max) {
        sv_remove(b, b->nodes[b->n_nodes - 1].variables);
        b->n_nodes--;
    }

    while (b->n_nodes)
        bb_next_node(b, 0);
}

/* solve, :450-533.  Returns -3 where the reference throws
 * solver_failure(no_solver_available) (r_size < sizeof(uint32)). */
static int
bb_solve(int maximize, ox_rc* R, int r_size, int bk_min, int bk_max)
{
    ox_bb b;
    memset(&b, 0, sizeof(b));
    b.maximize = maximize;
    b.b_min = bk_min;
    b.b_max = bk_max;
    b.n_items = r_size;
    b.items = (ox_bb_item*)calloc((size_t)r_size, sizeof(ox_bb_item));
    b.elem = (unsigned)r_size;
    b.free_head = OX_NPOS;
    b.pool_cap = (unsigned)r_size * 1024u;
    b.pool = (signed char*)malloc(b.pool_cap);
    b.pool_size = 0;
    b.cap_nodes = 64;
    b.nodes = (ox_bb_node*)malloc(sizeof(ox_bb_node) * (size_t)b.cap_nodes);

    if (sizeof(unsigned) > (size_t)r_size) {
        free(b.items);
        free(b.pool);
        free(b.nodes);
        return -3;
    }

    b.solution = sv_emplace(&b);

    for (int i = 0; i != r_size; ++i) {
        b.items[i].r = R[i].value;
        b.items[i].variable = R[i].id;
        b.items[i].original_factor = R[i].f;
        b.items[i].factor = abs(R[i].f);
        b.items[i].result = 0;
    }

    memset(&b.pool[b.solution], 0, b.elem);

    ox_std_sort(b.items, r_size, sizeof(ox_bb_item), bb_item_less,
                &b.maximize);

    for (int i = 0; i != r_size; ++i) {
        if (b.items[i].original_factor < 0) {
            b.b_min += b.items[i].factor;
            b.b_max += b.items[i].factor;
        }
    }

    b.items[r_size - 1].sum_z = b.items[r_size - 1].r;
    for (int i = r_size - 2; i >= 0; --i)
        b.items[i].sum_z = b.items[i + 1].sum_z + b.items[i].r;

    if (b.b_max == b.b_min)
        bb_solve_equality(&b);
    else
        bb_solve_inequality(&b);

    for (int i = 0; i != r_size; ++i) {
        if (b.items[i].original_factor > 0)
            b.items[i].result = b.pool[b.solution + (unsigned)i] == 0 ? 0 : 1;
        else
            b.items[i].result = b.pool[b.solution + (unsigned)i] == 0 ? 1 : 0;
    }

    ox_std_sort(b.items, r_size, sizeof(ox_bb_item),
                maximize ? bb_result_less_max : bb_result_less_min, NULL);

    int middle = 0;
    while (middle != r_size && b.items[middle].result != 0)
        ++middle;

    for (int i = 0; i != r_size; ++i) {
        R[i].value = b.items[i].r;
        R[i].id = b.items[i].variable;
        R[i].f = b.items[i].original_factor;
    }

    int ret;
    if (middle == r_size)
        ret = b.items[0].result == 0 ? -1 : r_size;
    else
        ret = middle - 1;

    free(b.items);
    free(b.pool);
    free(b.nodes);
    return ret;
}

int
ox_bb_solve(int maximize, double* value, int* factor, int* id, int n, int bmin,
            int bmax)
{
    ox_rc* R = (ox_rc*)malloc(sizeof(ox_rc) * (size_t)n);
    for (int i = 0; i != n; ++i) {
        R[i].value = value[i];
        R[i].id = id[i];
        R[i].f = factor[i];
    }
    int ret = bb_solve(maximize, R, n, bmin, bmax);
    for (int i = 0; i != n; ++i) {
        value[i] = R[i].value;
        id[i] = R[i].id;
        factor[i] = R[i].f;
    }
    free(R);
    return ret;
}

/* ======================================================================== */
/* Problem: sparse_matrix + per-solver bounds                               */
/* ======================================================================== */

enum
{
    Z_BRANCH_AND_BOUND = 0,
    Z_EXHAUSTIVE,
    Z_UNCONSTRAINED_BRANCH_AND_BOUND,
    Z_UNCONSTRAINED_EXHAUSTIVE,
    Z_LINEAR
};

enum
{
    FAM_01 = 0,
    FAM_101 = 1,
    FAM_Z = 2
};

struct ox_problem
{
    int m, n, nnz;
    int problem_type, family, maximize;
    /* CSR: row value id == position (sparse-matrix.hpp:144) */
    int* row_ptr;
    int* row_col;
    int* A; /* factor per nnz id (row-major) */
    /* CSC: (value = row-major nnz id, row); rows ascending inside a column
     * (stable_sort by column, sparse-matrix.hpp:155-174) */
    int* col_ptr;
    int* col_val;
    int* col_row;
    int* cmin;
    int* cmax; /* merged constraint min/max as given */
    int* bmin;
    int* bmax; /* solver bounds after the constructor's clamps */
    int* ztype;
    int** ex_flat;
    int* ex_solutions;
    int max_row;
    int n_obj;
    int* obj_var;
    double* obj_factor;
    double cost_constant;
    double* cost; /* default_cost_type::linear_elements, itm-common.hpp:1006-1016 */
    double tail_value;
};

int
ox_which_problem_type(int nnz, const int* factor, int has_inequality)
{
    /* coefficient_type, lib/src/problem.hpp:72-86 */
    int coefficient = 0;
    for (int i = 0; i != nnz; ++i) {
        if (factor[i] == -1 && coefficient == 0)
            coefficient = 1;
        else if (factor[i] <= -2 || factor[i] >= 2) {
            coefficient = 2;
            break;
        }
    }
    if (!has_inequality)
        return coefficient == 0 ? OX_EQ_01
                                : coefficient == 1 ? OX_EQ_101 : OX_EQ_Z;
    return coefficient == 0 ? OX_INEQ_01
                            : coefficient == 1 ? OX_INEQ_101 : OX_INEQ_Z;
}

ox_problem*
ox_problem_create(int m, int n, const int* row_ptr, const int* col,
                  const int* factor, const int* cmin, const int* cmax,
                  int problem_type, int maximize, int n_obj,
                  const int* obj_var, const double* obj_factor,
                  double cost_constant)
{
    ox_problem* pb = (ox_problem*)calloc(1, sizeof(ox_problem));
    const int nnz = row_ptr[m];

    pb->m = m;
    pb->n = n;
    pb->nnz = nnz;
    pb->problem_type = problem_type;
    pb->maximize = maximize;
    pb->family = (problem_type == OX_EQ_01 || problem_type == OX_INEQ_01)
                   ? FAM_01
                   : (problem_type == OX_EQ_101 || problem_type == OX_INEQ_101)
                       ? FAM_101
                       : FAM_Z;

    pb->row_ptr = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    pb->row_col = (int*)malloc(sizeof(int) * (size_t)(nnz + 1));
    pb->A = (int*)malloc(sizeof(int) * (size_t)(nnz + 1));
    memcpy(pb->row_ptr, row_ptr, sizeof(int) * (size_t)(m + 1));
    memcpy(pb->row_col, col, sizeof(int) * (size_t)nnz);
    memcpy(pb->A, factor, sizeof(int) * (size_t)nnz);

    /* CSC by counting sort on column = stable w.r.t. row-major order */
    pb->col_ptr = (int*)calloc((size_t)(n + 1), sizeof(int));
    pb->col_val = (int*)malloc(sizeof(int) * (size_t)(nnz + 1));
    pb->col_row = (int*)malloc(sizeof(int) * (size_t)(nnz + 1));
    for (int e = 0; e != nnz; ++e)
        pb->col_ptr[col[e] + 1]++;
    for (int j = 0; j != n; ++j)
        pb->col_ptr[j + 1] += pb->col_ptr[j];
    {
        int* fill = (int*)malloc(sizeof(int) * (size_t)(n + 1));
        memcpy(fill, pb->col_ptr, sizeof(int) * (size_t)(n + 1));
        for (int k = 0; k != m; ++k)
            for (int e = row_ptr[k]; e != row_ptr[k + 1]; ++e) {
                const int pos = fill[col[e]]++;
                pb->col_val[pos] = e;
                pb->col_row[pos] = k;
            }
        free(fill);
    }

    pb->cmin = (int*)malloc(sizeof(int) * (size_t)m);
    pb->cmax = (int*)malloc(sizeof(int) * (size_t)m);
    pb->bmin = (int*)malloc(sizeof(int) * (size_t)m);
    pb->bmax = (int*)malloc(sizeof(int) * (size_t)m);
    pb->ztype = (int*)malloc(sizeof(int) * (size_t)m);
    pb->ex_flat = (int**)calloc((size_t)m, sizeof(int*));
    pb->ex_solutions = (int*)calloc((size_t)m, sizeof(int));
    memcpy(pb->cmin, cmin, sizeof(int) * (size_t)m);
    memcpy(pb->cmax, cmax, sizeof(int) * (size_t)m);

    /* compute_reduced_costs_vector_size, itm-common.hpp:235-245 */
    pb->max_row = 0;
    for (int k = 0; k != m; ++k)
        if (row_ptr[k + 1] - row_ptr[k] > pb->max_row)
            pb->max_row = row_ptr[k + 1] - row_ptr[k];

    for (int k = 0; k != m; ++k) {
        const int len = row_ptr[k + 1] - row_ptr[k];
        const int* f = factor + row_ptr[k];
        pb->ztype[k] = Z_LINEAR;

        switch (problem_type) {
        case OX_EQ_01: /* itm-solver-equalities-01.cpp:84-96 */
        case OX_EQ_101: /* itm-solver-equalities-101.cpp:98-112 */
            pb->bmin[k] = cmin[k];
            pb->bmax[k] = cmin[k];
            break;
        case OX_INEQ_01: /* itm-solver-inequalities-01.cpp:88-106 */
            if (cmin[k] == cmax[k]) {
                pb->bmin[k] = cmin[k];
                pb->bmax[k] = cmax[k];
            } else {
                pb->bmin[k] = cmin[k] > 0 ? cmin[k] : 0;
                pb->bmax[k] = len < cmax[k] ? len : cmax[k];
            }
            break;
        case OX_INEQ_101: { /* itm-solver-inequalities-101.cpp:99-129 */
            int lower = 0, upper = 0;
            for (int i = 0; i != len; ++i) {
                if (f[i] > 0)
                    upper++;
                else
                    lower++;
            }
            if (cmin[k] == cmax[k]) {
                pb->bmin[k] = cmin[k];
                pb->bmax[k] = cmax[k];
            } else {
                pb->bmin[k] = -lower > cmin[k] ? -lower : cmin[k];
                pb->bmax[k] = upper < cmax[k] ? upper : cmax[k];
            }
        } break;
        default: { /* itm-solver-inequalities-Z.cpp:125-183 */
            int lower = 0, upper = 0;
            for (int i = 0; i != len; ++i) {
                if (f[i] > 0)
                    upper += f[i];
                else
                    lower += f[i];

                if (f[i] < -1 || f[i] > 1)
                    pb->ztype[k] = Z_BRANCH_AND_BOUND;
            }

            if (pb->ztype[k] == Z_BRANCH_AND_BOUND && len < 10)
                pb->ztype[k] = Z_EXHAUSTIVE;

            if (cmin[k] == cmax[k]) {
                pb->bmin[k] = cmin[k];
                pb->bmax[k] = cmax[k];
            } else {
                if (lower >= cmin[k] || upper <= cmax[k]) {
                    if (pb->ztype[k] == Z_BRANCH_AND_BOUND)
                        pb->ztype[k] = Z_UNCONSTRAINED_BRANCH_AND_BOUND;
                    else if (pb->ztype[k] == Z_EXHAUSTIVE)
                        pb->ztype[k] = Z_UNCONSTRAINED_EXHAUSTIVE;
                }

                pb->bmin[k] = lower > cmin[k] ? lower : cmin[k];
                pb->bmax[k] = upper < cmax[k] ? upper : cmax[k];
            }

            if (pb->ztype[k] == Z_EXHAUSTIVE ||
                pb->ztype[k] == Z_UNCONSTRAINED_EXHAUSTIVE)
                pb->ex_solutions[k] =
                  ex_build(f, len, pb->bmin[k], pb->bmax[k], &pb->ex_flat[k]);
        } break;
        }
    }

    /* bx_ensures(b[i].min <= b[i].max), itm-solver-inequalities-Z.cpp:182 and
     * the like: the reference terminates on an unrealisable row */
    for (int k = 0; k != m; ++k) {
        if (pb->bmin[k] > pb->bmax[k]) {
            fprintf(stderr, "oracle: constraint %d is unrealisable (%d > %d)\n", k,
                    pb->bmin[k], pb->bmax[k]);
            pb->n_obj = 0;
            ox_problem_destroy(pb);
            return NULL;
        }
    }

    pb->n_obj = n_obj;
    pb->obj_var = (int*)malloc(sizeof(int) * (size_t)(n_obj + 1));
    pb->obj_factor = (double*)malloc(sizeof(double) * (size_t)(n_obj + 1));
    memcpy(pb->obj_var, obj_var, sizeof(int) * (size_t)n_obj);
    memcpy(pb->obj_factor, obj_factor, sizeof(double) * (size_t)n_obj);
    pb->cost_constant = cost_constant;

    pb->cost = (double*)calloc((size_t)(n + 1), sizeof(double));
    for (int i = 0; i != n_obj; ++i)
        pb->cost[obj_var[i]] += (double)obj_factor[i];

    return pb;
}

void
ox_problem_set_tail_value(ox_problem* pb, double v)
{
    pb->tail_value = v;
}

void
ox_problem_destroy(ox_problem* pb)
{
    if (!pb)
        return;
    for (int k = 0; k != pb->m; ++k)
        free(pb->ex_flat[k]);
    free(pb->ex_flat);
    free(pb->ex_solutions);
    free(pb->row_ptr);
    free(pb->row_col);
    free(pb->A);
    free(pb->col_ptr);
    free(pb->col_val);
    free(pb->col_row);
    free(pb->cmin);
    free(pb->cmax);
    free(pb->bmin);
    free(pb->bmax);
    free(pb->ztype);
    free(pb->obj_var);
    free(pb->obj_factor);
    free(pb->cost);
    free(pb);
}

/* ======================================================================== */
/* Run state: Solver (P, pi, R) + x + compute_order                         */
/* ======================================================================== */

struct ox_run
{
    ox_problem* pb;
    ox_params p;
    uint32_t rng;
    double* norm_cost;
    double theta, delta;
    double* P;
    double* pi;
    ox_rc* R; /* max_row + 1 entries: index max_row models the reference's
                 one-past-the-end read (SURVEY.md 7 hard part 3) as a slot
                 that is never written, i.e. 0.0 */
    uint64_t* x;
    int x_blocks;
    /* compute_order, itm-common.hpp:627-915 */
    int* oR;
    int oR_len;
    ox_pair* m_order;
    int m_order_len;
    int order;
    int use_cycle;
    int64_t updates;
    int64_t sweeps;
};

static int
xget(const ox_run* r, int i)
{
    return (int)((r->x[i / 64] >> (i % 64)) & 1u);
}

static void
xset(ox_run* r, int i)
{
    r->x[i / 64] |= (uint64_t)1 << (i % 64);
}

static void
xunset(ox_run* r, int i)
{
    r->x[i / 64] &= ~((uint64_t)1 << (i % 64));
}

static void
xinvert(ox_run* r, int i)
{
    r->x[i / 64] ^= (uint64_t)1 << (i % 64);
}

/* factor(value): 01 -> 1, 101/Z -> A[value] */
static int
row_factor(const ox_problem* pb, int value)
{
    return pb->family == FAM_01 ? 1 : pb->A[value];
}

/* is_valid_constraint, itm-common.hpp:76-89 (returns v through *v_out) */
static int
row_value(const ox_run* r, int k)
{
    const ox_problem* pb = r->pb;
    int v = 0;
    for (int e = pb->row_ptr[k]; e != pb->row_ptr[k + 1]; ++e)
        v += row_factor(pb, e) * xget(r, pb->row_col[e]);
    return v;
}

static int
is_valid_constraint(const ox_run* r, int k)
{
    const int v = row_value(r, k);
    return r->pb->bmin[k] <= v && v <= r->pb->bmax[k];
}

/* compute_violated_constraints, itm-common.hpp:102-115 */
static int
compute_violated_constraints(ox_run* r)
{
    r->oR_len = 0;
    for (int k = 0; k != r->pb->m; ++k)
        if (!is_valid_constraint(r, k))
            r->oR[r->oR_len++] = k;
    return r->oR_len;
}

/* local_compute_violated_constraints, itm-common.hpp:882-891 */
static int
local_compute_violated_constraints(ox_run* r)
{
    int remaining = 0;
    for (int k = 0; k != r->pb->m; ++k)
        if (!is_valid_constraint(r, k))
            ++remaining;
    return remaining;
}

/* infeasibility_local_compute_violated_constraints, itm-common.hpp:893-914 */
static int
infeasibility_local_compute_violated_constraints(ox_run* r)
{
    r->m_order_len = 0;
    for (int k = 0; k != r->pb->m; ++k) {
        const int v = row_value(r, k);
        if (r->pb->bmin[k] > v) {
            r->m_order[r->m_order_len].first = k;
            r->m_order[r->m_order_len].second = r->pb->bmin[k] - v;
            r->m_order_len++;
        } else if (r->pb->bmax[k] < v) {
            r->m_order[r->m_order_len].first = k;
            r->m_order[r->m_order_len].second = v - r->pb->bmax[k];
            r->m_order_len++;
        }
    }
    return r->m_order_len;
}

/* random_shuffle_unique, itm-common.hpp:117-132 */
static void
random_shuffle_unique(ox_rc* begin, ox_rc* end, uint32_t* rng)
{
    ox_rc* ret = begin++;
    for (; begin != end; ++begin) {
        if (ret->value != begin->value) {
            ox_std_shuffle(ret, begin - ret, sizeof(ox_rc), rng);
            ret = begin;
        }
    }
    ox_std_shuffle(ret, begin - ret, sizeof(ox_rc), rng);
}

/* calculator_sort, itm-common.hpp:134-148 */
static void
calculator_sort(const ox_run* r, ox_rc* R, int r_size, uint32_t* rng)
{
    if (r_size > 1) {
        ox_std_sort(R, r_size, sizeof(ox_rc),
                    r->pb->maximize ? rc_less_max : rc_less_min, NULL);
        random_shuffle_unique(R, R + r_size, rng);
    }
}

/* stop_iterating with rng, itm-common.hpp:150-163 */
static int
stop_iterating_rng(const ox_run* r, double value, uint32_t* rng)
{
    if (value == 0)
        return ox_bernoulli(0.5, rng);
    return r->pb->maximize ? value < 0 : value > 0;
}

/* stop_iterating without rng, itm-common.hpp:182-190 */
static int
stop_iterating(const ox_run* r, double value)
{
    return r->pb->maximize ? value < 0 : value > 0;
}

/* compute_reduced_costs: 01 itm-solver-inequalities-01.cpp:166-191,
 * 101 itm-solver-inequalities-101.cpp:199-233, Z
 * itm-solver-inequalities-Z.cpp:253-293.  Returns c_size through *c_size. */
static int
compute_reduced_costs(ox_run* r, int k, int* c_size_out)
{
    const ox_problem* pb = r->pb;
    int r_size = 0, c_size = 0;

    for (int e = pb->row_ptr[k]; e != pb->row_ptr[k + 1]; ++e) {
        const int j = pb->row_col[e];

        if (pb->family == FAM_01) {
            double sum_a_pi = 0, sum_a_p = 0;
            for (int t = pb->col_ptr[j]; t != pb->col_ptr[j + 1]; ++t) {
                sum_a_pi += r->pi[pb->col_row[t]];
                sum_a_p += r->P[pb->col_val[t]];
            }
            r->R[r_size].id = r_size;
            r->R[r_size].f = 1;
            r->R[r_size].value = r->norm_cost[j] - sum_a_pi - sum_a_p;
        } else if (pb->family == FAM_101) {
            double sum_a_pi = 0, sum_a_p = 0;
            for (int t = pb->col_ptr[j]; t != pb->col_ptr[j + 1]; ++t) {
                const double a = (double)pb->A[pb->col_val[t]];
                sum_a_pi += a * r->pi[pb->col_row[t]];
                sum_a_p += a * r->P[pb->col_val[t]];
            }
            r->R[r_size].id = r_size;
            r->R[r_size].value = r->norm_cost[j] - sum_a_pi - sum_a_p;
            r->R[r_size].f = pb->A[e];
            if (r->R[r_size].f < 0) {
                r->R[r_size].value = -r->R[r_size].value;
                ++c_size;
            }
        } else {
            double sum_a_pi_p = 0;
            for (int t = pb->col_ptr[j]; t != pb->col_ptr[j + 1]; ++t) {
                const double a = fabs((double)pb->A[pb->col_val[t]]);
                sum_a_pi_p += a * (r->pi[pb->col_row[t]] + r->P[pb->col_val[t]]);
            }
            r->R[r_size].id = r_size;
            r->R[r_size].f = pb->A[e];
            r->R[r_size].value = r->norm_cost[j] - sum_a_pi_p;
        }
        ++r_size;
    }

    *c_size_out = c_size;
    return r_size;
}

/* local_compute_reduced_cost, itm-solver-inequalities-Z.cpp:295-306 */
static double
local_compute_reduced_cost(const ox_run* r, int variable)
{
    const ox_problem* pb = r->pb;
    double sum_a_pi_p = 0;
    for (int t = pb->col_ptr[variable]; t != pb->col_ptr[variable + 1]; ++t) {
        const double a = fabs((double)pb->A[pb->col_val[t]]);
        sum_a_pi_p += a * (r->pi[pb->col_row[t]] + r->P[pb->col_val[t]]);
    }
    return r->norm_cost[variable] - sum_a_pi_p;
}

/* select_variables: 01 itm-solver-inequalities-01.cpp:193-206 (equalities-01
 * :173-182 is the bkmin == bkmax branch), 101
 * itm-solver-inequalities-101.cpp:235-248 (equalities-101 :197-206). */
static int
select_variables_01_101(ox_run* r, int r_size, int c_size, int bkmin, int bkmax)
{
    const int fam = r->pb->family;

    if (bkmin == bkmax) {
        const int b = bkmin + c_size;
        return (b < r_size ? b : r_size) - 1;
    }

    if (fam == FAM_01) {
        bkmin = bkmin < r_size ? bkmin : r_size;
        bkmax = bkmax < r_size ? bkmax : r_size;
    } else {
        bkmin += c_size;
        bkmax = (bkmax + c_size) < r_size ? (bkmax + c_size) : r_size;
    }

    for (int i = bkmin; i <= bkmax; ++i)
        if (stop_iterating_rng(r, r->R[i].value, &r->rng))
            return i - 1;

    return bkmax - 1;
}

/* select_variables / select_variables_101 of the Z solver (identical
 * bodies), itm-solver-inequalities-Z.cpp:308-344 */
static int
select_variables_Z(ox_run* r, int r_size, int bkmin, int bkmax)
{
    int sum = 0;
    int best = -2;

    for (int i = -1; i < r_size; ++i) {
        if (bkmin <= sum && sum <= bkmax)
            best = i;

        if (best != -2 && i - 1 < r_size &&
            stop_iterating(r, r->R[i + 1].value))
            break;

        sum += r->R[i + 1].f;
    }

    return best;
}

static int
is_signbit_change(double lhs, double rhs)
{
    return (signbit(lhs) != 0) != (signbit(rhs) != 0);
}

/* affect, itm-common.hpp:382-467 (01 and 101 solvers) */
static int
affect(ox_run* r, int k, int selected, int r_size, double kappa, double delta)
{
    const ox_problem* pb = r->pb;
    const double one = 1, two = 2;
    const double middle = (two + one) / two;
    const int it = pb->row_ptr[k];
    const double old_pi = r->pi[k];
    double d = delta;

    if (selected < 0) {
        d += (kappa / (one - kappa)) * (r->R[0].value / two);

        for (int i = 0; i != r_size; ++i) {
            const int var = it + r->R[i].id;
            if (r->R[i].f < 0) {
                xset(r, pb->row_col[var]);
                r->P[var] += d;
            } else {
                xunset(r, pb->row_col[var]);
                r->P[var] -= d;
            }
        }
    } else if (selected + 1 >= r_size) {
        d += (kappa / (one - kappa)) * (r->R[selected].value * middle);

        for (int i = 0; i != r_size; ++i) {
            const int var = it + r->R[i].id;
            if (r->R[i].f < 0) {
                xunset(r, pb->row_col[var]);
                r->P[var] -= d;
            } else {
                xset(r, pb->row_col[var]);
                r->P[var] += d;
            }
        }
    } else {
        r->pi[k] += ((r->R[selected].value + r->R[selected + 1].value) / two);
        d += (kappa / (one - kappa)) *
             (r->R[selected + 1].value - r->R[selected].value);

        int i = 0;
        for (; i <= selected; ++i) {
            const int var = it + r->R[i].id;
            if (r->R[i].f < 0) {
                xunset(r, pb->row_col[var]);
                r->P[var] -= d;
            } else {
                xset(r, pb->row_col[var]);
                r->P[var] += d;
            }
        }

        for (; i != r_size; ++i) {
            const int var = it + r->R[i].id;
            if (r->R[i].f < 0) {
                xset(r, pb->row_col[var]);
                r->P[var] += d;
            } else {
                xunset(r, pb->row_col[var]);
                r->P[var] -= d;
            }
        }
    }

    return is_signbit_change(old_pi, r->pi[k]);
}

/* local_affect, itm-solver-inequalities-Z.cpp:346-439 */
static int
local_affect(ox_run* r, int k, int selected, int r_size, double kappa,
             double delta)
{
    const ox_problem* pb = r->pb;
    const double one = 1, two = 2;
    const double middle = (two + one) / two;
    const int it = pb->row_ptr[k];
    const double old_pi = r->pi[k];
    const double d = (kappa / (one - kappa)) + delta;

    if (selected < 0) {
        r->pi[k] += r->R[0].value / two;

        for (int i = 0; i != r_size; ++i) {
            const int var = it + r->R[i].id;
            xunset(r, pb->row_col[var]);
            r->P[var] -= d;

            const double repair =
              local_compute_reduced_cost(r, pb->row_col[var]);
            if (repair <= 0)
                r->P[var] = r->P[var] + repair - d;
        }
    } else if (selected + 1 >= r_size) {
        r->pi[k] += r->R[selected].value * middle;

        for (int i = 0; i != r_size; ++i) {
            const int var = it + r->R[i].id;
            xset(r, pb->row_col[var]);
            r->P[var] += d;

            const double repair =
              local_compute_reduced_cost(r, pb->row_col[var]);
            if (repair >= 0)
                r->P[var] = r->P[var] - repair + d;
        }
    } else {
        r->pi[k] += ((r->R[selected].value + r->R[selected + 1].value) / two);

        int i = 0;
        for (; i <= selected; ++i) {
            const int var = it + r->R[i].id;
            xset(r, pb->row_col[var]);
            r->P[var] += d;

            const double repair =
              local_compute_reduced_cost(r, pb->row_col[var]);
            if (repair >= 0)
                r->P[var] = r->P[var] - repair + d;
        }

        for (; i != r_size; ++i) {
            const int var = it + r->R[i].id;
            xunset(r, pb->row_col[var]);
            r->P[var] -= d;

            const double repair =
              local_compute_reduced_cost(r, pb->row_col[var]);
            if (repair <= 0)
                r->P[var] = r->P[var] + repair - d;
        }
    }

    return is_signbit_change(old_pi, r->pi[k]);
}

/* One constraint update = the loop body of compute_update_row /
 * push_and_compute_update_row: 01 itm-solver-inequalities-01.cpp:208-296,
 * 101 itm-solver-inequalities-101.cpp:250-347,
 * Z itm-solver-inequalities-Z.cpp:441-564.  push != 0 applies the objective
 * amplifier before the sort. */
static int
update_row(ox_run* r, int k, double kappa, double delta, double theta,
           int push, double obj_amp)
{
    const ox_problem* pb = r->pb;
    const int it = pb->row_ptr[k];
    int c_size = 0;

    r->updates++;

    /* decrease_preference, itm-solver-inequalities-01.cpp:158-164 */
    for (int e = pb->row_ptr[k]; e != pb->row_ptr[k + 1]; ++e)
        r->P[e] *= theta;

    const int r_size = compute_reduced_costs(r, k, &c_size);

    if (push) {
        if (pb->family == FAM_01) {
            /* cost indexed by the LOCAL position R[i].id, reference quirk:
             * itm-solver-inequalities-01.cpp:236-237 */
            for (int i = 0; i != r_size; ++i)
                r->R[i].value += obj_amp * r->norm_cost[r->R[i].id];
        } else {
            /* itm-solver-inequalities-101.cpp:274-276, -Z.cpp:467-469 */
            for (int i = 0; i != r_size; ++i)
                r->R[i].value +=
                  obj_amp * r->norm_cost[pb->row_col[it + r->R[i].id]];
        }
    }

    if (pb->family != FAM_Z) {
        calculator_sort(r, r->R, r_size, &r->rng);
        const int selected =
          select_variables_01_101(r, r_size, c_size, pb->bmin[k], pb->bmax[k]);
        return affect(r, k, selected, r_size, kappa, delta);
    }

    int selected = -2;
    switch (pb->ztype[k]) {
    case Z_BRANCH_AND_BOUND:
        calculator_sort(r, r->R, r_size, &r->rng);
        selected = bb_solve(pb->maximize, r->R, r_size, pb->bmin[k], pb->bmax[k]);
        break;
    case Z_EXHAUSTIVE:
        selected = ex_solve(pb->maximize, pb->ex_flat[k], pb->ex_solutions[k],
                            r->R, r_size);
        break;
    case Z_UNCONSTRAINED_BRANCH_AND_BOUND:
        calculator_sort(r, r->R, r_size, &r->rng);
        selected = select_variables_Z(r, r_size, pb->bmin[k], pb->bmax[k]);
        if (selected == -2)
            selected =
              bb_solve(pb->maximize, r->R, r_size, pb->bmin[k], pb->bmax[k]);
        break;
    case Z_UNCONSTRAINED_EXHAUSTIVE:
        calculator_sort(r, r->R, r_size, &r->rng);
        selected = select_variables_Z(r, r_size, pb->bmin[k], pb->bmax[k]);
        if (selected == -2)
            selected = ex_solve(pb->maximize, pb->ex_flat[k],
                                pb->ex_solutions[k], r->R, r_size);
        break;
    case Z_LINEAR:
    default:
        calculator_sort(r, r->R, r_size, &r->rng);
        selected = select_variables_Z(r, r_size, pb->bmin[k], pb->bmax[k]);
        break;
    }

    return local_affect(r, k, selected, r_size, kappa, delta);
}

/* ---- compute_order ----------------------------------------------------- */

static int
pi_less_decr(const void* a, const void* b, void* ctx)
{
    const double* pi = (const double*)ctx;
    return pi[*(const int*)b] < pi[*(const int*)a];
}

static int
pi_less_incr(const void* a, const void* b, void* ctx)
{
    const double* pi = (const double*)ctx;
    return pi[*(const int*)a] < pi[*(const int*)b];
}

/* compute_order::init, itm-common.hpp:660-680 */
static void
order_init(ox_run* r)
{
    switch (r->order) {
    case OX_ORDER_INFEASIBILITY_DECR:
    case OX_ORDER_INFEASIBILITY_INCR:
        infeasibility_local_compute_violated_constraints(r);
        break;
    case OX_ORDER_PI_SIGN_CHANGE:
        /* std::iota over R, whose size is whatever it currently is: n at
         * construction (itm-common.hpp:636), valid only when n <= m. */
        for (int i = 0; i != r->oR_len; ++i)
            r->oR[i] = i;
        break;
    default:
        compute_violated_constraints(r);
        break;
    }
}

/* compute_order::run / push_and_run, itm-common.hpp:682-880 */
static int
order_run(ox_run* r, double kappa, double delta, double theta, int push,
          double obj_amp)
{
    int pi_changed = 0;
    int remaining = 0;

    r->sweeps++;

    if (push && r->use_cycle) {
        /* next_state, itm-common.hpp:644-658 */
        int next = r->order + 1;
        r->order = next >= OX_ORDER_CYCLE ? 0 : next;
    }

    switch (r->order) {
    case OX_ORDER_REVERSING:
        for (int i = r->oR_len - 1; i >= 0; --i)
            update_row(r, r->oR[i], kappa, delta, theta, push, obj_amp);
        return compute_violated_constraints(r);
    case OX_ORDER_RANDOM_SORTING:
        ox_std_shuffle(r->oR, r->oR_len, sizeof(int), &r->rng);
        for (int i = 0; i != r->oR_len; ++i)
            update_row(r, r->oR[i], kappa, delta, theta, push, obj_amp);
        return compute_violated_constraints(r);
    case OX_ORDER_INFEASIBILITY_DECR:
    case OX_ORDER_INFEASIBILITY_INCR:
        ox_std_sort(r->m_order, r->m_order_len, sizeof(ox_pair),
                    r->order == OX_ORDER_INFEASIBILITY_DECR ? pair_less_decr
                                                            : pair_less_incr,
                    NULL);
        for (int i = 0; i != r->m_order_len; ++i)
            update_row(r, r->m_order[i].first, kappa, delta, theta, push,
                       obj_amp);
        return infeasibility_local_compute_violated_constraints(r);
    case OX_ORDER_LAGRANGIAN_DECR:
    case OX_ORDER_LAGRANGIAN_INCR:
        ox_std_sort(r->oR, r->oR_len, sizeof(int),
                    r->order == OX_ORDER_LAGRANGIAN_DECR ? pi_less_decr
                                                         : pi_less_incr,
                    r->pi);
        for (int i = 0; i != r->oR_len; ++i)
            update_row(r, r->oR[i], kappa, delta, theta, push, obj_amp);
        return compute_violated_constraints(r);
    case OX_ORDER_PI_SIGN_CHANGE:
        ox_std_shuffle(r->oR, r->oR_len, sizeof(int), &r->rng);
        for (int i = 0; i != r->oR_len; ++i)
            if (update_row(r, r->oR[i], kappa, delta, theta, push, obj_amp))
                pi_changed = 1;
        remaining = local_compute_violated_constraints(r);
        if (!pi_changed && remaining == 0)
            return 0;
        return remaining;
    case OX_ORDER_NONE:
    default:
        for (int i = 0; i != r->oR_len; ++i)
            update_row(r, r->oR[i], kappa, delta, theta, push, obj_amp);
        return compute_violated_constraints(r);
    }
}

/* ---- costs -------------------------------------------------------------- */

/* normalize_costs, itm-common.hpp:967-998 with make_l1/l2/loo_norm
 * :1095-1127.  cost_norm "random" (:1025-1082) is not restated. */
static int
normalize_costs(const ox_problem* pb, int cost_norm, double* out)
{
    const int n = pb->n;
    memcpy(out, pb->cost, sizeof(double) * (size_t)n);

    switch (cost_norm) {
    case 0: /* none */
        return 0;
    case 1: /* random */
        return -1;
    case 2: { /* l1 */
        double div = 0;
        for (int i = 0; i != n; ++i)
            div += fabs(out[i]);
        if (isnormal(div))
            for (int i = 0; i != n; ++i)
                out[i] /= div;
        return 0;
    }
    case 3: { /* l2 (sum of squares, no square root: reference behaviour) */
        double div = 0;
        for (int i = 0; i != n; ++i)
            div += out[i] * out[i];
        if (isnormal(div))
            for (int i = 0; i != n; ++i)
                out[i] /= div;
        return 0;
    }
    default: { /* loo: max element, NOT max abs (itm-common.hpp:1119-1127) */
        double div = out[0];
        for (int i = 1; i < n; ++i)
            if (div < out[i]) /* std::max_element keeps the first maximum */
                div = out[i];
        if (isnormal(div))
            for (int i = 0; i != n; ++i)
                out[i] /= div;
        return 0;
    }
    }
}

/* compute_delta, itm-common.hpp:917-933 with Cost::min :1084-1093 */
static double
compute_delta(const double* c, double theta, int n)
{
    double mini = DBL_MAX;
    for (int i = 0; i != n; ++i)
        if (c[i])
            mini = mini < fabs(c[i]) ? mini : fabs(c[i]);
    return mini - theta * mini;
}

/* default_cost_type::results, itm-common.hpp:1140-1147 */
static double
results(const ox_run* r)
{
    const ox_problem* pb = r->pb;
    double v = pb->cost_constant;
    for (int i = 0; i != pb->n_obj; ++i)
        if (xget(r, pb->obj_var[i]))
            v += pb->obj_factor[i];
    return v;
}

/* ---- x initialisation --------------------------------------------------- */

/* init_x, itm-common.hpp:202-221 */
static int
init_x(int maximize, double cost, int value_if_cost_0)
{
    if (!maximize) {
        if (cost < 0)
            return 1;
        if (cost == 0)
            return value_if_cost_0;
    } else {
        if (cost > 0)
            return 1;
        if (cost == 0)
            return value_if_cost_0;
    }
    return 0;
}

typedef struct ox_pre_rc
{
    float value;
    int factor;
    int id;
} ox_pre_rc;

static int
pre_less_min(const void* a, const void* b, void* ctx)
{
    (void)ctx;
    return ((const ox_pre_rc*)a)->value < ((const ox_pre_rc*)b)->value;
}

static int
pre_less_max(const void* a, const void* b, void* ctx)
{
    (void)ctx;
    return ((const ox_pre_rc*)a)->value > ((const ox_pre_rc*)b)->value;
}

/* init_with_pre_solve, itm-common.hpp:284-374; exactly one of pessimistic /
 * optimistic is filled (solver_functor passes an empty bit_array for the
 * other, itm-solver-common.hpp:104-113).  Uses the ORIGINAL costs and the
 * merged constraints' raw min/max. */
static void
init_with_pre_solve(ox_run* r, int optimistic, double init_random)
{
    const ox_problem* pb = r->pb;
    ox_pre_rc* R =
      (ox_pre_rc*)malloc(sizeof(ox_pre_rc) * (size_t)(pb->max_row + 1));

    for (int k = 0; k != pb->m; ++k) {
        if (!ox_bernoulli(init_random, &r->rng))
            continue;

        const int r_size = pb->row_ptr[k + 1] - pb->row_ptr[k];
        for (int i = 0; i != r_size; ++i) {
            const int e = pb->row_ptr[k] + i;
            R[i].value = (float)pb->cost[pb->row_col[e]];
            R[i].factor = pb->A[e];
            R[i].id = pb->row_col[e];
        }

        ox_std_shuffle(R, r_size, sizeof(ox_pre_rc), &r->rng);
        ox_std_sort(R, r_size, sizeof(ox_pre_rc),
                    pb->maximize ? pre_less_max : pre_less_min, NULL);

        int sum = 0;
        int best = -2;

        if (!optimistic) {
            for (int i = -1; i < r_size; ++i) {
                if (pb->cmin[k] <= sum && sum <= pb->cmax[k]) {
                    best = i;
                    break;
                }
                if (i + 1 < r_size)
                    sum += R[i + 1].factor;
            }
        } else {
            for (int i = -1; i < r_size; ++i) {
                if (pb->cmin[k] <= sum && sum <= pb->cmax[k])
                    best = i;

                if (best != -2 && i + 1 < r_size &&
                    (pb->maximize ? R[i + 1].value < 0 : R[i + 1].value > 0))
                    break;

                if (i + 1 < r_size)
                    sum += R[i + 1].factor;
            }
        }

        int i = 0;
        for (; i <= best; ++i)
            xset(r, R[i].id);
        for (; i != r_size; ++i)
            xunset(r, R[i].id);
    }

    free(R);
}

/* ---- run ---------------------------------------------------------------- */

ox_run*
ox_run_begin(ox_problem* pb, const ox_params* p, uint32_t seed)
{
    ox_run* r = (ox_run*)calloc(1, sizeof(ox_run));
    const int n = pb->n, m = pb->m;

    r->pb = pb;
    r->p = *p;
    /* linear_congruential_engine::seed: 0 mod m is mapped to 1 */
    r->rng = seed % OX_RNG_M;
    if (r->rng == 0)
        r->rng = 1;

    r->norm_cost = (double*)malloc(sizeof(double) * (size_t)(n + 1));
    if (normalize_costs(pb, p->cost_norm, r->norm_cost) != 0) {
        free(r->norm_cost);
        free(r);
        return NULL;
    }

    r->theta = p->theta;
    r->delta =
      p->delta < 0 ? compute_delta(r->norm_cost, r->theta, n) : p->delta;

    r->P = (double*)calloc((size_t)(pb->nnz + 1), sizeof(double));
    r->pi = (double*)calloc((size_t)(m + 1), sizeof(double));
    r->R = (ox_rc*)calloc((size_t)(pb->max_row + 2), sizeof(ox_rc));
    r->R[pb->max_row].value = pb->tail_value;
    r->x_blocks = n / 64 + 1; /* bit-array.hpp:60-66 */
    r->x = (uint64_t*)calloc((size_t)r->x_blocks, sizeof(uint64_t));

    /* compute_order ctor, itm-common.hpp:634-642: R and m_order start with
     * variable_number zero entries */
    {
        const int cap = (n > m ? n : m) + 1;
        r->oR = (int*)calloc((size_t)cap, sizeof(int));
        r->m_order = (ox_pair*)calloc((size_t)cap, sizeof(ox_pair));
        r->oR_len = n;
        r->m_order_len = n;
        r->use_cycle = p->order == OX_ORDER_CYCLE;
        r->order = r->use_cycle ? OX_ORDER_NONE : p->order;
    }

    /* itm-solver-common.hpp:99-123 */
    switch (p->init_policy) {
    case 1: /* pessimistic_solve */
        init_with_pre_solve(r, 0, 1.0);
        break;
    case 2: /* optimistic_solve */
        init_with_pre_solve(r, 1, 1.0);
        break;
    default: /* bastert, itm-common.hpp:255-267 on the ORIGINAL costs */
        for (int i = 0; i != n; ++i)
            if (init_x(pb->maximize, pb->cost[i], 0))
                xset(r, i);
            else
                xunset(r, i);
        break;
    }

    for (int i = 0; i != n; ++i)
        if (ox_bernoulli(p->init_policy_random, &r->rng))
            xinvert(r, i);

    order_init(r);

    return r;
}

void
ox_run_end(ox_run* r)
{
    if (!r)
        return;
    free(r->norm_cost);
    free(r->P);
    free(r->pi);
    free(r->R);
    free(r->x);
    free(r->oR);
    free(r->m_order);
    free(r);
}

int
ox_run_info(ox_run* r, double* delta, double* theta)
{
    if (delta)
        *delta = r->delta;
    if (theta)
        *theta = r->theta;
    return 0;
}

int
ox_run_sweep(ox_run* r, int push, double kappa, double obj_amp)
{
    return order_run(r, kappa, r->delta, r->theta, push, obj_amp);
}

int
ox_run_state(ox_run* r, uint64_t* x_words, double* pi, double* P,
             uint32_t* rng_state, int* order_list, int* order_len,
             double* norm_cost)
{
    const int infeas = r->order == OX_ORDER_INFEASIBILITY_DECR ||
                       r->order == OX_ORDER_INFEASIBILITY_INCR;
    if (x_words)
        memcpy(x_words, r->x, sizeof(uint64_t) * (size_t)r->x_blocks);
    if (pi)
        memcpy(pi, r->pi, sizeof(double) * (size_t)r->pb->m);
    if (P)
        memcpy(P, r->P, sizeof(double) * (size_t)r->pb->nnz);
    if (rng_state)
        *rng_state = r->rng;
    if (norm_cost)
        memcpy(norm_cost, r->norm_cost, sizeof(double) * (size_t)r->pb->n);
    if (order_len)
        *order_len = infeas ? r->m_order_len : r->oR_len;
    if (order_list) {
        if (infeas)
            memcpy(order_list, r->m_order,
                   sizeof(ox_pair) * (size_t)r->m_order_len);
        else
            memcpy(order_list, r->oR, sizeof(int) * (size_t)r->oR_len);
    }
    return 0;
}

int
ox_run_set_x(ox_run* r, const uint64_t* x_words)
{
    memcpy(r->x, x_words, sizeof(uint64_t) * (size_t)r->x_blocks);
    order_init(r);
    return 0;
}

int
ox_run_set_rng(ox_run* r, uint32_t state)
{
    r->rng = state;
    return 0;
}

double
ox_run_objective(ox_run* r)
{
    return results(r);
}

int64_t
ox_run_updates(ox_run* r)
{
    return r->updates;
}

/* ======================================================================== */
/* solver_functor::operator(), itm-solver-common.hpp:63-225                 */
/* ======================================================================== */

typedef struct ox_best
{
    uint64_t* x;
    double value;
    int64_t loop;
    int remaining;
    int has_x;
} ox_best;

/* store_if_better(x, int remaining, i), itm-solver-common.hpp:242-250 */
static void
store_if_better_remaining(ox_run* r, ox_best* b, int remaining, int64_t i)
{
    if (b->remaining > remaining) {
        memcpy(b->x, r->x, sizeof(uint64_t) * (size_t)r->x_blocks);
        b->has_x = 1;
        b->loop = i;
        b->remaining = remaining;
    }
}

/* store_if_better(x, double current, i), itm-solver-common.hpp:252-261 */
static void
store_if_better_value(ox_run* r, ox_best* b, double current, int64_t i)
{
    if (r->pb->maximize ? current > b->value : current < b->value) {
        memcpy(b->x, r->x, sizeof(uint64_t) * (size_t)r->x_blocks);
        b->has_x = 1;
        b->loop = i;
        b->remaining = 0;
        b->value = current;
    }
}

/* The loop of solver_functor::operator() (itm-solver-common.hpp:135-225) on a
 * run that is ready to sweep.  optimizer = 1 gives one call of a worker of
 * optimize_functor::operator() instead (itm-optimizer-common.hpp:658-750): it
 * starts at kappa_start (what reinit returned), the push sweep runs at kappa =
 * pushing_k_factor and every push restarts the schedule from kappa_start.
 * `best` tracks the best solution of the run (the optimizer offers each one to
 * the recorder as it goes; callers that need that do it per sweep). */
static int
solve_loop(ox_run* r, double kappa_start, int optimizer, ox_best* best)
{
    const ox_problem* pb = r->pb;
    const ox_params* p = &r->p;
    int status = OX_STATUS_UNINITIALIZED;
    int best_remaining = INT_MAX;

    const double kappa_step = p->kappa_step;
    const double kappa_max = p->kappa_max;
    const double alpha = p->alpha;
    const double theta = r->theta;
    const double delta = r->delta;
    const double pushing_k_factor = p->pushing_k_factor;
    const double pushing_objective_amplifier = p->pushing_objective_amplifier;
    const long w_limit = (long)p->w;

    int start_push = 0;
    double kappa = kappa_start;

    for (int64_t i = 0; i != p->limit; ++i) {
        int remaining = order_run(r, kappa, delta, theta, 0, 0.0);

        if (remaining == 0) {
            store_if_better_value(r, best, results(r), i);
            best_remaining = remaining;
            start_push = 1;
            break;
        }

        if (remaining < best_remaining) {
            store_if_better_remaining(r, best, remaining, i);
            best_remaining = remaining;
        }

        if (i > w_limit)
            kappa += kappa_step * pow((double)remaining / (double)pb->m, alpha);

        if (kappa > kappa_max) {
            status = OX_STATUS_KAPPA_MAX_REACHED;
            break;
        }
    }

    if (!start_push) {
        status = OX_STATUS_LIMIT_REACHED;
    } else {
        for (int push = 0; push < p->pushes_limit; ++push) {
            int remaining =
              order_run(r, optimizer ? pushing_k_factor : pushing_k_factor * kappa,
                        delta, theta, 1, pushing_objective_amplifier);

            if (remaining == 0)
                store_if_better_value(
                  r, best, results(r),
                  -(int64_t)push * p->pushing_iteration_limit - 1);

            if (optimizer)
                kappa = kappa_start;
            for (int iter = 0; iter < p->pushing_iteration_limit; ++iter) {
                remaining = order_run(r, kappa, delta, theta, 0, 0.0);

                if (remaining == 0) {
                    store_if_better_value(
                      r, best, results(r),
                      -(int64_t)push * p->pushing_iteration_limit - iter - 1);
                    break;
                }

                if (iter > p->w)
                    kappa += kappa_step *
                             pow((double)remaining / (double)pb->m, alpha);

                if (kappa > kappa_max)
                    break;
            }
        }
    }

    if (best->remaining == 0)
        status = OX_STATUS_SUCCESS;
    return status;
}

int
ox_solve(ox_problem* pb, const ox_params* p, uint32_t seed, int8_t* x,
         double* value, int* remaining_out, int64_t* loop, int64_t* updates,
         int64_t* sweeps)
{
    ox_run* r = ox_run_begin(pb, p, seed);
    if (!r)
        return OX_STATUS_INTERNAL_ERROR;

    ox_best best;
    best.x = (uint64_t*)calloc((size_t)r->x_blocks, sizeof(uint64_t));
    best.value = pb->maximize ? -INFINITY : INFINITY; /* bad_value */
    best.loop = 0;
    best.remaining = INT_MAX;
    best.has_x = 0;

    const int status = solve_loop(r, p->kappa_min, 0, &best);

    if (x) {
        for (int i = 0; i != pb->n; ++i)
            x[i] = best.has_x ? (int8_t)((best.x[i / 64] >> (i % 64)) & 1u) : 0;
    }
    if (value)
        *value = best.has_x ? best.value : 0.0;
    if (remaining_out)
        *remaining_out = best.remaining;
    if (loop)
        *loop = best.loop;
    if (updates)
        *updates = r->updates;
    if (sweeps)
        *sweeps = r->sweeps;

    free(best.x);
    ox_run_end(r);
    return status;
}

/* solve_loop on an existing run (after ox_run_set_rng / ox_run_set_x): the
 * Solver state (P, pi) is whatever the run holds, as for a worker of the
 * optimizer between two calls. */
int
ox_run_solve(ox_run* r, double kappa_start, int optimizer, uint64_t* best_x_words,
             double* value, int* remaining_out, int64_t* loop)
{
    ox_best best;
    best.x = (uint64_t*)calloc((size_t)r->x_blocks, sizeof(uint64_t));
    best.value = r->pb->maximize ? -INFINITY : INFINITY;
    best.loop = 0;
    best.remaining = INT_MAX;
    best.has_x = 0;

    const int status = solve_loop(r, kappa_start, optimizer, &best);

    if (best_x_words)
        memcpy(best_x_words, best.x, sizeof(uint64_t) * (size_t)r->x_blocks);
    if (value)
        *value = best.has_x ? best.value : 0.0;
    if (remaining_out)
        *remaining_out = best.remaining;
    if (loop)
        *loop = best.loop;
    free(best.x);
    return status;
}

/* ======================================================================== */
/* Optimizer population: local_context, storage,          */
/* best_solution_recorder -- itm-optimizer-common.hpp:52-594                */
/* ======================================================================== */

typedef struct ox_member
{
    uint64_t* x;
    double value;
    uint64_t hash;
    int64_t loop;
    int remaining;
} ox_member;

struct ox_pop
{
    ox_problem* pb;
    ox_params p;
    int size;
    int blocks;
    int threads;
    int* indices;
    ox_member* data;
    uint64_t* bastert;
    uint64_t* random;
    double* kappa_append;
};

/* local_context, itm-optimizer-common.hpp:52-92 */
struct ox_pop_ctx
{
    uint32_t rng;
    ox_normal_dist choose_sol;
    ox_normal_dist variable_p;
    ox_normal_dist value_p;
    int bad_lo, bad_hi;
    double bastert_insertion;
    double improve_start, improve_increase, improve_stop;
    unsigned thread_id;
};

static uint32_t
engine_seed(uint64_t seed)
{
    /* linear_congruential_engine::seed: value mod m, 0 is mapped to 1 */
    uint32_t s = (uint32_t)(seed % OX_RNG_M);
    return s == 0 ? 1u : s;
}

static int
wget(const uint64_t* x, int i)
{
    return (int)((x[i / 64] >> (i % 64)) & 1u);
}

/* bit_array_hash, bit-array.hpp:410-423 (std::hash<size_t> is the identity) */
uint64_t
ox_bit_array_hash(int n, const uint64_t* x)
{
    uint64_t ret = 0;
    for (int i = 0, e = n / 64 + 1; i != e; ++i)
        ret ^= x[i] + 0x9e3779b9u + (ret << 6) + (ret >> 2);
    return ret;
}

/* default_cost_type::results, itm-common.hpp:1140-1147 */
static double
pop_results(const ox_problem* pb, const uint64_t* x)
{
    double v = pb->cost_constant;
    for (int i = 0; i != pb->n_obj; ++i)
        if (wget(x, pb->obj_var[i]))
            v += pb->obj_factor[i];
    return v;
}

static int
pop_remaining(const ox_problem* pb, const uint64_t* x)
{
    int remaining = 0;
    for (int k = 0; k != pb->m; ++k) {
        int sum = 0;
        for (int e = pb->row_ptr[k]; e != pb->row_ptr[k + 1]; ++e)
            if (wget(x, pb->row_col[e]))
                sum += pb->A[e];
        if (!(pb->cmin[k] <= sum && sum <= pb->cmax[k]))
            ++remaining;
    }
    return remaining;
}

static int
is_better_solution(int maximize, double lhs, double rhs)
{
    return maximize ? lhs > rhs : lhs < rhs;
}

/* storage::sort's comparator, itm-optimizer-common.hpp:425-441 */
static int
pop_index_less(const void* a, const void* b, void* ctx)
{
    const struct ox_pop* pop = (const struct ox_pop*)ctx;
    const ox_member* m1 = &pop->data[*(const int*)a];
    const ox_member* m2 = &pop->data[*(const int*)b];

    if (m1->remaining < m2->remaining)
        return 1;
    if (m1->remaining == m2->remaining)
        return is_better_solution(pop->pb->maximize, m1->value, m2->value);
    return 0;
}

static void
pop_sort(struct ox_pop* pop)
{
    ox_std_sort(pop->indices, pop->size, sizeof(int), pop_index_less, pop);
}

/* init_with_random, itm-common.hpp:269-282 */
static void
pop_init_with_random(uint64_t* x, uint32_t* rng, int n, double p)
{
    for (int i = 0; i != n; ++i) {
        if (ox_bernoulli(p, rng))
            x[i / 64] |= (uint64_t)1 << (i % 64);
        else
            x[i / 64] &= ~((uint64_t)1 << (i % 64));
    }
}

/* init_with_pre_solve with both vectors present, itm-common.hpp:284-374 */
static void
pop_pre_solve(const ox_problem* pb, uint64_t* xp, uint64_t* xo, uint32_t* rng,
              double init_random)
{
    ox_pre_rc* R =
      (ox_pre_rc*)malloc(sizeof(ox_pre_rc) * (size_t)(pb->max_row + 1));

    for (int k = 0; k != pb->m; ++k) {
        if (!ox_bernoulli(init_random, rng))
            continue;

        const int r_size = pb->row_ptr[k + 1] - pb->row_ptr[k];
        for (int i = 0; i != r_size; ++i) {
            const int e = pb->row_ptr[k] + i;
            R[i].value = (float)pb->cost[pb->row_col[e]];
            R[i].factor = pb->A[e];
            R[i].id = pb->row_col[e];
        }

        ox_std_shuffle(R, r_size, sizeof(ox_pre_rc), rng);
        ox_std_sort(R, r_size, sizeof(ox_pre_rc),
                    pb->maximize ? pre_less_max : pre_less_min, NULL);

        {
            int sum = 0, best = -2;
            for (int i = -1; i < r_size; ++i) {
                if (pb->cmin[k] <= sum && sum <= pb->cmax[k]) {
                    best = i;
                    break;
                }
                if (i + 1 < r_size)
                    sum += R[i + 1].factor;
            }
            int i = 0;
            for (; i <= best; ++i)
                xp[R[i].id / 64] |= (uint64_t)1 << (R[i].id % 64);
            for (; i != r_size; ++i)
                xp[R[i].id / 64] &= ~((uint64_t)1 << (R[i].id % 64));
        }
        {
            int sum = 0, best = -2;
            for (int i = -1; i < r_size; ++i) {
                if (pb->cmin[k] <= sum && sum <= pb->cmax[k])
                    best = i;
                if (best != -2 && i + 1 < r_size &&
                    (pb->maximize ? R[i + 1].value < 0 : R[i + 1].value > 0))
                    break;
                if (i + 1 < r_size)
                    sum += R[i + 1].factor;
            }
            int i = 0;
            for (; i <= best; ++i)
                xo[R[i].id / 64] |= (uint64_t)1 << (R[i].id % 64);
            for (; i != r_size; ++i)
                xo[R[i].id / 64] &= ~((uint64_t)1 << (R[i].id % 64));
        }
    }
    free(R);
}

/* best_solution_recorder / storage constructors,
 * itm-optimizer-common.hpp:162-226,467-480, on random_engine(master_seed). */
ox_pop*
ox_pop_create(ox_problem* pb, const ox_params* p, uint64_t master_seed,
              int threads, uint32_t* rng_state)
{
    struct ox_pop* pop = (struct ox_pop*)calloc(1, sizeof(struct ox_pop));
    uint32_t rng = engine_seed(master_seed);
    const int n = pb->n;

    pop->pb = pb;
    pop->p = *p;
    pop->size = p->init_population_size;
    pop->blocks = n / 64 + 1;
    pop->threads = threads;
    pop->indices = (int*)calloc((size_t)pop->size, sizeof(int));
    pop->data = (ox_member*)calloc((size_t)pop->size, sizeof(ox_member));
    pop->bastert = (uint64_t*)calloc((size_t)pop->blocks, sizeof(uint64_t));
    pop->random = (uint64_t*)calloc((size_t)pop->blocks, sizeof(uint64_t));
    pop->kappa_append = (double*)calloc((size_t)(threads > 0 ? threads : 1), sizeof(double));
    for (int i = 0; i != pop->size; ++i)
        pop->data[i].x = (uint64_t*)calloc((size_t)pop->blocks, sizeof(uint64_t));

    for (int i = 0; i != n; ++i)
        if (init_x(pb->maximize, pb->cost[i], 0))
            pop->bastert[i / 64] |= (uint64_t)1 << (i % 64);

    for (int i = 0, e = pop->size / 2; i != e; ++i) {
        memcpy(pop->data[i].x, pop->bastert, sizeof(uint64_t) * (size_t)pop->blocks);
        const double pr = clampd((double)i / (5 * e), 0.0, 1.0);
        for (int v = 0; v != n; ++v)
            if (ox_bernoulli(pr, &rng))
                pop->data[i].x[v / 64] ^= (uint64_t)1 << (v % 64);
    }

    for (int i = pop->size / 2, e = pop->size; i + 1 < e; i += 2) {
        pop_init_with_random(pop->data[i].x, &rng, n, 0.2);
        pop_init_with_random(pop->data[i + 1].x, &rng, n, 0.8);
        pop_pre_solve(pb, pop->data[i].x, pop->data[i + 1].x, &rng,
                      clampd((double)i / (5 * e), 0.0, 1.0));
    }

    for (int i = 0; i != pop->size; ++i) {
        ox_member* mb = &pop->data[i];
        mb->hash = ox_bit_array_hash(n, mb->x);
        mb->value = pop_results(pb, mb->x);
        mb->remaining = pop_remaining(pb, mb->x);
        mb->loop = 0;
        pop->indices[i] = i;
    }
    pop_sort(pop);

    if (rng_state)
        *rng_state = rng;
    return pop;
}

void
ox_pop_destroy(ox_pop* pop)
{
    if (!pop)
        return;
    for (int i = 0; i != pop->size; ++i)
        free(pop->data[i].x);
    free(pop->indices);
    free(pop->data);
    free(pop->bastert);
    free(pop->random);
    free(pop->kappa_append);
    free(pop);
}

int
ox_pop_size(ox_pop* pop)
{
    return pop->size;
}

void
ox_pop_member(ox_pop* pop, int i, uint64_t* x_words, double* value,
              int* remaining, uint64_t* hash, int64_t* loop)
{
    const ox_member* mb = &pop->data[pop->indices[i]];
    if (x_words)
        memcpy(x_words, mb->x, sizeof(uint64_t) * (size_t)pop->blocks);
    if (value)
        *value = mb->value;
    if (remaining)
        *remaining = mb->remaining;
    if (hash)
        *hash = mb->hash;
    if (loop)
        *loop = mb->loop;
}

ox_pop_ctx*
ox_pop_ctx_create(ox_pop* pop, unsigned thread_id, uint64_t seed)
{
    struct ox_pop_ctx* c = (struct ox_pop_ctx*)calloc(1, sizeof(struct ox_pop_ctx));
    const ox_params* p = &pop->p;
    c->rng = engine_seed(seed);
    c->choose_sol = (ox_normal_dist){ p->init_crossover_solution_selection_mean,
                                      p->init_crossover_solution_selection_stddev, 0.0, 0 };
    c->variable_p = (ox_normal_dist){ p->init_mutation_variable_mean,
                                      p->init_mutation_variable_stddev, 0.0, 0 };
    c->value_p = (ox_normal_dist){ p->init_mutation_value_mean,
                                   p->init_mutation_value_stddev, 0.0, 0 };
    c->bad_lo = p->init_population_size / 5;
    c->bad_hi = p->init_population_size - 1;
    c->bastert_insertion = p->init_crossover_bastert_insertion;
    c->improve_start = p->init_kappa_improve_start;
    c->improve_increase = p->init_kappa_improve_increase;
    c->improve_stop = p->init_kappa_improve_stop;
    c->thread_id = thread_id;
    return c;
}

void
ox_pop_ctx_destroy(ox_pop_ctx* c)
{
    free(c);
}

uint32_t
ox_pop_ctx_rng(ox_pop_ctx* c)
{
    return c->rng;
}

void
ox_pop_ctx_set_rng(ox_pop_ctx* c, uint32_t state)
{
    c->rng = state;
}

/* storage::choose_a_solution, itm-optimizer-common.hpp:151-160 */
static int
choose_a_solution(struct ox_pop* pop, struct ox_pop_ctx* c)
{
    double value;
    do
        value = ox_normal_draw(&c->choose_sol, &c->rng);
    while (value < 0 || value > 1);
    return (int)(pop->size * value);
}

/* storage::crossover(ctx, x, first, second), :359-371 */
static void
crossover_blocks(struct ox_pop* pop, struct ox_pop_ctx* c, uint64_t* x,
                 const uint64_t* first, const uint64_t* second)
{
    for (int i = 0; i != pop->blocks; ++i) {
        const uint64_t x1 = first[i], x2 = second[i];
        x[i] = ((x1 ^ x2) & ox_uniform_u64(0, UINT64_MAX, &c->rng)) ^ x1;
    }
}

/* storage::crossover(ctx, x), :373-421 */
void
ox_pop_crossover(ox_pop* pop, ox_pop_ctx* c, uint64_t* x)
{
    const size_t bytes = sizeof(uint64_t) * (size_t)pop->blocks;

    if (ox_bernoulli(c->bastert_insertion, &c->rng)) {
        const int first = pop->indices[choose_a_solution(pop, c)];
        if (ox_bernoulli(0.5, &c->rng)) {
            if (ox_bernoulli(0.5, &c->rng))
                crossover_blocks(pop, c, x, pop->data[first].x, pop->bastert);
            else
                memcpy(pop->data[first].x, pop->bastert, bytes);
        } else {
            pop_init_with_random(pop->random, &c->rng, pop->pb->n, 0.5);
            if (ox_bernoulli(0.5, &c->rng))
                crossover_blocks(pop, c, x, pop->data[first].x, pop->random);
            else
                memcpy(pop->data[first].x, pop->random, bytes);
        }
    } else {
        const int first = pop->indices[choose_a_solution(pop, c)];
        int second = pop->indices[choose_a_solution(pop, c)];
        while (first == second)
            second = pop->indices[choose_a_solution(pop, c)];
        crossover_blocks(pop, c, x, pop->data[first].x, pop->data[second].x);
    }
}

/* best_solution_recorder::mutation, :494-526.  bit_array::set(i, bool)
 * (bit-array.hpp:143-148) writes `val` only into bit 0 of a block and clears
 * any other bit, whatever val is: reproduced. */
void
ox_pop_mutation(ox_pop* pop, ox_pop_ctx* c, uint64_t* x)
{
    if (c->value_p.mean == 0.0 && c->value_p.stddev == 0.0)
        return;

    double val_p, var_p;
    do
        var_p = ox_normal_draw(&c->variable_p, &c->rng);
    while (var_p <= 0.0 || var_p >= 1.0);
    do
        val_p = ox_normal_draw(&c->value_p, &c->rng);
    while (val_p < 0.0 || val_p > 1.0);

    for (int i = 0, e = pop->pb->n; i != e; ++i) {
        if (ox_bernoulli(var_p, &c->rng)) {
            const uint64_t val = (uint64_t)ox_bernoulli(val_p, &c->rng);
            const uint64_t mask = (uint64_t)1 << (i % 64);
            uint64_t* ref = &x[i / 64];
            *ref ^= (val ^ *ref) & mask;
        }
    }
}

/* best_solution_recorder::reinit, :528-554 */
double
ox_pop_reinit(ox_pop* pop, ox_pop_ctx* c, double kappa_min, double kappa_max,
              uint64_t* x)
{
    double kappa = kappa_min;
    double* ka = &pop->kappa_append[c->thread_id];

    if (*ka < c->improve_stop) {
        *ka += c->improve_increase;
        kappa = kappa_min + (kappa_max - kappa_min) * *ka;
    } else {
        *ka = c->improve_start;
        ox_pop_crossover(pop, c, x);
    }
    ox_pop_mutation(pop, c, x);
    return kappa;
}

/* storage::insert's common part: choose_a_bad_solution, replace_result,
 * sort (:228-300). */
static void
pop_replace(struct ox_pop* pop, struct ox_pop_ctx* c, const uint64_t* x,
            double value, uint64_t hash, int64_t loop, int remaining)
{
    const int pos = (int)ox_uniform_u64((uint64_t)c->bad_lo, (uint64_t)c->bad_hi, &c->rng);
    ox_member* mb = &pop->data[pop->indices[pos]];
    memcpy(mb->x, x, sizeof(uint64_t) * (size_t)pop->blocks);
    mb->value = value;
    mb->hash = hash;
    mb->loop = loop;
    mb->remaining = remaining;
    pop_sort(pop);
}

/* best_solution_recorder::try_update, :571-584 */
void
ox_pop_try_update(ox_pop* pop, ox_pop_ctx* c, const uint64_t* x, double value,
                  int64_t loop)
{
    const uint64_t hash = ox_bit_array_hash(pop->pb->n, x);
    for (int i = 0; i != pop->size; ++i)
        if (pop->data[i].remaining == 0 && pop->data[i].value == value &&
            pop->data[i].hash == hash)
            return;
    pop_replace(pop, c, x, value, hash, loop, 0);
}

/* best_solution_recorder::try_advance, :556-569 */
void
ox_pop_try_advance(ox_pop* pop, ox_pop_ctx* c, const uint64_t* x, int remaining,
                   int64_t loop)
{
    const uint64_t hash = ox_bit_array_hash(pop->pb->n, x);
    for (int i = 0; i != pop->size; ++i)
        if (pop->data[i].remaining == remaining && pop->data[i].hash == hash)
            return;
    pop_replace(pop, c, x, pop_results(pop->pb, x), hash, loop, remaining);
}

/* ======================================================================== */
/* optimize_functor::operator() at thread = 1,                              */
/* itm-optimizer-common.hpp:596-752 (set up like optimize_problem, :795-826) */
/* ======================================================================== */

/* generate_seed, itm-common.hpp:1457-1484, continuing the master engine */
static void
ox_generate_seed_from_state(uint32_t* master, unsigned count, uint64_t* out)
{
    if (count == 0)
        return;
    out[0] = ox_uniform_u64(0, UINT64_MAX, master);
    for (unsigned i = 1; i != count; ++i) {
        out[i] = ox_uniform_u64(0, UINT64_MAX, master);
        unsigned j = i - 1;
        while (j > 0) {
            if (out[j] == out[i]) {
                out[j] = ox_uniform_u64(0, UINT64_MAX, master);
                j = i;
            }
            --j;
        }
    }
}

struct ox_opt_trace
{
    int rounds; /* snapshots taken (restarts + 1) */
    int m, nnz, blocks, size;
    uint64_t* x;    /* [rounds][blocks] */
    uint32_t* rng;  /* [rounds] */
    double* kappa;  /* [rounds] */
    double* pi;     /* [rounds][m] */
    double* P;      /* [rounds][nnz] */
    uint64_t* mx;   /* [rounds][size][blocks] */
    double* mvalue; /* [rounds][size] */
    int* mremaining;
    uint64_t* mhash;
    int64_t* mloop;
    int sweeps, sweep_cap;
    int* sweep_round;
    int* sweep_push;
    int* sweep_len;
    double* sweep_kappa;
    int cur_round;
};

static void
trace_sweep(struct ox_opt_trace* t, const ox_run* r, int push, double kappa)
{
    if (t->sweeps == t->sweep_cap) {
        t->sweep_cap = t->sweep_cap ? 2 * t->sweep_cap : 1024;
        t->sweep_round = (int*)realloc(t->sweep_round, sizeof(int) * (size_t)t->sweep_cap);
        t->sweep_push = (int*)realloc(t->sweep_push, sizeof(int) * (size_t)t->sweep_cap);
        t->sweep_len = (int*)realloc(t->sweep_len, sizeof(int) * (size_t)t->sweep_cap);
        t->sweep_kappa = (double*)realloc(t->sweep_kappa, sizeof(double) * (size_t)t->sweep_cap);
    }
    const int pairs =
      r->order == OX_ORDER_INFEASIBILITY_DECR || r->order == OX_ORDER_INFEASIBILITY_INCR;
    t->sweep_round[t->sweeps] = t->cur_round;
    t->sweep_push[t->sweeps] = push;
    t->sweep_len[t->sweeps] = pairs ? r->m_order_len : r->oR_len;
    t->sweep_kappa[t->sweeps] = kappa;
    t->sweeps++;
}

static void
trace_snapshot(struct ox_opt_trace* t, int slot, const ox_run* r, const struct ox_pop* pop,
               double kappa)
{
    const ox_problem* pb = r->pb;
    memcpy(t->x + (size_t)slot * t->blocks, r->x, sizeof(uint64_t) * (size_t)t->blocks);
    t->rng[slot] = r->rng;
    t->kappa[slot] = kappa;
    memcpy(t->pi + (size_t)slot * t->m, r->pi, sizeof(double) * (size_t)t->m);
    memcpy(t->P + (size_t)slot * t->nnz, r->P, sizeof(double) * (size_t)t->nnz);
    for (int i = 0; i != t->size; ++i) {
        const ox_member* mb = &pop->data[pop->indices[i]];
        const size_t o = (size_t)slot * t->size + i;
        memcpy(t->mx + o * t->blocks, mb->x, sizeof(uint64_t) * (size_t)t->blocks);
        t->mvalue[o] = mb->value;
        t->mremaining[o] = mb->remaining;
        t->mhash[o] = mb->hash;
        t->mloop[o] = mb->loop;
    }
    (void)pb;
}

/* One sweep of the traced worker: the engine is the worker's (local_context::rng
 * is the engine the Solver holds by reference). */
static int
opt_sweep(ox_run* r, struct ox_opt_trace* t, double kappa, int push, double amp)
{
    trace_sweep(t, r, push, kappa);
    return order_run(r, kappa, r->delta, r->theta, push, amp);
}

static void
opt_try_update(ox_run* r, struct ox_pop* pop, struct ox_pop_ctx* c, int64_t loop)
{
    c->rng = r->rng;
    ox_pop_try_update(pop, c, r->x, results(r), loop);
    r->rng = c->rng;
}

ox_opt_trace*
ox_opt_trace_run(ox_problem* pb, const ox_params* p, uint32_t seed, int rounds)
{
    /* optimize_problem: random_engine rng(seed); best_solution_recorder(rng, ...);
     * seeds = generate_seed(rng, thread); functor(ctx, 0, seeds[0]) */
    uint32_t master = 0;
    struct ox_pop* pop = ox_pop_create(pb, p, seed, 1, &master);
    uint64_t wseed = 0;
    ox_generate_seed_from_state(&master, 1, &wseed);
    struct ox_pop_ctx* c = ox_pop_ctx_create(pop, 0, wseed);

    /* the functor's Solver, x = 0, compute_order(order, n) */
    ox_params q = *p;
    q.init_policy_random = 0.0;
    ox_run* r = ox_run_begin(pb, &q, 1);
    if (!r) {
        ox_pop_ctx_destroy(c);
        ox_pop_destroy(pop);
        return NULL;
    }
    r->p = *p;
    memset(r->x, 0, sizeof(uint64_t) * (size_t)r->x_blocks);
    r->rng = c->rng;
    {
        const int n = pb->n;
        r->oR_len = n;
        r->m_order_len = n;
        for (int i = 0; i != n; ++i) {
            r->oR[i] = 0;
            r->m_order[i].first = 0;
            r->m_order[i].second = 0;
        }
    }

    struct ox_opt_trace* t = (struct ox_opt_trace*)calloc(1, sizeof(struct ox_opt_trace));
    t->rounds = rounds + 1;
    t->m = pb->m;
    t->nnz = pb->nnz;
    t->blocks = r->x_blocks;
    t->size = pop->size;
    t->x = (uint64_t*)calloc((size_t)t->rounds * t->blocks, sizeof(uint64_t));
    t->rng = (uint32_t*)calloc((size_t)t->rounds, sizeof(uint32_t));
    t->kappa = (double*)calloc((size_t)t->rounds, sizeof(double));
    t->pi = (double*)calloc((size_t)t->rounds * t->m, sizeof(double));
    t->P = (double*)calloc((size_t)t->rounds * t->nnz + 1, sizeof(double));
    t->mx = (uint64_t*)calloc((size_t)t->rounds * t->size * t->blocks, sizeof(uint64_t));
    t->mvalue = (double*)calloc((size_t)t->rounds * t->size, sizeof(double));
    t->mremaining = (int*)calloc((size_t)t->rounds * t->size, sizeof(int));
    t->mhash = (uint64_t*)calloc((size_t)t->rounds * t->size, sizeof(uint64_t));
    t->mloop = (int64_t*)calloc((size_t)t->rounds * t->size, sizeof(int64_t));

    const double kappa_step = p->kappa_step, kappa_max = p->kappa_max, alpha = p->alpha;
    const long w_limit = (long)p->w;

    for (int round = 0; round <= rounds; ++round) {
        t->cur_round = round;
        c->rng = r->rng;
        const double kappa_start = ox_pop_reinit(pop, c, p->kappa_min, p->kappa_max, r->x);
        r->rng = c->rng;
        double kappa = kappa_start;
        order_init(r);
        trace_snapshot(t, round, r, pop, kappa);
        if (round == rounds)
            break;

        int best_remaining = INT_MAX;
        for (int64_t i = 0; i != p->limit; ++i) {
            const int remaining = opt_sweep(r, t, kappa, 0, 0.0);
            if (remaining == 0) {
                opt_try_update(r, pop, c, i);
                best_remaining = 0;
                break;
            } else if (remaining < best_remaining)
                best_remaining = remaining;

            if (i > w_limit)
                kappa += kappa_step * pow((double)remaining / (double)pb->m, alpha);
            if (kappa > kappa_max)
                break;
        }

        if (best_remaining > 0) {
            c->rng = r->rng;
            ox_pop_try_advance(pop, c, r->x, best_remaining, p->limit);
            r->rng = c->rng;
            continue;
        }

        for (int push = 0; push < p->pushes_limit; ++push) {
            int remaining =
              opt_sweep(r, t, p->pushing_k_factor, 1, p->pushing_objective_amplifier);
            if (remaining == 0)
                opt_try_update(r, pop, c, -(int64_t)push * p->pushing_iteration_limit - 1);

            kappa = kappa_start;
            for (int iter = 0; iter < p->pushing_iteration_limit; ++iter) {
                remaining = opt_sweep(r, t, kappa, 0, 0.0);
                if (remaining == 0) {
                    opt_try_update(r, pop, c,
                                   -(int64_t)push * p->pushing_iteration_limit - iter - 1);
                    break;
                }
                if (iter > p->w)
                    kappa += kappa_step * pow((double)remaining / (double)pb->m, alpha);
                if (kappa > kappa_max)
                    break;
            }
        }
    }

    ox_run_end(r);
    ox_pop_ctx_destroy(c);
    ox_pop_destroy(pop);
    return t;
}

void
ox_opt_trace_free(ox_opt_trace* t)
{
    if (!t)
        return;
    free(t->x);
    free(t->rng);
    free(t->kappa);
    free(t->pi);
    free(t->P);
    free(t->mx);
    free(t->mvalue);
    free(t->mremaining);
    free(t->mhash);
    free(t->mloop);
    free(t->sweep_round);
    free(t->sweep_push);
    free(t->sweep_len);
    free(t->sweep_kappa);
    free(t);
}

int
ox_opt_trace_rounds(ox_opt_trace* t)
{
    return t->rounds;
}

int
ox_opt_trace_round(ox_opt_trace* t, int r, uint64_t* x_words, uint32_t* rng_state, double* kappa,
                   double* pi, double* P)
{
    if (r < 0 || r >= t->rounds)
        return -1;
    if (x_words)
        memcpy(x_words, t->x + (size_t)r * t->blocks, sizeof(uint64_t) * (size_t)t->blocks);
    if (rng_state)
        *rng_state = t->rng[r];
    if (kappa)
        *kappa = t->kappa[r];
    if (pi)
        memcpy(pi, t->pi + (size_t)r * t->m, sizeof(double) * (size_t)t->m);
    if (P)
        memcpy(P, t->P + (size_t)r * t->nnz, sizeof(double) * (size_t)t->nnz);
    return 0;
}

int
ox_opt_trace_member(ox_opt_trace* t, int r, int i, uint64_t* x_words, double* value,
                    int* remaining, uint64_t* hash, int64_t* loop)
{
    if (r < 0 || r >= t->rounds || i < 0 || i >= t->size)
        return -1;
    const size_t o = (size_t)r * t->size + i;
    if (x_words)
        memcpy(x_words, t->mx + o * t->blocks, sizeof(uint64_t) * (size_t)t->blocks);
    if (value)
        *value = t->mvalue[o];
    if (remaining)
        *remaining = t->mremaining[o];
    if (hash)
        *hash = t->mhash[o];
    if (loop)
        *loop = t->mloop[o];
    return 0;
}

int
ox_opt_trace_sweeps(ox_opt_trace* t)
{
    return t->sweeps;
}

int
ox_opt_trace_sweep_log(ox_opt_trace* t, int* round, int* push, double* kappa, int* len)
{
    memcpy(round, t->sweep_round, sizeof(int) * (size_t)t->sweeps);
    memcpy(push, t->sweep_push, sizeof(int) * (size_t)t->sweeps);
    memcpy(kappa, t->sweep_kappa, sizeof(double) * (size_t)t->sweeps);
    memcpy(len, t->sweep_len, sizeof(int) * (size_t)t->sweeps);
    return 0;
}
