/* bxg.h -- C ABI of the B200 constraint-sweep engine (libbaryonyx_b200.so).
 *
 * The reference (quesnel/baryonyx v0.5.0) has no FFI: its hot path sits behind
 * a compile-time "Solver" concept.  This header is the boundary
 * a maintainer binds instead: plain C, opaque handles, caller-owned host
 * buffers, int status returns (0 = ok, negative = bxg_status), no exceptions
 * across the boundary.  One handle is used by one host thread at a time; each
 * batch owns one CUDA stream.  Every entry point cites the reference interface
 * it replaces (paths relative to the reference root).
 *
 * There is no CPU fallback: every compute call fails with BXG_E_CUDA when no
 * CUDA device is usable.
 */
#ifndef BXG_H
#define BXG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BXG_VERSION 2

typedef enum bxg_status
{
    BXG_OK = 0,
    BXG_E_INVALID = -1,      /* bad argument */
    BXG_E_CUDA = -2,         /* CUDA runtime error / no device */
    BXG_E_NO_SOLVER = -3,    /* solver_error_tag::no_solver_available, core:258-263 */
    BXG_E_UNREALISABLE = -4, /* solver_error_tag::unrealisable_constraint */
    BXG_E_MEMORY = -5,       /* solver_error_tag::not_enough_memory */
    BXG_E_UNSUPPORTED = -6   /* feature of the reference outside this engine */
} bxg_status;

/* baryonyx::problem_solver_type, lib/include/baryonyx/core:542-550 */
enum
{
    BXG_EQUALITIES_01 = 0,
    BXG_EQUALITIES_101 = 1,
    BXG_EQUALITIES_Z = 2,
    BXG_INEQUALITIES_01 = 3,
    BXG_INEQUALITIES_101 = 4,
    BXG_INEQUALITIES_Z = 5
};

/* baryonyx::result_status, core:692-701 */
enum
{
    BXG_STATUS_SUCCESS = 0,
    BXG_STATUS_INTERNAL_ERROR = 1,
    BXG_STATUS_UNINITIALIZED = 2,
    BXG_STATUS_TIME_LIMIT_REACHED = 3,
    BXG_STATUS_KAPPA_MAX_REACHED = 4,
    BXG_STATUS_LIMIT_REACHED = 5,
    BXG_STATUS_EMPTY_CONTEXT = 6
};

/* POD mirror of baryonyx::solver_parameters, core:294-462 (same field order;
 * enum members carry the reference's integer values). */
typedef struct bxg_params
{
    double time_limit;
    double theta;
    double delta;
    double kappa_min;
    double kappa_step;
    double kappa_max;
    double alpha;
    double pushing_k_factor;
    double pushing_objective_amplifier;
    double init_policy_random;
    double init_crossover_bastert_insertion;
    double init_crossover_solution_selection_mean;
    double init_crossover_solution_selection_stddev;
    double init_mutation_variable_mean;
    double init_mutation_variable_stddev;
    double init_mutation_value_mean;
    double init_mutation_value_stddev;
    double init_kappa_improve_start;
    double init_kappa_improve_increase;
    double init_kappa_improve_stop;
    double w;
    uint32_t seed;
    int32_t thread;
    int64_t limit;
    int32_t print_level;
    int32_t pushes_limit;
    int32_t pushing_iteration_limit;
    int32_t init_population_size;
    int32_t pre_order;
    int32_t order;
    int32_t float_type;
    int32_t init_policy;
    int32_t cost_norm;
    int32_t mode;
    int32_t preprocessor;
    int32_t observer;
    int32_t storage;
    int32_t solver;
    int32_t debug;
} bxg_params;

typedef struct bxg_problem bxg_problem;
typedef struct bxg_batch bxg_batch;

int bxg_version(void);
/* Message of the last failing call on this thread. */
const char* bxg_last_error(void);
int bxg_device_count(void);

/* solver_parameters defaults, core:410-461 */
void bxg_default_params(bxg_params* p);
/* context_set_solver_parameters, lib/src/private.cpp:51-172 (validation and
 * clamping; w becomes a sweep count). */
void bxg_clamp_params(const bxg_params* in, bxg_params* out);

/* ---- problem -------------------------------------------------------------
 * Replaces the Solver constructors (itm-solver-inequalities-01.cpp:72-107,
 * -equalities-01.cpp:68-98, -inequalities-101.cpp:82-130,
 * -equalities-101.cpp:80-113, -inequalities-Z.cpp:98-189) and
 * sparse_matrix<int> (sparse-matrix.hpp:86-186).  Input = the merged
 * constraints make_merged_constraints returns (merged-constraint.cpp:297-343)
 * in CSR form: row k holds col[row_ptr[k]..row_ptr[k+1]) with integer factors
 * and cmin[k] <= row <= cmax[k] (INT_MIN / INT_MAX for open sides); the
 * objective as (variable, factor) pairs in file order plus its constant.
 * Uploads the structure to `device` and builds the CSC index there. */
int bxg_problem_create(int m, int n, const int32_t* row_ptr, const int32_t* col,
                       const int32_t* factor, const int32_t* cmin,
                       const int32_t* cmax, int problem_type, int maximize,
                       int n_obj, const int32_t* obj_var,
                       const double* obj_factor, double cost_constant,
                       int device, bxg_problem** out);
void bxg_problem_destroy(bxg_problem* pb);
int bxg_problem_info(const bxg_problem* pb, int* m, int* n, int* nnz,
                     int* max_row, int* problem_type);
/* problem::which_problem_type, lib/src/problem.hpp:72-122 */
int bxg_which_problem_type(int nnz, const int32_t* factor, int has_inequality);

/* ---- batch of lanes --------------------------------------------------------
 * A lane is one independent Solver instance + its x, random engine and
 * compute_order state (what one worker of the reference owns,
 * itm-optimizer-common.hpp:620-656).  `p` must be clamped.  Allocates
 * P[n_lanes][nnz], pi[n_lanes][m], x, R scratch on the problem's device and
 * computes the normalised costs and delta once (normalize_costs / compute_delta,
 * itm-common.hpp:917-998). */
int bxg_batch_create(bxg_problem* pb, int n_lanes, const bxg_params* p,
                     bxg_batch** out);
void bxg_batch_destroy(bxg_batch* b);
int bxg_batch_info(const bxg_batch* b, int* n_lanes, double* delta,
                   double* theta, int64_t* device_bytes);
/* Value modelled in the slot one past the reduced-cost scratch, which the
 * reference reads for its longest row (undefined behaviour there;
 * itm-solver-inequalities-01.cpp:198-203).  Default 0.0. */
int bxg_batch_set_tail_value(bxg_batch* b, double v);

/* Start a run on every lane: random_engine(seed), P = pi = 0, x from the init
 * policy followed by the Bernoulli(init_policy_random) flips, then
 * compute_order::init (solver_functor::operator(), itm-solver-common.hpp:63-133).
 * seeds has n_lanes entries.  reset_solver = 0 keeps P and pi (a worker of the
 * optimizer never resets them between restarts). */
int bxg_batch_begin(bxg_batch* b, const uint32_t* seeds, int reset_solver);

/* Parity taps on one lane (any pointer may be NULL).  x_words: n/64+1 words
 * (bit_array layout, bit-array.hpp:60-66); P is nnz values in the reference's
 * row-major order; order_list = compute_order::R, or m_order as (k, amount)
 * pairs for the infeasibility orders. */
int bxg_batch_get_state(bxg_batch* b, int lane, uint64_t* x_words, double* pi,
                        double* P, uint32_t* rng_state, int32_t* order_list,
                        int32_t* order_len, double* norm_cost);
/* set_x also re-runs compute_order::init for the lane. */
int bxg_batch_set_state(bxg_batch* b, int lane, const uint64_t* x_words,
                        const double* pi, const double* P,
                        const uint32_t* rng_state);

/* One sweep on every lane = compute_order::run (push = 0) or push_and_run
 * (push = 1), itm-common.hpp:682-880, i.e. compute_update_row /
 * push_and_compute_update_row over the lane's order list followed by the
 * violated-row scan.  kappa: n_lanes host values; remaining: n_lanes host
 * outputs (may be NULL). */
int bxg_batch_sweep(bxg_batch* b, int push, const double* kappa,
                    double objective_amplifier, int32_t* remaining);

/* default_cost_type::results (itm-common.hpp:1140-1147) of the lanes' current x. */
int bxg_batch_objective(bxg_batch* b, double* value);

/* Whole run of solver_functor::operator() (itm-solver-common.hpp:135-225) on
 * every lane, on the device: sweeps with the kappa schedule, then the push
 * phase, tracking the best x.  Requires bxg_batch_begin.  Host outputs, each
 * n_lanes long, any may be NULL: status (result_status), value, remaining
 * constraints of the best, loop index of the best. */
int bxg_batch_solve(bxg_batch* b, int32_t* status, double* value,
                    int32_t* remaining, int64_t* loop);
/* Best x of a lane as bytes (n entries 0/1), like solution::variables. */
int bxg_batch_best(bxg_batch* b, int lane, int8_t* x);
/* Index of the best lane (feasible first, then objective), with its value. */
int bxg_batch_argbest(bxg_batch* b, int* lane, double* value, int* remaining);

/* Work counters since batch creation: constraint updates (order-list entries
 * processed), sweeps, and algorithmic bytes (SURVEY.md 8d formula). */
int bxg_batch_counters(bxg_batch* b, int64_t* updates, int64_t* sweeps,
                       double* algorithmic_bytes);
int bxg_batch_reset_counters(bxg_batch* b);
/* Updates by selection path of calculator_sort / select_variables
 * (itm-common.hpp:134-148): [0] exact single-thread ordering, [1] CTA-wide
 * sort, [2] two smallest values, [3] boundary at the sign change; and the SM
 * cycles spent inside them (one thread's clock); phase_cycles splits the
 * cycles of paths [2] and [3] into reduced-cost gather, column sums, tie
 * detection + boundary values, and affect.  Each output holds 4 values, any
 * may be NULL. */
int bxg_batch_path_counters(bxg_batch* b, int64_t* updates, int64_t* cycles,
                            int64_t* phase_cycles);
/* Parameter sweeps: per-lane values of the parameters the reference's
 * auto-tuners vary from run to run (manual-optimizer.cpp:130-160: theta, delta,
 * kappa-min, kappa-step, init-policy-random; nlopt-optimizer.cpp:101-103),
 * so that one batch runs the whole grid at once.  Arrays of n_lanes entries;
 * a NULL array keeps the batch's value for every lane, all NULL removes the
 * per-lane values.  Every lane's set goes through the clamps of
 * context_set_solver_parameters (private.cpp:51-172); delta < 0 = automatic
 * (compute_delta with the lane's theta, itm-common.hpp:917-933).  Takes
 * effect at the next bxg_batch_begin / bxg_batch_sweep. */
int bxg_batch_set_lane_params(bxg_batch* b, const double* theta,
                              const double* delta, const double* kappa_min,
                              const double* kappa_step,
                              const double* init_policy_random);
/* The grid of the reference's manual auto-tuner (manual-optimizer.cpp:31-111,
 * 113-160): five tables of five values -- entry 0 the caller's own parameter,
 * entry 1 a fixed fraction of 1/5, the others built on theta's table as the
 * reference builds them (:62-87, as written there) -- and the 5^5 = 3125
 * points in the order its odometer visits them (`next()`, :98-111: the last
 * dimension, init_policy_random, runs fastest).  Output: the five per-lane
 * arrays bxg_batch_set_lane_params takes, `capacity` entries each (any may be
 * NULL); points [first, first + capacity) of the walk are written, *count
 * receives how many (0 past the end), so a grid larger than a batch is taken
 * in slices.  `tables`, if not NULL, receives the 5 x 5 value tables (theta,
 * delta, kappa_min, kappa_step, init_policy_random).  Host only: no device
 * is touched. */
int bxg_manual_grid(const bxg_params* base, int64_t first, int64_t capacity,
                    double* theta, double* delta, double* kappa_min,
                    double* kappa_step, double* init_policy_random,
                    int64_t* count, double* tables);
/* Device time of the last bxg_batch_sweep / bxg_batch_solve, CUDA events on
 * the batch's stream, milliseconds. */
int bxg_batch_last_ms(bxg_batch* b, float* ms);
/* The same summed over all timed launches since the last
 * bxg_batch_reset_counters (sweeps, solves, the run half of optimizer rounds). */
int bxg_batch_device_ms(bxg_batch* b, double* total_ms);
/* Launches issued by this batch since creation (kernels of this library). */
int bxg_batch_launches(bxg_batch* b, int64_t* launches);

/* Device-resident variant of begin+solve for benchmarking / pipelines: x0 per
 * lane already in device memory (uint32 words, bxg_batch_words32 per lane). */
int bxg_batch_words32(const bxg_batch* b);
int bxg_batch_begin_device(bxg_batch* b, const uint32_t* d_x0,
                           const uint32_t* d_rng, int reset_solver);

/* ---- optimizer population --------------------------------------------------
 * Replaces best_solution_recorder / storage / local_context
 * (lib/src/itm-optimizer-common.hpp:52-594): `init_population_size` candidate
 * vectors ordered by (remaining constraints, objective) and, per worker, the
 * engine and distributions that pick parents, mutation rates and the member
 * to replace.  The lanes of a batch are the workers.  The streams of the
 * reference's std::default_random_engine and distributions are reproduced, so
 * every operation returns what the reference returns for the same seeds.
 * Vectors are bit_array blocks: n/64+1 64-bit words (bit-array.hpp:60-66). */
typedef struct bxg_population bxg_population;

/* best_solution_recorder's constructor (:467-480) on random_engine(master_seed):
 * the initial members (storage's constructor, :162-226), scored on the device. */
int bxg_population_create(bxg_problem* pb, const bxg_params* p,
                          uint64_t master_seed, int workers,
                          bxg_population** out);
void bxg_population_destroy(bxg_population* pop);
int bxg_population_size(const bxg_population* pop);
/* state of the master engine (after the constructor) */
int bxg_population_master_state(bxg_population* pop, uint32_t* state);
/* generate_seed (lib/src/itm-common.hpp:1457-1484) continuing the master engine,
 * as optimize_problem does right after the constructor (:817): the seeds of
 * `count` workers (bxg_population_worker_seed takes them as they are) */
int bxg_population_generate_seeds(bxg_population* pop, int count, uint64_t* seeds);
/* storage::get_best(i), :347-350: member at sorted position i */
int bxg_population_member(bxg_population* pop, int i, uint64_t* x_words,
                          double* value, int* remaining, uint64_t* hash,
                          int64_t* loop);
/* local_context(ctx, worker, seed), :67-91 */
int bxg_population_worker_seed(bxg_population* pop, int worker, uint64_t seed);
/* read and / or replace the state of a worker's engine (either may be NULL) */
int bxg_population_worker_state(bxg_population* pop, int worker,
                                uint32_t* get_state, const uint32_t* set_state);
/* best_solution_recorder::reinit, :528-554: x in / out, returns the kappa the
 * restart begins with */
int bxg_population_reinit(bxg_population* pop, int worker, double kappa_min,
                          double kappa_max, uint64_t* x_words, double* kappa);
/* storage::crossover(ctx, x), :373-421 */
int bxg_population_crossover(bxg_population* pop, int worker, uint64_t* x_words);
/* best_solution_recorder::mutation, :494-526 */
int bxg_population_mutation(bxg_population* pop, int worker, uint64_t* x_words);
/* best_solution_recorder::try_update, :571-584 */
int bxg_population_try_update(bxg_population* pop, int worker,
                              const uint64_t* x_words, double value,
                              int64_t loop);
/* best_solution_recorder::try_advance, :556-569 */
int bxg_population_try_advance(bxg_population* pop, int worker,
                               const uint64_t* x_words, int remaining,
                               int64_t loop);
/* remaining constraints (raw merged bounds) and default_cost_type::results of
 * `count` vectors, on the device */
int bxg_population_score(bxg_population* pop, const uint64_t* x_words, int count,
                         int32_t* remaining, double* value);
/* One call of optimize_functor::operator()'s loop body (:658-750) for every
 * lane, the lanes being the workers, on the device:
 *   reinit (:528-554)       the scalar draws (branch, parents, crossover masks,
 *                           mutation rates: std::normal_distribution needs
 *                           glibc's log) on the host from the engine state the
 *                           lane ended its last call with; the crossover blocks
 *                           and the per-variable mutation loop on the device
 *                           (the Bernoulli stream consumed in parallel with the
 *                           engine's jump-ahead), x never leaves the device;
 *   compute_order::init, the sweeps with the kappa schedule, the push phase;
 *   try_update (:571-584)   for EVERY solution met, where the reference offers
 *                           it, inside the kernel: hash, can_be_inserted, the
 *                           victim drawn from the lane's engine in the middle
 *                           of the run, replace_result, sort -- the population
 *                           lives in device memory, an insert holds its lock;
 *   try_advance (:556-569)  with the run's last x when no solution was met.
 * With one lane this is the reference's optimizer at thread = 1, restart by
 * restart (tests/golden/optimizer_*.npz are traces of the compiled reference).
 * With several lanes the inserts interleave in lock order, as the reference's
 * threads do in mutex order.  The Solver state (P, pi) and x persist between
 * calls; first_round = 1 starts from x = 0, P = pi = 0 and the workers' seeds.
 * A finite time_limit of the batch's parameters stops the lanes at a sweep
 * boundary (stop_task).  Host outputs of n_lanes entries, any may be NULL:
 * status (0 a solution was met, 5 none, 3 stopped by the deadline), the value
 * of the last solution, the smallest number of violated rows.
 * bxg_optimize_round = bxg_optimize_reinit + bxg_optimize_run; the two halves
 * are exposed as parity taps (state after reinit = state at the first sweep). */
int bxg_optimize_round(bxg_batch* b, bxg_population* pop, int first_round,
                       int32_t* status, double* value, int32_t* remaining);
int bxg_optimize_reinit(bxg_batch* b, bxg_population* pop, int first_round,
                        double* kappa);
int bxg_optimize_run(bxg_batch* b, bxg_population* pop, int32_t* status,
                     double* value, int32_t* remaining);
/* optimize_problem's driver loop (:828-857) over a batch: rounds of
 * bxg_optimize_round until `time_limit` seconds have passed (<= 0: no limit) or
 * `max_rounds` rounds are done (<= 0: no limit; one of the two is required).
 * The deadline also reaches the kernels (stop_task).  `update`, when given, is
 * called about once a second with what the reference hands ctx.update
 * (:839-853): remaining constraints, value and loop of the best member, seconds
 * since the call, restarts so far.  With a communicator the incumbent exchange
 * runs every `exchange_every` rounds (0: never). */
typedef void (*bxg_update_fn)(int remaining, double value, int64_t loop,
                              double duration, int64_t restarts, void* user);
typedef struct bxg_comm bxg_comm;
int bxg_optimize(bxg_batch* b, bxg_population* pop, bxg_comm* comm,
                 double time_limit, int64_t max_rounds, int exchange_every,
                 bxg_update_fn update, void* user, int64_t* rounds_done);
/* rounds run on this population; seconds (device clock, from the first round)
 * until a lane first offered a solution, -1 if none yet */
int bxg_population_stats(bxg_population* pop, int64_t* rounds,
                         double* first_solution_seconds);

/* ---- multi-GPU ---------------------------------------------------------------
 * The optimizer layer shards over workers: every rank owns its
 * lanes and its copy of the population, built from the same master seed.  The
 * only exchange is the incumbent: one ncclAllGather of every rank's best member
 * (scores + vector) on the batch's stream, then each rank adopts the box-wide
 * best -- fewest remaining constraints, then objective, then lowest rank -- in
 * place of its worst member unless it already holds it.  No host
 * synchronisation: the next round queues behind it.  NCCL is bound at run
 * time (libnccl.so.2).  The reference has no counterpart: its workers share one
 * address space (best_solution_recorder, :459-594). */
int bxg_comm_unique_id(void* id128); /* ncclGetUniqueId: rank 0 creates, the caller distributes */
int bxg_comm_init(const void* id128, int rank, int nranks, int device, bxg_comm** out);
void bxg_comm_destroy(bxg_comm* c);
int bxg_comm_info(const bxg_comm* c, int* rank, int* nranks, int64_t* exchanges);
/* winner_rank may be NULL (then the call does not wait for the device) */
int bxg_exchange_best(bxg_batch* b, bxg_population* pop, bxg_comm* c, int* winner_rank);

/* bit_array_hash, bit-array.hpp:410-423 */
uint64_t bxg_bit_array_hash(int n, const uint64_t* x_words);

/* ---- LP reader ---------------------------------------------------------------
 * From LP text to the arguments of bxg_problem_create: what the reference
 * obtains with make_problem (lib/src/parser.cpp:1065-1273), problem(raw_problem)
 * (lib/src/problem.hpp:58-122) and make_merged_constraints
 * (lib/src/merged-constraint.cpp:52-175,297-343).  Same dialect and the same
 * conventions: variables numbered by first appearance, coefficients and
 * right-hand sides truncated to int, rows over the same elements merged into
 * one [min, max] row.  pre_order = the reference's pre_constraint_order
 * (core:296-330): 0 none (rows in file order), 1 memory (+ elements of a row
 * sorted by variable), 2-7 rows grouped by type in the named order, 8-11
 * p1-p4 (rows and elements by cost / degree ratio); BXG_LP_RAW (-1): the parse
 * alone, no rows merged or reported, as the input of bxg_lp_preprocess (the
 * reference merges AFTER its preprocessor: rows the preprocessor removes
 * cannot make the merge fail).  Quadratic objectives are
 * BXG_E_INVALID.  Host code, no device needed.
 * On a malformed file bxg_last_error() carries the reference's
 * file_format_error_tag name. */
typedef struct bxg_lp bxg_lp;
#define BXG_LP_RAW (-1)
int bxg_lp_read(const char* path, int pre_order, bxg_lp** out);
int bxg_lp_parse(const char* text, size_t length, int pre_order, bxg_lp** out);
void bxg_lp_destroy(bxg_lp* lp);
int bxg_lp_info(const bxg_lp* lp, int* m, int* n, int* nnz, int* problem_type,
                int* maximize, double* cost_constant, int* n_obj);
/* row_ptr[m+1], col / factor[nnz], cmin / cmax[m] (INT_MIN / INT_MAX: open side) */
int bxg_lp_rows(const bxg_lp* lp, int32_t* row_ptr, int32_t* col,
                int32_t* factor, int32_t* cmin, int32_t* cmax);
/* objective elements in the order of first appearance */
int bxg_lp_objective(const bxg_lp* lp, int32_t* obj_var, double* obj_factor);
/* name (owned by the handle), bounds and variable_type (0 real, 1 binary, 2 general) */
const char* bxg_lp_variable(const bxg_lp* lp, int i, int* min, int* max, int* type);
/* the same for all n variables at once: min / max / type[n]; the names as n
 * NUL-terminated strings back to back (needed = their total size; a null
 * buffer only asks for it) */
int bxg_lp_variables(const bxg_lp* lp, int32_t* min, int32_t* max, int32_t* type);
int bxg_lp_names(const bxg_lp* lp, char* buffer, size_t capacity, size_t* needed);

/* ---- preprocessor ----
 * What the reference runs in front of the solver when
 * solver_parameters::preprocessor is "all" (lib/src/lpcore.cpp:89-130):
 * bxg_lp_preprocess = baryonyx::preprocess (lib/src/preprocessor.cpp:893-903),
 * bxg_lp_affect = baryonyx::affect (:876-891), bxg_lp_split = baryonyx::split
 * (:857-874; the branch optimizer's step, lib/src/branch-optimizer.cpp:162-163).
 * Variables fixed by their bounds, by rows with one free element or rows that
 * force all their elements, and variables no live row uses; the result is a
 * new handle over the remaining variables (renumbered in order), the rows that
 * still have two or more free elements (right-hand sides reduced), the
 * objective constant raised by the costs of the variables set to one, merged
 * with pre_order like bxg_lp_read.  The fixed variables of a handle, those of
 * the handle it was derived from first: bxg_lp_affected_count / bxg_lp_affected
 * (result::affected_vars, core:727).  Same shrunk problem as the reference on
 * its LP fixtures (tests/test_preprocess.py).  A row that cannot hold is
 * BXG_E_INVALID (the reference's bx_reach()).  Host code, no device needed;
 * the "preprocessed.lp" file the reference leaves in the working directory
 * is not written. */
int bxg_lp_preprocess(const bxg_lp* lp, int pre_order, bxg_lp** out);
int bxg_lp_affect(const bxg_lp* lp, int variable, int value, int pre_order, bxg_lp** out);
int bxg_lp_split(const bxg_lp* lp, int variable, int pre_order, bxg_lp** out_true, bxg_lp** out_false);
int bxg_lp_affected_count(const bxg_lp* lp, int* count);
const char* bxg_lp_affected(const bxg_lp* lp, int i, int* value);

#ifdef __cplusplus
}
#endif
#endif
