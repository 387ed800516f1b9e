/* bx_oracle.h -- CPU restatement of the reference's constraint-sweep path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load liboracle.so; the
 * product (baryonyx_b200/) never does.
 *
 * Plain C restatement of quesnel/baryonyx v0.5.0 (/root/reference) for the
 * Wedelin/Bastert sweep: solvers 01 / 101 / Z, affect, calculator_sort,
 * compute_order, the solver_functor loop, plus the libstdc++ 13 algorithms
 * whose implementation-defined behaviour shapes the results (std::sort,
 * std::shuffle, bernoulli / uniform_int distributions, minstd_rand0).  Each
 * function cites the reference file:line it follows.
 *
 * Parity pin: the reference's own tests hold no trajectory-level vectors for
 * this path; the pin is the unmodified reference compiled
 * here as oracle/_ref/libbxref.so (tests/test_oracle_vs_ref.py) and the
 * golden vectors generated from it (tests/golden/).
 */
#ifndef BX_ORACLE_H
#define BX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field layout as bxref_params (oracle/ref_harness/bxref.h): mirror of
 * baryonyx::solver_parameters, lib/include/baryonyx/core:294-462. */
typedef struct ox_params
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
} ox_params;

/* problem_solver_type, lib/include/baryonyx/core:542-550 */
enum
{
    OX_EQ_01 = 0,
    OX_EQ_101 = 1,
    OX_EQ_Z = 2,
    OX_INEQ_01 = 3,
    OX_INEQ_101 = 4,
    OX_INEQ_Z = 5
};

/* constraint_order, core:348-359 */
enum
{
    OX_ORDER_NONE = 0,
    OX_ORDER_REVERSING,
    OX_ORDER_RANDOM_SORTING,
    OX_ORDER_INFEASIBILITY_DECR,
    OX_ORDER_INFEASIBILITY_INCR,
    OX_ORDER_LAGRANGIAN_DECR,
    OX_ORDER_LAGRANGIAN_INCR,
    OX_ORDER_PI_SIGN_CHANGE,
    OX_ORDER_CYCLE
};

/* result_status, core:692-701 */
enum
{
    OX_STATUS_SUCCESS = 0,
    OX_STATUS_INTERNAL_ERROR,
    OX_STATUS_UNINITIALIZED,
    OX_STATUS_TIME_LIMIT_REACHED,
    OX_STATUS_KAPPA_MAX_REACHED,
    OX_STATUS_LIMIT_REACHED,
    OX_STATUS_EMPTY_CONTEXT
};

void ox_default_params(ox_params* p);
/* context_set_solver_parameters, lib/src/private.cpp:51-172 */
void ox_clamp_params(const ox_params* in, ox_params* out);

/* manual_course + the walk of optimize(), lib/src/manual-optimizer.cpp:31-160:
 * the 5^5 parameter sets of the manual auto-tuner in the order it tries them
 * (out arrays of 3125 entries each, tables 5 x 5); returns 3125. */
int ox_manual_grid(const ox_params* p, double* theta, double* delta, double* kappa_min,
                   double* kappa_step, double* init_policy_random, double* tables);

/* ---- libstdc++ 13 restatements ---------------------------------------- */
uint32_t ox_minstd_next(uint32_t* state);
int ox_bernoulli(double p, uint32_t* state);
uint64_t ox_uniform_u64(uint64_t a, uint64_t b, uint32_t* state);
void ox_shuffle_int(int* a, int n, uint32_t* state);
void ox_sort_rc(double* value, int* id, int n, int descending);
void ox_sort_pairs(int* first, int* second, int n, int descending);
void ox_normal(double mean, double stddev, int count, double* out,
               uint32_t* state);

/* ---- problem ----------------------------------------------------------- */
typedef struct ox_problem ox_problem;
typedef struct ox_run ox_run;

/* Merged constraints in CSR (what make_merged_constraints returns,
 * lib/src/merged-constraint.cpp:297-343), objective elements in file order. */
ox_problem* ox_problem_create(int m, int n, const int* row_ptr, const int* col,
                              const int* factor, const int* cmin,
                              const int* cmax, int problem_type, int maximize,
                              int n_obj, const int* obj_var,
                              const double* obj_factor, double cost_constant);
void ox_problem_destroy(ox_problem* pb);
/* Value modelled at R[longest row], the slot the reference reads one past its
 * allocation (undefined behaviour there; itm-solver-inequalities-01.cpp:198-203,
 * -101.cpp:240-245).  Default 0.0: the read then costs one Bernoulli draw. */
void ox_problem_set_tail_value(ox_problem* pb, double v);
/* which_problem_type, lib/src/problem.hpp:72-122: 0/1/2 coefficient class
 * and whether only equalities exist. */
int ox_which_problem_type(int nnz, const int* factor, int has_inequality);

/* ---- step-level run (solver_functor set-up, itm-solver-common.hpp:63-133)
 * p must already be clamped. */
ox_run* ox_run_begin(ox_problem* pb, const ox_params* p, uint32_t seed);
void ox_run_end(ox_run* run);
int ox_run_info(ox_run* run, double* delta, double* theta);
int ox_run_sweep(ox_run* run, int push, double kappa, double obj_amp);
int ox_run_state(ox_run* run, uint64_t* x_words, double* pi, double* P,
                 uint32_t* rng_state, int* order_list, int* order_len,
                 double* norm_cost);
int ox_run_set_x(ox_run* run, const uint64_t* x_words);
int ox_run_set_rng(ox_run* run, uint32_t state);
double ox_run_objective(ox_run* run);
int64_t ox_run_updates(ox_run* run);

/* ---- whole solve (solver_functor::operator(), itm-solver-common.hpp:63-225;
 * the wall-clock time limit is not modelled). Returns result_status. */
int ox_solve(ox_problem* pb, const ox_params* p, uint32_t seed, int8_t* x,
             double* value, int* remaining, int64_t* loop, int64_t* updates,
             int64_t* sweeps);

/* The same loop on an existing run (after ox_run_set_rng / ox_run_set_x), P and
 * pi as the run holds them.  optimizer = 1: one call of a worker of
 * optimize_functor::operator() (itm-optimizer-common.hpp:658-750) starting at
 * kappa_start.  best_x_words: n/64+1 blocks. */
int ox_run_solve(ox_run* run, double kappa_start, int optimizer,
                 uint64_t* best_x_words, double* value, int* remaining,
                 int64_t* loop);

/* ---- Z sub-solvers (known-answer tests lib/test/lib.cpp:376-563) -------- */
int ox_bb_solve(int maximize, double* value, int* factor, int* id, int n,
                int bmin, int bmax);
int ox_exhaustive_solve(int maximize, const int* cst_factor, double* value,
                        int* factor, int* id, int n, int bmin, int bmax);

/* ---- optimizer population ------------------------------
 * local_context, storage and best_solution_recorder of
 * lib/src/itm-optimizer-common.hpp:52-594.  `p` must be clamped. */
typedef struct ox_pop ox_pop;
typedef struct ox_pop_ctx ox_pop_ctx;
ox_pop* ox_pop_create(ox_problem* pb, const ox_params* p, uint64_t master_seed,
                      int threads, uint32_t* rng_state);
void ox_pop_destroy(ox_pop* pop);
int ox_pop_size(ox_pop* pop);
/* member at sorted position i (storage::get_best(i)) */
void ox_pop_member(ox_pop* pop, int i, uint64_t* x_words, double* value,
                   int* remaining, uint64_t* hash, int64_t* loop);
ox_pop_ctx* ox_pop_ctx_create(ox_pop* pop, unsigned thread_id, uint64_t seed);
void ox_pop_ctx_destroy(ox_pop_ctx* c);
uint32_t ox_pop_ctx_rng(ox_pop_ctx* c);
void ox_pop_ctx_set_rng(ox_pop_ctx* c, uint32_t state);
double ox_pop_reinit(ox_pop* pop, ox_pop_ctx* c, double kappa_min,
                     double kappa_max, uint64_t* x_words);
void ox_pop_mutation(ox_pop* pop, ox_pop_ctx* c, uint64_t* x_words);
void ox_pop_crossover(ox_pop* pop, ox_pop_ctx* c, uint64_t* x_words);
void ox_pop_try_update(ox_pop* pop, ox_pop_ctx* c, const uint64_t* x_words,
                       double value, int64_t loop);
void ox_pop_try_advance(ox_pop* pop, ox_pop_ctx* c, const uint64_t* x_words,
                        int remaining, int64_t loop);
uint64_t ox_bit_array_hash(int n, const uint64_t* x_words);

/* optimize_functor::operator() at thread = 1 (itm-optimizer-common.hpp:596-752),
 * set up like optimize_problem (:795-826) from ctx.parameters.seed = seed and run
 * for `rounds` restarts.  Same layout as bxref_opt_trace (ref_harness/bxref.h):
 * snapshot r is the state at the first sweep of restart r. */
typedef struct ox_opt_trace ox_opt_trace;
ox_opt_trace* ox_opt_trace_run(ox_problem* pb, const ox_params* p, uint32_t seed, int rounds);
void ox_opt_trace_free(ox_opt_trace* t);
int ox_opt_trace_rounds(ox_opt_trace* t);
int ox_opt_trace_round(ox_opt_trace* t, int r, uint64_t* x_words, uint32_t* rng_state,
                       double* kappa, double* pi, double* P);
int ox_opt_trace_member(ox_opt_trace* t, int r, int i, uint64_t* x_words, double* value,
                        int* remaining, uint64_t* hash, int64_t* loop);
int ox_opt_trace_sweeps(ox_opt_trace* t);
int ox_opt_trace_sweep_log(ox_opt_trace* t, int* round, int* push, double* kappa, int* len);

#ifdef __cplusplus
}
#endif
#endif
