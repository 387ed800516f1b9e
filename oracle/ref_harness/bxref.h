/* bxref.h -- C entry points of the reference harness (oracle/_ref/libbxref.so).
 *
 * TEST INFRASTRUCTURE ONLY.  This library is the UNMODIFIED reference
 * (quesnel/baryonyx, /root/reference) compiled where it lies by
 * oracle/Makefile, plus the thin harness in this directory that reaches the
 * reference's templates (solver classes, compute_order, solver_functor,
 * optimize_problem) and exposes them through plain C functions for ctypes.
 * Nothing of the product (baryonyx_b200/) may link or load it.
 */
#ifndef BXREF_H
#define BXREF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of baryonyx::solver_parameters (lib/include/baryonyx/core:294-462),
 * flattened to a POD.  Enum fields carry the reference's integer values. */
typedef struct bxref_params
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
} bxref_params;

void bxref_default_params(bxref_params* p);

/* Problem handle: parse LP (reference parser), unpreprocess, clamp parameters
 * (context_set_solver_parameters), make_merged_constraints. */
void* bxref_open(const char* lp_path, const bxref_params* p);
/* the same behind baryonyx::preprocess and a chain of baryonyx::affect calls
 * (lib/src/preprocessor.cpp:876-903); null + err on a contract failure */
void* bxref_open_pp(const char* lp_path, const bxref_params* p, int n_affect,
                    const int* variable, const int* value, char* err, int err_len);
int bxref_affected_count(void* h);
const char* bxref_affected(void* h, int i, int* value);
/* The grid of the manual auto-tuner (manual-optimizer.cpp:31-160): the
 * parameter sets its optimize() hands to itm::optimize, in its order; returns
 * how many there are (5^5), the first `capacity` are written; tables = the
 * 5 x 5 value tables of manual_course. */
int bxref_manual_grid(const bxref_params* p, double* theta, double* delta, double* kappa_min,
                      double* kappa_step, double* init_policy_random, int capacity, double* tables);
/* The same from arrays (merged-constraint form), without the LP parser. */
void* bxref_open_arrays(int m, int n, const int* row_ptr, const int* col,
                        const int* factor, const int* cmin, const int* cmax,
                        int maximize, int n_obj, const int* obj_var,
                        const double* obj_factor, double cost_constant,
                        const bxref_params* p);
void bxref_close(void* h);
/* problem_type = baryonyx::problem_solver_type value; maximize = 1 if
 * objective_function_type::maximize. */
int bxref_info(void* h, int* m, int* n, int* nnz, int* problem_type,
               int* maximize, double* cost_constant, int* n_obj);
int bxref_get_params(void* h, bxref_params* out); /* after clamping */
/* merged constraints, CSR */
int bxref_export_rows(void* h, int* row_ptr, int* col, int* factor, int* cmin,
                      int* cmax);
/* objective elements in file order */
int bxref_export_objective(void* h, int* var, double* factor);
/* variable i of the parsed problem: name (valid until the next call on this
 * thread), bounds, baryonyx::variable_type value */
const char* bxref_variable(void* h, int i, int* min, int* max, int* type);

/* Step-level run (replicates the set-up of solver_functor::operator(),
 * lib/src/itm-solver-common.hpp:63-133, with the reference's own Solver and
 * compute_order objects; the caller drives kappa). */
void* bxref_run_begin(void* h, uint32_t seed);
void bxref_run_end(void* run);
int bxref_run_info(void* run, double* delta, double* theta);
int bxref_run_sweep(void* run, int push, double kappa, double obj_amp);
/* value the reference reads one past its R array (undefined behaviour in the
 * reference, itm-solver-inequalities-01.cpp:198-203) */
double bxref_run_tail_value(void* run);
/* Any pointer may be NULL. x_words has n/64+1 entries; P is nnz entries in
 * the reference's row-major order; order list = compute_order::R (or m_order
 * for the infeasibility orders, as (k, amount) pairs flattened). */
int bxref_run_state(void* run, uint64_t* x_words, double* pi, double* P,
                    uint32_t* rng_state, int* order_list, int* order_len,
                    double* norm_cost);
int bxref_run_set_x(void* run, const uint64_t* x_words); /* + compute.init */
int bxref_run_set_rng(void* run, uint32_t state);
double bxref_run_objective(void* run); /* original_costs.results(x, const) */

/* Whole solve through the reference's own solve_problem<> with a counting
 * Solver subclass.  Returns result_status; x is n bytes (0/1). */
int bxref_solve(void* h, uint32_t seed, int8_t* x, double* value,
                int* remaining, int64_t* updates, int64_t* sweeps);
/* Public API path: baryonyx::solve() without harness types (x n bytes). */
int bxref_solve_public(void* h, uint32_t seed, int8_t* x, double* value);
/* Multithreaded optimizer with the counting subclass. */
int bxref_optimize(void* h, uint32_t seed, int threads, double time_limit,
                   int8_t* x, double* value, int* remaining, int64_t* updates,
                   int64_t* sweeps, double* seconds);

/* Of the last bxref_optimize: seconds until the best member first had no violated
 * row (-1: never; 1 s granularity, the reference's own ctx.update report) and
 * the restarts its workers began. */
void bxref_optimize_report(double* first_feasible_seconds, int64_t* restarts);

/* Trace of the optimizer at thread = 1: the reference's own optimize_functor
 * (itm-optimizer-common.hpp:596-752) on the calling thread, set up as
 * optimize_problem does (:795-826), stopped after `rounds` restarts.  Snapshot
 * r (0 <= r <= rounds) is taken at the first sweep of restart r (0-based), i.e.
 * after reinit + compute_order::init: snapshot `rounds` holds everything the
 * first `rounds` restarts did.  Members are in sorted order. */
void* bxref_opt_trace(void* h, uint32_t seed, int rounds);
void bxref_opt_trace_free(void* t);
int bxref_opt_trace_rounds(void* t); /* snapshots taken */
int bxref_opt_trace_round(void* t, int r, uint64_t* x_words, uint32_t* rng_state,
                          double* kappa, double* pi, double* P);
int bxref_opt_trace_member(void* t, int r, int i, uint64_t* x_words, double* value,
                           int* remaining, uint64_t* hash, int64_t* loop);
/* one entry per sweep of the traced restarts: restart, push flag, kappa, length
 * of the order list */
int bxref_opt_trace_sweeps(void* t);
int bxref_opt_trace_sweep_log(void* t, int* round, int* push, double* kappa, int* len);

/* libstdc++ primitives the path depends on (pins for the C restatement). */
void bxref_std_sort(double* value, int* id, int n, int descending);
void bxref_std_sort_pairs(int* first, int* second, int n, int descending);
void bxref_std_shuffle(int* a, int n, uint32_t* rng_state);
int bxref_std_bernoulli(double p, uint32_t* rng_state);
uint64_t bxref_std_uniform_u64(uint64_t a, uint64_t b, uint32_t* rng_state);
void bxref_std_normal(double mean, double stddev, int count, double* out,
                      uint32_t* rng_state);
double bxref_std_pow(double a, double b);
/* generate_seed (itm-common.hpp:1457-1484) from random_engine(master_seed). */
void bxref_generate_seed(uint64_t master_seed, unsigned count, uint64_t* out);
/* Sub-solver known-answer access (lib/test/lib.cpp:376-563). */
int bxref_bb_solve(int maximize, double* value, int* factor, int* id, int n,
                   int bmin, int bmax);
int bxref_exhaustive_solve(int maximize, const int* cst_factor, double* value,
                           int* factor, int* id, int n, int bmin, int bmax);

#ifdef __cplusplus
}
#endif
#endif
