/* bxref_pop.h -- C view of the reference's optimizer population
 * (best_solution_recorder / storage / local_context,
 * /root/reference/lib/src/itm-optimizer-common.hpp:52-594), instantiated
 * unmodified by harness_population.cpp.  TEST INFRASTRUCTURE ONLY. */
#ifndef BXREF_POP_H
#define BXREF_POP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* best_solution_recorder(rng(master_seed), threads, cost, constant, n,
 * constraints, parameters.init_population_size) on the problem of a bxref_open
 * handle.  rng_state: engine state after the constructor. */
void* bxref_pop_create(void* h, uint64_t master_seed, int threads, uint32_t* rng_state);
void bxref_pop_destroy(void* pop);
int bxref_pop_size(void* pop);
/* Member at sorted position i (storage::get_best(i)). */
void bxref_pop_member(void* pop, int i, uint64_t* x_words, double* value,
                      int* remaining, uint64_t* hash, int64_t* loop);
/* local_context(ctx, thread_id, seed). */
void* bxref_pop_ctx_create(void* pop, unsigned thread_id, uint64_t seed);
void bxref_pop_ctx_destroy(void* ctx);
uint32_t bxref_pop_ctx_rng(void* ctx);
/* best_solution_recorder::reinit -> kappa; x in/out. */
double bxref_pop_reinit(void* pop, void* ctx, double kappa_min, double kappa_max,
                        uint64_t* x_words);
void bxref_pop_mutation(void* pop, void* ctx, uint64_t* x_words);
void bxref_pop_crossover(void* pop, void* ctx, uint64_t* x_words);
void bxref_pop_try_update(void* pop, void* ctx, const uint64_t* x_words, double value,
                          int64_t loop);
void bxref_pop_try_advance(void* pop, void* ctx, const uint64_t* x_words, int remaining,
                           int64_t loop);
uint64_t bxref_bit_array_hash(int n, const uint64_t* x_words);

#ifdef __cplusplus
}
#endif
#endif
