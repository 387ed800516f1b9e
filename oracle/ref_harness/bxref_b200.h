/* bxref_b200.h -- the reference's solve API over the B200 engine
 * (integration/itm_b200.cpp compiled against the reference's headers).
 * TEST INFRASTRUCTURE. */
#ifndef BXREF_B200_H
#define BXREF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* baryonyx::itm::solve_b200(ctx{seed}, problem, restarts, device) on the problem
 * of a bxref_open handle.  Returns result_status, or -100 - solver_error_tag. */
int bxref_solve_b200(void* h, uint32_t seed, int restarts, int device, int8_t* x,
                     double* value, int* remaining, int64_t* loop);
/* baryonyx::itm::optimize_b200(ctx{seed, time_limit, storage}, problem, workers,
 * device, max_rounds): the reference's optimize over the engine.  calls[3] =
 * how often ctx.start / ctx.update / ctx.finish were invoked.  Returns
 * result_status, or -100 - solver_error_tag. */
int bxref_optimize_b200(void* h, uint32_t seed, int workers, int device, double time_limit,
                        int64_t max_rounds, int storage, int8_t* x, double* value,
                        int* remaining, int* n_solutions, int* calls);
#ifdef __cplusplus
}
#endif
#endif
