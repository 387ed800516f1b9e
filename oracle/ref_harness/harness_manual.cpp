// harness_manual.cpp -- test infrastructure (see bxref.h).  Pulls the
// reference's manual auto-tuner translation unit in so that its file-local
// `manual_course` (the parameter grid and its odometer) is reachable; the file
// is read from /root/reference/lib/src at build time, never copied.  The one
// external function of that file is renamed here: the library already holds
// the reference's own object of it.
#define manual_optimize bxref_manual_optimize_second_copy
#include "manual-optimizer.cpp"
#undef manual_optimize

#include "bxref.h"

extern "C" int
bxref_manual_grid(const bxref_params* p, double* theta, double* delta, double* kappa_min,
                  double* kappa_step, double* init_policy_random, int capacity, double* tables)
{
    // manual-optimizer.cpp:116-121: the course the reference's optimize() walks
    manual_course array(p->theta, p->delta, p->kappa_min, p->kappa_step, p->init_policy_random, 5);
    if (tables)
        for (int i = 0; i != 5; ++i) {
            tables[0 * 5 + i] = array.theta[i];
            tables[1 * 5 + i] = array.delta[i];
            tables[2 * 5 + i] = array.kappa_min[i];
            tables[3 * 5 + i] = array.kappa_step[i];
            tables[4 * 5 + i] = array.init_policy_random[i];
        }
    int count = 0;
    do { // :130-137: what every pass hands to itm::optimize
        if (count < capacity) {
            theta[count] = array.theta[array.iterators[0]];
            delta[count] = array.delta[array.iterators[1]];
            kappa_min[count] = array.kappa_min[array.iterators[2]];
            kappa_step[count] = array.kappa_step[array.iterators[3]];
            init_policy_random[count] = array.init_policy_random[array.iterators[4]];
        }
        ++count;
    } while (array.next());
    return count;
}
