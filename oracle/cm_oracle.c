/*
 * cm_oracle.c -- CPU oracle for the contrast-maximization hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see cm_oracle_impl.h).  Builds the fp32 oracle
 * (operation-for-operation restatement of the reference's eager CPU path)
 * and an fp64 twin of the same algorithm used for triangulation.
 *
 * Parity pin: the reference ships no tests or golden vectors (SURVEY.md §4,
 * §8c); this oracle is pinned against the reference itself, imported live in
 * the build container, by tests/test_oracle_vs_reference_live.py and through the
 * committed vectors under tests/golden/ (made by tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "cm_oracle.h"

/* ACC = type of the image / flow-gradient accumulators.  ACC == REAL restates the reference (its sums are fp32 in
 * index order).  The third instance below (suffix _f32x: fp32 per-event arithmetic, double accumulators, each sum
 * rounded to fp32 once) is NOT the reference: it is the summation-order-free yardstick the tests use to show that what
 * separates the CUDA path from the reference is the order of fp32 additions and nothing else. */
#define REAL float
#define ACC float
#define FN(n) n##_f32
#define R_FMA fmaf
#define R_FLOOR floorf
#define R_FABS fabsf
#define R_RINT rintf
#include "cm_oracle_impl.h"
#undef REAL
#undef ACC
#undef FN
#undef R_FMA
#undef R_FLOOR
#undef R_FABS
#undef R_RINT

#define REAL float
#define ACC double
#define FN(n) n##_f32x
#define R_FMA fmaf
#define R_FLOOR floorf
#define R_FABS fabsf
#define R_RINT rintf
#include "cm_oracle_impl.h"
#undef REAL
#undef ACC
#undef FN
#undef R_FMA
#undef R_FLOOR
#undef R_FABS
#undef R_RINT

#define REAL double
#define ACC double
#define FN(n) n##_f64
#define R_FMA fma
#define R_FLOOR floor
#define R_FABS fabs
#define R_RINT rint
#include "cm_oracle_impl.h"

int orc_num_slots(const orc_cfg *c, int linear)
{
    return linear ? linear_slots_f32(c, NULL) : build_slots_f32(c, NULL);
}
int orc_version(void) { return 1; }

/* OpenMP team size of the timed CPU baseline: launchers such as torchrun export OMP_NUM_THREADS=1, so bench.py sets it
   explicitly and reports what it got */
#ifdef _OPENMP
#include <omp.h>
int orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int orc_set_threads(int n) { (void)n; return 1; }
#endif
