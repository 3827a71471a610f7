/* Hand-written stand-in for the reference's autoconf-generated config.h.
 * The seven values are the AC_DEFINE defaults of the reference's configure.ac:38-44;
 * autoconf is not installed in this image, so the reference's own ./configure cannot run.
 * No BLAS / MKL: the reference then uses its scalar matrix::multiply loop
 * (src/matrix_cache.cpp:48-54) and std::exp (src/probability.cpp:120). */
#ifndef CAFE_B200_REF_CONFIG_H
#define CAFE_B200_REF_CONFIG_H
#define NUM_OPTIMIZER_INITIALIZATION_ATTEMPTS 100
#define LAMBDA_PERTURBATION_STEP_SIZE 50
#define OPTIMIZER_STRATEGY_SIMILARITY_CUTOFF
#define PHASED_OPTIMIZER_PHASE1_ATTEMPTS 4
#define OPTIMIZER_LOW_PRECISION 1e-3
#define OPTIMIZER_HIGH_PRECISION 1e-6
#endif
