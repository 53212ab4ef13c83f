/* Builds the two instantiations of brax_step_impl.h (float32: the timed CPU baseline; float64: pins the C text to
 * the NumPy float64 oracle in tests/test_oracle_c.py). TEST INFRASTRUCTURE ONLY -- see the header of the template.
 *   gcc -O3 -fno-math-errno -fno-trapping-math -ffp-contract=off -fopenmp -shared -fPIC oracle/brax_step.c -o oracle/_build/libbraxstep.so -lm
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL float
#define SUFFIX _f32
#define SQRT sqrtf
#define ATAN2 atan2f
#include "brax_step_impl.h"
#undef REAL
#undef SUFFIX
#undef SQRT
#undef ATAN2

#define REAL double
#define SUFFIX _f64
#define SQRT sqrt
#define ATAN2 atan2
#include "brax_step_impl.h"
#undef REAL
#undef SUFFIX
#undef SQRT
#undef ATAN2

int brax_step_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
