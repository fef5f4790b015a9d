/* Build shim (test infrastructure, not product code): SINGLE precision twin of force_f64.h.
 * Mirrors what includes/solver_precision.h:10-22 yields for PRECISION == SINGLE_PRECISION. */
#ifndef SOLVER_PRECISION_H
#define SOLVER_PRECISION_H
#define SINGLE_PRECISION (1)
#define DOUBLE_PRECISION (2)
#define PRECISION (SINGLE_PRECISION)
#define T_P float
#define prc(x) x##f
#define pprc(x) f##x
#endif
