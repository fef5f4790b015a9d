/* Build shim (test infrastructure, not product code).
 * Force-included (-include) ahead of every reference translation unit when oracle/build_ref.sh
 * compiles /root/reference in DOUBLE precision.  The reference selects precision with a hard-coded
 * #define in includes/solver_precision.h:8; that tree is read-only, so we pre-define the header's
 * include guard and supply the three macros it would have produced for the double build
 * (includes/solver_precision.h:10-22). */
#ifndef SOLVER_PRECISION_H
#define SOLVER_PRECISION_H
#define SINGLE_PRECISION (1)
#define DOUBLE_PRECISION (2)
#define PRECISION (DOUBLE_PRECISION)
#define T_P double
#define prc(x) x
#define pprc(x) x
#endif
