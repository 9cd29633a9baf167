/*
 * pano_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle: a plain-C restatement of the grid fluid step of
 * msiglreith/panopaea (examples/dec_fluid.rs, panopaea/src/dec/grid.rs,
 * panopaea/src/pcg.rs, panopaea/src/math/{interp,linear_view}.rs).  The Rust
 * reference cannot be compiled in this image (no rustc/cargo, dependencies not
 * vendored), so this file is "kind: port" everywhere it is timed.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (panopaea_b200/)
 * never does.
 *
 * Pinning: checked against the reference's own golden vectors
 * (panopaea/src/dec/grid.rs:428-482 divergence, :484-515 Laplacian) in
 * tests/test_oracle_golden.py.  advect, advect_mac, the CG loop, dot and
 * max-norm have NO reference test: parity unpinned there except through an
 * independent numpy restatement (oracle/np_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

typedef struct { size_t y0, y1, x0, x1; } orc_rect;

/* 0 serial, 1 reference-faithful threading, 2 all passes parallel */
static int g_threads = 0;
EXPORT void orc_set_threading(int mode) { g_threads = mode; }
EXPORT int orc_get_threading(void) { return g_threads; }

#ifdef _OPENMP
#include <omp.h>
EXPORT int orc_max_threads(void) { return omp_get_max_threads(); }
EXPORT void orc_set_num_threads(int n) { omp_set_num_threads(n); }
#else
EXPORT int orc_max_threads(void) { return 1; }
EXPORT void orc_set_num_threads(int n) { (void)n; }
#endif

#define REAL double
#define FN(x) x##_f64
#define FMIN fmin
#define FMAX fmax
#define FLOOR floor
#include "pano_oracle_body.inc"
#undef REAL
#undef FN
#undef FMIN
#undef FMAX
#undef FLOOR

#include "pano_oracle_mg.inc"
#include "pano_oracle3.inc"   /* Grid3d specification (f64), uses the f64 body above */

#define REAL float
#define FN(x) x##_f32
#define FMIN fminf
#define FMAX fmaxf
#define FLOOR floorf
#include "pano_oracle_body.inc"
#undef REAL
#undef FN
#undef FMIN
#undef FMAX
#undef FLOOR
