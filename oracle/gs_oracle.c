/*
 * gs_oracle.c -- CPU oracle for the 3D Gaussian Splatting rasterizer hot path (plain C + OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product (robosimgs_b200/) never links or calls it.
 *
 * PARITY UNPINNED -- see the header of gs_oracle_impl.h and DESIGN.md section "Oracle".
 *
 * Two instantiations of the same source: *_f32 (arithmetic type of the CUDA path; used as the
 * CPU baseline) and *_f64 (gradient/tolerance reference).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SFX _f32
#define REAL_IS_FLOAT 1
#define SQRT sqrtf
#define CEIL ceilf
#define EXP expf
#include "gs_oracle_impl.h"
#undef REAL
#undef SFX
#undef REAL_IS_FLOAT
#undef SQRT
#undef CEIL
#undef EXP

#define REAL double
#define SFX _f64
#define REAL_IS_FLOAT 0
#define SQRT sqrt
#define CEIL ceil
#define EXP exp
#include "gs_oracle_impl.h"

#ifdef _OPENMP
#include <omp.h>
int gso_num_threads(void) { return omp_get_max_threads(); }
void gso_set_num_threads(int n) { omp_set_num_threads(n); }
#else
int gso_num_threads(void) { return 1; }
void gso_set_num_threads(int n) { (void)n; }
#endif
