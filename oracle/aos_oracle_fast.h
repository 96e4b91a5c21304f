/* aos_oracle_fast.h -- spatially indexed / multi-threaded variants of the oracle's quadratic or full-image loops.
 * TEST INFRASTRUCTURE ONLY (see aos_oracle.h).  Every function here returns exactly what the literal loop in
 * aos_oracle_seed.c / aos_oracle_gvd.c returns (tests/test_oracle_fast_cpu.py compares them on every golden and on
 * random cases); they exist so that the oracle finishes BASELINE config 3 (20000 x 20000 cells, 200 M points) in
 * about a minute -- for bit-exact parity tests at full size and as the same-config CPU arm of bench.py. */
#ifndef AOS_ORACLE_FAST_H
#define AOS_ORACLE_FAST_H
#include <stddef.h>
#include <stdint.h>
#include "aos_oracle.h"

int orc_fast_enabled(void);
int orc_fast_threads(void);
int orc_fast_skip_labels(void);

typedef void (*orc_range_fn)(size_t lo, size_t hi, void *arg);
/* fn over [0, n) split into contiguous chunks, one per thread (orc_fast_threads()). */
void orc_parallel_for(size_t n, orc_range_fn fn, void *arg);

void orc_fast_bin_points(const orc_seed_params *p, const float *points, size_t n, size_t stride_floats, int w, int h,
                         double ox, double oy, int8_t *grid);
void orc_fast_inflate(const int8_t *in, int w, int h, int cells, int8_t *out);
void orc_fast_open_cross(const int8_t *in, int w, int h, int8_t *out);
int orc_fast_thin_zhangsuen(int8_t *grid, int w, int h);
/* max pairwise integer squared distance of n cells (linear indices x + y*w): convex hull, then hull pairs */
int64_t orc_fast_max_pair_d2(const int32_t *cells, int n, int w);

/* uniform hash grid over points added in increasing index order */
typedef struct {
  double cell;
  size_t cap;        /* power of two */
  int64_t *key;
  int32_t *head;     /* per slot: most recently added point of that cell, -1 = empty slot */
  int32_t *next;     /* per point */
  double *xy;        /* per point (copied) */
  int n, ncap;
} orc_sgrid;
void orc_sgrid_init(orc_sgrid *g, double cell, int expected_points);
void orc_sgrid_free(orc_sgrid *g);
int orc_sgrid_add(orc_sgrid *g, double x, double y);   /* returns the point's index */
/* indices of all points in the (2*rings+1)^2 cells around (x, y), unordered; returns the count (out may be NULL to count) */
int orc_sgrid_gather(const orc_sgrid *g, double x, double y, int rings, int32_t *out, int max_out);
void orc_sort_i32(int32_t *a, int n);
#endif
