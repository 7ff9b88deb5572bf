/* sirocco_tables.c -- index-free, bisection-based readers of the SIROCCO tables
 * (include/pluto_b200_tables.h; replaces Src/LineDriven/line_connect.c:43-262). */
#include "pluto_b200_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define LINELENGTH 400      /* line_connect.c:24 */

/* the reference's predicate (line_connect.c:133-134): fabs(1.0 - (xin / x[n])) < tol */
static int matches(double xin, double xn, double tol) { return fabs(1.0 - (xin / xn)) < tol; }

/* All n in [beg, end] with matches(xin, x[n]), ascending, as [*lo, *hi] (empty: *lo > *hi).
 * x ascending and positive over [beg, end] (the zone centres of a radial / polar grid): the matching
 * zones are contiguous around the bisection point.  Anything else falls back to a linear scan,
 * which is still O(n1 + n2) per row instead of O(n1 n2). */
typedef struct { int sorted; } axis_info;

static axis_info axis_check(const double *x, int beg, int end) {
  axis_info a = {1};
  if (!(x[beg] > 0.0)) a.sorted = 0;
  for (int n = beg; n < end && a.sorted; n++) if (!(x[n + 1] > x[n])) a.sorted = 0;
  return a;
}

/* returns the number of matches and their indices (ascending) in idx (at most cap are stored) */
static int find_matches(const double *x, int beg, int end, axis_info a, double xin, double tol, int *idx, int cap) {
  int n = 0;
  if (!a.sorted) {
    for (int q = beg; q <= end; q++) if (matches(xin, x[q], tol)) { if (n < cap) idx[n] = q; n++; }
    return n;
  }
  int lo = beg, hi = end;                 /* first q with x[q] >= xin */
  while (lo < hi) { int mid = lo + (hi - lo) / 2; if (x[mid] < xin) lo = mid + 1; else hi = mid; }
  int first = lo;
  while (first > beg && matches(xin, x[first - 1], tol)) first--;
  if (!matches(xin, x[first], tol)) {     /* the bisection point itself may be just outside */
    if (first + 1 <= end && matches(xin, x[first + 1], tol)) first++;
    else return 0;
  }
  for (int q = first; q <= end && matches(xin, x[q], tol); q++) { if (n < cap) idx[n] = q; n++; }
  return n;
}

int pb200_flux_file_nangles(const char *path) {
  FILE *f = fopen(path, "r");
  char aline[LINELENGTH];
  long ii;
  if (!f) return -1;
  if (fgets(aline, LINELENGTH, f) == NULL || fgets(aline, LINELENGTH, f) == NULL) { fclose(f); return -2; }
  fclose(f);
  if (sscanf(aline, "%*s %*s %ld", &ii) != 1) return -2;
  return (int)ii;
}

#define MAXM 64

/* rows `ii jj [inwind] x1 x2 v0 ... v{nval-1}`: the shared body of the two readers */
static long read_rows(FILE *f, const pb200_table_grid *g, int skip_third, int nval, int take_log10, double *out) {
  const double tol = 1e-6;
  const long plane = (long)g->nx1_tot * g->nx2_tot;
  axis_info a1 = axis_check(g->x1, g->ibeg, g->iend), a2 = axis_check(g->x2, g->jbeg, g->jend);
  long ii, jj, icount = 0;
  double x1in, x2in, temp;
  int I[MAXM], J[MAXM];
  while (fscanf(f, "%ld ", &ii) != EOF) {
    int got = skip_third ? fscanf(f, "%ld %*d %le %le", &jj, &x1in, &x2in) : fscanf(f, "%ld %le %le", &jj, &x1in, &x2in);
    if (got != 3) return -3;
    int nI = find_matches(g->x1, g->ibeg, g->iend, a1, x1in / g->unit_length, tol, I, MAXM);
    int nJ = find_matches(g->x2, g->jbeg, g->jend, a2, x2in, tol, J, MAXM);
    if (nI > MAXM || nJ > MAXM) return -3;       /* a tolerance wider than the zones: not a table for this grid */
    if (nI > 0 && nJ > 0) {
      /* DOM_LOOP order (j outer, i inner): one set of values per matching zone */
      for (int q = 0; q < nJ; q++) for (int p = 0; p < nI; p++) {
        for (int m = 0; m < nval; m++) {
          if (fscanf(f, "%le", &temp) != 1) return -3;
          out[m * plane + (long)J[q] * g->nx1_tot + I[p]] = take_log10 ? log10(temp) : temp;
        }
        icount++;
      }
    } else {
      for (int m = 0; m < nval; m++) if (fscanf(f, "%le", &temp) != 1) return -3;
    }
  }
  return icount;
}

long pb200_read_flux_file(const char *path, const pb200_table_grid *g, int nangles, double *out) {
  FILE *f = fopen(path, "r");
  char aline[LINELENGTH];
  long ii;
  if (!f) return -1;
  if (fgets(aline, LINELENGTH, f) == NULL || fgets(aline, LINELENGTH, f) == NULL) { fclose(f); return -2; }
  if (sscanf(aline, "%*s %*s %ld", &ii) != 1) { fclose(f); return -2; }
  if (ii != nangles) { fclose(f); return -4; }
  long n = read_rows(f, g, 1, nangles, 0, out);
  fclose(f);
  return n;
}

long pb200_read_mfit_file(const char *path, const pb200_table_grid *g, int *mpoints, double *t_fit, double *m_fit) {
  FILE *f = fopen(path, "r");
  char aline[LINELENGTH];
  long ii;
  double temp;
  if (!f) return -1;
  if (fgets(aline, LINELENGTH, f) == NULL) { fclose(f); return -2; }
  if (sscanf(aline, "%*s %ld", &ii) != 1) { fclose(f); return -2; }
  *mpoints = (int)ii;
  if (!t_fit) { fclose(f); return 0; }
  if (fscanf(f, "%*s ") != 0) { fclose(f); return -2; }
  for (int m = 0; m < *mpoints; m++) {
    if (fscanf(f, "%le", &temp) != 1) { fclose(f); return -3; }
    t_fit[m] = log10(temp);
  }
  long n = read_rows(f, g, 0, *mpoints, 1, m_fit);
  fclose(f);
  return n;
}
