/* sirocco_tables.c -- index-free, bisection-based readers of the SIROCCO tables
 * (include/pluto_b200_tables.h; replaces Src/LineDriven/line_connect.c:43-262). */
#include "pluto_b200_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define LINELENGTH 400      /* line_connect.c:24 */

/* the reference's predicates: read_sirocco_fluxes (line_connect.c:133-134) fabs(1.0 - (xin / x[n])) < tol,
 * read_sirocco_heatcool (line_connect.c:344-345,449-450) fabs((xin - x[n]) / xin) < tol */
static int matches_mode(int mode, double xin, double xn, double tol) {
  return mode == 0 ? fabs(1.0 - (xin / xn)) < tol : fabs((xin - xn) / xin) < tol;
}
#define matches(xin, xn, tol) matches_mode(a.mode, xin, xn, tol)

/* All n in [beg, end] with matches(xin, x[n]), ascending, as [*lo, *hi] (empty: *lo > *hi).
 * x ascending and positive over [beg, end] (the zone centres of a radial / polar grid): the matching
 * zones are contiguous around the bisection point.  Anything else falls back to a linear scan,
 * which is still O(n1 + n2) per row instead of O(n1 n2). */
typedef struct { int sorted, mode; } axis_info;   /* mode: which predicate (0 fluxes, 1 heatcool) */

static axis_info axis_check(const double *x, int beg, int end, int mode) {
  axis_info a = {1, mode};
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
  axis_info a1 = axis_check(g->x1, g->ibeg, g->iend, 0), a2 = axis_check(g->x2, g->jbeg, g->jend, 0);
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

/* ---- read_sirocco_heatcool(): py_heatcool.dat and prefactors.dat (line_connect.c:318-497) ----
 * one row per LINE (fgets + sscanf), tolerance 1e-5, every matching zone takes the row's values */
static long read_lines(FILE *f, const pb200_table_grid *g, int which, double *a0, double *a1) {
  const double tol = 1e-5;
  const long plane = (long)g->nx1_tot * g->nx2_tot;
  char aline[LINELENGTH];
  axis_info a1i = axis_check(g->x1, g->ibeg, g->iend, 1), a2i = axis_check(g->x2, g->jbeg, g->jend, 1);
  long icount = 0;
  int I[MAXM], J[MAXM];
  while (fgets(aline, LINELENGTH, f) != NULL) {
    int ii, jj, nwords;
    double rcen, thetacen, v[16];
    if (which == 0) {   /* py_heatcool.dat: i j rcen thetacen vol t_e t_r xi ne heat*4 cool*3 dens n_h */
      nwords = sscanf(aline, "%d %d %le %le %le %le %le %le %le %le %le %le %le %le %le %le %le %le", &ii, &jj, &rcen, &thetacen,
                      &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8], &v[9], &v[10], &v[11], &v[12], &v[13]);
      if (nwords != 18) return -3;
    } else {            /* prefactors.dat: i rcen j thetacen dens comp_h comp_c xray_h brem_c line_c xi_ion */
      nwords = sscanf(aline, "%d %le %d %le %le %le %le %le %le %le %le", &ii, &rcen, &jj, &thetacen,
                      &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6]);
      if (nwords != 11) return -3;
    }
    int nI = find_matches(g->x1, g->ibeg, g->iend, a1i, rcen / g->unit_length, tol, I, MAXM);
    int nJ = find_matches(g->x2, g->jbeg, g->jend, a2i, thetacen, tol, J, MAXM);
    if (nI > MAXM || nJ > MAXM) return -3;
    for (int q = 0; q < nJ; q++) for (int p = 0; p < nI; p++) {
      long o = (long)J[q] * g->nx1_tot + I[p];
      icount++;
      if (which == 0) {
        double xi = v[3], t_r = v[2];                       /* line_connect.c:353-356 */
        a0[o] = xi; a1[o] = t_r;
        if (a0[o] < 1.0) a0[o] = 1.0;
        if (a1[o] < 1.e3) a1[o] = 1.e3;
      } else {                                              /* line_connect.c:458-463: comp_h, comp_c, xray_h, line_c, brem_c, xi_ion */
        a0[0 * plane + o] = v[1]; a0[1 * plane + o] = v[2]; a0[2 * plane + o] = v[3];
        a0[3 * plane + o] = v[5]; a0[4 * plane + o] = v[4]; a0[5 * plane + o] = v[6];
      }
    }
  }
  return icount;
}

long pb200_read_heatcool_file(const char *path, const pb200_table_grid *g, double *xi, double *t_r) {
  FILE *f = fopen(path, "r");
  char aline[LINELENGTH];
  if (!f) return -1;
  if (fgets(aline, LINELENGTH, f) == NULL) { fclose(f); return -2; }
  long n = read_lines(f, g, 0, xi, t_r);
  fclose(f);
  return n;
}

long pb200_read_prefactors_file(const char *path, const pb200_table_grid *g, double *pre) {
  FILE *f = fopen(path, "r");
  char aline[LINELENGTH];
  if (!f) return -1;
  if (fgets(aline, LINELENGTH, f) == NULL) { fclose(f); return -2; }
  long n = read_lines(f, g, 1, pre, NULL);
  fclose(f);
  return n;
}
