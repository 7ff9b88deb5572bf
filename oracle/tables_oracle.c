/* tables_oracle.c -- CPU restatement of the reference's SIROCCO table readers
 * (Src/LineDriven/line_connect.c:43-262, read_sirocco_fluxes): every row of a file is matched
 * against ALL interior zones (DOM_LOOP), O(rows x zones), exactly like the reference.
 *
 * TEST INFRASTRUCTURE ONLY (the checker of pluto_sirocco_b200/csrc/sirocco_tables.c in
 * tests/test_sirocco_tables.py).  Parity status: restates the loop structure, the fscanf formats
 * and the matching predicate of the reference line by line; pinned indirectly - the golden
 * line-driven-wind runs of tests/golden (ldw_*.npz) were produced by the reference reading the very
 * files tests/common.py writes, and this reader returns the arrays those files were written from.
 */
#include <math.h>
#include <stdio.h>

#define LINELENGTH 400

typedef struct {
  int nx1_tot, nx2_tot, ibeg, iend, jbeg, jend;
  const double *x1, *x2;
  double unit_length;
} tgrid;

/* line_connect.c:108-165 for one axis file; out[nangles][nx2_tot][nx1_tot] */
long ref_read_flux_file(const char *path, const tgrid *g, int *nangles, double *out) {
  FILE *fptr = fopen(path, "r");
  char aline[LINELENGTH];
  long ii, jj;
  double x1in, x2in, temp, tol = 1e-6;
  int icount = 0, match = 0;
  if (fptr == NULL) return -1;
  if (fgets(aline, LINELENGTH, fptr) == NULL || fgets(aline, LINELENGTH, fptr) == NULL) { fclose(fptr); return -2; }
  if (sscanf(aline, "%*s %*s %ld", &ii) != 1) { fclose(fptr); return -2; }
  *nangles = (int)ii;
  long plane = (long)g->nx1_tot * g->nx2_tot;
  while (fscanf(fptr, "%ld ", &ii) != EOF) {
    if (fscanf(fptr, "%ld %*d %le %le", &jj, &x1in, &x2in) == 3) {
      for (int j = g->jbeg; j <= g->jend; j++) for (int i = g->ibeg; i <= g->iend; i++) {   /* DOM_LOOP */
        if (fabs(1.0 - (x1in / g->unit_length / g->x1[i])) < tol && fabs(1.0 - (x2in / g->x2[j])) < tol) {
          for (int iflux = 0; iflux < *nangles; iflux++) {
            if (fscanf(fptr, "%le", &temp) == 1) out[iflux * plane + (long)j * g->nx1_tot + i] = temp;
            else { fclose(fptr); return -3; }
          }
          match = 1;
          icount++;
        }
      }
      if (match == 0) {
        for (int iflux = 0; iflux < *nangles; iflux++) if (fscanf(fptr, "%le", &temp) != 1) { fclose(fptr); return -3; }
      }
      match = 0;
    } else { fclose(fptr); return -3; }
  }
  fclose(fptr);
  return icount;
}

/* line_connect.c:185-256 */
long ref_read_mfit_file(const char *path, const tgrid *g, int *mpoints, double *t_fit, double *m_fit) {
  FILE *fptr = fopen(path, "r");
  char aline[LINELENGTH];
  long ii, jj;
  double x1in, x2in, temp, tol = 1e-6;
  int icount = 0, match = 0;
  if (fptr == NULL) return -1;
  if (fgets(aline, LINELENGTH, fptr) == NULL) { fclose(fptr); return -2; }
  if (sscanf(aline, "%*s %ld", &ii) != 1) { fclose(fptr); return -2; }
  *mpoints = (int)ii;
  if (fscanf(fptr, "%*s ") != 0) { fclose(fptr); return -2; }
  for (int m = 0; m < *mpoints; m++) {
    if (fscanf(fptr, "%le", &temp) == 1) t_fit[m] = log10(temp);
    else { fclose(fptr); return -3; }
  }
  long plane = (long)g->nx1_tot * g->nx2_tot;
  while (fscanf(fptr, "%ld ", &ii) != EOF) {
    if (fscanf(fptr, "%ld %le %le", &jj, &x1in, &x2in) == 3) {
      for (int j = g->jbeg; j <= g->jend; j++) for (int i = g->ibeg; i <= g->iend; i++) {
        if (fabs(1.0 - (x1in / g->unit_length / g->x1[i])) < tol && fabs(1.0 - (x2in / g->x2[j])) < tol) {
          for (int m = 0; m < *mpoints; m++) {
            if (fscanf(fptr, "%le", &temp) == 1) m_fit[m * plane + (long)j * g->nx1_tot + i] = log10(temp);
            else { fclose(fptr); return -3; }
          }
          match = 1;
          icount++;
        }
      }
      if (match == 0) {
        for (int m = 0; m < *mpoints; m++) if (fscanf(fptr, "%le", &temp) != 1) { fclose(fptr); return -3; }
      }
      match = 0;
    } else { fclose(fptr); return -3; }
  }
  fclose(fptr);
  return icount;
}

/* line_connect.c:318-361: py_heatcool.dat; xi, t_r [nx2_tot][nx1_tot] */
long ref_read_heatcool_file(const char *path, const tgrid *g, double *xi_out, double *tr_out) {
  FILE *fptr_hc = fopen(path, "r");
  char aline[LINELENGTH];
  int ii, jj, nwords, icount = 0;
  double rcen, thetacen, vol, t_e, t_r, xi, ne, heat_xray, heat_comp, heat_lines, heat_ff, cool_comp, cool_lines, cool_ff, dens, n_h;
  double tol = 1e-5;
  if (fptr_hc == NULL) return -1;
  if (fgets(aline, LINELENGTH, fptr_hc) == NULL) { fclose(fptr_hc); return -2; }
  while (fgets(aline, LINELENGTH, fptr_hc) != NULL) {
    nwords = sscanf(aline, "%d %d %le %le %le %le %le %le %le %le %le %le %le %le %le %le %le %le",
                    &ii, &jj, &rcen, &thetacen, &vol, &t_e, &t_r, &xi, &ne, &heat_xray, &heat_comp, &heat_lines, &heat_ff,
                    &cool_comp, &cool_lines, &cool_ff, &dens, &n_h);
    if (nwords == 18) {
      for (int j = g->jbeg; j <= g->jend; j++) for (int i = g->ibeg; i <= g->iend; i++) {
        if (fabs(((rcen / g->unit_length) - g->x1[i]) / (rcen / g->unit_length)) < tol &&
            fabs((thetacen - g->x2[j]) / thetacen) < tol) {
          long o = (long)j * g->nx1_tot + i;
          icount++;
          xi_out[o] = xi;
          tr_out[o] = t_r;
          if (xi_out[o] < 1.0) xi_out[o] = 1.0;
          if (tr_out[o] < 1.e3) tr_out[o] = 1.e3;
        }
      }
    } else { fclose(fptr_hc); return -3; }
  }
  fclose(fptr_hc);
  return icount;
}

/* line_connect.c:420-466: prefactors.dat; pre[6][nx2_tot][nx1_tot] = comp_h, comp_c, xray_h, line_c, brem_c, xi_ion */
long ref_read_prefactors_file(const char *path, const tgrid *g, double *pre) {
  FILE *fptr = fopen(path, "r");
  char aline[LINELENGTH];
  int ii, jj, nwords, icount = 0;
  double rcen, thetacen, dens, comp_h_pre, comp_c_pre, xray_h_pre, brem_c_pre, line_c_pre, xi_ion_pre, tol = 1e-5;
  long plane = (long)g->nx1_tot * g->nx2_tot;
  if (fptr == NULL) return -1;
  if (fgets(aline, LINELENGTH, fptr) == NULL) { fclose(fptr); return -2; }
  while (fgets(aline, LINELENGTH, fptr) != NULL) {
    nwords = sscanf(aline, "%d %le %d %le %le %le %le %le %le %le %le", &ii, &rcen, &jj, &thetacen, &dens, &comp_h_pre,
                    &comp_c_pre, &xray_h_pre, &brem_c_pre, &line_c_pre, &xi_ion_pre);
    if (nwords == 11) {
      for (int j = g->jbeg; j <= g->jend; j++) for (int i = g->ibeg; i <= g->iend; i++) {
        if (fabs(((rcen / g->unit_length) - g->x1[i]) / (rcen / g->unit_length)) < tol &&
            fabs((thetacen - g->x2[j]) / thetacen) < tol) {
          long o = (long)j * g->nx1_tot + i;
          icount++;
          pre[0 * plane + o] = comp_h_pre; pre[1 * plane + o] = comp_c_pre; pre[2 * plane + o] = xray_h_pre;
          pre[3 * plane + o] = line_c_pre; pre[4 * plane + o] = brem_c_pre; pre[5 * plane + o] = xi_ion_pre;
        }
      }
    } else { fclose(fptr); return -3; }
  }
  fclose(fptr);
  return icount;
}
