/* pluto_b200_tables.h -- host-side ingestion of the SIROCCO tables of the line-driven-wind coupling.
 *
 * Replaces the readers of Src/LineDriven/line_connect.c:43-262 (read_sirocco_fluxes: the three
 * directional_flux_{r,theta,phi}.dat files and the force-multiplier fit M_UV_data.dat) and :267-497
 * (read_sirocco_heatcool: py_heatcool.dat, prefactors.dat).  The
 * reference walks ALL interior zones for every row of a file to find the zone whose centre matches
 * the row's coordinates (line_connect.c:132-150, :223-243): O(rows x zones), 2.7e11 coordinate tests
 * per file on the 1024 x 512 grid.  These functions keep the file formats, the parsing (the same
 * fscanf conversions) and the matching predicate, and find the matching zones by bisection on the
 * two coordinate arrays: O(rows x log(zones)).  Results are identical to the reference's readers,
 * including rows that match no zone (skipped) and rows whose coordinates match several zones (the
 * reference then consumes one set of values per matching zone; so do these, up to 64 matching
 * zones per axis - beyond that the table cannot be told apart on this grid and -3 is returned).
 *
 * Plain C, no CUDA: the arrays land in host memory in the reference's own layout
 * [table][k][j][i] (k extent NX3_TOT = 1 for the 2-D problem) and go to the device through
 * pb200_ldw_set_fluxes() / pb200_ldw_set_mfit() (pluto_b200.h).  The drop-in shim binds them behind the
 * reference's own entry point: --wrap=read_sirocco_fluxes (Src/prototypes.h:254, called from
 * Src/main.c:168,189), taken when PB200_FAST_TABLES=1.
 */
#ifndef PLUTO_B200_TABLES_H
#define PLUTO_B200_TABLES_H

#ifdef __cplusplus
extern "C" {
#endif

/* the part of the reference's Grid the readers use (grid->x[IDIR], grid->x[JDIR], IBEG..JEND) */
typedef struct pb200_table_grid {
  int nx1_tot, nx2_tot;      /* NX1_TOT, NX2_TOT */
  int ibeg, iend, jbeg, jend;/* IBEG, IEND, JBEG, JEND (DOM_LOOP bounds) */
  const double *x1, *x2;     /* grid->x[IDIR], grid->x[JDIR] (zone centres incl. ghosts) */
  double unit_length;        /* UNIT_LENGTH */
} pb200_table_grid;

/* Number of angular bins announced by the 2nd header line of a directional flux file
 * (line_connect.c:92-93); < 0 on error (-1 no file, -2 bad header). */
int pb200_flux_file_nangles(const char *path);

/* One directional flux file into out[nangles][nx2_tot][nx1_tot] (zones that no row matches keep
 * what out held).  Returns the number of zones filled (the reference's icount), < 0 on error:
 * -1 no file, -2 bad header, -3 truncated / malformed row, -4 nangles differs from the file's. */
long pb200_read_flux_file(const char *path, const pb200_table_grid *g, int nangles, double *out);

/* M_UV_data.dat (line_connect.c:185-256): *mpoints from the header; t_fit[mpoints] = log10(t);
 * m_fit[mpoints][nx2_tot][nx1_tot] = log10(M).  Call with t_fit == NULL to get *mpoints only.
 * Returns the number of zones filled, < 0 on error as above. */
long pb200_read_mfit_file(const char *path, const pb200_table_grid *g, int *mpoints, double *t_fit,
                          double *m_fit);

/* read_sirocco_heatcool() (line_connect.c:267-497), the two files of a restarted run; one row per
 * line, tolerance 1e-5, predicate fabs((x_row - x_zone)/x_row) < tol:
 *   py_heatcool.dat -> xi[nx2_tot][nx1_tot] (floored at 1) and t_r[...] (floored at 1e3), line_connect.c:334-361;
 *   prefactors.dat  -> pre[6][nx2_tot][nx1_tot] = comp_h, comp_c, xray_h, line_c, brem_c, xi_ion, :433-466.
 * Zones no row matches keep what the arrays held (the reference presets the prefactors to 1).
 * Return the number of zone assignments (the reference's icount), < 0 on error as above. */
long pb200_read_heatcool_file(const char *path, const pb200_table_grid *g, double *xi, double *t_r);
long pb200_read_prefactors_file(const char *path, const pb200_table_grid *g, double *pre);

#ifdef __cplusplus
}
#endif
#endif
