/* pluto_shim.c -- the reference-side binding of libplutob200.so.
 *
 * Compiled TOGETHER WITH THE REFERENCE (it includes the reference's own pluto.h and the
 * problem's definitions.h) and linked INSTEAD OF Src/Time_Stepping/rk_step.o, so that every
 * other reference object - main(), the pluto.ini parser, set_grid, init.c, output - stays
 * unmodified.  It replaces exactly one symbol:
 *
 *     int AdvanceStep (Data *d, timeStep *Dts, Grid *grid)      Src/prototypes.h:5
 *                                                               Src/Time_Stepping/rk_step.c:29
 *
 * Modes (environment):
 *   default            strict drop-in: d->Vc on the host is authoritative; every call does
 *                      H2D(d->Vc) -> AdvanceStep on the GPU -> D2H(d->Vc).
 *   PB200_RESIDENT=1   the state stays in HBM between steps; d->Vc is refreshed only before
 *                      WriteData()/Analysis() (link with -Wl,--wrap=WriteData,--wrap=Analysis).
 */
#include "pluto.h"
#include "pluto_b200.h"

static pb200_ctx *s_ctx = NULL;
static int s_resident = 0, s_dirty = 0;
static int s_host_bc = 0;   /* boundaries are filled by the reference's own Boundary() on the host */

static int translate_limiter(void) {
#ifdef LIMITER
#if LIMITER == DEFAULT
  return PB200_LIM_DEFAULT;
#elif LIMITER == FLAT_LIM
  return PB200_LIM_FLAT;
#elif LIMITER == MINMOD_LIM
  return PB200_LIM_MINMOD;
#elif LIMITER == VANLEER_LIM
  return PB200_LIM_VANLEER;
#elif LIMITER == MC_LIM
  return PB200_LIM_MC;
#elif LIMITER == VANALBADA_LIM
  return PB200_LIM_VANALBADA;
#elif LIMITER == OSPRE_LIM
  return PB200_LIM_OSPRE;
#elif LIMITER == UMIST_LIM
  return PB200_LIM_UMIST;
#else
  return -1;
#endif
#else
  return PB200_LIM_DEFAULT;
#endif
}

#if BODY_FORCE != NO
/* Evaluate a user body-force callback once on the reference's own grid arrays and hand the
 * result over as the smallest strided table: axes the values do not depend on get stride 0
 * (a constant g is one double, a Phi(x2) potential NX2_TOT doubles).                        */
typedef int (*bf_setter)(pb200_ctx *, int, const double *, long, long, long, long);
static void shim_set_table(bf_setter set, int sel, const double *full)
{
  long n[3] = {NX1_TOT, NX2_TOT, NX3_TOT}, st_full[3] = {1, NX1_TOT, (long)NX1_TOT*NX2_TOT};
  long dep[3] = {0, 0, 0}, st[3], m[3], i, j, k, cnt = 1;
  double *tab;
  for (k = 0; k < n[2]; k++) for (j = 0; j < n[1]; j++) for (i = 0; i < n[0]; i++) {
    double q = full[k*st_full[2] + j*st_full[1] + i];
    if (q != full[k*st_full[2] + j*st_full[1]])  dep[0] = 1;
    if (q != full[k*st_full[2] + i])             dep[1] = 1;
    if (q != full[j*st_full[1] + i])             dep[2] = 1;
  }
  for (i = 0; i < 3; i++) { m[i] = dep[i] ? n[i] : 1; st[i] = dep[i] ? cnt : 0; cnt *= m[i]; }
  tab = (double *) malloc(cnt*sizeof(double));
  for (k = 0; k < m[2]; k++) for (j = 0; j < m[1]; j++) for (i = 0; i < m[0]; i++)
    tab[i*st[0] + j*st[1] + k*st[2]] = full[k*st_full[2] + j*st_full[1] + i];
  if (set(s_ctx, sel, tab, cnt, st[0], st[1], st[2]) != PB200_OK) {
    print ("! AdvanceStep(): body-force table: %s\n", pb200_last_error());
    QUIT_PLUTO(1);
  }
  free(tab);
}

static void shim_body_force(Data *d, Grid *grid)
{
  long ntot = (long)NX1_TOT*NX2_TOT*NX3_TOT, o;
  int i, j, k, nv, c;
  double *x1 = grid->x[IDIR], *x2 = grid->x[JDIR], *x3 = grid->x[KDIR];
  double *full = (double *) malloc(3*ntot*sizeof(double));
#if (BODY_FORCE & VECTOR)
  double v[NVAR], g[3];
  TOT_LOOP(k,j,i) {                      /* rhs_source.c:256,367,429 */
    NVAR_LOOP(nv) v[nv] = d->Vc[nv][k][j][i];
    g[0] = g[1] = g[2] = 0.0;
    BodyForceVector(v, g, x1[i], x2[j], x3[k]);
    o = ((long)k*NX2_TOT + j)*NX1_TOT + i;
    for (c = 0; c < 3; c++) full[c*ntot + o] = g[c];
  }
  for (c = 0; c < 3; c++) shim_set_table(pb200_set_body_force_vector, c, full + c*ntot);
#endif
#if (BODY_FORCE & POTENTIAL)
  for (c = 0; c < 4; c++) {              /* rhs.c:168-182, rhs_source.c:279,382,441 */
    if (c > DIMENSIONS) break;
    TOT_LOOP(k,j,i) {
      o = ((long)k*NX2_TOT + j)*NX1_TOT + i;
      full[o] = BodyForcePotential(c == 1 ? grid->xr[IDIR][i] : x1[i],
                                   c == 2 ? grid->xr[JDIR][j] : x2[j],
                                   c == 3 ? grid->xr[KDIR][k] : x3[k]);
    }
    shim_set_table(pb200_set_body_force_potential, c, full);
  }
#endif
  free(full);
}
#endif

#if LINE_DRIVEN_WIND != NO
/* read_sirocco_fluxes() (Src/LineDriven/line_connect.c:43-262, called from Src/main.c:168,189) with the
 * bisection-based readers of sirocco_tables.c: same files, same globals (NFLUX_ANGLES, flux_{r,t,p}_UV,
 * MPOINTS, t_fit, M_UV_fit), O(rows log zones) instead of O(rows x zones) - 1.2 s instead of 54 s per
 * file at 512 x 256 zones.  Linked with --wrap=read_sirocco_fluxes; taken when PB200_FAST_TABLES=1
 * (round 1: opt-in, the default stays the reference's own reader).  The reference leaves the zones no
 * row matches (the ghost zones) as malloc() returned them; here they are zero.                     */
#include "pluto_b200_tables.h"
void __real_read_sirocco_fluxes (Data *d, Grid *grid);
void __wrap_read_sirocco_fluxes (Data *d, Grid *grid)
{
  const char *env = getenv("PB200_FAST_TABLES");
  if (env == NULL || atoi(env) == 0) { __real_read_sirocco_fluxes (d, grid); return; }
  static const char *names[3] = {"directional_flux_r.dat", "directional_flux_theta.dat", "directional_flux_phi.dat"};
  double *****dst[3] = {&flux_r_UV, &flux_t_UV, &flux_p_UV};
  pb200_table_grid tg;
  tg.nx1_tot = NX1_TOT; tg.nx2_tot = NX2_TOT;
  tg.ibeg = IBEG; tg.iend = IEND; tg.jbeg = JBEG; tg.jend = JEND;
  tg.x1 = grid->x[IDIR]; tg.x2 = grid->x[JDIR];
  tg.unit_length = UNIT_LENGTH;
  if (NX3_TOT != 1) { print ("! read_sirocco_fluxes (libplutob200): the tables are 2-D (NX3_TOT = 1)\n"); QUIT_PLUTO(1); }
  print ("> read_sirocco_fluxes: libplutob200 table readers\n");
  for (int ax = 0; ax < 3; ax++) {
    int na = pb200_flux_file_nangles (names[ax]);
    if (na == -1) { print ("No flux file %s\n", names[ax]); continue; }    /* line_connect.c:80-83 */
    if (na < 1) { print ("! %s: flux header improperly formatted\n", names[ax]); QUIT_PLUTO(1); }
    if (ax == 0) NFLUX_ANGLES = na;
    else if (na != NFLUX_ANGLES) { print ("! %s does not agree in NFLUX_ANGLES\n", names[ax]); QUIT_PLUTO(1); }
    *dst[ax] = ARRAY_4D(NFLUX_ANGLES, NX3_TOT, NX2_TOT, NX1_TOT, double);
    memset ((*dst[ax])[0][0][0], 0, sizeof(double)*(size_t)NFLUX_ANGLES*NX3_TOT*NX2_TOT*NX1_TOT);
    long n = pb200_read_flux_file (names[ax], &tg, NFLUX_ANGLES, (*dst[ax])[0][0][0]);
    if (n < 0) { print ("! %s: error %ld in reading flux file\n", names[ax], n); QUIT_PLUTO(1); }
    print ("Read %d fluxes for %ld cells\n", NFLUX_ANGLES, n);
  }
  if (g_inputParam[KRAD] == 999 && g_inputParam[ALPHARAD] == 999) {          /* line_connect.c:185-256 */
    int mp = 0;
    long n = pb200_read_mfit_file ("M_UV_data.dat", &tg, &mp, NULL, NULL);
    if (n == -1) { print ("No force multiplier file\n"); return; }
    if (n < 0 || mp < 1) { print ("! M_UV_data.dat: bad header\n"); QUIT_PLUTO(1); }
    MPOINTS = mp;
    M_UV_fit = ARRAY_4D(MPOINTS, NX3_TOT, NX2_TOT, NX1_TOT, double);
    memset (M_UV_fit[0][0][0], 0, sizeof(double)*(size_t)MPOINTS*NX3_TOT*NX2_TOT*NX1_TOT);
    t_fit = calloc (MPOINTS, sizeof(double));
    n = pb200_read_mfit_file ("M_UV_data.dat", &tg, &mp, t_fit, M_UV_fit[0][0][0]);
    if (n < 0) { print ("! M_UV_data.dat: error %ld in reading force multiplier file\n", n); QUIT_PLUTO(1); }
    print ("Read %d points to M vs t fits for %ld cells\n", MPOINTS, n);
  }
}

/* LINE_DRIVEN_WIND SIROCCO_MODE: parameters of the cv_idl problem and the flux tables that
 * read_sirocco_fluxes() left in the globals flux_{r,t,p}_UV[NFLUX_ANGLES][k][j][i]
 * (Src/globals.h:193-200, Src/main.c:157-200).  ARRAY_4D payloads are contiguous.           */
static void shim_line_driven_wind(void)
{
  pb200_ldw_config l;
  memset(&l, 0, sizeof(l));
  l.nangles       = NFLUX_ANGLES;
  l.userdef_bc    = 1;       /* the UserDefBoundary() of Test_Problems/LineDrivenWind/cv_idl */
  l.unit_length   = UNIT_LENGTH;
  l.unit_velocity = UNIT_VELOCITY;
  l.unit_density  = UNIT_DENSITY;
  l.mu        = g_inputParam[MU];
  l.krad      = g_inputParam[KRAD];
  l.alpharad  = g_inputParam[ALPHARAD];
  l.dfloor    = g_inputParam[DFLOOR];
  l.rho_0     = g_inputParam[RHO_0];
  l.rho_alpha = g_inputParam[RHO_ALPHA];
  l.cent_mass = g_inputParam[CENT_MASS];
  l.disk_mdot = g_inputParam[DISK_MDOT];
  l.lx        = g_inputParam[L_star]*g_inputParam[f_x];
  l.tx        = g_inputParam[T_x];
  if (flux_r_UV == NULL || flux_t_UV == NULL) {
    print ("! AdvanceStep(): no sirocco flux tables (directional_flux_*.dat) were read\n");
    QUIT_PLUTO(1);
  }
  if (pb200_ldw_enable(s_ctx, &l) != PB200_OK ||
      pb200_ldw_set_fluxes(s_ctx, flux_r_UV[0][0][0], flux_t_UV[0][0][0],
                           flux_p_UV != NULL ? flux_p_UV[0][0][0] : NULL) != PB200_OK) {
    print ("! AdvanceStep(): line-driven wind set-up failed: %s\n", pb200_last_error());
    QUIT_PLUTO(1);
  }
  if (l.krad == 999 && l.alpharad == 999) {     /* M(t) fit of M_UV_data.dat (line_connect.c:185-256) */
    if (M_UV_fit == NULL || t_fit == NULL ||
        pb200_ldw_set_mfit(s_ctx, MPOINTS, t_fit, M_UV_fit[0][0][0]) != PB200_OK) {
      print ("! AdvanceStep(): force-multiplier fit missing or not accepted: %s\n", pb200_last_error());
      QUIT_PLUTO(1);
    }
  }
}
#endif

static void shim_init(Data *d, Grid *grid) {
  pb200_config cfg;
  int dir;
  pb200_config_default(&cfg);
#if PHYSICS != HD || EOS != IDEAL
#error "libplutob200: only PHYSICS HD with EOS IDEAL is on the B200 path"
#endif
  cfg.dimensions = DIMENSIONS;
  cfg.geometry   = GEOMETRY;          /* same codes, Src/pluto.h:34-37 */
  cfg.nghost     = GetNghost();
  cfg.ntracer    = NTRACER;
#if RECONSTRUCTION == FLAT
  cfg.reconstruction = PB200_FLAT;
#elif RECONSTRUCTION == LINEAR
  cfg.reconstruction = PB200_LINEAR;
#elif RECONSTRUCTION == PARABOLIC
  cfg.reconstruction = PB200_PARABOLIC;
#else
#error "libplutob200: RECONSTRUCTION must be FLAT, LINEAR or PARABOLIC"
#endif
#if TIME_STEPPING == EULER
  cfg.time_stepping = PB200_EULER;
#elif TIME_STEPPING == RK2
  cfg.time_stepping = PB200_RK2;
#elif TIME_STEPPING == RK3
  cfg.time_stepping = PB200_RK3;
#else
#error "libplutob200: TIME_STEPPING must be EULER, RK2 or RK3"
#endif
  cfg.limiter = translate_limiter();
#if CHAR_LIMITING == YES
  cfg.char_limiting = 1;
#endif
#if SHOCK_FLATTENING == MULTID
  cfg.shock_flattening = 1;
#elif SHOCK_FLATTENING != NO
#error "libplutob200: SHOCK_FLATTENING must be NO or MULTID"
#endif
#if ENTROPY_SWITCH == ALWAYS || ENTROPY_SWITCH == SELECTIVE
  cfg.entropy_switch = ENTROPY_SWITCH;      /* same codes, Src/pluto.h:60-61 */
#elif ENTROPY_SWITCH != NO
#error "libplutob200: ENTROPY_SWITCH must be NO, SELECTIVE or ALWAYS"
#endif
#if (BODY_FORCE & VECTOR)
  cfg.body_force |= PB200_BF_VECTOR;
#endif
#if (BODY_FORCE & POTENTIAL)
  cfg.body_force |= PB200_BF_POTENTIAL;
#endif
  /* SetSolver() stored a function pointer (Src/HD/set_solver.c:4-58) */
  if      (d->fluidRiemannSolver == &HLLC_Solver) cfg.solver = PB200_HLLC;
  else if (d->fluidRiemannSolver == &HLL_Solver)  cfg.solver = PB200_HLL;
  else if (d->fluidRiemannSolver == &LF_Solver)   cfg.solver = PB200_TVDLF;
  else {
    print ("! AdvanceStep(): this Riemann solver is not available in libplutob200\n");
    QUIT_PLUTO(1);
  }
  /* USERDEF sides (other than the line-driven-wind problem's, which has device versions) and
     INTERNAL_BOUNDARY are arbitrary host code: fall back to the reference's Boundary() per stage */
#if LINE_DRIVEN_WIND == NO
  for (dir = 0; dir < DIMENSIONS; dir++)
    if (grid->lbound[dir] == USERDEF || grid->rbound[dir] == USERDEF) s_host_bc = 1;
#if INTERNAL_BOUNDARY == YES
  s_host_bc = 1;
#endif
#endif
  if (getenv("PB200_HOST_BOUNDARY")) s_host_bc = atoi(getenv("PB200_HOST_BOUNDARY"));
  for (dir = 0; dir < 3; dir++) {
    cfg.nx[dir]   = grid->np_int[dir];
    cfg.xbeg[dir] = grid->xbeg[dir];
    cfg.xend[dir] = grid->xend[dir];
    cfg.bc[2*dir]     = s_host_bc ? PB200_BC_NEIGHBOUR : grid->lbound[dir];   /* same codes, Src/pluto.h:163-170 */
    cfg.bc[2*dir + 1] = s_host_bc ? PB200_BC_NEIGHBOUR : grid->rbound[dir];
  }
  cfg.gamma          = g_gamma;
  cfg.small_density  = g_smallDensity;
  cfg.small_pressure = g_smallPressure;
  if (getenv("PB200_DEVICE")) cfg.device = atoi(getenv("PB200_DEVICE"));
  s_resident = getenv("PB200_RESIDENT") && atoi(getenv("PB200_RESIDENT"));
  if (pb200_create(&cfg, &s_ctx) != PB200_OK) {
    print ("! AdvanceStep(): pb200_create: %s\n", pb200_last_error());
    QUIT_PLUTO(1);
  }
  for (dir = 0; dir < DIMENSIONS; dir++) {   /* the reference's own grid arrays */
    if (pb200_set_grid(s_ctx, dir, grid->xl[dir], grid->xr[dir], grid->dx[dir]) != PB200_OK) {
      print ("! AdvanceStep(): pb200_set_grid: %s\n", pb200_last_error());
      QUIT_PLUTO(1);
    }
  }
#if BODY_FORCE != NO
  shim_body_force(d, grid);
#endif
#if LINE_DRIVEN_WIND != NO
  shim_line_driven_wind();
#if COOLING == BLONDIN
  {   /* Data tables of the BLONDIN module (Src/structs.h:621-643, contiguous ARRAY_3D payloads) */
    const double *tabs[7] = {d->comp_h_pre[0][0], d->comp_c_pre[0][0], d->xray_h_pre[0][0], d->line_c_pre[0][0],
                             d->brem_c_pre[0][0], d->sirocco_xi[0][0], d->sirocco_t_r[0][0]};
    if (pb200_cooling_set_tables(s_ctx, tabs) != PB200_OK) {
      print ("! AdvanceStep(): pb200_cooling_set_tables failed\n");
      QUIT_PLUTO(1);
    }
  }
#endif
#endif
  print ("> AdvanceStep() runs on the GPU (libplutob200 v%d, %s mode%s)\n", pb200_version(),
         s_resident ? "resident" : "strict host-buffer",
         s_host_bc ? ", boundaries by the host's Boundary()/UserDefBoundary()" : "");
}

int AdvanceStep (Data *d, timeStep *Dts, Grid *grid)
{
  pb200_step_info info;
  int rc;
  double *vc = d->Vc[0][0][0];   /* contiguous [NVAR][NX3_TOT][NX2_TOT][NX1_TOT], Src/arrays.c:251 */

  if (s_ctx == NULL) {
    shim_init(d, grid);
    /* page-lock the contiguous payload of d->Vc (ARRAY_4D, Src/arrays.c:251-330) for the copies */
    pb200_host_register(vc, (size_t)NVAR*NX3_TOT*NX2_TOT*NX1_TOT*sizeof(double));
    if (s_resident) pb200_upload_vc(s_ctx, vc);
  }
  if (s_host_bc) {
    /* per stage: D2H, the reference's Boundary() (-> UserDefBoundary) on d->Vc, H2D, stage */
    int s, ns = pb200_nstages(s_ctx);
    rc = pb200_step_begin(s_ctx, g_dt);
    for (s = 1; s <= ns && rc == PB200_OK; s++) {
      g_intStage = s;
      if (s > 1 || s_resident) rc = pb200_stage_download(s_ctx, s, vc);
      if (rc != PB200_OK) break;
      Boundary (d, 0, grid);
      rc = pb200_stage_upload(s_ctx, s, vc);
      if (rc == PB200_OK) rc = pb200_stage(s_ctx, s);
    }
    if (rc == PB200_OK) rc = pb200_step_end(s_ctx, &info);
    if (rc == PB200_OK) rc = pb200_download_vc(s_ctx, vc);
  } else if (s_resident) {
    rc = pb200_advance_step(s_ctx, g_dt, &info);
    s_dirty = 1;
  } else {
    rc = pb200_advance_step_host(s_ctx, vc, g_dt, &info);
  }
  if (rc != PB200_OK) {
    print ("! AdvanceStep(): %s\n", pb200_last_error());
    QUIT_PLUTO(1);
  }
  Dts->invDt_hyp = MAX(Dts->invDt_hyp, info.invDt_hyp);   /* update_stage.c:320,391 */
  g_maxMach      = MAX(g_maxMach, info.maxMach);          /* hll_speed.c:89 */
  return 0;
}

/* resident mode: refresh d->Vc just before the reference reads it */
void pb200_shim_sync_to_host(const Data *d)
{
  if (s_ctx != NULL && s_resident && s_dirty) {
    pb200_download_vc(s_ctx, d->Vc[0][0][0]);
    s_dirty = 0;
  }
}
/* resident mode with COOLING BLONDIN: the source step runs on the device copy as well
 * (link with -Wl,--wrap=SplitSource); host-buffer mode keeps the reference's own SplitSource() */
void __real_SplitSource (Data *, double, timeStep *, Grid *);
void __wrap_SplitSource (Data *d, double dt, timeStep *Dts, Grid *grid)
{
#if COOLING == BLONDIN && LINE_DRIVEN_WIND != NO
  if (s_ctx != NULL && s_resident) {
    if (pb200_split_source(s_ctx, dt, g_time) != PB200_OK) {
      print ("! SplitSource(): pb200_split_source failed\n");
      QUIT_PLUTO(1);
    }
    s_dirty = 1;
    return;
  }
#endif
  __real_SplitSource(d, dt, Dts, grid);
}

void __real_WriteData (const Data *, Output *, Grid *);
void __wrap_WriteData (const Data *d, Output *output, Grid *grid)
{
  pb200_shim_sync_to_host(d);
  __real_WriteData(d, output, grid);
}
void __real_Analysis (const Data *, Grid *);
void __wrap_Analysis (const Data *d, Grid *grid)
{
  pb200_shim_sync_to_host(d);
  __real_Analysis(d, grid);
}
