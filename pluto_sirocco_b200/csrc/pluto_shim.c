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
 *   PB200_NGPUS=N      slab-decompose the grid over N GPUs of the box from this one host thread
 *                      (pb200_multi_*: NCCL halo exchange and dt reduction inside the library);
 *                      the replacement of the reference's MPI layer (Src/Parallel, boundary.c:139-158,
 *                      main.c:288,547) for the unmodified serial C driver.  Cartesian 2-D/3-D path.
 *   PB200_DEVICES=a,b,..   device ordinal of every rank (may repeat: ranks sharing a GPU)
 *   PB200_HOST_BOUNDARY=1  force the reference's own Boundary() on the host per stage.
 *   PB200_LDW_CVIDL_BC=0|1 line-driven wind: never / always use the device copies of cv_idl's
 *                      UserDefBoundary(); default: use them only if they REPRODUCE the linked
 *                      UserDefBoundary() on the initial state and on a perturbed copy of it.
 *
 * No silent wrong answers: user callbacks the device path cannot honour are detected at start-up
 * (state-dependent BodyForceVector, FLAG_INTERNAL_BOUNDARY zones, a UserDefBoundary() that is not
 * cv_idl's) and either routed through the host (slow, correct) or refused with QUIT_PLUTO.
 */
#include "pluto.h"
#include "pluto_b200.h"

static pb200_ctx *s_ctx = NULL;
static pb200_multi *s_multi = NULL;   /* PB200_NGPUS > 1 */
static int s_resident = 0, s_dirty = 0;
static int s_host_bc = 0;   /* boundaries are filled by the reference's own Boundary() on the host */
static int s_bf_time_dependent = 0;   /* BodyForce*() reads g_time: tables are re-evaluated every step */
static int s_ldw_device_bc = 0;       /* the device copies of cv_idl's UserDefBoundary() are in use */

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
static void shim_set_table(int vector, int sel, const double *full)
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
  int rc;
  if (s_multi != NULL) rc = vector ? pb200_multi_set_body_force_vector(s_multi, sel, tab, cnt, st[0], st[1], st[2])
                                   : pb200_multi_set_body_force_potential(s_multi, sel, tab, cnt, st[0], st[1], st[2]);
  else                 rc = vector ? pb200_set_body_force_vector(s_ctx, sel, tab, cnt, st[0], st[1], st[2])
                                   : pb200_set_body_force_potential(s_ctx, sel, tab, cnt, st[0], st[1], st[2]);
  if (rc != PB200_OK) {
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
  for (c = 0; c < 3; c++) shim_set_table(1, c, full + c*ntot);
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
    shim_set_table(0, c, full);
  }
#endif
  free(full);
}

/* The device path tabulates the body force by POSITION.  Probe the user's callbacks once: a force
 * that changes with the state v[] cannot be tabulated (refused), one that changes with g_time is
 * re-tabulated before every step.                                                               */
static void shim_probe_body_force(Data *d, Grid *grid)
{
  int i, j, k, nv, c, n = 0, dep_v = 0, dep_t = 0;
  double t_save = g_time;
  TOT_LOOP(k,j,i) {
    if ((n++) % 97 != 0) continue;                  /* a sample of zones is enough */
    double x1 = grid->x[IDIR][i], x2 = grid->x[JDIR][j], x3 = grid->x[KDIR][k];
#if (BODY_FORCE & VECTOR)
    double v[NVAR], w[NVAR], g0[3] = {0,0,0}, g1[3] = {0,0,0}, g2[3] = {0,0,0};
    NVAR_LOOP(nv) { v[nv] = d->Vc[nv][k][j][i]; w[nv] = 1.37*v[nv] + 0.11; }
    BodyForceVector(v, g0, x1, x2, x3);
    BodyForceVector(w, g1, x1, x2, x3);
    g_time = t_save + 1.2345;
    BodyForceVector(v, g2, x1, x2, x3);
    g_time = t_save;
    for (c = 0; c < 3; c++) { if (g1[c] != g0[c]) dep_v = 1; if (g2[c] != g0[c]) dep_t = 1; }
#endif
#if (BODY_FORCE & POTENTIAL)
    double p0 = BodyForcePotential(x1, x2, x3), p1;
    g_time = t_save + 1.2345;
    p1 = BodyForcePotential(x1, x2, x3);
    g_time = t_save;
    if (p1 != p0) dep_t = 1;
#endif
  }
  if (dep_v) {
    print ("! AdvanceStep(): BodyForceVector() depends on the state v[]; libplutob200 tabulates the body\n");
    print ("!                force by position and cannot honour that.  Not supported on the GPU path.\n");
    QUIT_PLUTO(1);
  }
  s_bf_time_dependent = dep_t;
  if (dep_t) print ("> AdvanceStep(): the body force depends on g_time: tables are re-evaluated every step\n");
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

/* read_sirocco_heatcool() (Src/LineDriven/line_connect.c:267-497, called from Src/main.c:176,197): the
 * restart branch (flag != 0) reads py_heatcool.dat and prefactors.dat with the same O(rows x zones)
 * search; with PB200_FAST_TABLES=1 the library's bisection readers fill the same Data arrays
 * (sirocco_xi, sirocco_t_r and the six *_pre tables, Src/structs.h:621-643).  Link with
 * --wrap=read_sirocco_heatcool.  flag == 0 (analytic initialisation) stays the reference's code. */
void __real_read_sirocco_heatcool (Data *d, Grid *grid, int flag);
void __wrap_read_sirocco_heatcool (Data *d, Grid *grid, int flag)
{
  const char *env = getenv("PB200_FAST_TABLES");
#if COOLING == NO
  __real_read_sirocco_heatcool (d, grid, flag); (void)env; return;
#else
  if (env == NULL || atoi(env) == 0 || flag == 0) { __real_read_sirocco_heatcool (d, grid, flag); return; }
  int i, j, k;
  long n;
  pb200_table_grid tg;
  tg.nx1_tot = NX1_TOT; tg.nx2_tot = NX2_TOT;
  tg.ibeg = IBEG; tg.iend = IEND; tg.jbeg = JBEG; tg.jend = JEND;
  tg.x1 = grid->x[IDIR]; tg.x2 = grid->x[JDIR];
  tg.unit_length = UNIT_LENGTH;
  if (NX3_TOT != 1) { print ("! read_sirocco_heatcool (libplutob200): the tables are 2-D (NX3_TOT = 1)\n"); QUIT_PLUTO(1); }
  print ("> read_sirocco_heatcool: libplutob200 table readers\n");
  n = pb200_read_heatcool_file ("py_heatcool.dat", &tg, d->sirocco_xi[0][0], d->sirocco_t_r[0][0]);
  if (n == -1) {                                       /* line_connect.c:309-323: no file -> analytic values */
    print ("NO py_heatcool file\n");
    DOM_LOOP(k,j,i) {
      double rho = d->Vc[RHO][k][j][i]*UNIT_DENSITY, r = grid->x[IDIR][i]*UNIT_LENGTH;
      double nH = rho/(1.43*CONST_mp), lx = g_inputParam[L_star]*g_inputParam[f_x];
      d->sirocco_xi[k][j][i]  = lx/nH/(r*r);
      d->sirocco_t_r[k][j][i] = g_inputParam[T_x];
    }
  } else if (n < 0) { print ("! py_heatcool file incorrectly formatted (error %ld)\n", n); QUIT_PLUTO(1); }
  else print ("Read in %ld py_heatcool entries\n", n);
  DOM_LOOP(k,j,i) {                                    /* line_connect.c:400-410 */
    d->comp_h_pre[k][j][i] = d->comp_c_pre[k][j][i] = d->xray_h_pre[k][j][i] = 1.0;
    d->line_c_pre[k][j][i] = d->brem_c_pre[k][j][i] = d->xi_ion_pre[k][j][i] = 1.0;
  }
  {
    size_t nz = (size_t)NX2_TOT*NX1_TOT;
    double *pre = (double *) malloc (6*nz*sizeof(double));
    double *dst[6] = {d->comp_h_pre[0][0], d->comp_c_pre[0][0], d->xray_h_pre[0][0], d->line_c_pre[0][0],
                      d->brem_c_pre[0][0], d->xi_ion_pre[0][0]};
    int q;
    for (q = 0; q < 6; q++) memcpy (pre + q*nz, dst[q], nz*sizeof(double));
    n = pb200_read_prefactors_file ("prefactors.dat", &tg, pre);
    if (n == -1) print ("NO prefactor file\n");
    else if (n < 0) { print ("! Prefactor file incorrectly formatted (error %ld)\n", n); QUIT_PLUTO(1); }
    else {
      for (q = 0; q < 6; q++) memcpy (dst[q], pre + q*nz, nz*sizeof(double));
      print ("Read in %ld prefactors\n", n);
    }
    free (pre);
  }
#endif
}

/* LINE_DRIVEN_WIND SIROCCO_MODE: parameters of the cv_idl problem and the flux tables that
 * read_sirocco_fluxes() left in the globals flux_{r,t,p}_UV[NFLUX_ANGLES][k][j][i]
 * (Src/globals.h:193-200, Src/main.c:157-200).  ARRAY_4D payloads are contiguous.           */
static void shim_line_driven_wind(int device_bc)
{
  pb200_ldw_config l;
  memset(&l, 0, sizeof(l));
  l.nangles       = NFLUX_ANGLES;
  l.userdef_bc    = device_bc;   /* device copies of the UserDefBoundary() of Test_Problems/LineDrivenWind/cv_idl */
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
#if EOS == ISOTHERMAL
  l.t_iso     = g_inputParam[T_ISO];          /* line_connect.c:851-855 */
#endif
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

static void shim_fill_config(pb200_config *pcfg, Data *d, Grid *grid) {
  pb200_config cfg;
  int dir;
  pb200_config_default(&cfg);
#if PHYSICS != HD || (EOS != IDEAL && EOS != ISOTHERMAL)
#error "libplutob200: PHYSICS HD with EOS IDEAL or ISOTHERMAL is on the B200 path"
#endif
#if EOS == ISOTHERMAL
  cfg.eos = PB200_EOS_ISOTHERMAL;
  cfg.iso_sound_speed = g_isoSoundSpeed;      /* set by the user's Init() (e.g. cv_iso/init.c:61-63) */
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
#if RING_AVERAGE > 1
  cfg.ring_average = RING_AVERAGE;
  cfg.ring_average_rec = RING_AVERAGE_REC;
#endif
#if SHOCK_FLATTENING == MULTID
  cfg.shock_flattening = 1;
#elif SHOCK_FLATTENING == ONED
  cfg.shock_flattening = 2;
#elif SHOCK_FLATTENING != NO
#error "libplutob200: SHOCK_FLATTENING must be NO, ONED or MULTID"
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
  else if (d->fluidRiemannSolver == &Roe_Solver)  cfg.solver = PB200_ROE;
#if EOS == IDEAL
  else if (d->fluidRiemannSolver == &TwoShock_Solver) cfg.solver = PB200_TWO_SHOCK;
  else if (d->fluidRiemannSolver == &AUSMp_Solver)    cfg.solver = PB200_AUSM;
#endif
  else {
    print ("! AdvanceStep(): this Riemann solver is not available in libplutob200\n");
    QUIT_PLUTO(1);
  }
  for (dir = 0; dir < 3; dir++) {
    cfg.nx[dir]   = grid->np_int[dir];
    cfg.xbeg[dir] = grid->xbeg[dir];
    cfg.xend[dir] = grid->xend[dir];
    cfg.bc[2*dir]     = s_host_bc ? PB200_BC_NEIGHBOUR : grid->lbound[dir];   /* same codes, Src/pluto.h:163-170 */
    cfg.bc[2*dir + 1] = s_host_bc ? PB200_BC_NEIGHBOUR : grid->rbound[dir];
  }
#if EOS == IDEAL
  cfg.gamma          = g_gamma;
#endif
  cfg.small_density  = g_smallDensity;
  cfg.small_pressure = g_smallPressure;
  if (getenv("PB200_DEVICE")) cfg.device = atoi(getenv("PB200_DEVICE"));
  *pcfg = cfg;
}

/* create the device context(s) for the current s_host_bc / device-boundary choice and hand over the
 * grid, the body-force tables and the line-driven-wind tables                                     */
static void shim_create(Data *d, Grid *grid, int ngpus, int ldw_device_bc) {
  pb200_config cfg;
  int dir, rc;
  shim_fill_config(&cfg, d, grid);
  if (ngpus > 1) {
    int devs[64], nd = 0;                      /* PB200_DEVICES=0,1,...: device ordinal of every rank */
    const char *dl = getenv("PB200_DEVICES");
    while (dl != NULL && *dl != '\0' && nd < 64) { devs[nd++] = atoi(dl); dl = strchr(dl, ','); if (dl) dl++; }
    rc = pb200_multi_create(&cfg, ngpus, nd == ngpus ? devs : NULL, &s_multi);
    if (rc != PB200_OK) { print ("! AdvanceStep(): pb200_multi_create (PB200_NGPUS=%d): %s\n", ngpus, pb200_last_error()); QUIT_PLUTO(1); }
    s_ctx = pb200_multi_ctx(s_multi, 0);
  } else if (pb200_create(&cfg, &s_ctx) != PB200_OK) {
    print ("! AdvanceStep(): pb200_create: %s\n", pb200_last_error());
    QUIT_PLUTO(1);
  }
  for (dir = 0; dir < DIMENSIONS; dir++) {   /* the reference's own grid arrays */
    rc = s_multi ? pb200_multi_set_grid(s_multi, dir, grid->xl[dir], grid->xr[dir], grid->dx[dir])
                 : pb200_set_grid(s_ctx, dir, grid->xl[dir], grid->xr[dir], grid->dx[dir]);
    if (rc != PB200_OK) { print ("! AdvanceStep(): pb200_set_grid: %s\n", pb200_last_error()); QUIT_PLUTO(1); }
  }
  if (s_multi == NULL && pb200_set_grid_uniform(s_ctx, grid->uniform) != PB200_OK) {
    print ("! AdvanceStep(): pb200_set_grid_uniform: %s\n", pb200_last_error()); QUIT_PLUTO(1);
  }
#if GEOMETRY != CARTESIAN
  if (s_multi == NULL) {     /* the reference's own geometry arrays (Src/set_geometry.c:49-61), not a re-evaluation */
    pb200_geometry geo;
    geo.dV = grid->dV[0][0];
    geo.A[0] = &grid->A[IDIR][0][0][-1]; geo.A[1] = &grid->A[JDIR][0][-1][0]; geo.A[2] = &grid->A[KDIR][-1][0][0];
    geo.dx_dl[0] = grid->dx_dl[IDIR][0]; geo.dx_dl[1] = grid->dx_dl[JDIR][0]; geo.dx_dl[2] = grid->dx_dl[KDIR][0];
    geo.rt = grid->rt; geo.s = grid->s; geo.sp = grid->sp;
    if (pb200_set_geometry(s_ctx, &geo) != PB200_OK) { print ("! AdvanceStep(): pb200_set_geometry: %s\n", pb200_last_error()); QUIT_PLUTO(1); }
  }
#endif
#if BODY_FORCE != NO
  shim_body_force(d, grid);
#endif
#if LINE_DRIVEN_WIND != NO
  shim_line_driven_wind(ldw_device_bc);
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
#else
  (void)ldw_device_bc;
#endif
}

#if LINE_DRIVEN_WIND != NO
/* Do the device copies of cv_idl's UserDefBoundary() reproduce the UserDefBoundary() this executable
 * was linked with?  Boundary(d, 0, grid) on the host against pb200_boundary() on the device, on the
 * initial state and on a perturbed copy (rarefied / cold / counter-streaming zones, so that the
 * floors, the mid-plane reset and the inflow switches of the user's code act).  Returns 1 if equal. */
static int shim_probe_ldw_boundary(Data *d, Grid *grid) {
  size_t n = (size_t)NVAR*NX3_TOT*NX2_TOT*NX1_TOT, q;
  double *vc = d->Vc[0][0][0];
  double *save = (double *) malloc (n*sizeof(double)), *dev = (double *) malloc (n*sizeof(double));
  int pass, ok = 1, i, j, k;
  memcpy (save, vc, n*sizeof(double));
  for (pass = 0; pass < 2 && ok; pass++) {
    if (pass == 1) DOM_LOOP(k,j,i) {
      if ((i + 2*j) % 3 == 0) d->Vc[RHO][k][j][i] *= 1.e-9;
#if HAVE_ENERGY
      if ((2*i + j) % 5 == 0) d->Vc[PRS][k][j][i] *= 1.e-9;
#endif
      if ((i + j) % 7 == 0)   d->Vc[VX1][k][j][i] = -d->Vc[VX1][k][j][i] - 0.3;
      if ((i + 3*j) % 4 == 0) d->Vc[VX2][k][j][i] = 0.7 - d->Vc[VX2][k][j][i];
    }
    if (pb200_upload_vc (s_ctx, vc) != PB200_OK || pb200_boundary (s_ctx) != PB200_OK ||
        pb200_download_vc (s_ctx, dev) != PB200_OK) { ok = 0; break; }
    Boundary (d, 0, grid);
    for (q = 0; q < n; q++) {
      double a = dev[q], b = vc[q], sc = fabs(b) > 1.e-300 ? fabs(b) : 1.e-300;
      if (!(fabs(a - b) <= 1.e-11*sc)) { ok = 0; break; }
    }
    memcpy (vc, save, n*sizeof(double));
  }
  free (save); free (dev);
  return ok;
}
#endif

static void shim_init(Data *d, Grid *grid) {
  int dir, ngpus = 1, ldw_device_bc = 0;
  s_resident = getenv("PB200_RESIDENT") && atoi(getenv("PB200_RESIDENT"));
  if (getenv("PB200_NGPUS")) ngpus = atoi(getenv("PB200_NGPUS"));
  if (ngpus < 1) ngpus = 1;
  /* USERDEF sides and INTERNAL_BOUNDARY are arbitrary host code: the reference's Boundary() per stage
     (the line-driven-wind problem has device versions, used only when they are verified below)   */
#if LINE_DRIVEN_WIND == NO
  for (dir = 0; dir < DIMENSIONS; dir++)
    if (grid->lbound[dir] == USERDEF || grid->rbound[dir] == USERDEF) s_host_bc = 1;
#if INTERNAL_BOUNDARY == YES
  s_host_bc = 1;
#endif
#else
  ldw_device_bc = 1;
  if (getenv("PB200_LDW_CVIDL_BC")) ldw_device_bc = atoi(getenv("PB200_LDW_CVIDL_BC")) != 0;
  if (!ldw_device_bc) s_host_bc = 1;
#endif
  if (getenv("PB200_HOST_BOUNDARY")) { s_host_bc = atoi(getenv("PB200_HOST_BOUNDARY")); if (s_host_bc) ldw_device_bc = 0; }
  if (ngpus > 1 && s_host_bc) {
    print ("! AdvanceStep(): PB200_NGPUS > 1 needs boundaries the device can fill (no USERDEF / INTERNAL_BOUNDARY)\n");
    QUIT_PLUTO(1);
  }
#if BODY_FORCE != NO
  shim_probe_body_force(d, grid);
#endif
  shim_create(d, grid, ngpus, ldw_device_bc);
#if LINE_DRIVEN_WIND != NO
  if (ldw_device_bc && getenv("PB200_LDW_CVIDL_BC") == NULL) {
    if (shim_probe_ldw_boundary(d, grid)) {
      print ("> AdvanceStep(): UserDefBoundary() verified against the device version (cv_idl)\n");
    } else {
      print ("> AdvanceStep(): UserDefBoundary() differs from the device version of cv_idl's:\n");
      print (">                boundaries are filled by the host's Boundary() every stage\n");
      pb200_destroy(s_ctx); s_ctx = NULL;
      s_host_bc = 1; ldw_device_bc = 0;
      shim_create(d, grid, 1, 0);
    }
  }
#endif
  s_ldw_device_bc = ldw_device_bc;
  print ("> AdvanceStep() runs on the GPU (libplutob200 v%d, %s mode%s", pb200_version(),
         s_resident ? "resident" : "strict host-buffer",
         s_host_bc ? ", boundaries by the host's Boundary()/UserDefBoundary()" : "");
  if (s_multi) print (", %d GPUs", pb200_multi_ngpus(s_multi));
  print (")\n");
}

/* FLAG_INTERNAL_BOUNDARY (Src/int_bound_reset.c:17-40, called from rhs.c:416-417): zones whose update is
 * frozen.  The reference re-reads d->flag every sweep; the host zeroes the flags at the top of every
 * step (main.c:258-261) and user code sets them in UserDefBoundary(side 0), i.e. inside Boundary().
 * In host-boundary mode the mask therefore travels to the device after every Boundary() call.    */
#if INTERNAL_BOUNDARY == YES
static unsigned char *s_ibmask = NULL;
static int shim_internal_boundary_mask(const Data *d)
{
  long ntot = (long)NX1_TOT*NX2_TOT*NX3_TOT, o = 0, any = 0;
  int i, j, k;
  if (s_ibmask == NULL) s_ibmask = (unsigned char *) malloc (ntot);
  TOT_LOOP(k,j,i) { s_ibmask[o] = (d->flag[k][j][i] & FLAG_INTERNAL_BOUNDARY) ? 1 : 0; any |= s_ibmask[o]; o++; }
  return pb200_set_internal_boundary_mask(s_ctx, any ? s_ibmask : NULL);
}
#endif

/* Host-boundary mode, stages > 1: user code that changes INTERIOR zones inside Boundary() converts them
 * itself (PrimToCons3D on 1-zone boxes, cv_idl/init.c:272-275); zones it changes without converting keep the
 * d->Uc of the previous stage in the reference.  Mark d->Uc with a NaN payload before Boundary(), collect the
 * zones whose mark is gone afterwards and hand exactly those to the device copy of d->Uc.            */
static void shim_mark_uc(Data *d)
{
  int i, j, k;
  union { double f; unsigned long long u; } mark;
  mark.u = 0x7ff8dead00b20000ull;
  DOM_LOOP(k,j,i) d->Uc[k][j][i][RHO] = mark.f;
}
static int shim_patch_uc(Data *d)
{
  static long *zone = NULL;
  static double *u = NULL;
  long n = 0, ntot = (long)NX1_TOT*NX2_TOT*NX3_TOT;
  int i, j, k, nv;
  union { double f; unsigned long long u; } q;
  if (zone == NULL) { zone = (long *) malloc (ntot*sizeof(long)); u = (double *) malloc (ntot*NVAR*sizeof(double)); }
  DOM_LOOP(k,j,i) {
    q.f = d->Uc[k][j][i][RHO];
    if (q.u == 0x7ff8dead00b20000ull) continue;
    zone[n] = ((long)k*NX2_TOT + j)*NX1_TOT + i;
    NVAR_LOOP(nv) u[n*NVAR + nv] = d->Uc[k][j][i][nv];
    n++;
  }
  return pb200_stage_patch_u(s_ctx, n, zone, u);
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
    if (s_resident) { if (s_multi) pb200_multi_upload_vc(s_multi, vc); else pb200_upload_vc(s_ctx, vc); }
  }
#if BODY_FORCE != NO
  if (s_bf_time_dependent) shim_body_force(d, grid);     /* g(t), Phi(t): tables for this step's g_time */
#endif
  if (s_host_bc) {
    /* per stage: D2H, the reference's Boundary() (-> UserDefBoundary) on d->Vc, H2D, stage */
    int s, ns = pb200_nstages(s_ctx);
    rc = pb200_step_begin(s_ctx, g_dt);
    for (s = 1; s <= ns && rc == PB200_OK; s++) {
      g_intStage = s;
      if (s > 1 || s_resident) rc = pb200_stage_download(s_ctx, s, vc);
      if (rc != PB200_OK) break;
      if (s > 1) shim_mark_uc(d);                 /* which zones of d->Uc does the user's Boundary() write? */
      Boundary (d, 0, grid);
      if (s > 1) { rc = shim_patch_uc(d); if (rc != PB200_OK) break; }
#if INTERNAL_BOUNDARY == YES
      rc = shim_internal_boundary_mask(d);
      if (rc != PB200_OK) break;
#endif
      rc = pb200_stage_upload(s_ctx, s, vc);
      if (rc == PB200_OK) rc = pb200_stage(s_ctx, s);
    }
    if (rc == PB200_OK) rc = pb200_step_end(s_ctx, &info);
    if (rc == PB200_OK) rc = pb200_download_vc(s_ctx, vc);
  } else if (s_multi) {
    if (s_resident) { rc = pb200_multi_advance_step(s_multi, g_dt, &info); s_dirty = 1; }
    else rc = pb200_multi_advance_step_host(s_multi, vc, g_dt, &info);
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
  /* update_stage.c:320 (1-D) / :389-392: invDt_hyp = MAX(invDt_hyp, max C_dt) / DIMENSIONS, where the
     old value survives from the previous step when COOLING != NO (main.c:406-415 resets Dts only every
     second step); info.invDt_hyp is max C_dt / DIMENSIONS already                               */
#if DIMENSIONS > 1
  Dts->invDt_hyp = MAX(Dts->invDt_hyp/(double)DIMENSIONS, info.invDt_hyp);
#else
  Dts->invDt_hyp = MAX(Dts->invDt_hyp, info.invDt_hyp);
#endif
  g_maxMach      = MAX(g_maxMach, info.maxMach);          /* hll_speed.c:89 */
  return 0;
}

/* resident mode: refresh d->Vc just before the reference reads it */
void pb200_shim_sync_to_host(const Data *d)
{
  if (s_ctx != NULL && s_resident && s_dirty) {
    if (s_multi) pb200_multi_download_vc(s_multi, d->Vc[0][0][0]);
    else pb200_download_vc(s_ctx, d->Vc[0][0][0]);
    s_dirty = 0;
  }
}
/* The source step (link with -Wl,--wrap=SplitSource).  Resident mode with COOLING BLONDIN and the
 * line-driven wind: BlondinCooling on the device copy.  Every other module (POWER_LAW, TABULATED, ...,
 * STS / RKL parabolic terms) is the reference's own host code: in resident mode the state is brought
 * down, the reference's SplitSource() runs on d->Vc, and the result goes back up, so the host copy is
 * never stale.  Host-buffer mode keeps the reference's SplitSource() on d->Vc as it is.           */
void __real_SplitSource (Data *, double, timeStep *, Grid *);
void __wrap_SplitSource (Data *d, double dt, timeStep *Dts, Grid *grid)
{
#if COOLING == BLONDIN && LINE_DRIVEN_WIND != NO
  if (s_ctx != NULL && s_resident && !s_host_bc) {
    if (pb200_split_source(s_ctx, dt, g_time) != PB200_OK) {
      print ("! SplitSource(): pb200_split_source failed\n");
      QUIT_PLUTO(1);
    }
    s_dirty = 1;
    return;
  }
#endif
#if COOLING != NO || (defined(PARABOLIC_FLUX) && PARABOLIC_FLUX != NO)
  if (s_ctx != NULL && s_resident) {
    pb200_shim_sync_to_host(d);
    __real_SplitSource(d, dt, Dts, grid);
    if (s_multi) pb200_multi_upload_vc(s_multi, d->Vc[0][0][0]); else pb200_upload_vc(s_ctx, d->Vc[0][0][0]);
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
