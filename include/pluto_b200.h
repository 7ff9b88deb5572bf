/* pluto_b200.h -- C ABI of libplutob200.so: the B200 (sm_100a) implementation of PLUTO's
 * unsplit finite-volume HD update as built in the sirocco-coupled fork.
 *
 * Plain C: opaque handle, plain pointers and sizes, no C++/torch types.  The reference has
 * no FFI layer; its seam for this path is the link-time symbol set.  Each
 * entry point below names the reference interface it stands in for.  INTEGRATION.md shows
 * the shim (pluto_sirocco_b200/csrc/pluto_shim.c) that maps the reference's
 * `int AdvanceStep(Data*, timeStep*, Grid*)` (Src/prototypes.h:5) onto these calls.
 *
 * All arrays are FP64.  Host-side state arrays use the reference layout of d->Vc:
 * Vc[nvar][NX3_TOT][NX2_TOT][NX1_TOT], i fastest, ghost zones included
 * (Src/arrays.c:251-330, Src/initialize.c:442).  Inactive dimensions have 1 zone, no ghosts.
 *
 * Error convention: functions return 0 on success, a negative PB200_E* code on failure
 * (pb200_last_error() gives text).  Recoverable cons->prim failures are floored on the
 * device exactly like Src/HD/mappers.c:139-218 and COUNTED (pb200_step_info.c2p_failures);
 * a NaN in the state makes the step return PB200_ENAN (reference: CheckNaN -> QUIT_PLUTO,
 * Src/Time_Stepping/update_stage.c:227).  There is no CPU fallback: without a CUDA device
 * pb200_create() fails with PB200_ENODEV.
 */
#ifndef PLUTO_B200_H
#define PLUTO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_VERSION 100

/* GEOMETRY: same values as Src/pluto.h:34-37.  The other option codes below are this
 * library's own; the shim translates the reference macros (pluto.h:66-70,254-287). */
#define PB200_CARTESIAN   1 /* pluto.h: CARTESIAN   */
#define PB200_CYLINDRICAL 2 /* pluto.h: CYLINDRICAL (r, z), 1-D / 2-D */
#define PB200_POLAR       3 /* pluto.h: POLAR (r, phi, z) */
#define PB200_SPHERICAL   4 /* pluto.h: SPHERICAL (r, theta, phi) */

#define PB200_FLAT      1   /* RECONSTRUCTION FLAT      */
#define PB200_LINEAR    2   /* RECONSTRUCTION LINEAR    (Src/States/plm_states.c)  */
#define PB200_PARABOLIC 3   /* RECONSTRUCTION PARABOLIC (Src/States/ppm_states.c, PPM_ORDER 4) */

#define PB200_EULER 1       /* TIME_STEPPING EULER */
#define PB200_RK2   2       /* TIME_STEPPING RK2  (Src/Time_Stepping/rk_step.c) */
#define PB200_RK3   3       /* TIME_STEPPING RK3 */

#define PB200_TVDLF 1       /* Solver tvdlf  -> LF_Solver   (Src/HD/tvdlf.c:38) */
#define PB200_HLL   2       /* Solver hll    -> HLL_Solver  (Src/HD/hll.c:30)   */
#define PB200_HLLC  3       /* Solver hllc   -> HLLC_Solver (Src/HD/hllc.c:28)  */
#define PB200_ROE   4       /* Solver roe    -> Roe_Solver  (Src/HD/roe.c:48; general path) */
#define PB200_TWO_SHOCK 5   /* Solver two_shock -> TwoShock_Solver (Src/HD/two_shock.c:28; EOS IDEAL, general path) */
#define PB200_AUSM  6       /* Solver ausm+  -> AUSMp_Solver (Src/HD/ausm.c:20; EOS IDEAL, general path) */

#define PB200_LIM_DEFAULT   0  /* LIMITER DEFAULT: MC rho, VL v, MM p (plm_states.c:202-244) */
#define PB200_LIM_FLAT      1
#define PB200_LIM_MINMOD    2
#define PB200_LIM_VANLEER   3
#define PB200_LIM_MC        4
#define PB200_LIM_VANALBADA 5
#define PB200_LIM_OSPRE     6
#define PB200_LIM_UMIST     7

/* boundary types: same values as Src/pluto.h:163-170 (OUTFLOW..USERDEF); PB200_BC_NEIGHBOUR = face owned by another
 * rank of the slab decomposition (ghosts arrive by halo exchange, not by a fill) */
#define PB200_BC_OUTFLOW      1
#define PB200_BC_REFLECTIVE   2
#define PB200_BC_AXISYMMETRIC 3
#define PB200_BC_EQTSYMMETRIC 4
#define PB200_BC_PERIODIC     5
#define PB200_BC_USERDEF      8
#define PB200_BC_POLARAXIS    9   /* PolarAxisBoundary(), Src/boundary.c:770-840 (general path) */
#define PB200_BC_NEIGHBOUR    100

/* BODY_FORCE bits: same values as Src/pluto.h (VECTOR 4, POTENTIAL 8) are NOT assumed; the
 * shim translates. */
#define PB200_EOS_IDEAL      0
#define PB200_EOS_ISOTHERMAL 1

#define PB200_BF_VECTOR    1
#define PB200_BF_POTENTIAL 2

#define PB200_OK        0
#define PB200_EINVAL   -1
#define PB200_ENODEV   -2
#define PB200_ECUDA    -3
#define PB200_ENOMEM   -4
#define PB200_ENAN     -5
#define PB200_ENOTSUP  -6

typedef struct pb200_ctx pb200_ctx;

/* Compile-time (definitions.h) and run-time (pluto.ini) options of one block of the grid. */
typedef struct pb200_config {
  int dimensions;        /* DIMENSIONS 1|2|3 */
  int geometry;          /* GEOMETRY */
  int nx[3];             /* interior zones of THIS block: NX1, NX2, NX3 */
  int nghost;            /* GetNghost(): 2 LINEAR, 3 PARABOLIC (Src/get_nghost.c:19) */
  int ntracer;           /* NTRACER */
  int reconstruction;    /* RECONSTRUCTION */
  int limiter;           /* LIMITER */
  int time_stepping;     /* TIME_STEPPING */
  int solver;            /* [Solver] Solver, Src/HD/set_solver.c:4-58 */
  int bc[6];             /* [Boundary] X1-beg, X1-end, X2-beg, ... */
  double gamma;          /* g_gamma */
  double small_density;  /* g_smallDensity  (Src/globals.h) */
  double small_pressure; /* g_smallPressure */
  double xbeg[3];        /* g_domBeg of this block */
  double xend[3];        /* g_domEnd of this block */
  int device;            /* CUDA device ordinal */
  int body_force;        /* BODY_FORCE: 0 NO, PB200_BF_VECTOR, PB200_BF_POTENTIAL or both (pluto.h:76-77) */
  int char_limiting;     /* CHAR_LIMITING YES (Src/States/plm_states.c:481) */
  int shock_flattening;  /* SHOCK_FLATTENING: 0 NO, 1 MULTID (Src/flag_shock.c:81), 2 ONED (Src/States/flatten.c:58; 4 ghost zones) */
  int entropy_switch;    /* ENTROPY_SWITCH: 0 NO, 1 SELECTIVE, 2 ALWAYS (same values as Src/pluto.h:60-61);
                            NVAR grows by ENTR (Src/entropy_switch.c, mappers.c:186-219, flag_shock.c:146,256) */
  int eos;               /* EOS: 0 IDEAL, PB200_EOS_ISOTHERMAL (Src/EOS/Isothermal: no energy equation, NVAR = 4 + NTRACER;
                            general path; e.g. Test_Problems/LineDrivenWind/cv_iso) */
  int ring_average;      /* RING_AVERAGE (Src/ring_average.c, pluto.h:479): chunk size at the axis, a power of two > 1; 0 / 1: off.
                            POLAR (axis at X1-beg) and SPHERICAL (axis at X2-beg / X2-end) with PB200_BC_POLARAXIS there
                            and a periodic phi direction; general path */
  int ring_average_rec;  /* RING_AVERAGE_REC 1 (none), 2 (van Leer), 5 (MP5, the reference's default when on; 3 ghost zones);
                            0: that default */
  double iso_sound_speed;/* g_isoSoundSpeed (EOS ISOTHERMAL) */
} pb200_config;

/* What one AdvanceStep leaves in timeStep / globals (Src/structs.h:372, globals.h). */
typedef struct pb200_step_info {
  double invDt_hyp;            /* Dts->invDt_hyp contribution of this step */
  double maxMach;              /* g_maxMach contribution */
  unsigned long long c2p_failures; /* zones floored by ConsToPrim */
  float  gpu_ms;               /* device time of the step (CUDA events) */
  int    launches;             /* kernels launched by the step */
} pb200_step_info;

const char *pb200_last_error(void);
int  pb200_version(void);

/* defaults matching Src/pluto.h / globals.h (g_gamma 5/3, small 1e-12, LIMITER DEFAULT) */
void pb200_config_default(pb200_config *cfg);

/* lifecycle: after Initialize() (Src/main.c:110) / before exit (Src/main.c:372) */
int  pb200_create(const pb200_config *cfg, pb200_ctx **out);
void pb200_destroy(pb200_ctx *ctx);

/* total zones per direction incl. ghosts (NX1_TOT, NX2_TOT, NX3_TOT) and NVAR */
int  pb200_shape(const pb200_ctx *ctx, int tot[3], int *nvar);

/* grid->xl / grid->xr / grid->dx of direction dir (np_tot entries each; dx may be NULL ->
 * xr-xl); default = uniform from xbeg/xend (Src/set_grid.c:405-412).  Needed before the
 * first step only for non-uniform grids. */
int  pb200_set_grid(pb200_ctx *ctx, int dir, const double *xl, const double *xr, const double *dx);

/* Geometry of the general (curvilinear / non-uniform) path.  By default the library evaluates the formulas of
 * Src/set_geometry.c:83-254 from xl / xr / dx itself (with the host's libm, bit-identical to the reference); a caller
 * that owns a Grid hands over the reference's own arrays instead, so that whatever the host set up is what the kernels
 * use.  All arrays in the reference's layout, ghost zones included:
 *   dV[NX3_TOT][NX2_TOT][NX1_TOT]                         grid->dV
 *   A[0] [NX3_TOT][NX2_TOT][NX1_TOT+1]  (first entry = index -1 along x1)   &grid->A[IDIR][0][0][-1]
 *   A[1] [NX3_TOT][NX2_TOT+1][NX1_TOT]  (first row   = index -1 along x2)   &grid->A[JDIR][0][-1][0]
 *   A[2] [NX3_TOT+1][NX2_TOT][NX1_TOT]  (first plane = index -1 along x3)   &grid->A[KDIR][-1][0][0]
 *   dx_dl[d][NX2_TOT][NX1_TOT]                            grid->dx_dl[d]
 *   rt[NX1_TOT], s[NX2_TOT], sp[NX2_TOT]                  grid->rt, grid->s, grid->sp
 * (Src/set_geometry.c:49-61).  Call before the first step. */
typedef struct pb200_geometry {
  const double *dV, *A[3], *dx_dl[3], *rt, *s, *sp;
} pb200_geometry;
int  pb200_set_geometry(pb200_ctx *ctx, const pb200_geometry *geo);
/* grid->uniform[d] (Src/set_grid.c:67-72, Src/runtime_setup.c:83-89): 1 when direction d is one uniform patch of
 * pluto.ini.  RECONSTRUCTION PARABOLIC picks its interface weights by it (PPM_CoefficientsSet, Src/States/
 * ppm_coeffs.c:88-136): closed forms on uniform directions, PPM_FindWeights otherwise.  Not called: a direction
 * counts as uniform when its dx is constant to round-off. */
int  pb200_set_grid_uniform(pb200_ctx *ctx, const int uniform[3]);
/* PPM_CoefficientsSet() of one direction (Src/States/ppm_coeffs.c:60-290), host only: interface weights w[ntot][4]
 * (v_{i+1/2} = sum_j w[i][j] v_{i-1+j}, entries 1 .. ntot-3) and the extremum coefficients h+ / h- [ntot]
 * (PPM_Q6_Coeffs) from the zone edges (and grid->dx; null: xr - xl) of the direction, the way the general path
 * evaluates them. */
int  pb200_ppm_coefficients(int geometry, int dir, int ntot, const double *xl, const double *xr, const double *dx,
                            int uniform, double *w, double *hp, double *hm);

/* BODY_FORCE tables.  The reference calls the user's BodyForceVector(v, g, x1, x2, x3) per zone
 * and sweep (Src/MHD/rhs_source.c:256,367) and BodyForcePotential(x1, x2, x3) at zone centres
 * and faces (Src/MHD/rhs.c:168-182, rhs_source.c:279,382).  Both are functions of position only
 * in every configuration on the path, so the caller evaluates them ONCE with the reference's
 * grid arrays and hands over strided tables:  value(i,j,k) = tab[i*si + j*sj + k*sk], indices
 * including ghost zones, a stride of 0 meaning "does not depend on that coordinate" (a constant
 * g is a 1-element table with si=sj=sk=0).  n = number of doubles in tab (copied to the device).
 *   vector:    comp 0..2 = g[IDIR], g[JDIR], g[KDIR] at (x1[i], x2[j], x3[k])
 *   potential: where 0 = Phi(x1[i],x2[j],x3[k]);  1,2,3 = Phi at the upper x1 / x2 / x3 face,
 *              i.e. Phi(x1p[i],x2[j],x3[k]), Phi(x1[i],x2p[j],x3[k]), Phi(x1[i],x2[j],x3p[k]).
 * Every table that cfg.body_force asks for must be set before the first step. */
int  pb200_set_body_force_vector(pb200_ctx *ctx, int comp, const double *tab, long n,
                                 long si, long sj, long sk);
int  pb200_set_body_force_potential(pb200_ctx *ctx, int where, const double *tab, long n,
                                    long si, long sj, long sk);

/* LINE_DRIVEN_WIND SIROCCO_MODE (Src/LineDriven/line_connect.c; definitions.h of
 * Test_Problems/LineDrivenWind/cv_idl).  The reference keeps the sirocco tables in globals
 * (flux_r_UV, flux_t_UV, flux_p_UV[NFLUX_ANGLES][k][j][i], Src/globals.h:197-217) filled by
 * read_sirocco_fluxes(); VGradCalc() runs once per UpdateStage (update_stage.c:139) and
 * LineForce() per zone and sweep (rhs_source.c:284-297,386-396).  pb200_ldw_enable() switches
 * both on for a general-path context, pb200_ldw_set_fluxes() uploads the tables (whole arrays
 * incl. ghost zones; call again whenever the reference re-reads them).  userdef_bc != 0 also
 * installs the device versions of that problem's UserDefBoundary(): the side == 0 floors and
 * mid-plane reset (cv_idl/init.c:199-316, INTERNAL_BOUNDARY YES) and the X1_BEG / X1_END / X2_BEG
 * fills (init.c:319-363) for every side whose type is PB200_BC_USERDEF. */
typedef struct pb200_ldw_config {
  int    nangles;                 /* NFLUX_ANGLES (<= 64) */
  int    userdef_bc;
  double unit_length, unit_velocity, unit_density;      /* UNIT_LENGTH, UNIT_VELOCITY, UNIT_DENSITY */
  double mu, krad, alpharad;      /* g_inputParam[MU], [KRAD], [ALPHARAD]; 999/999 selects the M(t) fit: pb200_ldw_set_mfit() */
  double dfloor, rho_0, rho_alpha, cent_mass, disk_mdot;   /* g_inputParam[DFLOOR], [RHO_0], [RHO_ALPHA], [CENT_MASS], [DISK_MDOT] (cgs) */
  double lx, tx;                  /* g_inputParam[L_star]*[f_x], g_inputParam[T_x] (BLONDIN cooling) */
  double t_iso;                   /* EOS ISOTHERMAL: g_inputParam[T_ISO], the temperature LineForce() uses (line_connect.c:851-855) */
} pb200_ldw_config;
int  pb200_ldw_enable(pb200_ctx *ctx, const pb200_ldw_config *ldw);
int  pb200_ldw_set_fluxes(pb200_ctx *ctx, const double *flux_r, const double *flux_t, const double *flux_p);
/* force-multiplier fit for KRAD = ALPHARAD = 999 (M_UV_data.dat, line_connect.c:185-256,849-853):
 * t_fit = log10(t) [mpoints], m_fit = log10(M) [mpoints][NX3_TOT][NX2_TOT][NX1_TOT] (the globals t_fit and
 * M_UV_fit, Src/globals.h:194-196) */
int  pb200_ldw_set_mfit(pb200_ctx *ctx, int mpoints, const double *t_fit, const double *m_fit);

/* COOLING BLONDIN.  pb200_split_source() is SplitSource(d, dt, Dts, grid) (Src/split_source.c:29,
 * called from Integrate, Src/main.c:479-485) for the BLONDIN module: BlondinCooling(d->Vc, d, dt)
 * (Src/Cooling/BLONDIN/cooling.c:50-170) on the device-resident d->Vc; g_time selects the analytic
 * ionisation parameter (g_time <= 3) or the sirocco tables.  pb200_cooling_set_tables() uploads the
 * per-zone tables of Data (Src/structs.h:621-643), each [NX3_TOT][NX2_TOT][NX1_TOT]:
 *   tabs[0..4] = comp_h_pre, comp_c_pre, xray_h_pre, line_c_pre, brem_c_pre   (NULL: 1.0, the
 *                defaults of read_sirocco_heatcool(), Src/LineDriven/line_connect.c:383-393)
 *   tabs[5..6] = sirocco_xi, sirocco_t_r   (NULL: analytic xi, T_x)
 * Needs pb200_ldw_enable() (units and L_x, T_x come from there). */
int  pb200_cooling_set_tables(pb200_ctx *ctx, const double *const tabs[7]);
int  pb200_split_source(pb200_ctx *ctx, double dt, double g_time);

/* INTERNAL_BOUNDARY YES: zones flagged FLAG_INTERNAL_BOUNDARY in d->flag keep a zero right-hand side in
 * every sweep (InternalBoundaryReset(), Src/int_bound_reset.c:17-40, called from Src/MHD/rhs.c:416-417; the
 * default INTERNAL_BOUNDARY_CFL YES leaves the signal speeds alone).  mask[NX3_TOT][NX2_TOT][NX1_TOT]: non-zero
 * = flagged; NULL = none.  The reference clears the flags every step (Src/main.c:258-261) and user code sets
 * them again inside Boundary(); call this whenever they change. */
int  pb200_set_internal_boundary_mask(pb200_ctx *ctx, const unsigned char *mask);

/* Test aid for the cooling module's arithmetic: the library's device versions of the C library's exp / log / log10 /
 * pow (GNU libc 2.39, x86-64 FMA variants - what BlondinCooling() of the reference build calls through libm) on host
 * arrays: which = 0 exp(x), 1 log(x), 2 log10(x), 3 pow(x, y).  Results equal the C library's bit for bit. */
int  pb200_libm_probe(int which, long n, const double *x, const double *y, double *out);

/* d->Vc  host -> device / device -> host (whole array incl. ghosts) */
int  pb200_upload_vc(pb200_ctx *ctx, const double *vc_host);
int  pb200_download_vc(pb200_ctx *ctx, double *vc_host);
/* device pointer of the current d->Vc mirror (for torch / NCCL interop, zero copy) */
double *pb200_device_vc(pb200_ctx *ctx);

/* Boundary(d, 0, grid) on the device mirror (Src/boundary.c:56).  On the Cartesian 2-D/3-D path the x1
 * ghost zones of standard sides (outflow, reflective-type, periodic) are never materialised: the sweep
 * kernels map them at load time, so ghost zones of a downloaded array are scratch. */
int  pb200_boundary(pb200_ctx *ctx);

/* AdvanceStep(d, Dts, grid) with g_dt = dt  (Src/Time_Stepping/rk_step.c:29) on the
 * device-resident state.  info may be NULL. */
int  pb200_advance_step(pb200_ctx *ctx, double dt, pb200_step_info *info);

/* Same call on HOST buffers (the strict drop-in: d->Vc is authoritative on the host):
 * H2D of vc_host, AdvanceStep, D2H back into vc_host.  For 3-D runs (any RK order) on the Cartesian path with
 * non-periodic x3 sides the three phases are pipelined over slabs of x3 planes (upload of slab
 * s+1, the RK stages on the slabs that are ready, download of finished slabs all overlap), so the
 * call costs about one PCIe direction; results are identical to pb200_advance_step().  Ghost zones
 * of vc_host are not written back.  PB200_HOST_PIPELINE=<planes per slab> (0: off). */
int  pb200_advance_step_host(pb200_ctx *ctx, double *vc_host, double dt, pb200_step_info *info);
/* Deep halo for host-buffer steps of a slab-decomposed grid (replaces the per-stage exchange of Src/boundary.c:139-158
 * when every step starts from host data anyway): a block is given nghost x nstages extra planes of its neighbours on
 * each cut face, steps WITHOUT any exchange (the cut faces get any fill type; what they contaminate never reaches the
 * owned planes within one step), and only its own planes [k0, k1) (interior x3 indices of this block) are downloaded
 * by pb200_advance_step_host() and counted in invDt_hyp / maxMach.  3-D Cartesian path. */
int  pb200_set_owned_planes(pb200_ctx *ctx, int k0, int k1);
/* cudaHostRegister / cudaHostUnregister of a caller-owned buffer (the reference's d->Vc payload):
 * page-locked memory makes the copies above asynchronous and full speed. */
int  pb200_host_register(void *ptr, size_t bytes);
int  pb200_host_unregister(void *ptr);

/* NextTimeStep() (Src/main.c:521-697) for COOLING NO, no parabolic terms.
 * Returns the new g_dt, or a negative value if dt < first_dt*1e-9 ("dt is too small"). */
double pb200_next_time_step(double invDt_hyp, double cfl, double cfl_max_var, double g_dt,
                            double first_dt);

/* main() loop body (Src/main.c:215-337) run nsteps times entirely from the device-resident
 * state: clip g_dt to tstop, AdvanceStep, g_time += g_dt, g_dt = NextTimeStep.
 * t and dt are in/out.  Returns the number of steps done (>=0) or a negative error. */
int  pb200_integrate(pb200_ctx *ctx, int nsteps, double tstop, double cfl, double cfl_max_var,
                     double first_dt, double *t, double *dt, pb200_step_info *last);

/* slab decomposition (replaces Src/Parallel/al_exchange_dim.c): device pointers/sizes of
 * the ghost and edge slabs of direction dir so the caller (NCCL send/recv) can move them.
 * lo_ghost/lo_edge/hi_edge/hi_ghost are offsets (in doubles) into variable 0 of the device
 * Vc of the array that the NEXT stage will sweep; count = doubles per variable slab;
 * var_stride = doubles between variables. */
int  pb200_halo_layout(const pb200_ctx *ctx, int dir, long *lo_ghost, long *lo_edge,
                       long *hi_edge, long *hi_ghost, long *count, long *var_stride);

/* Stage-level entry points for a caller that must act between stages (halo exchange,
 * UserDefBoundary): begin -> [stage(s): exchange halos of pb200_stage_array(), then
 * pb200_stage(s)] -> end.  pb200_advance_step == begin; for s in 1..nstages: stage(s); end. */
int  pb200_step_begin(pb200_ctx *ctx, double dt);
double *pb200_stage_array(pb200_ctx *ctx, int stage);  /* device Vc swept by stage s */
int  pb200_stage(pb200_ctx *ctx, int stage);
/* pb200_stage in three parts, for overlapping the halo exchange with compute (replaces the
 * blocking exchange inside Boundary(), Src/boundary.c:139-158):
 *   _boundary  fills the physical boundaries of the array stage s sweeps;
 *   _begin     runs the sweeps that do not read the ghost planes of the outermost active
 *              direction (3-D: the fused x1+x2 kernel; 1-D/2-D: nothing);
 *   _finish    runs the sweeps that do.
 * A slab-decomposed caller posts the exchange between _boundary and _begin and waits for it
 * before _finish.  pb200_stage(s) == _boundary; _begin; _finish. */
int  pb200_stage_boundary(pb200_ctx *ctx, int stage);
int  pb200_stage_begin(pb200_ctx *ctx, int stage);
int  pb200_stage_finish(pb200_ctx *ctx, int stage);
int  pb200_step_end(pb200_ctx *ctx, pb200_step_info *info);
/* Host round trip of the array stage s sweeps (whole d->Vc layout incl. ghosts), for callers whose
 * boundary conditions are arbitrary host code: the shim downloads it, runs the reference's own
 * Boundary(d, 0, grid) -> UserDefBoundary() (Src/boundary.c:56, Src/prototypes.h:217) on the host
 * copy and uploads it again before pb200_stage(s).  Slow by construction (PCIe per stage); it is
 * the correctness path for user plug-ins, not a CPU implementation of the update. */
int  pb200_stage_download(pb200_ctx *ctx, int stage, double *vc_host);
int  pb200_stage_upload(pb200_ctx *ctx, int stage, const double *vc_host);
/* Same mode, stages > 1: conservative states the caller's Boundary() wrote into d->Uc (user code that changes
 * interior zones converts them itself with PrimToCons3D on 1-zone boxes, e.g. cv_idl/init.c:272-275; zones it
 * changes WITHOUT converting keep the previous stage's d->Uc, and so they do here).  zone[n] = offsets
 * k*NX2_TOT*NX1_TOT + j*NX1_TOT + i, u[n][NVAR] in the reference's variable order.  The general path keeps
 * d->Uc on the device between stages; the Cartesian path keeps none (cons(V) is recomputed) and ignores this. */
int  pb200_stage_patch_u(pb200_ctx *ctx, long n, const long *zone, const double *u);
int  pb200_nstages(const pb200_ctx *ctx);
/* Measurement aid (the reference's FUNCTION_CLOCK_PROFILE, Src/pluto.h:414-419, rk_step.c:
 * 51-55): record CUDA events around every sweep kernel of the following steps;
 * pb200_kernel_times() returns, for the last finished step, the device time [ms], sweep
 * direction and RK stage of each sweep launch (return value = number of launches, <= max). */
int  pb200_set_profiling(pb200_ctx *ctx, int on);
int  pb200_kernel_times(const pb200_ctx *ctx, int max, float *ms, int *dir, int *stage);
/* CUDA stream (cudaStream_t as void*) all kernels of ctx are launched on */
void *pb200_stream(pb200_ctx *ctx);

/* ---- several GPUs of one box from ONE host thread -------------------------------------------
 * Replaces the reference's parallel layer for this path when the host stays the reference's own C
 * driver: the domain decomposition of Src/Parallel/al_decompose.c:40,125-158 becomes a 1-D slab split
 * along the outermost active direction, the ghost-zone exchange inside Boundary()
 * (Src/boundary.c:139-158 -> Src/Parallel/al_exchange_dim.c:64-90) one grouped ncclSend/ncclRecv batch
 * of packed edge planes per RK stage on a communication stream per device (overlapped with the fused
 * x1+x2 kernel), and the MPI_Allreduce(MAX) of Src/main.c:288,547 one ncclAllReduce(ncclMax) per step.
 * cfg describes the GLOBAL grid; devices = NULL means devices 0..ngpus-1.  All host arrays handed to
 * the pb200_multi_* calls are GLOBAL arrays in the reference's layout; results are bit-identical to
 * the single-GPU path.  NCCL (libnccl.so.2) is loaded at run time when ngpus > 1.  Cartesian
 * fast path only (the 2-D general-path problems are single-GPU: replicas only). */
typedef struct pb200_multi pb200_multi;
int  pb200_multi_create(const pb200_config *cfg, int ngpus, const int *devices, pb200_multi **out);
void pb200_multi_destroy(pb200_multi *m);
int  pb200_multi_ngpus(const pb200_multi *m);
/* the per-device context of a rank (for per-rank calls such as pb200_set_profiling) and the planes
 * [offset, offset + count) of the split direction it owns (interior index) */
pb200_ctx *pb200_multi_ctx(pb200_multi *m, int rank);
int  pb200_multi_slab(const pb200_multi *m, int rank, int *offset, int *count);
int  pb200_multi_set_grid(pb200_multi *m, int dir, const double *xl, const double *xr, const double *dx);
int  pb200_multi_set_body_force_vector(pb200_multi *m, int comp, const double *tab, long n,
                                       long si, long sj, long sk);
int  pb200_multi_set_body_force_potential(pb200_multi *m, int where, const double *tab, long n,
                                          long si, long sj, long sk);
int  pb200_multi_upload_vc(pb200_multi *m, const double *vc_host);
int  pb200_multi_download_vc(pb200_multi *m, double *vc_host);
int  pb200_multi_advance_step(pb200_multi *m, double dt, pb200_step_info *info);
int  pb200_multi_advance_step_host(pb200_multi *m, double *vc_host, double dt, pb200_step_info *info);
int  pb200_multi_integrate(pb200_multi *m, int nsteps, double tstop, double cfl, double cfl_max_var,
                           double first_dt, double *t, double *dt, pb200_step_info *last);

#ifdef __cplusplus
}
#endif
#endif
