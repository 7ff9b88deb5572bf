/* gen_oracle.c -- CPU restatement of the reference's unsplit HD update on GENERAL grids
 * (the checker for the curvilinear / line-driven-wind rows of SURVEY.md 8a: a4, a8-a14).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * build, link or call this file.  The product (pluto_sirocco_b200/, libplutob200.so) never does.
 *
 * Parity status: PINNED against the compiled, unmodified reference (oracle/_ref/<cfg>/pluto built
 * by oracle/build_ref.py with the user files of oracle/problems/) through the golden dumps of
 * tests/golden/ (tests/test_gen_oracle_golden.py).
 *
 * Scope: PHYSICS HD, EOS IDEAL or ISOTHERMAL, GEOMETRY CARTESIAN / CYLINDRICAL / POLAR / SPHERICAL,
 * DIMENSIONS 1-3, uniform or non-uniform grids (grid->xl/xr are inputs: the reference's own
 * set_grid.c output), RECONSTRUCTION LINEAR with every LIMITER and CHAR_LIMITING NO/YES, or
 * PARABOLIC (order 4, CHAR_LIMITING NO/YES) with the general-grid weights of ppm_coeffs.c,
 * SHOCK_FLATTENING NO / MULTID / ONED, ENTROPY_SWITCH NO / SELECTIVE / ALWAYS, NTRACER >= 0,
 * BODY_FORCE VECTOR / POTENTIAL, TIME_STEPPING EULER/RK2/RK3, Solver tvdlf / hll / hllc / roe / two_shock / ausm+,
 * outflow / reflective / axisymmetric / eqtsymmetric / periodic / polaraxis boundaries plus the
 * user-defined boundaries of the line-driven-wind problems (cv_idl, cv_iso), RING_AVERAGE
 * (ring_average.c, RING_AVERAGE_REC 1 / 2 / 5), LINE_DRIVEN_WIND (VGradCalc + LineForce, power law
 * or M(t) fit) and COOLING BLONDIN.
 * The CUDA path runs every one of these fixtures too (tests/test_gpu_gen.py).
 *
 * Plain C17, scalar, pencil by pencil like the reference so that the operation order is the
 * reference's; build with -ffp-contract=off.  Each function cites the reference file:line.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RHO 0
#define VX1 1
#define VX2 2
#define VX3 3
#define PRS 4
#define NFLX 5
#define NVMAX 10

#define FLAG_MINMOD 1      /* pluto.h:212-221 */
#define FLAG_FLAT 2
#define FLAG_HLL 4
#define FLAG_ENTROPY 8
#define FLAG_CONS2PRIM_FAIL 64

#define CARTESIAN 1
#define CYLINDRICAL 2
#define POLAR 3
#define SPHERICAL 4

#define MAXV(a, b) ((a) >= (b) ? (a) : (b))
#define MINV(a, b) ((a) <= (b) ? (a) : (b))
#define ABS_MIN(a, b) (fabs(a) < fabs(b) ? (a) : (b))
#define MINMOD_LIMITER(a, b) ((a) * (b) > 0.0 ? (fabs(a) < fabs(b) ? (a) : (b)) : 0.0)

typedef struct gen_cfg {
  int ndim, nx[3], ng;
  int ntracer, entropy;     /* NTRACER; ENTROPY_SWITCH: 0 NO, 1 SELECTIVE, 2 ALWAYS (pluto.h:60-61) */
  int geometry;             /* CARTESIAN 1, CYLINDRICAL 2, POLAR 3, SPHERICAL 4 (pluto.h:34-37) */
  int limiter;              /* 0 DEFAULT, 1 FLAT, 2 MINMOD, 3 VANLEER, 4 MC, 5 VANALBADA, 6 OSPRE, 7 UMIST */
  int char_limiting;        /* CHAR_LIMITING */
  int flattening;           /* SHOCK_FLATTENING MULTID */
  int rk, solver;           /* 1 EULER 2 RK2 3 RK3 ; 1 tvdlf 2 hll 3 hllc 4 roe 5 two_shock 6 ausm+ */
  int bc[6];                /* pluto.h:163-170; 8 userdef -> ldw_bc != 0 selects the built-in LDW fills */
  double gamma, small_dn, small_pr;
  const double *xl[3], *xr[3];   /* grid->xl, grid->xr incl. ghosts (np_tot each) */
  int body_force;           /* bit 0: VECTOR, bit 1: POTENTIAL (tables bf_phi below) */
  const double *bf_g[3];    /* BodyForceVector at zone centres, [k][j][i] incl. ghosts */
  /* --- line-driven wind (Src/LineDriven/line_connect.c), COOLING BLONDIN: see ldw section --- */
  int ldw;                  /* LINE_DRIVEN_WIND != NO */
  int ldw_bc;               /* user-defined boundaries / floors of Test_Problems/LineDrivenWind/cv_idl/init.c */
  int nangles;              /* NFLUX_ANGLES */
  const double *flux_r, *flux_t, *flux_p;   /* [nangles][k][j][i] */
  double unit_length, unit_velocity, unit_density;
  double mu, krad, alpharad, t_iso;
  double dfloor, rho0, rho_alpha, cent_mass, disk_mdot;   /* g_inputParam[] of the LDW problem */
  int cooling;              /* COOLING BLONDIN */
  const double *cool_tab[8]; /* comp_h_pre, comp_c_pre, xray_h_pre, line_c_pre, brem_c_pre, xi, T_r? (see ldw section) */
  double lx, tx;            /* L_star*f_x, T_x */
  int mpoints;              /* MPOINTS of the force-multiplier fit (KRAD = ALPHARAD = 999), line_connect.c:185-256 */
  const double *t_fit;      /* log10(t), MPOINTS entries */
  const double *m_fit;      /* log10(M), [MPOINTS][k][j][i] */
  /* --- EOS ISOTHERMAL (Src/EOS/Isothermal/eos.c): no energy equation, NFLX = 4, tracers from index 4 --- */
  int iso;                  /* 0: EOS IDEAL, 1: EOS ISOTHERMAL */
  double iso_cs;            /* g_isoSoundSpeed */
  int flatten_oned;         /* SHOCK_FLATTENING ONED (States/flatten.c); `flattening` above is MULTID */
  int ppm;                  /* RECONSTRUCTION PARABOLIC (PPM_ORDER 4) */
  int uniform[3];           /* grid->uniform[d] (set_grid.c:67-72): one uniform patch along d */
  const double *bf_phi[4];  /* body_force bit 1 (POTENTIAL): BodyForcePotential at zone centres [0] and at the
                               x1 / x2 / x3 upper faces [1..3], [k][j][i] incl. ghosts */
  /* --- RING_AVERAGE (Src/ring_average.c; Zhang et al. 2019): chunk size at the axis (> 1: on) and
         RING_AVERAGE_REC 1 / 2 / 5 (pluto.h:479-489); needs a POLARAXIS boundary (bc code 9) --- */
  int ring_average, ring_rec;
} gen_cfg;

#define NF(c) ((c)->iso ? 4 : NFLX)                     /* NFLX of the configuration (mod_defs.h) */
#define A2(c, v) ((c)->iso ? (c)->iso_cs * (c)->iso_cs : (c)->gamma * (v)[PRS] / (v)[RHO])   /* SoundSpeed2 */

typedef struct {
  int tot[3], beg[3], end[3], nvar;
  long sj, sk, sv;
  double *x[3], *xl[3], *xr[3], *dx[3], *xgc[3], *inv_dx[3];
  double *cp[3], *cm[3], *wp[3], *wm[3], *dp[3], *dm[3];   /* PLM_CoefficientsSet */
  double *rt, *s, *sp, *dmu;
  double *dV;          /* [k][j][i] */
  double *A[3];        /* A[d] with one extra layer at index -1 along d */
  long Aoff[3], Asj[3], Ask[3];
  double *dx_dl[3];    /* [j][i] */
  double (*pwp[3])[4];  /* PPM interface weights wp[i][-1..2] (PPM_CoefficientsSet, order 4) */
  double *php[3], *phm[3];   /* PPM_Q6_Coeffs */
  int *csize;          /* grid->ring_av_csize[]: per i (POLAR) or j (SPHERICAL), 0 outside the interior */
} geom_t;

/* ---------------------------------------------------------------------------------------
 *  Grid-derived arrays: set_grid.c:135-138 (cell centres), set_geometry.c:20-290,
 *  States/plm_coeffs.c:66-88
 * --------------------------------------------------------------------------------------- */
static double A_at(const geom_t *g, int d, int k, int j, int i) {
  return g->A[d][g->Aoff[d] + (long)k * g->Ask[d] + (long)j * g->Asj[d] + i];
}

static geom_t *geom_new(const gen_cfg *c) {
  geom_t *g = calloc(1, sizeof(geom_t));
  g->nvar = NF(c) + c->ntracer + (c->entropy ? 1 : 0);
  for (int d = 0; d < 3; d++) {
    int act = d < c->ndim;
    int ng = act ? c->ng : 0, nx = act ? c->nx[d] : 1;
    g->tot[d] = nx + 2 * ng;
    g->beg[d] = ng;
    g->end[d] = ng + nx - 1;
  }
  g->sj = g->tot[0];
  g->sk = (long)g->tot[0] * g->tot[1];
  g->sv = g->sk * g->tot[2];
  for (int d = 0; d < 3; d++) {
    int n = g->tot[d];
    g->x[d] = calloc(n, 8); g->xl[d] = calloc(n, 8); g->xr[d] = calloc(n, 8); g->dx[d] = calloc(n, 8);
    g->xgc[d] = calloc(n, 8); g->inv_dx[d] = calloc(n, 8);
    g->cp[d] = calloc(n, 8); g->cm[d] = calloc(n, 8); g->wp[d] = calloc(n, 8); g->wm[d] = calloc(n, 8);
    g->dp[d] = calloc(n, 8); g->dm[d] = calloc(n, 8);
    for (int i = 0; i < n; i++) {
      g->xl[d][i] = c->xl[d][i];
      g->xr[d][i] = c->xr[d][i];
      g->dx[d][i] = c->xr[d][i] - c->xl[d][i];
      g->x[d][i] = 0.5 * (c->xl[d][i] + c->xr[d][i]);     /* set_grid.c:137 */
    }
  }
  return g;
}

/* grid->dx is an INPUT of the reference's geometry (set_grid.c fills it together with xl/xr and
 * xr - xl is not always bit-identical to it), so the caller may override it. */
static void ppm_coeffs_set(const gen_cfg *c, geom_t *g);
static void ring_size(const gen_cfg *c, geom_t *g);
static void geom_finish(const gen_cfg *c, geom_t *g, const double *const dxin[3]) {
  int n1 = g->tot[0], n2 = g->tot[1], n3 = g->tot[2];
  for (int d = 0; d < 3; d++)
    if (dxin && dxin[d]) for (int i = 0; i < g->tot[d]; i++) g->dx[d][i] = dxin[d][i];
  double *x1 = g->x[0], *x2 = g->x[1], *dx1 = g->dx[0], *dx2 = g->dx[1], *dx3 = g->dx[2];
  double *x1p = g->xr[0], *x1m = g->xl[0], *x2p = g->xr[1], *x2m = g->xl[1];
  g->rt = calloc(n1, 8); g->s = calloc(n2, 8); g->sp = calloc(n2, 8); g->dmu = calloc(n2, 8);
  for (int i = 0; i < n1; i++) {   /* set_geometry.c:83-97 */
    double xL = x1m[i], xR = x1p[i];
    if (c->geometry == CARTESIAN) { g->xgc[0][i] = x1[i]; g->rt[i] = x1[i]; }
    else if (c->geometry == CYLINDRICAL || c->geometry == POLAR) {
      g->xgc[0][i] = x1[i] + dx1[i] * dx1[i] / (12.0 * x1[i]);
      g->rt[i] = x1[i];
    } else {
      g->xgc[0][i] = x1[i] + 2.0 * x1[i] * dx1[i] * dx1[i] / (12.0 * x1[i] * x1[i] + dx1[i] * dx1[i]);
      g->rt[i] = (xR * xR * xR - xL * xL * xL) / (xR * xR - xL * xL) / 1.5;
    }
  }
  for (int j = 0; j < n2; j++) {   /* set_geometry.c:103-116 */
    double xL = x2m[j], xR = x2p[j];
    if (c->geometry != SPHERICAL) g->xgc[1][j] = x2[j];
    else {
      g->xgc[1][j] = sin(xR) - sin(xL) + xL * cos(xL) - xR * cos(xR);
      g->xgc[1][j] /= cos(xL) - cos(xR);
      g->sp[j] = fabs(sin(xR));
      g->s[j] = fabs(sin(x2[j]));
      g->dmu[j] = fabs(cos(xL) - cos(xR));
    }
  }
  for (int k = 0; k < n3; k++) g->xgc[2][k] = g->x[2][k];
  /* volumes and areas: DIM_EXPAND keeps only the factors of the active dimensions */
  int nd = c->ndim;
  g->dV = calloc(g->sv, 8);
  for (int k = 0; k < n3; k++) for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) {
    double v;
    if (c->geometry == CARTESIAN) {
      v = dx1[i]; if (nd > 1) v = v * dx2[j]; if (nd > 2) v = v * dx3[k];
    } else if (c->geometry == CYLINDRICAL || c->geometry == POLAR) {   /* set_geometry.c:124-129 */
      double dVr = fabs(x1[i]) * dx1[i];
      v = dVr; if (nd > 1) v = v * dx2[j]; if (nd > 2) v = v * (c->geometry == POLAR ? dx3[k] : 1.0);
    } else {
      double dVr = fabs(x1p[i] * x1p[i] * x1p[i] - x1m[i] * x1m[i] * x1m[i]) / 3.0;
      double dmu = fabs(cos(x2m[j]) - cos(x2p[j]));
      v = dVr; if (nd > 1) v = v * dmu; if (nd > 2) v = v * dx3[k];
    }
    g->dV[k * g->sk + j * g->sj + i] = v;
  }
  for (int d = 0; d < 3; d++) {
    int e1 = n1 + (d == 0), e2 = n2 + (d == 1), e3 = n3 + (d == 2);
    g->Asj[d] = e1; g->Ask[d] = (long)e1 * e2;
    g->Aoff[d] = (d == 0) ? 1 : (d == 1 ? g->Asj[d] : g->Ask[d]);
    g->A[d] = calloc((long)e1 * e2 * e3, 8);
  }
  for (int k = 0; k < n3; k++) for (int j = 0; j < n2; j++) for (int i = -1; i < n1; i++) {
    double a;
    if (c->geometry == CARTESIAN) { a = 1.0; if (nd > 1) a = a * dx2[j]; if (nd > 2) a = a * dx3[k]; }
    else if (c->geometry == CYLINDRICAL || c->geometry == POLAR) {   /* set_geometry.c:150-161 */
      a = (i == -1) ? fabs(x1m[0]) : fabs(x1p[i]);
      if (nd > 1) a = a * dx2[j]; if (nd > 2) a = a * (c->geometry == POLAR ? dx3[k] : 1.0);
    } else {
      double dmu = fabs(cos(x2m[j]) - cos(x2p[j]));
      a = (i == -1) ? x1m[0] * x1m[0] : x1p[i] * x1p[i];
      if (nd > 1) a = a * dmu; if (nd > 2) a = a * dx3[k];
    }
    g->A[0][g->Aoff[0] + k * g->Ask[0] + j * g->Asj[0] + i] = a;
  }
  for (int k = 0; k < n3; k++) for (int j = -1; j < n2; j++) for (int i = 0; i < n1; i++) {
    double a;
    if (c->geometry == CARTESIAN || c->geometry == POLAR) { a = dx1[i]; if (nd > 1) a = a * 1.0; if (nd > 2) a = a * dx3[k]; }
    else if (c->geometry == CYLINDRICAL) { a = fabs(x1[i]); if (nd > 1) a = a * dx1[i]; if (nd > 2) a = a * 1.0; }   /* set_geometry.c:181-182 */
    else {
      a = fabs(x1[i]) * dx1[i];
      if (nd > 1) a = a * ((j == -1) ? fabs(sin(x2m[0])) : fabs(sin(x2p[j])));
      if (nd > 2) a = a * dx3[k];
    }
    g->A[1][g->Aoff[1] + k * g->Ask[1] + j * g->Asj[1] + i] = a;
  }
  for (int k = -1; k < n3; k++) for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) {
    double a;
    if (c->geometry == CARTESIAN) { a = dx1[i]; if (nd > 1) a = a * dx2[j]; if (nd > 2) a = a * 1.0; }
    else if (c->geometry == CYLINDRICAL) a = 1.0;   /* set_geometry.c:203-204 */
    else { a = fabs(x1[i]) * dx1[i]; if (nd > 1) a = a * dx2[j]; if (nd > 2) a = a * 1.0; }
    g->A[2][g->Aoff[2] + k * g->Ask[2] + j * g->Asj[2] + i] = a;
  }
  for (int d = 0; d < 3; d++) g->dx_dl[d] = calloc((long)n1 * n2, 8);
  for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) {   /* set_geometry.c:236-254 */
    long o = (long)j * n1 + i;
    g->dx_dl[0][o] = 1.0;
    if (c->geometry == CARTESIAN) { g->dx_dl[1][o] = 1.0; g->dx_dl[2][o] = 1.0; }
    else if (c->geometry == CYLINDRICAL) g->dx_dl[1][o] = 1.0;
    else if (c->geometry == POLAR) { g->dx_dl[1][o] = 1.0 / x1[i]; g->dx_dl[2][o] = 1.0; }
    else { g->dx_dl[1][o] = 1.0 / g->rt[i]; g->dx_dl[2][o] = dx2[j] / (g->rt[i] * g->dmu[j]); }
  }
  for (int d = 0; d < 3; d++) {
    for (int i = 0; i < g->tot[d]; i++) g->inv_dx[d][i] = 1.0 / g->dx[d][i];
    /* plm_coeffs.c:66-88 (first and last zone excluded) */
    double *dx = g->dx[d], *xgc = g->xgc[d], *xr = g->xr[d];
    for (int i = 1; i <= g->tot[d] - 2; i++) {
      g->wp[d][i] = dx[i] / (xgc[i + 1] - xgc[i]);
      g->wm[d][i] = dx[i] / (xgc[i] - xgc[i - 1]);
      g->cp[d][i] = (xgc[i + 1] - xgc[i]) / (xr[i] - xgc[i]);
      g->cm[d][i] = (xgc[i] - xgc[i - 1]) / (xgc[i] - xr[i - 1]);
      g->dp[d][i] = (xr[i] - xgc[i]) / dx[i];
      g->dm[d][i] = (xgc[i] - xr[i - 1]) / dx[i];
    }
  }
  if (c->ppm) ppm_coeffs_set(c, g);
  ring_size(c, g);
}

/* ---------------------------------------------------------------------------------------
 *  PPM (order 4) coefficients on general grids: States/ppm_coeffs.c:60-290 (PPM_CoefficientsSet),
 *  :300-420 (PPM_FindWeights: B.w = xi^k by LU decomposition), :520-570 (PPM_Q6_Coeffs),
 *  Math_Tools/math_lu_decomp.c (LUDecompose / LUBackSubst), Math_Tools/math_quadrature.c:30-78,356-409
 *  (5-point Gauss rule for the sin(theta) moments).  All four geometries.
 * --------------------------------------------------------------------------------------- */
#define POLY_2(a0, a1, a2, x) (a0 + x * (a1 + x * a2))
#define POLY_4(a0, a1, a2, a3, a4, x) (a0 + x * (a1 + x * (a2 + x * (a3 + x * a4))))
#define POLY_6(a0, a1, a2, a3, a4, a5, a6, x) (a0 + x * (a1 + x * (a2 + x * (a3 + x * (a4 + x * (a5 + x * a6))))))
/* GaussQuadrature(&BetaTheta, NULL, xb, xe, 1, 5): int_xb^xe (x - x0)^k sin(x) dx */
static double gauss5_beta_theta(double xb0, double xe0, double x0, int k) {
  double one_third = 1.0 / 3.0, ten_seventh = 10.0 / 7.0, z[5], w[5];
  z[0] = 0.0;
  z[1] = sqrt(5.0 - 2.0 * sqrt(ten_seventh)) * one_third;
  z[2] = -sqrt(5.0 - 2.0 * sqrt(ten_seventh)) * one_third;
  z[3] = sqrt(5.0 + 2.0 * sqrt(ten_seventh)) * one_third;
  z[4] = -sqrt(5.0 + 2.0 * sqrt(ten_seventh)) * one_third;
  w[0] = 128.0 / 225.0;
  w[1] = (322.0 + 13.0 * sqrt(70.0)) / 900.0;
  w[2] = (322.0 + 13.0 * sqrt(70.0)) / 900.0;
  w[3] = (322.0 - 13.0 * sqrt(70.0)) / 900.0;
  w[4] = (322.0 - 13.0 * sqrt(70.0)) / 900.0;
  double dx = (xe0 - xb0) / (double)1, I = 0.0;
  double xb = xb0 + 0 * dx, xe = xb + dx, Isub = 0.0;
  for (int n = 0; n < 5; n++) {
    double x = 0.5 * (xe - xb) * z[n] + (xe + xb) * 0.5;
    Isub += w[n] * (pow(x - x0, k) * sin(x));
  }
  Isub *= 0.5 * (xe - xb);
  I += Isub;
  return I;
}
static int lu_decompose(double a[8][8], int n, int *indx, double *d) {
  int imax = 0;
  double big, dum, sum, temp, vv[8];
  *d = 1.0;
  for (int i = 0; i < n; i++) {
    big = 0.0;
    for (int j = 0; j < n; j++) if ((temp = fabs(a[i][j])) > big) big = temp;
    if (big == 0.0) return 0;
    vv[i] = 1.0 / big;
  }
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < j; i++) {
      sum = a[i][j];
      for (int k = 0; k < i; k++) sum -= a[i][k] * a[k][j];
      a[i][j] = sum;
    }
    big = 0.0;
    for (int i = j; i < n; i++) {
      sum = a[i][j];
      for (int k = 0; k < j; k++) sum -= a[i][k] * a[k][j];
      a[i][j] = sum;
      if ((dum = vv[i] * fabs(sum)) >= big) { big = dum; imax = i; }
    }
    if (j != imax) {
      for (int k = 0; k < n; k++) { dum = a[imax][k]; a[imax][k] = a[j][k]; a[j][k] = dum; }
      *d = -(*d);
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (a[j][j] == 0.0) a[j][j] = 1.0e-20;
    if (j != n - 1) {
      dum = 1.0 / (a[j][j]);
      for (int i = j + 1; i < n; i++) a[i][j] *= dum;
    }
  }
  return 1;
}
static void lu_backsubst(double a[8][8], int n, const int *indx, double *b) {
  int ii = 0;
  double sum;
  for (int i = 0; i < n; i++) {
    int ip = indx[i];
    sum = b[ip];
    b[ip] = b[i];
    if (ii) for (int j = ii - 1; j <= i - 1; j++) sum -= a[i][j] * b[j];
    else if (sum) ii = i + 1;
    b[i] = sum;
  }
  for (int i = n - 1; i >= 0; i--) {
    sum = b[i];
    for (int j = i + 1; j < n; j++) sum -= a[i][j] * b[j];
    b[i] = sum / a[i][i];
  }
}
static void ppm_coeffs_set(const gen_cfg *c, geom_t *g) {
  const int iL = 1, iR = 2, n = 4;
  for (int d = 0; d < c->ndim; d++) {
    int nt = g->tot[d];
    g->pwp[d] = calloc(nt, sizeof(*g->pwp[d]));
    g->php[d] = calloc(nt, 8); g->phm[d] = calloc(nt, 8);
    const int radial = (d == 0 && (c->geometry == CYLINDRICAL || c->geometry == POLAR));
    const int sph_r = (d == 0 && c->geometry == SPHERICAL), sph_t = (d == 1 && c->geometry == SPHERICAL);
    for (int i = 0; i < nt; i++) {   /* PPM_Q6_Coeffs */
      g->php[d][i] = 3.0; g->phm[d][i] = 3.0;
      if (radial) {
        g->php[d][i] = 3.0 + 0.5 * g->dx[0][i] / g->x[0][i];
        g->phm[d][i] = 3.0 - 0.5 * g->dx[0][i] / g->x[0][i];
      } else if (sph_r) {
        double r = g->x[0][i], dr = g->dx[0][i];
        double den = 20.0 * r * r + dr * dr;
        g->php[d][i] = 3.0 + 2.0 * dr * (10.0 * r + dr) / den;
        g->phm[d][i] = 3.0 - 2.0 * dr * (10.0 * r - dr) / den;
      } else if (sph_t && i > 0) {   /* the reference reads thp[-1] for i = 0; that entry is never used */
        const double *thp = g->xr[1], *dth = g->dx[1];
        double cp = cos(thp[i]), sp = sin(thp[i]);
        double cm = cos(thp[i - 1]), sm = sin(thp[i - 1]);
        double dmu = cm - cp;
        double dmu_t = sm - sp;
        g->php[d][i] = dth[i] * (dmu_t + dth[i] * cp) / (dth[i] * (sp + sm) - 2.0 * dmu);
        g->phm[d][i] = -dth[i] * (dmu_t + dth[i] * cm) / (dth[i] * (sp + sm) - 2.0 * dmu);
      }
    }
    int beg = iL, end = nt - 1 - iR;
    if (!c->uniform[d] || sph_t) {   /* PPM_FindWeights (the meridional direction always, ppm_coeffs.c:268-269) */
      for (int i = beg; i <= end; i++) {
        double beta[8][8], a[16], dd;
        int indx[16];
        double rc = g->x[d][i];
        int jb = i - iL, je = i + iR;
        for (int j = jb; j <= je; j++) {
          double rp = g->xr[d][j], rm = g->xl[d][j], vol;
          if (sph_t) {
            vol = cos(rm) - cos(rp);
            for (int k = 0; k < n; k++) beta[k][j - jb] = gauss5_beta_theta(rm, rp, rc, k);
            beta[0][j - jb] /= vol; beta[1][j - jb] /= vol; beta[2][j - jb] /= vol; beta[3][j - jb] /= vol;
          } else if (sph_r) {
            vol = (rp * rp * rp - rm * rm * rm) / 3.0;
            for (int k = 0; k < n; k++) {
              beta[k][j - jb] = pow(rp - rc, k + 1) * ((k * k + 3.0 * k + 2.0) * rp * rp + 2.0 * rc * (k + 1.0) * rp + 2.0 * rc * rc)
                              - pow(rm - rc, k + 1) * ((k * k + 3.0 * k + 2.0) * rm * rm + 2.0 * rc * (k + 1.0) * rm + 2.0 * rc * rc);
              beta[k][j - jb] /= (k + 3.0) * (k + 2.0) * (k + 1.0) * vol;
            }
          } else if (!radial) {
            vol = (rp - rm);
            for (int k = 0; k < n; k++) beta[k][j - jb] = (pow(rp - rc, k + 1) - pow(rm - rc, k + 1)) / (k + 1.0) / vol;
          } else {
            vol = (rp * rp - rm * rm) / 2.0;
            for (int k = 0; k < n; k++) {
              beta[k][j - jb] = pow(rp - rc, k + 1) * ((k + 1.0) * rp + rc) - pow(rm - rc, k + 1) * ((k + 1.0) * rm + rc);
              beta[k][j - jb] /= (k + 2.0) * (k + 1.0) * vol;
            }
          }
        }
        lu_decompose(beta, n, indx, &dd);
        double rp = g->xr[d][i];
        a[0] = 1.0;
        for (int k = 1; k < n; k++) a[k] = a[k - 1] * (rp - rc);
        lu_backsubst(beta, n, indx, a);
        for (int j = 0; j < n; j++) g->pwp[d][i][j] = a[j];
      }
      continue;
    }
    for (int i = beg; i <= end; i++) {   /* PPM_CartCoeff */
      g->pwp[d][i][0] = -1.0 / 12.0; g->pwp[d][i][1] = 7.0 / 12.0;
      g->pwp[d][i][2] = 7.0 / 12.0;  g->pwp[d][i][3] = -1.0 / 12.0;
    }
    if (radial) {                        /* ppm_coeffs.c:150-165 */
      for (int i = beg; i <= end; i++) {
        double rp = g->xr[0][i], dr = g->dx[0][i];
        double i1 = rp / dr, i2 = i1 * i1;
        double den = 24.0 * POLY_2(4.0, -15.0, 5.0, i2);
        g->pwp[d][i][0] = POLY_4(-12.0, -1.0, 30.0, -1.0, -10.0, i1) / den;
        g->pwp[d][i][1] = POLY_4(60.0, -27.0, -210.0, 13.0, 70.0, i1) / den;
        g->pwp[d][i][2] = POLY_4(60.0, 27.0, -210.0, -13.0, 70.0, i1) / den;
        g->pwp[d][i][3] = POLY_4(-12.0, 1.0, 30.0, 1.0, -10.0, i1) / den;
      }
    }
    if (sph_r) {                         /* ppm_coeffs.c:216-229 */
      for (int i = beg; i <= end; i++) {
        double rp = g->xr[0][i], dr = g->dx[0][i];
        double i1 = fabs(rp / dr), i2 = i1 * i1;
        double den = 36.0 * POLY_4(16.0, -60.0, 150.0, -85.0, 15.0, i2);
        g->pwp[d][i][0] = -POLY_2(7, -9, 3, i1) / den * POLY_6(12, 16, -30, -48.0, 23, 48, 15, i1);
        g->pwp[d][i][1] = POLY_2(1, -3, 3, i1) / den * POLY_6(372, 1008.0, 510, -720, -487, 144, 105, i1);
        g->pwp[d][i][2] = POLY_2(1, 3, 3.0, i1) / den * POLY_6(372, -1008, 510, 720, -487, -144, 105, i1);
        g->pwp[d][i][3] = -POLY_2(7, 9, 3, i1) / den * POLY_6(12, -16, -30, 48, 23, -48, 15, i1);
      }
    }
  }
}

static void geom_free(geom_t *g) {
  for (int d = 0; d < 3; d++) {
    free(g->x[d]); free(g->xl[d]); free(g->xr[d]); free(g->dx[d]); free(g->xgc[d]); free(g->inv_dx[d]);
    free(g->cp[d]); free(g->cm[d]); free(g->wp[d]); free(g->wm[d]); free(g->dp[d]); free(g->dm[d]);
    free(g->A[d]); free(g->dx_dl[d]); free(g->pwp[d]); free(g->php[d]); free(g->phm[d]);
  }
  free(g->rt); free(g->s); free(g->sp); free(g->dmu); free(g->dV); free(g->csize);
  free(g);
}

/* ---------------------------------------------------------------------------------------
 *  Limiters: States/plm_coeffs.h:72-152.  uniform = UNIFORM_CARTESIAN_GRID (GEOMETRY CARTESIAN)
 * --------------------------------------------------------------------------------------- */
static double lim_apply(int kind, int uniform, double dvp, double dvm, double cp, double cm) {
  double dv = 0.0;
  switch (kind) {
    case 1: dv = 0.0; break;
    case 2: dv = (dvp * dvm > 0.0 ? ABS_MIN(dvp, dvm) : 0.0); break;
    case 3:
      if (uniform) dv = (dvp * dvm > 0.0 ? 2.0 * dvp * dvm / (dvp + dvm) : 0.0);
      else dv = (dvp * dvm > 0.0 ? (dvp) * (dvm) * (cp * (dvm) + cm * (dvp)) /
                 ((dvp) * (dvp) + (dvm) * (dvm) + (cp + cm - 2.0) * (dvp) * (dvm)) : 0.0);
      break;
    case 4:
      if (dvp * dvm > 0.0) {
        double qc = 0.5 * (dvm + dvp);
        double scrh = uniform ? 2.0 * ABS_MIN(dvp, dvm) : ABS_MIN((dvp) * cp, (dvm) * cm);
        dv = ABS_MIN(qc, scrh);
      }
      break;
    case 5:
      if (dvp * dvm > 0.0) {
        double dpp = dvp * dvp, dmm = dvm * dvm;
        dv = (dvp * (dmm + 1.e-18) + dvm * (dpp + 1.e-18)) / (dpp + dmm + 1.e-18);
      }
      break;
    case 6:
      if (uniform) dv = (dvp * dvm > 0.0 ? 1.5 * dvp * dvm * (dvm + dvp) / (dvp * dvp + dvm * dvm + dvp * dvm) : 0.0);
      else if (dvp * dvm > 0.0) {
        double den = 2.0 * (dvp) * (dvp) + 2.0 * (dvm) * (dvm) + (cp + cm - 2.0) * (dvp) * (dvm);
        dv = dvp * dvm * ((1.0 + cp) * (dvm) + (1.0 + cm) * (dvp)) / den;
      }
      break;
    case 7:
      if (dvp * dvm > 0.0) {
        double ddp = 0.25 * (dvp + 3.0 * dvm), ddm = 0.25 * (dvm + 3.0 * dvp);
        double d2 = 2.0 * ABS_MIN(dvp, dvm);
        d2 = ABS_MIN(d2, ddp);
        dv = ABS_MIN(d2, ddm);
      }
      break;
    case 8:   /* SET_GM_LIMITER, plm_coeffs.h:96-100 */
      if (dvp * dvm > 0.0) {
        double qc = 0.5 * (dvm + dvp), scrh = ABS_MIN((dvp) * (cp), (dvm) * (cm));
        dv = ABS_MIN(qc, scrh);
      }
      break;
  }
  return dv;
}

/* HD/mappers.c:26-56 (+ scalars incl. ENTR: u = rho*v) */
static void prim_to_cons(const gen_cfg *c, int nvar, const double *v, double *u) {
  double gmm1 = c->gamma - 1.0, rho = v[RHO];
  u[RHO] = rho;
  u[VX1] = rho * v[VX1];
  u[VX2] = rho * v[VX2];
  u[VX3] = rho * v[VX3];
  if (!c->iso) {
    u[PRS] = v[VX1] * v[VX1] + v[VX2] * v[VX2] + v[VX3] * v[VX3];
    u[PRS] = 0.5 * rho * u[PRS] + v[PRS] / gmm1;
  }
  for (int nv = NF(c); nv < nvar; nv++) u[nv] = rho * v[nv];
}

/* HD/mappers.c:98-290 with ENTROPY_SWITCH */
static int cons_to_prim(const gen_cfg *c, int nvar, double *u, double *v, uint16_t *flag) {
  int fail = 0;
  double gmm1 = c->gamma - 1.0;
  int ENTR = nvar - 1;
  double m2 = u[VX1] * u[VX1] + u[VX2] * u[VX2] + u[VX3] * u[VX3];
  if (u[RHO] < 0.0) { u[RHO] = c->small_dn; *flag |= FLAG_CONS2PRIM_FAIL; fail = 1; }
  double rho = v[RHO] = u[RHO];
  double tau = 1.0 / u[RHO];
  v[VX1] = u[VX1] * tau;
  v[VX2] = u[VX2] * tau;
  v[VX3] = u[VX3] * tau;
  double kin = 0.5 * m2 / u[RHO];
  if (c->iso) {   /* mappers.c: no energy, no pressure */
    for (int nv = NF(c); nv < nvar; nv++) v[nv] = u[nv] * tau;
    return fail;
  }
  if (u[PRS] < 0.0) { u[PRS] = c->small_pr / gmm1 + kin; *flag |= FLAG_CONS2PRIM_FAIL; fail = 1; }
  int use_entropy = c->entropy && (*flag & FLAG_ENTROPY);
  if (use_entropy) {
    double rhog1 = pow(rho, gmm1);
    v[PRS] = u[ENTR] * rhog1;
    if (v[PRS] < 0.0) { v[PRS] = c->small_pr; *flag |= FLAG_CONS2PRIM_FAIL; fail = 1; }
    u[PRS] = v[PRS] / gmm1 + kin;
  } else {
    v[PRS] = gmm1 * (u[PRS] - kin);
    if (v[PRS] < 0.0) {
      v[PRS] = c->small_pr;
      u[PRS] = v[PRS] / gmm1 + kin;
      *flag |= FLAG_CONS2PRIM_FAIL;
      fail = 1;
    }
    if (c->entropy) u[ENTR] = v[PRS] / pow(rho, gmm1);
  }
  for (int nv = NF(c); nv < nvar; nv++) v[nv] = u[nv] * tau;
  return fail;
}

/* 1-D work arrays of one pencil (Sweep, tools.c:261-335) */
typedef struct {
  double (*v)[NVMAX], (*vp)[NVMAX], (*vm)[NVMAX], (*flux)[NVMAX], (*rhs)[NVMAX], (*fA)[NVMAX];
  double *press, *cmax;
  uint16_t *flag;
} sweep_t;

static void sweep_alloc(sweep_t *s, int n) {
  s->v = calloc(n + 4, sizeof(*s->v)); s->vp = calloc(n + 4, sizeof(*s->vp)); s->vm = calloc(n + 4, sizeof(*s->vm));
  s->flux = calloc(n + 4, sizeof(*s->flux)); s->rhs = calloc(n + 4, sizeof(*s->rhs)); s->fA = calloc(n + 4, sizeof(*s->fA));
  s->press = calloc(n + 4, 8); s->cmax = calloc(n + 4, 8); s->flag = calloc(n + 4, sizeof(uint16_t));
  s->fA += 1;     /* fA[-1] is used by the curvilinear rhs */
  s->flux += 1; s->press += 1; s->cmax += 1;
}
static void sweep_free(sweep_t *s) {
  free(s->v); free(s->vp); free(s->vm); free(s->flux - 1); free(s->rhs); free(s->fA - 1);
  free(s->press - 1); free(s->cmax - 1); free(s->flag);
}

/* Flatten(), States/flatten.c:58-130 (HD: EPS2 0.33, OME1 0.75, OME2 10) */
static void flatten_oned(const gen_cfg *c, const geom_t *g, sweep_t *s, int dir, int beg, int end) {
  int nvar = g->nvar, VXn = 1 + dir;
  {
    const int P = c->iso ? RHO : PRS;
    int fb = MAXV(beg, 3), fe = MINV(end, g->tot[dir] - 4);
    double *f_t = calloc(g->tot[dir] + 4, 8);
    for (int i = fb - 1; i <= fe + 1; i++) {
      double dp = s->v[i + 1][P] - s->v[i - 1][P];
      double min_p = MINV(s->v[i + 1][P], s->v[i - 1][P]);
      double d2p = s->v[i + 2][P] - s->v[i - 2][P];
      double scrh = fabs(dp) / min_p;
      if (scrh < 0.33 || (s->v[i + 1][VXn] > s->v[i - 1][VXn])) f_t[i] = 0.0;
      else {
        scrh = 10.0 * (fabs(dp / d2p) - 0.75);
        scrh = MINV(1.0, scrh);
        f_t[i] = MAXV(0.0, scrh);
      }
    }
    for (int i = fb; i <= fe; i++) {
      int sj = (s->v[i + 1][P] < s->v[i - 1][P] ? 1 : -1);
      double fj = MAXV(f_t[i], f_t[i + sj]);
      for (int nv = 0; nv < nvar; nv++) {
        double vf = s->v[i][nv] * fj, scrh = 1.0 - fj;
        s->vm[i][nv] = vf + s->vm[i][nv] * scrh;
        s->vp[i][nv] = vf + s->vp[i][nv] * scrh;
      }
    }
    free(f_t);
  }
}

/* States/plm_states.c:83-337 (CHAR_LIMITING NO) and :481-690 (CHAR_LIMITING YES) */
static void states(const gen_cfg *c, const geom_t *g, sweep_t *s, int dir, int beg, int end) {
  int nvar = g->nvar;
  int uniform = (c->geometry == CARTESIAN);   /* plm_coeffs.h:23-29 */
  int VXn = 1 + dir, VXt = 1 + (dir + 1) % 3, VXb = 1 + (dir + 2) % 3;
  double dvp[NVMAX], dvm[NVMAX], dv_lim[NVMAX];
  for (int i = beg; i <= end; i++) {
    double cp, cm, wp, wm, dp, dm;
    const double *v = s->v[i];
    if (uniform) {
      cp = cm = 2.0; dp = dm = 0.5; wp = wm = 1.0;
      for (int nv = 0; nv < nvar; nv++) { dvp[nv] = s->v[i + 1][nv] - v[nv]; dvm[nv] = v[nv] - s->v[i - 1][nv]; }
    } else {
      cp = g->cp[dir][i]; cm = g->cm[dir][i]; wp = g->wp[dir][i]; wm = g->wm[dir][i];
      dp = g->dp[dir][i]; dm = g->dm[dir][i];
      for (int nv = 0; nv < nvar; nv++) {
        dvp[nv] = (s->v[i + 1][nv] - v[nv]) * wp;
        dvm[nv] = (v[nv] - s->v[i - 1][nv]) * wm;
      }
    }
    if (!c->char_limiting) {
      if (c->flattening) {   /* plm_states.c:177-195 */
        if (s->flag[i] & FLAG_FLAT) {
          for (int nv = 0; nv < nvar; nv++) s->vp[i][nv] = s->vm[i][nv] = v[nv];
          continue;
        } else if (s->flag[i] & FLAG_MINMOD) {
          for (int nv = 0; nv < nvar; nv++) {
            dv_lim[nv] = lim_apply(2, uniform, dvp[nv], dvm[nv], cp, cm);
            s->vp[i][nv] = v[nv] + dv_lim[nv] * dp;
            s->vm[i][nv] = v[nv] - dv_lim[nv] * dm;
          }
          continue;
        }
      }
      for (int nv = 0; nv < nvar; nv++) {
        int kind;
        if (c->limiter == 0) {
          if (nv == RHO) kind = 4; else if (!c->iso && nv == PRS) kind = 2; else if (nv >= NF(c)) kind = 4; else kind = 3;
        } else kind = c->limiter;
        dv_lim[nv] = lim_apply(kind, uniform, dvp[nv], dvm[nv], cp, cm);
      }
      for (int nv = 0; nv < nvar; nv++) {
        s->vp[i][nv] = v[nv] + dv_lim[nv] * dp;
        s->vm[i][nv] = v[nv] - dv_lim[nv] * dm;
      }
    } else {
      /* SoundSpeed2 (eos.c:33), PrimEigenvectors (eigenv.c:92-200), PrimToChar (eigenv.c:575-616) */
      const int nf = NF(c);
      double a2 = A2(c, v);
      double cs = sqrt(a2), rhocs = v[RHO] * cs, rho_cs = v[RHO] / cs;
      double R[NFLX][NFLX], kstp[NFLX], cpk[NFLX], cmk[NFLX], dwp[NFLX], dwm[NFLX], dw_lim[NFLX];
      memset(R, 0, sizeof(R));
      R[RHO][0] = 0.5 * rho_cs; R[VXn][0] = -0.5;
      R[RHO][1] = 0.5 * rho_cs; R[VXn][1] = 0.5;
      if (!c->iso) { R[PRS][0] = 0.5 * rhocs; R[PRS][1] = 0.5 * rhocs; R[RHO][2] = 1.0; R[VXt][3] = 1.0; R[VXb][4] = 1.0; }
      else { R[VXt][2] = 1.0; R[VXb][3] = 1.0; }   /* eigenv.c:175-178 */
      double L0n = -1.0, L0p = 1.0 / rhocs, L1n = 1.0, L1p = 1.0 / rhocs, L2p = -1.0 / a2;
      for (int k = 0; k < NFLX; k++) kstp[k] = 2.0;
      kstp[0] = kstp[1] = 1.0;
      for (int k = 0; k < NFLX; k++) {
        if (uniform) cpk[k] = cmk[k] = kstp[k];
        else { cpk[k] = (2.0 - cp) + (cp - 1.0) * kstp[k]; cmk[k] = (2.0 - cm) + (cm - 1.0) * kstp[k]; }
      }
      if (!c->iso) {
        dwm[0] = L0n * dvm[VXn] + L0p * dvm[PRS]; dwm[1] = L1n * dvm[VXn] + L1p * dvm[PRS];
        dwm[2] = dvm[RHO] + L2p * dvm[PRS]; dwm[3] = dvm[VXt]; dwm[4] = dvm[VXb];
        dwp[0] = L0n * dvp[VXn] + L0p * dvp[PRS]; dwp[1] = L1n * dvp[VXn] + L1p * dvp[PRS];
        dwp[2] = dvp[RHO] + L2p * dvp[PRS]; dwp[3] = dvp[VXt]; dwp[4] = dvp[VXb];
      } else {   /* eigenv.c:182-196 (LL[0][RHO] = LL[1][RHO] = 1/rho_cs), PrimToChar eigenv.c:600-605 */
        double Lr = 1.0 / rho_cs;
        dwm[0] = Lr * dvm[RHO] + L0n * dvm[VXn]; dwm[1] = Lr * dvm[RHO] + L1n * dvm[VXn];
        dwm[2] = dvm[VXt]; dwm[3] = dvm[VXb];
        dwp[0] = Lr * dvp[RHO] + L0n * dvp[VXn]; dwp[1] = Lr * dvp[RHO] + L1n * dvp[VXn];
        dwp[2] = dvp[VXt]; dwp[3] = dvp[VXb];
      }
      if (c->flattening && (s->flag[i] & FLAG_FLAT)) {
        for (int k = nf; k--;) dw_lim[k] = 0.0;
      } else if (c->flattening && (s->flag[i] & FLAG_MINMOD)) {
        for (int k = nf; k--;) dw_lim[k] = lim_apply(2, uniform, dwp[k], dwm[k], cp, cm);
      } else {
        for (int k = nf; k--;) {
          if (c->limiter == 0) dw_lim[k] = lim_apply(8, uniform, dwp[k], dwm[k], cpk[k], cmk[k]);
          else dw_lim[k] = lim_apply(c->limiter, uniform, dwp[k], dwm[k], cp, cm);
        }
      }
      for (int nv = nf; nv--;) {
        double dc = 0.0;
        for (int k = 0; k < nf; k++) dc += dw_lim[k] * R[nv][k];
        if (dvp[nv] * dvm[nv] > 0.0) {
          double d2v = ABS_MIN(cp * dvp[nv], cm * dvm[nv]);
          dv_lim[nv] = MINMOD_LIMITER(d2v, dc);
        } else dv_lim[nv] = 0.0;
      }
      for (int nv = nf; nv < nvar; nv++)
        dv_lim[nv] = lim_apply(c->limiter == 0 ? 4 : c->limiter, uniform, dvp[nv], dvm[nv], cp, cm);
      for (int nv = nvar; nv--;) {
        s->vp[i][nv] = v[nv] + dv_lim[nv] * dp;
        s->vm[i][nv] = v[nv] - dv_lim[nv] * dm;
      }
    }
  }
  if (c->flatten_oned) flatten_oned(c, g, s, dir, beg, end);
}

/* States/ppm_states.c:66-232 (CHAR_LIMITING NO, PPM_ORDER 4) */
static void states_ppm(const gen_cfg *c, const geom_t *g, sweep_t *s, int dir, int beg, int end) {
  int nvar = g->nvar;
  for (int i = beg - 1; i <= end; i++) {   /* unique interface value, clipped between the cell averages */
    const double *wp = g->pwp[dir][i];
    for (int nv = 0; nv < nvar; nv++) {
      s->vp[i][nv] = wp[0] * s->v[i - 1][nv] + wp[1] * s->v[i][nv] + wp[2] * s->v[i + 1][nv] + wp[3] * s->v[i + 2][nv];
      double dv = s->v[i + 1][nv] - s->v[i][nv];
      double dvp = s->vp[i][nv] - s->v[i][nv];
      s->vp[i][nv] = s->v[i][nv] + MINMOD_LIMITER(dvp, dv);
    }
  }
  for (int i = beg; i <= end; i++) for (int nv = 0; nv < nvar; nv++) s->vm[i][nv] = s->vp[i - 1][nv];
  for (int i = beg; i <= end; i++) {
    const double *v = s->v[i];
    if (c->flattening) {
      if (s->flag[i] & FLAG_FLAT) {
        for (int nv = 0; nv < nvar; nv++) s->vp[i][nv] = s->vm[i][nv] = v[nv];
        continue;
      } else if (s->flag[i] & FLAG_MINMOD) {
        for (int nv = 0; nv < nvar; nv++) {
          double dvp = (s->v[i + 1][nv] - v[nv]) * g->wp[dir][i];
          double dvm = (v[nv] - s->v[i - 1][nv]) * g->wm[dir][i];
          double dv = MINMOD_LIMITER(dvp, dvm);
          s->vp[i][nv] = v[nv] + dv * g->dp[dir][i];
          s->vm[i][nv] = v[nv] - dv * g->dm[dir][i];
        }
        continue;
      }
    }
    double hp = g->php[dir][i], hm = g->phm[dir][i];
    double cm = (hm + 1.0) / (hp - 1.0);
    double cp = (hp + 1.0) / (hm - 1.0);
    for (int nv = 0; nv < nvar; nv++) {
      double dvp = s->vp[i][nv] - v[nv];
      double dvm = s->vm[i][nv] - v[nv];
      if (dvp * dvm >= 0.0) dvp = dvm = 0.0;
      else {
        if (fabs(dvp) >= cm * fabs(dvm)) dvp = -cm * dvm;
        else if (fabs(dvm) >= cp * fabs(dvp)) dvm = -cp * dvp;
      }
      s->vp[i][nv] = v[nv] + dvp;
      s->vm[i][nv] = v[nv] + dvm;
    }
  }
  if (c->flatten_oned) flatten_oned(c, g, s, dir, beg, end);   /* ppm_states.c:229-231 */
}

/* PrimToChar (eigenv.c:575-616) with the left eigenvectors of PrimEigenvectors (eigenv.c:92-200) of zone v */
static void prim_to_char(const gen_cfg *c, int nvar, int dir, const double *v, const double *dv, double *w) {
  int VXn = 1 + dir, VXt = 1 + (dir + 1) % 3, VXb = 1 + (dir + 2) % 3;
  double a2 = A2(c, v), cs = sqrt(a2), rhocs = v[RHO] * cs, rho_cs = v[RHO] / cs;
  if (!c->iso) {
    double L0p = 1.0 / rhocs, L2p = -1.0 / a2;
    w[0] = -1.0 * dv[VXn] + L0p * dv[PRS];
    w[1] = 1.0 * dv[VXn] + L0p * dv[PRS];
    w[2] = dv[RHO] + L2p * dv[PRS];
    w[3] = dv[VXt];
    w[4] = dv[VXb];
  } else {
    double Lr = 1.0 / rho_cs;
    w[0] = Lr * dv[RHO] + -1.0 * dv[VXn];
    w[1] = Lr * dv[RHO] + 1.0 * dv[VXn];
    w[2] = dv[VXt];
    w[3] = dv[VXb];
  }
  for (int nv = NF(c); nv < nvar; nv++) w[nv] = dv[nv];
}

/* States/ppm_states.c:280-590 (CHAR_LIMITING YES, PPM_ORDER 4, PARABOLIC_LIM 1) */
static void states_ppm_char(const gen_cfg *c, const geom_t *g, sweep_t *s, int dir, int beg, int end) {
  int nvar = g->nvar, nf = NF(c);
  int VXn = 1 + dir, VXt = 1 + (dir + 1) % 3, VXb = 1 + (dir + 2) % 3;
  int ntot = g->tot[dir];
  double (*dvF)[NVMAX] = calloc(ntot + 4, sizeof(*dvF));
  double (*vppm4)[NVMAX] = calloc(ntot + 4, sizeof(*vppm4));
  for (int i = beg - 2; i <= end + 1; i++) for (int nv = 0; nv < nvar; nv++) dvF[i][nv] = s->v[i + 1][nv] - s->v[i][nv];
  for (int i = beg - 1; i <= end; i++) {
    const double *wp = g->pwp[dir][i];
    for (int nv = 0; nv < nvar; nv++)
      vppm4[i][nv] = wp[0] * s->v[i - 1][nv] + wp[1] * s->v[i][nv] + wp[2] * s->v[i + 1][nv] + wp[3] * s->v[i + 2][nv];
  }
  for (int i = beg; i <= end; i++) {
    const double *v = s->v[i];
    if (c->flattening) {
      if (s->flag[i] & FLAG_FLAT) {
        for (int nv = 0; nv < nvar; nv++) s->vp[i][nv] = s->vm[i][nv] = v[nv];
        continue;
      } else if (s->flag[i] & FLAG_MINMOD) {
        for (int nv = 0; nv < nvar; nv++) {
          double dp = dvF[i][nv] * g->wp[dir][i];
          double dm = dvF[i - 1][nv] * g->wm[dir][i];
          double dv = MINMOD_LIMITER(dp, dm);
          s->vp[i][nv] = v[nv] + dv * g->dp[dir][i];
          s->vm[i][nv] = v[nv] - dv * g->dm[dir][i];
        }
        continue;
      }
    }
    double dvp[NVMAX], dvm[NVMAX], dwp[NVMAX], dwm[NVMAX], dwp1[NVMAX], dwm1[NVMAX];
    for (int nv = 0; nv < nvar; nv++) { dvp[nv] = vppm4[i][nv] - v[nv]; dvm[nv] = vppm4[i - 1][nv] - v[nv]; }
    prim_to_char(c, nvar, dir, v, dvp, dwp);
    prim_to_char(c, nvar, dir, v, dvm, dwm);
    prim_to_char(c, nvar, dir, v, dvF[i - 1], dwm1);
    prim_to_char(c, nvar, dir, v, dvF[i], dwp1);
    for (int k = 0; k < nvar; k++) {
      dwp[k] = MINMOD_LIMITER(dwp[k], dwp1[k]);
      dwm[k] = MINMOD_LIMITER(dwm[k], -dwm1[k]);
    }
    double hp = g->php[dir][i], hm = g->phm[dir][i];
    double cm = (hm + 1.0) / (hp - 1.0);
    double cp = (hp + 1.0) / (hm - 1.0);
    /* right eigenvectors (eigenv.c:140-178) */
    double a2 = A2(c, v), cs = sqrt(a2), rhocs = v[RHO] * cs, rho_cs = v[RHO] / cs;
    double R[NFLX][NFLX];
    memset(R, 0, sizeof(R));
    R[RHO][0] = 0.5 * rho_cs; R[VXn][0] = -0.5;
    R[RHO][1] = 0.5 * rho_cs; R[VXn][1] = 0.5;
    if (!c->iso) { R[PRS][0] = 0.5 * rhocs; R[PRS][1] = 0.5 * rhocs; R[RHO][2] = 1.0; R[VXt][3] = 1.0; R[VXb][4] = 1.0; }
    else { R[VXt][2] = 1.0; R[VXb][3] = 1.0; }
    for (int nv = 0; nv < nf; nv++) {
      double dp = 0.0, dm = 0.0;
      for (int k = 0; k < nf; k++) { dp += dwp[k] * R[nv][k]; dm += dwm[k] * R[nv][k]; }
      dvp[nv] = dp;
      dvm[nv] = dm;
    }
    for (int nv = nf; nv < nvar; nv++) { dvp[nv] = dwp[nv]; dvm[nv] = dwm[nv]; }
    for (int nv = 0; nv < nvar; nv++) {
      if (dvp[nv] * dvm[nv] >= 0.0) dvp[nv] = dvm[nv] = 0.0;
      else {
        if (fabs(dvp[nv]) >= cm * fabs(dvm[nv])) dvp[nv] = -cm * dvm[nv];
        else if (fabs(dvm[nv]) >= cp * fabs(dvp[nv])) dvm[nv] = -cp * dvp[nv];
      }
      s->vp[i][nv] = v[nv] + dvp[nv];
      s->vm[i][nv] = v[nv] + dvm[nv];
    }
    if (s->vp[i][RHO] < 0.0 || s->vm[i][RHO] < 0.0) {
      dvp[RHO] = 0.5 * (MINMOD_LIMITER(dvF[i][RHO], dvF[i - 1][RHO]));
      dvm[RHO] = -dvp[RHO];
      s->vp[i][RHO] = v[RHO] + dvp[RHO];
      s->vm[i][RHO] = v[RHO] + dvm[RHO];
    }
    if (!c->iso && (s->vp[i][PRS] < 0.0 || s->vm[i][PRS] < 0.0)) {
      dvp[PRS] = 0.5 * (MINMOD_LIMITER(dvF[i][PRS], dvF[i - 1][PRS]));
      dvm[PRS] = -dvp[PRS];
      s->vp[i][PRS] = v[PRS] + dvp[PRS];
      s->vm[i][PRS] = v[PRS] + dvm[PRS];
    }
  }
  free(dvF); free(vppm4);
}

/* HD/hllc.c:28-178, HD/hll.c:30-96, HD/tvdlf.c:38-130, HD/hll_speed.c:76-90, HD/fluxes.c:36-47,
 * adv_flux.c:47-134 (scalars + entropy) */
static void riemann(const gen_cfg *c, const geom_t *g, sweep_t *s, int dir, int beg, int end, double *maxMach) {
  int nvar = g->nvar;
  int VXn = 1 + dir, VXt = 1 + (dir + 1) % 3, VXb = 1 + (dir + 2) % 3;
  for (int i = beg; i <= end; i++) {
    const double *vL = s->vp[i], *vR = s->vm[i + 1];
    double uL[NVMAX], uR[NVMAX], fL[NFLX], fR[NFLX];
    prim_to_cons(c, nvar, vL, uL);
    prim_to_cons(c, nvar, vR, uR);
    const int nf = NF(c);
    double a2L = A2(c, vL);
    double a2R = A2(c, vR);
    fL[RHO] = uL[VXn]; fL[VX1] = uL[VX1] * vL[VXn]; fL[VX2] = uL[VX2] * vL[VXn];
    fL[VX3] = uL[VX3] * vL[VXn];
    fR[RHO] = uR[VXn]; fR[VX1] = uR[VX1] * vR[VXn]; fR[VX2] = uR[VX2] * vR[VXn];
    fR[VX3] = uR[VX3] * vR[VXn];
    double pL, pR;
    if (!c->iso) {
      fL[PRS] = (uL[PRS] + vL[PRS]) * vL[VXn]; fR[PRS] = (uR[PRS] + vR[PRS]) * vR[VXn];
      pL = vL[PRS]; pR = vR[PRS];
    } else { pL = a2L * vL[RHO]; pR = a2R * vR[RHO]; }   /* fluxes.c:47-48 */
    double *flux = s->flux[i];
    if (c->solver == 1) {
      double vRL[NVMAX];
      for (int nv = 0; nv < nvar; nv++) vRL[nv] = 0.5 * (vL[nv] + vR[nv]);
      vRL[VXn] = 0.5 * (fabs(vL[VXn]) + fabs(vR[VXn]));
      double a2 = A2(c, vRL);
      double a = sqrt(a2);
      double cmin = vRL[VXn] - a, cmaxv = vRL[VXn] + a;
      s->cmax[i] = MAXV(fabs(cmaxv), fabs(cmin));
      *maxMach = MAXV(*maxMach, fabs(vRL[VXn]) / sqrt(a2));
      for (int nv = nf; nv--;) flux[nv] = 0.5 * (fL[nv] + fR[nv] - s->cmax[i] * (uR[nv] - uL[nv]));
      s->press[i] = 0.5 * (pL + pR);
    } else if (c->solver == 6) {
      /* HD/ausm.c:20-110 (AUSM+, EOS IDEAL only) */
      const double alpha = 3.0 / 16.0, beta = 0.125, gm = c->gamma;
      double aL = sqrt(gm * vL[PRS] / vL[RHO]);
      double aR = sqrt(gm * vR[PRS] / vR[RHO]);
      double asL2 = vL[VX1] * vL[VX1] + vL[VX2] * vL[VX2] + vL[VX3] * vL[VX3];
      asL2 = aL * aL / (gm - 1.0) + 0.5 * asL2;
      asL2 *= 2.0 * (gm - 1.0) / (gm + 1.0);
      double asR2 = vR[VX1] * vR[VX1] + vR[VX2] * vR[VX2] + vR[VX3] * vR[VX3];
      asR2 = aR * aR / (gm - 1.0) + 0.5 * asR2;
      asR2 *= 2.0 * (gm - 1.0) / (gm + 1.0);
      double asL = sqrt(asL2), asR = sqrt(asR2);
      double atL = asL2 / MAXV(asL, fabs(vL[VXn]));
      double atR = asR2 / MAXV(asR, fabs(vR[VXn]));
      double a = MINV(atL, atR);
      double ML = vL[VXn] / a, MpL, PpL, MR, MmR, PmR;
      if (fabs(ML) >= 1.0) { MpL = 0.5 * (ML + fabs(ML)); PpL = ML > 0.0 ? 1.0 : 0.0; }
      else {
        MpL = 0.25 * (ML + 1.0) * (ML + 1.0) + beta * (ML * ML - 1.0) * (ML * ML - 1.0);
        PpL = 0.25 * (ML + 1.0) * (ML + 1.0) * (2.0 - ML) + alpha * ML * (ML * ML - 1.0) * (ML * ML - 1.0);
      }
      MR = vR[VXn] / a;
      if (fabs(MR) >= 1.0) { MmR = 0.5 * (MR - fabs(MR)); PmR = MR > 0.0 ? 0.0 : 1.0; }
      else {
        MmR = -0.25 * (MR - 1.0) * (MR - 1.0) - beta * (MR * MR - 1.0) * (MR * MR - 1.0);
        PmR = 0.25 * (MR - 1.0) * (MR - 1.0) * (2.0 + MR) - alpha * MR * (MR * MR - 1.0) * (MR * MR - 1.0);
      }
      double m = MpL + MmR;
      double mp = 0.5 * (m + fabs(m));
      double mm = 0.5 * (m - fabs(m));
      s->press[i] = PpL * vL[PRS] + PmR * vR[PRS];
      flux[RHO] = a * (mp * uL[RHO] + mm * uR[RHO]);
      flux[VX1] = a * (mp * uL[VX1] + mm * uR[VX1]);
      flux[VX2] = a * (mp * uL[VX2] + mm * uR[VX2]);
      flux[VX3] = a * (mp * uL[VX3] + mm * uR[VX3]);
      flux[PRS] = a * (mp * (uL[PRS] + vL[PRS]) + mm * (uR[PRS] + vR[PRS]));
      s->cmax[i] = MAXV(fabs(vL[VXn]) + aL, fabs(vR[VXn]) + aR);
      *maxMach = MAXV(fabs(ML), *maxMach);
      *maxMach = MAXV(fabs(MR), *maxMach);
    } else if (c->solver == 5) {
      /* HD/two_shock.c:28-243 (EOS IDEAL only; MAX_ITER 5, small_p = small_rho = 1e-9) */
      const double small_p = 1.e-9, small_rho = 1.e-9;
      double g1_g = 0.5 * (c->gamma + 1.0) / c->gamma;
      if (c->flattening && ((s->flag[i] & FLAG_HLL) || (s->flag[i + 1] & FLAG_HLL))) {   /* two_shock.c:66-88 */
        double aL = sqrt(a2L), aR = sqrt(a2R);          /* HLL_Speed, hll_speed.c:76-90 */
        double cl = MINV(vL[VXn] - aL, vR[VXn] - aR);
        double cr = MAXV(vL[VXn] + aL, vR[VXn] + aR);
        double scrh = fabs(vL[VXn]) + fabs(vR[VXn]);
        scrh /= aL + aR;
        *maxMach = MAXV(scrh, *maxMach);
        double cs = MAXV(fabs(cl), fabs(cr));
        s->cmax[i] = cs;
        cl = MINV(0.0, cl);
        cr = MAXV(0.0, cr);
        double scrh1 = 1.0 / (cr - cl);
        for (int nv = nf; nv--;) {
          flux[nv] = cl * cr * (uR[nv] - uL[nv]) + cr * fL[nv] - cl * fR[nv];
          flux[nv] *= scrh1;
        }
        s->press[i] = (cr * pL - cl * pR) * scrh1;
      } else {
        const double *ql = vL, *qr = vR, *qs;
        double cl = sqrt(c->gamma * ql[PRS] * ql[RHO]);
        double cr = sqrt(c->gamma * qr[PRS] * qr[RHO]);
        double taul = 1.0 / ql[RHO], taur = 1.0 / qr[RHO];
        double vxl = 0.0, vxr = 0.0, scrh1, scrh2, scrh3, scrh4, dp;
        double pstar = qr[PRS] - ql[PRS] - cr * (qr[VXn] - ql[VXn]);
        pstar = ql[PRS] + pstar * cl / (cl + cr);
        pstar = MAXV(small_p, pstar);
        for (int iter = 1; iter <= 5; iter++) {
          vxl = cl * sqrt(1.0 + g1_g * (pstar - ql[PRS]) / ql[PRS]);
          vxr = cr * sqrt(1.0 + g1_g * (pstar - qr[PRS]) / qr[PRS]);
          scrh1 = vxl * vxl;
          scrh1 = 2.0 * scrh1 * vxl / (scrh1 + cl * cl);
          scrh2 = vxr * vxr;
          scrh2 = 2.0 * scrh2 * vxr / (scrh2 + cr * cr);
          scrh3 = ql[VXn] - (pstar - ql[PRS]) / vxl;
          scrh4 = qr[VXn] + (pstar - qr[PRS]) / vxr;
          dp = scrh1 * scrh2 / (scrh1 + scrh2) * (scrh4 - scrh3);
          pstar -= dp;
          pstar = MAXV(small_p, pstar);
          if (fabs(dp / pstar) < 1.e-6) break;
        }
        scrh3 = ql[VXn] - (pstar - ql[PRS]) / vxl;
        scrh4 = qr[VXn] + (pstar - qr[PRS]) / vxr;
        double ustar = 0.5 * (scrh3 + scrh4);
        double sigma, taus, cs, zs;
        if (ustar > 0.0) { sigma = 1.0; taus = taul; cs = cl * taul; zs = vxl; qs = ql; }
        else { sigma = -1.0; taus = taur; cs = cr * taur; zs = vxr; qs = qr; }
        double rho_star = taus - (pstar - qs[PRS]) / (zs * zs);
        rho_star = MAXV(small_rho, 1.0 / rho_star);
        double cstar = sqrt(c->gamma * pstar / rho_star);
        double lambda_s, lambda_star;
        if (pstar < qs[PRS]) {
          lambda_s = cs - sigma * qs[VXn];
          lambda_star = cstar - sigma * ustar;
        } else {
          lambda_s = lambda_star = zs * taus - sigma * qs[VXn];
        }
        double vS[NVMAX], uS[NVMAX];
        for (int nv = 0; nv < nvar; nv++) vS[nv] = 0.0;
        if (lambda_star > 0.0) { vS[RHO] = rho_star; vS[VXn] = ustar; vS[PRS] = pstar; }
        else if (lambda_s < 0.0) { vS[RHO] = qs[RHO]; vS[VXn] = qs[VXn]; vS[PRS] = qs[PRS]; }
        else {
          scrh1 = MAXV(lambda_s - lambda_star, lambda_s + lambda_star);
          scrh1 = MAXV(1.e-12, scrh1);
          double zeta = 0.5 * (1.0 + (lambda_s + lambda_star) / scrh1);
          vS[RHO] = zeta * rho_star + (1.0 - zeta) * qs[RHO];
          vS[VXn] = zeta * ustar + (1.0 - zeta) * qs[VXn];
          vS[PRS] = zeta * pstar + (1.0 - zeta) * qs[PRS];
        }
        vS[VXt] = qs[VXt];
        vS[VXb] = qs[VXb];
        prim_to_cons(c, NFLX, vS, uS);
        double a2S = c->gamma * vS[PRS] / vS[RHO];
        flux[RHO] = uS[VXn]; flux[VX1] = uS[VX1] * vS[VXn]; flux[VX2] = uS[VX2] * vS[VXn];
        flux[VX3] = uS[VX3] * vS[VXn]; flux[PRS] = (uS[PRS] + vS[PRS]) * vS[VXn];
        s->press[i] = vS[PRS];
        cstar = sqrt(a2S);
        scrh1 = fabs(vS[VXn]) / cstar;
        *maxMach = MAXV(scrh1, *maxMach);
        s->cmax[i] = fabs(vS[VXn]) + cstar;
      }
    } else if (c->solver == 4) {
      /* HD/roe.c:48-346 (ROE_AVERAGE YES, the file's default) */
      const double delta = 1.e-7;
      double gmm1 = c->gamma - 1.0, gmm1_inv = 1.0 / gmm1;
      double Rc[NFLX][NFLX], lambda[NFLX], alambda[NFLX], eta[NFLX], dv[NFLX], um[NFLX];
      memset(Rc, 0, sizeof(Rc));
      int done = 0;
      if (c->flattening && ((s->flag[i] & FLAG_HLL) || (s->flag[i + 1] & FLAG_HLL))) {   /* roe.c:101-116 */
        double aL = sqrt(a2L), aR = sqrt(a2R);          /* HLL_Speed, hll_speed.c:76-90 */
        double bmin = MINV(vL[VXn] - aL, vR[VXn] - aR);
        double bmax = MAXV(vL[VXn] + aL, vR[VXn] + aR);
        double scrh = fabs(vL[VXn]) + fabs(vR[VXn]);
        scrh /= aL + aR;
        *maxMach = MAXV(scrh, *maxMach);
        double a = MAXV(fabs(bmin), fabs(bmax));
        s->cmax[i] = a;
        bmin = MINV(0.0, bmin);
        bmax = MAXV(0.0, bmax);
        scrh = 1.0 / (bmax - bmin);
        for (int nv = nf; nv--;) {
          flux[nv] = bmin * bmax * (uR[nv] - uL[nv]) + bmax * fL[nv] - bmin * fR[nv];
          flux[nv] *= scrh;
        }
        s->press[i] = (bmax * pL - bmin * pR) * scrh;
        done = 1;
      }
      if (!done) {
        const double *ql = vL, *qr = vR;
        double a2, a, h = 0.0, vel2 = 0.0;
        for (int nv = nf; nv--;) dv[nv] = qr[nv] - ql[nv];
        double sq = sqrt(qr[RHO] / ql[RHO]);
        um[RHO] = ql[RHO] * sq;
        sq = 1.0 / (1.0 + sq);
        double cq = 1.0 - sq;
        um[VX1] = sq * ql[VX1] + cq * qr[VX1];
        um[VX2] = sq * ql[VX2] + cq * qr[VX2];
        um[VX3] = sq * ql[VX3] + cq * qr[VX3];
        if (!c->iso) {
          vel2 = um[VX1] * um[VX1] + um[VX2] * um[VX2] + um[VX3] * um[VX3];
          double hl = 0.5 * (ql[VX1] * ql[VX1] + ql[VX2] * ql[VX2] + ql[VX3] * ql[VX3]);
          hl += a2L * gmm1_inv;
          double hr = 0.5 * (qr[VX1] * qr[VX1] + qr[VX2] * qr[VX2] + qr[VX3] * qr[VX3]);
          hr += a2R * gmm1_inv;
          h = sq * hl + cq * hr;
          a2 = gmm1 * (h - 0.5 * vel2);
          a = sqrt(a2);
        } else {
          a2 = 0.5 * (a2L + a2R);
          a = sqrt(a2);
        }
        int nn = 0;                         /* u - c_s */
        lambda[nn] = um[VXn] - a;
        if (!c->iso) eta[nn] = 0.5 / a2 * (dv[PRS] - dv[VXn] * um[RHO] * a);
        else eta[nn] = 0.5 * (dv[RHO] - um[RHO] * dv[VXn] / a);
        Rc[RHO][nn] = 1.0; Rc[VXn][nn] = um[VXn] - a; Rc[VXt][nn] = um[VXt]; Rc[VXb][nn] = um[VXb];
        if (!c->iso) Rc[PRS][nn] = h - um[VXn] * a;
        nn = 1;                             /* u + c_s */
        lambda[nn] = um[VXn] + a;
        if (!c->iso) eta[nn] = 0.5 / a2 * (dv[PRS] + dv[VXn] * um[RHO] * a);
        else eta[nn] = 0.5 * (dv[RHO] + um[RHO] * dv[VXn] / a);
        Rc[RHO][nn] = 1.0; Rc[VXn][nn] = um[VXn] + a; Rc[VXt][nn] = um[VXt]; Rc[VXb][nn] = um[VXb];
        if (!c->iso) Rc[PRS][nn] = h + um[VXn] * a;
        if (!c->iso) {                      /* u (entropy wave) */
          nn = 2;
          lambda[nn] = um[VXn];
          eta[nn] = dv[RHO] - dv[PRS] / a2;
          Rc[RHO][nn] = 1.0; Rc[VX1][nn] = um[VX1]; Rc[VX2][nn] = um[VX2]; Rc[VX3][nn] = um[VX3];
          Rc[PRS][nn] = 0.5 * vel2;
        }
        nn++;                               /* u (shear waves) */
        lambda[nn] = um[VXn];
        eta[nn] = um[RHO] * dv[VXt];
        Rc[VXt][nn] = 1.0;
        if (!c->iso) Rc[PRS][nn] = um[VXt];
        nn++;
        lambda[nn] = um[VXn];
        eta[nn] = um[RHO] * dv[VXb];
        Rc[VXb][nn] = 1.0;
        if (!c->iso) Rc[PRS][nn] = um[VXb];
        s->cmax[i] = fabs(um[VXn]) + a;
        *maxMach = MAXV(fabs(um[VXn] / a), *maxMach);
        if (c->ndim > 1) {                  /* roe.c:262-287: HLL inside strong shocks */
          double scrh;
          if (!c->iso) { scrh = fabs(ql[PRS] - qr[PRS]); scrh /= MINV(ql[PRS], qr[PRS]); }
          else { scrh = fabs(ql[RHO] - qr[RHO]); scrh /= MINV(ql[RHO], qr[RHO]); scrh *= a * a; }
          if (scrh > 0.5 && (qr[VXn] < ql[VXn])) {
            double bmin = MINV(0.0, lambda[0]);
            double bmax = MAXV(0.0, lambda[1]);
            double scrh1 = 1.0 / (bmax - bmin);
            for (int nv = nf; nv--;) {
              flux[nv] = bmin * bmax * (uR[nv] - uL[nv]) + bmax * fL[nv] - bmin * fR[nv];
              flux[nv] *= scrh1;
            }
            s->press[i] = (bmax * pL - bmin * pR) * scrh1;
            done = 1;
          }
        }
        if (!done) {
          for (int nv = nf; nv--;) alambda[nv] = fabs(lambda[nv]);
          if (alambda[0] <= delta) alambda[0] = 0.5 * lambda[0] * lambda[0] / delta + 0.5 * delta;   /* entropy fix */
          if (alambda[1] <= delta) alambda[1] = 0.5 * lambda[1] * lambda[1] / delta + 0.5 * delta;
          for (int nv = nf; nv--;) {
            flux[nv] = fL[nv] + fR[nv];
            for (int k = nf; k--;) flux[nv] -= alambda[k] * eta[k] * Rc[nv][k];
            flux[nv] *= 0.5;
          }
          s->press[i] = 0.5 * (pL + pR);
        }
      }
    } else {
      double aL = sqrt(a2L), aR = sqrt(a2R);
      double SL = MINV(vL[VXn] - aL, vR[VXn] - aR);
      double SR = MAXV(vL[VXn] + aL, vR[VXn] + aR);
      double scrh = fabs(vL[VXn]) + fabs(vR[VXn]);
      scrh /= aL + aR;
      *maxMach = MAXV(scrh, *maxMach);
      s->cmax[i] = MAXV(fabs(SL), fabs(SR));
      if (SL > 0.0) {
        for (int nv = nf; nv--;) flux[nv] = fL[nv];
        s->press[i] = pL;
      } else if (SR < 0.0) {
        for (int nv = nf; nv--;) flux[nv] = fR[nv];
        s->press[i] = pR;
      } else if (c->solver == 2 ||
                 (c->flattening && ((s->flag[i] & FLAG_HLL) || (s->flag[i + 1] & FLAG_HLL)))) {
        scrh = 1.0 / (SR - SL);
        for (int nv = nf; nv--;) {
          flux[nv] = SL * SR * (uR[nv] - uL[nv]) + SR * fL[nv] - SL * fR[nv];
          flux[nv] *= scrh;
        }
        s->press[i] = (SR * pL - SL * pR) * scrh;
      } else {
        double usL[NFLX], usR[NFLX];
        double vxr = vR[VXn], vxl = vL[VXn];
        double vs;
        if (c->iso) {   /* hllc.c:137-150 */
          scrh = 1.0 / (SR - SL);
          double rho = (SR * uR[RHO] - SL * uL[RHO] - fR[RHO] + fL[RHO]) * scrh;
          double mx = (SR * uR[VXn] - SL * uL[VXn] - fR[VXn] + fL[VXn]) * scrh;
          usL[RHO] = usR[RHO] = rho;
          usL[VXn] = usR[VXn] = mx;
          vs = (SR * fL[RHO] - SL * fR[RHO] + SR * SL * (uR[RHO] - uL[RHO]));
          vs *= scrh;
          vs /= rho;
          usL[VXt] = rho * vL[VXt]; usR[VXt] = rho * vR[VXt];
          usL[VXb] = rho * vL[VXb]; usR[VXb] = rho * vR[VXb];
        } else {
        double qL = vL[PRS] + uL[VXn] * (vL[VXn] - SL);     /* hllc.c:112-116 */
        double qR = vR[PRS] + uR[VXn] * (vR[VXn] - SR);
        double wL = vL[RHO] * (vL[VXn] - SL);
        double wR = vR[RHO] * (vR[VXn] - SR);
        vs = (qR - qL) / (wR - wL);
        usL[RHO] = uL[RHO] * (SL - vxl) / (SL - vs);
        usR[RHO] = uR[RHO] * (SR - vxr) / (SR - vs);
        usL[VXn] = usL[RHO] * vs;       usR[VXn] = usR[RHO] * vs;
        usL[VXt] = usL[RHO] * vL[VXt];  usR[VXt] = usR[RHO] * vR[VXt];
        usL[VXb] = usL[RHO] * vL[VXb];  usR[VXb] = usR[RHO] * vR[VXb];
        usL[PRS] = uL[PRS] / vL[RHO] + (vs - vxl) * (vs + vL[PRS] / (vL[RHO] * (SL - vxl)));
        usR[PRS] = uR[PRS] / vR[RHO] + (vs - vxr) * (vs + vR[PRS] / (vR[RHO] * (SR - vxr)));
        usL[PRS] *= usL[RHO];
        usR[PRS] *= usR[RHO];
        }
        if (vs >= 0.0) {
          for (int nv = nf; nv--;) flux[nv] = fL[nv] + SL * (usL[nv] - uL[nv]);
          s->press[i] = pL;
        } else {
          for (int nv = nf; nv--;) flux[nv] = fR[nv] + SR * (usR[nv] - uR[nv]);
          s->press[i] = pR;
        }
      }
    }
    const double *ts = flux[RHO] > 0.0 ? vL : vR;
    for (int nv = nf; nv < nvar; nv++) flux[nv] = flux[RHO] * ts[nv];
    if (c->entropy) {
      int ENTR = nvar - 1;
      if (flux[RHO] >= 0.0) flux[ENTR] = vL[ENTR] * flux[RHO];
      else flux[ENTR] = vR[ENTR] * flux[RHO];
    }
  }
}

/* ---------------------------------------------------------------------------------------
 *  Boundaries: boundary.c:228-459 (side order, full transverse range), fills :617-767,
 *  FlipSign :503-610 (reflective: vn; axisymmetric: vn and vphi; eqtsymmetric: vn)
 * --------------------------------------------------------------------------------------- */
static double *g_Uc_for_floor = NULL;   /* Uc, while Boundary() runs inside stages >= 2 */
static void ldw_userdef_side(const gen_cfg *c, const geom_t *g, double *Vc, int side);
static void ldw_internal_floor(const gen_cfg *c, const geom_t *g, double *Vc);

/* PolarAxisBoundary(), boundary.c:770-840: ghost zones across the axis take the zone half a turn away, the two
 * components that change sign through the axis flipped */
static void polar_axis_boundary(const gen_cfg *c, const geom_t *g, double *Vc, int side) {
  int nvar = g->nvar, dir = side / 2, hi = side & 1;
  int pdir = c->geometry == POLAR ? 1 : 2;                 /* phi: x2 (POLAR), x3 (SPHERICAL) */
  int nphi = g->end[pdir] - g->beg[pdir] + 1;
  int lo[3] = {0, 0, 0}, up[3] = {g->tot[0] - 1, g->tot[1] - 1, g->tot[2] - 1};
  if (hi) { lo[dir] = g->end[dir] + 1; up[dir] = g->tot[dir] - 1; }
  else { lo[dir] = 0; up[dir] = g->beg[dir] - 1; }
  for (int k = lo[2]; k <= up[2]; k++)
    for (int j = lo[1]; j <= up[1]; j++)
      for (int i = lo[0]; i <= up[0]; i++) {
        int src[3] = {i, j, k};
        src[pdir] += nphi / 2;
        if (src[pdir] > g->end[pdir]) src[pdir] -= nphi;
        src[dir] = hi ? 2 * g->end[dir] - src[dir] + 1 : 2 * g->beg[dir] - src[dir] - 1;
        long o = k * g->sk + j * g->sj + i, os = src[2] * g->sk + src[1] * g->sj + src[0];
        for (int nv = 0; nv < nvar; nv++) Vc[nv * g->sv + o] = Vc[nv * g->sv + os];
        Vc[(1 + dir) * g->sv + o] *= -1.0;                 /* POLAR: VX1, VX2; SPHERICAL: VX2, VX3 */
        Vc[(1 + pdir) * g->sv + o] *= -1.0;
      }
}

static void boundary(const gen_cfg *c, const geom_t *g, double *Vc) {
  int nvar = g->nvar;
  if (c->ldw_bc) ldw_internal_floor(c, g, Vc);        /* boundary.c:126-128, side == 0 */
  for (int side = 0; side < 2 * c->ndim; side++) {
    int type = c->bc[side], dir = side / 2, hi = side & 1;
    if (type == 8) { if (c->ldw_bc) ldw_userdef_side(c, g, Vc, side); continue; }
    if (type == 9) { polar_axis_boundary(c, g, Vc, side); continue; }
    if (type != 1 && type != 2 && type != 3 && type != 4 && type != 5) continue;
    int nb = g->beg[dir], ne = g->end[dir], nxd = ne - nb + 1;
    int lo[3] = {0, 0, 0}, up[3] = {g->tot[0] - 1, g->tot[1] - 1, g->tot[2] - 1};
    if (hi) { lo[dir] = ne + 1; up[dir] = g->tot[dir] - 1; }
    else { lo[dir] = 0; up[dir] = nb - 1; }
    for (int nv = 0; nv < nvar; nv++) {
      double sgn = 1.0;
      if ((type == 2 || type == 3 || type == 4) && nv == 1 + dir) sgn = -1.0;
      if (type == 3 && c->geometry != CARTESIAN && nv == (c->geometry == POLAR ? VX2 : VX3)) sgn = -1.0;   /* boundary.c:560-575: iVPHI */
      int mirror = (type == 2 || type == 3 || type == 4);
      double *q = Vc + nv * g->sv;
      for (int k = lo[2]; k <= up[2]; k++)
        for (int j = lo[1]; j <= up[1]; j++)
          for (int i = lo[0]; i <= up[0]; i++) {
            int idx[3] = {i, j, k};
            int n = idx[dir], src;
            if (type == 1) src = hi ? ne : nb;
            else if (type == 5) src = hi ? n - nxd : n + nxd;
            else src = hi ? 2 * ne + 1 - n : 2 * nb - 1 - n;
            idx[dir] = src;
            double val = q[idx[2] * g->sk + idx[1] * g->sj + idx[0]];
            q[k * g->sk + j * g->sj + i] = mirror ? sgn * val : val;
          }
    }
  }
  if (c->entropy) {      /* ComputeEntropy, entropy_switch.c:14-39, Entropy eos.c:77-103 */
    int ENTR = nvar - 1;
    for (long o = 0; o < g->sv; o++) Vc[ENTR * g->sv + o] = Vc[PRS * g->sv + o] / pow(Vc[RHO * g->sv + o], c->gamma);
  }
}

/* flag_shock.c:81-260 */
static void flag_shock(const gen_cfg *c, const geom_t *g, const double *Vc, uint16_t *flag) {
  const double *pt = Vc + PRS * g->sv;
  double *pt_iso = NULL;
  if (c->iso) {   /* flag_shock.c:138-139 */
    pt_iso = malloc(g->sv * 8);
    for (long o = 0; o < g->sv; o++) pt_iso[o] = Vc[RHO * g->sv + o] * c->iso_cs * c->iso_cs;
    pt = pt_iso;
  }
  const double *vx[3] = {Vc + VX1 * g->sv, Vc + VX2 * g->sv, Vc + VX3 * g->sv};
  long st[3] = {1, g->sj, g->sk};
  if (c->entropy) for (long o = 0; o < g->sv; o++) flag[o] |= FLAG_ENTROPY;
  int inc[3] = {c->ndim > 0, c->ndim > 1, c->ndim > 2};
  for (int k = inc[2]; k < g->tot[2] - inc[2]; k++)
    for (int j = inc[1]; j < g->tot[1] - inc[1]; j++)
      for (int i = inc[0]; i < g->tot[0] - inc[0]; i++) {
        long o = k * g->sk + j * g->sj + i;
        int idx[3] = {i, j, k};
        double dvx[3] = {0, 0, 0}, divv;
        for (int d = 0; d < c->ndim; d++) {
          if (c->geometry == CARTESIAN) dvx[d] = (vx[d][o + st[d]] - vx[d][o - st[d]]) / g->dx[d][idx[d]];
          else {
            int im[3] = {i, j, k};
            im[d] -= 1;
            dvx[d] = A_at(g, d, k, j, i) * (vx[d][o + st[d]] + vx[d][o]) -
                     A_at(g, d, im[2], im[1], im[0]) * (vx[d][o - st[d]] + vx[d][o]);
          }
        }
        divv = dvx[0];
        if (c->ndim > 1) divv = divv + dvx[1];
        if (c->ndim > 2) divv = divv + dvx[2];
        if (c->geometry != CARTESIAN) divv = divv / g->dV[o];
        if (divv < 0.0) {
          double pt_min = pt[o], gradp = 0.0;
          for (int d = 0; d < c->ndim; d++) {
            double m = MINV(pt[o + st[d]], pt[o - st[d]]);
            pt_min = MINV(pt_min, m);
          }
          for (int d = 0; d < c->ndim; d++) {
            double dpx = fabs(pt[o + st[d]] - pt[o - st[d]]);
            gradp = (d == 0) ? dpx : gradp + dpx;
          }
          if (c->flattening && gradp > 5.0 * pt_min) {
            flag[o] |= FLAG_HLL | FLAG_MINMOD;
            for (int d = 0; d < c->ndim; d++) { flag[o + st[d]] |= FLAG_MINMOD; flag[o - st[d]] |= FLAG_MINMOD; }
          }
          if (c->entropy == 1 && gradp > 0.05 * pt_min) {   /* SELECTIVE: flag_shock.c:256-267, EPS_PSHOCK_ENTROPY */
            flag[o] &= ~FLAG_ENTROPY;
            for (int d = 0; d < c->ndim; d++) { flag[o + st[d]] &= ~FLAG_ENTROPY; flag[o - st[d]] &= ~FLAG_ENTROPY; }
          }
        }
      }
  free(pt_iso);
}

/* ---------------------------------------------------------------------------------------
 *  Line-driven wind and BLONDIN cooling hooks (filled in below)
 * --------------------------------------------------------------------------------------- */
static void ldw_vgrad_calc(const gen_cfg *c, const geom_t *g, const double *Vc, double *dvds);
static void ldw_line_force(const gen_cfg *c, const geom_t *g, const double *v, const double *dvds, long o, double *grad);

/* ---------------------------------------------------------------------------------------
 *  RING_AVERAGE: Src/ring_average.c (RingAverageSize :620-735, RingAverageCons :27-131,
 *  RingAverageReconstruct :172-400), MP5_Reconstruct / Median of Src/reconstruct.c:38-93,296-304
 * --------------------------------------------------------------------------------------- */
static void ring_size(const gen_cfg *c, geom_t *g) {
  int rd = c->geometry == POLAR ? 0 : 1;                   /* rings are counted along r (POLAR) or theta (SPHERICAL) */
  g->csize = calloc(g->tot[rd], sizeof(int));
  if (c->ring_average <= 1) return;
  int cs;
  if (c->bc[2 * rd] == 9) {                                /* axis at the lower boundary */
    cs = c->ring_average;
    for (int n = g->beg[rd]; n <= g->end[rd]; n++) { g->csize[n] = cs; if (cs > 1) cs >>= 1; }
  } else for (int n = g->beg[rd]; n <= g->end[rd]; n++) g->csize[n] = 1;
  if (c->geometry == SPHERICAL) {                          /* south pole */
    if (c->bc[3] == 9) {
      cs = c->ring_average;
      for (int n = g->end[1]; n >= g->beg[1]; n--) { g->csize[n] = MAXV(g->csize[n], cs); if (cs > 1) cs >>= 1; }
    } else for (int n = g->end[1]; n >= g->beg[1]; n--) g->csize[n] = MAXV(g->csize[n], 1);
  }
}

static void ring_average_cons(const gen_cfg *c, const geom_t *g, double *Uc) {
  int nvar = g->nvar;
  double uav[NVMAX], dVav;
  if (c->geometry == POLAR) {
    if (g->csize[g->beg[0]] == 1) return;
    int i = g->beg[0];
    while (i <= g->end[0] && g->csize[i] > 1) {
      for (int k = g->beg[2]; k <= g->end[2]; k++)
        for (int j = g->beg[1]; j <= g->end[1]; j += g->csize[i]) {
          dVav = 0.0;
          for (int nv = 0; nv < nvar; nv++) uav[nv] = 0.0;
          for (int j1 = j; j1 < j + g->csize[i]; j1++) {
            long o = k * g->sk + j1 * g->sj + i;
            dVav += g->dV[o];
            for (int nv = 0; nv < nvar; nv++) uav[nv] += Uc[o * nvar + nv] * g->dV[o];
          }
          for (int j1 = j; j1 < j + g->csize[i]; j1++) {
            long o = k * g->sk + j1 * g->sj + i;
            for (int nv = 0; nv < nvar; nv++) Uc[o * nvar + nv] = uav[nv] / dVav;
          }
        }
      i++;
    }
  } else {
    if (g->csize[g->beg[1]] == 1 && g->csize[g->end[1]] == 1) return;
    for (int pole = 0; pole < 2; pole++) {
      int j = pole ? g->end[1] : g->beg[1];
      while (j >= g->beg[1] && j <= g->end[1] && g->csize[j] > 1) {
        for (int i = g->beg[0]; i <= g->end[0]; i++)
          for (int k = g->beg[2]; k <= g->end[2]; k += g->csize[j]) {
            dVav = 0.0;
            for (int nv = 0; nv < nvar; nv++) uav[nv] = 0.0;
            for (int k1 = k; k1 < k + g->csize[j]; k1++) {
              long o = k1 * g->sk + j * g->sj + i;
              dVav += g->dV[o];
              for (int nv = 0; nv < nvar; nv++) uav[nv] += Uc[o * nvar + nv] * g->dV[o];
            }
            for (int k1 = k; k1 < k + g->csize[j]; k1++) {
              long o = k1 * g->sk + j * g->sj + i;
              for (int nv = 0; nv < nvar; nv++) Uc[o * nvar + nv] = uav[nv] / dVav;
            }
          }
        j += pole ? -1 : 1;
      }
    }
  }
}

static double mp5_reconstruct(const double *F, int j) {
  const double alpha = 4.0, epsm = 1.e-12;
  double f = 2.0 * F[j - 2] - 13.0 * F[j - 1] + 47.0 * F[j] + 27.0 * F[j + 1] - 3.0 * F[j + 2];
  f /= 60.0;
  double fMP = F[j] + MINMOD_LIMITER(F[j + 1] - F[j], alpha * (F[j] - F[j - 1]));
  if ((f - F[j]) * (f - fMP) <= epsm) return f;
  double d2m = F[j - 2] + F[j] - 2.0 * F[j - 1];
  double d2 = F[j - 1] + F[j + 1] - 2.0 * F[j];
  double d2p = F[j] + F[j + 2] - 2.0 * F[j + 1];
  double scrh1 = MINMOD_LIMITER(4.0 * d2 - d2p, 4.0 * d2p - d2);
  double scrh2 = MINMOD_LIMITER(d2, d2p);
  double dMMp = MINMOD_LIMITER(scrh1, scrh2);
  scrh1 = MINMOD_LIMITER(4.0 * d2m - d2, 4.0 * d2 - d2m);
  scrh2 = MINMOD_LIMITER(d2, d2m);
  double dMMm = MINMOD_LIMITER(scrh1, scrh2);
  double fUL = F[j] + alpha * (F[j] - F[j - 1]);
  double fAV = 0.5 * (F[j] + F[j + 1]);
  double fMD = fAV - 0.5 * dMMp;
  double fLC = 0.5 * (3.0 * F[j] - F[j - 1]) + 4.0 / 3.0 * dMMm;
  scrh1 = MINV(F[j], F[j + 1]); scrh1 = MINV(scrh1, fMD);
  scrh2 = MINV(F[j], fUL); scrh2 = MINV(scrh2, fLC);
  double fmin = MAXV(scrh1, scrh2);
  scrh1 = MAXV(F[j], F[j + 1]); scrh1 = MAXV(scrh1, fMD);
  scrh2 = MAXV(F[j], fUL); scrh2 = MAXV(scrh2, fLC);
  double fmax = MINV(scrh1, scrh2);
  return f + MINMOD_LIMITER(fmin - f, fmax - f);             /* Median(f, fmin, fmax) */
}

/* chunk averages -> states on the reduced grid -> parabola through (vam, va, vap) evaluated at the sub-zone edges */
static void ring_reconstruct(const gen_cfg *c, const geom_t *g, sweep_t *s, int dir, int chunk_size) {
  if (c->ring_rec == 1 || chunk_size == 1) return;
  int nvar = g->nvar, ngh = c->ng;
  int dbeg = g->beg[dir], dend = g->end[dir], nphi = dend - dbeg + 1;
  int nchunks = nphi / chunk_size, cbeg = dbeg, cend = dbeg + nchunks - 1;
  int nmax = g->tot[dir] + 8;
  double (*va)[NVMAX] = calloc(nmax, sizeof(*va)), (*vap)[NVMAX] = calloc(nmax, sizeof(*vap)), (*vam)[NVMAX] = calloc(nmax, sizeof(*vam));
  for (int j = dbeg; j <= dend; j += chunk_size) {
    int ja = (j - dbeg) / chunk_size + dbeg;
    for (int nv = 0; nv < nvar; nv++) va[ja][nv] = s->v[j][nv];
  }
  for (int j = 0; j < cbeg; j++) for (int nv = 0; nv < nvar; nv++) va[j][nv] = va[j + nchunks][nv];
  for (int j = cend + 1; j <= cend + ngh; j++) for (int nv = 0; nv < nvar; nv++) va[j][nv] = va[j - nchunks][nv];
  if (c->ring_rec == 2) {
    for (int j = cbeg - 1; j <= cend + 1; j++)
      for (int nv = 0; nv < nvar; nv++) {
        double dvap = va[j + 1][nv] - va[j][nv], dvam = va[j][nv] - va[j - 1][nv];
        double dva = (dvap * dvam > 0.0 ? 2.0 * dvap * dvam / (dvap + dvam) : 0.0);     /* VANLEER_LIMITER */
        vap[j][nv] = va[j][nv] + 0.5 * dva;
        vam[j][nv] = va[j][nv] - 0.5 * dva;
      }
  } else {      /* RING_AVERAGE_REC 5 */
    int nphi_tot = nchunks + 2 * ngh;
    double *qfwd = calloc(nmax, 8), *qbck = calloc(nmax, 8);
    for (int nv = 0; nv < nvar; nv++) {
      for (int j = 0; j < nphi_tot; j++) { qfwd[j] = va[j][nv]; qbck[j] = va[nphi_tot - j - 1][nv]; }
      for (int j = cbeg - 1; j <= cend; j++) {
        vap[j][nv] = mp5_reconstruct(qfwd, j);
        vam[j][nv] = mp5_reconstruct(qbck, cend - (j - cbeg));
      }
    }
    free(qfwd); free(qbck);
  }
  for (int j = dbeg; j <= dend; j++) {
    int ja = (j - dbeg) / chunk_size + dbeg;
    int k = (j - dbeg) % chunk_size + 1;
    double xp = k / (double)chunk_size, xm = (k - 1.0) / (double)chunk_size;
    for (int nv = 0; nv < nvar; nv++) {
      double A = 3.0 * ((vap[ja][nv] + vam[ja][nv]) - 2.0 * va[ja][nv]);
      double B = -4.0 * vam[ja][nv] - 2.0 * vap[ja][nv] + 6.0 * va[ja][nv];
      double Cc = vam[ja][nv];
      s->vp[j][nv] = A * xp * xp + B * xp + Cc;
      s->vm[j][nv] = A * xm * xm + B * xm + Cc;
    }
  }
  for (int j = 0; j < dbeg; j++) for (int nv = 0; nv < nvar; nv++) { s->vp[j][nv] = s->vp[j + nphi][nv]; s->vm[j][nv] = s->vm[j + nphi][nv]; }
  for (int j = dend + 1; j <= dend + ngh; j++) for (int nv = 0; nv < nvar; nv++) { s->vp[j][nv] = s->vp[j - nphi][nv]; s->vm[j][nv] = s->vm[j - nphi][nv]; }
  free(va); free(vap); free(vam);
}

/* Time_Stepping/update_stage.c:35-394 + MHD/rhs.c:84-420 + MHD/rhs_source.c:101-470 */
static void update_stage(const gen_cfg *c, const geom_t *g, const double *Vc, double *Uc, const uint16_t *flag,
                         double *C_dt, double *dvds, double dt, int stage, double *invDt_hyp, double *maxMach) {
  int nvar = g->nvar;
  int nmax = MAXV(g->tot[0], MAXV(g->tot[1], g->tot[2]));
  sweep_t s;
  sweep_alloc(&s, nmax);
  if (c->ndim > 1 && stage == 1) memset(C_dt, 0, sizeof(double) * g->sv);
  if (c->ldw) ldw_vgrad_calc(c, g, Vc, dvds);         /* update_stage.c:116-118 */
  double *inv_dl = calloc(nmax, 8);
  double *phi_p = calloc(nmax, 8);
  for (int dir = 0; dir < c->ndim; dir++) {
    int VXn = 1 + dir;
    const int iMPHI = (c->geometry == POLAR) ? VX2 : VX3;   /* pluto.h: iVPHI per geometry */
    int ntot = g->tot[dir], nbeg = g->beg[dir], nend = g->end[dir];
    long st = dir == 0 ? 1 : (dir == 1 ? g->sj : g->sk);
    int t1 = (dir + 1) % 3, t2 = (dir + 2) % 3;
    int idx[3];
    /* BOX_TRANSVERSE_LOOP: the direction after dir varies fastest (macros.h / rbox.c) */
    for (idx[t2] = g->beg[t2]; idx[t2] <= g->end[t2]; idx[t2]++)
      for (idx[t1] = g->beg[t1]; idx[t1] <= g->end[t1]; idx[t1]++) {
        idx[dir] = 0;
        long base = idx[2] * g->sk + idx[1] * g->sj + idx[0];
        for (int n = 0; n < ntot; n++) {
          for (int nv = 0; nv < nvar; nv++) s.v[n][nv] = Vc[nv * g->sv + base + n * st];
          s.flag[n] = flag[base + n * st];
        }
        if (c->ppm && c->char_limiting) states_ppm_char(c, g, &s, dir, nbeg - 1, nend + 1);
        else if (c->ppm) states_ppm(c, g, &s, dir, nbeg - 1, nend + 1);
        else states(c, g, &s, dir, nbeg - 1, nend + 1);
        /* RingAverageReconstruct after States() in the phi sweep, update_stage.c:230-234 */
        int ring_line = 0;
        if (c->ring_average > 1 && c->geometry == POLAR && dir == 1) ring_line = g->csize[idx[0]];
        if (c->ring_average > 1 && c->geometry == SPHERICAL && dir == 2) ring_line = g->csize[idx[1]];
        if (ring_line > 1) ring_reconstruct(c, g, &s, dir, ring_line);
        riemann(c, g, &s, dir, nbeg - 1, nend, maxMach);
        if (c->body_force & 2) {   /* TotalFlux(): flux[ENG] += flux[RHO] phi_p at the faces (rhs.c:171-179,525,581,621) */
          for (int n = nbeg - 1; n <= nend; n++) {
            phi_p[n] = c->bf_phi[1 + dir][base + n * st];
            if (!c->iso) s.flux[n][PRS] += s.flux[n][RHO] * phi_p[n];
          }
        }
        /* ---- RightHandSide ---- */
        int i = idx[0], j = idx[1], k = idx[2];
        if (c->geometry == CARTESIAN) {
          for (int n = nbeg; n <= nend; n++) {
            double scrh = dt / g->dx[dir][n];
            for (int nv = 0; nv < nvar; nv++) s.rhs[n][nv] = -scrh * (s.flux[n][nv] - s.flux[n - 1][nv]);
            s.rhs[n][VXn] -= scrh * (s.press[n] - s.press[n - 1]);
          }
        } else {
          /* TotalFlux rhs.c:530-600: fA = F A, fA[iMPHI] *= |x1p| (r) or |sp| (theta) */
          for (int n = nbeg - 1; n <= nend; n++) {
            int q[3] = {i, j, k};
            q[dir] = n;
            double A = A_at(g, dir, q[2], q[1], q[0]);
            for (int nv = 0; nv < nvar; nv++) s.fA[n][nv] = s.flux[n][nv] * A;
            if (dir == 0) s.fA[n][iMPHI] *= fabs(g->xr[0][n]);
            else if (dir == 1 && c->geometry == SPHERICAL) s.fA[n][iMPHI] *= fabs(g->sp[n]);
          }
          for (int n = nbeg; n <= nend; n++) {
            int q[3] = {i, j, k};
            q[dir] = n;
            long o = q[2] * g->sk + q[1] * g->sj + q[0];
            double dtdV = dt / g->dV[o], dtdl;
            if (dir == 0) dtdl = dt / g->dx[0][n];
            else dtdl = dt / g->dx[dir][n] * g->dx_dl[dir][(long)q[1] * g->tot[0] + q[0]];
            for (int nv = 0; nv < nvar; nv++) s.rhs[n][nv] = -dtdV * (s.fA[n][nv] - s.fA[n - 1][nv]);
            s.rhs[n][VXn] -= dtdl * (s.press[n] - s.press[n - 1]);
            if (dir == 0) s.rhs[n][iMPHI] /= fabs(g->x[0][n]);
            else if (dir == 1 && c->geometry == SPHERICAL) s.rhs[n][iMPHI] /= fabs(g->s[n]);
          }
        }
        /* ---- RightHandSideSource ---- */
        for (int n = nbeg; n <= nend; n++) {
          int q[3] = {i, j, k};
          q[dir] = n;
          long o = q[2] * g->sk + q[1] * g->sj + q[0];
          double vc[NVMAX];
          const double *vg = s.v[n];
          if (c->geometry == SPHERICAL && dir == 0) {          /* rhs_source.c:229-241 */
            double r_1 = 1.0 / g->x[0][n];
            for (int nv = 0; nv < nvar; nv++) vc[nv] = 0.5 * (s.vp[n][nv] + s.vm[n][nv]);
            vg = vc;
            double vphi = vc[VX3];
            double Sm = vc[RHO] * (vc[VX2] * vc[VX2] + vphi * vphi);
            s.rhs[n][VX1] += dt * Sm * r_1;
          } else if ((c->geometry == CYLINDRICAL || c->geometry == POLAR) && dir == 0) {   /* rhs_source.c:201-227 */
            double r_1 = 1.0 / g->x[0][n];
            for (int nv = 0; nv < nvar; nv++) vc[nv] = 0.5 * (s.vp[n][nv] + s.vm[n][nv]);
            vg = vc;
            double vphi = vc[iMPHI];
            s.rhs[n][VX1] += dt * (vc[RHO] * vphi * vphi - 0.0) * r_1;
          } else if (c->geometry == SPHERICAL && dir == 1) {   /* rhs_source.c:311-357 */
            double r_1 = 1.0 / g->rt[i];
            double ct = 1.0 / tan(g->x[1][n]);
            for (int nv = 0; nv < nvar; nv++) vc[nv] = s.v[n][nv];
            vg = vc;
            double vphi = vc[VX3];
            double Sm = vc[RHO] * (-vc[VX2] * vc[VX1] + ct * vphi * vphi);
            s.rhs[n][VX2] += dt * Sm * r_1;
          }
          double gv[3];
          for (int pass = 0; pass < 3; pass++) {
            /* pass 0: BodyForceVector (rhs_source.c:253-272,360-376,428-440);
             * pass 1: BodyForcePotential (rhs_source.c:274-279,378-383,442-447);
             * pass 2: LineForce, same pattern as pass 0 (rhs_source.c:284-297,386-396,448-458) */
            if (pass == 1) {
              if (!(c->body_force & 2)) continue;
              double dtdx;
              if (dir == 0) dtdx = dt / g->dx[0][n];
              else if (dir == 1) {
                double scrh = dt;
                if (c->geometry == POLAR) scrh /= g->x[0][i];
                else if (c->geometry == SPHERICAL) scrh /= g->rt[i];
                dtdx = scrh / g->dx[1][n];
              } else {
                double scrh = dt;
                if (c->geometry == SPHERICAL) scrh *= g->dx[1][j] / (g->rt[i] * g->dmu[j]);
                dtdx = scrh / g->dx[2][n];
              }
              s.rhs[n][VXn] -= dtdx * vg[RHO] * (phi_p[n] - phi_p[n - 1]);
              if (!c->iso) s.rhs[n][PRS] -= c->bf_phi[0][o] * s.rhs[n][RHO];
              continue;
            }
            if (pass == 0) {
              if (!(c->body_force & 1)) continue;
              gv[0] = c->bf_g[0][o]; gv[1] = c->bf_g[1][o]; gv[2] = c->bf_g[2][o];
            } else {
              if (!c->ldw) continue;
              ldw_line_force(c, g, vg, dvds, o, gv);
            }
            const int en = !c->iso;   /* IF_ENERGY */
            s.rhs[n][VXn] += dt * vg[RHO] * gv[dir];
            if (en) s.rhs[n][PRS] += dt * 0.5 * (s.flux[n][RHO] + s.flux[n - 1][RHO]) * gv[dir];
            if (dir == 0 && c->ndim == 1) {
              s.rhs[n][VX2] += dt * vg[RHO] * gv[1];
              if (en) s.rhs[n][PRS] += dt * vg[RHO] * vg[VX2] * gv[1];
              s.rhs[n][VX3] += dt * vg[RHO] * gv[2];
              if (en) s.rhs[n][PRS] += dt * vg[RHO] * vg[VX3] * gv[2];
            }
            if (dir == 1 && c->ndim == 2) {
              s.rhs[n][VX3] += dt * vg[RHO] * gv[2];
              if (en) s.rhs[n][PRS] += dt * vg[RHO] * vg[VX3] * gv[2];
            }
          }
        }
        for (int n = nbeg; n <= nend; n++) {
          double *U = Uc + (base + n * st) * nvar;
          for (int nv = 0; nv < nvar; nv++) U[nv] += s.rhs[n][nv];
        }
        double ring_q = 1.0;        /* update_stage.c:305-311 */
        if (ring_line > 1) ring_q = 1.0 / ring_line;
        /* GetInverse_dl, set_geometry.c:303-375 */
        for (int n = 0; n < ntot; n++) {
          inv_dl[n] = g->inv_dx[dir][n];
          if ((c->geometry == SPHERICAL || c->geometry == POLAR) && dir == 1) inv_dl[n] = g->inv_dx[1][n] * (1.0 / g->x[0][i]);
          if (c->geometry == SPHERICAL && dir == 2) inv_dl[n] = g->inv_dx[2][n] * (1.0 / g->x[0][i]) / sin(g->x[1][j]);
        }
        if (c->ndim > 1) {
          if (stage == 1)
            for (int n = nbeg; n <= nend; n++)
              C_dt[base + n * st] += 0.5 * (s.cmax[n - 1] + s.cmax[n]) * inv_dl[n] * ring_q;
        } else {
          for (int n = nbeg - 1; n <= nend; n++) *invDt_hyp = MAXV(*invDt_hyp, s.cmax[n] * inv_dl[n]);
        }
      }
  }
  if (c->ndim > 1 && stage == 1) {
    for (int k = g->beg[2]; k <= g->end[2]; k++)
      for (int j = g->beg[1]; j <= g->end[1]; j++)
        for (int i = g->beg[0]; i <= g->end[0]; i++)
          *invDt_hyp = MAXV(*invDt_hyp, C_dt[k * g->sk + j * g->sj + i]);
    *invDt_hyp /= (double)c->ndim;
  }
  free(inv_dl);
  free(phi_p);
  sweep_free(&s);
}

/* ---------------------------------------------------------------------------------------
 *  AdvanceStep(): Time_Stepping/rk_step.c:29-322
 * --------------------------------------------------------------------------------------- */
typedef struct gen_ctx {
  gen_cfg c;
  geom_t *g;
  double *Uc, *U0, *C_dt, *dvds;
  uint16_t *flag;
} gen_ctx;

void *gen_create(const gen_cfg *c, const double *dx1, const double *dx2, const double *dx3) {
  gen_ctx *x = calloc(1, sizeof(gen_ctx));
  x->c = *c;
  x->g = geom_new(c);
  const double *dxin[3] = {dx1, dx2, dx3};
  geom_finish(c, x->g, dxin);
  long n = x->g->sv * x->g->nvar;
  x->Uc = calloc(n, 8); x->U0 = calloc(n, 8); x->C_dt = calloc(x->g->sv, 8);
  x->flag = calloc(x->g->sv, sizeof(uint16_t));
  if (c->ldw) x->dvds = calloc((long)c->nangles * x->g->sv, 8);
  return x;
}
void gen_destroy(void *p) {
  gen_ctx *x = p;
  geom_free(x->g);
  free(x->Uc); free(x->U0); free(x->C_dt); free(x->flag); free(x->dvds);
  free(x);
}
int gen_nvar(void *p) { return ((gen_ctx *)p)->g->nvar; }
void gen_boundary(void *p, double *Vc) { gen_ctx *x = p; boundary(&x->c, x->g, Vc); }
/* copies of derived geometry for the tests: which = 0 dV, 1..3 A[d] (without the -1 layer) */
void gen_get_geometry(void *p, int which, double *out) {
  gen_ctx *x = p;
  const geom_t *g = x->g;
  for (int k = 0; k < g->tot[2]; k++) for (int j = 0; j < g->tot[1]; j++) for (int i = 0; i < g->tot[0]; i++) {
    long o = k * g->sk + j * g->sj + i;
    out[o] = which == 0 ? g->dV[o] : A_at(g, which - 1, k, j, i);
  }
}

/* PPM coefficients of direction d for the tests: w[tot][4], hp[tot], hm[tot] */
void gen_get_ppm(void *p, int d, double *w, double *hp, double *hm) {
  gen_ctx *x = p;
  const geom_t *g = x->g;
  if (!g->pwp[d]) return;
  for (int i = 0; i < g->tot[d]; i++) {
    for (int j = 0; j < 4; j++) w[4 * i + j] = g->pwp[d][i][j];
    hp[i] = g->php[d][i];
    hm[i] = g->phm[d][i];
  }
}

int gen_advance_step(void *p, double *Vc, double dt, double *invDt_hyp, double *maxMach) {
  gen_ctx *x = p;
  const gen_cfg *c = &x->c;
  const geom_t *g = x->g;
  int nvar = g->nvar, nfail = 0;
  long ntot = g->sv * nvar;
  memset(x->flag, 0, g->sv * sizeof(uint16_t));          /* main.c:258-261 */
  double v[NVMAX];
  if (c->ring_average > 1) {      /* rk_step.c:115-119: PrimToCons3D, RingAverageCons, ConsToPrim3D before Boundary() */
    for (int k = g->beg[2]; k <= g->end[2]; k++)
      for (int j = g->beg[1]; j <= g->end[1]; j++)
        for (int i = g->beg[0]; i <= g->end[0]; i++) {
          long o = k * g->sk + j * g->sj + i;
          for (int nv = 0; nv < nvar; nv++) v[nv] = Vc[nv * g->sv + o];
          prim_to_cons(c, nvar, v, x->Uc + o * nvar);
        }
    ring_average_cons(c, g, x->Uc);
    for (int k = g->beg[2]; k <= g->end[2]; k++)
      for (int j = g->beg[1]; j <= g->end[1]; j++)
        for (int i = g->beg[0]; i <= g->end[0]; i++) {
          long o = k * g->sk + j * g->sj + i;
          nfail += cons_to_prim(c, nvar, x->Uc + o * nvar, v, &x->flag[o]);
          for (int nv = 0; nv < nvar; nv++) Vc[nv * g->sv + o] = v[nv];
        }
  }
  for (int stage = 1; stage <= c->rk; stage++) {
    g_Uc_for_floor = (stage > 1) ? x->Uc : NULL;    /* stage 1 re-derives all of Uc right after */
    boundary(c, g, Vc);
    g_Uc_for_floor = NULL;
    if (stage == 1) {
      if (c->flattening || c->entropy) flag_shock(c, g, Vc, x->flag);   /* rk_step.c:123-125 */
      for (int k = g->beg[2]; k <= g->end[2]; k++)
        for (int j = g->beg[1]; j <= g->end[1]; j++)
          for (int i = g->beg[0]; i <= g->end[0]; i++) {
            long o = k * g->sk + j * g->sj + i;
            for (int nv = 0; nv < nvar; nv++) v[nv] = Vc[nv * g->sv + o];
            prim_to_cons(c, nvar, v, x->Uc + o * nvar);
          }
      memcpy(x->U0, x->Uc, ntot * 8);
    }
    double hyp = 0.0;
    update_stage(c, g, Vc, x->Uc, x->flag, x->C_dt, x->dvds, dt, stage, &hyp, maxMach);
    *invDt_hyp = MAXV(*invDt_hyp, hyp);
    double w0 = 0, wc = 1;
    if (stage == 2) { w0 = c->rk == 2 ? 0.5 : 0.75; wc = c->rk == 2 ? 0.5 : 0.25; }
    for (int k = g->beg[2]; k <= g->end[2]; k++)
      for (int j = g->beg[1]; j <= g->end[1]; j++)
        for (int i = g->beg[0]; i <= g->end[0]; i++) {
          long o = k * g->sk + j * g->sj + i;
          double *U = x->Uc + o * nvar, *U0 = x->U0 + o * nvar;
          if (stage == 2) for (int nv = 0; nv < nvar; nv++) U[nv] = w0 * U0[nv] + wc * U[nv];
          if (stage == 3) for (int nv = 0; nv < nvar; nv++) U[nv] = (1.0 / 3.0) * (U0[nv] + 2.0 * U[nv]);
          if (c->ring_average > 1) continue;      /* RingAverageCons comes between the combination and ConsToPrim3D */
          nfail += cons_to_prim(c, nvar, U, v, &x->flag[o]);
          for (int nv = 0; nv < nvar; nv++) Vc[nv * g->sv + o] = v[nv];
        }
    if (c->ring_average > 1) {                    /* rk_step.c:167-169,238-240,306-308 */
      ring_average_cons(c, g, x->Uc);
      for (int k = g->beg[2]; k <= g->end[2]; k++)
        for (int j = g->beg[1]; j <= g->end[1]; j++)
          for (int i = g->beg[0]; i <= g->end[0]; i++) {
            long o = k * g->sk + j * g->sj + i;
            nfail += cons_to_prim(c, nvar, x->Uc + o * nvar, v, &x->flag[o]);
            for (int nv = 0; nv < nvar; nv++) Vc[nv * g->sv + o] = v[nv];
          }
    }
  }
  return nfail;
}

/* ---------------------------------------------------------------------------------------
 *  Line-driven wind (LINE_DRIVEN_WIND SIROCCO_MODE), Src/LineDriven/line_connect.c, and the user
 *  boundaries of Test_Problems/LineDrivenWind/cv_idl/init.c:157-316
 * --------------------------------------------------------------------------------------- */
#define CONST_amu 1.66053886e-24    /* pluto.h:299-321 */
#define CONST_c 2.99792458e10
#define CONST_G 6.6726e-8
#define CONST_kB 1.3806505e-16
#define CONST_mp 1.67262171e-24
#define CONST_PI 3.14159265358979
#define CONST_sigma 5.67051e-5
#define CONST_sigmaT 6.6524e-25

static double kelvin(const gen_cfg *c) { return c->unit_velocity * c->unit_velocity * CONST_amu / CONST_kB; }  /* pluto.h:560 */

/* UserDefBoundary(side == 0): density / pressure floors over TOT_LOOP and the mid-plane reset
 * (init.c:199-316).  Uc of a floored zone is re-derived like PrimToCons3D(d->Vc, d->Uc, 1-zone box). */
static void ldw_internal_floor(const gen_cfg *c, const geom_t *g, double *Vc) {
  int nvar = g->nvar, TRC = NF(c);
  const int en = !c->iso;      /* the #if EOS != ISOTHERMAL blocks of init.c */
  double KELVIN = kelvin(c), mu = c->mu;
  double dfloor = c->dfloor / c->unit_density;
  double rho_0 = c->rho0 / c->unit_density;
  double tfloor = 5.e2, pfloor = dfloor * tfloor / (KELVIN * mu);
  double r_WD = g->xl[0][g->beg[0]];                 /* g_domBeg[IDIR] */
  double gm_cgs = CONST_G * c->cent_mass;
  double gm_code = gm_cgs / (c->unit_length * c->unit_velocity * c->unit_velocity);
  const double *x1 = g->xgc[0], *x2 = g->xgc[1];
  int jlast = g->end[1];     /* j == np_int[JDIR] + 2 with the 3 ghost zones of this configuration */
  for (int k = 0; k < g->tot[2]; k++) for (int j = 0; j < g->tot[1]; j++) for (int i = 0; i < g->tot[0]; i++) {
    long o = k * g->sk + j * g->sj + i;
    double *rho = Vc + RHO * g->sv + o, *prs = Vc + PRS * g->sv + o;
    double *v1 = Vc + VX1 * g->sv + o, *v2 = Vc + VX2 * g->sv + o, *v3 = Vc + VX3 * g->sv + o;
    int convert = 0;
    if (*rho < dfloor) {
      if (*rho < 0.0) *rho = dfloor;
      double cs = en ? sqrt(c->gamma * *prs / *rho) : 0.0;
      double dfact = *rho / dfloor;
      *rho = dfloor;
      *v1 = dfact * *v1; *v2 = dfact * *v2; *v3 = dfact * *v3;
      if (en) {
        *prs = pow(cs, 2) * *rho / c->gamma;
        double temp = *prs / *rho * KELVIN * mu;
        if (temp < tfloor) { temp = tfloor; *prs = *rho * temp / (KELVIN * mu); }
      }
      Vc[TRC * g->sv + o] = 0.0;
      convert = 1;
    }
    if (en && *prs < pfloor) { *prs = pfloor; convert = 1; }
    if (convert && g_Uc_for_floor) {
      double v[NVMAX];
      for (int nv = 0; nv < nvar; nv++) v[nv] = Vc[nv * g->sv + o];
      prim_to_cons(c, nvar, v, g_Uc_for_floor + o * nvar);
    }
    if (j == jlast) {     /* the coordinate test of init.c:283-284 holds on a single-block grid */
      double r = x1[i], theta = x2[j], rcyl = r * sin(theta);
      double rho_mid = rho_0 * pow((r / r_WD), -1.0 * c->rho_alpha);
      *v2 = (*rho * *v2) / rho_mid;
      *rho = rho_mid;
      *v1 = 0.0;
      *v3 = sqrt(gm_code / r) * sin(theta);
      if (en) {
        double teff = pow(3.0 * gm_cgs * c->disk_mdot / (8.0 * CONST_PI * CONST_sigma), 0.25);
        teff *= pow(r_WD * c->unit_length, -0.75);
        double temp = teff * pow(r_WD / rcyl, 0.75) * pow(1.0 - sqrt(r_WD / rcyl), 0.25);
        *prs = rho_mid * temp / (KELVIN * mu);
      }
      Vc[TRC * g->sv + o] = 1.0;
    }
  }
}

/* UserDefBoundary(X1_BEG / X1_END / X2_BEG), init.c:319-363 */
static void ldw_userdef_side(const gen_cfg *c, const geom_t *g, double *Vc, int side) {
  int nvar = g->nvar;
  int IBEG = g->beg[0], IEND = g->end[0], JBEG = g->beg[1];
  for (int k = 0; k < g->tot[2]; k++) {
    if (side == 0) {          /* X1_BEG */
      for (int j = 0; j < g->tot[1]; j++) for (int i = 0; i < IBEG; i++) {
        long o = k * g->sk + j * g->sj + i, os = k * g->sk + j * g->sj + IBEG;
        for (int nv = 0; nv < nvar; nv++) Vc[nv * g->sv + o] = Vc[nv * g->sv + os];
        Vc[VX1 * g->sv + o] = MINV(Vc[VX1 * g->sv + o], 0.0);
      }
    } else if (side == 1) {   /* X1_END */
      for (int j = 0; j < g->tot[1]; j++) for (int i = IEND + 1; i < g->tot[0]; i++) {
        long o = k * g->sk + j * g->sj + i, os = k * g->sk + j * g->sj + IEND;
        for (int nv = 0; nv < nvar; nv++) Vc[nv * g->sv + o] = Vc[nv * g->sv + os];
        Vc[VX1 * g->sv + o] = MAXV(Vc[VX1 * g->sv + o], 0.0);
      }
    } else if (side == 2) {   /* X2_BEG: reflective velocities, outflow density and pressure */
      for (int j = 0; j < JBEG; j++) for (int i = 0; i < g->tot[0]; i++) {
        long o = k * g->sk + j * g->sj + i, os = k * g->sk + (2 * JBEG - j - 1) * g->sj + i;
        long ob = k * g->sk + JBEG * g->sj + i;
        for (int nv = 0; nv < nvar; nv++) Vc[nv * g->sv + o] = Vc[nv * g->sv + os];
        Vc[VX2 * g->sv + o] *= -1.0;
        Vc[RHO * g->sv + o] = Vc[RHO * g->sv + ob];
        if (!c->iso) Vc[PRS * g->sv + o] = Vc[PRS * g->sv + ob];
      }
    }
  }
}

/* bilinear(), line_connect.c:746-767 */
static void bilinear(const double x11[2], const double x22[2], const double v11[2], const double v12[2],
                     const double v21[2], const double v22[2], const double test[2], double ans[2]) {
  double fracx1 = (test[0] - x11[0]) / (x22[0] - x11[0]);
  double fracx2 = (test[1] - x11[1]) / (x22[1] - x11[1]);
  double temp1 = (1.0 - fracx1) * v11[0] + fracx1 * v21[0];
  double temp2 = (1.0 - fracx1) * v12[0] + fracx1 * v22[0];
  ans[0] = (1.0 - fracx2) * temp1 + fracx2 * temp2;
  temp1 = (1.0 - fracx1) * v11[1] + fracx1 * v21[1];
  temp2 = (1.0 - fracx1) * v12[1] + fracx1 * v22[1];
  ans[1] = (1.0 - fracx2) * temp1 + fracx2 * temp2;
}

/* VGradCalc(), line_connect.c:504-744: dvds[iangle][k][j][i] over the interior */
static void ldw_vgrad_calc(const gen_cfg *c, const geom_t *g, const double *Vc, double *dvds) {
  const double UL = c->unit_length, UV = c->unit_velocity;
  const double *x1 = g->x[0], *x2 = g->x[1];
  const double *V1 = Vc + VX1 * g->sv, *V2 = Vc + VX2 * g->sv;
  long sj = g->sj;
  for (int ia = 0; ia < c->nangles; ia++)
    for (int k = g->beg[2]; k <= g->end[2]; k++) for (int j = g->beg[1]; j <= g->end[1]; j++)
      for (int i = g->beg[0]; i <= g->end[0]; i++) {
        long o = k * g->sk + j * g->sj + i;
        double x11[2], x22[2], v11[2], v12[2], v22[2], v21[2], loc[2], ans1[2], ans2[2];
        x11[0] = (x1[i - 1] + x1[i]) / 2.0 * UL;
        x11[1] = (x2[j - 1] + x2[j]) / 2.0;
        x22[0] = (x1[i + 1] + x1[i]) / 2.0 * UL;
        x22[1] = (x2[j + 1] + x2[j]) / 2.0;
        double maxds = fabs(x22[0] - x11[0]);
        if (maxds > (x1[i] * UL * fabs(x22[1] - x11[1]))) maxds = x1[i] * UL * fabs(x22[1] - x11[1]);
        maxds /= 2.0;
        double fr = c->flux_r[ia * g->sv + o], ft = c->flux_t[ia * g->sv + o], fp = c->flux_p[ia * g->sv + o];
        double mod_flux = sqrt(pow(fr, 2) + pow(ft, 2) + pow(fp, 2));
        double r_off = 0, t_off = 0, ds;
        double theta_angle = (ia + 0.5) * (2.0 * CONST_PI) / 36.0;
        if (mod_flux == 0.0) ds = -999;
        else {
          double x = x1[i] * sin(x2[j]) * UL, z = x1[i] * cos(x2[j]) * UL;
          double dx1 = maxds * sin(theta_angle), dx2 = maxds * cos(theta_angle);
          ds = sqrt(dx1 * dx1 + dx2 * dx2);
          r_off = sqrt((x + dx1) * (x + dx1) + (z + dx2) * (z + dx2));
          t_off = atan((x + dx1) / (z + dx2));
        }
        v11[0] = (V1[o - sj - 1] + V1[o - sj] + V1[o - 1] + V1[o]) / 4.0;
        v11[1] = (V2[o - sj - 1] + V2[o - sj] + V2[o - 1] + V2[o]) / 4.0;
        v12[0] = (V1[o + sj - 1] + V1[o - 1] + V1[o + sj] + V1[o]) / 4.0;
        v12[1] = (V2[o + sj - 1] + V2[o - 1] + V2[o + sj] + V2[o]) / 4.0;
        v22[0] = (V1[o + sj] + V1[o + sj + 1] + V1[o + 1] + V1[o]) / 4.0;
        v22[1] = (V2[o + sj] + V2[o + sj + 1] + V2[o + 1] + V2[o]) / 4.0;
        v21[0] = (V1[o + 1] + V1[o - sj + 1] + V1[o - sj] + V1[o]) / 4.0;
        v21[1] = (V2[o + 1] + V2[o - sj + 1] + V2[o - sj] + V2[o]) / 4.0;
        loc[0] = x1[i] * UL;
        loc[1] = x2[j];
        bilinear(x11, x22, v11, v12, v21, v22, loc, ans1);
        double vx1 = (ans1[0] * UV * sin(x2[j]) + ans1[1] * UV * cos(x2[j]));
        double vz1 = (ans1[0] * UV * cos(x2[j]) - ans1[1] * UV * sin(x2[j]));
        double out;
        if (ds == -999) out = -999;
        else {
          loc[0] = r_off; loc[1] = t_off;
          bilinear(x11, x22, v11, v12, v21, v22, loc, ans2);
          double vx2 = (ans2[0] * UV * sin(loc[1]) + ans2[1] * UV * cos(loc[1]));
          double vz2 = (ans2[0] * UV * cos(loc[1]) - ans2[1] * UV * sin(loc[1]));
          double v1 = sin(theta_angle) * vx1 + cos(theta_angle) * vz1;
          double v2 = sin(theta_angle) * vx2 + cos(theta_angle) * vz2;
          out = fabs((v2 - v1) / ds);
        }
        dvds[ia * g->sv + o] = out;
      }
}

/* linterp(), line_connect.c:781-811 */
static double linterp(double x, const double *xarray, const double *yarray, int nelem) {
  int idx = 0;
  double result;
  while (idx < nelem && xarray[idx] < x) idx++;
  if (idx == 0) result = pow(10.0, yarray[0]);
  else if (idx >= nelem) result = pow(10.0, yarray[nelem - 1]);
  else {
    double x_low = xarray[idx - 1], x_high = xarray[idx], y_low = yarray[idx - 1], y_high = yarray[idx];
    double slope = (y_high - y_low) / (x_high - x_low);
    result = pow(10.0, y_low + slope * (x - x_low));
  }
  if (isnan(result)) result = 0.0;
  return result;
}

/* LineForce(), line_connect.c:815-903 (KRAD/ALPHARAD power law, or the per-zone M(t) fit when both are 999) */
static void ldw_line_force(const gen_cfg *c, const geom_t *g, const double *v, const double *dvds, long o, double *grad) {
  double sigma_e = CONST_sigmaT / CONST_amu / 1.18;
  double rho = v[RHO] * c->unit_density;
  double T = c->iso ? c->t_iso : v[PRS] / v[RHO] * kelvin(c) * c->mu;   /* line_connect.c:851-855 */
  double v_th = sqrt((2.0 * CONST_kB * T) / CONST_mp);
  double M_max = 4400.;
  double UNIT_ACC = c->unit_velocity * c->unit_velocity / c->unit_length;
  grad[0] = grad[1] = grad[2] = 0.0;
  int fit = (c->krad == 999 && c->alpharad == 999);
  double M_UV_array[64];
  if (fit) for (int ii = 0; ii < c->mpoints; ii++) M_UV_array[ii] = c->m_fit[ii * g->sv + o];
  for (int ia = 0; ia < c->nangles; ia++) {
    double flux_r = c->flux_r[ia * g->sv + o], flux_t = c->flux_t[ia * g->sv + o], M_UV;
    double dv = dvds[ia * g->sv + o];
    if (dv > 0.0) {
      double t_UV = sigma_e * rho * v_th / dv;
      if (fit) M_UV = linterp(log10(t_UV), c->t_fit, M_UV_array, c->mpoints);
      else M_UV = c->krad * pow(t_UV, c->alpharad);
      if (M_UV > M_max) M_UV = M_max;
    } else M_UV = 0.0;
    grad[0] += ((1.0 + M_UV) * sigma_e * flux_r / CONST_c) / UNIT_ACC;
    grad[1] += ((1.0 + M_UV) * sigma_e * flux_t / CONST_c) / UNIT_ACC;
  }
}

/* ---------------------------------------------------------------------------------------
 *  COOLING BLONDIN: BlondinCooling(), heatcool(), ne_rat(), zfunc(), zbrent()
 *  Src/Cooling/BLONDIN/cooling.c:50-330 (SplitSource, Src/split_source.c:53)
 * --------------------------------------------------------------------------------------- */
typedef struct {
  double comp_c_pre, comp_h_pre, line_c_pre, brem_c_pre, xray_h_pre;
  double nH, ne, xi, tx, sqxi, sqsqxi, n, E, hc_init, dt_share;
} cool_t;

static double ne_rat(double T) {
  if (T < 1.5e4) return 1e-2 + pow(10, (-51.59417133 + 12.27740153 * log10(T)));
  else if (T >= 1.5e4 && T < 3.3e4) return pow(10, (-3.80749689 + 0.86092628 * log10(T)));
  return 1.21;
}
static double heatcool(cool_t *q, double T) {
  double sqT = sqrt(T);
  q->ne = q->nH * ne_rat(T);
  double comp_heat = q->comp_h_pre * (8.9e-36 * q->xi * q->tx);
  double comp_cool = q->comp_c_pre * (8.9e-36 * q->xi * (4.0 * T));
  double xray_heat = q->xray_h_pre * (1.5e-21 * (q->sqsqxi / sqT));
  double line_cool = q->line_c_pre * ((1e-16 * exp(-1.3e5 / T) / q->sqxi / T) + fmin(fmin(1e-24, 5e-27 * sqT), 1.5e-17 / T));
  double brem_cool = q->brem_c_pre * (3.3e-27 * sqT);
  return q->nH * (q->ne * comp_heat + q->nH * xray_heat - q->ne * comp_cool - q->ne * line_cool - q->ne * brem_cool);
}
static double zfunc(cool_t *q, double temp) {
  return (temp * q->n * CONST_kB / (2.0 / 3.0)) - q->E - q->dt_share * (q->hc_init + heatcool(q, temp)) / 2.0;
}
static double zbrent(cool_t *z, int which, double x1, double x2, double tol) {
#define FUNC(x) (which ? zfunc(z, (x)) : heatcool(z, (x)))
  const double EPS = 3.0e-8;
  double a = x1, b = x2, c = x2, d = 0.0, e = 0.0;
  double fa = FUNC(a), fb = FUNC(b), fc = fb, p, q, r, s, tol1, xm;
  if (fb * fa > 0.0) return b;
  for (int iter = 1; iter <= 100; iter++) {
    if (fb * fc > 0.0) { c = a; fc = fa; e = d = b - a; }
    if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    tol1 = 2.0 * EPS * fabs(b) + 0.5 * tol;
    xm = 0.5 * (c - b);
    if (fabs(xm) <= tol1 || fb == 0.0) return b;
    if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
      s = fb / fa;
      if (a == c) { p = 2.0 * xm * s; q = 1.0 - s; }
      else {
        q = fa / fc; r = fb / fc;
        p = s * (2.0 * xm * q * (q - r) - (b - a) * (r - 1.0));
        q = (q - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q = -q;
      p = fabs(p);
      double min1 = 3.0 * xm * q - fabs(tol1 * q), min2 = fabs(e * q);
      if (2.0 * p < (min1 < min2 ? min1 : min2)) { e = d; d = p / q; }
      else { d = xm; e = d; }
    } else { d = xm; e = d; }
    a = b; fa = fb;
    if (fabs(d) > tol1) b += d;
    else b += (xm > 0.0 ? fabs(tol1) : -fabs(tol1));
    fb = FUNC(b);
  }
  return b;
#undef FUNC
}

/* tabs: comp_h_pre, comp_c_pre, xray_h_pre, line_c_pre, brem_c_pre, sirocco_xi, sirocco_t_r  [k][j][i] */
void gen_blondin_cooling(void *p, double *Vc, double dt, double g_time, const double *const tabs[7]) {
  gen_ctx *x = p;
  const gen_cfg *c = &x->c;
  const geom_t *g = x->g;
  double UNIT_TIME = c->unit_length / c->unit_velocity;
  double UNIT_PRESSURE = c->unit_density * c->unit_velocity * c->unit_velocity;
  double KELVIN = kelvin(c), mu = c->mu, lx = c->lx;
  double minT = 1.e4;
  cool_t q;
  q.dt_share = dt * UNIT_TIME;
  for (int k = g->beg[2]; k <= g->end[2]; k++) for (int j = g->beg[1]; j <= g->end[1]; j++)
    for (int i = g->beg[0]; i <= g->end[0]; i++) {
      long o = k * g->sk + j * g->sj + i;
      q.comp_h_pre = tabs[0][o]; q.comp_c_pre = tabs[1][o]; q.xray_h_pre = tabs[2][o];
      q.line_c_pre = tabs[3][o]; q.brem_c_pre = tabs[4][o];
      double r = g->x[0][i] * c->unit_length;
      double rho = Vc[RHO * g->sv + o] * c->unit_density;
      double pr = Vc[PRS * g->sv + o];
      double T = pr / Vc[RHO * g->sv + o] * KELVIN * mu;
      q.E = (pr * UNIT_PRESSURE) / (c->gamma - 1);
      q.nH = rho / (1.43 * CONST_mp);
      if (g_time <= 3.0) { q.xi = lx / q.nH / r / r; q.tx = c->tx; }
      else { q.xi = tabs[5][o]; q.tx = tabs[6][o]; }
      q.n = rho / (mu * CONST_mp);
      T = q.E * (2.0 / 3.0) / (q.n * CONST_kB);
      if (T < minT) continue;
      q.sqxi = sqrt(q.xi);
      q.sqsqxi = pow(q.xi, 0.25);
      q.hc_init = heatcool(&q, T);
      double t_l = T * 0.9, t_u = T * 1.1, T_f;
      double test = zfunc(&q, t_l) * zfunc(&q, t_u);
      while (test > 0 && test == test) { t_l *= 0.9; t_u *= 1.1; test = zfunc(&q, t_l) * zfunc(&q, t_u); }
      if (test != test) T_f = T;
      else {
        T_f = zbrent(&q, 1, t_l, t_u, 1.0);
        double hc_final = heatcool(&q, T_f);
        if (hc_final * q.hc_init < 0.0) T_f = zbrent(&q, 0, fmin(T_f, T), fmax(T_f, T), 1.0);
      }
      T_f = MAXV(T_f, minT);
      double E_f = T_f / (2.0 / 3.0) * (q.n * CONST_kB);
      Vc[PRS * g->sv + o] = E_f * (c->gamma - 1) / UNIT_PRESSURE;
    }
}
