// gen_kernels.cuh -- the GENERAL-GRID path of the HD update: curvilinear geometry, non-uniform
// grids, characteristic limiting, MULTID shock flattening, the entropy switch, body forces and
// the line-driven-wind sources.
//
// The grids this path serves are small (the line-driven-wind problem is 1024 x 512 zones, 30 MB
// of state), so every intermediate array stays resident in the 126 MB L2 and the update is split
// into simple, one-thread-per-zone (or per-face) kernels that mirror the reference's stages:
//   gen_entropy   ComputeEntropy           Src/entropy_switch.c:14-39 (after Boundary)
//   gen_shock / gen_flags   FlagShock      Src/flag_shock.c:81-260 (gather form of the scatter)
//   gen_p2c       PrimToCons3D + U0 copy   Src/Time_Stepping/rk_step.c:129-130
//   gen_states    States (PLM)             Src/States/plm_states.c:83-337 and :481-690
//   gen_riemann   Riemann solver + AdvectFlux  Src/HD/{hllc,hll,tvdlf}.c, Src/adv_flux.c:47-134
//   gen_rhs       RightHandSide + Source + U += rhs + C_dt   Src/MHD/rhs.c:84-420,
//                 Src/MHD/rhs_source.c:101-470, Src/Time_Stepping/update_stage.c:283,303-322
//   gen_finish    RK combination + ConsToPrim3D + dt reduction  rk_step.c:235-237,304,
//                 Src/HD/mappers.c:98-290, update_stage.c:389-392
// Arrays are SoA [var][k][j][i] like d->Vc, threads map to i fastest: every access is coalesced.
// The sweep direction is a RUN-TIME argument (only global-memory indices depend on it); the
// number of variables is a template parameter so that per-zone vectors live in registers.
#pragma once
#include <type_traits>

#include "pb200_kernels.cuh"

// gen_vgrad: active bins per loop iteration and the resident 64-thread blocks per SM its register budget is cut for.
// Measured on C4, one box, ms per step (profiles/r02_v4_gen_tiles.txt): 1 bin 0.820, 2 bins / 128 registers 0.801,
// 3 bins / 128 registers 0.806, 3 bins / 188 registers (uncapped: 5 blocks per SM) 0.928; loop before: 0.880
#ifndef PB_VGRAD_NB
#define PB_VGRAD_NB 2
#endif
#ifndef PB_VGRAD_MINB
#define PB_VGRAD_MINB 8
#endif

namespace pb {

enum { GF_MINMOD = 1, GF_FLAT = 2, GF_HLL = 4, GF_ENTROPY = 8, GF_C2P_FAIL = 64 };   // pluto.h:212-221
enum { GEO_CARTESIAN = 1, GEO_CYLINDRICAL = 2, GEO_POLAR = 3, GEO_SPHERICAL = 4 };    // Src/pluto.h:34-37

// LINE_DRIVEN_WIND SIROCCO_MODE (Src/LineDriven/line_connect.c) + the user boundaries of
// Test_Problems/LineDrivenWind/cv_idl/init.c
struct LdwDev {
  int on, userdef_bc, nangles;
  const double *flux_r, *flux_t, *flux_p;   // [nangles][k][j][i]
  const unsigned long long *mask;           // bit a of zone o: |flux| != 0 in angular bin a (gen_ldw_mask)
  double *gline;                            // line force [2][k][j][i]: g_r for the r sweep, g_theta for the theta sweep
  const double *sin_a, *cos_a;              // sin/cos((a + 1/2) 2 pi / 36), libm values from the host
  const double *inv_ca;                     // 1 / sqrt(sin_a^2 + cos_a^2) (the bin's part of 1/ds, line_connect.c:566)
  const double *sin_t, *cos_t;              // sin/cos(x2[j])
  const double *xgc1, *xgc2;                // grid->xgc (mid-plane reset uses the centroids)
  double UL, UV, UD;                        // UNIT_LENGTH, UNIT_VELOCITY, UNIT_DENSITY
  double kelvin_mu;                         // KELVIN * mu
  double krad, alpharad;
  int alpha_m06;                            // ALPHARAD == -0.6 (the value cv_idl ships): dvds^0.6 through pow_three_fifths()
  double t_iso;                             // > 0: EOS ISOTHERMAL, the temperature of LineForce() is g_inputParam[T_ISO]
  int mpoints;                              // > 0: force multiplier from the per-zone M(t) fit (KRAD = ALPHARAD = 999)
  const double *t_fit, *m_fit;              // log10(t) [mpoints]; log10(M) [mpoints][k][j][i]
  double sigma_e, unit_acc;                 // sigma_T/amu/1.18 ; UNIT_ACCELERATION
  double dfloor, pfloor, tfloor, rho_0, rho_alpha, r_WD, gm_code, teff_wd;   // init.c:175-197,296-300
};

struct GenDev {
  Dev d;
  int nvar, geometry, limiter, char_lim, flatten, entropy, solver;
  // EOS ISOTHERMAL (Src/EOS/Isothermal/eos.c): no energy equation, NFLX = 4, scalars start at index 4, p = cs2 rho
  int iso;
  double cs2;                          // g_isoSoundSpeed^2
  int flatten_oned;                    // SHOCK_FLATTENING ONED (States/flatten.c); `flatten` is MULTID
  // RECONSTRUCTION PARABOLIC (PPM_ORDER 4): interface weights [tot][4] and h+ / h- per direction (States/ppm_coeffs.c)
  int ppm;
  const double *pw[3], *php[3], *phm[3];
  // RING_AVERAGE (Src/ring_average.c): chunk size at the axis (> 1: on), RING_AVERAGE_REC 1 / 2 / 5 and
  // grid->ring_av_csize[] per i (POLAR) or j (SPHERICAL)
  int ring_average, ring_rec;
  const int *csize;
  // 1-D grid arrays per direction (np_tot entries): grid->x, xr, dx, inv_dx and PLM_Coeffs
  const double *x[3], *xr[3], *dx[3], *inv_dx[3];
  const double *cp[3], *cm[3], *wp[3], *wm[3], *dp[3], *dm[3];
  const double *rt, *s, *sp;           // grid->rt[i], s[j], sp[j]
  const double *cot, *sin2;            // 1/tan(x2[j]) and sin(x2[j]) with the host's libm (rhs_source.c:350, set_geometry.c:344)
  const double *dV;                    // grid->dV[k][j][i]
  const double *A[3];                  // grid->A[d], one extra layer at index -1 along d
  long Aoff[3], Asj[3], Ask[3];
  const double *dx_dl[3];              // grid->dx_dl[d][j][i]
  LdwDev ldw;
};

struct GenArgs {
  double *V;             // d->Vc (updated in place like the reference)
  double *U, *U0;        // d->Uc, U0 (SoA)
  double *VP, *VM;       // stateL.v = vp, stateR.v - 1 = vm of the current direction
  double *F;             // sweep.flux [nvar] + press + cmax  (nvar + 2 arrays)
  double *cdt;           // C_dt
  unsigned short *flag;  // d->flag
  unsigned char *shock;  // zones with div v < 0 and grad p > 5 p_min
  const double *dt;
  unsigned long long *red;
  double w0, wc;
  int comb, stage, dir;
  const unsigned char *ibmask;   // FLAG_INTERNAL_BOUNDARY zones: rhs = 0 (int_bound_reset.c:34-35), null: none
  // fused sweeps with the line-driven wind: the r sweep leaves its centre state (rho, p) and mean mass flux per zone
  // in cen[3][zones]; gen_vgrad takes the line force from them and the theta sweep adds the r sweep's force terms
  // (the force of a sweep needs the states of that sweep, and the r sweep's states are not stored any more)
  double *cen;
  int defer;
};

// index of the pressure in a per-zone vector of NV variables; with EOS ISOTHERMAL there is none (the expressions that
// use it are never evaluated then; the clamp only keeps the constant index inside the array for NV = 4)
template <int NV> __host__ __device__ constexpr int pidx() { return NV > 4 ? 4 : 0; }
PB_D int gen_nflx(const GenDev &g) { return g.iso ? 4 : 5; }

PB_D double gen_A(const GenDev &g, int dir, int k, int j, int i) {
  return __ldg(g.A[dir] + g.Aoff[dir] + (long)k * g.Ask[dir] + (long)j * g.Asj[dir] + i);
}

// ---- limiters on general grids (States/plm_coeffs.h:72-152) ------------------------------
PB_D double gen_lim(int kind, bool uniform, double dvp, double dvm, double cp, double cm) {
  if (!(dvp * dvm > 0.0)) return 0.0;
  switch (kind) {
    case LIM_MINMOD: return absmin(dvp, dvm);
    case LIM_VANLEER:
      if (uniform) return 2.0 * dvp * dvm / (dvp + dvm);
      return dvp * dvm * (cp * dvm + cm * dvp) / (dvp * dvp + dvm * dvm + (cp + cm - 2.0) * dvp * dvm);
    case LIM_MC: {
      double qc = 0.5 * (dvm + dvp);
      double scrh = uniform ? 2.0 * absmin(dvp, dvm) : absmin(dvp * cp, dvm * cm);
      return absmin(qc, scrh);
    }
    case LIM_VANALBADA: {
      double pp = dvp * dvp, mm = dvm * dvm;
      return (dvp * (mm + 1.e-18) + dvm * (pp + 1.e-18)) / (pp + mm + 1.e-18);
    }
    case LIM_OSPRE:
      if (uniform) return 1.5 * dvp * dvm * (dvm + dvp) / (dvp * dvp + dvm * dvm + dvp * dvm);
      return dvp * dvm * ((1.0 + cp) * dvm + (1.0 + cm) * dvp) /
             (2.0 * dvp * dvp + 2.0 * dvm * dvm + (cp + cm - 2.0) * dvp * dvm);
    case LIM_UMIST: {
      double ddp = 0.25 * (dvp + 3.0 * dvm), ddm = 0.25 * (dvm + 3.0 * dvp);
      double d2 = 2.0 * absmin(dvp, dvm);
      d2 = absmin(d2, ddp);
      return absmin(d2, ddm);
    }
    case 8: {  // SET_GM_LIMITER
      double qc = 0.5 * (dvm + dvp), scrh = absmin(dvp * cp, dvm * cm);
      return absmin(qc, scrh);
    }
    default: return 0.0;   // LIM_FLAT
  }
}

// zone (i,j,k) of a launch over the box [lo, hi] (inclusive), i fastest
PB_D bool gen_zone(const int lo[3], const int hi[3], int &i, int &j, int &k) {
  long n1 = hi[0] - lo[0] + 1, n2 = hi[1] - lo[1] + 1, n3 = hi[2] - lo[2] + 1;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1 * n2 * n3) return false;
  i = lo[0] + (int)(t % n1);
  j = lo[1] + (int)((t / n1) % n2);
  k = lo[2] + (int)(t / (n1 * n2));
  return true;
}
struct GenBox { int lo[3], hi[3]; };

// ---- ComputeEntropy over the whole array ------------------------------------------------
static __global__ void gen_entropy(GenDev g, double *V) {
  long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= g.d.sv) return;
  const int ENTR = g.nvar - 1;
  V[ENTR * g.d.sv + o] = V[iPRS * g.d.sv + o] / pow(V[o], g.d.gas.gamma);   // eos.c:95
}

// ---- FlagShock, gather form: (a) shock indicator per zone, (b) flags from the neighbours ---
static __global__ void gen_shock(GenDev g, GenArgs a, GenBox b) {
  int i, j, k;
  if (!gen_zone(b.lo, b.hi, i, j, k)) return;
  const Dev &d = g.d;
  const long o = (long)k * d.sk + (long)j * d.sj + i;
  const long st[3] = {1, d.sj, d.sk};
  const int idx[3] = {i, j, k};
  // EOS ISOTHERMAL: pt = cs2 rho (flag_shock.c:138-139); the same factor on every zone of the stencil
  auto pt = [&](long q) { return g.iso ? a.V[q] * g.cs2 : a.V[iPRS * d.sv + q]; };
  double divv = 0.0;
  for (int dir = 0; dir < d.ndim; dir++) {
    const double *vx = a.V + (1 + dir) * d.sv;
    double dv;
    if (g.geometry == GEO_CARTESIAN) dv = (vx[o + st[dir]] - vx[o - st[dir]]) / __ldg(g.dx[dir] + idx[dir]);
    else {
      int m[3] = {i, j, k};
      m[dir] -= 1;
      dv = gen_A(g, dir, k, j, i) * (vx[o + st[dir]] + vx[o]) - gen_A(g, dir, m[2], m[1], m[0]) * (vx[o - st[dir]] + vx[o]);
    }
    divv = (dir == 0) ? dv : divv + dv;
  }
  if (g.geometry != GEO_CARTESIAN) divv = divv / __ldg(g.dV + o);
  unsigned char sh = 0;
  if (divv < 0.0) {
    double pt_min = pt(o), gradp = 0.0;
    for (int dir = 0; dir < d.ndim; dir++) {
      pt_min = fmin(pt_min, fmin(pt(o + st[dir]), pt(o - st[dir])));
      double dp = fabs(pt(o + st[dir]) - pt(o - st[dir]));
      gradp = (dir == 0) ? dp : gradp + dp;
    }
    sh = (gradp > 5.0 * pt_min) ? 1 : 0;      // EPS_PSHOCK_FLATTEN, flag_shock.c:69-70
    if (gradp > 0.05 * pt_min) sh |= 2;       // EPS_PSHOCK_ENTROPY, flag_shock.c:73-74 (ENTROPY_SWITCH SELECTIVE)
  }
  a.shock[o] = sh;
}

static __global__ void gen_flags(GenDev g, GenArgs a) {
  const Dev &d = g.d;
  long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= d.sv) return;
  int i = (int)(o % d.tot[0]), j = (int)((o / d.sj) % d.tot[1]), k = (int)(o / d.sk);
  const long st[3] = {1, d.sj, d.sk};
  const int idx[3] = {i, j, k};
  unsigned short f = g.entropy ? GF_ENTROPY : 0;
  unsigned char near = a.shock[o];         // own indicator bits and the neighbours' (the reference scatters)
  for (int dir = 0; dir < d.ndim; dir++) {
    if (idx[dir] > 0) near |= a.shock[o - st[dir]];
    if (idx[dir] < d.tot[dir] - 1) near |= a.shock[o + st[dir]];
  }
  if (g.flatten) {
    if (a.shock[o] & 1) f |= GF_HLL | GF_MINMOD;
    if (near & 1) f |= GF_MINMOD;
  }
  if (g.entropy == 1 && (near & 2)) f &= ~GF_ENTROPY;    // SELECTIVE: unflag shocked zones and their neighbours
  a.flag[o] = f;
}

// ---- PrimToCons3D + RBoxCopy(U0) over the interior --------------------------------------
template <int NV>
static __global__ void gen_p2c(GenDev g, GenArgs a, GenBox b) {
  int i, j, k;
  if (!gen_zone(b.lo, b.hi, i, j, k)) return;
  const Dev &d = g.d;
  const long o = (long)k * d.sk + (long)j * d.sj + i;
  double v[NV], u[NV];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) v[nv] = a.V[nv * d.sv + o];
  // exact restatement (no reciprocal sharing): these values seed U0
  const double rho = v[iRHO];
  u[0] = rho; u[1] = rho * v[1]; u[2] = rho * v[2]; u[3] = rho * v[3];
  double k2 = v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  if (!g.iso) u[pidx<NV>()] = 0.5 * rho * k2 + v[pidx<NV>()] / d.gas.gmm1;
#pragma unroll
  for (int nv = 4; nv < NV; nv++) if (nv >= gen_nflx(g)) u[nv] = rho * v[nv];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) { a.U[nv * d.sv + o] = u[nv]; a.U0[nv * d.sv + o] = u[nv]; }
}

// ---- States: PLM on general grids, primitive or characteristic limiting -------------------
// vp / vm of ONE zone (States(), plm_states.c): shared by gen_states (one kernel per reference stage) and by
// the fused sweep kernel gen_sweep / gen_vgrad.  n = the zone's index along dir, st = its stride, o = its offset.
// Flatten(), States/flatten.c:58-130 (HD: EPS2 0.33, OME1 0.75, OME2 10), zones max(beg,3)..min(end,tot-4): the
// states are pulled towards the zone value by f = max(f_t[i], f_t[i + s]), s pointing down the pressure gradient
template <int NV>
PB_D void gen_flatten_oned(const GenDev &g, const double *__restrict__ V, int dir, int n, long st, long o,
                           const double (&v)[NV], double (&vpo)[NV], double (&vmo)[NV]) {
  const Dev &d = g.d;
  if (!(g.flatten_oned && n >= 3 && n <= d.tot[dir] - 4)) return;
  const double *Pv = V + (g.iso ? 0 : iPRS) * d.sv, *Vn = V + (1 + dir) * d.sv;
  auto f_t = [&](long q) {
    const double dpq = Pv[q + st] - Pv[q - st];
    const double min_p = fmin(Pv[q + st], Pv[q - st]);
    const double d2p = Pv[q + 2 * st] - Pv[q - 2 * st];
    double scrh = fabs(dpq) / min_p;
    if (scrh < 0.33 || (Vn[q + st] > Vn[q - st])) return 0.0;
    scrh = 10.0 * (fabs(dpq / d2p) - 0.75);
    scrh = fmin(1.0, scrh);
    return fmax(0.0, scrh);
  };
  const long sj = (Pv[o + st] < Pv[o - st]) ? st : -st;
  const double fj = fmax(f_t(o), f_t(o + sj));
  const double om = 1.0 - fj;
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    const double vf = v[nv] * fj;
    vmo[nv] = vf + vmo[nv] * om;
    vpo[nv] = vf + vpo[nv] * om;
  }
}

// PPM states of zone n (PPM_ORDER 4, PARABOLIC_LIM 1): States/ppm_states.c:66-232 (CHAR_LIMITING NO) and :280-590
// (CHAR_LIMITING YES).  v+ of a zone needs the interface values at n+1/2 (zones n-1..n+2) and n-1/2 (n-2..n+1).
PB_D double gen_minmod(double a, double b) { return a * b > 0.0 ? (fabs(a) < fabs(b) ? a : b) : 0.0; }

template <int NV>
PB_D void gen_prim_to_char(const GenDev &g, int dir, const double (&v)[NV], double cs, const double (&dv)[NV], double (&w)[NV]) {
  // PrimToChar (HD/eigenv.c:575-616) with the left eigenvectors of the zone (eigenv.c:92-200)
  constexpr int P = pidx<NV>();
  const double n_ = dir == 0 ? dv[1] : (dir == 1 ? dv[2] : dv[3]);
  const double t_ = dir == 0 ? dv[2] : (dir == 1 ? dv[3] : dv[1]);
  const double b_ = dir == 0 ? dv[3] : (dir == 1 ? dv[1] : dv[2]);
  if (!g.iso) {
    const double L0p = 1.0 / (v[iRHO] * cs), L2p = -1.0 / (cs * cs);
    w[0] = -1.0 * n_ + L0p * dv[P];
    w[1] = 1.0 * n_ + L0p * dv[P];
    w[2] = dv[iRHO] + L2p * dv[P];
    w[3] = t_;
    if (NV > 4) w[pidx<NV>()] = b_;
  } else {
    const double Lr = 1.0 / (v[iRHO] / cs);
    w[0] = Lr * dv[iRHO] + -1.0 * n_;
    w[1] = Lr * dv[iRHO] + 1.0 * n_;
    w[2] = t_;
    w[3] = b_;
  }
  const int nf = gen_nflx(g);
#pragma unroll
  for (int nv = 4; nv < NV; nv++) if (nv >= nf) w[nv] = dv[nv];
}

template <int NV>
PB_D void gen_zone_states_ppm(const GenDev &g, const double *__restrict__ V, const unsigned short *__restrict__ flag, int dir,
                              int n, long st, long o, double (&v)[NV], double (&vpo)[NV], double (&vmo)[NV]) {
  const Dev &d = g.d;
  constexpr int P = pidx<NV>();
  const double *wq = g.pw[dir] + 4 * (long)n;
  const double a0 = __ldg(wq - 4), a1 = __ldg(wq - 3), a2 = __ldg(wq - 2), a3 = __ldg(wq - 1);   // interface n-1/2
  const double b0 = __ldg(wq), b1 = __ldg(wq + 1), b2 = __ldg(wq + 2), b3 = __ldg(wq + 3);       // interface n+1/2
  const double hp = __ldg(g.php[dir] + n), hm = __ldg(g.phm[dir] + n);
  const double cm = (hm + 1.0) / (hp - 1.0), cp = (hp + 1.0) / (hm - 1.0);
  const unsigned short fl = g.flatten ? flag[o] : 0;
  double dfp[NV], dfm[NV], ip[NV], im[NV], vl1[NV];   // v(n+1) - v(n), v(n) - v(n-1), 4th-order interface values, v(n-1)
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    const double *q = V + nv * d.sv + o;
    const double m2 = q[-2 * st], m1 = q[-st], c0 = q[0], p1 = q[st], p2 = q[2 * st];
    v[nv] = c0;
    vl1[nv] = m1;
    dfp[nv] = p1 - c0;
    dfm[nv] = c0 - m1;
    ip[nv] = b0 * m1 + b1 * c0 + b2 * p1 + b3 * p2;
    im[nv] = a0 * m2 + a1 * m1 + a2 * c0 + a3 * p1;
  }
  if (fl & GF_FLAT) {
#pragma unroll
    for (int nv = 0; nv < NV; nv++) vpo[nv] = vmo[nv] = v[nv];
  } else if (fl & GF_MINMOD) {                  // PLM weights of plm_coeffs.c on the flagged zones
    const bool uniform = g.geometry == GEO_CARTESIAN;
    const double wp = uniform ? 1.0 : __ldg(g.wp[dir] + n), wm = uniform ? 1.0 : __ldg(g.wm[dir] + n);
    const double dp = uniform ? 0.5 : __ldg(g.dp[dir] + n), dm = uniform ? 0.5 : __ldg(g.dm[dir] + n);
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      const double dv = gen_minmod(dfp[nv] * wp, dfm[nv] * wm);
      vpo[nv] = v[nv] + dv * dp;
      vmo[nv] = v[nv] - dv * dm;
    }
  } else if (!g.char_lim) {
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      // interface values clipped between the two cell averages (ppm_states.c:118-126), then the parabola limiter
      const double m1 = vl1[nv];
      const double vr = v[nv] + gen_minmod(ip[nv] - v[nv], dfp[nv]);
      const double vl = m1 + gen_minmod(im[nv] - m1, dfm[nv]);
      double dvp = vr - v[nv], dvm = vl - v[nv];
      if (dvp * dvm >= 0.0) dvp = dvm = 0.0;
      else if (fabs(dvp) >= cm * fabs(dvm)) dvp = -cm * dvm;
      else if (fabs(dvm) >= cp * fabs(dvp)) dvm = -cp * dvp;
      vpo[nv] = v[nv] + dvp;
      vmo[nv] = v[nv] + dvm;
    }
  } else {
    const double a2s = g.iso ? g.cs2 : d.gas.gamma * v[P] / v[iRHO];
    const double cs = sqrt(a2s), rhocs = v[iRHO] * cs, rho_cs = v[iRHO] / cs;
    double dvp[NV], dvm[NV], dwp[NV], dwm[NV], dwp1[NV], dwm1[NV];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) { dvp[nv] = ip[nv] - v[nv]; dvm[nv] = im[nv] - v[nv]; }
    gen_prim_to_char<NV>(g, dir, v, cs, dvp, dwp);
    gen_prim_to_char<NV>(g, dir, v, cs, dvm, dwm);
    gen_prim_to_char<NV>(g, dir, v, cs, dfm, dwm1);
    gen_prim_to_char<NV>(g, dir, v, cs, dfp, dwp1);
#pragma unroll
    for (int q = 0; q < NV; q++) {
      dwp[q] = gen_minmod(dwp[q], dwp1[q]);
      dwm[q] = gen_minmod(dwm[q], -dwm1[q]);
    }
    // dv = R dw with the right eigenvectors of eigenv.c:140-178
    double pn, pt_, pb_, mn, mt_, mb_;
    if (!g.iso) {
      dvp[iRHO] = (dwp[0] * (0.5 * rho_cs) + dwp[1] * (0.5 * rho_cs)) + dwp[2];
      dvm[iRHO] = (dwm[0] * (0.5 * rho_cs) + dwm[1] * (0.5 * rho_cs)) + dwm[2];
      dvp[P] = dwp[0] * (0.5 * rhocs) + dwp[1] * (0.5 * rhocs);
      dvm[P] = dwm[0] * (0.5 * rhocs) + dwm[1] * (0.5 * rhocs);
      pn = dwp[0] * -0.5 + dwp[1] * 0.5; pt_ = dwp[3]; pb_ = dwp[P];
      mn = dwm[0] * -0.5 + dwm[1] * 0.5; mt_ = dwm[3]; mb_ = dwm[P];
    } else {
      dvp[iRHO] = dwp[0] * (0.5 * rho_cs) + dwp[1] * (0.5 * rho_cs);
      dvm[iRHO] = dwm[0] * (0.5 * rho_cs) + dwm[1] * (0.5 * rho_cs);
      pn = dwp[0] * -0.5 + dwp[1] * 0.5; pt_ = dwp[2]; pb_ = dwp[3];
      mn = dwm[0] * -0.5 + dwm[1] * 0.5; mt_ = dwm[2]; mb_ = dwm[3];
    }
    dvp[1] = dir == 0 ? pn : (dir == 1 ? pb_ : pt_);
    dvp[2] = dir == 0 ? pt_ : (dir == 1 ? pn : pb_);
    dvp[3] = dir == 0 ? pb_ : (dir == 1 ? pt_ : pn);
    dvm[1] = dir == 0 ? mn : (dir == 1 ? mb_ : mt_);
    dvm[2] = dir == 0 ? mt_ : (dir == 1 ? mn : mb_);
    dvm[3] = dir == 0 ? mb_ : (dir == 1 ? mt_ : mn);
    const int nf = gen_nflx(g);
#pragma unroll
    for (int nv = 4; nv < NV; nv++) if (nv >= nf) { dvp[nv] = dwp[nv]; dvm[nv] = dwm[nv]; }
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      if (dvp[nv] * dvm[nv] >= 0.0) dvp[nv] = dvm[nv] = 0.0;
      else if (fabs(dvp[nv]) >= cm * fabs(dvm[nv])) dvp[nv] = -cm * dvm[nv];
      else if (fabs(dvm[nv]) >= cp * fabs(dvp[nv])) dvm[nv] = -cp * dvp[nv];
      vpo[nv] = v[nv] + dvp[nv];
      vmo[nv] = v[nv] + dvm[nv];
    }
    if (vpo[iRHO] < 0.0 || vmo[iRHO] < 0.0) {       // ppm_states.c:560-575: back to a minmod slope
      const double h = 0.5 * gen_minmod(dfp[iRHO], dfm[iRHO]);
      vpo[iRHO] = v[iRHO] + h;
      vmo[iRHO] = v[iRHO] + -h;
    }
    if (!g.iso && (vpo[P] < 0.0 || vmo[P] < 0.0)) {
      const double h = 0.5 * gen_minmod(dfp[P], dfm[P]);
      vpo[P] = v[P] + h;
      vmo[P] = v[P] + -h;
    }
  }
  gen_flatten_oned<NV>(g, V, dir, n, st, o, v, vpo, vmo);
}

template <int NV>
PB_D void gen_zone_states(const GenDev &g, const double *__restrict__ V, const unsigned short *__restrict__ flag, int dir,
                          int n, long st, long o, double (&v)[NV], double (&vpo)[NV], double (&vmo)[NV]) {
  const Dev &d = g.d;
  const bool uniform = g.geometry == GEO_CARTESIAN;     // UNIFORM_CARTESIAN_GRID, plm_coeffs.h:23-29
  double dvp[NV], dvm[NV], dvl[NV];
  double cp = 2.0, cm = 2.0, dp = 0.5, dm = 0.5, wp = 1.0, wm = 1.0;
  if (!uniform) {
    cp = __ldg(g.cp[dir] + n); cm = __ldg(g.cm[dir] + n); wp = __ldg(g.wp[dir] + n); wm = __ldg(g.wm[dir] + n);
    dp = __ldg(g.dp[dir] + n); dm = __ldg(g.dm[dir] + n);
  }
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    v[nv] = V[nv * d.sv + o];
    double fp = V[nv * d.sv + o + st] - v[nv], fm = v[nv] - V[nv * d.sv + o - st];
    dvp[nv] = uniform ? fp : fp * wp;
    dvm[nv] = uniform ? fm : fm * wm;
  }
  const unsigned short fl = g.flatten ? flag[o] : 0;
  if (!g.char_lim) {
    if (fl & GF_FLAT) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) dvl[nv] = 0.0;
    } else if (fl & GF_MINMOD) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) dvl[nv] = gen_lim(LIM_MINMOD, uniform, dvp[nv], dvm[nv], cp, cm);
    } else {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        int kind = g.limiter;
        if (kind == LIM_DEFAULT) kind = (nv == iRHO || nv >= gen_nflx(g)) ? LIM_MC : ((!g.iso && nv == iPRS) ? LIM_MINMOD : LIM_VANLEER);
        dvl[nv] = gen_lim(kind, uniform, dvp[nv], dvm[nv], cp, cm);
      }
    }
  } else {
    // SoundSpeed2, PrimEigenvectors (HD/eigenv.c:92-200), PrimToChar (:575-616)
    constexpr int P = pidx<NV>();
    const double a2 = g.iso ? g.cs2 : d.gas.gamma * v[P] / v[iRHO];
    const double cs = sqrt(a2), rhocs = v[iRHO] * cs, rho_cs = v[iRHO] / cs;
    const double L0p = 1.0 / rhocs, L2p = -1.0 / a2;
    // dvp/dvm of the normal, tangent, bitangent velocity: select without dynamic register indexing
    const double pn = dir == 0 ? dvp[1] : (dir == 1 ? dvp[2] : dvp[3]);
    const double pt_ = dir == 0 ? dvp[2] : (dir == 1 ? dvp[3] : dvp[1]);
    const double pb_ = dir == 0 ? dvp[3] : (dir == 1 ? dvp[1] : dvp[2]);
    const double mn = dir == 0 ? dvm[1] : (dir == 1 ? dvm[2] : dvm[3]);
    const double mt_ = dir == 0 ? dvm[2] : (dir == 1 ? dvm[3] : dvm[1]);
    const double mb_ = dir == 0 ? dvm[3] : (dir == 1 ? dvm[1] : dvm[2]);
    double dwp[NFLX], dwm[NFLX], dwl[NFLX];
    const int nf = gen_nflx(g);
    if (!g.iso) {
      dwm[0] = -1.0 * mn + L0p * dvm[P]; dwm[1] = 1.0 * mn + L0p * dvm[P];
      dwm[2] = dvm[iRHO] + L2p * dvm[P]; dwm[3] = mt_; dwm[4] = mb_;
      dwp[0] = -1.0 * pn + L0p * dvp[P]; dwp[1] = 1.0 * pn + L0p * dvp[P];
      dwp[2] = dvp[iRHO] + L2p * dvp[P]; dwp[3] = pt_; dwp[4] = pb_;
    } else {   // eigenv.c:182-196 (LL[0][RHO] = LL[1][RHO] = 1/rho_cs), PrimToChar eigenv.c:600-605
      const double Lr = 1.0 / rho_cs;
      dwm[0] = Lr * dvm[iRHO] + -1.0 * mn; dwm[1] = Lr * dvm[iRHO] + 1.0 * mn; dwm[2] = mt_; dwm[3] = mb_; dwm[4] = 0.0;
      dwp[0] = Lr * dvp[iRHO] + -1.0 * pn; dwp[1] = Lr * dvp[iRHO] + 1.0 * pn; dwp[2] = pt_; dwp[3] = pb_; dwp[4] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < NFLX; q++) {
      if (q >= nf) { dwl[q] = 0.0; continue; }
      if (fl & GF_FLAT) dwl[q] = 0.0;
      else if (fl & GF_MINMOD) dwl[q] = gen_lim(LIM_MINMOD, uniform, dwp[q], dwm[q], cp, cm);
      else if (g.limiter == LIM_DEFAULT) {
        const double kstp = q < 2 ? 1.0 : 2.0;
        const double cpk = uniform ? kstp : (2.0 - cp) + (cp - 1.0) * kstp;
        const double cmk = uniform ? kstp : (2.0 - cm) + (cm - 1.0) * kstp;
        dwl[q] = gen_lim(8, uniform, dwp[q], dwm[q], cpk, cmk);
      } else dwl[q] = gen_lim(g.limiter, uniform, dwp[q], dwm[q], cp, cm);
    }
    // dv = sum_k dw_lim[k] R[nv][k]: the reference adds all five products, zeros included
    double dc[NFLX];
    double dcn, dct, dcb;
    if (!g.iso) {
      dc[iRHO] = (((dwl[0] * (0.5 * rho_cs) + dwl[1] * (0.5 * rho_cs)) + dwl[2] * 1.0) + dwl[3] * 0.0) + dwl[4] * 0.0;
      dc[iPRS] = (((dwl[0] * (0.5 * rhocs) + dwl[1] * (0.5 * rhocs)) + dwl[2] * 0.0) + dwl[3] * 0.0) + dwl[4] * 0.0;
      dcn = dwl[0] * -0.5 + dwl[1] * 0.5; dct = dwl[3]; dcb = dwl[4];
    } else {   // R[RHO][0,1] = rho_cs/2, R[VXn][0,1] = -+1/2, R[VXt][2] = R[VXb][3] = 1 (eigenv.c:175-178); 4 products each
      dc[iRHO] = ((dwl[0] * (0.5 * rho_cs) + dwl[1] * (0.5 * rho_cs)) + dwl[2] * 0.0) + dwl[3] * 0.0;
      dc[iPRS] = 0.0;
      dcn = ((dwl[0] * -0.5 + dwl[1] * 0.5) + dwl[2] * 0.0) + dwl[3] * 0.0; dct = dwl[2]; dcb = dwl[3];
    }
    dc[1] = dir == 0 ? dcn : (dir == 1 ? dcb : dct);
    dc[2] = dir == 0 ? dct : (dir == 1 ? dcn : dcb);
    dc[3] = dir == 0 ? dcb : (dir == 1 ? dct : dcn);
#pragma unroll
    for (int nv = 0; nv < (NV < NFLX ? NV : NFLX); nv++) {
      if (nv >= nf) continue;
      if (dvp[nv] * dvm[nv] > 0.0) {
        double d2v = absmin(cp * dvp[nv], cm * dvm[nv]);
        dvl[nv] = (d2v * dc[nv] > 0.0) ? absmin(d2v, dc[nv]) : 0.0;
      } else dvl[nv] = 0.0;
    }
#pragma unroll
    for (int nv = 4; nv < NV; nv++)
      if (nv >= nf) dvl[nv] = gen_lim(g.limiter == LIM_DEFAULT ? LIM_MC : g.limiter, uniform, dvp[nv], dvm[nv], cp, cm);
  }
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    vpo[nv] = v[nv] + dvl[nv] * dp;
    vmo[nv] = v[nv] - dvl[nv] * dm;
  }
  gen_flatten_oned<NV>(g, V, dir, n, st, o, v, vpo, vmo);
}

// ---- RING_AVERAGE (Src/ring_average.c): reconstruction on the reduced grid; RingAverageCons is gen_ring below ----
// chunk size of the ring zone (i, j) belongs to; 1 = not averaged
PB_D int gen_ring_of(const GenDev &g, int i, int j) {
  if (g.ring_average <= 1) return 1;
  const int cs = __ldg(g.csize + (g.geometry == GEO_POLAR ? i : j));
  return cs > 1 ? cs : 1;
}
PB_D int gen_ring_dir(const GenDev &g) { return g.geometry == GEO_POLAR ? 1 : 2; }     // phi: x2 (POLAR), x3 (SPHERICAL)

// Monotonicity-preserving reconstruction of Suresh & Huynh as MP5_Reconstruct() has it (Src/reconstruct.c:38-93,
// MP5_ALPHA 4, Median :296-304): right-edge value of the middle one of five zone averages
PB_D double gen_mp5(double fm2, double fm1, double f0, double fp1, double fp2) {
  const double alpha = 4.0, epsm = 1.e-12;
  double f = 2.0 * fm2 - 13.0 * fm1 + 47.0 * f0 + 27.0 * fp1 - 3.0 * fp2;
  f /= 60.0;
  const double fMP = f0 + gen_minmod(fp1 - f0, alpha * (f0 - fm1));
  if ((f - f0) * (f - fMP) <= epsm) return f;
  const double d2m = fm2 + f0 - 2.0 * fm1, d2 = fm1 + fp1 - 2.0 * f0, d2p = f0 + fp2 - 2.0 * fp1;
  double s1 = gen_minmod(4.0 * d2 - d2p, 4.0 * d2p - d2), s2 = gen_minmod(d2, d2p);
  const double dMMp = gen_minmod(s1, s2);
  s1 = gen_minmod(4.0 * d2m - d2, 4.0 * d2 - d2m); s2 = gen_minmod(d2, d2m);
  const double dMMm = gen_minmod(s1, s2);
  const double fUL = f0 + alpha * (f0 - fm1), fAV = 0.5 * (f0 + fp1);
  const double fMD = fAV - 0.5 * dMMp, fLC = 0.5 * (3.0 * f0 - fm1) + 4.0 / 3.0 * dMMm;
  s1 = fmin(f0, fp1); s1 = fmin(s1, fMD);
  s2 = fmin(f0, fUL); s2 = fmin(s2, fLC);
  const double lo = fmax(s1, s2);
  s1 = fmax(f0, fp1); s1 = fmax(s1, fMD);
  s2 = fmax(f0, fUL); s2 = fmax(s2, fLC);
  const double hi = fmin(s1, s2);
  return f + gen_minmod(lo - f, hi - f);
}

// RingAverageReconstruct() (ring_average.c:172-400) for zone n of a phi line whose ring has chunk size cs > 1: the chunk
// averages (periodic in phi) are reconstructed on the reduced grid (RING_AVERAGE_REC 2: van Leer, 5: MP5) and the parabola
// through (vam, va, vap) is evaluated at the edges of the zone inside its chunk.  Ghost zones take the periodic image.
template <int NV>
PB_D void gen_ring_states(const GenDev &g, const double *__restrict__ V, int dir, int n, long st, long o, int cs,
                          double (&v)[NV], double (&vpo)[NV], double (&vmo)[NV]) {
  const Dev &d = g.d;
  const int dbeg = d.beg[dir], dend = d.end[dir], nphi = dend - dbeg + 1, nch = nphi / cs;
  int jw = n;
  if (jw < dbeg) jw += nphi; else if (jw > dend) jw -= nphi;
  const int ja = (jw - dbeg) / cs, kk = (jw - dbeg) % cs + 1;
  const long base = o - (long)n * st;
  const double xp = kk / (double)cs, xm = (kk - 1.0) / (double)cs;
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    const double *q = V + nv * d.sv + base;
    v[nv] = q[(long)n * st];
    auto VA = [&](int m) {
      int c = (ja + m) % nch;
      if (c < 0) c += nch;
      return q[(long)(dbeg + c * cs) * st];
    };
    const double va = VA(0);
    double vap, vam;
    if (g.ring_rec == 2) {
      const double dvap = VA(1) - va, dvam = va - VA(-1);
      const double dva = dvap * dvam > 0.0 ? 2.0 * dvap * dvam / (dvap + dvam) : 0.0;     // VANLEER_LIMITER
      vap = va + 0.5 * dva;
      vam = va - 0.5 * dva;
    } else {
      const double m2 = VA(-2), m1 = VA(-1), p1 = VA(1), p2 = VA(2);
      vap = gen_mp5(m2, m1, va, p1, p2);
      vam = gen_mp5(p2, p1, va, m1, m2);
    }
    const double A = 3.0 * ((vap + vam) - 2.0 * va);
    const double B = -4.0 * vam - 2.0 * vap + 6.0 * va;
    vpo[nv] = A * xp * xp + B * xp + vam;
    vmo[nv] = A * xm * xm + B * xm + vam;
  }
}

template <int NV>
static __global__ void gen_states(GenDev g, GenArgs a, GenBox b) {
  int i, j, k;
  if (!gen_zone(b.lo, b.hi, i, j, k)) return;
  const Dev &d = g.d;
  const int dir = a.dir;
  const long st = dir == 0 ? 1 : (dir == 1 ? d.sj : d.sk);
  const long o = (long)k * d.sk + (long)j * d.sj + i;
  const int n = dir == 0 ? i : (dir == 1 ? j : k);
  double v[NV], vp[NV], vm[NV];
  const int ring = (g.ring_average > 1 && g.ring_rec != 1 && dir == gen_ring_dir(g)) ? gen_ring_of(g, i, j) : 1;
  if (ring > 1) gen_ring_states<NV>(g, a.V, dir, n, st, o, ring, v, vp, vm);      // update_stage.c:230-234
  else if (g.ppm) gen_zone_states_ppm<NV>(g, a.V, a.flag, dir, n, st, o, v, vp, vm);
  else gen_zone_states<NV>(g, a.V, a.flag, dir, n, st, o, v, vp, vm);
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    a.VP[nv * d.sv + o] = vp[nv];
    a.VM[nv * d.sv + o] = vm[nv];
  }
}

// ---- EOS ISOTHERMAL Riemann solvers: HD/tvdlf.c:100-130, HD/hll.c:72-96, HD/hllc.c:119-178 with the
// "#if EOS == ISOTHERMAL" star state (hllc.c:137-150), fluxes.c:36-48 (p = cs2 rho), hll_speed.c:76-90.
// q: (rho, v_n, v_t, v_b, scalars...) in sweep-local order; f[0..3] mass and momentum fluxes (no pressure in f[1]).
// Written with the reference's own operations (plain divisions): the isothermal problems are small grids.
template <int NV>
PB_D double riemann_iso(const double (&vL)[NV], const double (&vR)[NV], double cs2, int solver, bool force_hll,
                        double (&f)[NV], double &prs, double &cmax) {
  double uL[4], uR[4], fL[4], fR[4];
  uL[0] = vL[0]; uL[1] = vL[0] * vL[1]; uL[2] = vL[0] * vL[2]; uL[3] = vL[0] * vL[3];
  uR[0] = vR[0]; uR[1] = vR[0] * vR[1]; uR[2] = vR[0] * vR[2]; uR[3] = vR[0] * vR[3];
  fL[0] = uL[1]; fL[1] = uL[1] * vL[1]; fL[2] = uL[2] * vL[1]; fL[3] = uL[3] * vL[1];
  fR[0] = uR[1]; fR[1] = uR[1] * vR[1]; fR[2] = uR[2] * vR[1]; fR[3] = uR[3] * vR[1];
  const double pL = cs2 * vL[0], pR = cs2 * vR[0];
  double machv;
  if (solver == SOLVER_TVDLF) {
    const double vn = 0.5 * (fabs(vL[1]) + fabs(vR[1]));
    const double a = sqrt(cs2);
    const double cmin = vn - a, cmaxv = vn + a;
    cmax = fmax(fabs(cmaxv), fabs(cmin));
    machv = fabs(vn) / sqrt(cs2);
#pragma unroll
    for (int nv = 0; nv < 4; nv++) f[nv] = 0.5 * (fL[nv] + fR[nv] - cmax * (uR[nv] - uL[nv]));
    prs = 0.5 * (pL + pR);
  } else {
    const double aL = sqrt(cs2), aR = sqrt(cs2);
    const double SL = fmin(vL[1] - aL, vR[1] - aR), SR = fmax(vL[1] + aL, vR[1] + aR);
    double scrh = fabs(vL[1]) + fabs(vR[1]);
    scrh /= aL + aR;
    machv = scrh;
    cmax = fmax(fabs(SL), fabs(SR));
    if (SL > 0.0) {
#pragma unroll
      for (int nv = 0; nv < 4; nv++) f[nv] = fL[nv];
      prs = pL;
    } else if (SR < 0.0) {
#pragma unroll
      for (int nv = 0; nv < 4; nv++) f[nv] = fR[nv];
      prs = pR;
    } else if (solver == SOLVER_HLL || force_hll) {
      scrh = 1.0 / (SR - SL);
#pragma unroll
      for (int nv = 0; nv < 4; nv++) {
        f[nv] = SL * SR * (uR[nv] - uL[nv]) + SR * fL[nv] - SL * fR[nv];
        f[nv] *= scrh;
      }
      prs = (SR * pL - SL * pR) * scrh;
    } else {
      scrh = 1.0 / (SR - SL);
      const double rho = (SR * uR[0] - SL * uL[0] - fR[0] + fL[0]) * scrh;
      const double mx = (SR * uR[1] - SL * uL[1] - fR[1] + fL[1]) * scrh;
      double vs = (SR * fL[0] - SL * fR[0] + SR * SL * (uR[0] - uL[0]));
      vs *= scrh;
      vs /= rho;
      if (vs >= 0.0) {
        const double us[4] = {rho, mx, rho * vL[2], rho * vL[3]};
#pragma unroll
        for (int nv = 0; nv < 4; nv++) f[nv] = fL[nv] + SL * (us[nv] - uL[nv]);
        prs = pL;
      } else {
        const double us[4] = {rho, mx, rho * vR[2], rho * vR[3]};
#pragma unroll
        for (int nv = 0; nv < 4; nv++) f[nv] = fR[nv] + SR * (us[nv] - uR[nv]);
        prs = pR;
      }
    }
  }
#pragma unroll
  for (int nv = 4; nv < NV; nv++) f[nv] = f[0] * (f[0] > 0.0 ? vL[nv] : vR[nv]);    // adv_flux.c:61-72
  return machv;
}

// ---- Roe_Solver (HD/roe.c:48-346, ROE_AVERAGE YES) and TwoShock_Solver (HD/two_shock.c:28-243), both equations of
// state for Roe, EOS IDEAL for two-shock.  q: (rho, v_n, v_t, v_b[, p], scalars...) in sweep-local order with NF = 4
// (isothermal) or 5 flux components; f[0..NF-1] without the pressure in f[1].  Small-grid options: written with the
// reference's own operations.
enum { SOLVER_ROE = 4, SOLVER_TWO_SHOCK = 5, SOLVER_AUSM = 6 };
template <int NV>
PB_D double riemann_roe_ts(const double (&vL)[NV], const double (&vR)[NV], bool iso, double cs2, double gamma, int solver,
                           bool force_hll, int ndim, double (&f)[NV], double &prs, double &cmax) {
  constexpr int P = NV > 4 ? 4 : 0;
  const int nf = iso ? 4 : 5;
  double uL[5], uR[5], fL[5], fR[5];
  uL[0] = vL[0]; uL[1] = vL[0] * vL[1]; uL[2] = vL[0] * vL[2]; uL[3] = vL[0] * vL[3];
  uR[0] = vR[0]; uR[1] = vR[0] * vR[1]; uR[2] = vR[0] * vR[2]; uR[3] = vR[0] * vR[3];
  fL[0] = uL[1]; fL[1] = uL[1] * vL[1]; fL[2] = uL[2] * vL[1]; fL[3] = uL[3] * vL[1];
  fR[0] = uR[1]; fR[1] = uR[1] * vR[1]; fR[2] = uR[2] * vR[1]; fR[3] = uR[3] * vR[1];
  const double gmm1 = gamma - 1.0;
  double a2L, a2R, pL, pR;
  uL[4] = uR[4] = fL[4] = fR[4] = 0.0;
  if (!iso) {
    uL[4] = vL[1] * vL[1] + vL[2] * vL[2] + vL[3] * vL[3];
    uL[4] = 0.5 * vL[0] * uL[4] + vL[P] / gmm1;
    uR[4] = vR[1] * vR[1] + vR[2] * vR[2] + vR[3] * vR[3];
    uR[4] = 0.5 * vR[0] * uR[4] + vR[P] / gmm1;
    a2L = gamma * vL[P] / vL[0]; a2R = gamma * vR[P] / vR[0];
    fL[4] = (uL[4] + vL[P]) * vL[1]; fR[4] = (uR[4] + vR[P]) * vR[1];
    pL = vL[P]; pR = vR[P];
  } else { a2L = a2R = cs2; pL = a2L * vL[0]; pR = a2R * vR[0]; }
  double fl[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, machv = 0.0;
  bool done = false;
  if (solver == SOLVER_AUSM) {      // AUSMp_Solver, HD/ausm.c:20-110 (EOS IDEAL; no HLL switch in flagged zones)
    const double alpha = 3.0 / 16.0, beta = 0.125;
    const double aL = sqrt(gamma * vL[P] / vL[0]), aR = sqrt(gamma * vR[P] / vR[0]);
    double asL2 = vL[1] * vL[1] + vL[2] * vL[2] + vL[3] * vL[3];
    asL2 = aL * aL / gmm1 + 0.5 * asL2;
    asL2 *= 2.0 * gmm1 / (gamma + 1.0);
    double asR2 = vR[1] * vR[1] + vR[2] * vR[2] + vR[3] * vR[3];
    asR2 = aR * aR / gmm1 + 0.5 * asR2;
    asR2 *= 2.0 * gmm1 / (gamma + 1.0);
    const double asL = sqrt(asL2), asR = sqrt(asR2);
    const double atL = asL2 / fmax(asL, fabs(vL[1])), atR = asR2 / fmax(asR, fabs(vR[1]));
    const double a = fmin(atL, atR);
    const double ML = vL[1] / a, MR = vR[1] / a;
    double MpL, PpL, MmR, PmR;
    if (fabs(ML) >= 1.0) { MpL = 0.5 * (ML + fabs(ML)); PpL = ML > 0.0 ? 1.0 : 0.0; }
    else {
      MpL = 0.25 * (ML + 1.0) * (ML + 1.0) + beta * (ML * ML - 1.0) * (ML * ML - 1.0);
      PpL = 0.25 * (ML + 1.0) * (ML + 1.0) * (2.0 - ML) + alpha * ML * (ML * ML - 1.0) * (ML * ML - 1.0);
    }
    if (fabs(MR) >= 1.0) { MmR = 0.5 * (MR - fabs(MR)); PmR = MR > 0.0 ? 0.0 : 1.0; }
    else {
      MmR = -0.25 * (MR - 1.0) * (MR - 1.0) - beta * (MR * MR - 1.0) * (MR * MR - 1.0);
      PmR = 0.25 * (MR - 1.0) * (MR - 1.0) * (2.0 + MR) - alpha * MR * (MR * MR - 1.0) * (MR * MR - 1.0);
    }
    const double m = MpL + MmR;
    const double mp = 0.5 * (m + fabs(m)), mm = 0.5 * (m - fabs(m));
    prs = PpL * vL[P] + PmR * vR[P];
#pragma unroll
    for (int nv = 0; nv < 4; nv++) fl[nv] = a * (mp * uL[nv] + mm * uR[nv]);
    fl[4] = a * (mp * (uL[4] + vL[P]) + mm * (uR[4] + vR[P]));
    cmax = fmax(fabs(vL[1]) + aL, fabs(vR[1]) + aR);
    machv = fmax(fabs(ML), fabs(MR));
    done = true;
  } else if (force_hll) {       // roe.c:101-116, two_shock.c:66-88: HLL in zones flagged by MULTID flattening
    const double aL = sqrt(a2L), aR = sqrt(a2R);
    double bmin = fmin(vL[1] - aL, vR[1] - aR), bmax = fmax(vL[1] + aL, vR[1] + aR);
    double scrh = fabs(vL[1]) + fabs(vR[1]);
    scrh /= aL + aR;
    machv = scrh;
    cmax = fmax(fabs(bmin), fabs(bmax));
    bmin = fmin(0.0, bmin);
    bmax = fmax(0.0, bmax);
    scrh = 1.0 / (bmax - bmin);
    for (int nv = nf; nv--;) { fl[nv] = bmin * bmax * (uR[nv] - uL[nv]) + bmax * fL[nv] - bmin * fR[nv]; fl[nv] *= scrh; }
    prs = (bmax * pL - bmin * pR) * scrh;
    done = true;
  }
  if (!done && solver == SOLVER_ROE) {
    const double delta = 1.e-7, gmm1_inv = 1.0 / gmm1;
    double Rc[5][5], lambda[5], alambda[5], eta[5], dv[5], um[4];
#pragma unroll
    for (int q = 0; q < 5; q++) {
#pragma unroll
      for (int r = 0; r < 5; r++) Rc[q][r] = 0.0;
      lambda[q] = eta[q] = dv[q] = 0.0;
    }
    for (int nv = nf; nv--;) dv[nv] = (nv == 4 ? vR[P] - vL[P] : vR[nv] - vL[nv]);
    double sq = sqrt(vR[0] / vL[0]);
    um[0] = vL[0] * sq;
    sq = 1.0 / (1.0 + sq);
    const double cq = 1.0 - sq;
    um[1] = sq * vL[1] + cq * vR[1];
    um[2] = sq * vL[2] + cq * vR[2];
    um[3] = sq * vL[3] + cq * vR[3];
    double a2, a, h = 0.0, vel2 = 0.0;
    if (!iso) {
      vel2 = um[1] * um[1] + um[2] * um[2] + um[3] * um[3];     // (global order VX1..VX3: a permutation of the same sum
      double hl = 0.5 * (vL[1] * vL[1] + vL[2] * vL[2] + vL[3] * vL[3]);   //  of three products; <= 1 ulp)
      hl += a2L * gmm1_inv;
      double hr = 0.5 * (vR[1] * vR[1] + vR[2] * vR[2] + vR[3] * vR[3]);
      hr += a2R * gmm1_inv;
      h = sq * hl + cq * hr;
      a2 = gmm1 * (h - 0.5 * vel2);
      a = sqrt(a2);
    } else { a2 = 0.5 * (a2L + a2R); a = sqrt(a2); }
    int nn = 0;                         // u - c_s
    lambda[nn] = um[1] - a;
    if (!iso) eta[nn] = 0.5 / a2 * (dv[4] - dv[1] * um[0] * a);
    else eta[nn] = 0.5 * (dv[0] - um[0] * dv[1] / a);
    Rc[0][nn] = 1.0; Rc[1][nn] = um[1] - a; Rc[2][nn] = um[2]; Rc[3][nn] = um[3];
    if (!iso) Rc[4][nn] = h - um[1] * a;
    nn = 1;                             // u + c_s
    lambda[nn] = um[1] + a;
    if (!iso) eta[nn] = 0.5 / a2 * (dv[4] + dv[1] * um[0] * a);
    else eta[nn] = 0.5 * (dv[0] + um[0] * dv[1] / a);
    Rc[0][nn] = 1.0; Rc[1][nn] = um[1] + a; Rc[2][nn] = um[2]; Rc[3][nn] = um[3];
    if (!iso) Rc[4][nn] = h + um[1] * a;
    if (!iso) {                         // u (entropy wave)
      nn = 2;
      lambda[nn] = um[1];
      eta[nn] = dv[0] - dv[4] / a2;
      Rc[0][nn] = 1.0; Rc[1][nn] = um[1]; Rc[2][nn] = um[2]; Rc[3][nn] = um[3];
      Rc[4][nn] = 0.5 * vel2;
    }
    nn++;                               // u (shear waves)
    lambda[nn] = um[1];
    eta[nn] = um[0] * dv[2];
    Rc[2][nn] = 1.0;
    if (!iso) Rc[4][nn] = um[2];
    nn++;
    lambda[nn] = um[1];
    eta[nn] = um[0] * dv[3];
    Rc[3][nn] = 1.0;
    if (!iso) Rc[4][nn] = um[3];
    cmax = fabs(um[1]) + a;
    machv = fabs(um[1] / a);
    if (ndim > 1) {                     // roe.c:262-287: HLL inside strong shocks
      double scrh;
      if (!iso) { scrh = fabs(vL[P] - vR[P]); scrh /= fmin(vL[P], vR[P]); }
      else { scrh = fabs(vL[0] - vR[0]); scrh /= fmin(vL[0], vR[0]); scrh *= a * a; }
      if (scrh > 0.5 && (vR[1] < vL[1])) {
        const double bmin = fmin(0.0, lambda[0]), bmax = fmax(0.0, lambda[1]);
        const double scrh1 = 1.0 / (bmax - bmin);
        for (int nv = nf; nv--;) { fl[nv] = bmin * bmax * (uR[nv] - uL[nv]) + bmax * fL[nv] - bmin * fR[nv]; fl[nv] *= scrh1; }
        prs = (bmax * pL - bmin * pR) * scrh1;
        done = true;
      }
    }
    if (!done) {
      for (int nv = nf; nv--;) alambda[nv] = fabs(lambda[nv]);
      if (alambda[0] <= delta) alambda[0] = 0.5 * lambda[0] * lambda[0] / delta + 0.5 * delta;   // entropy fix
      if (alambda[1] <= delta) alambda[1] = 0.5 * lambda[1] * lambda[1] / delta + 0.5 * delta;
      for (int nv = nf; nv--;) {
        fl[nv] = fL[nv] + fR[nv];
        for (int kk = nf; kk--;) fl[nv] -= alambda[kk] * eta[kk] * Rc[nv][kk];
        fl[nv] *= 0.5;
      }
      prs = 0.5 * (pL + pR);
    }
  } else if (!done) {
    // TwoShock_Solver, two_shock.c:90-243 (EOS IDEAL; MAX_ITER 5, small_p = small_rho = 1e-9)
    const double small_p = 1.e-9, small_rho = 1.e-9;
    const double g1_g = 0.5 * (gamma + 1.0) / gamma;
    const double cl = sqrt(gamma * vL[P] * vL[0]), cr = sqrt(gamma * vR[P] * vR[0]);
    const double taul = 1.0 / vL[0], taur = 1.0 / vR[0];
    double vxl = 0.0, vxr = 0.0, scrh1, scrh2, scrh3, scrh4, dp;
    double pstar = vR[P] - vL[P] - cr * (vR[1] - vL[1]);
    pstar = vL[P] + pstar * cl / (cl + cr);
    pstar = fmax(small_p, pstar);
    for (int iter = 1; iter <= 5; iter++) {
      vxl = cl * sqrt(1.0 + g1_g * (pstar - vL[P]) / vL[P]);
      vxr = cr * sqrt(1.0 + g1_g * (pstar - vR[P]) / vR[P]);
      scrh1 = vxl * vxl;
      scrh1 = 2.0 * scrh1 * vxl / (scrh1 + cl * cl);
      scrh2 = vxr * vxr;
      scrh2 = 2.0 * scrh2 * vxr / (scrh2 + cr * cr);
      scrh3 = vL[1] - (pstar - vL[P]) / vxl;
      scrh4 = vR[1] + (pstar - vR[P]) / vxr;
      dp = scrh1 * scrh2 / (scrh1 + scrh2) * (scrh4 - scrh3);
      pstar -= dp;
      pstar = fmax(small_p, pstar);
      if (fabs(dp / pstar) < 1.e-6) break;
    }
    scrh3 = vL[1] - (pstar - vL[P]) / vxl;
    scrh4 = vR[1] + (pstar - vR[P]) / vxr;
    const double ustar = 0.5 * (scrh3 + scrh4);
    const bool left = ustar > 0.0;
    const double sigma = left ? 1.0 : -1.0, taus = left ? taul : taur, cs = left ? cl * taul : cr * taur, zs = left ? vxl : vxr;
    const double qs0 = left ? vL[0] : vR[0], qs1 = left ? vL[1] : vR[1], qsp = left ? vL[P] : vR[P];
    const double qst = left ? vL[2] : vR[2], qsb = left ? vL[3] : vR[3];
    double rho_star = taus - (pstar - qsp) / (zs * zs);
    rho_star = fmax(small_rho, 1.0 / rho_star);
    const double cstar0 = sqrt(gamma * pstar / rho_star);
    double lambda_s, lambda_star;
    if (pstar < qsp) { lambda_s = cs - sigma * qs1; lambda_star = cstar0 - sigma * ustar; }
    else lambda_s = lambda_star = zs * taus - sigma * qs1;
    double vS0, vS1, vSp;
    if (lambda_star > 0.0) { vS0 = rho_star; vS1 = ustar; vSp = pstar; }
    else if (lambda_s < 0.0) { vS0 = qs0; vS1 = qs1; vSp = qsp; }
    else {
      scrh1 = fmax(lambda_s - lambda_star, lambda_s + lambda_star);
      scrh1 = fmax(1.e-12, scrh1);
      const double zeta = 0.5 * (1.0 + (lambda_s + lambda_star) / scrh1);
      vS0 = zeta * rho_star + (1.0 - zeta) * qs0;
      vS1 = zeta * ustar + (1.0 - zeta) * qs1;
      vSp = zeta * pstar + (1.0 - zeta) * qsp;
    }
    const double uS1 = vS0 * vS1, uS2 = vS0 * qst, uS3 = vS0 * qsb;
    double uS4 = vS1 * vS1 + qst * qst + qsb * qsb;
    uS4 = 0.5 * vS0 * uS4 + vSp / gmm1;
    const double a2S = gamma * vSp / vS0;
    fl[0] = uS1; fl[1] = uS1 * vS1; fl[2] = uS2 * vS1; fl[3] = uS3 * vS1; fl[4] = (uS4 + vSp) * vS1;
    prs = vSp;
    const double cstar = sqrt(a2S);
    machv = fabs(vS1) / cstar;
    cmax = fabs(vS1) + cstar;
  }
#pragma unroll
  for (int nv = 0; nv < 4; nv++) f[nv] = fl[nv];
  if (!iso) f[P] = fl[4];
#pragma unroll
  for (int nv = 4; nv < NV; nv++) if (nv >= nf) f[nv] = f[0] * (f[0] > 0.0 ? vL[nv] : vR[nv]);    // adv_flux.c:61-72
  return machv;
}

// ---- Riemann solver + AdvectFlux at the face between zone n and n+1 -------------------------
// vLg / vRg: left / right state in GLOBAL variable order; F: flux in global order, F[NV] = pressure, F[NV+1] = cmax
// X: the Roe / two-shock / AUSM+ code is compiled in (the fused sweeps are instantiated with and without it, so that the
// HLL-family kernels keep their register budget)
template <int NV, bool X = true>
PB_D double gen_face(const GenDev &g, int dir, const double (&vLg)[NV], const double (&vRg)[NV], bool hll, double (&F)[NV + 2]) {
  const int gn = 1 + dir, gt = 1 + (dir + 1) % 3, gb = 1 + (dir + 2) % 3;
  double vL[NV], vR[NV];     // sweep-local order (n, t, b)
  vL[0] = vLg[0]; vR[0] = vRg[0];
  vL[1] = dir == 0 ? vLg[1] : (dir == 1 ? vLg[2] : vLg[3]); vR[1] = dir == 0 ? vRg[1] : (dir == 1 ? vRg[2] : vRg[3]);
  vL[2] = dir == 0 ? vLg[2] : (dir == 1 ? vLg[3] : vLg[1]); vR[2] = dir == 0 ? vRg[2] : (dir == 1 ? vRg[3] : vRg[1]);
  vL[3] = dir == 0 ? vLg[3] : (dir == 1 ? vLg[1] : vLg[2]); vR[3] = dir == 0 ? vRg[3] : (dir == 1 ? vRg[1] : vRg[2]);
#pragma unroll
  for (int nv = 4; nv < NV; nv++) { vL[nv] = vLg[nv]; vR[nv] = vRg[nv]; }
  if (g.iso || (X && g.solver >= SOLVER_ROE)) {
    double fl[NV], prs, cmx;
    double mv;
    if constexpr (X) {
      mv = g.solver >= SOLVER_ROE ? riemann_roe_ts<NV>(vL, vR, g.iso != 0, g.cs2, g.d.gas.gamma, g.solver, hll, g.d.ndim, fl, prs, cmx)
                                  : riemann_iso<NV>(vL, vR, g.cs2, g.solver, hll, fl, prs, cmx);
    } else mv = riemann_iso<NV>(vL, vR, g.cs2, g.solver, hll, fl, prs, cmx);
    if (g.entropy) fl[NV - 1] = fl[0] * sel(fl[0] >= 0.0, vL[NV - 1], vR[NV - 1]);    // adv_flux.c:131-134
    F[0] = fl[0];
    F[1] = dir == 0 ? fl[1] : (dir == 1 ? fl[3] : fl[2]);
    F[2] = dir == 0 ? fl[2] : (dir == 1 ? fl[1] : fl[3]);
    F[3] = dir == 0 ? fl[3] : (dir == 1 ? fl[2] : fl[1]);
#pragma unroll
    for (int nv = 4; nv < NV; nv++) F[nv] = fl[nv];
    F[NV] = prs;
    F[NV + 1] = cmx;
    return mv;
  }
  Face<NV> Ff;
  Ratio mach;
  mach.init();
  if (g.solver == SOLVER_TVDLF) riemann<NV, SOLVER_TVDLF>(vL, vR, g.d.gas, Ff, mach);
  else if (g.solver == SOLVER_HLL) riemann<NV, SOLVER_HLL>(vL, vR, g.d.gas, Ff, mach);
  else riemann<NV, SOLVER_HLLC>(vL, vR, g.d.gas, Ff, mach, true, hll);
  if (g.entropy) {    // adv_flux.c:131-134: ">=" for the entropy, ">" for the other scalars
    Ff.f[NV - 1] = Ff.f[iRHO] * sel(Ff.f[iRHO] >= 0.0, vL[NV - 1], vR[NV - 1]);
  }
  F[0] = Ff.f[0];
  // local -> global: F[gn] = f[1], F[gt] = f[2], F[gb] = f[3]
  F[1] = dir == 0 ? Ff.f[1] : (dir == 1 ? Ff.f[3] : Ff.f[2]);
  F[2] = dir == 0 ? Ff.f[2] : (dir == 1 ? Ff.f[1] : Ff.f[3]);
  F[3] = dir == 0 ? Ff.f[3] : (dir == 1 ? Ff.f[2] : Ff.f[1]);
#pragma unroll
  for (int nv = 4; nv < NV; nv++) F[nv] = Ff.f[nv];
  F[NV] = Ff.prs;
  F[NV + 1] = Ff.cmax;
  (void)gn; (void)gt; (void)gb;
  return mach.value();
}

template <int NV>
static __global__ void gen_riemann(GenDev g, GenArgs a, GenBox b) {
  int i, j, k;
  double machv = 0.0;
  const Dev &d = g.d;
  if (gen_zone(b.lo, b.hi, i, j, k)) {
    const int dir = a.dir;
    const long st = dir == 0 ? 1 : (dir == 1 ? d.sj : d.sk);
    const long o = (long)k * d.sk + (long)j * d.sj + i;
    double vL[NV], vR[NV], F[NV + 2];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) { vL[nv] = a.VP[nv * d.sv + o]; vR[nv] = a.VM[nv * d.sv + o + st]; }
    const bool hll = g.flatten && ((a.flag[o] & GF_HLL) || (a.flag[o + st] & GF_HLL));
    machv = gen_face<NV>(g, dir, vL, vR, hll, F);
    if ((d.bf_kind & 2) && !g.iso) F[pidx<NV>()] += F[iRHO] * bf_at(d, 4 + dir, i, j, k);   // TotalFlux(), rhs.c:171-179,525
    const long nz = d.sv;
#pragma unroll
    for (int nv = 0; nv < NV + 2; nv++) a.F[nv * nz + o] = F[nv];
  }
  machv = warp_max(machv);
  if ((threadIdx.x & 31) == 0 && machv > 0.0) atomic_max_pos(a.red + 1, machv);
}

// ---- line-driven wind ----------------------------------------------------------------------
// bilinear(), line_connect.c:746-767
PB_D void gen_bilinear(const double (&x11)[2], const double (&x22)[2], const double (&v11)[2], const double (&v12)[2],
                       const double (&v21)[2], const double (&v22)[2], double t0, double t1, double (&ans)[2]) {
  const double f1 = (t0 - x11[0]) / (x22[0] - x11[0]);
  const double f2 = (t1 - x11[1]) / (x22[1] - x11[1]);
  double a = (1.0 - f1) * v11[0] + f1 * v21[0];
  double b = (1.0 - f1) * v12[0] + f1 * v22[0];
  ans[0] = (1.0 - f2) * a + f2 * b;
  a = (1.0 - f1) * v11[1] + f1 * v21[1];
  b = (1.0 - f1) * v12[1] + f1 * v22[1];
  ans[1] = (1.0 - f2) * a + f2 * b;
}

// The only use VGradCalc() makes of the flux tables is the test |F| != 0 per zone and bin
// (line_connect.c:560-575): the tables change when new SIROCCO fluxes arrive, not per step, so
// the test is taken once per hand-over and kept as one bit per bin - 8 B per zone in place of
// 3 x 36 x 8 B read by every gen_vgrad launch.
static __global__ void gen_ldw_mask(LdwDev w, long nz, unsigned long long *mask) {
  const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nz) return;
  unsigned long long m = 0ull;
  for (int ia = 0; ia < w.nangles; ia++) {
    const double fr = w.flux_r[ia * nz + o], ft = w.flux_t[ia * nz + o], fp = w.flux_p ? w.flux_p[ia * nz + o] : 0.0;
    if (sqrt(fr * fr + ft * ft + fp * fp) != 0.0) m |= 1ull << ia;
  }
  mask[o] = m;
}

// LineForce(), line_connect.c:815-903, split into its per-zone and per-bin parts.
// Per zone and sweep: S = sigma_e rho v_th of the sweep's centre state (the r sweep passes
// (vp + vm)/2, the theta sweep the zone value: rhs_source.c:229-232) and, for the KRAD / ALPHARAD
// power law, kS = k S^alpha.
struct LdwZone { double S, kS; };
PB_D LdwZone gen_ldw_zone(const LdwDev &w, double rho_code, double prs_code) {
  const double rho = rho_code * w.UD;
  const double T = w.t_iso > 0.0 ? w.t_iso : prs_code / rho_code * w.kelvin_mu;    // EOS ISOTHERMAL: T_ISO (line_connect.c:851-855)
  const double v_th = sqrt((2.0 * 1.3806505e-16 * T) / 1.67262171e-24);
  LdwZone z;
  z.S = w.sigma_e * rho * v_th;
  z.kS = w.mpoints > 0 ? 0.0 : w.krad * pow(z.S, w.alpharad);
  return z;
}
// Per bin: the force multiplier M(t), capped at M_max = 4400.  D = dvds^(-alpha) for the power law
// (M = k (S / dvds)^alpha = kS D), D = dvds itself for the per-zone fit; 0 where dvds <= 0.
PB_D double gen_ldw_M(const LdwDev &w, const LdwZone &z, double D, long o, long nz) {
  if (w.mpoints <= 0) return fmin(z.kS * D, 4400.0);
  if (!(D > 0.0)) return 0.0;
  // linterp(log10 t, t_fit, M_UV_fit[.][zone]), line_connect.c:781-811
  const double x = log10(z.S / D);
  int idx = 0;
  while (idx < w.mpoints && __ldg(w.t_fit + idx) < x) idx++;
  double y;
  if (idx == 0) y = __ldg(w.m_fit + o);
  else if (idx >= w.mpoints) y = __ldg(w.m_fit + (long)(w.mpoints - 1) * nz + o);
  else {
    const double xl = __ldg(w.t_fit + idx - 1), xh = __ldg(w.t_fit + idx);
    const double yl = __ldg(w.m_fit + (long)(idx - 1) * nz + o), yh = __ldg(w.m_fit + (long)idx * nz + o);
    y = yl + (yh - yl) / (xh - xl) * (x - xl);
  }
  double M = pow(10.0, y);
  if (!(M == M)) M = 0.0;
  return fmin(M, 4400.0);
}

// VGradCalc(), line_connect.c:504-744: one thread per zone, looping over the angular bins so that
// everything that does not depend on the bin (vertex velocities, interpolation box, the velocity at
// the cell centre) is computed once; the offsets that the reference tabulates once
// (dvds_r/t/mod_offset) are recomputed, they only depend on the grid.
// The kernel also takes the sums of LineForce() while the gradient of a bin is still in a register:
// the reference stores dvds_array[36][zones] and every sweep walks it again together with the flux
// tables (3 x 155 MB per sweep on the 1024 x 512 grid, more than L2 holds); here the 36-bin tables
// are read ONCE per stage and the sweeps pick up one number per zone.  It runs after States() of
// the r sweep because that sweep's force uses (vp + vm)/2 as its centre state.
template <int NV, int NBINS = PB_VGRAD_NB, int MINB = PB_VGRAD_MINB>
static __global__ void __launch_bounds__(64, MINB) gen_vgrad(GenDev g, GenArgs a, GenBox b, int fused) {
  int i, j, k;
  if (!gen_zone(b.lo, b.hi, i, j, k)) return;
  const Dev &d = g.d;
  const LdwDev &w = g.ldw;
  const long o = (long)k * d.sk + (long)j * d.sj + i;
  const long sj = d.sj;
  const double *x1 = g.x[0], *x2 = g.x[1];
  const double *V1 = a.V + 1 * d.sv, *V2 = a.V + 2 * d.sv;
  double x11[2], x22[2], v11[2], v12[2], v22[2], v21[2], ans1[2], ans2[2];
  const double x1i = __ldg(x1 + i), x2j = __ldg(x2 + j);
  x11[0] = (__ldg(x1 + i - 1) + x1i) / 2.0 * w.UL;
  x11[1] = (__ldg(x2 + j - 1) + x2j) / 2.0;
  x22[0] = (__ldg(x1 + i + 1) + x1i) / 2.0 * w.UL;
  x22[1] = (__ldg(x2 + j + 1) + x2j) / 2.0;
  double maxds = fabs(x22[0] - x11[0]);
  const double arc = x1i * w.UL * fabs(x22[1] - x11[1]);
  if (maxds > arc) maxds = arc;
  maxds /= 2.0;
  const double st = __ldg(w.sin_t + j), ct = __ldg(w.cos_t + j);
  v11[0] = (V1[o - sj - 1] + V1[o - sj] + V1[o - 1] + V1[o]) / 4.0;
  v11[1] = (V2[o - sj - 1] + V2[o - sj] + V2[o - 1] + V2[o]) / 4.0;
  v12[0] = (V1[o + sj - 1] + V1[o - 1] + V1[o + sj] + V1[o]) / 4.0;
  v12[1] = (V2[o + sj - 1] + V2[o - 1] + V2[o + sj] + V2[o]) / 4.0;
  v22[0] = (V1[o + sj] + V1[o + sj + 1] + V1[o + 1] + V1[o]) / 4.0;
  v22[1] = (V2[o + sj] + V2[o + sj + 1] + V2[o + 1] + V2[o]) / 4.0;
  v21[0] = (V1[o + 1] + V1[o - sj + 1] + V1[o - sj] + V1[o]) / 4.0;
  v21[1] = (V2[o + 1] + V2[o - sj + 1] + V2[o - sj] + V2[o]) / 4.0;
  gen_bilinear(x11, x22, v11, v12, v21, v22, x1i * w.UL, x2j, ans1);
  const double vx1 = (ans1[0] * w.UV * st + ans1[1] * w.UV * ct);
  const double vz1 = (ans1[0] * w.UV * ct - ans1[1] * w.UV * st);
  const double x = x1i * st * w.UL, z = x1i * ct * w.UL;
  const unsigned long long mk = __ldg(w.mask + o);
  const long nz = d.sv;
  // centre states of the two sweeps (rhs_source.c:229-232)
  double rho_c, prs_c;
  if (fused) {     // left by the fused r sweep (gen_sweep, dir 0)
    rho_c = a.cen[o];
    prs_c = a.cen[nz + o];
  } else {
    rho_c = 0.5 * (a.VP[iRHO * nz + o] + a.VM[iRHO * nz + o]);
    prs_c = g.iso ? 0.0 : 0.5 * (a.VP[iPRS * nz + o] + a.VM[iPRS * nz + o]);
  }
  const LdwZone zr = gen_ldw_zone(w, rho_c, prs_c);
  const LdwZone zt = gen_ldw_zone(w, a.V[iRHO * nz + o], g.iso ? 0.0 : a.V[iPRS * nz + o]);
  const double coef = w.sigma_e / (2.99792458e10 * w.unit_acc);
  double g_r = 0.0, g_t = 0.0;
  // per zone: the reciprocals the 36 bins share (<= 1 ulp each against the reference's divisions)
  const double inv_b0 = 1.0 / (x22[0] - x11[0]), inv_b1 = 1.0 / (x22[1] - x11[1]);
  const double inv_maxds = 1.0 / maxds;
  const double vUV[2][4] = {{v11[0] * w.UV, v12[0] * w.UV, v21[0] * w.UV, v22[0] * w.UV},
                            {v11[1] * w.UV, v12[1] * w.UV, v21[1] * w.UV, v22[1] * w.UV}};
  // dvds^(-alpha) (power law) or dvds itself (fit mode) of ONE ACTIVE bin, 0 when the gradient vanishes.
  // MODE 0: fit mode, 1: ALPHARAD = -0.6 (pow_three_fifths), 2: any exponent (exp / log).
  // FAST: straight-line code (no branch, so that the bins of a group interleave): the small-angle series for
  // t_off - theta_j and the 3/5 power; `slow` reports the (rare) bins that need atan() or exp / log, which are then
  // re-evaluated with FAST = false - same operations as before for every bin, chosen per bin instead of per branch.
  auto binD = [&](auto mode, auto fast, int ia, bool &slow) -> double {
    constexpr int MODE = decltype(mode)::value;
    constexpr bool FAST = decltype(fast)::value;
    const double sa = __ldg(w.sin_a + ia), ca = __ldg(w.cos_a + ia);
    const double dx1 = maxds * sa, dx2 = maxds * ca;
    const double X = x + dx1, Z = z + dx2;
    // r_off = sqrt(X^2 + Z^2); sin / cos of t_off = atan(X / Z) (principal branch: cos > 0) are X / r, Z / r
    const double s2 = X * X + Z * Z;
    const double rs = rsqrt_fast(s2);
    const double r_off = s2 * rs;
    const double co = fabs(Z) * rs, so = (Z < 0.0 ? -X : X) * rs;
    // t_off = atan(X / Z) = theta_j + atan(y), y = tan(t_off - theta_j) = (z dx1 - x dx2) / (z Z + x X): the offset
    // point is at most half a zone away, so |y| << 1 and the odd series converges in a few terms (|y| < 0.06:
    // next term y^14/15 < 6e-19); anything else takes atan() itself
    const double yy = (z * dx1 - x * dx2) * rcp_fast(z * Z + x * X);
    const bool small = Z > 0.0 && fabs(yy) < 0.06;
    double t_off;
    if (FAST || small) {
      const double y2 = yy * yy;
      const double p = 1.0 + y2 * (-1.0 / 3.0 + y2 * (1.0 / 5.0 + y2 * (-1.0 / 7.0 + y2 * (1.0 / 9.0 + y2 * (-1.0 / 11.0 + y2 * (1.0 / 13.0))))));
      t_off = x2j + yy * p;
    } else t_off = atan(X / Z);
    // bilinear() of the two velocity components at (r_off, t_off), line_connect.c:746-767
    const double f1 = (r_off - x11[0]) * inv_b0, f2 = (t_off - x11[1]) * inv_b1;
    const double a0 = (1.0 - f1) * vUV[0][0] + f1 * vUV[0][2], b0 = (1.0 - f1) * vUV[0][1] + f1 * vUV[0][3];
    const double a1 = (1.0 - f1) * vUV[1][0] + f1 * vUV[1][2], b1 = (1.0 - f1) * vUV[1][1] + f1 * vUV[1][3];
    const double q0 = (1.0 - f2) * a0 + f2 * b0, q1 = (1.0 - f2) * a1 + f2 * b1;     // ans2[] * UNIT_VELOCITY
    const double vx2 = (q0 * so + q1 * co);
    const double vz2 = (q0 * co - q1 * so);
    const double v1 = sa * vx1 + ca * vz1;
    const double v2 = sa * vx2 + ca * vz2;
    // ds = sqrt(dx1^2 + dx2^2) = maxds sqrt(sa^2 + ca^2): the bin's constant comes from the host table
    const double out = fabs(v2 - v1) * (inv_maxds * __ldg(w.inv_ca + ia));
    // LineForce() needs M = k (sigma_e rho v_th / dvds)^alpha per bin and SWEEP (the sweeps pass
    // different centre states); the bin-dependent factor dvds^(-alpha) is taken once and serves
    // both, so that a zone costs one pow() per sweep instead of 36.
    // exp(-alpha log x): |log x| is O(10) here, within a few ulp of pow(x, -alpha) at half its cost
    const bool in_range = out >= 1e-12 && out <= 1e12;
    double D;
    if (MODE == 0) D = out;                                   // fit mode keeps dvds itself
    else if (MODE == 1 && (FAST || in_range)) D = pow_three_fifths(out);
    else if (FAST) D = 0.0;                                   // MODE 2: exp / log in the second pass
    else D = out > 0.0 ? exp(-w.alpharad * log(out)) : 0.0;
    slow = FAST && out > 0.0 && (!small || (MODE == 1 && !in_range) || MODE == 2);
    return out > 0.0 ? D : 0.0;
  };
  // Only the bins with flux (mask bits) are visited: a bin without flux adds (1 + M(0)) coef * 0 to both sums.  NB active
  // bins per iteration: their (streaming) flux loads are issued together, the NB gradient evaluations are straight-line
  // and independent - the kernel waits on load and FP64 latencies, not on a pipe - and the sums take the bins in order.
  // The last group of a zone is padded with copies of its first bin (discarded).
  auto sweep_bins = [&](auto mode) {
    constexpr int NB = NBINS;
    unsigned long long rem = w.nangles < 64 ? (mk & ((1ull << w.nangles) - 1ull)) : mk;
    while (rem) {
      int ib[NB];
      bool on[NB], slow[NB];
      double fa[NB], ta[NB], Dq[NB];
#pragma unroll
      for (int q = 0; q < NB; q++) {
        on[q] = rem != 0ull;
        ib[q] = on[q] ? __ffsll((long long)rem) - 1 : ib[0];
        rem &= rem - 1ull;                          // clears the lowest set bit; 0 stays 0
      }
#pragma unroll
      for (int q = 0; q < NB; q++) {
        fa[q] = on[q] ? __ldg(w.flux_r + ib[q] * nz + o) : 0.0;
        ta[q] = on[q] ? __ldg(w.flux_t + ib[q] * nz + o) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < NB; q++) Dq[q] = binD(mode, std::true_type{}, ib[q], slow[q]);
#pragma unroll
      for (int q = 0; q < NB; q++) {
        bool dummy;
        if (slow[q] && on[q]) Dq[q] = binD(mode, std::false_type{}, ib[q], dummy);
      }
#pragma unroll
      for (int q = 0; q < NB; q++) {
        if (on[q]) {
          // ((1 + M) sigma_e F / c) / UNIT_ACCELERATION with the two constant divisions folded into one
          // factor (<= 1 ulp per term; 144 FP64 divisions per zone and sweep otherwise)
          g_r += ((1.0 + gen_ldw_M(w, zr, Dq[q], o, nz)) * coef) * fa[q];
          g_t += ((1.0 + gen_ldw_M(w, zt, Dq[q], o, nz)) * coef) * ta[q];
        }
      }
    }
  };
  if (w.mpoints > 0) sweep_bins(std::integral_constant<int, 0>{});
  else if (w.alpha_m06) sweep_bins(std::integral_constant<int, 1>{});
  else sweep_bins(std::integral_constant<int, 2>{});
  w.gline[o] = g_r;
  w.gline[nz + o] = g_t;
}

// UserDefBoundary(side == 0) of cv_idl (init.c:199-316): floors over the WHOLE array and the
// mid-plane reset of the last active theta row; Uc of a floored zone is re-derived when the
// call sits inside stage >= 2 (PrimToCons3D on a 1-zone box, init.c:272-275)
template <int NV>
static __global__ void gen_ldw_floor(GenDev g, GenArgs a, int update_U) {
  const Dev &d = g.d;
  const LdwDev &w = g.ldw;
  long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= d.sv) return;
  const int i = (int)(o % d.tot[0]), j = (int)((o / d.sj) % d.tot[1]);
  double v[NV];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) v[nv] = a.V[nv * d.sv + o];
  bool convert = false;
  constexpr int P = pidx<NV>();
  const bool en = !g.iso;                   // the "#if EOS != ISOTHERMAL" blocks of init.c
  const int TRC = gen_nflx(g);
  if (v[iRHO] < w.dfloor) {
    if (v[iRHO] < 0.0) v[iRHO] = w.dfloor;
    const double cs = en ? sqrt(d.gas.gamma * v[P] / v[iRHO]) : 0.0;
    const double dfact = v[iRHO] / w.dfloor;
    v[iRHO] = w.dfloor;
    v[1] = dfact * v[1]; v[2] = dfact * v[2]; v[3] = dfact * v[3];
    if (en) {
      v[P] = (cs * cs) * v[iRHO] / d.gas.gamma;
      double temp = v[P] / v[iRHO] * w.kelvin_mu;
      if (temp < w.tfloor) { temp = w.tfloor; v[P] = v[iRHO] * temp / w.kelvin_mu; }
    }
#pragma unroll
    for (int nv = 4; nv < NV; nv++) if (nv == TRC) v[nv] = 0.0;
    convert = true;
  }
  if (en && v[P] < w.pfloor) { v[P] = w.pfloor; convert = true; }
  if (convert && update_U) {
    const double rho = v[iRHO];
    a.U[o] = rho;
    a.U[1 * d.sv + o] = rho * v[1];
    a.U[2 * d.sv + o] = rho * v[2];
    a.U[3 * d.sv + o] = rho * v[3];
    if (en) a.U[4 * d.sv + o] = 0.5 * rho * (v[1] * v[1] + v[2] * v[2] + v[3] * v[3]) + v[P] / d.gas.gmm1;
#pragma unroll
    for (int nv = 4; nv < NV; nv++) if (nv >= TRC) a.U[nv * d.sv + o] = rho * v[nv];
  }
  if (j == d.end[1]) {
    const double r = __ldg(w.xgc1 + i), theta = __ldg(w.xgc2 + j);
    const double sth = sin(theta), rcyl = r * sth;
    const double rho_mid = w.rho_0 * pow(r / w.r_WD, -1.0 * w.rho_alpha);
    v[2] = (v[iRHO] * v[2]) / rho_mid;
    v[iRHO] = rho_mid;
    v[1] = 0.0;
    v[3] = sqrt(w.gm_code / r) * sth;
    if (en) {
      const double temp = w.teff_wd * pow(w.r_WD / rcyl, 0.75) * pow(1.0 - sqrt(w.r_WD / rcyl), 0.25);
      v[P] = rho_mid * temp / w.kelvin_mu;
    }
#pragma unroll
    for (int nv = 4; nv < NV; nv++) if (nv == TRC) v[nv] = 1.0;
    convert = true;
  }
  if (convert) {
#pragma unroll
    for (int nv = 0; nv < NV; nv++) a.V[nv * d.sv + o] = v[nv];
  }
}

// UserDefBoundary(X1_BEG / X1_END / X2_BEG) of cv_idl (init.c:319-363)
static __global__ void gen_ldw_side(GenDev g, double *V, int side) {
  const Dev &d = g.d;
  const int ng = d.beg[side / 2];
  int ext[3] = {d.tot[0], d.tot[1], d.tot[2]};
  ext[side / 2] = ng;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)ext[0] * ext[1] * ext[2]) return;
  int c0 = (int)(t % ext[0]), c1 = (int)((t / ext[0]) % ext[1]), c2 = (int)(t / ((long)ext[0] * ext[1]));
  long o, os, ob = 0;
  if (side == 0) { o = (long)c2 * d.sk + (long)c1 * d.sj + c0; os = (long)c2 * d.sk + (long)c1 * d.sj + d.beg[0]; }
  else if (side == 1) { o = (long)c2 * d.sk + (long)c1 * d.sj + d.end[0] + 1 + c0; os = (long)c2 * d.sk + (long)c1 * d.sj + d.end[0]; }
  else {
    o = (long)c2 * d.sk + (long)c1 * d.sj + c0;
    os = (long)c2 * d.sk + (long)(2 * d.beg[1] - c1 - 1) * d.sj + c0;
    ob = (long)c2 * d.sk + (long)d.beg[1] * d.sj + c0;
  }
  for (int nv = 0; nv < g.nvar; nv++) V[nv * d.sv + o] = V[nv * d.sv + os];
  if (side == 0) V[1 * d.sv + o] = fmin(V[1 * d.sv + o], 0.0);
  else if (side == 1) V[1 * d.sv + o] = fmax(V[1 * d.sv + o], 0.0);
  else {
    V[2 * d.sv + o] *= -1.0;
    V[o] = V[ob];
    if (!g.iso) V[iPRS * d.sv + o] = V[iPRS * d.sv + ob];
  }
}

// ---- RightHandSide + RightHandSideSource + U += rhs + C_dt -------------------------------
// rhs of ONE zone from the fluxes of its two faces along dir (fp: upper face n+1/2, fm: lower face, global variable
// order, pressure and cmax passed separately) and the centre state vg (stateC->v, or (vp + vm)/2 for the spherical
// r sweep, rhs_source.c:229-232).  cdt_c: the zone's C_dt term of this direction (DIMENSIONS > 1);
// inv_max: the 1-D invDt_hyp candidates.  Shared by gen_rhs and the fused gen_sweep.
template <int NV>
PB_D void gen_zone_rhs(const GenDev &g, const GenArgs &a, int dir, int i, int j, int k, long o, double (&fp)[NV],
                       double (&fm)[NV], double pp, double pm, double cp_, double cm_, const double (&vg)[NV],
                       double (&rhs)[NV], double &cdt_c, double &inv_max, int ldw_mode = 0, double *mf_out = nullptr) {
  // ldw_mode 1: leave the line force of this (r) sweep to the theta sweep; 2: add it here (theta sweep)
  const Dev &d = g.d;
  const long nz = d.sv;
  const int n = dir == 0 ? i : (dir == 1 ? j : k);
  const double dt = *a.dt;
  {
    const double frp = fp[iRHO], frm = fm[iRHO];      // mass fluxes (not area weighted)
    double dpn;   // pressure-gradient term of the normal momentum
    if (g.geometry == GEO_CARTESIAN) {
      const double scrh = dt / __ldg(g.dx[dir] + n);
#pragma unroll
      for (int nv = 0; nv < NV; nv++) rhs[nv] = -scrh * (fp[nv] - fm[nv]);
      dpn = scrh * (pp - pm);
    } else {
      // TotalFlux (rhs.c:530-600): fA = F A, fA[iMPHI] *= |x1p| (r sweep) or |sp| (theta sweep)
      int m[3] = {i, j, k};
      m[dir] -= 1;
      const double Ap = gen_A(g, dir, k, j, i), Am = gen_A(g, dir, m[2], m[1], m[0]);
#pragma unroll
      for (int nv = 0; nv < NV; nv++) { fp[nv] = fp[nv] * Ap; fm[nv] = fm[nv] * Am; }
      // angular momentum: iMPHI = VX2 (POLAR: r, phi, z) or VX3 (CYLINDRICAL r, z [, phi]; SPHERICAL r, theta, phi)
      const bool mphi2 = g.geometry == GEO_POLAR;
      if (dir == 0) {
        const double ap = fabs(__ldg(g.xr[0] + n)), am = fabs(__ldg(g.xr[0] + n - 1));
        if (mphi2) { fp[2] *= ap; fm[2] *= am; } else { fp[3] *= ap; fm[3] *= am; }
      } else if (dir == 1 && g.geometry == GEO_SPHERICAL) { fp[3] *= fabs(__ldg(g.sp + n)); fm[3] *= fabs(__ldg(g.sp + n - 1)); }
      const double dtdV = dt / __ldg(g.dV + o);
      double dtdl = dt / __ldg(g.dx[dir] + n);
      if (dir != 0) dtdl = dtdl * __ldg(g.dx_dl[dir] + (long)j * d.tot[0] + i);
#pragma unroll
      for (int nv = 0; nv < NV; nv++) rhs[nv] = -dtdV * (fp[nv] - fm[nv]);
      dpn = dtdl * (pp - pm);
      if (dir == 0) { if (mphi2) rhs[2] /= fabs(__ldg(g.x[0] + n)); else rhs[3] /= fabs(__ldg(g.x[0] + n)); }
      else if (dir == 1 && g.geometry == GEO_SPHERICAL) rhs[3] /= fabs(__ldg(g.s + n));
    }
    double sn = 0.0;   // source of the normal momentum
    if (g.geometry == GEO_SPHERICAL && dir == 0) {
      const double r_1 = 1.0 / __ldg(g.x[0] + n);
      const double Sm = vg[iRHO] * (vg[2] * vg[2] + vg[3] * vg[3]);
      sn = dt * Sm * r_1;
    } else if ((g.geometry == GEO_CYLINDRICAL || g.geometry == GEO_POLAR) && dir == 0) {   // rhs_source.c:201-227
      const double r_1 = 1.0 / __ldg(g.x[0] + n);
      const double vphi = g.geometry == GEO_POLAR ? vg[2] : vg[3];
      sn = dt * (vg[iRHO] * vphi * vphi - 0.0) * r_1;
    } else if (g.geometry == GEO_SPHERICAL && dir == 1) {
      const double r_1 = 1.0 / __ldg(g.rt + i);
      const double ct = __ldg(g.cot + n);          // 1/tan(x2[j]) with the host's libm tan(), like the reference
      const double Sm = vg[iRHO] * (-vg[2] * vg[1] + ct * vg[3] * vg[3]);
      sn = dt * Sm * r_1;
    }
    // accumulate in the reference's order: flux difference, pressure gradient, geometry, forces
    double rn = (dir == 0 ? rhs[1] : (dir == 1 ? rhs[2] : rhs[3])) - dpn;
    if ((g.geometry == GEO_SPHERICAL && dir <= 1) || ((g.geometry == GEO_CYLINDRICAL || g.geometry == GEO_POLAR) && dir == 0)) rn += sn;
    for (int pass = 0; pass < 3; pass++) {
      // pass 0: BodyForceVector (rhs_source.c:253-272,360-376,428-440); pass 1: BodyForcePotential (:274-279,378-383,
      // 442-447); pass 2: LineForce, same pattern as pass 0 (:284-297,386-396,448-458)
      double gv[3];
      if (pass == 1) {
        if (!(d.bf_kind & 2)) continue;
        double dtdx;
        if (dir == 0) dtdx = dt / __ldg(g.dx[0] + n);
        else if (dir == 1) {
          double scrh = dt;
          if (g.geometry == GEO_POLAR) scrh /= __ldg(g.x[0] + i);
          else if (g.geometry == GEO_SPHERICAL) scrh /= __ldg(g.rt + i);
          dtdx = scrh / __ldg(g.dx[1] + n);
        } else {
          double scrh = dt;
          if (g.geometry == GEO_SPHERICAL) scrh *= __ldg(g.dx_dl[2] + (long)j * d.tot[0] + i);   // dx2[j] / (rt[i] dmu[j])
          dtdx = scrh / __ldg(g.dx[2] + n);
        }
        int m[3] = {i, j, k};
        m[dir] -= 1;
        const double php = bf_at(d, 4 + dir, i, j, k), phm = bf_at(d, 4 + dir, m[0], m[1], m[2]);
        rn -= dtdx * vg[iRHO] * (php - phm);
        if (!g.iso) rhs[pidx<NV>()] -= bf_at(d, 3, i, j, k) * rhs[iRHO];
        continue;
      }
      if (pass == 0) {
        if (!(d.bf_kind & 1)) continue;
        gv[0] = bf_at(d, 0, i, j, k); gv[1] = bf_at(d, 1, i, j, k); gv[2] = bf_at(d, 2, i, j, k);
      } else {
        if (!g.ldw.on || ldw_mode == 1) continue;
        gv[0] = g.ldw.gline[o]; gv[1] = g.ldw.gline[nz + o]; gv[2] = 0.0;   // LineForce() sums taken by gen_vgrad
      }
      const double gd = dir == 0 ? gv[0] : (dir == 1 ? gv[1] : gv[2]);
      constexpr int P = pidx<NV>();
      const bool en = !g.iso;                // IF_ENERGY (rhs_source.c)
      rn += dt * vg[iRHO] * gd;
      if (en) rhs[P] += dt * 0.5 * (frp + frm) * gd;
      if (dir == 0 && d.ndim == 1) {
        rhs[2] += dt * vg[iRHO] * gv[1];
        if (en) rhs[P] += dt * vg[iRHO] * vg[2] * gv[1];
        rhs[3] += dt * vg[iRHO] * gv[2];
        if (en) rhs[P] += dt * vg[iRHO] * vg[3] * gv[2];
      }
      if (dir == 1 && d.ndim == 2) {
        rhs[3] += dt * vg[iRHO] * gv[2];
        if (en) rhs[P] += dt * vg[iRHO] * vg[3] * gv[2];
      }
    }
    if (dir == 0) rhs[1] = rn; else if (dir == 1) rhs[2] = rn; else rhs[3] = rn;
    if (mf_out) *mf_out = 0.5 * (frp + frm);
    if (ldw_mode == 2) {   // the r sweep's line force (rhs_source.c:284-297) with that sweep's centre density and mass flux
      const double g_r = g.ldw.gline[o];
      rhs[1] += dt * a.cen[o] * g_r;
      if (!g.iso) rhs[pidx<NV>()] += dt * a.cen[2 * nz + o] * g_r;
    }
    if (a.ibmask && a.ibmask[o]) {   // InternalBoundaryReset(), rhs.c:416-417
#pragma unroll
      for (int nv = 0; nv < NV; nv++) rhs[nv] = 0.0;
    }
    // GetInverse_dl (set_geometry.c:303-375) and C_dt (update_stage.c:303-322)
    double inv_dl = __ldg(g.inv_dx[dir] + n);
    if ((g.geometry == GEO_SPHERICAL || g.geometry == GEO_POLAR) && dir == 1) inv_dl = inv_dl * (1.0 / __ldg(g.x[0] + i));
    if (g.geometry == GEO_SPHERICAL && dir == 2) inv_dl = inv_dl * (1.0 / __ldg(g.x[0] + i)) / __ldg(g.sin2 + j);
    if (d.ndim > 1) {
      double q = 1.0;                                   // update_stage.c:305-311
      if (g.ring_average > 1 && dir == gen_ring_dir(g)) q = 1.0 / gen_ring_of(g, i, j);
      cdt_c = 0.5 * (cm_ + cp_) * inv_dl * q;
    } else {
      // 1-D: every stage, faces IBEG-1..IEND with inv_dl of the face's left zone
      inv_max = cp_ * inv_dl;
      if (n == d.beg[0]) inv_max = fmax(inv_max, cm_ * __ldg(g.inv_dx[0] + n - 1));
    }
  }
}

template <int NV>
static __global__ void gen_rhs(GenDev g, GenArgs a, GenBox b) {
  int i, j, k;
  const Dev &d = g.d;
  double inv_max = 0.0;
  if (gen_zone(b.lo, b.hi, i, j, k)) {
    const int dir = a.dir;
    const long st = dir == 0 ? 1 : (dir == 1 ? d.sj : d.sk);
    const long o = (long)k * d.sk + (long)j * d.sj + i;
    const long nz = d.sv;
    double rhs[NV], fp[NV], fm[NV];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) { fp[nv] = a.F[nv * nz + o]; fm[nv] = a.F[nv * nz + o - st]; }
    const double pp = a.F[NV * nz + o], pm = a.F[NV * nz + o - st];
    const double cp_ = a.F[(NV + 1) * nz + o], cm_ = a.F[(NV + 1) * nz + o - st];
    // centre state (stateC->v) and, for the spherical r sweep, vc = (vp + vm)/2  (rhs_source.c:229-232)
    double vg[NV];
    if (g.geometry != GEO_CARTESIAN && dir == 0) {       // spherical, cylindrical, polar r sweep: (vp + vm)/2
#pragma unroll
      for (int nv = 0; nv < NV; nv++) vg[nv] = 0.5 * (a.VP[nv * nz + o] + a.VM[nv * nz + o]);
    } else {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) vg[nv] = a.V[nv * nz + o];
    }
    double cdt_c = 0.0;
    gen_zone_rhs<NV>(g, a, dir, i, j, k, o, fp, fm, pp, pm, cp_, cm_, vg, rhs, cdt_c, inv_max);
#pragma unroll
    for (int nv = 0; nv < NV; nv++) a.U[nv * nz + o] += rhs[nv];
    if (d.ndim > 1 && a.stage == 1) a.cdt[o] = (dir == 0) ? cdt_c : a.cdt[o] + cdt_c;
  }
  if (g.d.ndim == 1) {
    inv_max = warp_max(inv_max);
    if ((threadIdx.x & 31) == 0 && inv_max > 0.0) atomic_max_pos(a.red + 0, inv_max);
  }
}

// ---- the fused sweep: States -> Riemann -> RightHandSide of one direction in ONE kernel -------------------
// (north star: "each directional sweep is a single fused kernel": the L/R states, the fluxes and the right-hand
// side never leave the SM; the VP / VM / F arrays of the multi-kernel form are not touched.)
// Thread <-> zone.  A block covers S consecutive zones along the sweep direction for L lanes across it
// (dir 0: S = 128, L = 1, one row of i;  dir 1, 2: S = 16 rows/planes of L = 32 zones of i, so that global accesses
// stay coalesced along i).  The first and last zone of a tile only supply their states: S - 2 zones per tile are
// updated.  vm of zone n+1 and the flux of face n-1/2 reach zone n through shared memory (two barriers).
// FIRST (stage 1, first direction) also does PrimToCons3D + the U0 copy (rk_step.c:129-130) of its zones.
template <int NV, int S, int L, bool X, int MB = 1>
static __global__ void __launch_bounds__(S * L, MB) gen_sweep(GenDev g, GenArgs a, int first) {
  __shared__ double sh[NV + 2][S * L];
  const Dev &d = g.d;
  const int dir = a.dir;
  const int tid = threadIdx.x, s = tid / L, l = tid % L;
  // tile origin: blockIdx.x tiles the lanes (i for dir != 0), blockIdx.y the sweep direction, blockIdx.z the rest
  int idx[3];
  const int n = d.beg[dir] - 1 + (int)blockIdx.y * (S - 2) + s;
  if (dir == 0) { idx[0] = n; idx[1] = d.beg[1] + (int)blockIdx.x; idx[2] = d.beg[2] + (int)blockIdx.z; }
  else if (dir == 1) { idx[0] = d.beg[0] + (int)blockIdx.x * L + l; idx[1] = n; idx[2] = d.beg[2] + (int)blockIdx.z; }
  else { idx[0] = d.beg[0] + (int)blockIdx.x * L + l; idx[1] = d.beg[1] + (int)blockIdx.z; idx[2] = n; }
  const int i = idx[0], j = idx[1], k = idx[2];
  const bool lane_ok = (dir == 0 || i <= d.end[0]) && (dir == 1 || j <= d.end[1]) && (dir == 2 || k <= d.end[2]);
  const bool st_ok = lane_ok && n <= d.end[dir] + 1;            // States(nbeg-1, nend+1)
  const long st = dir == 0 ? 1 : (dir == 1 ? d.sj : d.sk);
  const long o = (long)k * d.sk + (long)j * d.sj + i;
  const long nz = d.sv;
  if (st_ok) {
    // the lines the Riemann and right-hand-side phases will read are requested (into L2) before the states are built:
    // a tile's warps wait on global-memory latency most of the time (ncu: long scoreboard), -2.4 % per C4 step
    auto pf = [&](const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
#pragma unroll
    for (int nv = 0; nv < NV; nv++) pf(a.U + nv * nz + o);
    pf(a.cdt + o);
    pf(g.dV + o);
    pf(g.A[dir] + g.Aoff[dir] + (long)k * g.Ask[dir] + (long)j * g.Asj[dir] + i);
    if (g.ldw.on) { pf(g.ldw.gline + o); pf(g.ldw.gline + nz + o); if (a.defer) { pf(a.cen + o); pf(a.cen + nz + o); pf(a.cen + 2 * nz + o); } }
  }
  double v[NV], vp[NV], vm[NV];
  if (st_ok) gen_zone_states<NV>(g, a.V, a.flag, dir, n, st, o, v, vp, vm);
  else {
#pragma unroll
    for (int nv = 0; nv < NV; nv++) v[nv] = vp[nv] = vm[nv] = 1.0;
  }
#pragma unroll
  for (int nv = 0; nv < NV; nv++) sh[nv][tid] = vm[nv];
  __syncthreads();
  const bool face_ok = st_ok && s < S - 1 && n <= d.end[dir];   // Riemann(nbeg-1, nend)
  double F[NV + 2], machv = 0.0;
  if (face_ok) {
    double vR[NV];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) vR[nv] = sh[nv][tid + L];
    const bool hll = g.flatten && ((a.flag[o] & GF_HLL) || (a.flag[o + st] & GF_HLL));
    machv = gen_face<NV, X>(g, dir, vp, vR, hll, F);
    if ((d.bf_kind & 2) && !g.iso) F[pidx<NV>()] += F[iRHO] * bf_at(d, 4 + dir, i, j, k);   // TotalFlux(), rhs.c:171-179,525
  } else {
#pragma unroll
    for (int nv = 0; nv < NV + 2; nv++) F[nv] = 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int nv = 0; nv < NV + 2; nv++) sh[nv][tid] = F[nv];
  __syncthreads();
  double inv_max = 0.0;
  if (face_ok && s >= 1 && n >= d.beg[dir]) {
    double fp[NV], fm[NV], rhs[NV], vg[NV];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) { fp[nv] = F[nv]; fm[nv] = sh[nv][tid - L]; }
    const double pp = F[NV], pm = sh[NV][tid - L], cp_ = F[NV + 1], cm_ = sh[NV + 1][tid - L];
    if (g.geometry != GEO_CARTESIAN && dir == 0) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) vg[nv] = 0.5 * (vp[nv] + vm[nv]);
    } else {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) vg[nv] = v[nv];
    }
    double cdt_c = 0.0, mf = 0.0;
    const int ldw_mode = (a.defer && g.ldw.on) ? (dir == 0 ? 1 : (dir == 1 ? 2 : 0)) : 0;
    gen_zone_rhs<NV>(g, a, dir, i, j, k, o, fp, fm, pp, pm, cp_, cm_, vg, rhs, cdt_c, inv_max, ldw_mode, &mf);
    if (ldw_mode == 1) { a.cen[o] = vg[iRHO]; a.cen[nz + o] = vg[pidx<NV>()]; a.cen[2 * nz + o] = mf; }
    if (first) {
      // PrimToCons3D + RBoxCopy(U0): exact restatement (no reciprocal sharing), these values seed U0
      double u[NV];
      const double rho = v[iRHO];
      u[0] = rho; u[1] = rho * v[1]; u[2] = rho * v[2]; u[3] = rho * v[3];
      double k2 = v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
      if (!g.iso) u[pidx<NV>()] = 0.5 * rho * k2 + v[pidx<NV>()] / d.gas.gmm1;
#pragma unroll
      for (int nv = 4; nv < NV; nv++) if (nv >= gen_nflx(g)) u[nv] = rho * v[nv];
#pragma unroll
      for (int nv = 0; nv < NV; nv++) { a.U0[nv * nz + o] = u[nv]; a.U[nv * nz + o] = u[nv] + rhs[nv]; }
    } else {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) a.U[nv * nz + o] += rhs[nv];
    }
    if (d.ndim > 1 && a.stage == 1) a.cdt[o] = (dir == 0) ? cdt_c : a.cdt[o] + cdt_c;
  }
  machv = warp_max(machv);
  if ((tid & 31) == 0 && machv > 0.0) atomic_max_pos(a.red + 1, machv);
  if (d.ndim == 1) {
    inv_max = warp_max(inv_max);
    if ((tid & 31) == 0 && inv_max > 0.0) atomic_max_pos(a.red + 0, inv_max);
  }
}

// ---- RING_AVERAGE, Src/ring_average.c ---------------------------------------------------------
// ConsToPrim (HD/mappers.c:98-290, entropy aware) of one zone; returns true when a floor was applied
template <int NV>
PB_D bool gen_c2p(const GenDev &g, double (&u)[NV], double (&v)[NV], unsigned short fl) {
  const Gas &gs = g.d.gas;
  constexpr int P = pidx<NV>();
  const double m2 = u[1] * u[1] + u[2] * u[2] + u[3] * u[3];
  bool bad = false;
  if (u[0] < 0.0) { u[0] = gs.small_dn; bad = true; }
  const double rho = u[0], tau = 1.0 / u[0];
  v[0] = rho; v[1] = u[1] * tau; v[2] = u[2] * tau; v[3] = u[3] * tau;
  const double kin = 0.5 * m2 / u[0];
  if (!g.iso) {
    if (u[P] < 0.0) { u[P] = gs.small_pr / gs.gmm1 + kin; bad = true; }
    if (g.entropy && (fl & GF_ENTROPY)) {
      const double rhog1 = pow(rho, gs.gmm1);
      v[P] = u[NV - 1] * rhog1;
      if (v[P] < 0.0) { v[P] = gs.small_pr; bad = true; }
      u[P] = v[P] / gs.gmm1 + kin;
    } else {
      v[P] = gs.gmm1 * (u[P] - kin);
      if (v[P] < 0.0) { v[P] = gs.small_pr; u[P] = v[P] / gs.gmm1 + kin; bad = true; }
      if (g.entropy) u[NV - 1] = v[P] / pow(rho, gs.gmm1);
    }
  }
#pragma unroll
  for (int nv = 4; nv < NV; nv++) if (nv >= gen_nflx(g)) v[nv] = u[nv] * tau;
  return bad;
}

// RingAverageCons() + ConsToPrim3D.  from_prim = 1: the round trip at the start of a step (rk_step.c:115-119:
// PrimToCons3D, RingAverageCons, ConsToPrim3D over the whole domain, averaged or not); 0: after a stage, on the
// rings only (gen_finish has left the combined U of those zones and converted all the others).
// One thread per zone; the first zone of a chunk sums its chunk in the reference's order and writes all its zones.
template <int NV>
static __global__ void gen_ring(GenDev g, GenArgs a, GenBox b, int from_prim) {
  int i, j, k;
  const Dev &d = g.d;
  int nfail = 0, nan = 0;
  if (gen_zone(b.lo, b.hi, i, j, k)) {
    const long nz = d.sv;
    const int cs = gen_ring_of(g, i, j);
    const int pd = gen_ring_dir(g);
    const int np = pd == 1 ? j : k;
    const long stp = pd == 1 ? d.sj : d.sk;
    const long o = (long)k * d.sk + (long)j * d.sj + i;
    const bool first = (np - d.beg[pd]) % cs == 0;
    if ((cs > 1 || from_prim) && first) {
      double uav[NV], dVav = 0.0;
#pragma unroll
      for (int nv = 0; nv < NV; nv++) uav[nv] = 0.0;
      for (int m = 0; m < cs; m++) {
        const long om = o + m * stp;
        double u[NV];
        if (from_prim) {
          double v[NV];
#pragma unroll
          for (int nv = 0; nv < NV; nv++) v[nv] = a.V[nv * nz + om];
          const double rho = v[iRHO];
          u[0] = rho; u[1] = rho * v[1]; u[2] = rho * v[2]; u[3] = rho * v[3];
          const double k2 = v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
          if (!g.iso) u[pidx<NV>()] = 0.5 * rho * k2 + v[pidx<NV>()] / d.gas.gmm1;
#pragma unroll
          for (int nv = 4; nv < NV; nv++) if (nv >= gen_nflx(g)) u[nv] = rho * v[nv];
        } else {
#pragma unroll
          for (int nv = 0; nv < NV; nv++) u[nv] = a.U[nv * nz + om];
        }
        if (cs == 1) {
#pragma unroll
          for (int nv = 0; nv < NV; nv++) uav[nv] = u[nv];
        } else {
          const double dv = __ldg(g.dV + om);
          dVav += dv;
#pragma unroll
          for (int nv = 0; nv < NV; nv++) uav[nv] += u[nv] * dv;
        }
      }
      if (cs > 1) {
#pragma unroll
        for (int nv = 0; nv < NV; nv++) uav[nv] = uav[nv] / dVav;
      }
      for (int m = 0; m < cs; m++) {
        const long om = o + m * stp;
        double u[NV], v[NV];
#pragma unroll
        for (int nv = 0; nv < NV; nv++) u[nv] = uav[nv];
        unsigned short fl = a.flag[om];
        if (gen_c2p<NV>(g, u, v, fl)) { fl |= GF_C2P_FAIL; nfail = 1; a.flag[om] = fl; }
        nan |= !(v[1] == v[1]) || !(v[0] == v[0]) || (!g.iso && !(v[pidx<NV>()] == v[pidx<NV>()]));
#pragma unroll
        for (int nv = 0; nv < NV; nv++) { a.U[nv * nz + om] = u[nv]; a.V[nv * nz + om] = v[nv]; }
      }
    }
  }
  block_reduce(0.0, 0.0, nfail, nan, false, a.red);
}

// ---- RK combination + ConsToPrim3D (entropy aware) + dt reduction ---------------------------
template <int NV>
static __global__ void gen_finish(GenDev g, GenArgs a, GenBox b) {
  int i, j, k;
  const Dev &d = g.d;
  double cmaxv = 0.0;
  int nfail = 0, nan = 0;
  if (gen_zone(b.lo, b.hi, i, j, k)) {
    const long o = (long)k * d.sk + (long)j * d.sj + i;
    const long nz = d.sv;
    double u[NV], v[NV];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) u[nv] = a.U[nv * nz + o];
    if (a.comb == 1) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) u[nv] = a.w0 * a.U0[nv * nz + o] + a.wc * u[nv];
    } else if (a.comb == 2) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) u[nv] = (1.0 / 3.0) * (a.U0[nv * nz + o] + 2.0 * u[nv]);
    }
    if (gen_ring_of(g, i, j) > 1) {
      // RING_AVERAGE: RingAverageCons comes between the combination and ConsToPrim3D (rk_step.c:167-169,238-240,
      // 306-308): gen_ring converts these zones
#pragma unroll
      for (int nv = 0; nv < NV; nv++) a.U[nv * nz + o] = u[nv];
    } else {
      // ConsToPrim, mappers.c:98-290
      unsigned short fl = a.flag[o];
      if (gen_c2p<NV>(g, u, v, fl)) { fl |= GF_C2P_FAIL; nfail = 1; a.flag[o] = fl; }
      nan = !(v[1] == v[1]) || !(v[0] == v[0]) || (!g.iso && !(v[pidx<NV>()] == v[pidx<NV>()]));
#pragma unroll
      for (int nv = 0; nv < NV; nv++) { a.U[nv * nz + o] = u[nv]; a.V[nv * nz + o] = v[nv]; }
    }
    if (a.stage == 1 && d.ndim > 1) cmaxv = a.cdt[o];
  }
  block_reduce(cmaxv, 0.0, nfail, nan, a.stage == 1 && g.d.ndim > 1, a.red);
}

}  // namespace pb
