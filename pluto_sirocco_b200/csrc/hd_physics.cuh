// hd_physics.cuh -- per-zone / per-interface FP64 building blocks of the HD update.
//
// Everything here works on a state in SWEEP-LOCAL order
//     q[0]=rho  q[1]=v_n  q[2]=v_t  q[3]=v_b  q[4]=prs  q[5..]=scalars (tracers, entropy)
// i.e. the n/t/b permutation that the reference applies through the globals VXn/VXt/VXb
// (Src/set_indexes.c:18) is done once when a zone is loaded, so the physics is written
// once.  Functions are register-only and __forceinline__; all loops over NV have
// compile-time bounds so nothing is spilled to local memory.
//
// Reference behaviour reproduced (tolerance contract: <=1e-12 relative per step, so FMA
// contraction and reciprocal sharing are allowed; operation ORDER is kept where it
// decides the result, e.g. limiter branches and the upwind choice):
//   limiters       Src/States/plm_coeffs.h:72-152
//   PLM states     Src/States/plm_states.c:141-258 (LIMITER DEFAULT: MC rho, VL v, MM p, MC scalars)
//   PPM4 states    Src/States/ppm_states.c:150-214, weights Src/States/ppm_coeffs.c:490-495
//   PrimToCons     Src/HD/mappers.c:44-56      ConsToPrim  Src/HD/mappers.c:118-218
//   Flux           Src/HD/fluxes.c:36-47       SoundSpeed2 Src/EOS/Ideal/eos.c:33
//   HLL_Speed      Src/HD/hll_speed.c:76-90    (Davis estimate + g_maxMach)
//   HLLC / HLL / LF  Src/HD/hllc.c:70-178, Src/HD/hll.c:72-96, Src/HD/tvdlf.c:100-130
//   AdvectFlux     Src/adv_flux.c:61-72
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD static inline
#endif

namespace pb {

enum { iRHO = 0, iVN = 1, iVT = 2, iVB = 3, iPRS = 4, NFLX = 5 };

// values shared with include/pluto_b200.h
enum Solver { SOLVER_TVDLF = 1, SOLVER_HLL = 2, SOLVER_HLLC = 3 };
enum Recon { RECON_FLAT = 1, RECON_LINEAR = 2, RECON_PARABOLIC = 3 };
enum Limiter { LIM_DEFAULT = 0, LIM_FLAT = 1, LIM_MINMOD = 2, LIM_VANLEER = 3, LIM_MC = 4,
               LIM_VANALBADA = 5, LIM_OSPRE = 6, LIM_UMIST = 7 };

struct Gas {
  double gamma;      // g_gamma
  double gmm1;       // gamma - 1
  double inv_gmm1;   // 1/(gamma-1)
  double small_dn;   // g_smallDensity
  double small_pr;   // g_smallPressure
};

// fast reciprocal/division: correctly rounded division costs ~2x the DP-pipe slots of
// this form; the relative error (<~2 ulp) is far inside the 1e-12 contract.
PB_HD double pb_div(double a, double b) { return a / b; }

PB_HD double absmin(double a, double b) { return fabs(a) < fabs(b) ? a : b; }

// ---- slope limiters on a uniform Cartesian grid (plm_coeffs.h:76-122) -------------------
PB_HD double lim_mm(double dp, double dm) { return dp * dm > 0.0 ? absmin(dp, dm) : 0.0; }
PB_HD double lim_vl(double dp, double dm) {
  return dp * dm > 0.0 ? pb_div(2.0 * dp * dm, dp + dm) : 0.0;
}
PB_HD double lim_mc(double dp, double dm) {
  if (dp * dm > 0.0) {
    double qc = 0.5 * (dm + dp), s = 2.0 * absmin(dp, dm);
    return absmin(qc, s);
  }
  return 0.0;
}
PB_HD double lim_va(double dp, double dm) {
  if (dp * dm > 0.0) {
    double pp = dp * dp, mm = dm * dm;
    return pb_div(dp * (mm + 1.e-18) + dm * (pp + 1.e-18), pp + mm + 1.e-18);
  }
  return 0.0;
}
PB_HD double lim_os(double dp, double dm) {
  return dp * dm > 0.0 ? pb_div(1.5 * dp * dm * (dm + dp), dp * dp + dm * dm + dp * dm) : 0.0;
}
PB_HD double lim_um(double dp, double dm) {
  if (dp * dm > 0.0) {
    double ddp = 0.25 * (dp + 3.0 * dm), ddm = 0.25 * (dm + 3.0 * dp);
    double d2 = 2.0 * absmin(dp, dm);
    d2 = absmin(d2, ddp);
    return absmin(d2, ddm);
  }
  return 0.0;
}
// general-grid forms (plm_coeffs.h:128-149), used with the curvilinear coefficients
PB_HD double lim_vl_g(double dp, double dm, double cp, double cm) {
  return dp * dm > 0.0
             ? pb_div(dp * dm * (cp * dm + cm * dp), dp * dp + dm * dm + (cp + cm - 2.0) * dp * dm)
             : 0.0;
}
PB_HD double lim_mc_g(double dp, double dm, double cp, double cm) {
  if (dp * dm > 0.0) {
    double qc = 0.5 * (dm + dp), s = absmin(dp * cp, dm * cm);
    return absmin(qc, s);
  }
  return 0.0;
}

template <int LIM>
PB_HD double lim_one(double dp, double dm) {
  if (LIM == LIM_FLAT) return 0.0;
  if (LIM == LIM_MINMOD) return lim_mm(dp, dm);
  if (LIM == LIM_VANLEER) return lim_vl(dp, dm);
  if (LIM == LIM_MC) return lim_mc(dp, dm);
  if (LIM == LIM_VANALBADA) return lim_va(dp, dm);
  if (LIM == LIM_OSPRE) return lim_os(dp, dm);
  if (LIM == LIM_UMIST) return lim_um(dp, dm);
  return 0.0;
}

// limited slope of variable nv (sweep-local index) given forward/backward differences
template <int LIM>
PB_HD double plm_slope(int nv, double dp, double dm) {
  if (LIM == LIM_DEFAULT) {
    if (nv == iRHO) return lim_mc(dp, dm);
    if (nv == iPRS) return lim_mm(dp, dm);
    if (nv >= NFLX) return lim_mc(dp, dm);
    return lim_vl(dp, dm);
  }
  return lim_one<LIM>(dp, dm);
}

// PLM, uniform Cartesian: vp = v + dv*1/2, vm = v - dv*1/2   (plm_states.c:155-162,256-257)
template <int NV, int LIM>
PB_HD void plm_zone(const double (&vm1)[NV], const double (&v0)[NV], const double (&vp1)[NV],
                    double (&vp)[NV], double (&vm)[NV]) {
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    double dvp = vp1[nv] - v0[nv];
    double dvm = v0[nv] - vm1[nv];
    double dv = plm_slope<LIM>(nv, dvp, dvm);
    vp[nv] = v0[nv] + dv * 0.5;
    vm[nv] = v0[nv] - dv * 0.5;
  }
}

// PPM order 4, uniform Cartesian: unique interface value at i+1/2 from (i-1,i,i+1,i+2),
// clipped to lie between v[i] and v[i+1]  (ppm_states.c:150-160)
PB_HD double ppm4_iface(double vm1, double v0, double vp1, double vp2) {
  const double w0 = -1.0 / 12.0, w1 = 7.0 / 12.0;
  double q = w0 * vm1 + w1 * v0 + w1 * vp1 + w0 * vp2;
  double dv = vp1 - v0;
  double dq = q - v0;
  return v0 + lim_mm(dq, dv);
}
// parabola extremum limiter with Cartesian h=3 -> cm=cp=2   (ppm_states.c:196-214)
PB_HD void ppm_parabola(double v0, double &vp, double &vm, double cp, double cm) {
  double dvp = vp - v0, dvm = vm - v0;
  if (dvp * dvm >= 0.0) {
    dvp = dvm = 0.0;
  } else {
    if (fabs(dvp) >= cm * fabs(dvm)) dvp = -cm * dvm;
    else if (fabs(dvm) >= cp * fabs(dvp)) dvm = -cp * dvp;
  }
  vp = v0 + dvp;
  vm = v0 + dvm;
}

// ---- primitive <-> conservative -------------------------------------------------------
template <int NV>
PB_HD void prim2cons(const double (&v)[NV], double (&u)[NV], const Gas &g) {
  double rho = v[iRHO];
  u[iRHO] = rho;
  u[iVN] = rho * v[iVN];
  u[iVT] = rho * v[iVT];
  u[iVB] = rho * v[iVB];
  double k2 = v[iVN] * v[iVN] + v[iVT] * v[iVT] + v[iVB] * v[iVB];
  u[iPRS] = 0.5 * rho * k2 + v[iPRS] * g.inv_gmm1;
#pragma unroll
  for (int nv = NFLX; nv < NV; nv++) u[nv] = rho * v[nv];
}

// returns the FLAG_CONS2PRIM_FAIL-style status (0 ok; bit0 rho<0, bit1 E<0, bit2 p<0);
// u is updated where the reference redefines it (mappers.c:139-218)
template <int NV>
PB_HD int cons2prim(double (&u)[NV], double (&v)[NV], const Gas &g) {
  int fail = 0;
  double m2 = u[iVN] * u[iVN] + u[iVT] * u[iVT] + u[iVB] * u[iVB];
  if (u[iRHO] < 0.0) {
    u[iRHO] = g.small_dn;
    fail |= 1;
  }
  double rho = u[iRHO];
  double tau = 1.0 / rho;
  v[iRHO] = rho;
  v[iVN] = u[iVN] * tau;
  v[iVT] = u[iVT] * tau;
  v[iVB] = u[iVB] * tau;
  double kin = 0.5 * m2 * tau;
  if (u[iPRS] < 0.0) {
    u[iPRS] = g.small_pr * g.inv_gmm1 + kin;
    fail |= 2;
  }
  double p = g.gmm1 * (u[iPRS] - kin);
  if (p < 0.0) {
    p = g.small_pr;
    u[iPRS] = p * g.inv_gmm1 + kin;
    fail |= 4;
  }
  v[iPRS] = p;
#pragma unroll
  for (int nv = NFLX; nv < NV; nv++) v[nv] = u[nv] * tau;
  return fail;
}

// ---- Riemann solvers ------------------------------------------------------------------
// Face state produced by a solver: flux of (rho, m_n, m_t, m_b, E [, scalars]) WITHOUT the
// pressure in the normal momentum (USE_PRS_GRADIENT YES, Src/MHD/rhs.c:79), the interface
// pressure, and the fastest signal speed.
template <int NV>
struct Face {
  double f[NV];
  double prs;
  double cmax;
};

template <int NV, int SOLVER>
PB_HD void riemann(const double (&vL)[NV], const double (&vR)[NV], const Gas &g, Face<NV> &o,
                   double &maxMach, bool force_hll = false) {
  double uL[NV], uR[NV];
  prim2cons<NV>(vL, uL, g);
  prim2cons<NV>(vR, uR, g);
  double a2L = g.gamma * pb_div(vL[iPRS], vL[iRHO]);
  double a2R = g.gamma * pb_div(vR[iPRS], vR[iRHO]);
  double fL[NFLX], fR[NFLX];
  fL[iRHO] = uL[iVN];
  fL[iVN] = uL[iVN] * vL[iVN];
  fL[iVT] = uL[iVT] * vL[iVN];
  fL[iVB] = uL[iVB] * vL[iVN];
  fL[iPRS] = (uL[iPRS] + vL[iPRS]) * vL[iVN];
  fR[iRHO] = uR[iVN];
  fR[iVN] = uR[iVN] * vR[iVN];
  fR[iVT] = uR[iVT] * vR[iVN];
  fR[iVB] = uR[iVB] * vR[iVN];
  fR[iPRS] = (uR[iPRS] + vR[iPRS]) * vR[iVN];
  double pL = vL[iPRS], pR = vR[iPRS];

  if (SOLVER == SOLVER_TVDLF) {
    // Rusanov flux on the arithmetic-mean state, |v_n| averaged  (tvdlf.c:100-129)
    double rho = 0.5 * (vL[iRHO] + vR[iRHO]);
    double prs = 0.5 * (vL[iPRS] + vR[iPRS]);
    double vn = 0.5 * (fabs(vL[iVN]) + fabs(vR[iVN]));
    double a2 = g.gamma * pb_div(prs, rho);
    double a = sqrt(a2);
    double cmin_ = vn - a, cmax_ = vn + a;
    double c = fmax(fabs(cmax_), fabs(cmin_));
    o.cmax = c;
    maxMach = fmax(maxMach, pb_div(fabs(vn), a));
#pragma unroll
    for (int nv = 0; nv < NFLX; nv++) o.f[nv] = 0.5 * (fL[nv] + fR[nv] - c * (uR[nv] - uL[nv]));
    o.prs = 0.5 * (pL + pR);
  } else {
    double aL = sqrt(a2L), aR = sqrt(a2R);
    double SL = fmin(vL[iVN] - aL, vR[iVN] - aR);
    double SR = fmax(vL[iVN] + aL, vR[iVN] + aR);
    maxMach = fmax(maxMach, pb_div(fabs(vL[iVN]) + fabs(vR[iVN]), aL + aR));
    o.cmax = fmax(fabs(SL), fabs(SR));
    if (SL > 0.0) {
#pragma unroll
      for (int nv = 0; nv < NFLX; nv++) o.f[nv] = fL[nv];
      o.prs = pL;
    } else if (SR < 0.0) {
#pragma unroll
      for (int nv = 0; nv < NFLX; nv++) o.f[nv] = fR[nv];
      o.prs = pR;
    } else if (SOLVER == SOLVER_HLL || force_hll) {
      double s = 1.0 / (SR - SL);
#pragma unroll
      for (int nv = 0; nv < NFLX; nv++)
        o.f[nv] = (SL * SR * (uR[nv] - uL[nv]) + SR * fL[nv] - SL * fR[nv]) * s;
      o.prs = (SR * pL - SL * pR) * s;
    } else {  // HLLC
      double qL = vL[iPRS] + uL[iVN] * (vL[iVN] - SL);
      double qR = vR[iPRS] + uR[iVN] * (vR[iVN] - SR);
      double wL = vL[iRHO] * (vL[iVN] - SL);
      double wR = vR[iRHO] * (vR[iVN] - SR);
      double vs = pb_div(qR - qL, wR - wL);
      if (vs >= 0.0) {
        double us[NFLX];
        double dS = SL - vL[iVN];
        us[iRHO] = pb_div(uL[iRHO] * dS, SL - vs);
        us[iVN] = us[iRHO] * vs;
        us[iVT] = us[iRHO] * vL[iVT];
        us[iVB] = us[iRHO] * vL[iVB];
        us[iPRS] = pb_div(uL[iPRS], vL[iRHO]) +
                   (vs - vL[iVN]) * (vs + pb_div(vL[iPRS], vL[iRHO] * dS));
        us[iPRS] *= us[iRHO];
#pragma unroll
        for (int nv = 0; nv < NFLX; nv++) o.f[nv] = fL[nv] + SL * (us[nv] - uL[nv]);
        o.prs = pL;
      } else {
        double us[NFLX];
        double dS = SR - vR[iVN];
        us[iRHO] = pb_div(uR[iRHO] * dS, SR - vs);
        us[iVN] = us[iRHO] * vs;
        us[iVT] = us[iRHO] * vR[iVT];
        us[iVB] = us[iRHO] * vR[iVB];
        us[iPRS] = pb_div(uR[iPRS], vR[iRHO]) +
                   (vs - vR[iVN]) * (vs + pb_div(vR[iPRS], vR[iRHO] * dS));
        us[iPRS] *= us[iRHO];
#pragma unroll
        for (int nv = 0; nv < NFLX; nv++) o.f[nv] = fR[nv] + SR * (us[nv] - uR[nv]);
        o.prs = pR;
      }
    }
  }
  // passive scalars: upwind on the sign of the mass flux  (adv_flux.c:61-72)
#pragma unroll
  for (int nv = NFLX; nv < NV; nv++) o.f[nv] = o.f[iRHO] * (o.f[iRHO] > 0.0 ? vL[nv] : vR[nv]);
}

}  // namespace pb
