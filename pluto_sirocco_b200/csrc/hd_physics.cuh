// hd_physics.cuh -- per-zone / per-interface FP64 building blocks of the HD update (device only).
//
// Everything here works on a state in SWEEP-LOCAL order
//     q[0]=rho  q[1]=v_n  q[2]=v_t  q[3]=v_b  q[4]=prs  q[5..]=scalars (tracers, entropy)
// i.e. the n/t/b permutation that the reference applies through the globals VXn/VXt/VXb
// (Src/set_indexes.c:18) is done once when a zone is loaded, so the physics is written
// once.  Functions are register-only and __forceinline__; all loops over NV have
// compile-time bounds so nothing is spilled to local memory.
//
// The step is FP64-issue bound on B200 (64 DP lanes/SM), so the code is written to MINIMISE
// DP instructions and to be BRANCH FREE (selects instead of divergent branches, so ptxas can
// interleave the independent dependency chains of a whole Riemann problem):
//   * 1/x and 1/sqrt(x) are MUFU.RCP64H / MUFU.RSQ64H seeds (2^-20, measured with
//     tools/probe_mufu.cu) + ONE cubic Newton step -> <= 1 ulp (2.2e-16), no slow-path branch.
//     CUDA's own a/b and sqrt() cost 8-9 DP instructions + a range check + a call.
//   * the sound speed and 1/rho of a state come from one rsqrt:  r = rsqrt(g p rho),
//     a = g p r, 1/rho = a r.
//   * HLLC evaluates conservative state, flux and star state only for the UPWIND side
//     (selected with SEL instructions), not for both.
//   * max Mach number is tracked as a (numerator, denominator) pair, compared by cross
//     multiplication; one division per thread at the end of the kernel.
// Tolerance contract: <= 1e-12 relative per step against the reference C build (measured
// ~1e-15): operation ORDER is kept where it decides the result (limiter branches, upwind
// choice, accumulation order), reciprocal sharing and FMA contraction are allowed.
//
// Reference behaviour reproduced:
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

#define PB_D __device__ __forceinline__

namespace pb {

enum { iRHO = 0, iVN = 1, iVT = 2, iVB = 3, iPRS = 4, NFLX = 5 };

// values shared with include/pluto_b200.h
enum Solver { SOLVER_TVDLF = 1, SOLVER_HLL = 2, SOLVER_HLLC = 3 };
enum Recon { RECON_FLAT = 1, RECON_LINEAR = 2, RECON_PARABOLIC = 3 };
enum Limiter { LIM_RT = -1,  // run-time choice (SweepArgs::limiter)
               LIM_DEFAULT = 0, LIM_FLAT = 1, LIM_MINMOD = 2, LIM_VANLEER = 3, LIM_MC = 4,
               LIM_VANALBADA = 5, LIM_OSPRE = 6, LIM_UMIST = 7 };

struct Gas {
  double gamma;      // g_gamma
  double gmm1;       // gamma - 1
  double inv_gmm1;   // 1/(gamma-1)
  double small_dn;   // g_smallDensity
  double small_pr;   // g_smallPressure
};

// ---- branch-free reciprocal / reciprocal square root ------------------------------------
// valid for normal, non-zero finite arguments (denormals flush: x -> inf/NaN like 1/0)
PB_D double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RCP64H, rel. error 2^-20
  double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);                          // cubic step: 2^-60 -> 1 ulp
}
PB_D double rsqrt_fast(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RSQ64H, rel. error 2^-20
  double t = x * r;
  double e = fma(-t, r, 1.0);
  return fma(fma(e, 0.375, 0.5), e * r, r);                // cubic step
}
PB_D double sqrt_fast(double x) {   // sqrt(0) = 0, sqrt(<0) = NaN like libm
  double s = x * rsqrt_fast(x);
  return x == 0.0 ? 0.0 : s;
}
PB_D double div_fast(double a, double b) { return a * rcp_fast(b); }
// x^(3/5) for 1e-12 <= x <= 1e12: with c = x^3, z -> c^(-1/5) by the division-free Newton step z <- z (6 - c z^5) / 5
// from a single-precision seed (ex2.approx(-0.6 lg2.approx(x)): relative error < 1e-5; each step maps e to 3 e^2:
// 3e-10, 3e-19), and c^(1/5) = c z^4.  ~20 FP64 operations where exp(0.6 log(x)) takes ~60; a few ulp like it.
PB_D double pow_three_fifths(double x) {
  float lx, sd;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lx) : "f"((float)x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sd) : "f"(-0.6f * lx));
  const double c = x * x * x;
  double z = (double)sd;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double z2 = z * z, z4 = z2 * z2;
    z = (z * 0.2) * fma(-c * z4, z, 6.0);
  }
  const double z2 = z * z;
  return c * (z2 * z2);
}

PB_D double sel(bool c, double a, double b) { return c ? a : b; }
PB_D double absmin(double a, double b) { return fabs(a) < fabs(b) ? a : b; }
// dp*dm > 0 for finite operands, without the DP multiply: same sign and both non-zero
PB_D bool same_sign_nz(double a, double b) {
  return ((__double2hiint(a) ^ __double2hiint(b)) >= 0) && (a != 0.0) && (b != 0.0);
}
// sign agreement only: enough for limiters whose value is 0 when either argument is 0
PB_D bool same_sign(double a, double b) { return (__double2hiint(a) ^ __double2hiint(b)) >= 0; }

// running maximum of a ratio n/d (d > 0) without dividing
struct Ratio {
  double n, d;
  PB_D void init() { n = 0.0; d = 1.0; }
  PB_D void update(double num, double den, bool on = true) {
    bool gt = (num * d > n * den) && on;
    n = sel(gt, num, n);
    d = sel(gt, den, d);
  }
  PB_D double value() const { return n / d; }
};

// ---- slope limiters on a uniform Cartesian grid (plm_coeffs.h:76-122) -------------------
// All return the HALF slope h = dv_lim/2, so that vp = v + h, vm = v - h
// (plm_states.c:155-162,256-257 with dp = dm = 1/2).
PB_D double half_mm(double dp, double dm) { return sel(same_sign(dp, dm), 0.5 * absmin(dp, dm), 0.0); }
PB_D double half_vl(double dp, double dm) {
  double pr = dp * dm;
  return sel(pr > 0.0, pr * rcp_fast(dp + dm), 0.0);       // (2 dp dm/(dp+dm))/2
}
PB_D double half_mc(double dp, double dm) {
  // absmin(0.5(dp+dm), 2 absmin(dp,dm))/2: the scalings by powers of two are exact
  double qc = 0.25 * (dm + dp);
  return sel(same_sign(dp, dm), absmin(absmin(dp, dm), qc), 0.0);
}
PB_D double half_va(double dp, double dm) {
  double pp = dp * dp, mm = dm * dm;
  double v = (dp * (mm + 1.e-18) + dm * (pp + 1.e-18)) * rcp_fast(pp + mm + 1.e-18);
  return sel(dp * dm > 0.0, 0.5 * v, 0.0);
}
PB_D double half_os(double dp, double dm) {
  double pr = dp * dm;
  double v = 0.75 * pr * (dm + dp) * rcp_fast(dp * dp + dm * dm + pr);
  return sel(pr > 0.0, v, 0.0);
}
PB_D double half_um(double dp, double dm) {
  double ddp = 0.25 * (dp + 3.0 * dm), ddm = 0.25 * (dm + 3.0 * dp);
  double d2 = 2.0 * absmin(dp, dm);
  d2 = absmin(d2, ddp);
  d2 = absmin(d2, ddm);
  return sel(dp * dm > 0.0, 0.5 * d2, 0.0);
}

// half slope of variable nv (sweep-local index) from the forward/backward differences.
// LIM == LIM_RT: run-time limiter `lim` (block-uniform switch).
template <int LIM>
PB_D double half_slope(int nv, double dp, double dm, int lim) {
  if (LIM == LIM_DEFAULT) {
    if (nv == iRHO) return half_mc(dp, dm);
    if (nv == iPRS) return half_mm(dp, dm);
    if (nv >= NFLX) return half_mc(dp, dm);
    return half_vl(dp, dm);
  }
  if (LIM == LIM_FLAT) return 0.0;
  if (LIM == LIM_MINMOD) return half_mm(dp, dm);
  if (LIM == LIM_VANLEER) return half_vl(dp, dm);
  if (LIM == LIM_MC) return half_mc(dp, dm);
  if (LIM == LIM_VANALBADA) return half_va(dp, dm);
  if (LIM == LIM_OSPRE) return half_os(dp, dm);
  if (LIM == LIM_UMIST) return half_um(dp, dm);
  switch (lim) {   // LIM_RT
    case LIM_FLAT: return 0.0;
    case LIM_MINMOD: return half_mm(dp, dm);
    case LIM_VANLEER: return half_vl(dp, dm);
    case LIM_MC: return half_mc(dp, dm);
    case LIM_VANALBADA: return half_va(dp, dm);
    case LIM_OSPRE: return half_os(dp, dm);
    case LIM_UMIST: return half_um(dp, dm);
    default: return half_slope<LIM_DEFAULT>(nv, dp, dm, 0);
  }
}

// PLM, uniform Cartesian, from the two one-sided differences of a zone
template <int NV, int LIM>
PB_D void plm_zone(const double (&v0)[NV], const double (&dvp)[NV], const double (&dvm)[NV],
                   double (&vp)[NV], double (&vm)[NV], int lim) {
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    double h = half_slope<LIM>(nv, dvp[nv], dvm[nv], lim);
    vp[nv] = v0[nv] + h;
    vm[nv] = v0[nv] - h;
  }
}

// PPM order 4, uniform Cartesian: unique interface value at i+1/2 from (i-1,i,i+1,i+2),
// clipped to lie between v[i] and v[i+1]  (ppm_states.c:150-160)
PB_D double ppm4_iface(double vm1, double v0, double vp1, double vp2) {
  const double w0 = -1.0 / 12.0, w1 = 7.0 / 12.0;
  double q = w0 * vm1 + w1 * v0 + w1 * vp1 + w0 * vp2;
  double dv = vp1 - v0;
  double dq = q - v0;
  return v0 + sel(dq * dv > 0.0, absmin(dq, dv), 0.0);
}
// parabola extremum limiter with Cartesian h=3 -> cm=cp=2   (ppm_states.c:196-214)
PB_D void ppm_parabola(double v0, double &vp, double &vm, double cp, double cm) {
  double dvp = vp - v0, dvm = vm - v0;
  const bool flat = dvp * dvm >= 0.0;
  const bool c1 = fabs(dvp) >= cm * fabs(dvm);
  const bool c2 = fabs(dvm) >= cp * fabs(dvp);
  double np = sel(c1, -cm * dvm, dvp);
  double nm = sel(!c1 && c2, -cp * dvp, dvm);
  vp = v0 + sel(flat, 0.0, np);
  vm = v0 + sel(flat, 0.0, nm);
}

// ---- primitive <-> conservative -------------------------------------------------------
template <int NV>
PB_D void prim2cons(const double (&v)[NV], double (&u)[NV], const Gas &g) {
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
// u is updated where the reference redefines it (mappers.c:139-218).  Branch free.
template <int NV>
PB_D int cons2prim(double (&u)[NV], double (&v)[NV], const Gas &g) {
  double m2 = u[iVN] * u[iVN] + u[iVT] * u[iVT] + u[iVB] * u[iVB];
  const bool f1 = u[iRHO] < 0.0;
  double rho = sel(f1, g.small_dn, u[iRHO]);
  u[iRHO] = rho;
  double tau = rcp_fast(rho);
  v[iRHO] = rho;
  v[iVN] = u[iVN] * tau;
  v[iVT] = u[iVT] * tau;
  v[iVB] = u[iVB] * tau;
  double kin = 0.5 * m2 * tau;
  const bool f2 = u[iPRS] < 0.0;
  double E = sel(f2, g.small_pr * g.inv_gmm1 + kin, u[iPRS]);
  double p = g.gmm1 * (E - kin);
  const bool f3 = p < 0.0;
  p = sel(f3, g.small_pr, p);
  E = sel(f3, p * g.inv_gmm1 + kin, E);
  u[iPRS] = E;
  v[iPRS] = p;
#pragma unroll
  for (int nv = NFLX; nv < NV; nv++) v[nv] = u[nv] * tau;
  return (f1 ? 1 : 0) | (f2 ? 2 : 0) | (f3 ? 4 : 0);
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

// sound speed a = sqrt(gamma p/rho) and r = 1/sqrt(gamma p rho)  (1/rho = a*r)
PB_D void sound(double rho, double p, double gamma, double &a, double &r) {
  double gp = gamma * p;
  r = rsqrt_fast(gp * rho);
  a = gp * r;
}

template <int NV>
PB_D void hll_average(const double (&vL)[NV], const double (&vR)[NV], double SL, double SR,
                      const Gas &g, Face<NV> &o) {
  double uL[NV], uR[NV];
  prim2cons<NV>(vL, uL, g);
  prim2cons<NV>(vR, uR, g);
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
  const bool supL = SL > 0.0, supR = SR < 0.0;
  double s = rcp_fast(SR - SL);
  double sLR = SL * SR;
#pragma unroll
  for (int nv = 0; nv < NFLX; nv++) {
    double h = (sLR * (uR[nv] - uL[nv]) + SR * fL[nv] - SL * fR[nv]) * s;
    o.f[nv] = sel(supL, fL[nv], sel(supR, fR[nv], h));
  }
  double ph = (SR * vL[iPRS] - SL * vR[iPRS]) * s;
  o.prs = sel(supL, vL[iPRS], sel(supR, vR[iPRS], ph));
}

// force_hll: MULTID shock flattening switches flagged interfaces to HLL (hllc.c:97-117)
template <int NV, int SOLVER>
PB_D void riemann(const double (&vL)[NV], const double (&vR)[NV], const Gas &g, Face<NV> &o,
                  Ratio &mach, bool mach_on = true, bool force_hll = false) {
  if (SOLVER == SOLVER_TVDLF) {
    // Rusanov flux on the arithmetic-mean state, |v_n| averaged  (tvdlf.c:100-129)
    double uL[NV], uR[NV];
    prim2cons<NV>(vL, uL, g);
    prim2cons<NV>(vR, uR, g);
    double rho = 0.5 * (vL[iRHO] + vR[iRHO]);
    double prs = 0.5 * (vL[iPRS] + vR[iPRS]);
    double vn = 0.5 * (fabs(vL[iVN]) + fabs(vR[iVN]));
    double a, r;
    sound(rho, prs, g.gamma, a, r);
    double c = fmax(fabs(vn + a), fabs(vn - a));
    o.cmax = c;
    mach.update(vn, a, mach_on);
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
#pragma unroll
    for (int nv = 0; nv < NFLX; nv++) o.f[nv] = 0.5 * (fL[nv] + fR[nv] - c * (uR[nv] - uL[nv]));
    o.prs = 0.5 * (vL[iPRS] + vR[iPRS]);
  } else {
    double aL, rL, aR, rR;
    sound(vL[iRHO], vL[iPRS], g.gamma, aL, rL);
    sound(vR[iRHO], vR[iPRS], g.gamma, aR, rR);
    const double SL = fmin(vL[iVN] - aL, vR[iVN] - aR);
    const double SR = fmax(vL[iVN] + aL, vR[iVN] + aR);
    mach.update(fabs(vL[iVN]) + fabs(vR[iVN]), aL + aR, mach_on);
    o.cmax = fmax(fabs(SL), fabs(SR));
    if (SOLVER == SOLVER_HLL) {
      hll_average<NV>(vL, vR, SL, SR, g, o);
    } else {  // HLLC, hllc.c:119-178, evaluated for the upwind side only
      const double dL = vL[iVN] - SL, dR = vR[iVN] - SR;
      const double wL = vL[iRHO] * dL, wR = vR[iRHO] * dR;
      const double qL = fma(wL, vL[iVN], vL[iPRS]);       // pL + mL (vL - SL)
      const double qR = fma(wR, vR[iVN], vR[iPRS]);
      const double vs = (qR - qL) * rcp_fast(wR - wL);
      const bool supL = SL > 0.0, supR = SR < 0.0;
      const bool left = supL || (!supR && vs >= 0.0);
      const bool sup = supL || supR;
      const double rho = sel(left, vL[iRHO], vR[iRHO]);
      const double vn = sel(left, vL[iVN], vR[iVN]);
      const double vt = sel(left, vL[iVT], vR[iVT]);
      const double vb = sel(left, vL[iVB], vR[iVB]);
      const double p = sel(left, vL[iPRS], vR[iPRS]);
      const double S = sel(left, SL, SR);
      const double w = sel(left, wL, wR);                  // rho (vn - S)
      const double irho = sel(left, aL * rL, aR * rR);     // 1/rho
      // conservative state and flux of the upwind side
      const double mn = rho * vn, mt = rho * vt, mb = rho * vb;
      const double k2 = vn * vn + vt * vt + vb * vb;
      const double E = 0.5 * rho * k2 + p * g.inv_gmm1;
      double f[NFLX];
      f[iRHO] = mn;
      f[iVN] = mn * vn;
      f[iVT] = mt * vn;
      f[iVB] = mb * vn;
      f[iPRS] = (E + p) * vn;
      // star state: us_rho = rho (S - vn)/(S - vs); the correction S (us - u) is switched
      // off (S -> 0, finite operands) where the face is supersonic
      const double den = sel(sup, 1.0, S - vs);
      const double usr = -w * rcp_fast(den);
      const double rw = rcp_fast(w);                       // -1/(rho (S - vn))
      const double usE = (E * irho + (vs - vn) * fma(-p, rw, vs)) * usr;
      const double Se = sel(sup, 0.0, S);
      const double sd = Se * (usr - rho);
      o.f[iRHO] = f[iRHO] + sd;
      o.f[iVN] = fma(Se, fma(usr, vs, -mn), f[iVN]);
      o.f[iVT] = fma(sd, vt, f[iVT]);
      o.f[iVB] = fma(sd, vb, f[iVB]);
      o.f[iPRS] = fma(Se, usE - E, f[iPRS]);
      o.prs = p;
      if (force_hll) {   // rare, block-divergent at most around shocks
        Face<NV> h;
        hll_average<NV>(vL, vR, SL, SR, g, h);
#pragma unroll
        for (int nv = 0; nv < NFLX; nv++) o.f[nv] = h.f[nv];
        o.prs = h.prs;
      }
    }
  }
  // passive scalars: upwind on the sign of the mass flux  (adv_flux.c:61-72)
#pragma unroll
  for (int nv = NFLX; nv < NV; nv++) o.f[nv] = o.f[iRHO] * sel(o.f[iRHO] > 0.0, vL[nv], vR[nv]);
}

}  // namespace pb
