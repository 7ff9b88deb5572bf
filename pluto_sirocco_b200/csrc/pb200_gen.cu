// pb200_gen.cu -- host side of the general-grid path (gen_kernels.cuh): geometry set-up and the
// stage sequencing of AdvanceStep() for curvilinear / characteristic-limited / flattened /
// entropy-switched / line-driven-wind configurations.
//
// Geometry: the reference derives dV, A, dx_dl, xgc, rt, s, sp and the PLM coefficients from
// grid->xl/xr/dx once at start-up (Src/set_geometry.c:20-290, Src/States/plm_coeffs.c:66-88).
// gen_setup() restates those formulas on the host with the same libm calls (set-up code, not
// the hot path) and uploads the arrays; every time step then runs on the device only.
#include <math.h>
#include <string.h>

#include <vector>

#include "pb200_internal.h"
#include "gen_kernels.cuh"

using namespace pb;

namespace {

template <typename T>
T *upload(pb200_ctx *c, const std::vector<T> &h) {
  T *p = nullptr;
  if (cudaMalloc(&p, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
  c->gen_allocs.push_back(p);      // released by pb200_gen_release() whatever happens next
  if (cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return p;
}
template <typename T>
T *dalloc(pb200_ctx *c, size_t n) {
  T *p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  c->gen_allocs.push_back(p);
  if (cudaMemset(p, 0, n * sizeof(T)) != cudaSuccess) return nullptr;
  return p;
}

}  // namespace

void pb200_fill_ldw(pb200_ctx *c, pb::GenDev &G);
static void fill_ldw(pb200_ctx *c, GenDev &G) { pb200_fill_ldw(c, G); }

void pb200_gen_release(pb200_ctx *c) {
  for (void *p : c->gen_allocs) cudaFree(p);
  c->gen_allocs.clear();
  c->gen_ready = false;
  delete c->gdev;
  c->gdev = nullptr;
}

static void ppm_coefficients(int geo, int d, int nt, const double *xl, const double *xr, const double *dx, const double *xc,
                             bool uniform, std::vector<double> &w, std::vector<double> &hp, std::vector<double> &hm);

// fused-sweep tile shapes (alternatives measured on the 1024 x 512 line-driven wind: profiles/r02_v4_gen_tiles.txt)
#ifndef PB_GEN_MB0
#define PB_GEN_MB0 4      // resident r-sweep tiles per SM the register budget is cut for (128 registers: 16 warps instead of 12)
#endif
#ifndef PB_GEN_S0
#define PB_GEN_S0 128
#endif
#ifndef PB_GEN_S1
#define PB_GEN_S1 16
#endif
#ifndef PB_GEN_L1
#define PB_GEN_L1 32
#endif

int pb200_gen_setup(pb200_ctx *c) {
  if (c->gen_ready) return PB200_OK;
  pb200_gen_release(c);
  c->gdev = new GenDev;
  c->gen_epoch++;
  GenDev &G = *c->gdev;
  memset(&G, 0, sizeof(G));
  G.d = c->dev;
  const Dev &D = c->dev;
  const int n1 = D.tot[0], n2 = D.tot[1], n3 = D.tot[2], nd = D.ndim;
  const int geo = c->cfg.geometry;
  G.nvar = c->nvar;
  G.geometry = geo;
  G.limiter = c->cfg.limiter;
  G.char_lim = c->cfg.char_limiting;
  G.flatten = c->cfg.shock_flattening == 1;          // MULTID
  G.flatten_oned = c->cfg.shock_flattening == 2;     // ONED
  G.entropy = c->cfg.entropy_switch;
  G.solver = c->cfg.solver;
  G.iso = c->cfg.eos == PB200_EOS_ISOTHERMAL;
  G.cs2 = c->cfg.iso_sound_speed * c->cfg.iso_sound_speed;

  std::vector<double> x[3], xgc[3], inv[3], cp[3], cm[3], wp[3], wm[3], dp[3], dm[3];
  for (int d = 0; d < 3; d++) {
    int n = D.tot[d];
    x[d].resize(n); xgc[d].resize(n); inv[d].resize(n);
    for (int i = 0; i < n; i++) x[d][i] = 0.5 * (c->xl[d][i] + c->xr[d][i]);   // set_grid.c:137
  }
  const std::vector<double> &x1 = x[0], &x2 = x[1], &dx1 = c->dx[0], &dx2 = c->dx[1], &dx3 = c->dx[2];
  const std::vector<double> &x1p = c->xr[0], &x1m = c->xl[0], &x2p = c->xr[1], &x2m = c->xl[1];
  std::vector<double> rt(n1), s(n2), sp(n2), dmu(n2);
  for (int i = 0; i < n1; i++) {                 // set_geometry.c:83-97
    double xL = x1m[i], xR = x1p[i];
    if (geo == PB200_CARTESIAN) { xgc[0][i] = x1[i]; rt[i] = x1[i]; }
    else if (geo == PB200_CYLINDRICAL || geo == PB200_POLAR) {
      xgc[0][i] = x1[i] + dx1[i] * dx1[i] / (12.0 * x1[i]);
      rt[i] = x1[i];
    } else {
      xgc[0][i] = x1[i] + 2.0 * x1[i] * dx1[i] * dx1[i] / (12.0 * x1[i] * x1[i] + dx1[i] * dx1[i]);
      rt[i] = (xR * xR * xR - xL * xL * xL) / (xR * xR - xL * xL) / 1.5;
    }
  }
  for (int j = 0; j < n2; j++) {                 // set_geometry.c:103-116
    double xL = x2m[j], xR = x2p[j];
    if (geo != PB200_SPHERICAL) xgc[1][j] = x2[j];
    else {
      xgc[1][j] = sin(xR) - sin(xL) + xL * cos(xL) - xR * cos(xR);
      xgc[1][j] /= cos(xL) - cos(xR);
      sp[j] = fabs(sin(xR));
      s[j] = fabs(sin(x2[j]));
      dmu[j] = fabs(cos(xL) - cos(xR));
    }
  }
  for (int k = 0; k < n3; k++) xgc[2][k] = x[2][k];
  // volumes and areas (DIM_EXPAND: factors of the active dimensions only), set_geometry.c:122-230
  std::vector<double> dV((size_t)D.sv);
  for (int k = 0; k < n3; k++) for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) {
    double v;
    if (geo == PB200_CARTESIAN) { v = dx1[i]; if (nd > 1) v = v * dx2[j]; if (nd > 2) v = v * dx3[k]; }
    else if (geo == PB200_CYLINDRICAL || geo == PB200_POLAR) {     // set_geometry.c:124-129
      double dVr = fabs(x1[i]) * dx1[i];
      v = dVr; if (nd > 1) v = v * dx2[j]; if (nd > 2) v = v * (geo == PB200_POLAR ? dx3[k] : 1.0);
    } else {
      double dVr = fabs(x1p[i] * x1p[i] * x1p[i] - x1m[i] * x1m[i] * x1m[i]) / 3.0;
      double dm_ = fabs(cos(x2m[j]) - cos(x2p[j]));
      v = dVr; if (nd > 1) v = v * dm_; if (nd > 2) v = v * dx3[k];
    }
    dV[(size_t)k * D.sk + (size_t)j * D.sj + i] = v;
  }
  std::vector<double> A[3];
  for (int d = 0; d < 3; d++) {
    int e1 = n1 + (d == 0), e2 = n2 + (d == 1), e3 = n3 + (d == 2);
    G.Asj[d] = e1;
    G.Ask[d] = (long)e1 * e2;
    G.Aoff[d] = d == 0 ? 1 : (d == 1 ? G.Asj[d] : G.Ask[d]);
    A[d].assign((size_t)e1 * e2 * e3, 0.0);
  }
  for (int k = 0; k < n3; k++) for (int j = 0; j < n2; j++) for (int i = -1; i < n1; i++) {
    double a;
    if (geo == PB200_CARTESIAN) { a = 1.0; if (nd > 1) a = a * dx2[j]; if (nd > 2) a = a * dx3[k]; }
    else if (geo == PB200_CYLINDRICAL || geo == PB200_POLAR) {     // set_geometry.c:150-161
      a = (i == -1) ? fabs(x1m[0]) : fabs(x1p[i]);
      if (nd > 1) a = a * dx2[j];
      if (nd > 2) a = a * (geo == PB200_POLAR ? dx3[k] : 1.0);
    } else {
      double dm_ = fabs(cos(x2m[j]) - cos(x2p[j]));
      a = (i == -1) ? x1m[0] * x1m[0] : x1p[i] * x1p[i];
      if (nd > 1) a = a * dm_;
      if (nd > 2) a = a * dx3[k];
    }
    A[0][G.Aoff[0] + (long)k * G.Ask[0] + (long)j * G.Asj[0] + i] = a;
  }
  for (int k = 0; k < n3; k++) for (int j = -1; j < n2; j++) for (int i = 0; i < n1; i++) {
    double a;
    if (geo == PB200_CARTESIAN || geo == PB200_POLAR) { a = dx1[i]; if (nd > 1) a = a * 1.0; if (nd > 2) a = a * dx3[k]; }
    else if (geo == PB200_CYLINDRICAL) { a = fabs(x1[i]); if (nd > 1) a = a * dx1[i]; if (nd > 2) a = a * 1.0; }   // set_geometry.c:181-182
    else {
      a = fabs(x1[i]) * dx1[i];
      if (nd > 1) a = a * ((j == -1) ? fabs(sin(x2m[0])) : fabs(sin(x2p[j])));
      if (nd > 2) a = a * dx3[k];
    }
    A[1][G.Aoff[1] + (long)k * G.Ask[1] + (long)j * G.Asj[1] + i] = a;
  }
  for (int k = -1; k < n3; k++) for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) {
    double a;
    if (geo == PB200_CARTESIAN) { a = dx1[i]; if (nd > 1) a = a * dx2[j]; }
    else if (geo == PB200_CYLINDRICAL) a = 1.0;      // set_geometry.c:203-204
    else { a = fabs(x1[i]) * dx1[i]; if (nd > 1) a = a * dx2[j]; }
    A[2][G.Aoff[2] + (long)k * G.Ask[2] + (long)j * G.Asj[2] + i] = a;
  }
  std::vector<double> dxdl[3];
  for (int d = 0; d < 3; d++) dxdl[d].assign((size_t)n1 * n2, 1.0);
  if (geo == PB200_SPHERICAL)
    for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) {      // set_geometry.c:250-254
      dxdl[1][(size_t)j * n1 + i] = 1.0 / rt[i];
      dxdl[2][(size_t)j * n1 + i] = dx2[j] / (rt[i] * dmu[j]);
    }
  if (geo == PB200_POLAR)
    for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) dxdl[1][(size_t)j * n1 + i] = 1.0 / x1[i];   // set_geometry.c:228
  for (int d = 0; d < 3; d++) {
    int n = D.tot[d];
    cp[d].assign(n, 2.0); cm[d].assign(n, 2.0); wp[d].assign(n, 1.0); wm[d].assign(n, 1.0);
    dp[d].assign(n, 0.5); dm[d].assign(n, 0.5);
    const std::vector<double> &dx = c->dx[d], &xr = c->xr[d];
    for (int i = 0; i < n; i++) inv[d][i] = 1.0 / dx[i];
    for (int i = 1; i <= n - 2; i++) {           // plm_coeffs.c:66-88
      wp[d][i] = dx[i] / (xgc[d][i + 1] - xgc[d][i]);
      wm[d][i] = dx[i] / (xgc[d][i] - xgc[d][i - 1]);
      cp[d][i] = (xgc[d][i + 1] - xgc[d][i]) / (xr[i] - xgc[d][i]);
      cm[d][i] = (xgc[d][i] - xgc[d][i - 1]) / (xgc[d][i] - xr[i - 1]);
      dp[d][i] = (xr[i] - xgc[d][i]) / dx[i];
      dm[d][i] = (xgc[d][i] - xr[i - 1]) / dx[i];
    }
  }
  if (c->geo_set) {
    // the caller's Grid arrays (grid->dV, A[], dx_dl[], rt, s, sp of Src/set_geometry.c) take the place of this
    // function's own evaluation of the same formulas: a user-modified geometry is honoured as it is
    dV = c->geo_dV;
    for (int d = 0; d < 3; d++) { A[d] = c->geo_A[d]; dxdl[d] = c->geo_dxdl[d]; }
    rt = c->geo_rt; s = c->geo_s; sp = c->geo_sp;
  }
  bool ok = true;
  for (int d = 0; d < 3; d++) {
    ok &= (G.x[d] = upload(c, x[d])) != nullptr;
    ok &= (G.xr[d] = upload(c, c->xr[d])) != nullptr;
    ok &= (G.dx[d] = upload(c, c->dx[d])) != nullptr;
    ok &= (G.inv_dx[d] = upload(c, inv[d])) != nullptr;
    ok &= (G.cp[d] = upload(c, cp[d])) != nullptr;
    ok &= (G.cm[d] = upload(c, cm[d])) != nullptr;
    ok &= (G.wp[d] = upload(c, wp[d])) != nullptr;
    ok &= (G.wm[d] = upload(c, wm[d])) != nullptr;
    ok &= (G.dp[d] = upload(c, dp[d])) != nullptr;
    ok &= (G.dm[d] = upload(c, dm[d])) != nullptr;
    ok &= (G.A[d] = upload(c, A[d])) != nullptr;
    ok &= (G.dx_dl[d] = upload(c, dxdl[d])) != nullptr;
  }
  G.ring_average = c->cfg.ring_average > 1 ? c->cfg.ring_average : 0;
  G.ring_rec = G.ring_average ? (c->cfg.ring_average_rec ? c->cfg.ring_average_rec : 5) : 1;
  {   // RingAverageSize(), ring_average.c:620-735: chunk sizes halve ring by ring away from the axis
    const int rd = geo == PB200_POLAR ? 0 : 1;
    std::vector<int> cs(D.tot[rd], 0);
    if (G.ring_average) {
      int q = G.ring_average;
      for (int n = D.beg[rd]; n <= D.end[rd]; n++) {
        cs[n] = c->cfg.bc[2 * rd] == PB200_BC_POLARAXIS ? q : 1;
        if (q > 1) q >>= 1;
      }
      if (geo == PB200_SPHERICAL) {
        q = G.ring_average;
        for (int n = D.end[1]; n >= D.beg[1]; n--) {
          cs[n] = std::max(cs[n], c->cfg.bc[3] == PB200_BC_POLARAXIS ? q : 1);
          if (q > 1) q >>= 1;
        }
      }
    }
    ok &= (G.csize = upload(c, cs)) != nullptr;
  }
  G.ppm = c->cfg.reconstruction == PB200_PARABOLIC;
  for (int d = 0; d < 3 && G.ppm; d++) {
    std::vector<double> w, hp, hm;
    ppm_coefficients(geo, d, D.tot[d], c->xl[d].data(), c->xr[d].data(), c->dx[d].data(), x[d].data(),
                     pb200_grid_is_uniform(c, d), w, hp, hm);
    ok &= (G.pw[d] = upload(c, w)) != nullptr;
    ok &= (G.php[d] = upload(c, hp)) != nullptr;
    ok &= (G.phm[d] = upload(c, hm)) != nullptr;
  }
  {
    std::vector<double> cot(n2, 0.0), sn2(n2, 1.0);
    if (geo == PB200_SPHERICAL && nd > 1)
      for (int j = 0; j < n2; j++) { cot[j] = 1.0 / tan(x2[j]); sn2[j] = sin(x2[j]); }
    ok &= (G.cot = upload(c, cot)) != nullptr;
    ok &= (G.sin2 = upload(c, sn2)) != nullptr;
  }
  ok &= (G.rt = upload(c, rt)) != nullptr;
  ok &= (G.s = upload(c, s)) != nullptr;
  ok &= (G.sp = upload(c, sp)) != nullptr;
  ok &= (G.dV = upload(c, dV)) != nullptr;
  const size_t nz = (size_t)D.sv, nv = (size_t)c->nvar;
  ok &= (c->gU = dalloc<double>(c, nz * nv)) != nullptr;
  ok &= (c->gU0 = dalloc<double>(c, nz * nv)) != nullptr;
  ok &= (c->gVP = dalloc<double>(c, nz * nv)) != nullptr;
  ok &= (c->gVM = dalloc<double>(c, nz * nv)) != nullptr;
  ok &= (c->gF = dalloc<double>(c, nz * (nv + 2))) != nullptr;
  ok &= (c->gflag = dalloc<unsigned short>(c, nz)) != nullptr;
  ok &= (c->gshock = dalloc<unsigned char>(c, nz)) != nullptr;
  ok &= (c->gcdt = dalloc<double>(c, nz)) != nullptr;
  // line-driven wind: libm sin/cos tables of the angular bins and of theta, centroids
  {
    std::vector<double> sa(64), ca(64), ica(64), st(n2), ct(n2);
    for (int a = 0; a < 64; a++) {
      double theta_angle = (a + 0.5) * (2.0 * 3.14159265358979) / 36.0;     // line_connect.c:563
      sa[a] = sin(theta_angle); ca[a] = cos(theta_angle);
      ica[a] = 1.0 / sqrt(sa[a] * sa[a] + ca[a] * ca[a]);
    }
    ok &= (G.ldw.inv_ca = upload(c, ica)) != nullptr;
    for (int j = 0; j < n2; j++) { st[j] = sin(x2[j]); ct[j] = cos(x2[j]); }
    ok &= (G.ldw.sin_a = upload(c, sa)) != nullptr;
    ok &= (G.ldw.cos_a = upload(c, ca)) != nullptr;
    ok &= (G.ldw.sin_t = upload(c, st)) != nullptr;
    ok &= (G.ldw.cos_t = upload(c, ct)) != nullptr;
    ok &= (G.ldw.xgc1 = upload(c, xgc[0])) != nullptr;
    ok &= (G.ldw.xgc2 = upload(c, xgc[1])) != nullptr;
  }
  // the uploads are pageable copies on the legacy stream and the first stage kernels follow on c->stream
  if (!ok || cudaDeviceSynchronize() != cudaSuccess) { pb200_gen_release(c); return PB200_ENOMEM; }
  c->gen_ready = true;
  return PB200_OK;
}

// constants of the line-driven-wind problem (cv_idl/init.c:175-197, line_connect.c:831-846)
void pb200_fill_ldw(pb200_ctx *c, pb::GenDev &G) {
  LdwDev &w = G.ldw;
  w.on = c->ldw_on;
  if (!c->ldw_on) return;
  const pb200_ldw_config &L = c->ldw;
  const double amu = 1.66053886e-24, kB = 1.3806505e-16, Gc = 6.6726e-8, sigma = 5.67051e-5, sigmaT = 6.6524e-25;
  const double PI = 3.14159265358979;
  w.userdef_bc = L.userdef_bc;
  w.nangles = L.nangles;
  w.flux_r = c->ldw_flux[0]; w.flux_t = c->ldw_flux[1]; w.flux_p = c->ldw_flux[2];
  w.gline = c->ldw_dvds;
  w.mask = c->ldw_mask;
  w.UL = L.unit_length; w.UV = L.unit_velocity; w.UD = L.unit_density;
  const double KELVIN = L.unit_velocity * L.unit_velocity * amu / kB;      // pluto.h:560
  w.kelvin_mu = KELVIN * L.mu;
  w.krad = L.krad; w.alpharad = L.alpharad;
  w.alpha_m06 = L.alpharad == -0.6 && !getenv("PB200_LDW_GENERIC_POW");
  w.t_iso = c->cfg.eos == PB200_EOS_ISOTHERMAL ? L.t_iso : 0.0;
  w.mpoints = c->ldw_mpoints; w.t_fit = c->ldw_tfit; w.m_fit = c->ldw_mfit;
  w.sigma_e = sigmaT / amu / 1.18;
  w.unit_acc = L.unit_velocity * L.unit_velocity / L.unit_length;
  w.dfloor = L.dfloor / L.unit_density;
  w.tfloor = 5.e2;
  w.pfloor = w.dfloor * w.tfloor / (KELVIN * L.mu);
  w.rho_0 = L.rho_0 / L.unit_density;
  w.rho_alpha = L.rho_alpha;
  w.r_WD = c->xl[0][c->dev.beg[0]];                 // g_domBeg[IDIR]
  const double gm_cgs = Gc * L.cent_mass;
  w.gm_code = gm_cgs / (L.unit_length * L.unit_velocity * L.unit_velocity);
  double teff = pow(3.0 * gm_cgs * L.disk_mdot / (8.0 * PI * sigma), 0.25);
  teff *= pow(w.r_WD * L.unit_length, -0.75);
  w.teff_wd = teff;
}

template <int NV>
static void gen_floor_nv(pb200_ctx *c, double *V) {
  GenDev G = *c->gdev;
  G.d = c->dev;
  fill_ldw(c, G);
  GenArgs a;
  memset(&a, 0, sizeof(a));
  a.V = V; a.U = c->gU;
  const int T = 128;
  gen_ldw_floor<NV><<<(unsigned)((c->dev.sv + T - 1) / T), T, 0, c->stream>>>(G, a, c->cur_stage > 1 ? 1 : 0);
  c->launches++;
}

int pb200_gen_internal_boundary(pb200_ctx *c, double *V) {
  if (!c->gen || !c->ldw_on || !c->ldw.userdef_bc) return PB200_OK;
  int rc = pb200_gen_setup(c);
  if (rc) return rc;
  if (c->nvar - (c->cfg.eos == PB200_EOS_ISOTHERMAL ? 4 : 5) < 1) return PB200_ENOTSUP;    // the problem carries a tracer
  switch (c->nvar) {
    case 5: gen_floor_nv<5>(c, V); break;
    case 6: gen_floor_nv<6>(c, V); break;
    case 7: gen_floor_nv<7>(c, V); break;
    case 8: gen_floor_nv<8>(c, V); break;
    default: return PB200_ENOTSUP;
  }
  return PB200_OK;
}

// the last act of Boundary(): ComputeEntropy over the whole array (boundary.c:488-493)
int pb200_gen_entropy(pb200_ctx *c, double *V) {
  if (!c->gen || !c->cfg.entropy_switch) return PB200_OK;
  int rc = pb200_gen_setup(c);
  if (rc) return rc;
  GenDev G = *c->gdev;
  G.d = c->dev;
  gen_entropy<<<(unsigned)((c->dev.sv + 127) / 128), 128, 0, c->stream>>>(G, V);
  c->launches++;
  return PB200_OK;
}

int pb200_gen_userdef_side(pb200_ctx *c, double *V, int side) {
  if (!c->gen || !c->ldw_on || !c->ldw.userdef_bc) return PB200_ENOTSUP;
  if (side > 2) return PB200_ENOTSUP;               // cv_idl defines X1_BEG, X1_END, X2_BEG only
  int rc = pb200_gen_setup(c);
  if (rc) return rc;
  GenDev G = *c->gdev;
  G.d = c->dev;
  const Dev &D = c->dev;
  int ext[3] = {D.tot[0], D.tot[1], D.tot[2]};
  ext[side / 2] = D.beg[side / 2];
  long n = (long)ext[0] * ext[1] * ext[2];
  gen_ldw_side<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(G, V, side);
  c->launches++;
  return PB200_OK;
}

extern "C" int pb200_ldw_enable(pb200_ctx *c, const pb200_ldw_config *l) {
  if (!c || !l) return pb200_fail(PB200_EINVAL, "null argument");
  if (!c->gen || c->cfg.geometry != PB200_SPHERICAL || c->dev.ndim != 2)
    return pb200_fail(PB200_ENOTSUP, "LINE_DRIVEN_WIND is built for GEOMETRY SPHERICAL, DIMENSIONS 2");
  if (l->nangles < 1 || l->nangles > 64) return pb200_fail(PB200_EINVAL, "NFLUX_ANGLES must be 1..64");
  cudaSetDevice(c->cfg.device);
  c->ldw = *l;
  c->ldw_on = true;
  size_t n = (size_t)l->nangles * c->dev.sv;
  for (int q = 0; q < 3; q++) {
    if (c->ldw_flux[q]) cudaFree(c->ldw_flux[q]);
    if (cudaMalloc(&c->ldw_flux[q], n * sizeof(double)) != cudaSuccess) return pb200_fail(PB200_ENOMEM, "flux tables: out of device memory");
    if (cudaMemset(c->ldw_flux[q], 0, n * sizeof(double)) != cudaSuccess) return pb200_fail(PB200_ECUDA, "flux tables: cudaMemset failed");
  }
  if (c->ldw_dvds) cudaFree(c->ldw_dvds);
  if (cudaMalloc(&c->ldw_dvds, 2 * c->dev.sv * sizeof(double)) != cudaSuccess) return pb200_fail(PB200_ENOMEM, "line force: out of device memory");
  if (cudaMemset(c->ldw_dvds, 0, 2 * c->dev.sv * sizeof(double)) != cudaSuccess) return pb200_fail(PB200_ECUDA, "line force: cudaMemset failed");
  if (c->ldw_mask) cudaFree(c->ldw_mask);
  if (cudaMalloc(&c->ldw_mask, c->dev.sv * sizeof(unsigned long long)) != cudaSuccess) return pb200_fail(PB200_ENOMEM, "flux mask: out of device memory");
  if (cudaMemset(c->ldw_mask, 0, c->dev.sv * sizeof(unsigned long long)) != cudaSuccess) return pb200_fail(PB200_ECUDA, "flux mask: cudaMemset failed");
  return PB200_OK;
}

// force-multiplier fit of LineForce() (KRAD = ALPHARAD = 999): t_fit = log10(t) [mpoints] and
// M_UV_fit = log10(M) [mpoints][k][j][i], as read_sirocco_fluxes() leaves them (line_connect.c:185-256)
extern "C" int pb200_ldw_set_mfit(pb200_ctx *c, int mpoints, const double *t_fit, const double *m_fit) {
  if (!c || !c->ldw_on || mpoints < 1 || mpoints > 64 || !t_fit || !m_fit)
    return pb200_fail(PB200_EINVAL, "pb200_ldw_set_mfit: call pb200_ldw_enable first; 1 <= MPOINTS <= 64");
  cudaSetDevice(c->cfg.device);
  if (c->ldw_tfit) cudaFree(c->ldw_tfit);
  if (c->ldw_mfit) cudaFree(c->ldw_mfit);
  c->ldw_tfit = c->ldw_mfit = nullptr;
  size_t n = (size_t)mpoints * c->dev.sv * sizeof(double);
  if (cudaMalloc(&c->ldw_tfit, mpoints * sizeof(double)) != cudaSuccess || cudaMalloc(&c->ldw_mfit, n) != cudaSuccess)
    return pb200_fail(PB200_ENOMEM, "force-multiplier fit: out of device memory");
  // the tables are read by kernels on c->stream (a non-blocking stream): a pageable cudaMemcpy may
  // return before its last DMA has landed, so wait for the device before handing them over
  if (cudaMemcpy(c->ldw_tfit, t_fit, mpoints * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(c->ldw_mfit, m_fit, n, cudaMemcpyHostToDevice) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
    return pb200_fail(PB200_ECUDA, "force-multiplier fit: upload failed");
  c->ldw_mpoints = mpoints;
  return PB200_OK;
}

extern "C" int pb200_ldw_set_fluxes(pb200_ctx *c, const double *fr, const double *ft, const double *fp) {
  if (!c || !c->ldw_on || !fr || !ft) return pb200_fail(PB200_EINVAL, "pb200_ldw_set_fluxes: call pb200_ldw_enable first; flux_r / flux_t must not be NULL");
  cudaSetDevice(c->cfg.device);
  size_t n = (size_t)c->ldw.nangles * c->dev.sv * sizeof(double);
  if (cudaMemcpy(c->ldw_flux[0], fr, n, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(c->ldw_flux[1], ft, n, cudaMemcpyHostToDevice) != cudaSuccess ||
      (fp ? cudaMemcpy(c->ldw_flux[2], fp, n, cudaMemcpyHostToDevice) : cudaMemset(c->ldw_flux[2], 0, n)) != cudaSuccess)
    return pb200_fail(PB200_ECUDA, "flux tables: upload failed");
  // pageable copies may still be in flight on the legacy stream; the mask kernel runs on c->stream
  if (cudaDeviceSynchronize() != cudaSuccess) return pb200_fail(PB200_ECUDA, "flux tables: upload failed");
  LdwDev w;
  memset(&w, 0, sizeof(w));
  w.nangles = c->ldw.nangles;
  w.flux_r = c->ldw_flux[0]; w.flux_t = c->ldw_flux[1]; w.flux_p = fp ? c->ldw_flux[2] : nullptr;
  gen_ldw_mask<<<(unsigned)((c->dev.sv + 127) / 128), 128, 0, c->stream>>>(w, c->dev.sv, c->ldw_mask);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return pb200_fail(PB200_ECUDA, "flux mask kernel failed");
  return PB200_OK;
}

// ---- RECONSTRUCTION PARABOLIC on general grids: PPM_CoefficientsSet(), States/ppm_coeffs.c:60-290 --------------------
// The interface weights w (v_{i+1/2} = sum_j w[j] v_{i-1+j}) and the extremum coefficients h+ / h- (PPM_Q6_Coeffs,
// :520-570) are grid constants, evaluated once on the host in the reference's own operation order - same libm,
// no FMA contraction in host code - so they are the doubles the reference holds.

// B w = b by Crout LU with implicit row scaling and partial pivoting, the algorithm (and operation order) of
// Math_Tools/math_lu_decomp.c LUDecompose() + LUBackSubst(); 4 unknowns
static void ppm_solve4(double (&m)[4][4], double (&b)[4]) {
  const int n = 4;
  int piv[4];
  double scale[4];
  for (int r = 0; r < n; r++) {
    double big = 0.0;
    for (int q = 0; q < n; q++) { const double t = fabs(m[r][q]); if (t > big) big = t; }
    scale[r] = 1.0 / big;
  }
  for (int col = 0; col < n; col++) {
    for (int r = 0; r < col; r++) {
      double acc = m[r][col];
      for (int q = 0; q < r; q++) acc -= m[r][q] * m[q][col];
      m[r][col] = acc;
    }
    double best = 0.0;
    int rbest = col;
    for (int r = col; r < n; r++) {
      double acc = m[r][col];
      for (int q = 0; q < col; q++) acc -= m[r][q] * m[q][col];
      m[r][col] = acc;
      const double merit = scale[r] * fabs(acc);
      if (merit >= best) { best = merit; rbest = r; }
    }
    if (rbest != col) {
      for (int q = 0; q < n; q++) std::swap(m[rbest][q], m[col][q]);
      scale[rbest] = scale[col];
    }
    piv[col] = rbest;
    if (m[col][col] == 0.0) m[col][col] = 1.0e-20;
    if (col != n - 1) {
      const double inv = 1.0 / m[col][col];
      for (int r = col + 1; r < n; r++) m[r][col] *= inv;
    }
  }
  int first = 0;                                   // forward substitution skipping leading zeros of b
  for (int r = 0; r < n; r++) {
    const int pr = piv[r];
    double acc = b[pr];
    b[pr] = b[r];
    if (first) for (int q = first - 1; q <= r - 1; q++) acc -= m[r][q] * b[q];
    else if (acc) first = r + 1;
    b[r] = acc;
  }
  for (int r = n - 1; r >= 0; r--) {
    double acc = b[r];
    for (int q = r + 1; q < n; q++) acc -= m[r][q] * b[q];
    b[r] = acc / m[r][r];
  }
}

// int_a^b (x - x0)^k sin(x) dx by one panel of the 5-point Gauss-Legendre rule: GaussQuadrature(&BetaTheta, ., a, b, 1, 5),
// Math_Tools/math_quadrature.c:30-78,356-409
static double ppm_theta_moment(double a, double b, double x0, int k) {
  const double third = 1.0 / 3.0, c107 = 10.0 / 7.0;
  const double node[5] = {0.0, sqrt(5.0 - 2.0 * sqrt(c107)) * third, -sqrt(5.0 - 2.0 * sqrt(c107)) * third,
                          sqrt(5.0 + 2.0 * sqrt(c107)) * third, -sqrt(5.0 + 2.0 * sqrt(c107)) * third};
  const double wgt[5] = {128.0 / 225.0, (322.0 + 13.0 * sqrt(70.0)) / 900.0, (322.0 + 13.0 * sqrt(70.0)) / 900.0,
                         (322.0 - 13.0 * sqrt(70.0)) / 900.0, (322.0 - 13.0 * sqrt(70.0)) / 900.0};
  const double h = (b - a) / (double)1;
  const double lo = a + 0 * h, hi = lo + h;
  double panel = 0.0;
  for (int q = 0; q < 5; q++) {
    const double x = 0.5 * (hi - lo) * node[q] + (hi + lo) * 0.5;
    panel += wgt[q] * (pow(x - x0, k) * sin(x));
  }
  panel *= 0.5 * (hi - lo);
  double total = 0.0;
  total += panel;
  return total;
}

static inline double poly2(double a0, double a1, double a2, double x) { return a0 + x * (a1 + x * a2); }
static inline double poly4(double a0, double a1, double a2, double a3, double a4, double x) {
  return a0 + x * (a1 + x * (a2 + x * (a3 + x * a4)));
}
static inline double poly6(double a0, double a1, double a2, double a3, double a4, double a5, double a6, double x) {
  return a0 + x * (a1 + x * (a2 + x * (a3 + x * (a4 + x * (a5 + x * a6)))));
}

// weights [tot][4], h+ [tot], h- [tot] of direction d; `uniform`: grid->uniform[d] (one uniform patch, set_grid.c:67-72)
static void ppm_coefficients(int geo, int d, int nt, const double *xl, const double *xr, const double *dx, const double *xc,
                             bool uniform, std::vector<double> &w, std::vector<double> &hp, std::vector<double> &hm) {
  const bool radial = d == 0 && (geo == PB200_CYLINDRICAL || geo == PB200_POLAR);
  const bool sph_r = d == 0 && geo == PB200_SPHERICAL, sph_t = d == 1 && geo == PB200_SPHERICAL;
  w.assign((size_t)nt * 4, 0.0); hp.assign(nt, 3.0); hm.assign(nt, 3.0);
  for (int i = 0; i < nt; i++) {                   // PPM_Q6_Coeffs, ppm_coeffs.c:520-570
    if (radial) {
      hp[i] = 3.0 + 0.5 * dx[i] / xc[i];
      hm[i] = 3.0 - 0.5 * dx[i] / xc[i];
    } else if (sph_r) {
      const double r = xc[i], dr = dx[i], den = 20.0 * r * r + dr * dr;
      hp[i] = 3.0 + 2.0 * dr * (10.0 * r + dr) / den;
      hm[i] = 3.0 - 2.0 * dr * (10.0 * r - dr) / den;
    } else if (sph_t && i > 0) {
      const double cp = cos(xr[i]), sp = sin(xr[i]), cm = cos(xr[i - 1]), sm = sin(xr[i - 1]);
      const double dmu = cm - cp, dmu_t = sm - sp;
      hp[i] = dx[i] * (dmu_t + dx[i] * cp) / (dx[i] * (sp + sm) - 2.0 * dmu);
      hm[i] = -dx[i] * (dmu_t + dx[i] * cm) / (dx[i] * (sp + sm) - 2.0 * dmu);
    }
  }
  const int first = 1, last = nt - 1 - 2;          // stencil i-1 .. i+2
  if (!uniform || sph_t) {                         // PPM_FindWeights, ppm_coeffs.c:300-420 (theta: always, :268-269)
    for (int i = first; i <= last; i++) {
      double beta[4][4], rhs[4];
      const double x0 = xc[i];
      for (int j = i - 1; j <= i + 2; j++) {
        const double rp = xr[j], rm = xl[j];
        const int col = j - (i - 1);
        if (sph_t) {
          const double vol = cos(rm) - cos(rp);
          for (int k = 0; k < 4; k++) beta[k][col] = ppm_theta_moment(rm, rp, x0, k);
          for (int k = 0; k < 4; k++) beta[k][col] /= vol;
        } else if (sph_r) {
          const double vol = (rp * rp * rp - rm * rm * rm) / 3.0;
          for (int k = 0; k < 4; k++) {
            beta[k][col] = pow(rp - x0, k + 1) * ((k * k + 3.0 * k + 2.0) * rp * rp + 2.0 * x0 * (k + 1.0) * rp + 2.0 * x0 * x0)
                         - pow(rm - x0, k + 1) * ((k * k + 3.0 * k + 2.0) * rm * rm + 2.0 * x0 * (k + 1.0) * rm + 2.0 * x0 * x0);
            beta[k][col] /= (k + 3.0) * (k + 2.0) * (k + 1.0) * vol;
          }
        } else if (!radial) {
          const double vol = rp - rm;
          for (int k = 0; k < 4; k++) beta[k][col] = (pow(rp - x0, k + 1) - pow(rm - x0, k + 1)) / (k + 1.0) / vol;
        } else {
          const double vol = (rp * rp - rm * rm) / 2.0;
          for (int k = 0; k < 4; k++) {
            beta[k][col] = pow(rp - x0, k + 1) * ((k + 1.0) * rp + x0) - pow(rm - x0, k + 1) * ((k + 1.0) * rm + x0);
            beta[k][col] /= (k + 2.0) * (k + 1.0) * vol;
          }
        }
      }
      rhs[0] = 1.0;
      for (int k = 1; k < 4; k++) rhs[k] = rhs[k - 1] * (xr[i] - x0);
      ppm_solve4(beta, rhs);
      for (int j = 0; j < 4; j++) w[(size_t)i * 4 + j] = rhs[j];
    }
    return;
  }
  for (int i = first; i <= last; i++) {            // PPM_CartCoeff, ppm_coeffs.c:468-509
    double *q = &w[(size_t)i * 4];
    q[0] = -1.0 / 12.0; q[1] = 7.0 / 12.0; q[2] = 7.0 / 12.0; q[3] = -1.0 / 12.0;
    if (radial) {                                  // ppm_coeffs.c:150-165
      const double i1 = xr[i] / dx[i], i2 = i1 * i1;
      const double den = 24.0 * poly2(4.0, -15.0, 5.0, i2);
      q[0] = poly4(-12.0, -1.0, 30.0, -1.0, -10.0, i1) / den;
      q[1] = poly4(60.0, -27.0, -210.0, 13.0, 70.0, i1) / den;
      q[2] = poly4(60.0, 27.0, -210.0, -13.0, 70.0, i1) / den;
      q[3] = poly4(-12.0, 1.0, 30.0, 1.0, -10.0, i1) / den;
    }
    if (sph_r) {                                   // ppm_coeffs.c:216-229
      const double i1 = fabs(xr[i] / dx[i]), i2 = i1 * i1;
      const double den = 36.0 * poly4(16.0, -60.0, 150.0, -85.0, 15.0, i2);
      q[0] = -poly2(7, -9, 3, i1) / den * poly6(12, 16, -30, -48.0, 23, 48, 15, i1);
      q[1] = poly2(1, -3, 3, i1) / den * poly6(372, 1008.0, 510, -720, -487, 144, 105, i1);
      q[2] = poly2(1, 3, 3.0, i1) / den * poly6(372, -1008, 510, 720, -487, -144, 105, i1);
      q[3] = -poly2(7, 9, 3, i1) / den * poly6(12, -16, -30, 48, 23, -48, 15, i1);
    }
  }
}

// host-only entry (no device needed): the coefficients of one direction, for callers and tests that want to look at them
extern "C" int pb200_ppm_coefficients(int geometry, int dir, int ntot, const double *xl, const double *xr, const double *dxin,
                                      int uniform, double *w, double *hp, double *hm) {
  if (geometry < PB200_CARTESIAN || geometry > PB200_SPHERICAL || dir < 0 || dir > 2 || ntot < 4 || !xl || !xr || !w || !hp || !hm)
    return pb200_fail(PB200_EINVAL, "pb200_ppm_coefficients: bad argument");
  std::vector<double> dx(ntot), xc(ntot), vw, vp, vm;
  for (int i = 0; i < ntot; i++) { dx[i] = dxin ? dxin[i] : xr[i] - xl[i]; xc[i] = 0.5 * (xl[i] + xr[i]); }   // set_grid.c:135-137
  ppm_coefficients(geometry, dir, ntot, xl, xr, dx.data(), xc.data(), uniform != 0, vw, vp, vm);
  std::copy(vw.begin(), vw.end(), w);
  std::copy(vp.begin(), vp.end(), hp);
  std::copy(vm.begin(), vm.end(), hm);
  return PB200_OK;
}

template <int NV>
static void launch_vgrad(const GenDev &G, const GenArgs &a, const GenBox &dom, int defer, unsigned nb, cudaStream_t st) {
  gen_vgrad<NV><<<nb, 64, 0, st>>>(G, a, dom, defer);
}

// SplitSource() for COOLING BLONDIN (Src/split_source.c:53): BlondinCooling(d->Vc, d, dt, ...)
template <int NV>
static void gen_stage_nv(pb200_ctx *c, int stage, double w0, double wc, int comb) {
  GenDev G = *c->gdev;
  G.d = c->dev;     // body-force tables may have been set after gen_setup
  fill_ldw(c, G);
  const Dev &D = c->dev;
  cudaStream_t st = c->stream;
  GenArgs a;
  a.V = c->V[c->cur];
  a.U = c->gU; a.U0 = c->gU0; a.VP = c->gVP; a.VM = c->gVM; a.F = c->gF;
  a.cdt = c->gcdt; a.flag = c->gflag; a.shock = c->gshock;
  a.dt = c->d_dt; a.red = c->d_red;
  a.w0 = w0; a.wc = wc; a.comb = comb; a.stage = stage; a.dir = 0;
  a.ibmask = c->ib_n ? c->d_ibmask : nullptr;
  // PB200_GEN_FUSED=0: the one-kernel-per-reference-stage form (gen_states -> gen_riemann -> gen_rhs through
  // the VP / VM / F arrays); default: one fused kernel per direction (gen_sweep)
  static const bool fused_env = !(getenv("PB200_GEN_FUSED") && atoi(getenv("PB200_GEN_FUSED")) == 0);
  // PARABOLIC and RING_AVERAGE: gen_states carries the PPM / ring states, gen_sweep does not
  const bool fused = fused_env && !G.ppm && !G.ring_average;
  const int T = 128;
  auto blocks = [&](const GenBox &b) {
    long n = (long)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
    return (unsigned)((n + T - 1) / T);
  };
  const unsigned ball = (unsigned)((D.sv + T - 1) / T);
  auto zones_of = [](const GenBox &b) {
    return (long)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
  };
  GenBox dom;
  for (int d = 0; d < 3; d++) { dom.lo[d] = D.beg[d]; dom.hi[d] = D.end[d]; }
  if (stage == 1) {
    if (G.flatten || G.entropy) {   // FlagShock, rk_step.c:123-125 (flags were zeroed by main.c:258-261)
      if (G.flatten || G.entropy == 1) {
        GenBox b;
        for (int d = 0; d < 3; d++) { int inc = d < D.ndim; b.lo[d] = inc; b.hi[d] = D.tot[d] - 1 - inc; }
        gen_shock<<<blocks(b), T, 0, st>>>(G, a, b);
        c->launches++;
      }
      gen_flags<<<ball, T, 0, st>>>(G, a);
      c->launches++;
    }
    if (!fused) {
      gen_p2c<NV><<<blocks(dom), T, 0, st>>>(G, a, dom);
      c->launches++;
    }
  }
  c->stage_uploaded = false;
  a.cen = c->gVP;                 // the fused form does not store VP: 3 of its arrays carry the r sweep's centre data
  a.defer = (fused && c->ldw_on && D.ndim >= 2) ? 1 : 0;
  for (int dir = 0; dir < D.ndim && fused; dir++) {
    // one kernel per direction: States -> Riemann -> RightHandSide (+ PrimToCons3D / U0 in the first one)
    a.dir = dir;
    const int nx = D.end[0] - D.beg[0] + 1, ny = D.end[1] - D.beg[1] + 1, nzz = D.end[2] - D.beg[2] + 1;
    if (c->ldw_on && ((dir == 0 && !a.defer) || (dir == 1 && a.defer))) {
      // VGradCalc (update_stage.c:138-140) + the sums of LineForce(): after the r sweep, whose centre state it needs
      launch_vgrad<NV>(G, a, dom, a.defer, (unsigned)((zones_of(dom) + 63) / 64), st);
      c->launches++;
    }
    const int first = (stage == 1 && dir == 0) ? 1 : 0;
    // tile shapes (zones along the sweep x lanes across it); S - 2 zones of a tile are updated
    constexpr int S0 = PB_GEN_S0, S1 = PB_GEN_S1, L1 = PB_GEN_L1;
    const dim3 g0(ny, (nx + S0 - 3) / (S0 - 2), nzz), g1((nx + L1 - 1) / L1, (ny + S1 - 3) / (S1 - 2), nzz),
        g2((nx + L1 - 1) / L1, (nzz + S1 - 3) / (S1 - 2), ny);
    if (G.solver >= SOLVER_ROE) {       // Roe / two-shock / AUSM+: the instantiation that carries them
      if (dir == 0) gen_sweep<NV, S0, 1, true, PB_GEN_MB0><<<g0, S0, 0, st>>>(G, a, first);
      else gen_sweep<NV, S1, L1, true><<<dir == 1 ? g1 : g2, S1 * L1, 0, st>>>(G, a, first);
    } else {
      if (dir == 0) gen_sweep<NV, S0, 1, false, PB_GEN_MB0><<<g0, S0, 0, st>>>(G, a, first);
      else gen_sweep<NV, S1, L1, false><<<dir == 1 ? g1 : g2, S1 * L1, 0, st>>>(G, a, first);
    }
    c->launches++;
  }
  for (int dir = 0; dir < D.ndim && !fused; dir++) {
    a.dir = dir;
    GenBox bs = dom, bf = dom;
    bs.lo[dir] = D.beg[dir] - 1; bs.hi[dir] = D.end[dir] + 1;     // States(nbeg-1, nend+1)
    bf.lo[dir] = D.beg[dir] - 1; bf.hi[dir] = D.end[dir];         // Riemann(nbeg-1, nend)
    gen_states<NV><<<blocks(bs), T, 0, st>>>(G, a, bs);
    if (dir == 0 && c->ldw_on) {                // VGradCalc (update_stage.c:138-140) + the sums of LineForce()
      launch_vgrad<NV>(G, a, dom, 0, (unsigned)((zones_of(dom) + 63) / 64), st);
      c->launches++;
    }
    gen_riemann<NV><<<blocks(bf), T, 0, st>>>(G, a, bf);
    gen_rhs<NV><<<blocks(dom), T, 0, st>>>(G, a, dom);
    c->launches += 3;
  }
  gen_finish<NV><<<blocks(dom), T, 0, st>>>(G, a, dom);
  c->launches++;
  if (G.ring_average) {        // RingAverageCons + ConsToPrim3D of the averaged rings (rk_step.c:167-169,238-240,306-308)
    gen_ring<NV><<<blocks(dom), T, 0, st>>>(G, a, dom, 0);
    c->launches++;
  }
}

// rk_step.c:115-119: PrimToCons3D, RingAverageCons, ConsToPrim3D before the first Boundary() of a step
template <int NV>
static void gen_ring_start_nv(pb200_ctx *c) {
  GenDev G = *c->gdev;
  G.d = c->dev;
  GenArgs a;
  memset(&a, 0, sizeof(a));
  a.V = c->V[c->cur];
  a.U = c->gU; a.flag = c->gflag; a.red = c->d_red;
  GenBox dom;
  for (int d = 0; d < 3; d++) { dom.lo[d] = c->dev.beg[d]; dom.hi[d] = c->dev.end[d]; }
  const long n = (long)(dom.hi[0] - dom.lo[0] + 1) * (dom.hi[1] - dom.lo[1] + 1) * (dom.hi[2] - dom.lo[2] + 1);
  gen_ring<NV><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(G, a, dom, 1);
  c->launches++;
}
int pb200_gen_ring_start(pb200_ctx *c) {
  if (!c->gen || !c->gen_ready || c->cfg.ring_average <= 1) return PB200_OK;
  switch (c->nvar) {
    case 4: gen_ring_start_nv<4>(c); break;
    case 5: gen_ring_start_nv<5>(c); break;
    case 6: gen_ring_start_nv<6>(c); break;
    case 7: gen_ring_start_nv<7>(c); break;
    default: return PB200_ENOTSUP;
  }
  return cudaGetLastError() == cudaSuccess ? PB200_OK : pb200_fail(PB200_ECUDA, "ring average: kernel launch failed");
}

// Host-boundary mode, stages > 1: user code that changes interior zones inside Boundary() converts them
// itself (PrimToCons3D on 1-zone boxes, e.g. cv_idl/init.c:272-275) - d->Uc of every other zone, converted
// or not, stays what the previous stage left.  The caller hands over exactly the zones its Boundary() wrote.
__global__ void gen_patch_u(double *U, long sv, int nvar, long n, const long *zone, const double *u) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  for (int nv = 0; nv < nvar; nv++) U[nv * sv + zone[t]] = u[t * nvar + nv];
}
int pb200_gen_patch_u(pb200_ctx *c, long n, const long *zone, const double *u) {
  if (!c->gen || !c->gen_ready || n <= 0) return PB200_OK;
  long *dz = nullptr;
  double *du = nullptr;
  if (cudaMalloc(&dz, n * sizeof(long)) != cudaSuccess || cudaMalloc(&du, n * c->nvar * sizeof(double)) != cudaSuccess) {
    if (dz) cudaFree(dz);
    return pb200_fail(PB200_ENOMEM, "pb200_stage_patch_u: out of device memory");
  }
  cudaMemcpyAsync(dz, zone, n * sizeof(long), cudaMemcpyHostToDevice, c->stream);
  cudaMemcpyAsync(du, u, n * c->nvar * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  gen_patch_u<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->gU, c->dev.sv, c->nvar, n, dz, du);
  cudaStreamSynchronize(c->stream);
  cudaFree(dz);
  cudaFree(du);
  return cudaGetLastError() == cudaSuccess ? PB200_OK : pb200_fail(PB200_ECUDA, "pb200_stage_patch_u failed");
}

// one stage of AdvanceStep() on the general path; Boundary() fills were enqueued by the caller
int pb200_gen_stage(pb200_ctx *c, int stage) {
  double w0 = 0.0, wc = 1.0;
  int comb = 0;
  if (stage == 2) {  // rk_step.c:18-24
    comb = 1;
    if (c->nstages == 2) { w0 = 0.5; wc = 0.5; } else { w0 = 0.75; wc = 0.25; }
  } else if (stage == 3) comb = 2;
  if (c->ldw_on && c->ldw.krad == 999 && c->ldw.alpharad == 999 && c->ldw_mpoints == 0)
    return pb200_fail(PB200_EINVAL, "KRAD = ALPHARAD = 999 needs the force-multiplier fit (pb200_ldw_set_mfit)");
  switch (c->nvar) {
    case 4: gen_stage_nv<4>(c, stage, w0, wc, comb); break;
    case 5: gen_stage_nv<5>(c, stage, w0, wc, comb); break;
    case 6: gen_stage_nv<6>(c, stage, w0, wc, comb); break;
    case 7: gen_stage_nv<7>(c, stage, w0, wc, comb); break;
    case 8: gen_stage_nv<8>(c, stage, w0, wc, comb); break;
    default: return PB200_ENOTSUP;
  }
  return cudaGetLastError() == cudaSuccess ? PB200_OK : pb200_fail(PB200_ECUDA, "general-grid stage: kernel launch failed");
}
