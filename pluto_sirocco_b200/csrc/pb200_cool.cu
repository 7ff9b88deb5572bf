// pb200_cool.cu -- COOLING BLONDIN on the device: SplitSource() -> BlondinCooling() (Src/split_source.c:29,
// Src/Cooling/BLONDIN/cooling.c:50-330).
//
// This translation unit is compiled with -fmad=false: the reference build (gcc -O3 -std=c17, baseline x86-64) does
// not contract a*b+c, and the Brent iteration of BlondinCooling stops on |dT| <= 1 K, so a single differently rounded
// operation can end the iteration one step apart (a pressure difference of up to the solver's tolerance, 2e-4).  With
// the same operation order, no contraction, IEEE sqrt / division and the C library's own exp / pow / log10
// (glibc_math.cuh) the device reproduces the host's sequence of iterates and the pressure bit for bit.
#include <cuda_runtime.h>
#include <math.h>

#include "pb200_internal.h"
#include "gen_kernels.cuh"
#include "glibc_math.cuh"

namespace pb {

// ---- COOLING BLONDIN: BlondinCooling(), Src/Cooling/BLONDIN/cooling.c:50-330 ---------------
struct CoolDev {
  const double *tab[7];   // comp_h_pre, comp_c_pre, xray_h_pre, line_c_pre, brem_c_pre, sirocco_xi, sirocco_t_r (null: 1 / unused)
  double dt_share;        // dt * UNIT_TIME
  double unit_pressure, lx, tx, mu;
  int analytic_xi;        // g_time <= 3.0 (cooling.c:99-106)
};
struct CoolZone {
  double comp_c_pre, comp_h_pre, line_c_pre, brem_c_pre, xray_h_pre;
  double nH, xi, tx, sqxi, sqsqxi, n, E, hc_init, dt_share;
};
PB_D double cool_ne_rat(double T) {
  if (T < 1.5e4) return 1e-2 + pbm::pow_glibc(10.0, (-51.59417133 + 12.27740153 * pbm::log10_glibc(T)));
  else if (T < 3.3e4) return pbm::pow_glibc(10.0, (-3.80749689 + 0.86092628 * pbm::log10_glibc(T)));
  return 1.21;
}
PB_D double cool_heatcool(const CoolZone &q, double T) {
  const double sqT = sqrt(T);
  const double ne = q.nH * cool_ne_rat(T);
  const double comp_heat = q.comp_h_pre * (8.9e-36 * q.xi * q.tx);
  const double comp_cool = q.comp_c_pre * (8.9e-36 * q.xi * (4.0 * T));
  const double xray_heat = q.xray_h_pre * (1.5e-21 * (q.sqsqxi / sqT));
  const double line_cool = q.line_c_pre * ((1e-16 * pbm::exp_glibc(-1.3e5 / T) / q.sqxi / T) + fmin(fmin(1e-24, 5e-27 * sqT), 1.5e-17 / T));
  const double brem_cool = q.brem_c_pre * (3.3e-27 * sqT);
  return q.nH * (ne * comp_heat + q.nH * xray_heat - ne * comp_cool - ne * line_cool - ne * brem_cool);
}
PB_D double cool_zfunc(const CoolZone &q, double temp) {
  return (temp * q.n * 1.3806505e-16 / (2.0 / 3.0)) - q.E - q.dt_share * (q.hc_init + cool_heatcool(q, temp)) / 2.0;
}
template <bool ZF>
__device__ double cool_zbrent(const CoolZone &z, double x1, double x2, double tol) {
  auto F = [&](double x) { return ZF ? cool_zfunc(z, x) : cool_heatcool(z, x); };
  const double EPS = 3.0e-8;
  double a = x1, b = x2, c = x2, d = 0.0, e = 0.0;
  double fa = F(a), fb = F(b), fc = fb, p, q, r, s, tol1, xm;
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
      const double min1 = 3.0 * xm * q - fabs(tol1 * q), min2 = fabs(e * q);
      if (2.0 * p < (min1 < min2 ? min1 : min2)) { e = d; d = p / q; }
      else { d = xm; e = d; }
    } else { d = xm; e = d; }
    a = b; fa = fb;
    if (fabs(d) > tol1) b += d;
    else b += (xm > 0.0 ? fabs(tol1) : -fabs(tol1));
    fb = F(b);
  }
  return b;
}

static __global__ void gen_blondin(GenDev g, double *V, CoolDev cd, GenBox b) {
  int i, j, k;
  if (!gen_zone(b.lo, b.hi, i, j, k)) return;
  const Dev &d = g.d;
  const long o = (long)k * d.sk + (long)j * d.sj + i;
  CoolZone q;
  q.dt_share = cd.dt_share;
  q.comp_h_pre = cd.tab[0] ? cd.tab[0][o] : 1.0;     // defaults of read_sirocco_heatcool(), line_connect.c:383-393
  q.comp_c_pre = cd.tab[1] ? cd.tab[1][o] : 1.0;
  q.xray_h_pre = cd.tab[2] ? cd.tab[2][o] : 1.0;
  q.line_c_pre = cd.tab[3] ? cd.tab[3][o] : 1.0;
  q.brem_c_pre = cd.tab[4] ? cd.tab[4][o] : 1.0;
  const double r = __ldg(g.x[0] + i) * g.ldw.UL;
  const double rho_code = V[o], pr = V[iPRS * d.sv + o];
  const double rho = rho_code * g.ldw.UD;
  q.E = (pr * cd.unit_pressure) / (d.gas.gamma - 1);
  q.nH = rho / (1.43 * 1.67262171e-24);
  if (cd.analytic_xi || !cd.tab[5] || !cd.tab[6]) { q.xi = cd.lx / q.nH / r / r; q.tx = cd.tx; }
  else { q.xi = cd.tab[5][o]; q.tx = cd.tab[6][o]; }
  q.n = rho / (cd.mu * 1.67262171e-24);
  const double T = q.E * (2.0 / 3.0) / (q.n * 1.3806505e-16);
  if (T < 1.e4) return;                               // g_minCoolingTemp
  q.sqxi = sqrt(q.xi);
  q.sqsqxi = pbm::pow_glibc(q.xi, 0.25);
  q.hc_init = cool_heatcool(q, T);
  double t_l = T * 0.9, t_u = T * 1.1, T_f;
  double test = cool_zfunc(q, t_l) * cool_zfunc(q, t_u);
  int guard = 0;
  while (test > 0 && test == test && guard++ < 4000) {
    t_l *= 0.9; t_u *= 1.1;
    test = cool_zfunc(q, t_l) * cool_zfunc(q, t_u);
  }
  if (test != test) T_f = T;
  else {
    T_f = cool_zbrent<true>(q, t_l, t_u, 1.0);
    const double hc_final = cool_heatcool(q, T_f);
    if (hc_final * q.hc_init < 0.0) T_f = cool_zbrent<false>(q, fmin(T_f, T), fmax(T_f, T), 1.0);
  }
  T_f = fmax(T_f, 1.e4);
  const double E_f = T_f / (2.0 / 3.0) * (q.n * 1.3806505e-16);
  V[iPRS * d.sv + o] = E_f * (d.gas.gamma - 1) / cd.unit_pressure;
}


}  // namespace pb

using namespace pb;

void pb200_fill_ldw(pb200_ctx *c, pb::GenDev &G);

extern "C" int pb200_cooling_set_tables(pb200_ctx *c, const double *const tabs[7]) {
  if (!c || !c->ldw_on) return pb200_fail(PB200_EINVAL, "pb200_cooling_set_tables: call pb200_ldw_enable first");
  cudaSetDevice(c->cfg.device);
  size_t n = (size_t)c->dev.sv * sizeof(double);
  for (int q = 0; q < 7; q++) {
    if (c->cool_tab[q]) { cudaFree(c->cool_tab[q]); c->cool_tab[q] = nullptr; }
    if (!tabs || !tabs[q]) continue;
    if (cudaMalloc(&c->cool_tab[q], n) != cudaSuccess) return pb200_fail(PB200_ENOMEM, "cooling tables: out of device memory");
    if (cudaMemcpy(c->cool_tab[q], tabs[q], n, cudaMemcpyHostToDevice) != cudaSuccess) return pb200_fail(PB200_ECUDA, "cooling tables: upload failed");
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return pb200_fail(PB200_ECUDA, "cooling tables: upload failed");
  return PB200_OK;
}

extern "C" int pb200_split_source(pb200_ctx *c, double dt, double g_time) {
  if (!c || !c->gen || !c->ldw_on)
    return pb200_fail(PB200_ENOTSUP, "pb200_split_source: COOLING BLONDIN needs a line-driven-wind context (pb200_ldw_enable)");
  cudaSetDevice(c->cfg.device);
  int rc = pb200_gen_setup(c);
  if (rc) return rc;
  GenDev G = *c->gdev;
  G.d = c->dev;
  pb200_fill_ldw(c, G);
  const pb200_ldw_config &L = c->ldw;
  CoolDev cd;
  for (int q = 0; q < 7; q++) cd.tab[q] = c->cool_tab[q];
  cd.dt_share = dt * (L.unit_length / L.unit_velocity);                     // dt * UNIT_TIME
  cd.unit_pressure = L.unit_density * L.unit_velocity * L.unit_velocity;    // UNIT_PRESSURE
  cd.lx = L.lx; cd.tx = L.tx; cd.mu = L.mu;
  cd.analytic_xi = g_time <= 3.0;
  GenBox dom;
  for (int d = 0; d < 3; d++) { dom.lo[d] = c->dev.beg[d]; dom.hi[d] = c->dev.end[d]; }
  long n = (long)(dom.hi[0] - dom.lo[0] + 1) * (dom.hi[1] - dom.lo[1] + 1) * (dom.hi[2] - dom.lo[2] + 1);
  gen_blondin<<<(unsigned)((n + 63) / 64), 64, 0, c->stream>>>(G, c->V[c->cur], cd, dom);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return PB200_ECUDA;
  return PB200_OK;
}


// Measurement / test aid: the device build of glibc_math.cuh on an array (0 exp(x), 1 log(x), 2 log10(x), 3 pow(x, y)),
// host arrays in and out.  tests/test_glibc_math.py compares it with the C library bit for bit.
static __global__ void libm_probe(int which, long n, const double *x, const double *y, double *out) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  out[t] = which == 0 ? pbm::exp_glibc(x[t]) : (which == 1 ? pbm::log_glibc(x[t]) : (which == 2 ? pbm::log10_glibc(x[t]) : pbm::pow_glibc(x[t], y[t])));
}
extern "C" int pb200_libm_probe(int which, long n, const double *x, const double *y, double *out) {
  if (which < 0 || which > 3 || n < 1 || !x || !y || !out) return pb200_fail(PB200_EINVAL, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return pb200_fail(PB200_ENODEV, "no CUDA device");
  double *d = nullptr;
  if (cudaMalloc(&d, 3 * n * sizeof(double)) != cudaSuccess) return pb200_fail(PB200_ENOMEM, "out of device memory");
  cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d + n, y, n * sizeof(double), cudaMemcpyHostToDevice);
  libm_probe<<<(unsigned)((n + 127) / 128), 128>>>(which, n, d, d + n, d + 2 * n);
  cudaError_t e = cudaMemcpy(out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? PB200_OK : pb200_fail(PB200_ECUDA, cudaGetErrorString(e));
}
