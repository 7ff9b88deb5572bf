// pb200.cu -- C ABI (include/pluto_b200.h) and stage sequencing of AdvanceStep().
//
// Host-side restatement of the control flow of Src/Time_Stepping/rk_step.c:29-322:
//   stage s:  Boundary(V_s)  ->  directional sweeps (x1 [, x2 [, x3]])  ->  V_{s+1}
// with the RK weights of rk_step.c:18-24 / :304.  All state stays in HBM between calls.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "pb200_internal.h"

using namespace pb;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(PB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
  } while (0)

int pb200_fail(int code, const char *msg) { return fail(code, msg); }
extern "C" const char *pb200_last_error(void) { return g_err.c_str(); }
extern "C" int pb200_version(void) { return PB200_VERSION; }

extern "C" void pb200_config_default(pb200_config *c) {
  memset(c, 0, sizeof(*c));
  c->dimensions = 1;
  c->geometry = PB200_CARTESIAN;
  c->nx[0] = c->nx[1] = c->nx[2] = 1;
  c->nghost = 2;
  c->reconstruction = PB200_LINEAR;
  c->limiter = PB200_LIM_DEFAULT;
  c->time_stepping = PB200_RK2;
  c->solver = PB200_HLLC;
  for (int s = 0; s < 6; s++) c->bc[s] = PB200_BC_OUTFLOW;
  c->gamma = 5.0 / 3.0;
  c->small_density = 1.e-12;
  c->small_pressure = 1.e-12;
  for (int d = 0; d < 3; d++) { c->xbeg[d] = 0.0; c->xend[d] = 1.0; }
}

static int upload_grid(pb200_ctx *c, int dir) {
  int n = c->dev.tot[dir];
  std::vector<double> inv(n);
  for (int i = 0; i < n; i++) inv[i] = 1.0 / c->dx[dir][i];  // grid->inv_dx
  CK(cudaMemcpy(c->d_invdx[dir], inv.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());   // pageable copy on the legacy stream; the sweeps run on a non-blocking stream
  return PB200_OK;
}

extern "C" int pb200_create(const pb200_config *cfg, pb200_ctx **out) {
  if (!cfg || !out) return fail(PB200_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->dimensions < 1 || cfg->dimensions > 3) return fail(PB200_EINVAL, "dimensions must be 1..3");
  if (cfg->geometry < PB200_CARTESIAN || cfg->geometry > PB200_SPHERICAL) return fail(PB200_EINVAL, "bad geometry");
  if (cfg->geometry == PB200_CYLINDRICAL && cfg->dimensions == 3)
    return fail(PB200_ENOTSUP, "GEOMETRY CYLINDRICAL is 1-D / 2-D (r, z) in the reference; use POLAR for (r, phi, z)");
  // curvilinear geometry, characteristic limiting, MULTID flattening and the entropy switch run
  // on the general-grid path (pb200_gen.cu)
  const bool iso = cfg->eos == PB200_EOS_ISOTHERMAL;
  if (cfg->eos != PB200_EOS_IDEAL && !iso) return fail(PB200_EINVAL, "eos must be IDEAL or ISOTHERMAL");
  if (iso && !(cfg->iso_sound_speed > 0.0)) return fail(PB200_EINVAL, "EOS ISOTHERMAL needs iso_sound_speed > 0 (g_isoSoundSpeed)");
  if (iso && cfg->entropy_switch) return fail(PB200_EINVAL, "ENTROPY_SWITCH needs an energy equation (EOS IDEAL)");
  const int ring = cfg->ring_average > 1 ? cfg->ring_average : 0;
  const int ring_rec = ring ? (cfg->ring_average_rec ? cfg->ring_average_rec : 5) : 1;     // pluto.h:483-489
  if (ring) {     // RingAverageSize(), ring_average.c:620-735, and what RingAverageReconstruct() relies on
    const int pd = cfg->geometry == PB200_POLAR ? 1 : 2;
    if (cfg->geometry != PB200_POLAR && cfg->geometry != PB200_SPHERICAL)
      return fail(PB200_EINVAL, "RING_AVERAGE cannot be used in this geometry (ring_average.c:50)");
    if (cfg->dimensions <= pd) return fail(PB200_EINVAL, "RING_AVERAGE needs the phi direction");
    if ((ring & (ring - 1)) != 0 || cfg->nx[pd] % ring != 0)
      return fail(PB200_EINVAL, "RING_AVERAGE must be a power of two that divides the number of phi zones");
    if (cfg->nx[pd] / ring < cfg->nghost)
      return fail(PB200_ENOTSUP, "RING_AVERAGE: fewer chunks on the innermost ring than ghost zones");
    if (cfg->bc[2 * pd] != PB200_BC_PERIODIC || cfg->bc[2 * pd + 1] != PB200_BC_PERIODIC)
      return fail(PB200_EINVAL, "RING_AVERAGE needs a periodic phi direction");
    if (ring_rec != 1 && ring_rec != 2 && ring_rec != 5) return fail(PB200_EINVAL, "RING_AVERAGE_REC must be 1, 2 or 5");
    if (cfg->entropy_switch) return fail(PB200_ENOTSUP, "RING_AVERAGE with ENTROPY_SWITCH is not built");
  }
  for (int s = 0; s < 2 * cfg->dimensions; s++)
    if (cfg->bc[s] == PB200_BC_POLARAXIS) {      // Boundary(), boundary.c:344-364
      const bool ok = (cfg->geometry == PB200_POLAR && s == 0) || (cfg->geometry == PB200_SPHERICAL && (s == 2 || s == 3));
      if (!ok) return fail(PB200_EINVAL, "polaraxis: X1-beg in POLAR, an X2 boundary in SPHERICAL geometry");
      const int pd = cfg->geometry == PB200_POLAR ? 1 : 2;
      if (cfg->dimensions <= pd || cfg->nx[pd] % 2) return fail(PB200_EINVAL, "polaraxis needs an even number of phi zones");
    }
  const bool gen = cfg->geometry != PB200_CARTESIAN || cfg->char_limiting || cfg->shock_flattening ||
                   cfg->entropy_switch || iso || cfg->solver >= PB200_ROE;
  if (gen && cfg->reconstruction == PB200_FLAT)
    return fail(PB200_ENOTSUP, "the general-grid path is built for RECONSTRUCTION LINEAR and PARABOLIC");
  if (cfg->ntracer < 0 || cfg->ntracer > 2) return fail(PB200_ENOTSUP, "ntracer must be 0..2");
  if (cfg->body_force < 0 || cfg->body_force > 3) return fail(PB200_EINVAL, "bad body_force");
  if (cfg->reconstruction < PB200_FLAT || cfg->reconstruction > PB200_PARABOLIC)
    return fail(PB200_EINVAL, "bad reconstruction");
  if (cfg->solver < PB200_TVDLF || cfg->solver > PB200_AUSM) return fail(PB200_EINVAL, "bad solver");
  if (cfg->solver >= PB200_TWO_SHOCK && iso) return fail(PB200_ENOTSUP, "two_shock and ausm+ need EOS IDEAL (Src/HD/set_solver.c:25-49)");
  if (cfg->time_stepping < PB200_EULER || cfg->time_stepping > PB200_RK3)
    return fail(PB200_EINVAL, "bad time_stepping");
  int need = cfg->reconstruction == PB200_PARABOLIC ? 3 : 2;
  if (cfg->shock_flattening < 0 || cfg->shock_flattening > 2) return fail(PB200_EINVAL, "shock_flattening: 0 NO, 1 MULTID, 2 ONED");
  if (cfg->shock_flattening == 2) need = 4;         // GetNghost(), Src/get_nghost.c:42-57
  if (ring && ring_rec > 2 && need < 3) need = 3;   // get_nghost.c:40
  if (cfg->nghost < need) return fail(PB200_EINVAL, "nghost too small for the reconstruction stencil");
  for (int d = 0; d < cfg->dimensions; d++)
    if (cfg->nx[d] < cfg->nghost) return fail(PB200_EINVAL, "nx < nghost in an active dimension");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PB200_ENODEV, "no CUDA device: libplutob200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(PB200_EINVAL, "bad device ordinal");
  CK(cudaSetDevice(cfg->device));

  pb200_ctx *c = new pb200_ctx();
  c->cfg = *cfg;
  c->nvar = (iso ? 4 : 5) + cfg->ntracer + (cfg->entropy_switch ? 1 : 0);
  c->gen = gen;
  c->gen_ready = false;
  c->gdev = nullptr;
  c->ldw_on = false;
  c->cur_stage = 0;
  c->d_ibmask = nullptr;
  c->geo_set = false;
  for (int d = 0; d < 3; d++) c->grid_uniform[d] = -1;
  c->graph_exec = nullptr;
  c->graph_sig = 0;
  c->graph_launches = 0;
  c->gen_epoch = 0;
  c->capturing = false;
  // PB200_GRAPH=0|1 overrides; default: grids below 4 M zones (above that a step is milliseconds of kernels)
  c->use_graph = -1;
  if (const char *p = getenv("PB200_GRAPH")) c->use_graph = atoi(p) != 0;
  c->stage_uploaded = false;
  c->d_iblist = nullptr;
  c->ib_n = 0;
  for (int q = 0; q < 3; q++) c->ldw_flux[q] = nullptr;
  c->ldw_dvds = nullptr;
  c->ldw_mask = nullptr;
  c->ldw_mpoints = 0;
  c->ldw_tfit = c->ldw_mfit = nullptr;
  for (int q = 0; q < 7; q++) c->cool_tab[q] = nullptr;
  Dev &D = c->dev;
  D.ndim = cfg->dimensions;
  for (int d = 0; d < 3; d++) {
    bool act = d < cfg->dimensions;
    int ng = act ? cfg->nghost : 0;
    int nx = act ? cfg->nx[d] : 1;
    D.tot[d] = nx + 2 * ng;
    D.beg[d] = ng;
    D.end[d] = ng + nx - 1;
  }
  c->own_k0 = 0;
  c->own_k1 = D.end[2] - D.beg[2] + 1;
  D.sj = D.tot[0];
  D.sk = (long)D.tot[0] * D.tot[1];
  D.sv = D.sk * D.tot[2];
  c->nzone = D.sv;
  c->vbytes = (size_t)c->nvar * c->nzone * sizeof(double);
  D.gas.gamma = cfg->gamma;
  D.gas.gmm1 = cfg->gamma - 1.0;
  D.gas.inv_gmm1 = 1.0 / (cfg->gamma - 1.0);
  D.gas.small_dn = cfg->small_density;
  D.gas.small_pr = cfg->small_pressure;

  c->nstages = cfg->time_stepping == PB200_EULER ? 1 : (cfg->time_stepping == PB200_RK2 ? 2 : 3);
  int ncopies = gen ? 1 : (c->nstages == 3 ? 3 : 2);   // the general path updates d->Vc in place
  for (int k = 0; k < 3; k++) c->V[k] = nullptr;
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < ncopies && e == cudaSuccess; k++) {
    e = cudaMalloc(&c->V[k], c->vbytes);
    if (e == cudaSuccess) e = cudaMemset(c->V[k], 0, c->vbytes);
  }
  c->acc = nullptr;
  c->cdt = nullptr;
  if (e == cudaSuccess && D.ndim > 1 && !gen) {
    e = cudaMalloc(&c->acc, c->vbytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->cdt, c->nzone * sizeof(double));
  }
  if (e == cudaSuccess) e = cudaMalloc(&c->d_dt, sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_red, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost(&c->h_red, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost(&c->h_dt, sizeof(double));
  for (int d = 0; d < 3 && e == cudaSuccess; d++) e = cudaMalloc(&c->d_invdx[d], D.tot[d] * sizeof(double));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
  if (e != cudaSuccess) {
    std::string m = std::string("allocation failed: ") + cudaGetErrorString(e);
    pb200_destroy(c);
    return fail(e == cudaErrorMemoryAllocation ? PB200_ENOMEM : PB200_ECUDA, m);
  }
  // uniform grid from xbeg/xend, ghost zones continue the spacing (Src/set_grid.c)
  for (int d = 0; d < 3; d++) {
    int n = D.tot[d];
    c->xl[d].resize(n);
    c->xr[d].resize(n);
    c->dx[d].resize(n);
    int nx = D.end[d] - D.beg[d] + 1;
    double dx = (cfg->xend[d] - cfg->xbeg[d]) / nx;
    for (int i = 0; i < n; i++) {
      c->xl[d][i] = cfg->xbeg[d] + (i - D.beg[d]) * dx;
      c->xr[d][i] = cfg->xbeg[d] + (i - D.beg[d] + 1) * dx;
      c->dx[d][i] = dx;  // Src/set_grid.c:410
    }
    int rc = upload_grid(c, d);
    if (rc) { pb200_destroy(c); return rc; }
    D.inv_dx[d] = c->d_invdx[d];
  }
  D.bf_kind = cfg->body_force;
  for (int q = 0; q < 7; q++) { D.bf_tab[q] = nullptr; c->d_bf[q] = nullptr; }
  // PB200_FUSE_BC=0: materialise the x1 ghost zones with bc_fill instead of mapping them at load time.
  // (Writing the x2/x3 ghost copies from the x3 sweep's store was tried too: it removes four more
  // launches but slows the HBM-bound x3 kernel by 5 %, a net loss.)
  bool fuse_env = true;
  if (const char *p = getenv("PB200_FUSE_BC")) fuse_env = atoi(p) != 0;
  D.nghost = cfg->nghost;
  for (int s = 0; s < 6; s++) {
    int t = cfg->bc[s];
    const bool on = s < 2 && !gen && D.ndim >= 2 && fuse_env;   // x1: virtual ghosts in 2-D and 3-D
    D.bc_fuse[s] = (on && s < 2 * D.ndim && (t == PB200_BC_OUTFLOW || t == PB200_BC_REFLECTIVE || t == PB200_BC_AXISYMMETRIC ||
                                   t == PB200_BC_EQTSYMMETRIC || t == PB200_BC_PERIODIC)) ? t : 0;
  }
  c->h2d = c->d2h = nullptr;
  for (int k = 0; k < 64; k++) c->ev_up[k] = c->ev_done[k] = nullptr;
  c->host_pipeline = 16;   // measured on 512^3: 16 planes per slab gives the best overlap (1.02 vs 0.61 Gzones/s unpipelined)
  if (const char *p = getenv("PB200_HOST_PIPELINE")) c->host_pipeline = atoi(p);
  c->cur = 0;
  c->in_step = false;
  c->launches = 0;
  *out = c;
  return PB200_OK;
}

extern "C" void pb200_destroy(pb200_ctx *c) {
  if (!c) return;
  pb200_gen_release(c);
  for (int q = 0; q < 3; q++) if (c->ldw_flux[q]) cudaFree(c->ldw_flux[q]);
  if (c->ldw_dvds) cudaFree(c->ldw_dvds);
  if (c->ldw_mask) cudaFree(c->ldw_mask);
  if (c->ldw_tfit) cudaFree(c->ldw_tfit);
  if (c->ldw_mfit) cudaFree(c->ldw_mfit);
  for (int q = 0; q < 7; q++) if (c->cool_tab[q]) cudaFree(c->cool_tab[q]);
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  if (c->d_ibmask) cudaFree(c->d_ibmask);
  if (c->d_iblist) cudaFree(c->d_iblist);
  for (int k = 0; k < 3; k++) if (c->V[k]) cudaFree(c->V[k]);
  if (c->acc) cudaFree(c->acc);
  if (c->cdt) cudaFree(c->cdt);
  if (c->d_dt) cudaFree(c->d_dt);
  if (c->d_red) cudaFree(c->d_red);
  if (c->h_red) cudaFreeHost(c->h_red);
  if (c->h_dt) cudaFreeHost(c->h_dt);
  for (int d = 0; d < 3; d++) if (c->d_invdx[d]) cudaFree(c->d_invdx[d]);
  for (int q = 0; q < 7; q++) if (c->d_bf[q]) cudaFree(c->d_bf[q]);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->h2d) cudaStreamDestroy(c->h2d);
  if (c->d2h) cudaStreamDestroy(c->d2h);
  for (int k = 0; k < 64; k++) { if (c->ev_up[k]) cudaEventDestroy(c->ev_up[k]); if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]); }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (int k = 0; k < 16; k++) {
    if (c->pev0[k]) cudaEventDestroy(c->pev0[k]);
    if (c->pev1[k]) cudaEventDestroy(c->pev1[k]);
  }
  delete c;
}

extern "C" int pb200_shape(const pb200_ctx *c, int tot[3], int *nvar) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  if (tot) for (int d = 0; d < 3; d++) tot[d] = c->dev.tot[d];
  if (nvar) *nvar = c->nvar;
  return PB200_OK;
}

// grid->uniform[d] of the reference is an input-file property (one uniform patch); when the caller has not passed it,
// a direction counts as uniform if its dx is constant to round-off
bool pb200_grid_is_uniform(const pb200_ctx *c, int dir) {
  if (c->grid_uniform[dir] >= 0) return c->grid_uniform[dir] != 0;
  const std::vector<double> &dx = c->dx[dir];
  for (size_t i = 1; i < dx.size(); i++)
    if (fabs(dx[i] - dx[0]) > 1e-12 * fabs(dx[0])) return false;
  return true;
}

// RECONSTRUCTION PARABOLIC: the marching kernels carry the uniform-grid weights of PPM_CartCoeff() (ppm_coeffs.c:468-509);
// a non-uniform direction needs the grid-dependent weights (ppm_coeffs.c:124-136), which the general path holds
static void ppm_route(pb200_ctx *c) {
  if (c->cfg.reconstruction != PB200_PARABOLIC || c->gen) return;
  for (int d = 0; d < c->dev.ndim; d++)
    if (!pb200_grid_is_uniform(c, d)) c->gen = true;
  if (c->gen)                                    // the general path reads real x1 ghost zones (no virtual ghosts)
    for (int s = 0; s < 6; s++) c->dev.bc_fuse[s] = 0;
}

extern "C" int pb200_set_grid(pb200_ctx *c, int dir, const double *xl, const double *xr, const double *dx) {
  if (!c || dir < 0 || dir > 2 || !xl || !xr) return fail(PB200_EINVAL, "bad argument");
  int n = c->dev.tot[dir];
  for (int i = 0; i < n; i++) {
    if (!(xr[i] > xl[i])) return fail(PB200_EINVAL, "grid: xr <= xl");
    c->xl[dir][i] = xl[i];
    c->xr[dir][i] = xr[i];
    c->dx[dir][i] = dx ? dx[i] : xr[i] - xl[i];
  }
  CK(cudaSetDevice(c->cfg.device));
  c->gen_ready = false;
  ppm_route(c);
  return upload_grid(c, dir);
}

extern "C" int pb200_set_grid_uniform(pb200_ctx *c, const int uniform[3]) {
  if (!c || !uniform) return fail(PB200_EINVAL, "null argument");
  for (int d = 0; d < 3; d++) c->grid_uniform[d] = uniform[d] ? 1 : 0;
  c->gen_ready = false;
  ppm_route(c);
  return PB200_OK;
}

extern "C" int pb200_set_geometry(pb200_ctx *c, const pb200_geometry *g) {
  if (!c || !g || !g->dV || !g->rt || !g->s || !g->sp) return fail(PB200_EINVAL, "null argument");
  const Dev &D = c->dev;
  const size_t n1 = D.tot[0], n2 = D.tot[1], n3 = D.tot[2];
  c->geo_dV.assign(g->dV, g->dV + n1 * n2 * n3);
  const size_t na[3] = {(n1 + 1) * n2 * n3, n1 * (n2 + 1) * n3, n1 * n2 * (n3 + 1)};
  for (int d = 0; d < 3; d++) {
    if (!g->A[d] || !g->dx_dl[d]) return fail(PB200_EINVAL, "null argument");
    c->geo_A[d].assign(g->A[d], g->A[d] + na[d]);
    c->geo_dxdl[d].assign(g->dx_dl[d], g->dx_dl[d] + n1 * n2);
  }
  c->geo_rt.assign(g->rt, g->rt + n1);
  c->geo_s.assign(g->s, g->s + n2);
  c->geo_sp.assign(g->sp, g->sp + n2);
  c->geo_set = true;
  c->gen_ready = false;
  return PB200_OK;
}

static int set_bf_table(pb200_ctx *c, int q, const double *tab, long n, long si, long sj, long sk) {
  if (!c || !tab || n < 1) return fail(PB200_EINVAL, "bad argument");
  const Dev &D = c->dev;
  long last = (long)(D.tot[0] - 1) * si + (long)(D.tot[1] - 1) * sj + (long)(D.tot[2] - 1) * sk;
  if (si < 0 || sj < 0 || sk < 0 || last >= n) return fail(PB200_EINVAL, "body-force table too small for its strides");
  CK(cudaSetDevice(c->cfg.device));
  if (c->d_bf[q]) { cudaFree(c->d_bf[q]); c->d_bf[q] = nullptr; }
  CK(cudaMalloc(&c->d_bf[q], n * sizeof(double)));
  CK(cudaMemcpy(c->d_bf[q], tab, n * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());   // pageable copy on the legacy stream; the sweeps run on a non-blocking stream
  c->dev.bf_tab[q] = c->d_bf[q];
  c->dev.bf_st[q][0] = si; c->dev.bf_st[q][1] = sj; c->dev.bf_st[q][2] = sk;
  return PB200_OK;
}
extern "C" int pb200_set_body_force_vector(pb200_ctx *c, int comp, const double *tab, long n, long si,
                                           long sj, long sk) {
  if (!c || comp < 0 || comp > 2) return fail(PB200_EINVAL, "bad component");
  if (!(c->cfg.body_force & PB200_BF_VECTOR)) return fail(PB200_EINVAL, "cfg.body_force has no VECTOR part");
  return set_bf_table(c, comp, tab, n, si, sj, sk);
}
extern "C" int pb200_set_body_force_potential(pb200_ctx *c, int where, const double *tab, long n, long si,
                                              long sj, long sk) {
  if (!c || where < 0 || where > 3) return fail(PB200_EINVAL, "bad table selector");
  if (!(c->cfg.body_force & PB200_BF_POTENTIAL)) return fail(PB200_EINVAL, "cfg.body_force has no POTENTIAL part");
  return set_bf_table(c, 3 + where, tab, n, si, sj, sk);
}

extern "C" int pb200_upload_vc(pb200_ctx *c, const double *h) {
  if (!c || !h) return fail(PB200_EINVAL, "null argument");
  CK(cudaSetDevice(c->cfg.device));
  CK(cudaMemcpyAsync(c->V[c->cur], h, c->vbytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PB200_OK;
}
extern "C" int pb200_download_vc(pb200_ctx *c, double *h) {
  if (!c || !h) return fail(PB200_EINVAL, "null argument");
  CK(cudaSetDevice(c->cfg.device));
  CK(cudaMemcpyAsync(h, c->V[c->cur], c->vbytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PB200_OK;
}
extern "C" double *pb200_device_vc(pb200_ctx *c) { return c ? c->V[c->cur] : nullptr; }
extern "C" void *pb200_stream(pb200_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int pb200_nstages(const pb200_ctx *c) { return c ? c->nstages : 0; }

// ---- Boundary() ------------------------------------------------------------------------
// sides: bit mask of the sides to fill; [k0, k1): absolute x3 plane range for the x1 / x2 sides
static int boundary_on(pb200_ctx *c, double *V, unsigned sides = 0x3f, int k0 = 0, int k1 = -1) {
  const Dev &D = c->dev;
  if (k1 < 0) k1 = D.tot[2];
  // INTERNAL_BOUNDARY YES: UserDefBoundary(d, NULL, 0, grid) comes first (boundary.c:126-128)
  int rc0 = pb200_gen_internal_boundary(c, V);
  if (rc0) return fail(rc0, "internal boundary (UserDefBoundary side 0) failed");
  for (int side = 0; side < 2 * D.ndim; side++) {
    int type = c->cfg.bc[side];
    if (type == PB200_BC_USERDEF) {
      // device versions exist for the line-driven-wind problem only; other user code must fill
      // its ghost zones itself between pb200_stage_boundary() and pb200_stage_begin()
      if (c->ldw_on && c->ldw.userdef_bc) {
        int rc = pb200_gen_userdef_side(c, V, side);
        if (rc) return fail(rc, "no device UserDefBoundary for this side");
      }
      continue;
    }
    if (type == PB200_BC_NEIGHBOUR || type == 0 || !(sides & (1u << side))) continue;
    if (side < 2 && D.bc_fuse[side]) continue;    // x1 ghosts are virtual in the fused x1+x2 kernel
    BcArgs b;
    b.k0 = k0; b.k1 = k1;
    b.V = V;
    b.side = side;
    b.type = type;
    b.nvar = c->nvar;
    b.nghost = c->cfg.nghost;
    for (int nv = 0; nv < 16; nv++) b.sign[nv] = 1.0;
    b.sign[1 + side / 2] = -1.0;  // FlipSign(): normal velocity (Src/boundary.c:503)
    if (type == PB200_BC_AXISYMMETRIC && c->cfg.geometry != PB200_CARTESIAN)
      b.sign[c->cfg.geometry == PB200_POLAR ? 2 : 3] = -1.0;  // iVPHI: VX2 (POLAR), VX3 otherwise (boundary.c:548, pluto.h)
    b.pdir = c->cfg.geometry == PB200_POLAR ? 1 : 2;
    if (type == PB200_BC_POLARAXIS) b.sign[1 + b.pdir] = -1.0;  // PolarAxisBoundary(): v_normal and v_phi change sign
    int ext[3] = {D.tot[0], D.tot[1], D.tot[2]};
    ext[side / 2] = b.nghost;
    if (side < 4) ext[2] = k1 - k0;
    long n = (long)ext[0] * ext[1] * ext[2];
    int nb = (int)((n + 255) / 256);
    bc_fill<<<nb, 256, 0, c->stream>>>(D, b);
    c->launches++;
  }
  int rce = pb200_gen_entropy(c, V);
  if (rce) return fail(rce, "ComputeEntropy failed");
  CK(cudaGetLastError());
  return PB200_OK;
}

extern "C" int pb200_boundary(pb200_ctx *c) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  CK(cudaSetDevice(c->cfg.device));
  int rc = boundary_on(c, c->V[c->cur]);
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->stream));
  return PB200_OK;
}

// ---- sweeps: one launcher per (NVAR, body force), see pb200_sweeps.cu -----------------------
static void launch_sweep(pb200_ctx *c, int dir, const SweepArgs &a) {
  static const pb200_sweep_fn tab[3][2] = {
      {pb200_launch_sweep_nv5_bf0, pb200_launch_sweep_nv5_bf1},
      {pb200_launch_sweep_nv6_bf0, pb200_launch_sweep_nv6_bf1},
      {pb200_launch_sweep_nv7_bf0, pb200_launch_sweep_nv7_bf1}};
  tab[c->nvar - 5][c->dev.bf_kind ? 1 : 0](c, dir, a);
}

// NaN screen of the array about to be swept is folded into the reduction cell by a tiny
// kernel over the dt-reduction result instead of a full pass: a NaN anywhere propagates to
// cmax of its faces, and fmax() drops NaNs, so test invDt/maxMach via the c2p counter.
// g_dt comes from a pinned host cell (device-visible under UVA), not from a kernel argument, so that a
// captured graph of the step can be replayed with a new dt
__global__ void reset_red(unsigned long long *red, double *dt, const double *dt_host) {
  red[0] = 0ull;
  red[1] = 0ull;
  red[2] = 0ull;
  red[3] = 0ull;
  *dt = *dt_host;
}

// ---- FLAG_INTERNAL_BOUNDARY --------------------------------------------------------------
// InternalBoundaryReset() (Src/int_bound_reset.c:17-40, called at the end of RightHandSide(),
// Src/MHD/rhs.c:416-417) zeroes the right-hand side of flagged zones in every sweep (fluxes AND
// sources; with the default INTERNAL_BOUNDARY_CFL YES the signal speeds stay), i.e. a flagged zone
// leaves a stage as  prim(w0 U0 + wc cons(V_in))  while its neighbours still see its face fluxes.
// Fast path: the sweeps run unchanged and ib_fix rewrites the flagged zones afterwards (they are
// few; no cost for the hot kernels when there are none).  General path: gen_rhs zeroes its rhs.
// The frozen states are computed BEFORE the last sweep of the stage (which may overwrite V^n in place:
// stage 2 of RK2 writes into the array that holds V^n) and scattered after it.
template <int NV>
__global__ void ib_gather(Dev d, const long *list, long n, const double *Vin, const double *V0, double *buf,
                          int comb, double w0, double wc) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long o = list[t];
  double v[NV], v0[NV], U[NV], vn[NV];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) { v[nv] = Vin[nv * d.sv + o]; v0[nv] = V0[nv * d.sv + o]; }
  prim2cons<NV>(v, U, d.gas);
  int nfail = 0, nan = 0;
  combine_c2p<NV>(U, v0, d.gas, comb, w0, wc, vn, nfail, nan, true);
#pragma unroll
  for (int nv = 0; nv < NV; nv++) buf[nv * n + t] = vn[nv];
}
__global__ void ib_scatter(Dev d, const long *list, long n, const double *buf, double *Vout, int nvar) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long o = list[t];
  for (int nv = 0; nv < nvar; nv++) Vout[nv * d.sv + o] = buf[nv * n + t];
}

extern "C" int pb200_set_internal_boundary_mask(pb200_ctx *c, const unsigned char *mask) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  CK(cudaSetDevice(c->cfg.device));
  if (c->d_iblist) { cudaFree(c->d_iblist); c->d_iblist = nullptr; }
  c->ib_n = 0;
  if (!mask) {
    if (c->d_ibmask) { cudaFree(c->d_ibmask); c->d_ibmask = nullptr; }
    return PB200_OK;
  }
  const Dev &D = c->dev;
  std::vector<long> list;
  for (int k = D.beg[2]; k <= D.end[2]; k++)
    for (int j = D.beg[1]; j <= D.end[1]; j++)
      for (int i = D.beg[0]; i <= D.end[0]; i++) {
        const long o = (long)k * D.sk + (long)j * D.sj + i;
        if (mask[o]) list.push_back(o);
      }
  if (!c->d_ibmask) CK(cudaMalloc(&c->d_ibmask, c->nzone));
  CK(cudaMemcpyAsync(c->d_ibmask, mask, c->nzone, cudaMemcpyHostToDevice, c->stream));
  if (!list.empty()) {
    CK(cudaMalloc(&c->d_iblist, list.size() * (sizeof(long) + c->nvar * sizeof(double))));   // list + the gather buffer
    CK(cudaMemcpyAsync(c->d_iblist, list.data(), list.size() * sizeof(long), cudaMemcpyHostToDevice, c->stream));
    c->ib_n = (long)list.size();
  }
  CK(cudaStreamSynchronize(c->stream));
  return PB200_OK;
}

static void internal_boundary_fix(pb200_ctx *c, const SweepArgs &a, bool after) {
  if (!c->ib_n || c->gen) return;
  const int nb = (int)((c->ib_n + 127) / 128);
  double *buf = (double *)(c->d_iblist + c->ib_n);
  if (after) ib_scatter<<<nb, 128, 0, c->stream>>>(c->dev, c->d_iblist, c->ib_n, buf, a.Vout, c->nvar);
  else switch (c->nvar) {
    case 5: ib_gather<5><<<nb, 128, 0, c->stream>>>(c->dev, c->d_iblist, c->ib_n, a.V, a.V0, buf, a.comb, a.w0, a.wc); break;
    case 6: ib_gather<6><<<nb, 128, 0, c->stream>>>(c->dev, c->d_iblist, c->ib_n, a.V, a.V0, buf, a.comb, a.w0, a.wc); break;
    default: ib_gather<7><<<nb, 128, 0, c->stream>>>(c->dev, c->d_iblist, c->ib_n, a.V, a.V0, buf, a.comb, a.w0, a.wc); break;
  }
  c->launches++;
}

extern "C" int pb200_step_begin(pb200_ctx *c, double dt) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  if (!(dt > 0.0)) return fail(PB200_EINVAL, "dt must be > 0");
  for (int q = 0; q < 7; q++) {   // every table BODY_FORCE needs has been handed over
    bool need = (q < 3) ? (c->cfg.body_force & PB200_BF_VECTOR) != 0
                        : ((c->cfg.body_force & PB200_BF_POTENTIAL) != 0 && (q == 3 || q - 4 < c->dev.ndim));
    if (need && !c->dev.bf_tab[q]) return fail(PB200_EINVAL, "BODY_FORCE table not set (pb200_set_body_force_*)");
  }
  CK(cudaSetDevice(c->cfg.device));
  c->launches = 0;
  c->nprof = 0;
  if (c->gen) {
    int rc = pb200_gen_setup(c);
    if (rc) return fail(rc, "general-grid set-up failed (out of device memory?)");
  }
  if (!c->capturing) {
    CK(cudaStreamSynchronize(c->stream));    // h_dt of the previous step has been consumed
    CK(cudaEventRecord(c->ev0, c->stream));
  }
  *c->h_dt = dt;
  reset_red<<<1, 1, 0, c->stream>>>(c->d_red, c->d_dt, c->h_dt);
  c->launches++;
  // array rotation: stage s sweeps stage_in[s] and writes stage_out[s]
  int A = c->cur, B = (c->cur + 1) % (c->nstages == 3 ? 3 : 2), C = (c->cur + 2) % 3;
  if (c->nstages == 1) { c->stage_in[1] = A; c->stage_out[1] = B; }
  else if (c->nstages == 2) { c->stage_in[1] = A; c->stage_out[1] = B; c->stage_in[2] = B; c->stage_out[2] = A; }
  else { c->stage_in[1] = A; c->stage_out[1] = B; c->stage_in[2] = B; c->stage_out[2] = C;
         c->stage_in[3] = C; c->stage_out[3] = A; }
  if (c->gen) for (int s = 1; s <= 3; s++) c->stage_in[s] = c->stage_out[s] = c->cur;
  c->in_step = true;
  return PB200_OK;
}

extern "C" double *pb200_stage_array(pb200_ctx *c, int stage) {
  if (!c || stage < 1 || stage > c->nstages) return nullptr;
  return c->V[c->stage_in[stage]];
}

extern "C" int pb200_stage_download(pb200_ctx *c, int stage, double *h) {
  if (!c || !h || !c->in_step || stage < 1 || stage > c->nstages) return fail(PB200_EINVAL, "bad argument");
  CK(cudaSetDevice(c->cfg.device));
  CK(cudaMemcpyAsync(h, c->V[c->stage_in[stage]], c->vbytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PB200_OK;
}
extern "C" int pb200_stage_upload(pb200_ctx *c, int stage, const double *h) {
  if (!c || !h || !c->in_step || stage < 1 || stage > c->nstages) return fail(PB200_EINVAL, "bad argument");
  CK(cudaSetDevice(c->cfg.device));
  CK(cudaMemcpyAsync(c->V[c->stage_in[stage]], h, c->vbytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->stage_uploaded = true;
  return PB200_OK;
}

extern "C" int pb200_set_owned_planes(pb200_ctx *c, int k0, int k1) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  const int nk = c->dev.end[2] - c->dev.beg[2] + 1;
  if (c->gen || c->dev.ndim != 3) return fail(PB200_ENOTSUP, "owned planes: 3-D Cartesian path");
  if (k0 < 0 || k1 > nk || k0 >= k1) return fail(PB200_EINVAL, "owned planes: need 0 <= k0 < k1 <= NX3");
  c->own_k0 = k0;
  c->own_k1 = k1;
  return PB200_OK;
}

extern "C" int pb200_stage_patch_u(pb200_ctx *c, long n, const long *zone, const double *u) {
  if (!c || !c->in_step || n < 0 || (n > 0 && (!zone || !u))) return fail(PB200_EINVAL, "bad argument");
  CK(cudaSetDevice(c->cfg.device));
  return pb200_gen_patch_u(c, n, zone, u);     // the Cartesian path keeps no d->Uc: cons(V) is recomputed
}

// A stage in three parts so that a slab-decomposed caller can overlap the halo exchange of the
// outermost direction with the sweeps that do not read those ghost planes:
//   pb200_stage_boundary: Boundary() fills of the physical sides;
//   pb200_stage_begin : (3-D only) the fused x1+x2 kernel, which touches interior planes only;
//   pb200_stage_finish: the sweeps that read the ghosts of the outermost active direction.
static int stage_args(pb200_ctx *c, int stage, SweepArgs &a) {
  if (!c || !c->in_step) return fail(PB200_EINVAL, "pb200_stage outside begin/end");
  if (stage < 1 || stage > c->nstages) return fail(PB200_EINVAL, "bad stage");
  a.V = c->V[c->stage_in[stage]];
  a.V0 = c->V[c->cur];
  a.acc = c->acc;
  a.Vout = c->V[c->stage_out[stage]];
  a.cdt = c->cdt;
  a.dt = c->d_dt;
  a.red = c->d_red;
  a.stage = stage;
  a.limiter = c->cfg.limiter;
  a.i0 = 0;
  a.k0 = 0;
  a.k1 = c->dev.end[2] - c->dev.beg[2] + 1;
  a.ko0 = c->own_k0;
  a.ko1 = c->own_k1;
  a.comb = 0; a.w0 = 0.0; a.wc = 1.0;
  if (stage == 2) {  // rk_step.c:18-24
    a.comb = 1;
    if (c->nstages == 2) { a.w0 = 0.5; a.wc = 0.5; } else { a.w0 = 0.75; a.wc = 0.25; }
  } else if (stage == 3) {
    a.comb = 2;
  }
  return PB200_OK;
}

extern "C" int pb200_stage_boundary(pb200_ctx *c, int stage) {
  SweepArgs a;
  int rc = stage_args(c, stage, a);
  if (rc) return rc;
  if (stage == 1) {                               // RING_AVERAGE: rk_step.c:115-119, ahead of the first Boundary()
    rc = pb200_gen_ring_start(c);
    if (rc) return rc;
  }
  c->cur_stage = stage;
  rc = boundary_on(c, c->V[c->stage_in[stage]]);  // Boundary(d, 0, grid), rk_step.c:121,213,285
  c->cur_stage = 0;
  return rc;
}

extern "C" int pb200_stage_begin(pb200_ctx *c, int stage) {
  SweepArgs a;
  int rc = stage_args(c, stage, a);
  if (rc) return rc;
  if (c->gen) return PB200_OK;
  if (c->dev.ndim == 3) launch_sweep(c, 1, a);    // x1 + x2 in one kernel: interior planes only
  CK(cudaGetLastError());
  return PB200_OK;
}

extern "C" int pb200_stage_finish(pb200_ctx *c, int stage) {
  SweepArgs a;
  int rc = stage_args(c, stage, a);
  if (rc) return rc;
  if (c->gen) {
    rc = pb200_gen_stage(c, stage);
    return rc;      // pb200_gen_stage set the error text
  }
  const Dev &D = c->dev;
  internal_boundary_fix(c, a, false);             // InternalBoundaryReset(): flagged zones keep w0 U0 + wc U ...
  if (D.ndim == 1) launch_sweep(c, 0, a);
  else if (D.ndim == 2) launch_sweep(c, 1, a);    // x1 + x2 (reads the x2 ghosts)
  else launch_sweep(c, 2, a);                     // x3
  internal_boundary_fix(c, a, true);              // ... written over what the sweeps left in those zones
  CK(cudaGetLastError());
  return PB200_OK;
}

extern "C" int pb200_stage(pb200_ctx *c, int stage) {
  int rc = pb200_stage_boundary(c, stage);
  if (rc) return rc;
  rc = pb200_stage_begin(c, stage);
  if (rc) return rc;
  return pb200_stage_finish(c, stage);
}

static int finish_info(pb200_ctx *c, pb200_step_info *info);
extern "C" int pb200_step_end(pb200_ctx *c, pb200_step_info *info) {
  if (!c || !c->in_step) return fail(PB200_EINVAL, "pb200_step_end outside a step");
  c->in_step = false;
  if (c->nstages == 1) c->cur = c->stage_out[1];  // EULER: result lives in the other array
  CK(cudaMemcpyAsync(c->h_red, c->d_red, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  return finish_info(c, info);
}

static unsigned long long graph_signature(const pb200_ctx *c) {
  unsigned long long h = 1469598103934665603ull;
  auto mix = [&h](const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t k = 0; k < n; k++) { h ^= b[k]; h *= 1099511628211ull; }
  };
  mix(&c->dev, sizeof(c->dev));
  mix(&c->cfg, sizeof(c->cfg));
  mix(&c->cur, sizeof(c->cur));
  mix(&c->own_k0, sizeof(int));
  mix(&c->own_k1, sizeof(int));
  mix(&c->ldw_on, sizeof(c->ldw_on));
  mix(&c->ldw, sizeof(c->ldw));
  mix(c->ldw_flux, sizeof(c->ldw_flux));
  mix(&c->ldw_dvds, sizeof(void *));
  mix(&c->ldw_mask, sizeof(void *));
  mix(&c->ldw_mpoints, sizeof(int));
  mix(&c->ldw_tfit, sizeof(void *));
  mix(&c->ldw_mfit, sizeof(void *));
  mix(&c->d_ibmask, sizeof(void *));
  mix(&c->d_iblist, sizeof(void *));
  mix(&c->ib_n, sizeof(long));
  mix(&c->gdev, sizeof(void *));
  mix(&c->gen_epoch, sizeof(int));
  mix(&c->gen_ready, sizeof(bool));
  mix(c->V, sizeof(c->V));
  return h;
}

static int finish_info(pb200_ctx *c, pb200_step_info *info) {
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  double invdt, mach;
  memcpy(&invdt, &c->h_red[0], 8);
  memcpy(&mach, &c->h_red[1], 8);
  if (c->dev.ndim > 1) invdt /= (double)c->dev.ndim;  // update_stage.c:392
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  for (int k = 0; k < c->nprof; k++) cudaEventElapsedTime(&c->pms[k], c->pev0[k], c->pev1[k]);
  if (info) {
    info->invDt_hyp = invdt;
    info->maxMach = mach;
    info->c2p_failures = c->h_red[2];
    info->gpu_ms = ms;
    info->launches = c->launches;
  }
  if (c->h_red[3] != 0ull || !(invdt > 0.0) || !isfinite(invdt))
    return fail(PB200_ENAN, "non-finite or zero signal speed: NaN in the state (CheckNaN)");
  return PB200_OK;
}

extern "C" int pb200_set_profiling(pb200_ctx *c, int on) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  c->profiling = on != 0;
  return PB200_OK;
}

extern "C" int pb200_kernel_times(const pb200_ctx *c, int max, float *ms, int *dir, int *stage) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  int n = c->nprof < max ? c->nprof : max;
  for (int k = 0; k < n; k++) {
    if (ms) ms[k] = c->pms[k];
    if (dir) dir[k] = c->pdir[k];
    if (stage) stage[k] = c->pstage[k];
  }
  return n;
}

// everything that ends up in a kernel argument of a step: if it changes, the captured graph is stale
static unsigned long long graph_signature(const pb200_ctx *c);

static int finish_info(pb200_ctx *c, pb200_step_info *info);

static int advance_step_graph(pb200_ctx *c, double dt, pb200_step_info *info) {
  if (!(dt > 0.0)) return fail(PB200_EINVAL, "dt must be > 0");
  CK(cudaSetDevice(c->cfg.device));
  if (c->gen) {
    int rc = pb200_gen_setup(c);      // allocations and uploads happen here, outside the capture
    if (rc) return fail(rc, "general-grid set-up failed (out of device memory?)");
  }
  const unsigned long long sig = graph_signature(c);
  if (!c->graph_exec || sig != c->graph_sig) {
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    int rc = pb200_step_begin(c, dt);
    for (int s = 1; s <= c->nstages && !rc; s++) rc = pb200_stage(c, s);
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMemcpyAsync(c->h_red, c->d_red, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(c->stream, &graph);
    c->capturing = false;
    c->in_step = false;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || e2 != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return fail(PB200_ECUDA, "CUDA graph capture of the step failed");
    }
    e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { c->graph_exec = nullptr; return fail(PB200_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    c->graph_sig = sig;
    c->graph_launches = c->launches;
  }
  *c->h_dt = dt;     // read by the first node of the graph
  CK(cudaEventRecord(c->ev0, c->stream));
  CK(cudaGraphLaunch(c->graph_exec, c->stream));
  c->launches = c->graph_launches;
  c->nprof = 0;
  if (c->nstages == 1) c->cur = c->stage_out[1];
  return finish_info(c, info);
}

extern "C" int pb200_advance_step(pb200_ctx *c, double dt, pb200_step_info *info) {
  if (!c) return fail(PB200_EINVAL, "null ctx");
  const bool small = c->nzone < 4L * 1000 * 1000;
  if ((c->use_graph == 1 || (c->use_graph < 0 && small)) && !c->profiling && c->nstages > 1 && !c->in_step) {
    for (int q = 0; q < 7; q++) {   // the table check of pb200_step_begin
      bool need = (q < 3) ? (c->cfg.body_force & PB200_BF_VECTOR) != 0
                          : ((c->cfg.body_force & PB200_BF_POTENTIAL) != 0 && (q == 3 || q - 4 < c->dev.ndim));
      if (need && !c->dev.bf_tab[q]) return fail(PB200_EINVAL, "BODY_FORCE table not set (pb200_set_body_force_*)");
    }
    return advance_step_graph(c, dt, info);
  }
  int rc = pb200_step_begin(c, dt);
  if (rc) return rc;
  for (int s = 1; s <= c->nstages; s++) {
    rc = pb200_stage(c, s);
    if (rc) { c->in_step = false; return rc; }
  }
  return pb200_step_end(c, info);
}

// AdvanceStep on a HOST d->Vc as a slab-wise pipeline (3-D, fast path, any RK order): the x3 planes
// travel up in slabs, stage 1 runs on slab s as soon as slab s+1 has arrived, stage 2 on slab s-1 as
// soon as stage 1 of slab s is done (stage 3 on slab s-2), and every finished slab travels back while the next ones are still
// being computed - upload, compute and download overlap, so the call costs about one PCIe
// direction instead of two plus the compute.  Same kernels, same arithmetic as pb200_advance_step.
static int advance_step_host_pipelined_body(pb200_ctx *c, double *h, double dt, pb200_step_info *info);
static int advance_step_host_pipelined(pb200_ctx *c, double *h, double dt, pb200_step_info *info) {
  int rc = advance_step_host_pipelined_body(c, h, dt, info);
  if (rc) {
    // asynchronous copies may still touch the caller's buffer: drain all three streams first
    const std::string msg = g_err;
    if (c->h2d) cudaStreamSynchronize(c->h2d);
    if (c->d2h) cudaStreamSynchronize(c->d2h);
    cudaStreamSynchronize(c->stream);
    c->in_step = false;
    g_err = msg;
  }
  return rc;
}
static int advance_step_host_pipelined_body(pb200_ctx *c, double *h, double dt, pb200_step_info *info) {
  const Dev &D = c->dev;
  const int nk = D.end[2] - D.beg[2] + 1, ng = c->cfg.nghost;
  int per = c->host_pipeline;
  if (nk / per > 64) per = (nk + 63) / 64;
  const int S = nk / per;            // the last slab takes the remainder (never thinner than `per`)
  auto rel0 = [&](int q) { return q * per; };
  auto rel1 = [&](int q) { return q == S - 1 ? nk : (q + 1) * per; };
  if (!c->h2d) {
    CK(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    for (int k = 0; k < 64; k++) {
      CK(cudaEventCreateWithFlags(&c->ev_up[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    }
  }
  int rc = pb200_step_begin(c, dt);
  if (rc) return rc;
  double *A = c->V[c->stage_in[1]];
  const size_t plane = (size_t)D.sk;
  auto kbeg = [&](int s) { return s == 0 ? 0 : D.beg[2] + rel0(s); };                // absolute planes of slab s,
  auto kend = [&](int s) { return s == S - 1 ? D.tot[2] : D.beg[2] + rel1(s); };     // ghosts ride with the end slabs
  for (int s = 0; s < S; s++) {   // uploads
    for (int nv = 0; nv < c->nvar; nv++) {
      size_t o = (size_t)nv * D.sv + (size_t)kbeg(s) * plane;
      CK(cudaMemcpyAsync(A + o, h + o, (size_t)(kend(s) - kbeg(s)) * plane * sizeof(double), cudaMemcpyHostToDevice, c->h2d));
    }
    CK(cudaEventRecord(c->ev_up[s], c->h2d));
  }
  const int NS = c->nstages;
  SweepArgs sa[4];
  for (int st = 1; st <= NS; st++) {
    rc = stage_args(c, st, sa[st]);
    if (rc) { c->in_step = false; return rc; }
  }
  double *R = c->V[c->stage_out[NS]];                                           // the step's result array
  auto run_slab = [&](int stage, int q) -> int {
    SweepArgs a = sa[stage];
    double *Vin = c->V[c->stage_in[stage]];
    const int k0 = rel0(q), k1 = rel1(q);                                       // relative to KBEG
    unsigned sides = 0x0f;                                                      // x1, x2 sides of these planes
    int r = boundary_on(c, Vin, sides, D.beg[2] + k0, D.beg[2] + k1);
    if (r) return r;
    if (q == 0) { r = boundary_on(c, Vin, 1u << 4); if (r) return r; }           // x3-beg ghost planes
    if (q == S - 1) { r = boundary_on(c, Vin, 1u << 5); if (r) return r; }       // x3-end ghost planes
    a.k0 = k0; a.k1 = k1;
    launch_sweep(c, 1, a);
    launch_sweep(c, 2, a);
    if (stage < NS) return PB200_OK;
    CK(cudaEventRecord(c->ev_done[q], c->stream));                              // last stage: slab q is final
    CK(cudaStreamWaitEvent(c->d2h, c->ev_done[q], 0));
    const int d0 = k0 > c->own_k0 ? k0 : c->own_k0, d1 = k1 < c->own_k1 ? k1 : c->own_k1;   // owned planes only
    for (int nv = 0; nv < c->nvar && d1 > d0; nv++) {
      size_t o = (size_t)nv * D.sv + (size_t)(D.beg[2] + d0) * plane;
      CK(cudaMemcpyAsync(h + o, R + o, (size_t)(d1 - d0) * plane * sizeof(double), cudaMemcpyDeviceToHost, c->d2h));
    }
    return PB200_OK;
  };
  // stage st works on slab s-(st-1): its x3 sweep reads the planes next to the slab, which the
  // stage before it produced one slab ahead earlier in the same iteration; a stage's output slab
  // is never one that a running earlier stage still reads (slabs are >= 2*nghost planes thick).
  for (int s = 0; s < S + NS - 1; s++) {
    if (s < S) CK(cudaStreamWaitEvent(c->stream, c->ev_up[s + 1 < S ? s + 1 : S - 1], 0));
    for (int st = 1; st <= NS; st++) {
      const int q = s - (st - 1);
      if (q < 0 || q >= S) continue;
      rc = run_slab(st, q);
      if (rc) { c->in_step = false; return rc; }
    }
  }
  CK(cudaGetLastError());
  rc = pb200_step_end(c, info);
  CK(cudaStreamSynchronize(c->d2h));
  (void)ng;
  return rc;
}

extern "C" int pb200_advance_step_host(pb200_ctx *c, double *vc_host, double dt, pb200_step_info *info) {
  if (!c || !vc_host) return fail(PB200_EINVAL, "null argument");
  CK(cudaSetDevice(c->cfg.device));
  {
    const Dev &D = c->dev;
    const int nk = D.end[2] - D.beg[2] + 1;
    const bool x3_ok = c->cfg.bc[4] != PB200_BC_PERIODIC && c->cfg.bc[5] != PB200_BC_PERIODIC &&
                       c->cfg.bc[4] != PB200_BC_NEIGHBOUR && c->cfg.bc[5] != PB200_BC_NEIGHBOUR &&
                       c->cfg.bc[4] != PB200_BC_USERDEF && c->cfg.bc[5] != PB200_BC_USERDEF;
    if (!c->gen && D.ndim == 3 && c->host_pipeline >= 2 * c->cfg.nghost &&
        nk >= 3 * c->host_pipeline && x3_ok && !c->profiling && !c->ib_n)
      return advance_step_host_pipelined(c, vc_host, dt, info);
  }
  CK(cudaMemcpyAsync(c->V[c->cur], vc_host, c->vbytes, cudaMemcpyHostToDevice, c->stream));
  int rc = pb200_advance_step(c, dt, info);
  if (rc) return rc;
  CK(cudaMemcpyAsync(vc_host, c->V[c->cur], c->vbytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PB200_OK;
}

// page-lock / unlock a host buffer (d->Vc of the reference is malloc'ed, Src/arrays.c:251) so that
// the copies of pb200_advance_step_host() run asynchronously at full PCIe speed
extern "C" int pb200_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(PB200_EINVAL, "null argument");
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return PB200_OK;
}
extern "C" int pb200_host_unregister(void *ptr) {
  if (!ptr) return fail(PB200_EINVAL, "null argument");
  CK(cudaHostUnregister(ptr));
  return PB200_OK;
}

extern "C" double pb200_next_time_step(double invDt_hyp, double cfl, double cfl_max_var, double g_dt,
                                       double first_dt) {
  // Src/main.c:601-697 with COOLING NO, PARABOLIC_FLUX NO, no particles
  double dt_hyp = 1.0 / invDt_hyp;
  dt_hyp *= cfl;
  double dtnext = dt_hyp;
  double lim = cfl_max_var * g_dt;
  dtnext = dtnext <= lim ? dtnext : lim;
  if (dtnext < first_dt * 1.e-9) return -1.0;
  if (cfl_max_var == 1.0) return g_dt;
  return dtnext;
}

extern "C" int pb200_integrate(pb200_ctx *c, int nsteps, double tstop, double cfl, double cfl_max_var,
                               double first_dt, double *t, double *dt, pb200_step_info *last) {
  if (!c || !t || !dt) return fail(PB200_EINVAL, "null argument");
  int done = 0;
  pb200_step_info info;
  memset(&info, 0, sizeof(info));
  for (int n = 0; n < nsteps; n++) {
    bool last_step = false;
    if ((*t + *dt) >= tstop * (1.0 - 1.e-8)) {  // Src/main.c:227-230
      *dt = tstop - *t;
      last_step = true;
    }
    int rc = pb200_advance_step(c, *dt, &info);
    if (rc) return rc;
    *t += *dt;
    double nd = pb200_next_time_step(info.invDt_hyp, cfl, cfl_max_var, *dt, first_dt);
    if (nd < 0.0) return fail(PB200_EINVAL, "NextTimeStep(): dt is too small");
    *dt = nd;
    done++;
    if (last_step) break;
  }
  if (last) *last = info;
  return done;
}

extern "C" int pb200_halo_layout(const pb200_ctx *c, int dir, long *lo_ghost, long *lo_edge,
                                 long *hi_edge, long *hi_ghost, long *count, long *var_stride) {
  if (!c || dir < 0 || dir >= c->dev.ndim) return fail(PB200_EINVAL, "bad argument");
  if (dir != c->dev.ndim - 1)
    return fail(PB200_ENOTSUP, "slab split is along the outermost active direction only (contiguous planes)");
  const Dev &D = c->dev;
  long plane = (dir == 2) ? D.sk : (dir == 1 ? D.sj : 1);
  int ng = c->cfg.nghost;
  if (lo_ghost) *lo_ghost = 0;
  if (lo_edge) *lo_edge = (long)D.beg[dir] * plane;
  if (hi_edge) *hi_edge = (long)(D.end[dir] - ng + 1) * plane;
  if (hi_ghost) *hi_ghost = (long)(D.end[dir] + 1) * plane;
  if (count) *count = (long)ng * plane;
  if (var_stride) *var_stride = D.sv;
  return PB200_OK;
}
