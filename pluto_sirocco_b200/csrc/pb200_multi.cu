// pb200_multi.cu -- slab decomposition across the GPUs of one box, driven from ONE host thread.
//
// This is the C-host replacement of the reference's parallel layer for the hot path:
//   Src/Parallel/al_decompose.c:40,125-158   MPI Cartesian decomposition
//        -> 1-D slab split along the OUTERMOST active direction (x3 in 3-D, x2 in 2-D): the ghost
//           planes of that direction are contiguous per variable in Vc[nv][k][j][i];
//   Src/boundary.c:139-158 + Src/Parallel/al_exchange_dim.c:64-90   per-variable MPI_Sendrecv pairs
//        -> per stage ONE grouped batch of ncclSend / ncclRecv of packed edge planes on a
//           communication stream per device, overlapped with the fused x1+x2 kernel (which reads
//           interior planes only);
//   Src/main.c:288,547   MPI_Allreduce(MAX) of g_maxMach / invDt_hyp
//        -> one ncclAllReduce(ncclMax) per step on the IEEE bit patterns (non-negative doubles order
//           like unsigned integers, so the result is exact and order independent).
// The reference's driver is single threaded (SURVEY 8b "Threading"), so the design is one host
// thread, N devices: a pb200_ctx per device (same kernels, same arithmetic as the single-GPU path,
// interior faces typed PB200_BC_NEIGHBOUR) and a single-process NCCL communicator
// (ncclCommInitAll).  NCCL is resolved at run time (dlopen of libnccl.so.2) so that the library
// neither needs NCCL for single-GPU use nor clashes with a copy the process already loaded.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "pb200_internal.h"

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;

const char *load_nccl() {
  if (g_nccl.h) return nullptr;
  const char *names[] = {getenv("PB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) return "libnccl.so.2 not found (set PB200_NCCL_LIB)";
#define PB_SYM(field, name)                                         \
  *(void **)(&g_nccl.field) = dlsym(h, name);                        \
  if (!g_nccl.field) { dlclose(h); return "NCCL symbol missing: " name; }
  PB_SYM(CommInitAll, "ncclCommInitAll")
  PB_SYM(CommDestroy, "ncclCommDestroy")
  PB_SYM(GroupStart, "ncclGroupStart")
  PB_SYM(GroupEnd, "ncclGroupEnd")
  PB_SYM(Send, "ncclSend")
  PB_SYM(Recv, "ncclRecv")
  PB_SYM(AllReduce, "ncclAllReduce")
  PB_SYM(GetErrorString, "ncclGetErrorString")
  PB_SYM(GetVersion, "ncclGetVersion")
#undef PB_SYM
  g_nccl.h = h;
  return nullptr;
}

}  // namespace

struct pb200_multi {
  int n = 0;
  int sdir = 0, ng = 0, ndim = 0, nvar = 0;
  pb200_config gcfg;
  bool periodic = false;
  std::vector<pb200_ctx *> ctx;
  std::vector<int> dev, count, offset;
  std::vector<ncclComm_t> comm;
  std::vector<cudaStream_t> cstream;            // communication stream per device
  std::vector<cudaEvent_t> ev_ready, ev_done;
  std::vector<double *> sbuf_lo, sbuf_hi, rbuf_lo, rbuf_hi;   // packed edge / ghost planes (all variables)
  long plane = 0;                               // doubles per plane of the split direction
  int gtot[3] = {1, 1, 1};                      // global NX*_TOT
  bool use_nccl = true;
  bool threads = true;      // one worker thread per device enqueues that device's share of a step (NCCL mode)
  int exchanges = 0;
};

#define MCK(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return pb200_fail(PB200_ECUDA, (std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); \
  } while (0)
#define NCK(call)                                                                              \
  do {                                                                                         \
    ncclResult_t r_ = (call);                                                                  \
    if (r_ != ncclSuccess)                                                                     \
      return pb200_fail(PB200_ECUDA, (std::string(#call) + ": " + g_nccl.GetErrorString(r_)).c_str()); \
  } while (0)

static int lo_of(const pb200_multi *m, int r) { return r > 0 ? r - 1 : (m->periodic ? m->n - 1 : -1); }
static int hi_of(const pb200_multi *m, int r) { return r < m->n - 1 ? r + 1 : (m->periodic ? 0 : -1); }

extern "C" void pb200_multi_destroy(pb200_multi *m) {
  if (!m) return;
  for (int r = 0; r < (int)m->ctx.size(); r++) {
    if (r < (int)m->dev.size()) cudaSetDevice(m->dev[r]);
    if (r < (int)m->comm.size() && m->comm[r]) g_nccl.CommDestroy(m->comm[r]);
    if (r < (int)m->cstream.size() && m->cstream[r]) cudaStreamDestroy(m->cstream[r]);
    if (r < (int)m->ev_ready.size() && m->ev_ready[r]) cudaEventDestroy(m->ev_ready[r]);
    if (r < (int)m->ev_done.size() && m->ev_done[r]) cudaEventDestroy(m->ev_done[r]);
    for (auto *v : {&m->sbuf_lo, &m->sbuf_hi, &m->rbuf_lo, &m->rbuf_hi})
      if (r < (int)v->size() && (*v)[r]) cudaFree((*v)[r]);
    if (m->ctx[r]) pb200_destroy(m->ctx[r]);
  }
  delete m;
}

extern "C" int pb200_multi_create(const pb200_config *gcfg, int ngpus, const int *devices, pb200_multi **out) {
  if (!gcfg || !out || ngpus < 1) return pb200_fail(PB200_EINVAL, "bad argument");
  *out = nullptr;
  if (gcfg->dimensions < 2 && ngpus > 1)
    return pb200_fail(PB200_ENOTSUP, "1-D grids are not decomposed (replicas only)");
  const bool gen = gcfg->geometry != PB200_CARTESIAN || gcfg->char_limiting || gcfg->shock_flattening || gcfg->entropy_switch ||
                   gcfg->eos != PB200_EOS_IDEAL || gcfg->solver >= PB200_ROE || gcfg->ring_average > 1;    // as pb200_create routes
  if (gen && ngpus > 1) return pb200_fail(PB200_ENOTSUP, "the general-grid path runs on one GPU (replicas only)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return pb200_fail(PB200_ENODEV, "no CUDA device: libplutob200 has no CPU fallback");
  // ranks may share a device (devices = {0, 0, ...}): the slab logic is then exercised on a single GPU,
  // with the edge planes moved by device copies instead of NCCL (which refuses duplicate devices)
  bool dup = false;
  for (int r = 0; r < ngpus; r++) {
    const int dr = devices ? devices[r] : r;
    if (dr < 0 || dr >= ndev) return pb200_fail(PB200_ENODEV, "more GPUs requested than visible");
    for (int q = 0; q < r; q++) dup = dup || (devices ? devices[q] : q) == dr;
  }
  const int sdir = gcfg->dimensions - 1, ng = gcfg->nghost;
  if (gcfg->nx[sdir] / ngpus < 2 * ng && ngpus > 1)
    return pb200_fail(PB200_EINVAL, "slabs would be thinner than 2*nghost planes");

  pb200_multi *m = new pb200_multi();
  m->n = ngpus; m->sdir = sdir; m->ng = ng; m->ndim = gcfg->dimensions; m->gcfg = *gcfg;
  m->periodic = gcfg->bc[2 * sdir] == PB200_BC_PERIODIC;
  // PB200_MULTI_EXCHANGE=copy: cudaMemcpyPeerAsync between the packed buffers instead of ncclSend/ncclRecv
  m->use_nccl = !dup && !(getenv("PB200_MULTI_EXCHANGE") && !strcmp(getenv("PB200_MULTI_EXCHANGE"), "copy"));
  // The caller stays single threaded (the reference's driver is); inside a step the library fans the launch work of
  // the N devices out to N short-lived worker threads, so that short steps (PPM + RK3 at 256^3: 4.8 ms of GPU time,
  // ~40 launches and as many stream / event calls per device) are not limited by one thread's enqueue rate
  // (measured at N = 2: 4.78 ms per step against 5.67 ms).  PB200_MULTI_THREADS=0: one thread.
  m->threads = m->use_nccl && !(getenv("PB200_MULTI_THREADS") && atoi(getenv("PB200_MULTI_THREADS")) == 0);
  m->ctx.assign(ngpus, nullptr);
  m->comm.assign(ngpus, nullptr);
  m->cstream.assign(ngpus, nullptr);
  m->ev_ready.assign(ngpus, nullptr);
  m->ev_done.assign(ngpus, nullptr);
  for (auto *v : {&m->sbuf_lo, &m->sbuf_hi, &m->rbuf_lo, &m->rbuf_hi}) v->assign(ngpus, nullptr);
  for (int d = 0; d < 3; d++) m->gtot[d] = d < gcfg->dimensions ? gcfg->nx[d] + 2 * ng : 1;
  const int base = gcfg->nx[sdir] / ngpus, rem = gcfg->nx[sdir] % ngpus;
  const double gdx = (gcfg->xend[sdir] - gcfg->xbeg[sdir]) / gcfg->nx[sdir];
  int off = 0;
  for (int r = 0; r < ngpus; r++) {
    m->dev.push_back(devices ? devices[r] : r);
    m->count.push_back(base + (r < rem ? 1 : 0));
    m->offset.push_back(off);
    off += m->count[r];
  }
  for (int r = 0; r < ngpus; r++) {
    pb200_config c = *gcfg;
    c.device = m->dev[r];
    c.nx[sdir] = m->count[r];
    c.xbeg[sdir] = gcfg->xbeg[sdir] + m->offset[r] * gdx;
    c.xend[sdir] = gcfg->xbeg[sdir] + (m->offset[r] + m->count[r]) * gdx;
    if (ngpus > 1) {
      if (lo_of(m, r) >= 0) c.bc[2 * sdir] = PB200_BC_NEIGHBOUR;
      if (hi_of(m, r) >= 0) c.bc[2 * sdir + 1] = PB200_BC_NEIGHBOUR;
    }
    int rc = pb200_create(&c, &m->ctx[r]);
    if (rc) { pb200_multi_destroy(m); return rc; }
  }
  m->nvar = m->ctx[0]->nvar;
  m->plane = sdir == 2 ? m->ctx[0]->dev.sk : m->ctx[0]->dev.sj;
  // every block uses the spacing of the UNDECOMPOSED grid (Src/set_grid.c:410), not (xend-xbeg)/n of
  // its own extent: otherwise the results are not bit-identical with a single-GPU run
  {
    const int ntot = m->gtot[sdir];
    std::vector<double> xl(ntot), xr(ntot), dx(ntot);
    for (int i = 0; i < ntot; i++) {
      xl[i] = gcfg->xbeg[sdir] + (i - ng) * gdx;
      xr[i] = gcfg->xbeg[sdir] + (i - ng + 1) * gdx;
      dx[i] = gdx;
    }
    for (int r = 0; r < ngpus; r++) {
      int rc = pb200_set_grid(m->ctx[r], sdir, xl.data() + m->offset[r], xr.data() + m->offset[r], dx.data() + m->offset[r]);
      if (rc) { pb200_multi_destroy(m); return rc; }
    }
  }
  if (ngpus > 1) {
    if (m->use_nccl) {
      if (const char *err = load_nccl()) { pb200_multi_destroy(m); return pb200_fail(PB200_ENOTSUP, err); }
      ncclResult_t nr = g_nccl.CommInitAll(m->comm.data(), ngpus, m->dev.data());
      if (nr != ncclSuccess) {
        std::string msg = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(nr);
        for (auto &c : m->comm) c = nullptr;
        pb200_multi_destroy(m);
        return pb200_fail(PB200_ECUDA, msg.c_str());
      }
    } else {
      for (int r = 0; r < ngpus; r++)
        for (int q = 0; q < ngpus; q++) {
          int can = 0;
          if (m->dev[r] != m->dev[q] && cudaDeviceCanAccessPeer(&can, m->dev[r], m->dev[q]) == cudaSuccess && can) {
            cudaSetDevice(m->dev[r]);
            cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[q], 0);
            if (e != cudaSuccess) cudaGetLastError();   // already enabled
          }
        }
    }
    const size_t bytes = (size_t)m->nvar * ng * m->plane * sizeof(double);
    for (int r = 0; r < ngpus; r++) {
      cudaError_t e = cudaSetDevice(m->dev[r]);
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->cstream[r], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_ready[r], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_done[r], cudaEventDisableTiming);
      for (auto *v : {&m->sbuf_lo, &m->sbuf_hi, &m->rbuf_lo, &m->rbuf_hi})
        if (e == cudaSuccess) e = cudaMalloc(&(*v)[r], bytes);
      if (e != cudaSuccess) {
        std::string msg = std::string("multi-GPU set-up: ") + cudaGetErrorString(e);
        pb200_multi_destroy(m);
        return pb200_fail(PB200_ECUDA, msg.c_str());
      }
    }
  }
  *out = m;
  return PB200_OK;
}

extern "C" int pb200_multi_ngpus(const pb200_multi *m) { return m ? m->n : 0; }
extern "C" pb200_ctx *pb200_multi_ctx(pb200_multi *m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }
extern "C" int pb200_multi_slab(const pb200_multi *m, int rank, int *offset, int *count) {
  if (!m || rank < 0 || rank >= m->n) return pb200_fail(PB200_EINVAL, "bad rank");
  if (offset) *offset = m->offset[rank];
  if (count) *count = m->count[rank];
  return PB200_OK;
}

// grid->xl / xr / dx of the GLOBAL grid (np_tot entries of direction dir)
extern "C" int pb200_multi_set_grid(pb200_multi *m, int dir, const double *xl, const double *xr, const double *dx) {
  if (!m || dir < 0 || dir > 2 || !xl || !xr) return pb200_fail(PB200_EINVAL, "bad argument");
  for (int r = 0; r < m->n; r++) {
    const int o = dir == m->sdir ? m->offset[r] : 0;
    int rc = pb200_set_grid(m->ctx[r], dir, xl + o, xr + o, dx ? dx + o : nullptr);
    if (rc) return rc;
    if (m->ctx[r]->gen && m->n > 1)
      return pb200_fail(PB200_ENOTSUP, "RECONSTRUCTION PARABOLIC on a non-uniform grid runs on the general path: one GPU");
  }
  return PB200_OK;
}

// BODY_FORCE tables with GLOBAL indices: value(i,j,k) = tab[i*si + j*sj + k*sk]; each rank gets the
// part of the table its planes address
static int multi_bf(pb200_multi *m, bool vec, int sel, const double *tab, long n, long si, long sj, long sk) {
  if (!m || !tab) return pb200_fail(PB200_EINVAL, "bad argument");
  const long st[3] = {si, sj, sk};
  for (int r = 0; r < m->n; r++) {
    const long shift = (long)m->offset[r] * st[m->sdir];
    int rc = vec ? pb200_set_body_force_vector(m->ctx[r], sel, tab + shift, n - shift, si, sj, sk)
                 : pb200_set_body_force_potential(m->ctx[r], sel, tab + shift, n - shift, si, sj, sk);
    if (rc) return rc;
  }
  return PB200_OK;
}
extern "C" int pb200_multi_set_body_force_vector(pb200_multi *m, int comp, const double *tab, long n, long si, long sj, long sk) {
  return multi_bf(m, true, comp, tab, n, si, sj, sk);
}
extern "C" int pb200_multi_set_body_force_potential(pb200_multi *m, int where, const double *tab, long n, long si, long sj, long sk) {
  return multi_bf(m, false, where, tab, n, si, sj, sk);
}

// d->Vc of the GLOBAL grid [nvar][NX3_TOT][NX2_TOT][NX1_TOT] <-> the slabs.  Each slab is uploaded
// with its ghost planes (the neighbours' edge planes of the host copy); only interior planes come back.
static int multi_copy(pb200_multi *m, double *h, bool up, bool sync) {
  const long gsv = (long)m->gtot[0] * m->gtot[1] * m->gtot[2];
  for (int r = 0; r < m->n; r++) {
    pb200_ctx *c = m->ctx[r];
    MCK(cudaSetDevice(m->dev[r]));
    double *V = c->V[c->cur];
    const long lsv = c->dev.sv;
    for (int nv = 0; nv < m->nvar; nv++) {
      if (up) {
        const long npl = c->dev.tot[m->sdir];
        MCK(cudaMemcpyAsync(V + nv * lsv, h + nv * gsv + (long)m->offset[r] * m->plane, (size_t)npl * m->plane * sizeof(double),
                            cudaMemcpyHostToDevice, c->stream));
      } else {
        const long o = (long)m->ng * m->plane;
        MCK(cudaMemcpyAsync(h + nv * gsv + (long)m->offset[r] * m->plane + o, V + nv * lsv + o,
                            (size_t)m->count[r] * m->plane * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      }
    }
  }
  if (sync)
    for (int r = 0; r < m->n; r++) {
      MCK(cudaSetDevice(m->dev[r]));
      MCK(cudaStreamSynchronize(m->ctx[r]->stream));
    }
  return PB200_OK;
}
extern "C" int pb200_multi_upload_vc(pb200_multi *m, const double *vc_host) {
  if (!m || !vc_host) return pb200_fail(PB200_EINVAL, "null argument");
  return multi_copy(m, const_cast<double *>(vc_host), true, true);
}
extern "C" int pb200_multi_download_vc(pb200_multi *m, double *vc_host) {
  if (!m || !vc_host) return pb200_fail(PB200_EINVAL, "null argument");
  return multi_copy(m, vc_host, false, true);
}

// pack / unpack of the edge planes of all variables: one strided device copy per face
static cudaError_t pack(double *buf, const double *V, long first_plane, const pb200_multi *m, const pb200_ctx *c, bool unpack) {
  const size_t width = (size_t)m->ng * m->plane * sizeof(double);
  const double *src = V + first_plane * m->plane;
  if (!unpack) return cudaMemcpy2DAsync(buf, width, src, (size_t)c->dev.sv * sizeof(double), width, m->nvar, cudaMemcpyDeviceToDevice, c->stream);
  return cudaMemcpy2DAsync(const_cast<double *>(src), (size_t)c->dev.sv * sizeof(double), buf, width, width, m->nvar, cudaMemcpyDeviceToDevice, c->stream);
}

// the whole step of ONE rank (NCCL mode), run by that rank's worker thread: every NCCL call is a per-thread group on
// the rank's own communicator, matched by the neighbours' threads
static int rank_step(pb200_multi *m, int r, double dt, pb200_step_info *ir, std::string *err) {
  pb200_ctx *c = m->ctx[r];
  const int ng = m->ng, lo = lo_of(m, r), hi = hi_of(m, r);
  const size_t cnt = (size_t)m->nvar * ng * m->plane;
  auto body = [&]() -> int {
    MCK(cudaSetDevice(m->dev[r]));
    int rc = pb200_step_begin(c, dt);
    if (rc) return rc;
    for (int s = 1; s <= c->nstages; s++) {
      rc = pb200_stage_boundary(c, s);
      if (rc) return rc;
      double *V = pb200_stage_array(c, s);
      const int npl = c->dev.tot[m->sdir];
      if (hi >= 0) MCK(pack(m->sbuf_hi[r], V, npl - 2 * ng, m, c, false));
      if (lo >= 0) MCK(pack(m->sbuf_lo[r], V, ng, m, c, false));
      MCK(cudaEventRecord(m->ev_ready[r], c->stream));
      MCK(cudaStreamWaitEvent(m->cstream[r], m->ev_ready[r], 0));
      NCK(g_nccl.GroupStart());
      if (hi >= 0) NCK(g_nccl.Send(m->sbuf_hi[r], cnt, ncclDouble, hi, m->comm[r], m->cstream[r]));
      if (lo >= 0) NCK(g_nccl.Recv(m->rbuf_lo[r], cnt, ncclDouble, lo, m->comm[r], m->cstream[r]));
      if (lo >= 0) NCK(g_nccl.Send(m->sbuf_lo[r], cnt, ncclDouble, lo, m->comm[r], m->cstream[r]));
      if (hi >= 0) NCK(g_nccl.Recv(m->rbuf_hi[r], cnt, ncclDouble, hi, m->comm[r], m->cstream[r]));
      NCK(g_nccl.GroupEnd());
      MCK(cudaEventRecord(m->ev_done[r], m->cstream[r]));
      rc = pb200_stage_begin(c, s);
      if (rc) return rc;
      MCK(cudaStreamWaitEvent(c->stream, m->ev_done[r], 0));
      if (lo >= 0) MCK(pack(m->rbuf_lo[r], V, 0, m, c, true));
      if (hi >= 0) MCK(pack(m->rbuf_hi[r], V, npl - ng, m, c, true));
      rc = pb200_stage_finish(c, s);
      if (rc) return rc;
    }
    NCK(g_nccl.AllReduce(c->d_red, c->d_red, 2, ncclUint64, ncclMax, m->comm[r], c->stream));
    return pb200_step_end(c, ir);
  };
  const int rc = body();
  if (rc) { *err = pb200_last_error(); c->in_step = false; }
  return rc;
}

static int advance_step_threaded(pb200_multi *m, double dt, pb200_step_info *info) {
  const int n = m->n;
  std::vector<pb200_step_info> ir(n);
  std::vector<std::string> err(n);
  std::vector<int> rc(n, 0);
  std::vector<std::thread> th;
  th.reserve(n);
  for (int r = 1; r < n; r++) th.emplace_back([&, r] { rc[r] = rank_step(m, r, dt, &ir[r], &err[r]); });
  rc[0] = rank_step(m, 0, dt, &ir[0], &err[0]);
  for (auto &t : th) t.join();
  m->exchanges += m->ctx[0]->nstages;
  pb200_step_info tot;
  memset(&tot, 0, sizeof(tot));
  for (int r = 0; r < n; r++) {
    if (rc[r]) return pb200_fail(rc[r], err[r].c_str());
    tot.invDt_hyp = ir[r].invDt_hyp > tot.invDt_hyp ? ir[r].invDt_hyp : tot.invDt_hyp;
    tot.maxMach = ir[r].maxMach > tot.maxMach ? ir[r].maxMach : tot.maxMach;
    tot.c2p_failures += ir[r].c2p_failures;
    tot.gpu_ms = ir[r].gpu_ms > tot.gpu_ms ? ir[r].gpu_ms : tot.gpu_ms;
    tot.launches += ir[r].launches;
  }
  if (info) *info = tot;
  return PB200_OK;
}

extern "C" int pb200_multi_advance_step(pb200_multi *m, double dt, pb200_step_info *info) {
  if (!m) return pb200_fail(PB200_EINVAL, "null ctx");
  if (m->n > 1 && m->threads) return advance_step_threaded(m, dt, info);
  const int n = m->n, ng = m->ng;
  int rc = PB200_OK;
  for (int r = 0; r < n && !rc; r++) rc = pb200_step_begin(m->ctx[r], dt);
  const int ns = m->ctx[0]->nstages;
  const size_t cnt = (size_t)m->nvar * ng * m->plane;
  for (int s = 1; s <= ns && !rc; s++) {
    // physical boundaries, then pack the edge planes of the array this stage sweeps
    for (int r = 0; r < n && !rc; r++) {
      pb200_ctx *c = m->ctx[r];
      MCK(cudaSetDevice(m->dev[r]));
      rc = pb200_stage_boundary(c, s);
      if (rc || n == 1) continue;
      const double *V = pb200_stage_array(c, s);
      const int npl = c->dev.tot[m->sdir];
      if (hi_of(m, r) >= 0) MCK(pack(m->sbuf_hi[r], V, npl - 2 * ng, m, c, false));
      if (lo_of(m, r) >= 0) MCK(pack(m->sbuf_lo[r], V, ng, m, c, false));
      MCK(cudaEventRecord(m->ev_ready[r], c->stream));
      MCK(cudaStreamWaitEvent(m->cstream[r], m->ev_ready[r], 0));
    }
    if (rc) break;
    if (n > 1 && !m->use_nccl) {
      // device copies: rank r writes its edge planes straight into the neighbours' receive buffers,
      // once both sides have packed (= have unpacked the previous stage's planes)
      const size_t bytes = cnt * sizeof(double);
      for (int r = 0; r < n; r++) {
        const int lo = lo_of(m, r), hi = hi_of(m, r);
        MCK(cudaSetDevice(m->dev[r]));
        if (hi >= 0) {
          MCK(cudaStreamWaitEvent(m->cstream[r], m->ev_ready[hi], 0));
          MCK(cudaMemcpyPeerAsync(m->rbuf_lo[hi], m->dev[hi], m->sbuf_hi[r], m->dev[r], bytes, m->cstream[r]));
        }
        if (lo >= 0) {
          MCK(cudaStreamWaitEvent(m->cstream[r], m->ev_ready[lo], 0));
          MCK(cudaMemcpyPeerAsync(m->rbuf_hi[lo], m->dev[lo], m->sbuf_lo[r], m->dev[r], bytes, m->cstream[r]));
        }
        MCK(cudaEventRecord(m->ev_done[r], m->cstream[r]));
      }
      m->exchanges++;
    } else if (n > 1) {
      // one grouped batch: "upward" traffic first (hi edge -> the upper neighbour's lo ghosts), then
      // "downward", so that two ranks that are each other's lo AND hi neighbour (periodic, n = 2) match
      NCK(g_nccl.GroupStart());
      for (int r = 0; r < n; r++) {
        const int lo = lo_of(m, r), hi = hi_of(m, r);
        if (hi >= 0) NCK(g_nccl.Send(m->sbuf_hi[r], cnt, ncclDouble, hi, m->comm[r], m->cstream[r]));
        if (lo >= 0) NCK(g_nccl.Recv(m->rbuf_lo[r], cnt, ncclDouble, lo, m->comm[r], m->cstream[r]));
      }
      for (int r = 0; r < n; r++) {
        const int lo = lo_of(m, r), hi = hi_of(m, r);
        if (lo >= 0) NCK(g_nccl.Send(m->sbuf_lo[r], cnt, ncclDouble, lo, m->comm[r], m->cstream[r]));
        if (hi >= 0) NCK(g_nccl.Recv(m->rbuf_hi[r], cnt, ncclDouble, hi, m->comm[r], m->cstream[r]));
      }
      NCK(g_nccl.GroupEnd());
      m->exchanges++;
      for (int r = 0; r < n; r++) {
        MCK(cudaSetDevice(m->dev[r]));
        MCK(cudaEventRecord(m->ev_done[r], m->cstream[r]));
      }
    }
    // the sweeps that do not read the ghost planes of the split direction run while the planes travel
    for (int r = 0; r < n && !rc; r++) {
      MCK(cudaSetDevice(m->dev[r]));
      rc = pb200_stage_begin(m->ctx[r], s);
    }
    for (int r = 0; r < n && !rc; r++) {
      pb200_ctx *c = m->ctx[r];
      MCK(cudaSetDevice(m->dev[r]));
      if (n > 1) {
        if (m->use_nccl) MCK(cudaStreamWaitEvent(c->stream, m->ev_done[r], 0));
        else {   // the ghost planes were written by the neighbours' copy streams
          if (lo_of(m, r) >= 0) MCK(cudaStreamWaitEvent(c->stream, m->ev_done[lo_of(m, r)], 0));
          if (hi_of(m, r) >= 0) MCK(cudaStreamWaitEvent(c->stream, m->ev_done[hi_of(m, r)], 0));
        }
        double *V = pb200_stage_array(c, s);
        const int npl = c->dev.tot[m->sdir];
        if (lo_of(m, r) >= 0) MCK(pack(m->rbuf_lo[r], V, 0, m, c, true));
        if (hi_of(m, r) >= 0) MCK(pack(m->rbuf_hi[r], V, npl - ng, m, c, true));
      }
      rc = pb200_stage_finish(c, s);
    }
  }
  if (rc) {
    for (int r = 0; r < n; r++) { cudaSetDevice(m->dev[r]); cudaStreamSynchronize(m->ctx[r]->stream); m->ctx[r]->in_step = false; }
    return rc;
  }
  if (n > 1 && m->use_nccl) {
    // MPI_Allreduce(MAX) of invDt_hyp and g_maxMach (main.c:547,288) on the device reduction cells
    NCK(g_nccl.GroupStart());
    for (int r = 0; r < n; r++)
      NCK(g_nccl.AllReduce(m->ctx[r]->d_red, m->ctx[r]->d_red, 2, ncclUint64, ncclMax, m->comm[r], m->ctx[r]->stream));
    NCK(g_nccl.GroupEnd());
  }
  pb200_step_info tot;
  memset(&tot, 0, sizeof(tot));
  int bad = PB200_OK;
  for (int r = 0; r < n; r++) {
    pb200_step_info ir;
    MCK(cudaSetDevice(m->dev[r]));
    int e = pb200_step_end(m->ctx[r], &ir);
    if (e && !bad) bad = e;
    tot.invDt_hyp = ir.invDt_hyp > tot.invDt_hyp ? ir.invDt_hyp : tot.invDt_hyp;
    tot.maxMach = ir.maxMach > tot.maxMach ? ir.maxMach : tot.maxMach;
    tot.c2p_failures += ir.c2p_failures;
    tot.gpu_ms = ir.gpu_ms > tot.gpu_ms ? ir.gpu_ms : tot.gpu_ms;
    tot.launches += ir.launches;
  }
  if (info) *info = tot;
  return bad;
}

// the same call on the host's d->Vc of the GLOBAL grid: the slabs travel up, the step runs, the
// interior planes travel back; the copies of the N devices overlap (one PCIe link each)
extern "C" int pb200_multi_advance_step_host(pb200_multi *m, double *vc_host, double dt, pb200_step_info *info) {
  if (!m || !vc_host) return pb200_fail(PB200_EINVAL, "null argument");
  int rc = multi_copy(m, vc_host, true, false);
  if (rc) return rc;
  rc = pb200_multi_advance_step(m, dt, info);
  if (rc) return rc;
  return multi_copy(m, vc_host, false, true);
}

extern "C" int pb200_multi_integrate(pb200_multi *m, int nsteps, double tstop, double cfl, double cfl_max_var,
                                     double first_dt, double *t, double *dt, pb200_step_info *last) {
  if (!m || !t || !dt) return pb200_fail(PB200_EINVAL, "null argument");
  int done = 0;
  pb200_step_info info;
  memset(&info, 0, sizeof(info));
  for (int k = 0; k < nsteps; k++) {
    bool last_step = false;
    if ((*t + *dt) >= tstop * (1.0 - 1.e-8)) {  // Src/main.c:227-230
      *dt = tstop - *t;
      last_step = true;
    }
    int rc = pb200_multi_advance_step(m, *dt, &info);
    if (rc) return rc;
    *t += *dt;
    double nd = pb200_next_time_step(info.invDt_hyp, cfl, cfl_max_var, *dt, first_dt);
    if (nd < 0.0) return pb200_fail(PB200_EINVAL, "NextTimeStep(): dt is too small");
    *dt = nd;
    done++;
    if (last_step) break;
  }
  if (last) *last = info;
  return done;
}
