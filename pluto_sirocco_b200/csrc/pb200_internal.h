// pb200_internal.h -- state shared by the translation units of libplutob200.so (not part of
// the ABI; the public interface is include/pluto_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/pluto_b200.h"
#include "pb200_kernels.cuh"

namespace pb { struct GenDev; }

struct pb200_ctx {
  pb::Dev dev;
  pb200_config cfg;
  int nvar;
  long nzone;        // zones incl. ghosts
  size_t vbytes;     // bytes of one [nvar] state array
  double *V[3];      // primitive state copies (A = current d->Vc, B, C)
  double *acc;       // conservative accumulator (DIMENSIONS > 1)
  double *cdt;       // C_dt
  double *d_dt;      // device g_dt
  unsigned long long *d_red;   // reduction cell: invDt bits, maxMach bits, #fail, NaN flag
  unsigned long long *h_red;   // pinned
  double *h_dt;                // pinned
  double *d_invdx[3];
  double *d_bf[7];   // body-force tables (Dev::bf_tab)
  std::vector<double> xl[3], xr[3], dx[3];
  // the caller's own Grid arrays (pb200_set_geometry): used by the general path instead of its own evaluation
  std::vector<double> geo_dV, geo_A[3], geo_dxdl[3], geo_rt, geo_s, geo_sp;
  bool geo_set;
  int grid_uniform[3];     // grid->uniform[d] (set_grid.c:67-72): 1 / 0 from pb200_set_grid_uniform, -1: inferred from dx
  cudaStream_t stream;
  cudaStream_t h2d, d2h;        // copy streams of the slab-wise host pipeline (pb200_advance_step_host)
  cudaEvent_t ev_up[64], ev_done[64];
  int host_pipeline;            // x3 planes per slab of that pipeline, 0: plain copy - step - copy
  cudaEvent_t ev0, ev1;
  int launches;
  int cur;           // index of the array holding d->Vc
  int nstages;
  int stage_in[4], stage_out[4];  // array indices per stage (1-based)
  bool in_step;
  // optional per-kernel timing (pb200_set_profiling)
  bool profiling;
  int nprof;
  cudaEvent_t pev0[16], pev1[16];
  int pdir[16], pstage[16];
  float pms[16];
  // general-grid path (pb200_gen.cu / gen_kernels.cuh)
  bool gen, gen_ready;
  pb::GenDev *gdev;
  double *gU, *gU0, *gVP, *gVM, *gF, *gcdt;
  unsigned short *gflag;
  unsigned char *gshock;
  std::vector<void *> gen_allocs;
  // line-driven wind
  bool ldw_on;
  pb200_ldw_config ldw;
  double *ldw_flux[3], *ldw_dvds;         // ldw_dvds: the two line-force arrays (g_r, g_theta) gen_vgrad leaves for the sweeps
  unsigned long long *ldw_mask;           // per zone: bins with a non-zero flux (gen_ldw_mask)
  int ldw_mpoints;                        // force-multiplier fit (0: power law)
  double *ldw_tfit, *ldw_mfit;
  double *cool_tab[7];                    // BLONDIN tables (null: defaults)
  int cur_stage;                          // stage whose Boundary() is being filled (0: outside a step)
  // FLAG_INTERNAL_BOUNDARY zones (Src/int_bound_reset.c): byte mask over all zones (general path) and
  // the list of flagged interior zones (fast path: ib_fix after the last sweep of a stage)
  // CUDA graph of one whole AdvanceStep (small grids are launch bound: Sod-400 has 7 launches of a few
  // microseconds per step, the line-driven wind 30): captured on the first pb200_advance_step() and
  // replayed while nothing that feeds a kernel argument has changed (graph_sig)
  cudaGraphExec_t graph_exec;
  unsigned long long graph_sig;
  int graph_launches, use_graph, gen_epoch;
  bool capturing;
  int own_k0, own_k1;                     // pb200_set_owned_planes (relative to KBEG); default: the whole interior
  bool stage_uploaded;                    // pb200_stage_upload() replaced the array the next stage sweeps
  unsigned char *d_ibmask;
  long *d_iblist;
  long ib_n;
};

int  pb200_fail(int code, const char *msg);   // sets pb200_last_error(), returns code
int  pb200_gen_setup(pb200_ctx *c);
void pb200_gen_release(pb200_ctx *c);
int  pb200_gen_stage(pb200_ctx *c, int stage);
int  pb200_gen_patch_u(pb200_ctx *c, long n, const long *zone, const double *u);
bool pb200_grid_is_uniform(const pb200_ctx *c, int dir);
int  pb200_gen_ring_start(pb200_ctx *c);
int  pb200_gen_internal_boundary(pb200_ctx *c, double *V);   // UserDefBoundary(side == 0)
int  pb200_gen_userdef_side(pb200_ctx *c, double *V, int side);
int  pb200_gen_entropy(pb200_ctx *c, double *V);               // ComputeEntropy at the end of Boundary()


// One launcher per (NVAR, body force) pair, each compiled in its own translation unit
// (pb200_sweeps.cu with -DPB_NV=.. -DPB_BF=..) so that the template instantiations build in
// parallel.  dir 0: x1 sweep of a 1-D grid; 1: fused x1+x2 march; 2: x3 march.
typedef void (*pb200_sweep_fn)(pb200_ctx *, int dir, const pb::SweepArgs &);
#define PB200_SWEEP_DECL(NV, BF) void pb200_launch_sweep_nv##NV##_bf##BF(pb200_ctx *, int, const pb::SweepArgs &);
PB200_SWEEP_DECL(5, 0) PB200_SWEEP_DECL(5, 1)
PB200_SWEEP_DECL(6, 0) PB200_SWEEP_DECL(6, 1)
PB200_SWEEP_DECL(7, 0) PB200_SWEEP_DECL(7, 1)
