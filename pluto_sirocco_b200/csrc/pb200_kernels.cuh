// pb200_kernels.cuh -- fused directional sweep kernels, boundary fills and reductions.
//
// Replaces the per-pencil loop of UpdateStage() (Src/Time_Stepping/update_stage.c:142-323)
// plus the PrimToCons3D / RBoxCopy / convex combination / ConsToPrim3D passes of
// AdvanceStep() (Src/Time_Stepping/rk_step.c:129-130,158,185,235-237,261,304,317).
//
// Data layout in HBM (all FP64, i fastest, same as the reference's d->Vc):
//   V*   [NVAR][NX3_TOT][NX2_TOT][NX1_TOT]  primitive state incl. ghost zones (2 or 3 copies)
//   acc  [NVAR][...same...]                 conservative accumulator  U + dt*(Rx + Ry)
//   cdt  [NX3_TOT][NX2_TOT][NX1_TOT]        C_dt partial sums (update_stage.c:314)
// U0 and Uc of the reference are NOT stored: cons(V) is recomputed in registers from the V
// that the sweep loads anyway.
//
// Kernels per RK stage:
//   1-D : sweep_x1     thread <-> zone along i, neighbours through shared memory
//   2-D : sweep_fused<1,FUSEX,LAST>   marches along x2, x1 sweep of every finished row fused in
//   3-D : sweep_fused<1,FUSEX,!LAST>  (x1+x2 -> acc, cdt)  then  sweep_fused<2,!FUSEX,LAST>
//         (x3 march: acc + Rz, RK combination with cons(V^n), cons->prim, store, C_dt max)
// Marching sweeps: thread <-> i (coalesced), each thread marches along the sweep direction
// keeping the zone value, the backward difference, the previous left state and the previous
// face flux in registers, so every limited slope and every interface flux is computed ONCE.
// All code inside the marching loop is branch free apart from block-uniform conditions.
// Accumulation order is the reference's: ((U + Rx) + Ry) + Rz, then w0*U0 + wc*U.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hd_physics.cuh"

namespace pb {

constexpr int BX = 128;  // threads per block along i

struct Dev {
  int tot[3];   // NX1_TOT, NX2_TOT, NX3_TOT
  int beg[3];   // IBEG, JBEG, KBEG
  int end[3];   // IEND, JEND, KEND
  int ndim;     // DIMENSIONS
  long sj, sk, sv;  // strides (zones): j, k, variable
  Gas gas;
  const double *inv_dx[3];  // 1/dx per zone index, per direction
  // BODY_FORCE (Src/MHD/rhs_source.c:253-281): strided tables evaluated once from the user's
  // BodyForceVector / BodyForcePotential.  value(i,j,k) = tab[q][i*st[q][0]+j*st[q][1]+k*st[q][2]]
  // (indices incl. ghosts; stride 0 = independent of that coordinate).
  //   bf_kind & 1 (VECTOR):    tab[0..2] = g[IDIR], g[JDIR], g[KDIR] at zone centres
  //   bf_kind & 2 (POTENTIAL): tab[3] = Phi at zone centres, tab[4+d] = Phi at the x_d upper face
  // x1 boundary types (0 = none / neighbour / userdef) for the VIRTUAL x1 ghost zones of the fused
  // x1+x2 kernel; entries 2..5 are unused
  int bc_fuse[6], nghost;
  int bf_kind;
  const double *bf_tab[7];
  long bf_st[7][3];
};

PB_D double bf_at(const Dev &d, int q, int i, int j, int k) {
  return __ldg(d.bf_tab[q] + i * d.bf_st[q][0] + j * d.bf_st[q][1] + k * d.bf_st[q][2]);
}

struct SweepArgs {
  const double *V;    // primitive array being swept (ghosts filled)
  const double *V0;   // primitive array at t^n (for the RK combination)
  double *acc;        // conservative accumulator
  double *Vout;       // primitive output of the stage (written by the LAST sweep)
  double *cdt;        // C_dt accumulator (stage 1, DIMENSIONS > 1)
  const double *dt;   // device pointer to g_dt
  unsigned long long *red;  // [0] invDt_hyp bits [1] maxMach bits [2] #cons2prim failures [3] NaN flag
  double w0, wc;
  int comb;    // 0: U ; 1: w0*U0 + wc*U (rk_step.c:236) ; 2: (U0 + 2U)/3 (rk_step.c:304)
  int stage;   // g_intStage
  int limiter;
  int i0;      // first interior zone (relative to IBEG) of this launch of the fused kernel
  int k0, k1;  // x3 planes [k0, k1) relative to KBEG covered by this launch (slab-wise host pipeline); 3-D
  // x3 planes [ko0, ko1) relative to KBEG whose zones count in the reductions (C_dt max, max Mach number).  The
  // whole interior by default; a block that carries a deep halo of planes it does not own (host-buffer steps of a
  // slab-decomposed grid without exchange, pb200_set_owned_planes) restricts them to its own planes, so that
  // invDt_hyp and g_maxMach equal those of the undecomposed grid.
  int ko0, ko1;
};

// local (sweep) component c=1,2,3 -> global velocity variable, Src/set_indexes.c:18-110
template <int DIR>
__host__ __device__ __forceinline__ constexpr int gvar(int c) {
  return c == 0 ? 0 : (c >= 4 ? c : ((c - 1 + DIR) % 3) + 1);
}
// inverse: global variable -> sweep-local component
template <int DIR>
__host__ __device__ __forceinline__ constexpr int lvar(int v) {
  return v == 0 ? 0 : (v >= 4 ? v : ((v - 1 - DIR + 3) % 3) + 1);
}

template <int DIR, int NV>
PB_D void load_zone(const double *__restrict__ V, long off, long sv, double (&q)[NV]) {
#pragma unroll
  for (int c = 0; c < NV; c++) q[c] = __ldg(V + gvar<DIR>(c) * sv + off);
}

PB_D void atomic_max_pos(unsigned long long *p, double x) {
  // valid for x >= 0: IEEE ordering == unsigned integer ordering
  atomicMax(p, (unsigned long long)__double_as_longlong(x));
}

PB_D double warp_max(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// end-of-kernel reduction: block-wide max of (C_dt, Mach), sum of cons2prim failures and the
// NaN flag -> one atomic each per block
template <int BXT>
PB_D void block_reduce_t(double cdt, double mach, int nfail, int nan, bool do_cdt,
                         unsigned long long *red) {
  __shared__ double sa[BXT / 32], sb[BXT / 32];
  __shared__ int sf[BXT / 32], sn[BXT / 32];
  cdt = warp_max(cdt);
  mach = warp_max(mach);
  nfail = __reduce_add_sync(0xffffffffu, nfail);
  nan = __reduce_or_sync(0xffffffffu, nan);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = cdt; sb[w] = mach; sf[w] = nfail; sn[w] = nan; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 1; k < BXT / 32; k++) {
      cdt = fmax(cdt, sa[k]); mach = fmax(mach, sb[k]); nfail += sf[k]; nan |= sn[k];
    }
    if (do_cdt && cdt > 0.0) atomic_max_pos(red + 0, cdt);
    if (mach > 0.0) atomic_max_pos(red + 1, mach);
    if (nfail) atomicAdd(red + 2, (unsigned long long)nfail);
    if (nan) atomicOr(red + 3, 1ull);
  }
}

PB_D void block_reduce(double cdt, double mach, int nfail, int nan, bool do_cdt, unsigned long long *red) {
  block_reduce_t<BX>(cdt, mach, nfail, nan, do_cdt, red);
}

// RK combination (rk_step.c:236,304) with U0 = cons(V^n); U in any component order
template <int NV>
PB_D void rk_combine(double (&U)[NV], const double (&v0z)[NV], const Gas &gas, int comb,
                     double w0, double wc) {
  double U0[NV];
  prim2cons<NV>(v0z, U0, gas);
  if (comb == 1) {
#pragma unroll
    for (int nv = 0; nv < NV; nv++) U[nv] = w0 * U0[nv] + wc * U[nv];
  } else {
    const double one_third = 1.0 / 3.0;
#pragma unroll
    for (int nv = 0; nv < NV; nv++) U[nv] = one_third * (U0[nv] + 2.0 * U[nv]);
  }
}

// RK combination + cons->prim of one zone
template <int NV>
PB_D void combine_c2p(double (&U)[NV], const double (&v0z)[NV], const Gas &gas, int comb,
                      double w0, double wc, double (&vn)[NV], int &nfail, int &nan, bool own) {
  if (comb) rk_combine<NV>(U, v0z, gas, comb, w0, wc);
  int fl = cons2prim<NV>(U, vn, gas);
  nfail += (own && fl != 0) ? 1 : 0;
  nan |= (own && !(vn[iPRS] == vn[iPRS])) ? 1 : 0;   // any NaN in U ends up in the pressure
}

// ------------------------------------------------------------------------------------
//  x1 sweep (DIMENSIONS == 1): thread per zone, neighbour exchange through shared memory
// ------------------------------------------------------------------------------------
template <int NV, int RECON, int SOLVER, int BF>
__global__ void __launch_bounds__(BX) sweep_x1(Dev d, SweepArgs a) {
  constexpr int LO = (RECON == RECON_PARABOLIC) ? 2 : 1;
  constexpr int USE = BX - 1 - LO;
  __shared__ double sm[NV + 2][BX];

  const int t = threadIdx.x;
  const int j = d.beg[1] + blockIdx.y;
  const int k = d.beg[2] + blockIdx.z;
  const int i = d.beg[0] + blockIdx.x * USE + t - LO;
  const int nx = d.tot[0];
  const long row = (long)k * d.sk + (long)j * d.sj;
  auto cl = [nx](int x) { return x < 0 ? 0 : (x >= nx ? nx - 1 : x); };
  const double dt = *a.dt;

  double v0[NV], vp[NV], vm[NV];
  load_zone<0, NV>(a.V, row + cl(i), d.sv, v0);
  if (RECON == RECON_PARABOLIC) {
    double m1[NV], p1[NV], p2[NV];
    load_zone<0, NV>(a.V, row + cl(i - 1), d.sv, m1);
    load_zone<0, NV>(a.V, row + cl(i + 1), d.sv, p1);
    load_zone<0, NV>(a.V, row + cl(i + 2), d.sv, p2);
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      vp[nv] = ppm4_iface(m1[nv], v0[nv], p1[nv], p2[nv]);
      sm[nv][t] = vp[nv];
    }
    __syncthreads();
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      vm[nv] = sm[nv][t > 0 ? t - 1 : 0];
      ppm_parabola(v0[nv], vp[nv], vm[nv], 2.0, 2.0);
    }
    __syncthreads();
  } else if (RECON == RECON_LINEAR) {
    double m1[NV], p1[NV], dvp[NV], dvm[NV];
    load_zone<0, NV>(a.V, row + cl(i - 1), d.sv, m1);
    load_zone<0, NV>(a.V, row + cl(i + 1), d.sv, p1);
#pragma unroll
    for (int nv = 0; nv < NV; nv++) { dvp[nv] = p1[nv] - v0[nv]; dvm[nv] = v0[nv] - m1[nv]; }
    plm_zone<NV, LIM_RT>(v0, dvp, dvm, vp, vm, a.limiter);
  } else {
#pragma unroll
    for (int nv = 0; nv < NV; nv++) vp[nv] = vm[nv] = v0[nv];
  }

  // right state of face i+1/2 is vm of zone i+1
#pragma unroll
  for (int nv = 0; nv < NV; nv++) sm[nv][t] = vm[nv];
  __syncthreads();
  double vR[NV];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) vR[nv] = sm[nv][t < BX - 1 ? t + 1 : t];
  __syncthreads();

  Face<NV> Fp, Fm;
  Ratio mach;
  mach.init();
  const bool face_ok = (t >= LO - 1) && (t <= BX - 2) && (i >= d.beg[0] - 1) && (i <= d.end[0]);
  riemann<NV, SOLVER>(vp, vR, d.gas, Fp, mach, face_ok);
  double phi_p = 0.0;
  if (BF && (d.bf_kind & 2)) {   // TotalFlux(): F_E += F_rho Phi at the face (rhs.c:524-526)
    phi_p = bf_at(d, 4, cl(i), j, k);
    Fp.f[iPRS] += Fp.f[iRHO] * phi_p;
  }

#pragma unroll
  for (int nv = 0; nv < NV; nv++) sm[nv][t] = Fp.f[nv];
  sm[NV][t] = Fp.prs;
  sm[NV + 1][t] = Fp.cmax;
  __syncthreads();
  const int tm = t > 0 ? t - 1 : 0;
#pragma unroll
  for (int nv = 0; nv < NV; nv++) Fm.f[nv] = sm[nv][tm];
  Fm.prs = sm[NV][tm];
  Fm.cmax = sm[NV + 1][tm];

  double cdt_max = 0.0;
  int nfail = 0, nan = 0;
  const double inv_dl = d.inv_dx[0][cl(i)];
  const bool own = t >= LO && t <= BX - 2 && i >= d.beg[0] && i <= d.end[0];
  {
    const double dtdx = dt * inv_dl;
    const long off = row + cl(i);
    double U[NV], vn[NV], v0z[NV];
    prim2cons<NV>(v0, U, d.gas);
#pragma unroll
    for (int nv = 0; nv < NV; nv++) U[nv] += -dtdx * (Fp.f[nv] - Fm.f[nv]);
    U[iVN] -= dtdx * (Fp.prs - Fm.prs);
    if (BF) {   // RightHandSideSource(), x1 sweep of a 1-D grid: all three components
      const int ii = cl(i);
      if (d.bf_kind & 1) {   // rhs_source.c:253-272
        const double g1 = bf_at(d, 0, ii, j, k), g2 = bf_at(d, 1, ii, j, k), g3 = bf_at(d, 2, ii, j, k);
        const double dr = dt * v0[iRHO];
        U[iVN] += dr * g1;
        U[iPRS] += dt * 0.5 * (Fp.f[iRHO] + Fm.f[iRHO]) * g1;
        U[iVT] += dr * g2;
        U[iPRS] += dr * v0[iVT] * g2;
        U[iVB] += dr * g3;
        U[iPRS] += dr * v0[iVB] * g3;
      }
      if (d.bf_kind & 2) {   // rhs_source.c:276-281
        const double phi_m = bf_at(d, 4, cl(i - 1), j, k);
        U[iVN] -= dtdx * v0[iRHO] * (phi_p - phi_m);
        U[iPRS] -= bf_at(d, 3, ii, j, k) * (-dtdx * (Fp.f[iRHO] - Fm.f[iRHO]));
      }
    }
    if (a.comb) load_zone<0, NV>(a.V0, off, d.sv, v0z);
    combine_c2p<NV>(U, v0z, d.gas, a.comb, a.w0, a.wc, vn, nfail, nan, own);
    if (own) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) a.Vout[nv * d.sv + off] = vn[nv];
    }
  }
  // update_stage.c:317-322: every stage, faces IBEG-1..IEND
  if (face_ok) cdt_max = Fp.cmax * inv_dl;
  block_reduce(cdt_max, mach.value(), nfail, nan, true, a.red);
}

// ------------------------------------------------------------------------------------
//  marching sweeps (x2 / x3) with an asynchronous prefetch ring, optionally fused with x1
// ------------------------------------------------------------------------------------
// Every global read of the loop goes through a ring in shared memory filled by cp.async
// (LDGSTS) DEPTH iterations ahead, so HBM latency is covered by a handful of warps per SM
// without spending registers.  With FUSEX the kernel also performs the x1 sweep of every
// finished row: neighbours' zone values are read straight from the ring (the rows stay
// resident two iterations longer), right states and fluxes are exchanged through two small
// shared arrays (2 barriers per row), so the x1 and x2 updates of a stage share ONE read of V
// and ONE write of the accumulator.
#ifndef PB_DEPTH
#define PB_DEPTH 2   // measured: 2 beats 3 (smaller ring, same latency cover) on B200
#endif
#ifndef PB_MINBLK
#define PB_MINBLK 3
#endif
#ifndef PB_UNROLL
#define PB_UNROLL 2
#endif
constexpr int DEPTH = PB_DEPTH;   // cp.async groups in flight
constexpr int MARCH_UNROLL = PB_UNROLL;

PB_D void cp_async8(double *sdst, const double *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
PB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
PB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RECON>
__host__ __device__ constexpr int recon_lead() { return RECON == RECON_PARABOLIC ? 2 : 1; }
template <int RECON>
__host__ __device__ constexpr int recon_xhalo() {
  return RECON == RECON_PARABOLIC ? 3 : (RECON == RECON_LINEAR ? 2 : 1);
}
template <bool FUSEX, int RECON>
__host__ __device__ constexpr int ring_slots() {
  return FUSEX ? DEPTH + recon_lead<RECON>() + 2 : DEPTH + 1;
}
// number of ring quantities a sweep needs (host and device must agree)
__host__ __device__ inline int ring_nq(int nv, bool first, int comb, bool cdt_in) {
  return nv + (first ? 0 : nv) + ((first && comb) ? nv : 0) + (cdt_in ? 1 : 0);
}
// The V rows stay in the ring while the x1 sweep of the fused kernel still needs them (ring_slots);
// the zone-aligned quantities (acc, V^n, cdt of the zone an iteration finishes) are consumed by the
// iteration they land for and live in a second, short ring of DEPTH + 1 slots.
template <bool FUSEX, int NV, int RECON, int BXT = BX>
__host__ __device__ inline size_t sweep_smem_bytes(int nq) {
  return ((size_t)ring_slots<FUSEX, RECON>() * NV * BXT + (size_t)(DEPTH + 1) * (nq - NV) * BXT +
          (FUSEX ? (size_t)(2 * NV + 2) * BXT : 0)) * sizeof(double);
}

enum BcType { BC_NONE = 0, BC_OUTFLOW = 1, BC_REFLECTIVE = 2, BC_AXISYMMETRIC = 3,
              BC_EQTSYMMETRIC = 4, BC_PERIODIC = 5, BC_USERDEF = 8, BC_POLARAXIS = 9, BC_NEIGHBOUR = 100 };

// BXT: threads per block.  The fused x1+x2 kernel loses 2*XH threads per block to the x1 halo, so
// on a 512-wide grid 128-thread blocks need 5 blocks (20 warps) per row where 192-thread blocks
// need 3 (18 warps): the launcher picks the width that runs the fewest warps.
// PB_REG128: run the fused x1+x2 kernel at 16 warps per SM (128 registers per thread, block widths
// 256 / 128 / 64) instead of 12 (168 registers, widths 192 / 160 / 128)
#ifndef PB_REG128
#define PB_REG128 0
#endif
__host__ __device__ constexpr int fused_minblk(bool fusex, int bxt) {
  return (PB_REG128 && fusex) ? 512 / bxt : (bxt == 128 ? PB_MINBLK : 2);
}
template <int DIR, bool FUSEX, bool LAST, int NV, int RECON, int SOLVER, int LIM, int BF, int BXT = BX>
__global__ void __launch_bounds__(BXT, fused_minblk(FUSEX, BXT)) sweep_fused(Dev d, SweepArgs a, int chunk) {
  static_assert(DIR == 1 || DIR == 2, "marching sweeps are x2/x3");
  constexpr int LEAD = recon_lead<RECON>();
  constexpr int XH = recon_xhalo<RECON>();
  constexpr int LO = FUSEX ? XH : 0, HI = FUSEX ? XH : 0;
  constexpr int USE = BXT - LO - HI;
  constexpr int RING = ring_slots<FUSEX, RECON>();
  constexpr bool FIRST = FUSEX;   // the x1(+x2) kernel starts the accumulation: U = cons(V)
  extern __shared__ double smem[];

  const int t = threadIdx.x;
  const int i = d.beg[0] + a.i0 + blockIdx.x * USE + t - LO;
  const bool own = (t >= LO) && (t < BXT - HI) && (i <= d.end[0]);
  const int ic = min(max(i, 0), d.tot[0] - 1);
  const int tr = blockIdx.y + ((DIR == 1 && d.ndim == 3) ? a.k0 : 0);  // transverse index (k for x2 sweeps, j for x3 sweeps)
  // x2 march of a 3-D grid: the whole block sits on one x3 plane (block uniform)
  const bool plane_own = !(DIR == 1 && d.ndim == 3) || (tr >= a.ko0 && tr < a.ko1);
  const long st = (DIR == 1) ? d.sj : d.sk;
  // x1 ghost zones are VIRTUAL in the fused kernel: a halo thread loads the zone Boundary() would have
  // copied into its ghost (outflow: the edge zone, reflective-type: the mirror zone with v_x1 flipped
  // after landing, periodic: the wrapped zone), so the x1 sides need no fill kernel at all
  int isrc = ic;
  bool xflip = false;
  if (FUSEX && (i < d.beg[0] || i > d.end[0])) {
    const int hi = i > d.end[0], type = d.bc_fuse[hi];
    const int nb = d.beg[0], ne = d.end[0], nx1 = ne - nb + 1;
    if (type == BC_OUTFLOW) isrc = hi ? ne : nb;
    else if (type == BC_PERIODIC) isrc = hi ? i - nx1 : i + nx1;
    else if (type != 0) { isrc = hi ? 2 * ne + 1 - i : 2 * nb - 1 - i; xflip = true; }
    isrc = min(max(isrc, 0), d.tot[0] - 1);
  }
  const long base = (DIR == 1) ? ((long)(d.beg[2] + tr) * d.sk + isrc) : ((long)(d.beg[1] + tr) * d.sj + isrc);
  const int cb = d.beg[DIR] + (DIR == 2 ? a.k0 : 0) + blockIdx.z * chunk;
  const int ce = min(cb + chunk - 1, DIR == 2 ? d.beg[2] + a.k1 - 1 : d.end[DIR]);
  const double dt = *a.dt;
  const double *__restrict__ inv_dx = d.inv_dx[DIR];
  const int lim = a.limiter;

  // The RK combination with U0 = cons(V^n) is applied by the kernel that STARTS the
  // accumulation (it is compute bound and has HBM bandwidth to spare for the extra read of
  // V^n):  acc = w0 U0 + wc (U + Rx + Ry);  the x3 kernel then adds wc Rz.
  const int comb = a.comb;
  const bool use_v0 = FIRST && comb != 0;
  const double wscale = (FIRST || comb == 0) ? 1.0 : (comb == 1 ? a.wc : 2.0 / 3.0);
  const bool cdt_on = d.ndim > 1 && a.stage == 1;
  const bool cdt_in = cdt_on && !FIRST;
  const int qA = NV, q0 = qA + (FIRST ? 0 : NV), qC = q0 + (use_v0 ? NV : 0);
  const int nq = qC + (cdt_in ? 1 : 0);
  constexpr int RING2 = DEPTH + 1;
  constexpr int slot_sz = NV * BXT;
  const int z_sz = (nq - NV) * BXT;
  double *ring = smem + t;                              // [RING][NV][BXT]   V rows
  double *ring2 = ring + RING * slot_sz - qA * BXT;     // [RING2][nq-NV][BXT] zone-aligned quantities, indexed from qA
  double *exm = smem + RING * slot_sz + RING2 * z_sz;   // FUSEX: [NV][BXT] right states (vm)
  double *exf = exm + NV * BXT;                          //        [NV+2][BXT] fluxes, prs, cmax

  const int n0 = cb - LEAD;  // first iteration: consumes V row cb, so the ring holds rows >= cb
  // iteration m consumes V row m+LEAD and the stored quantities of zone m-1
  auto issue = [&](int m, int s, int s2) {
    if (m <= ce + 1) {
      double *slot = ring + s * slot_sz;
      double *zslot = ring2 + s2 * z_sz;
      const long oV = base + (long)(m + LEAD) * st;
#pragma unroll
      for (int c = 0; c < NV; c++) cp_async8(slot + c * BXT, a.V + gvar<DIR>(c) * d.sv + oV);
      const int z = m - 1;
      if (z >= cb && own) {
        const long oz = base + (long)z * st;
        if (!FIRST) {
#pragma unroll
          for (int v = 0; v < NV; v++) cp_async8(zslot + (qA + v) * BXT, a.acc + v * d.sv + oz);
        }
        if (use_v0) {
#pragma unroll
          for (int v = 0; v < NV; v++) cp_async8(zslot + (q0 + v) * BXT, a.V0 + v * d.sv + oz);
        }
        if (cdt_in) cp_async8(zslot + qC * BXT, a.cdt + oz);
      }
    }
    cp_async_commit();
  };
  auto wrap = [](int s) { return s >= RING ? s - RING : s; };
  auto wrap2 = [](int s) { return s >= RING2 ? s - RING2 : s; };

  Ratio mach;
  mach.init();
  double cdt_max = 0.0;
  int nfail = 0, nan = 0;
  double v0[NV], dvm[NV];          // PLM/FLAT carry: zone n and its backward difference
  double vm1[NV], vp1[NV], qm[NV]; // PPM carry: zones n-1, n+1, interface value at n-1/2
  double vpL[NV];                  // left state of face n-1/2 (= vp of zone n-1)
  Face<NV> Fm;                     // face n-3/2

#pragma unroll
  for (int g = 0; g < DEPTH; g++) issue(n0 + g, g, g);
  {
    double b0[NV];
    load_zone<DIR, NV>(a.V, base + (long)(n0 - 1) * st, d.sv, b0);
    load_zone<DIR, NV>(a.V, base + (long)n0 * st, d.sv, v0);
    if (RECON == RECON_PARABOLIC) {
      // iteration n0 = cb-2 only builds the interface value at cb-3/2; its own states
      // (which would need row cb-4) are never used, so qm may hold anything finite
      load_zone<DIR, NV>(a.V, base + (long)(n0 + 1) * st, d.sv, vp1);
#pragma unroll
      for (int nv = 0; nv < NV; nv++) { vm1[nv] = b0[nv]; qm[nv] = v0[nv]; }
    } else {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) dvm[nv] = v0[nv] - b0[nv];
    }
  }
#pragma unroll
  for (int nv = 0; nv < NV; nv++) { vpL[nv] = v0[nv]; Fm.f[nv] = 0.0; }
  Fm.prs = 0.0;
  Fm.cmax = 0.0;
  // BODY_FORCE carries: centre density of zone n (x3 march; the fused kernel has the whole row
  // in the ring) and the potential at the face behind
  double rho_c = v0[iRHO], phi_m = 0.0;
  const int jt = (DIR == 1) ? 0 : d.beg[1] + tr, kt = (DIR == 1) ? d.beg[2] + tr : 0;

  int sc = 0, sc2 = 0;  // ring slots consumed by this iteration
#pragma unroll MARCH_UNROLL
  for (int n = n0; n <= ce + 1; n++) {
    // ---- data of this iteration has landed in the ring; refill the slot DEPTH ahead ----
    cp_async_wait<DEPTH - 1>();
    if (FUSEX && xflip) {    // mirror ghost: flip v_x1 of the row that just landed (once per row)
      double *w = ring + sc * slot_sz + lvar<DIR>(1) * BXT;
      *w = -*w;
    }
    const double *slot = ring + sc * slot_sz;
    const double *zslot = ring2 + sc2 * z_sz;
    double vin[NV];
#pragma unroll
    for (int c = 0; c < NV; c++) vin[c] = slot[c * BXT];
    const int z = n - 1;          // zone finished by this iteration
    const bool fin = n >= cb + 1;  // block-uniform
    double U[NV], v0z[NV], cin = 0.0;   // U, v0z: GLOBAL variable order
    if (fin) {
      if (!FIRST) {
#pragma unroll
        for (int v = 0; v < NV; v++) U[v] = zslot[(qA + v) * BXT];
      }
      if (use_v0) {
#pragma unroll
        for (int v = 0; v < NV; v++) v0z[v] = zslot[(q0 + v) * BXT];
      }
      if (cdt_in) cin = zslot[qC * BXT];
    }
    const double inv_dl = __ldg(inv_dx + max(z, 0));
    issue(n + DEPTH, wrap(sc + DEPTH), wrap2(sc2 + DEPTH));

    // ---- reconstruct zone n along DIR ----
    double vp[NV], vm[NV];
    if (RECON == RECON_PARABOLIC) {
      // rows: vm1 = n-1, v0 = n, vp1 = n+1, vin = n+2
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        double q = ppm4_iface(vm1[nv], v0[nv], vp1[nv], vin[nv]);
        vp[nv] = q;
        vm[nv] = qm[nv];
        qm[nv] = q;
        ppm_parabola(v0[nv], vp[nv], vm[nv], 2.0, 2.0);
      }
    } else if (RECON == RECON_LINEAR) {
      double dvp[NV];
#pragma unroll
      for (int nv = 0; nv < NV; nv++) dvp[nv] = vin[nv] - v0[nv];
      plm_zone<NV, LIM>(v0, dvp, dvm, vp, vm, lim);
#pragma unroll
      for (int nv = 0; nv < NV; nv++) dvm[nv] = dvp[nv];
    } else {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) vp[nv] = vm[nv] = v0[nv];
    }

    // ---- face n-1/2 along DIR (result unused before n = cb) ----
    // With FUSEX the call is placed between the two barriers of the x1 sweep, next to the x1
    // Riemann problem: two independent dependency chains for the scheduler to interleave.
    Face<NV> Fp;
    // x3 march: face n-1/2 counts if one of its two zones (n-1, n) is owned
    const bool face_own = (DIR == 2) ? (n - d.beg[2] >= a.ko0 && n - d.beg[2] <= a.ko1) : plane_own;
    if (!FUSEX || !fin) riemann<NV, SOLVER>(vpL, vm, d.gas, Fp, mach, own && n >= cb && face_own);
    // zone z = n-1 sits at (ic, z, kt) for the x2 march and (ic, jt, z) for the x3 march
    const int zj = (DIR == 1) ? max(z, 0) : jt, zk = (DIR == 1) ? kt : max(z, 0);
    double phi_p = 0.0;
    if (BF && (d.bf_kind & 2)) phi_p = bf_at(d, 4 + DIR, ic, zj, zk);

    if (fin) {
      double cx = 0.0;
      double vx[NV];   // FUSEX: centre state of zone z, global component order
      if (FUSEX) {
        // ---- x1 sweep of row z: thread <-> zone; neighbours' values straight from the ring
        int sz = sc - (1 + LEAD);
        sz = sz < 0 ? sz + RING : sz;
        const double *rz = smem + sz * slot_sz;   // row z, all threads
        const int tm = t > 0 ? t - 1 : 0, tp = t < BXT - 1 ? t + 1 : BXT - 1;
        double xp[NV], xm[NV];                    // x1-local = global component order
#pragma unroll
        for (int v = 0; v < NV; v++) vx[v] = rz[lvar<DIR>(v) * BXT + t];
        if (RECON == RECON_PARABOLIC) {
          const int tp2 = t < BXT - 2 ? t + 2 : BXT - 1;
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const double *r = rz + lvar<DIR>(v) * BXT;
            xp[v] = ppm4_iface(r[tm], vx[v], r[tp], r[tp2]);
            exm[v * BXT + t] = xp[v];   // interface value at t+1/2
          }
          __syncthreads();
#pragma unroll
          for (int v = 0; v < NV; v++) {
            xm[v] = exm[v * BXT + tm];
            ppm_parabola(vx[v], xp[v], xm[v], 2.0, 2.0);
          }
          __syncthreads();
        } else if (RECON == RECON_LINEAR) {
          double dp[NV], dm[NV];
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const double *r = rz + lvar<DIR>(v) * BXT;
            dp[v] = r[tp] - vx[v];
            dm[v] = vx[v] - r[tm];
          }
          plm_zone<NV, LIM>(vx, dp, dm, xp, xm, lim);
        } else {
#pragma unroll
          for (int v = 0; v < NV; v++) xp[v] = xm[v] = vx[v];
        }
#pragma unroll
        for (int v = 0; v < NV; v++) exm[v * BXT + t] = xm[v];
        __syncthreads();
        double xr[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) xr[v] = exm[v * BXT + tp];
        Face<NV> Gp;
        const bool xface = t >= LO - 1 && t < BXT - HI && i >= d.beg[0] - 1 && i <= d.end[0];
        riemann<NV, SOLVER>(xp, xr, d.gas, Gp, mach, xface && plane_own);
        riemann<NV, SOLVER>(vpL, vm, d.gas, Fp, mach, own && plane_own);
        double phx = 0.0;
        if (BF && (d.bf_kind & 2)) {   // TotalFlux(): F_E += F_rho Phi(x1p) (rhs.c:524-526)
          phx = bf_at(d, 4, ic, zj, zk);
          Gp.f[iPRS] += Gp.f[iRHO] * phx;
        }
#pragma unroll
        for (int v = 0; v < NV; v++) exf[v * BXT + t] = Gp.f[v];
        exf[NV * BXT + t] = Gp.prs;
        exf[(NV + 1) * BXT + t] = Gp.cmax;
        __syncthreads();
        const double idx1 = __ldg(d.inv_dx[0] + ic);
        const double dtdx1 = dt * idx1;
        prim2cons<NV>(vx, U, d.gas);
#pragma unroll
        for (int v = 0; v < NV; v++) U[v] += -dtdx1 * (Gp.f[v] - exf[v * BXT + tm]);
        U[iVN] -= dtdx1 * (Gp.prs - exf[NV * BXT + tm]);
        cx = 0.5 * (exf[(NV + 1) * BXT + tm] + Gp.cmax) * idx1;
        if (BF) {   // RightHandSideSource(), x1 sweep (rhs_source.c:253-281)
          if (d.bf_kind & 1) {
            const double g1 = bf_at(d, 0, ic, zj, zk);
            U[iVN] += dt * vx[iRHO] * g1;
            U[iPRS] += dt * 0.5 * (Gp.f[iRHO] + exf[iRHO * BXT + tm]) * g1;
          }
          if (d.bf_kind & 2) {
            const double phm = bf_at(d, 4, max(ic - 1, 0), zj, zk);
            U[iVN] -= dtdx1 * vx[iRHO] * (phx - phm);
            U[iPRS] -= bf_at(d, 3, ic, zj, zk) * (-dtdx1 * (Gp.f[iRHO] - exf[iRHO * BXT + tm]));
          }
        }
      }
      if (BF && (d.bf_kind & 2)) Fp.f[iPRS] += Fp.f[iRHO] * phi_p;

      // ---- finish zone z: add this direction, combine, cons->prim ----
      const double dtdx = dt * inv_dl * wscale;
      const long oz = base + (long)z * st;
#pragma unroll
      for (int c = 0; c < NV; c++) U[gvar<DIR>(c)] += -dtdx * (Fp.f[c] - Fm.f[c]);
      U[gvar<DIR>(iVN)] -= dtdx * (Fp.prs - Fm.prs);
      if (BF) {   // RightHandSideSource(), x2 / x3 sweep (rhs_source.c:360-384 and the x3 twin)
        const double dts = dt * wscale;
        double rz, vbz;   // centre density and the velocity along the next (inactive) direction
        if (FUSEX) { rz = vx[iRHO]; vbz = vx[3]; } else { rz = rho_c; vbz = 0.0; }
        if (d.bf_kind & 1) {
          const double gn = bf_at(d, DIR, ic, zj, zk);
          U[gvar<DIR>(iVN)] += dts * rz * gn;
          U[iPRS] += dts * 0.5 * (Fp.f[iRHO] + Fm.f[iRHO]) * gn;
          if (DIR == 1 && d.ndim == 2) {   // !INCLUDE_KDIR: g[KDIR] is added by the x2 sweep
            const double g3 = bf_at(d, 2, ic, zj, zk);
            U[3] += dts * rz * g3;
            U[iPRS] += dts * rz * vbz * g3;
          }
        }
        if (d.bf_kind & 2) {
          U[gvar<DIR>(iVN)] -= dtdx * rz * (phi_p - phi_m);
          U[iPRS] -= bf_at(d, 3, ic, zj, zk) * (-dtdx * (Fp.f[iRHO] - Fm.f[iRHO]));
        }
      }
      if (LAST) {
        double vn[NV];
        combine_c2p<NV>(U, v0z, d.gas, FIRST ? comb : 0, a.w0, a.wc, vn, nfail, nan, own);
        if (own) {
#pragma unroll
          for (int v = 0; v < NV; v++) a.Vout[v * d.sv + oz] = vn[v];
        }
      } else {
        if (use_v0) rk_combine<NV>(U, v0z, d.gas, comb, a.w0, a.wc);
        if (own) {
#pragma unroll
          for (int v = 0; v < NV; v++) a.acc[v * d.sv + oz] = U[v];
        }
      }
      if (cdt_on) {
        double c = 0.5 * (Fm.cmax + Fp.cmax) * inv_dl;
        if (FUSEX) c = cx + c;
        else if (!FIRST) c = cin + c;
        const bool zone_own = (DIR == 2) ? (z - d.beg[2] >= a.ko0 && z - d.beg[2] < a.ko1) : plane_own;
        if (LAST) cdt_max = fmax(cdt_max, (own && zone_own) ? c : 0.0);
        else if (own) a.cdt[oz] = c;
      }
    } else {
      if (BF && (d.bf_kind & 2)) Fp.f[iPRS] += Fp.f[iRHO] * phi_p;
      if (FUSEX) __syncthreads();   // rows landed so far become visible to the neighbours
    }
    Fm = Fp;
    if (BF) { phi_m = phi_p; rho_c = v0[iRHO]; }
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      vpL[nv] = vp[nv];
      if (RECON == RECON_PARABOLIC) { vm1[nv] = v0[nv]; v0[nv] = vp1[nv]; vp1[nv] = vin[nv]; }
      else v0[nv] = vin[nv];
    }
    sc = wrap(sc + 1);
    sc2 = wrap2(sc2 + 1);
  }
  cp_async_wait<0>();
  block_reduce_t<BXT>(cdt_max, mach.value(), nfail, nan, a.stage == 1 && LAST, a.red);
}

// ------------------------------------------------------------------------------------
//  physical boundaries  (Src/boundary.c:228-459 dispatch, :617-767 fills, :503 FlipSign)
// ------------------------------------------------------------------------------------

struct BcArgs {
  double *V;
  int side;       // 0..5 = X1_BEG, X1_END, X2_BEG, ...
  int type;
  int nvar;
  int nghost;
  double sign[16];
  int k0, k1;     // restrict the fill of an x1 / x2 side to the x3 planes [k0, k1) (absolute indices)
  int pdir;       // BC_POLARAXIS: the phi direction (x2 POLAR, x3 SPHERICAL), PolarAxisBoundary() boundary.c:770-840
};

static __global__ void bc_fill(Dev d, BcArgs b) {
  const int dir = b.side >> 1;
  const bool hi = b.side & 1;
  // extents of the ghost box: nghost layers along dir, full transverse range
  int ext[3] = {d.tot[0], d.tot[1], d.tot[2]};
  ext[dir] = b.nghost;
  const int koff = dir < 2 ? b.k0 : 0;
  if (dir < 2) ext[2] = b.k1 - b.k0;
  long ntot = (long)ext[0] * ext[1] * ext[2];
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ntot) return;
  int c[3];
  c[0] = (int)(idx % ext[0]);
  c[1] = (int)((idx / ext[0]) % ext[1]);
  c[2] = (int)(idx / ((long)ext[0] * ext[1])) + koff;
  const int g = c[dir];
  const int nb = d.beg[dir], ne = d.end[dir], nx = ne - nb + 1;
  const int n = hi ? ne + 1 + g : nb - 1 - g;
  int src;
  if (b.type == BC_OUTFLOW) src = hi ? ne : nb;
  else if (b.type == BC_PERIODIC) src = hi ? n - nx : n + nx;
  else src = hi ? 2 * ne + 1 - n : 2 * nb - 1 - n;
  const bool flip = (b.type == BC_REFLECTIVE || b.type == BC_AXISYMMETRIC || b.type == BC_EQTSYMMETRIC || b.type == BC_POLARAXIS);
  c[dir] = n;
  long dst_off = (long)c[2] * d.sk + (long)c[1] * d.sj + c[0];
  c[dir] = src;
  if (b.type == BC_POLARAXIS) {      // the mirror zone lies half a turn away
    const int nphi = d.end[b.pdir] - d.beg[b.pdir] + 1;
    c[b.pdir] += nphi / 2;
    if (c[b.pdir] > d.end[b.pdir]) c[b.pdir] -= nphi;
  }
  long src_off = (long)c[2] * d.sk + (long)c[1] * d.sj + c[0];
  for (int nv = 0; nv < b.nvar; nv++) {
    double q = b.V[nv * d.sv + src_off];
    if (flip) q = b.sign[nv] * q;
    b.V[nv * d.sv + dst_off] = q;
  }
}

}  // namespace pb
