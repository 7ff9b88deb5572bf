// pb200_kernels.cuh -- fused directional sweep kernels, boundary fills and reductions.
//
// Replaces the per-pencil loop of UpdateStage() (Src/Time_Stepping/update_stage.c:142-323)
// plus the PrimToCons3D / RBoxCopy / convex combination / ConsToPrim3D passes of
// AdvanceStep() (Src/Time_Stepping/rk_step.c:129-130,158,185,235-237,261,304,317).
//
// Data layout in HBM (all FP64, i fastest, same as the reference's d->Vc):
//   V*   [NVAR][NX3_TOT][NX2_TOT][NX1_TOT]  primitive state incl. ghost zones (2 or 3 copies)
//   acc  [NVAR][...same...]                 conservative accumulator  U + dt*(Rx [+Ry])
//   cdt  [NX3_TOT][NX2_TOT][NX1_TOT]        C_dt partial sums (update_stage.c:314)
// U0 and Uc of the reference are NOT stored: cons(V) is recomputed in registers from the V
// that the sweep loads anyway, so a stage moves (stage 1) 40 B in + 40 B out through HBM in
// 1-D, and V once per direction + one accumulator round trip in 2-D/3-D.
//
// One kernel per direction; every interface flux and every limited slope is computed ONCE:
//   x1 sweep  : thread <-> zone along i; L/R states and fluxes are exchanged between
//               neighbouring threads through shared memory (2 halo threads of 128).
//   x2/x3 sweep: thread <-> i (coalesced), each thread MARCHES along j (or k) keeping the
//               stencil, the previous left state and the previous flux in registers.
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
};

struct SweepArgs {
  const double *V;    // primitive array being swept (ghosts filled)
  const double *V0;   // primitive array at t^n (for the RK combination)
  double *acc;        // conservative accumulator
  double *Vout;       // primitive output of the stage (written by the LAST sweep)
  double *cdt;        // C_dt accumulator (stage 1, DIMENSIONS > 1)
  const double *dt;   // device pointer to g_dt
  unsigned long long *red;  // [0] invDt_hyp bits  [1] maxMach bits  [2] #cons2prim failures
  double w0, wc;
  int comb;    // 0: U ; 1: w0*U0 + wc*U (rk_step.c:236) ; 2: (U0 + 2U)/3 (rk_step.c:304)
  int first;   // this sweep starts the accumulation: U = cons(V)
  int last;    // this sweep finishes the stage: combination + cons->prim + store Vout
  int stage;   // g_intStage
  int limiter;
};

// local (sweep) component c=1,2,3 -> global velocity variable, Src/set_indexes.c:18-110
template <int DIR>
__device__ __forceinline__ constexpr int gvar(int c) {
  return c == 0 ? 0 : (c >= 4 ? c : ((c - 1 + DIR) % 3) + 1);
}

template <int DIR, int NV>
__device__ __forceinline__ void load_zone(const double *__restrict__ V, long off, long sv,
                                          double (&q)[NV]) {
#pragma unroll
  for (int c = 0; c < NV; c++) q[c] = __ldg(V + gvar<DIR>(c) * sv + off);
}
template <int DIR, int NV>
__device__ __forceinline__ void store_zone(double *__restrict__ V, long off, long sv,
                                           const double (&q)[NV]) {
#pragma unroll
  for (int c = 0; c < NV; c++) V[gvar<DIR>(c) * sv + off] = q[c];
}

__device__ __forceinline__ void atomic_max_pos(unsigned long long *p, double x) {
  // valid for x >= 0: IEEE ordering == unsigned integer ordering
  atomicMax(p, (unsigned long long)__double_as_longlong(x));
}

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// block-wide max of two values -> global atomics (one per block)
__device__ __forceinline__ void block_reduce_max2(double a, double b, bool do_a,
                                                  unsigned long long *red) {
  __shared__ double sa[BX / 32], sb[BX / 32];
  a = warp_max(a);
  b = warp_max(b);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 1; k < BX / 32; k++) { a = fmax(a, sa[k]); b = fmax(b, sb[k]); }
    if (do_a && a > 0.0) atomic_max_pos(red + 0, a);
    if (b > 0.0) atomic_max_pos(red + 1, b);
  }
}

template <int LIM_RT_DUMMY = 0>
__device__ __forceinline__ double slope_rt(int lim, int nv, double dp, double dm) {
  switch (lim) {
    case LIM_FLAT: return 0.0;
    case LIM_MINMOD: return lim_mm(dp, dm);
    case LIM_VANLEER: return lim_vl(dp, dm);
    case LIM_MC: return lim_mc(dp, dm);
    case LIM_VANALBADA: return lim_va(dp, dm);
    case LIM_OSPRE: return lim_os(dp, dm);
    case LIM_UMIST: return lim_um(dp, dm);
    default: return plm_slope<LIM_DEFAULT>(nv, dp, dm);
  }
}

template <int NV>
__device__ __forceinline__ void plm_rt(int lim, const double (&vm1)[NV], const double (&v0)[NV],
                                       const double (&vp1)[NV], double (&vp)[NV],
                                       double (&vm)[NV]) {
  if (lim == LIM_DEFAULT) {
    plm_zone<NV, LIM_DEFAULT>(vm1, v0, vp1, vp, vm);
  } else {
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      double dvp = vp1[nv] - v0[nv], dvm = v0[nv] - vm1[nv];
      double dv = slope_rt<>(lim, nv, dvp, dvm);
      vp[nv] = v0[nv] + dv * 0.5;
      vm[nv] = v0[nv] - dv * 0.5;
    }
  }
}

// Finish one zone: rhs from the two faces, accumulate, optionally combine + cons->prim.
// vz = primitive state of the zone (sweep-local order), off = linear zone offset.
template <int DIR, int NV>
__device__ __forceinline__ void finish_zone(const Dev &d, const SweepArgs &a, long off,
                                            const double (&vz)[NV], const Face<NV> &Fm,
                                            const Face<NV> &Fp, double dtdx, double inv_dl,
                                            double &cdt_max) {
  double U[NV];
  if (a.first) {
    prim2cons<NV>(vz, U, d.gas);
  } else {
    load_zone<DIR, NV>(a.acc, off, d.sv, U);
  }
#pragma unroll
  for (int nv = 0; nv < NV; nv++) U[nv] += -dtdx * (Fp.f[nv] - Fm.f[nv]);
  U[iVN] -= dtdx * (Fp.prs - Fm.prs);

  if (a.last) {
    if (a.comb) {
      double v0[NV], U0[NV];
      load_zone<DIR, NV>(a.V0, off, d.sv, v0);
      prim2cons<NV>(v0, U0, d.gas);
      if (a.comb == 1) {
#pragma unroll
        for (int nv = 0; nv < NV; nv++) U[nv] = a.w0 * U0[nv] + a.wc * U[nv];
      } else {
        const double one_third = 1.0 / 3.0;
#pragma unroll
        for (int nv = 0; nv < NV; nv++) U[nv] = one_third * (U0[nv] + 2.0 * U[nv]);
      }
    }
    double vn[NV];
    int fail = cons2prim<NV>(U, vn, d.gas);
    if (fail) atomicAdd(a.red + 2, 1ull);
    store_zone<DIR, NV>(a.Vout, off, d.sv, vn);
  } else {
    store_zone<DIR, NV>(a.acc, off, d.sv, U);
  }
  // inverse time step, update_stage.c:303-316 (DIMENSIONS > 1, predictor only)
  if (d.ndim > 1 && a.stage == 1) {
    double c = 0.5 * (Fm.cmax + Fp.cmax) * inv_dl;
    if (!a.first) c = a.cdt[off] + c;
    if (a.last) cdt_max = fmax(cdt_max, c);
    else a.cdt[off] = c;
  }
}

// ------------------------------------------------------------------------------------
//  x1 sweep: thread per zone, neighbour exchange through shared memory
// ------------------------------------------------------------------------------------
template <int NV, int RECON, int SOLVER>
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
    double m1[NV], p1[NV];
    load_zone<0, NV>(a.V, row + cl(i - 1), d.sv, m1);
    load_zone<0, NV>(a.V, row + cl(i + 1), d.sv, p1);
    plm_rt<NV>(a.limiter, m1, v0, p1, vp, vm);
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
  double mach = 0.0;
  riemann<NV, SOLVER>(vp, vR, d.gas, Fp, mach);
  const bool face_ok = (t >= LO - 1) && (t <= BX - 2) && (i >= d.beg[0] - 1) && (i <= d.end[0]);
  if (!face_ok) mach = 0.0;

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
  const double inv_dl = d.inv_dx[0][cl(i)];
  if (t >= LO && t <= BX - 2 && i >= d.beg[0] && i <= d.end[0]) {
    finish_zone<0, NV>(d, a, row + i, v0, Fm, Fp, dt * inv_dl, inv_dl, cdt_max);
  }
  if (d.ndim == 1) {  // update_stage.c:317-322: every stage, faces IBEG-1..IEND
    if (face_ok) cdt_max = Fp.cmax * inv_dl;
  }
  const bool want_dt = (d.ndim == 1) || (a.stage == 1 && a.last);
  block_reduce_max2(cdt_max, mach, want_dt, a.red);
}

// ------------------------------------------------------------------------------------
//  marching sweeps (x2 / x3) with an asynchronous prefetch ring, optionally fused with x1
// ------------------------------------------------------------------------------------
// thread <-> i (coalesced); each thread marches along direction DIR keeping the stencil, the
// previous left state and the previous flux in registers.  Every global read of the loop goes
// through a per-thread ring in shared memory filled by cp.async (LDGSTS) DEPTH iterations
// ahead, so HBM latency is covered by a handful of warps per SM without spending registers.
// With FUSEX the kernel also performs the x1 sweep of every finished row: the three
// neighbour exchanges (centre states, right states, fluxes) go through shared memory, so
// the x1 and x2 updates of a stage share ONE read of V and ONE write of the accumulator.
constexpr int RING = 4;    // ring slots
constexpr int DEPTH = 3;   // cp.async groups in flight

__device__ __forceinline__ void cp_async8(double *sdst, const double *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// number of ring quantities a sweep needs (host and device must agree)
__host__ __device__ inline int ring_nq(int nv, bool first, bool last, int comb, bool cdt_in) {
  return nv + (first ? 0 : nv) + ((last && comb) ? nv : 0) + (cdt_in ? 1 : 0);
}

template <int DIR, bool FUSEX, int NV, int RECON, int SOLVER>
__global__ void __launch_bounds__(BX, 3) sweep_fused(Dev d, SweepArgs a, int chunk) {
  static_assert(DIR == 1 || DIR == 2, "marching sweeps are x2/x3");
  constexpr int LEAD = (RECON == RECON_PARABOLIC) ? 2 : 1;
  constexpr int XH = (RECON == RECON_PARABOLIC) ? 3 : (RECON == RECON_LINEAR ? 2 : 1);
  constexpr int LO = FUSEX ? XH : 0, HI = FUSEX ? XH : 0;
  constexpr int USE = BX - LO - HI;
  extern __shared__ double smem[];

  const int t = threadIdx.x;
  const int i = d.beg[0] + blockIdx.x * USE + t - LO;
  const bool own = (t >= LO) && (t < BX - HI) && (i <= d.end[0]);
  const int ic = min(max(i, 0), d.tot[0] - 1);
  const int tr = blockIdx.y;  // transverse index (k for x2 sweeps, j for x3 sweeps)
  const long st = (DIR == 1) ? d.sj : d.sk;
  const long base = (DIR == 1) ? ((long)(d.beg[2] + tr) * d.sk + ic) : ((long)(d.beg[1] + tr) * d.sj + ic);
  const int cb = d.beg[DIR] + blockIdx.z * chunk;
  const int ce = min(cb + chunk - 1, d.end[DIR]);
  const double dt = *a.dt;
  const double *__restrict__ inv_dx = d.inv_dx[DIR];

  const bool first = FUSEX ? true : (a.first != 0);
  const bool last = a.last != 0;
  const bool use_v0 = last && a.comb != 0;
  const bool cdt_on = d.ndim > 1 && a.stage == 1;
  const bool cdt_in = cdt_on && !first;
  const int qA = NV, q0 = qA + (first ? 0 : NV), qC = q0 + (use_v0 ? NV : 0);
  const int nq = qC + (cdt_in ? 1 : 0);
  double *ring = smem;                                  // [RING][nq][BX]
  double *exv = smem + RING * nq * BX;                  // FUSEX: [NV][BX] centre states
  double *exm = exv + NV * BX;                          //        [NV][BX] right states (vm)
  double *exf = exm + NV * BX;                          //        [NV+2][BX] fluxes, prs, cmax

  const int n0 = cb - 1;  // first zone to reconstruct
  auto issue = [&](int m) {  // prefetch the global data iteration m will consume
    if (m <= ce + 1) {
      double *slot = ring + ((m - n0) % RING) * nq * BX + t;
      const long oV = base + (long)(m + LEAD) * st;
#pragma unroll
      for (int c = 0; c < NV; c++) cp_async8(slot + c * BX, a.V + gvar<DIR>(c) * d.sv + oV);
      const int z = m - 1;
      if (z >= cb && own) {
        const long oz = base + (long)z * st;
        if (!first) {
#pragma unroll
          for (int c = 0; c < NV; c++) cp_async8(slot + (qA + c) * BX, a.acc + gvar<DIR>(c) * d.sv + oz);
        }
        if (use_v0) {
#pragma unroll
          for (int c = 0; c < NV; c++) cp_async8(slot + (q0 + c) * BX, a.V0 + gvar<DIR>(c) * d.sv + oz);
        }
        if (cdt_in) cp_async8(slot + qC * BX, a.cdt + oz);
      }
    }
    cp_async_commit();
  };

  double mach = 0.0, cdt_max = 0.0;
  double vm1[NV], v0[NV], vp1[NV], vp2[NV];
  double vpL[NV], vp[NV], vm[NV];
  double qm[NV];  // PPM: interface value at n-1/2
  Face<NV> Fm, Fp;

#pragma unroll
  for (int g = 0; g < DEPTH; g++) issue(n0 + g);
  if (RECON == RECON_PARABOLIC) {
    double a0[NV];
    load_zone<DIR, NV>(a.V, base + (long)(cb - 3) * st, d.sv, a0);
    load_zone<DIR, NV>(a.V, base + (long)(cb - 2) * st, d.sv, vm1);
    load_zone<DIR, NV>(a.V, base + (long)(cb - 1) * st, d.sv, v0);
    load_zone<DIR, NV>(a.V, base + (long)(cb)*st, d.sv, vp1);
#pragma unroll
    for (int nv = 0; nv < NV; nv++) qm[nv] = ppm4_iface(a0[nv], vm1[nv], v0[nv], vp1[nv]);
  } else {
    load_zone<DIR, NV>(a.V, base + (long)(cb - 2) * st, d.sv, vm1);
    load_zone<DIR, NV>(a.V, base + (long)(cb - 1) * st, d.sv, v0);
  }

  for (int n = n0; n <= ce + 1; n++) {
    // ---- data of this iteration has landed in the ring; refill the slot DEPTH ahead ----
    cp_async_wait<DEPTH - 1>();
    const double *slot = ring + ((n - n0) % RING) * nq * BX + t;
    double vin[NV];
#pragma unroll
    for (int c = 0; c < NV; c++) vin[c] = slot[c * BX];
    const int z = n - 1;  // zone finished by this iteration
    const bool fin = z >= cb;  // block-uniform
    double U[NV], v0z[NV], cin = 0.0;
    if (fin && own) {
      if (!first) {
#pragma unroll
        for (int c = 0; c < NV; c++) U[c] = slot[(qA + c) * BX];
      }
      if (use_v0) {
#pragma unroll
        for (int c = 0; c < NV; c++) v0z[c] = slot[(q0 + c) * BX];
      }
      if (cdt_in) cin = slot[qC * BX];
    }
    issue(n + DEPTH);

    // ---- reconstruct zone n along DIR ----
    if (RECON == RECON_PARABOLIC) {
      if (n > n0) {
#pragma unroll
        for (int nv = 0; nv < NV; nv++) vp1[nv] = vp2[nv];
      }
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        vp2[nv] = vin[nv];
        double q = ppm4_iface(vm1[nv], v0[nv], vp1[nv], vp2[nv]);
        vp[nv] = q;
        vm[nv] = qm[nv];
        qm[nv] = q;
        ppm_parabola(v0[nv], vp[nv], vm[nv], 2.0, 2.0);
      }
    } else if (RECON == RECON_LINEAR) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) vp1[nv] = vin[nv];
      plm_rt<NV>(a.limiter, vm1, v0, vp1, vp, vm);
    } else {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) { vp1[nv] = vin[nv]; vp[nv] = vm[nv] = v0[nv]; }
    }

    // ---- face n-1/2 along DIR ----
    if (n >= cb && own) riemann<NV, SOLVER>(vpL, vm, d.gas, Fp, mach);

    if (fin) {
      double cx = 0.0;
      if (FUSEX) {
        // ---- x1 sweep of row z: thread <-> zone, neighbours through shared memory ----
        double vx[NV];  // this zone in x1 (= global) component order
#pragma unroll
        for (int c = 0; c < NV; c++) vx[gvar<DIR>(c)] = vm1[c];
#pragma unroll
        for (int nv = 0; nv < NV; nv++) exv[nv * BX + t] = vx[nv];
        __syncthreads();
        const int tm = t > 0 ? t - 1 : 0, tp = t < BX - 1 ? t + 1 : BX - 1;
        double xp[NV], xm[NV];
        if (RECON == RECON_PARABOLIC) {
          const int tp2 = t < BX - 2 ? t + 2 : BX - 1;
#pragma unroll
          for (int nv = 0; nv < NV; nv++) {
            xp[nv] = ppm4_iface(exv[nv * BX + tm], vx[nv], exv[nv * BX + tp], exv[nv * BX + tp2]);
            exm[nv * BX + t] = xp[nv];   // interface value at t+1/2
          }
          __syncthreads();
#pragma unroll
          for (int nv = 0; nv < NV; nv++) {
            xm[nv] = exm[nv * BX + tm];
            ppm_parabola(vx[nv], xp[nv], xm[nv], 2.0, 2.0);
          }
          __syncthreads();
        } else if (RECON == RECON_LINEAR) {
          double m1[NV], p1[NV];
#pragma unroll
          for (int nv = 0; nv < NV; nv++) { m1[nv] = exv[nv * BX + tm]; p1[nv] = exv[nv * BX + tp]; }
          plm_rt<NV>(a.limiter, m1, vx, p1, xp, xm);
        } else {
#pragma unroll
          for (int nv = 0; nv < NV; nv++) xp[nv] = xm[nv] = vx[nv];
        }
#pragma unroll
        for (int nv = 0; nv < NV; nv++) exm[nv * BX + t] = xm[nv];
        __syncthreads();
        double xr[NV];
#pragma unroll
        for (int nv = 0; nv < NV; nv++) xr[nv] = exm[nv * BX + tp];
        Face<NV> Gp;
        double machx = 0.0;
        riemann<NV, SOLVER>(xp, xr, d.gas, Gp, machx);
        if (t >= LO - 1 && t < BX - HI && i >= d.beg[0] - 1 && i <= d.end[0]) mach = fmax(mach, machx);
#pragma unroll
        for (int nv = 0; nv < NV; nv++) exf[nv * BX + t] = Gp.f[nv];
        exf[NV * BX + t] = Gp.prs;
        exf[(NV + 1) * BX + t] = Gp.cmax;
        __syncthreads();
        if (own) {
          const double idx1 = d.inv_dx[0][i];
          const double dtdx1 = dt * idx1;
          double ux[NV];
          prim2cons<NV>(vx, ux, d.gas);
#pragma unroll
          for (int nv = 0; nv < NV; nv++) ux[nv] += -dtdx1 * (Gp.f[nv] - exf[nv * BX + tm]);
          ux[iVN] -= dtdx1 * (Gp.prs - exf[NV * BX + tm]);
          cx = 0.5 * (exf[(NV + 1) * BX + tm] + Gp.cmax) * idx1;
#pragma unroll
          for (int c = 0; c < NV; c++) U[c] = ux[gvar<DIR>(c)];   // back to DIR-local order
        }
      } else if (first && own) {
        prim2cons<NV>(vm1, U, d.gas);
      }

      if (own && n >= cb + 1) {
        // ---- finish zone z: add this direction, combine, cons->prim ----
        const double inv_dl = inv_dx[z];
        const double dtdx = dt * inv_dl;
        const long oz = base + (long)z * st;
#pragma unroll
        for (int nv = 0; nv < NV; nv++) U[nv] += -dtdx * (Fp.f[nv] - Fm.f[nv]);
        U[iVN] -= dtdx * (Fp.prs - Fm.prs);
        if (last) {
          if (use_v0) {
            double U0[NV];
            prim2cons<NV>(v0z, U0, d.gas);
            if (a.comb == 1) {
#pragma unroll
              for (int nv = 0; nv < NV; nv++) U[nv] = a.w0 * U0[nv] + a.wc * U[nv];
            } else {
              const double one_third = 1.0 / 3.0;
#pragma unroll
              for (int nv = 0; nv < NV; nv++) U[nv] = one_third * (U0[nv] + 2.0 * U[nv]);
            }
          }
          double vn[NV];
          int fl = cons2prim<NV>(U, vn, d.gas);
          if (fl) atomicAdd(a.red + 2, 1ull);
          store_zone<DIR, NV>(a.Vout, oz, d.sv, vn);
        } else {
          store_zone<DIR, NV>(a.acc, oz, d.sv, U);
        }
        if (cdt_on) {
          double c = 0.5 * (Fm.cmax + Fp.cmax) * inv_dl;
          if (FUSEX) c = cx + c;
          else if (!first) c = cin + c;
          if (last) cdt_max = fmax(cdt_max, c);
          else a.cdt[oz] = c;
        }
      }
    }
    if (n >= cb) Fm = Fp;
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      vpL[nv] = vp[nv];
      vm1[nv] = v0[nv];
      v0[nv] = vp1[nv];
    }
  }
  cp_async_wait<0>();
  block_reduce_max2(cdt_max, mach, a.stage == 1 && last, a.red);
}

// ------------------------------------------------------------------------------------
//  physical boundaries  (Src/boundary.c:228-459 dispatch, :617-767 fills, :503 FlipSign)
// ------------------------------------------------------------------------------------
enum BcType { BC_NONE = 0, BC_OUTFLOW = 1, BC_REFLECTIVE = 2, BC_AXISYMMETRIC = 3,
              BC_EQTSYMMETRIC = 4, BC_PERIODIC = 5, BC_USERDEF = 8, BC_NEIGHBOUR = 100 };

struct BcArgs {
  double *V;
  int side;       // 0..5 = X1_BEG, X1_END, X2_BEG, ...
  int type;
  int nvar;
  int nghost;
  double sign[16];
};

__global__ void bc_fill(Dev d, BcArgs b) {
  const int dir = b.side >> 1;
  const bool hi = b.side & 1;
  // extents of the ghost box: nghost layers along dir, full transverse range
  int ext[3] = {d.tot[0], d.tot[1], d.tot[2]};
  ext[dir] = b.nghost;
  long ntot = (long)ext[0] * ext[1] * ext[2];
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ntot) return;
  int c[3];
  c[0] = (int)(idx % ext[0]);
  c[1] = (int)((idx / ext[0]) % ext[1]);
  c[2] = (int)(idx / ((long)ext[0] * ext[1]));
  const int g = c[dir];
  const int nb = d.beg[dir], ne = d.end[dir], nx = ne - nb + 1;
  const int n = hi ? ne + 1 + g : nb - 1 - g;
  int src;
  if (b.type == BC_OUTFLOW) src = hi ? ne : nb;
  else if (b.type == BC_PERIODIC) src = hi ? n - nx : n + nx;
  else src = hi ? 2 * ne + 1 - n : 2 * nb - 1 - n;
  const bool flip = (b.type == BC_REFLECTIVE || b.type == BC_AXISYMMETRIC || b.type == BC_EQTSYMMETRIC);
  c[dir] = n;
  long dst_off = (long)c[2] * d.sk + (long)c[1] * d.sj + c[0];
  c[dir] = src;
  long src_off = (long)c[2] * d.sk + (long)c[1] * d.sj + c[0];
  for (int nv = 0; nv < b.nvar; nv++) {
    double q = b.V[nv * d.sv + src_off];
    if (flip) q = b.sign[nv] * q;
    b.V[nv * d.sv + dst_off] = q;
  }
}

}  // namespace pb
