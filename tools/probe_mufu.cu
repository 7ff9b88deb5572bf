// probe_mufu.cu -- measures the accuracy of MUFU.RCP64H / MUFU.RSQ64H seeds and of the
// branch-free Newton refinements used in hd_physics.cuh (run on the GPU box).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rsq_seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ double atomicMaxD(double *a, double v) {
  unsigned long long *p = (unsigned long long *)a, old = *p, assumed;
  do { assumed = old; if (__longlong_as_double(assumed) >= v) break; old = atomicCAS(p, assumed, __double_as_longlong(v)); } while (assumed != old);
  return __longlong_as_double(old);
}
__global__ void probe(double *out, long n) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  double e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (; i < n; i += (long)gridDim.x * blockDim.x) {
    // x sweeps [1,4) finely plus a wide exponent range
    double x = 1.0 + 3.0 * ((double)i / (double)n);
    double sc = ldexp(1.0, (int)(i % 41) * 10 - 200);
    double xs = x * sc;
    double y0 = rcp_seed(xs), ex = 1.0 / xs;
    e[0] = fmax(e[0], fabs(y0 - ex) / ex);
    double q = fma(-xs, y0, 1.0);
    double y1 = fma(y0, q, y0);                 // quadratic
    e[1] = fmax(e[1], fabs(y1 - ex) / ex);
    double y1c = fma(y0, fma(q, q, q), y0);     // cubic
    e[2] = fmax(e[2], fabs(y1c - ex) / ex);
    double q2 = fma(-xs, y1c, 1.0);
    double y2 = fma(y1c, q2, y1c);              // cubic + quadratic
    e[3] = fmax(e[3], fabs(y2 - ex) / ex);
    double r0 = rsq_seed(xs), er = 1.0 / sqrt(xs);
    e[4] = fmax(e[4], fabs(r0 - er) / er);
    double t = xs * r0, ee = fma(-t, r0, 1.0);
    double r1 = fma(fma(ee, 0.375, 0.5), ee * r0, r0);   // cubic
    e[5] = fmax(e[5], fabs(r1 - er) / er);
    double s = xs * r1;                                   // sqrt
    double s2 = fma(fma(-s, s, xs), 0.5 * r1, s);         // one correction
    double es = sqrt(xs);
    e[6] = fmax(e[6], fabs(s - es) / es);
    e[7] = fmax(e[7], fabs(s2 - es) / es);
  }
  for (int k = 0; k < 8; k++) atomicMaxD(out + k, e[k]);
}
int main() {
  double *d, h[8];
  cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  probe<<<148 * 8, 256>>>(d, 1L << 30);
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const char *nm[8] = {"rcp seed", "rcp quad", "rcp cubic", "rcp cubic+quad", "rsq seed", "rsq cubic", "sqrt = x*rsq", "sqrt corrected"};
  for (int k = 0; k < 8; k++) printf("%-16s max rel err %.3e (2^%.1f)\n", nm[k], h[k], log2(h[k]));
  return 0;
}
