// probe_pow35.cu -- accuracy of pow_three_fifths() (csrc/hd_physics.cuh) against long-double powl over 1e-12 .. 1e12, next to
// exp(0.6 log x):  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_p35_probe tools/probe_pow35.cu
#include <cstdio>
#include <cmath>
#define PB_D __device__ __forceinline__
namespace t {
PB_D double pow_three_fifths(double x) {
  float lx, sd;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lx) : "f"((float)x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sd) : "f"(-0.6f * lx));
  const double c = x * x * x;
  double z = (double)sd;
  for (int it = 0; it < 2; it++) { const double z2 = z * z, z4 = z2 * z2; z = (z * 0.2) * fma(-c * z4, z, 6.0); }
  const double z2 = z * z;
  return c * (z2 * z2);
}
}
__global__ void k(const double *x, double *y, double *y2, int n) { int i = blockIdx.x*blockDim.x+threadIdx.x; if (i<n) { y[i] = t::pow_three_fifths(x[i]); y2[i] = exp(0.6*log(x[i])); } }
int main() {
  const int n = 1<<20; double *x, *y, *y2; cudaMallocManaged(&x, n*8); cudaMallocManaged(&y, n*8); cudaMallocManaged(&y2, n*8);
  for (int i = 0; i < n; i++) x[i] = pow(10.0, -12.0 + 24.0 * (i + 0.37) / n);
  k<<<n/256,256>>>(x,y,y2,n); cudaDeviceSynchronize();
  double e1=0,e2=0; for (int i=0;i<n;i++){ long double r = powl((long double)x[i], 0.6L); e1 = fmax(e1, fabs((double)((y[i]-r)/r))); e2 = fmax(e2, fabs((double)((y2[i]-r)/r))); }
  printf("max rel err: three_fifths %.3e  exp-log %.3e (eps %.3e)\n", e1, e2, 2.2e-16);
}
