// probe_xu.cu -- throughput of the XU-pipe instructions the FP64 kernels use on sm_100a: 64-bit reciprocal / rsqrt
// seeds (MUFU.RCP64H / RSQ64H), 32-bit MUFU, and the conversions between 64-bit and 32-bit types.  Eight independent
// chains per thread, 16 warps per SM: cycles per warp instruction and scheduler.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_xu tools/probe_xu.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) probe(double *out, int n, double seed) {
  double x[8];
  float f[8];
  int q[8];
#pragma unroll
  for (int c = 0; c < 8; c++) { x[c] = seed + threadIdx.x + c; f[c] = (float)x[c]; q[c] = threadIdx.x + c; }
#pragma unroll 1
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (MODE == 0) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x[c]));
        if (MODE == 1) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x[c]));
        if (MODE == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
        if (MODE == 3) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
        if (MODE == 4) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(x[c])); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x[c]) : "f"(f[c])); }
        if (MODE == 5) { asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(x[c]) : "r"(q[c])); asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(q[c]) : "d"(x[c])); }
        if (MODE == 6) asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(x[c]));
        if (MODE == 7) asm volatile("cvt.rzi.f64.f64 %0, %0;" : "+d"(x[c]));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) s += x[c] + f[c] + q[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int per_iter, double *out) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int n = 2000, blocks = sms * 4;      // 4 blocks of 4 warps per SM: 4 warps per scheduler
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<blocks, 128>>>(out, 10, 1.5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 128>>>(out, n, 1.5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * khz * 1e3;
  const double winst_per_sched = (double)n * 4 * 8 * per_iter * 4;   // 4 warps per scheduler
  printf("%-34s %7.2f cycles per warp instruction and scheduler (%.3f ms, clock attr %d MHz)\n", name, cycles / winst_per_sched, ms, khz / 1000);
}

int main() {
  double *out;
  cudaMalloc(&out, 148 * 8 * 128 * sizeof(double));
  run<6>("DFMA", 1, out);
  run<0>("MUFU.RCP64H (rcp.approx.f64)", 1, out);
  run<1>("MUFU.RSQ64H (rsqrt.approx.f64)", 1, out);
  run<2>("MUFU.RCP (f32)", 1, out);
  run<3>("MUFU.LG2 (f32)", 1, out);
  run<4>("F2F f64->f32 + f32->f64 (pair)", 2, out);
  run<5>("I2F.F64.S32 + F2I.S32.F64 (pair)", 2, out);
  run<7>("FRND.F64 (cvt.rzi.f64.f64)", 1, out);
  return 0;
}
