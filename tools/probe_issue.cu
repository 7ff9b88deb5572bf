// probe_issue.cu -- how do FP64 instructions share the issue port with ALU / FSEL / LDS / IMAD work on
// sm_100a?  Each kernel runs R independent DFMA chains per thread and interleaves K non-FP64
// instructions per DFMA; the time per DFMA (in SM cycles per warp instruction and scheduler) tells
// whether the other instructions hide in the second cycle of the half-rate FP64 pipe (time stays at
// 2 cycles per DFMA) or cost issue slots of their own (2 + K).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_issue tools/probe_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int MODE, int K>
__global__ void __launch_bounds__(192) probe(double *out, double a, double b, int n, unsigned *sink) {
  extern __shared__ double sm[];
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  unsigned y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
  float f0 = y0, f1 = y1;
  double tl = 0;
  sm[threadIdx.x] = x0;
  __syncthreads();
  const unsigned sp = (unsigned)__cvta_generic_to_shared(sm + (threadIdx.x & 31));
#pragma unroll 1
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#define DF(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
#define OTHER(y, z)                                                                             \
  if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y) : "r"(z), "r"(it));      \
  if (MODE == 2) asm volatile("slct.f32.s32 %0, %0, %1, %2;" : "+f"(f0) : "f"(f1), "r"(y)); \
  if (MODE == 3) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(tl) : "r"(sp) : "memory"); \
  if (MODE == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y) : "r"(z), "r"(it));          \
  if (MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(y) : "r"(z));                          \
  if (MODE == 6) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(y));
      DF(x0) if (K >= 1) { OTHER(y0, y1) } if (K >= 2) { OTHER(y1, y2) } if (K >= 3) { OTHER(y2, y3) } if (K >= 4) { OTHER(y3, y0) }
      DF(x1) if (K >= 1) { OTHER(y2, y3) } if (K >= 2) { OTHER(y3, y0) } if (K >= 3) { OTHER(y0, y1) } if (K >= 4) { OTHER(y1, y2) }
      DF(x2) if (K >= 1) { OTHER(y0, y1) } if (K >= 2) { OTHER(y1, y2) } if (K >= 3) { OTHER(y2, y3) } if (K >= 4) { OTHER(y3, y0) }
      DF(x3) if (K >= 1) { OTHER(y2, y3) } if (K >= 2) { OTHER(y3, y0) } if (K >= 3) { OTHER(y0, y1) } if (K >= 4) { OTHER(y1, y2) }
      DF(x4) if (K >= 1) { OTHER(y0, y1) } if (K >= 2) { OTHER(y1, y2) } if (K >= 3) { OTHER(y2, y3) } if (K >= 4) { OTHER(y3, y0) }
      DF(x5) if (K >= 1) { OTHER(y2, y3) } if (K >= 2) { OTHER(y3, y0) } if (K >= 3) { OTHER(y0, y1) } if (K >= 4) { OTHER(y1, y2) }
      DF(x6) if (K >= 1) { OTHER(y0, y1) } if (K >= 2) { OTHER(y1, y2) } if (K >= 3) { OTHER(y2, y3) } if (K >= 4) { OTHER(y3, y0) }
      DF(x7) if (K >= 1) { OTHER(y2, y3) } if (K >= 2) { OTHER(y3, y0) } if (K >= 3) { OTHER(y0, y1) } if (K >= 4) { OTHER(y1, y2) }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + f0 + tl;
  if ((y0 ^ y1 ^ y2 ^ y3) == 0x12345) *sink = 1;
}

// pure non-FP64 kernel: throughput of the other instruction alone
template <int MODE>
__global__ void __launch_bounds__(192) other_only(double *out, int n, unsigned *sink) {
  extern __shared__ double sm[];
  unsigned y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
  float f0 = y0, f1 = y1;
  double tl = 0;
  int it = 0;
  sm[threadIdx.x] = y0;
  __syncthreads();
  const unsigned sp = (unsigned)__cvta_generic_to_shared(sm + (threadIdx.x & 31));
#pragma unroll 1
  for (it = 0; it < n; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      OTHER(y0, y1) OTHER(y1, y2) OTHER(y2, y3) OTHER(y3, y0)
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + tl;
  if ((y0 ^ y1 ^ y2 ^ y3) == 0x12345) *sink = 1;
}

template <typename F>
static float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int nsm = p.multiProcessorCount;
  double *out;
  unsigned *sink;
  cudaMalloc(&out, sizeof(double) * nsm * 4 * 256 * 8);
  cudaMalloc(&sink, 4);
  printf("%s, %d SMs, %.0f MHz nominal\n", p.name, nsm, khz / 1000.0);
  const char *names[] = {"none", "LOP3", "FSEL", "LDS.64", "IMAD", "IADD", "SHFL"};
  for (int bps = 1; bps <= 4; bps++) {   // blocks of 192 threads per SM: 6, 12, 18, 24 warps per SM
    const int nb = nsm * bps;
    printf("--- %d warps per SM (%.1f per scheduler)\n", bps * 6, bps * 1.5);
    auto report = [&](const char *nm, int k, float ms) {
      // DFMA warp instructions per scheduler: 6 warps*bps/4 schedulers * ITER*32
      double dfma = bps * 6 / 4.0 * ITER * 32.0;
      double cyc = ms * 1e-3 * 1.965e9;   // assumes boost clock; relative numbers are what matter
      printf("  DFMA + %d %-6s : %8.3f ms  %.2f cycles per DFMA (at 1965 MHz)\n", k, nm, ms, cyc / dfma);
    };
#define RUN(M, K) report(names[M], K, timeit([&] { probe<M, K><<<nb, 192, 2048>>>(out, 1.0000001, 1e-9, ITER, sink); }));
    RUN(0, 0)
    RUN(1, 1) RUN(1, 2) RUN(1, 3) RUN(1, 4)
    RUN(2, 1) RUN(2, 2) RUN(2, 4)
    RUN(3, 1) RUN(3, 2)
    RUN(4, 1) RUN(4, 2) RUN(4, 4)
    RUN(5, 1) RUN(5, 2)
    RUN(6, 1) RUN(6, 2)
    auto rep2 = [&](const char *nm, float ms) {
      double n = bps * 6 / 4.0 * ITER * 64.0;
      printf("  %-6s alone   : %8.3f ms  %.2f cycles per instruction\n", nm, ms, ms * 1e-3 * 1.965e9 / n);
    };
#define RUN2(M) rep2(names[M], timeit([&] { other_only<M><<<nb, 192, 2048>>>(out, ITER, sink); }));
    RUN2(1) RUN2(2) RUN2(3) RUN2(4) RUN2(5) RUN2(6)
  }
  // dependent-issue latency of DFMA: one chain
  return 0;
}
