// Microbenchmarks that decide the fused-kernel design on B200 (sm_100a):
//  (1) DFMA throughput / dependent-issue latency
//  (2) F2F.F64.F32 (float->double) throughput
//  (3) DMMA m8n8k4 f64 throughput, alone and together with DFMA (shared pipe?)
//  (4) fp64 exp() cost
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096

__global__ void k_dfma(double *out, int ilp_sel) {
  double a[8], b = 1.0000001, c = 0.5;
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma_dep(double *out) {  // single dependent chain
  double a = threadIdx.x, b = 1.0000001, c = 0.5;
  for (int it = 0; it < ITERS * 8; it++) a = fma(a, b, c);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

__global__ void k_f2f(double *out, const float *in) {
  float f[8];
  for (int i = 0; i < 8; i++) f[i] = in[threadIdx.x + i];
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      // conversion whose input changes every iteration so that it cannot be hoisted
      f[i] = __int_as_float(__float_as_int(f[i]) ^ (it & 1));
      s[i] += (double)f[i];
    }
  }
  double t = 0;
  for (int i = 0; i < 8; i++) t += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__global__ void k_f2f_int(double *out, const float *in) {  // integer re-encoding instead of F2F
  float f[8];
  for (int i = 0; i < 8; i++) f[i] = in[threadIdx.x + i];
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      unsigned u = __float_as_uint(f[i]) ^ (it & 1);
      f[i] = __uint_as_float(u);
      unsigned hi = ((u >> 3) & 0x0fffffffu) + 0x38000000u + (u & 0x80000000u);
      unsigned lo = u << 29;
      s[i] += __hiloint2double(hi, lo);
    }
  }
  double t = 0;
  for (int i = 0; i < 8; i++) t += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double *out, int with_dfma) {
  double c[8][2];
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  double x[4] = {1, 2, 3, 4};
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
    if (with_dfma) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = fma(x[i], b, a);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  for (int i = 0; i < 4; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double *out) {
  double a[4];
  for (int i = 0; i < 4; i++) a[i] = -1e-3 * (threadIdx.x + i);
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = exp(a[i]) - 1.0001;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double ghz = p.clockRate * 1e-6;
  printf("%s: %d SMs, %.3f GHz nominal\n", p.name, sms, ghz);
  const int T = 512, B = sms * 4;
  double *out; float *in;
  cudaMalloc(&out, sizeof(double) * T * B);
  cudaMalloc(&in, sizeof(float) * (T + 8));
  cudaMemset(in, 0x3f, sizeof(float) * (T + 8));
  double n = (double)T * B * ITERS;
  float ms;
  ms = timeit([&] { k_dfma<<<B, T>>>(out, 0); });
  printf("DFMA ilp8      : %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", ms, 2 * n * 8 / ms / 1e9,
         n * 8 / (ms * 1e-3) / sms / (ghz * 1e9));
  ms = timeit([&] { k_dfma_dep<<<sms, 32>>>(out); });
  printf("DFMA dependent : %8.3f ms  -> %.2f clk per dependent FMA (1 warp/SM, nominal clk)\n", ms,
         ms * 1e-3 * ghz * 1e9 / (ITERS * 8.0));
  ms = timeit([&] { k_f2f<<<B, T>>>(out, in); });
  printf("F2F+DADD ilp8  : %8.3f ms  %6.1f conv/clk/SM (upper bound incl. DADD+LOP)\n", ms,
         n * 8 / (ms * 1e-3) / sms / (ghz * 1e9));
  ms = timeit([&] { k_f2f_int<<<B, T>>>(out, in); });
  printf("int-conv+DADD  : %8.3f ms  %6.1f conv/clk/SM\n", ms, n * 8 / (ms * 1e-3) / sms / (ghz * 1e9));
  ms = timeit([&] { k_dmma<<<B, T>>>(out, 0); });
  double dm = (double)(T / 32) * B * ITERS * 8;
  printf("DMMA only      : %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", ms, 2 * dm * 256 / ms / 1e9,
         dm * 256 / (ms * 1e-3) / sms / (ghz * 1e9));
  float ms2 = timeit([&] { k_dmma<<<B, T>>>(out, 1); });
  printf("DMMA + 16 DFMA : %8.3f ms  (DFMA alone would take %.3f ms)\n", ms2,
         (double)T * B * ITERS * 16 / (n * 8) * timeit([&] { k_dfma<<<B, T>>>(out, 0); }));
  ms = timeit([&] { k_exp<<<B, T>>>(out); });
  printf("exp(f64)       : %8.3f ms  %6.2f exp/clk/SM  (= %.1f DFMA-equivalents each)\n", ms,
         n / (ms * 1e-3) / sms / (ghz * 1e9), 64.0 / (n / (ms * 1e-3) / sms / (ghz * 1e9)));
  return 0;
}
