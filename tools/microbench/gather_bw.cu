// What HBM bandwidth can the template gather reach?  Each "item" reads a window of
// W floats from each of 16 randomly chosen rows of a [nrow][ld] fp32 grid
// (the access pattern of chunk_kernel's stage 1), accumulates a weighted sum in
// fp64 and writes one value per lane.  Variants differ in how a warp walks the
// window and how many loads a lane keeps in flight.
//   mode 0: 8 rows x 16 B in flight per lane, two bursts (chunk_kernel v3)
//   mode 1: 16 rows x 16 B in flight per lane
//   mode 2: row after row, 4 x 16 B of ONE row per lane (2 KB contiguous per warp)
//   mode 3: like 0 but fp32 accumulation (no cvt) -- instruction-count probe
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(128) gather_kernel(const float *grid, int64_t ld, const int *ids,
                                                      const int *start, int W, int nwarps,
                                                      double *out) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (wg >= nwarps) return;
  const int *id = ids + wg * 16;
  const int s0 = start[wg];
  double tot = 0;
  if (MODE == 2) {
    // W = 4 * 128 floats handled as 4 float4 per lane per row
    double acc[4][4] = {};
    for (int j = 0; j < 16; j++) {
      const float4 *row = reinterpret_cast<const float4 *>(grid + (int64_t)id[j] * ld + s0);
      float4 v[4];
#pragma unroll
      for (int it = 0; it < 4; it++) v[it] = (lane + 32 * it) * 4 < W ? ldg_stream(row + lane + 32 * it) : make_float4(0, 0, 0, 0);
      const double w = 0.0625 + j;
#pragma unroll
      for (int it = 0; it < 4; it++) {
        acc[it][0] = fma(w, (double)v[it].x, acc[it][0]);
        acc[it][1] = fma(w, (double)v[it].y, acc[it][1]);
        acc[it][2] = fma(w, (double)v[it].z, acc[it][2]);
        acc[it][3] = fma(w, (double)v[it].w, acc[it][3]);
      }
    }
    for (int it = 0; it < 4; it++) tot += acc[it][0] + acc[it][1] + acc[it][2] + acc[it][3];
  } else {
    for (int i0 = lane * 4; i0 < W; i0 += 128) {
      if (MODE == 3) {
        float acc[4] = {0, 0, 0, 0};
        float4 v[8];
        for (int h = 0; h < 2; h++) {
#pragma unroll
          for (int j = 0; j < 8; j++) v[j] = ldg_stream(reinterpret_cast<const float4 *>(grid + (int64_t)id[h * 8 + j] * ld + s0 + i0));
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float w = 0.0625f + j;
            acc[0] = fmaf(w, v[j].x, acc[0]); acc[1] = fmaf(w, v[j].y, acc[1]);
            acc[2] = fmaf(w, v[j].z, acc[2]); acc[3] = fmaf(w, v[j].w, acc[3]);
          }
        }
        tot += acc[0] + acc[1] + acc[2] + acc[3];
      } else {
        double acc[4] = {0, 0, 0, 0};
        constexpr int NB = MODE == 1 ? 16 : 8;
        float4 v[NB];
        for (int h = 0; h < 16 / NB; h++) {
#pragma unroll
          for (int j = 0; j < NB; j++) v[j] = ldg_stream(reinterpret_cast<const float4 *>(grid + (int64_t)id[h * NB + j] * ld + s0 + i0));
#pragma unroll
          for (int j = 0; j < NB; j++) {
            const double w = 0.0625 + j;
            acc[0] = fma(w, (double)v[j].x, acc[0]); acc[1] = fma(w, (double)v[j].y, acc[1]);
            acc[2] = fma(w, (double)v[j].z, acc[2]); acc[3] = fma(w, (double)v[j].w, acc[3]);
          }
        }
        tot += acc[0] + acc[1] + acc[2] + acc[3];
      }
    }
  }
  out[wg * 32 + lane] = tot;
}

template <int MODE>
static void run(const char *name, const float *grid, int64_t ld, const int *ids, const int *start,
                int W, int nwarps, double *out, int minb_hint) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = (nwarps + 3) / 4;
  for (int it = 0; it < 3; it++) gather_kernel<MODE><<<blocks, 128>>>(grid, ld, ids, start, W, nwarps, out);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int it = 0; it < reps; it++) gather_kernel<MODE><<<blocks, 128>>>(grid, ld, ids, start, W, nwarps, out);
  cudaEventRecord(e1);
  CHECK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)nwarps * 16 * W * 4;
  printf("%-44s W=%4d warps=%7d  %8.1f us  %7.1f GB/s\n", name, W, nwarps, 1e3 * ms / reps, bytes * reps / (ms * 1e-3) / 1e9);
  (void)minb_hint;
}

int main() {
  const int64_t nrow = 28600, ld = 6240;  // one DESI-b bank: 0.71 GB
  float *grid; CHECK(cudaMalloc(&grid, nrow * ld * 4));
  CHECK(cudaMemset(grid, 0, nrow * ld * 4));
  const int maxw = 1 << 18;
  int *h_ids = (int *)malloc(maxw * 16 * 4), *h_start = (int *)malloc(maxw * 4);
  srand(1);
  for (int W : {512, 2048}) {
    // items = 16 random rows; consecutive warps of an item take consecutive windows
    const int chunks = 6144 / W;
    const int nwarps = 1024 * chunks;  // 1024 items
    for (int it = 0; it < 1024; it++) {
      int rows[16];
      for (int j = 0; j < 16; j++) rows[j] = rand() % nrow;
      for (int c = 0; c < chunks; c++) {
        for (int j = 0; j < 16; j++) h_ids[(it * chunks + c) * 16 + j] = rows[j];
        h_start[it * chunks + c] = c * W;
      }
    }
    int *ids, *start; double *out;
    CHECK(cudaMalloc(&ids, nwarps * 16 * 4)); CHECK(cudaMalloc(&start, nwarps * 4));
    CHECK(cudaMalloc(&out, (size_t)nwarps * 32 * 8));
    CHECK(cudaMemcpy(ids, h_ids, nwarps * 16 * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(start, h_start, nwarps * 4, cudaMemcpyHostToDevice));
    run<0>("mode0: 8 rows x16B in flight, fp64 acc", grid, ld, ids, start, W, nwarps, out, 0);
    run<1>("mode1: 16 rows x16B in flight, fp64 acc", grid, ld, ids, start, W, nwarps, out, 0);
    if (W == 512) run<2>("mode2: row by row, 4x16B of one row, fp64", grid, ld, ids, start, W, nwarps, out, 0);
    run<3>("mode3: 8 rows x16B in flight, fp32 acc", grid, ld, ids, start, W, nwarps, out, 0);
    cudaFree(ids); cudaFree(start); cudaFree(out);
  }
  return 0;
}
