// Library-level plumbing (error text, launch counter) and the host-buffer
// spline entry points that mirror the reference's cffi module `_spliner`
// (reference ffibuilder.py:10-17, src/spliner.c:7-108).
#include <math.h>
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include <cuda.h>
#include <string.h>

#include "common.cuh"

#ifndef RVS_TMA_COLS
#define RVS_TMA_COLS 64
#endif
#define RVS_TMA_COLS_HOST RVS_TMA_COLS
#ifndef RVS_TMA_ROWS
#define RVS_TMA_ROWS 8
#endif
#define RVS_TMA_ROWS_HOST RVS_TMA_ROWS

namespace rvs {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

cudaError_t ensure_dyn_smem_ptr(const void *kern, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> cur;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  size_t &have = cur[std::make_pair(dev, kern)];
  if (smem > have) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    have = smem;
  }
  return cudaSuccess;
}

constexpr int SPL_HALO = 32;
constexpr int SPL_CHUNK = 33;  // rows per thread (odd)

// forward elimination of rows [k0,k1) of each thread, warm-up from k0-HALO
__global__ void spline_forward_kernel(const double *y, const double *h, const double *hinv,
                                      const double *winv, int n, double *d) {
  const int m = n - 2;
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * SPL_CHUNK;
  if (k0 >= m) return;
  const int k1 = min(m, k0 + SPL_CHUNK);
  const int ks = max(0, k0 - SPL_HALO);
  double dd = 0;
  double bl = (y[ks + 1] - y[ks]) * hinv[ks];
  for (int k = ks; k < k1; k++) {
    const double br = (y[k + 2] - y[k + 1]) * hinv[k + 1];
    dd = (6 * (br - bl) - h[k] * dd) * winv[k];
    if (k >= k0) d[k] = dd;
    bl = br;
  }
}

__global__ void spline_backward_kernel(const double *y, const double *h, const double *hinv,
                                       const double *cp, const double *d, int n, double *z) {
  const int m = n - 2;
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * SPL_CHUNK;
  if (blockIdx.x == 0 && threadIdx.x == 0) { z[0] = 0; z[n - 1] = 0; }
  if (k0 >= m) return;
  const int k1 = min(m, k0 + SPL_CHUNK);
  const int ke = min(m, k1 + SPL_HALO);
  double zz = 0;
  for (int k = ke - 1; k >= k0; k--) {
    zz = d[k] - cp[k] * zz;
    if (k < k1) z[k + 1] = zz;
  }
}

__global__ void spline_coeff_kernel(const double *y, const double *z, const double *h,
                                    const double *hinv, int n, double *A, double *B, double *C,
                                    double *D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const double t1 = hinv[i] * (1. / 6), t2 = h[i] * (1. / 6);
  A[i] = z[i + 1] * t1;
  B[i] = z[i] * t1;
  C[i] = y[i + 1] * hinv[i] - z[i + 1] * t2;
  D[i] = y[i] * hinv[i] - z[i] * t2;
}

__global__ void spline_eval_kernel(const double *ex, int nex, int n, const double *x,
                                   const double *A, const double *B, const double *C,
                                   const double *D, int log_step, double q0, double qstep,
                                   double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nex) return;
  const double e = ex[i];
  int p = (int)(((log_step ? log(e) : e) - q0) / qstep);
  p = max(0, min(p, n - 2));
  const double dl = e - x[p], dr = x[p + 1] - e;
  out[i] = A[p] * dl * dl * dl + B[p] * dr * dr * dr + C[p] * dl + D[p] * dr;
}

struct DevBuf {
  double *p = nullptr;
  explicit DevBuf(size_t n) { if (cudaMalloc(&p, n * sizeof(double)) != cudaSuccess) p = nullptr; }
  ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace rvs

extern "C" const char *rvs_last_error(void) { return rvs::g_err; }
extern "C" int rvs_version(void) { return 100; }
extern "C" int64_t rvs_launch_count(void) { return rvs::g_launches.load(); }

namespace rvs {
struct ProfRec { int stage; cudaEvent_t e0, e1; };
static bool g_prof = false;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_rec;
static cudaEvent_t g_prof_open[ST_COUNT][8];   // begin events by (stage, stream slot)
static cudaStream_t g_prof_stream[ST_COUNT][8];
bool prof_on() { return g_prof; }
void prof_begin(int stage, cudaStream_t st) {
  if (!g_prof) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < 8; i++)
    if (!g_prof_open[stage][i]) { g_prof_open[stage][i] = e; g_prof_stream[stage][i] = st; return; }
  cudaEventDestroy(e);
}
void prof_end(int stage, cudaStream_t st) {
  if (!g_prof) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < 8; i++)
    if (g_prof_open[stage][i] && g_prof_stream[stage][i] == st) {
      cudaEvent_t e1;
      if (cudaEventCreate(&e1) == cudaSuccess) {
        cudaEventRecord(e1, st);
        g_prof_rec.push_back({stage, g_prof_open[stage][i], e1});
      }
      g_prof_open[stage][i] = nullptr;
      return;
    }
}
}  // namespace rvs

extern "C" int rvs_profile_timeline(double *out, int max_records) {
  using namespace rvs;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto &r : g_prof_rec) {
    if (n < max_records && !g_prof_rec.empty()) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, g_prof_rec[0].e0, r.e0);
      cudaEventElapsedTime(&b, g_prof_rec[0].e0, r.e1);
      out[3 * n] = r.stage; out[3 * n + 1] = a; out[3 * n + 2] = b;
      n++;
    }
  }
  for (auto &r : g_prof_rec) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof_rec.clear();
  return n;
}

extern "C" int rvs_gridbox_init(rvs_gridbox *box, const void *d_grid, int64_t ld, int ndim,
                                const int32_t *len) {
  using namespace rvs;
  RVS_REQUIRE(box && d_grid && len, RVS_E_ARG, "rvs_gridbox_init: null pointer");
  RVS_REQUIRE(ndim == 4, RVS_E_ARG, "rvs_gridbox_init: ndim=%d (the box gather is 4-D)", ndim);
  RVS_REQUIRE(ld > 0 && ld % 4 == 0 && ((uintptr_t)d_grid & 15) == 0, RVS_E_ARG,
              "rvs_gridbox_init: rows must be 16-byte aligned");
  for (int i = 0; i < 4; i++)
    RVS_REQUIRE(len[i] >= 2, RVS_E_ARG, "rvs_gridbox_init: len[%d]=%d", i, len[i]);
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                               const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  RVS_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  RVS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, RVS_E_CUDA,
              "rvs_gridbox_init: cuTensorMapEncodeTiled not available");
  const cuuint64_t row = (cuuint64_t)ld * 4;
  const cuuint64_t dims[5] = {(cuuint64_t)ld, (cuuint64_t)len[3], (cuuint64_t)len[2],
                              (cuuint64_t)len[1], (cuuint64_t)len[0]};
  const cuuint64_t strides[4] = {row, row * len[3], row * len[3] * len[2],
                                 row * len[3] * len[2] * len[1]};
  const cuuint32_t boxdim[5] = {(cuuint32_t)RVS_TMA_COLS_HOST, 2, 2, RVS_TMA_ROWS_HOST / 4, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  alignas(64) CUtensorMap tm;
  const CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                                    const_cast<void *>(d_grid), dims, strides, boxdim, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RVS_REQUIRE(r == CUDA_SUCCESS, RVS_E_CUDA, "rvs_gridbox_init: cuTensorMapEncodeTiled -> %d",
              (int)r);
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  memcpy(box->tmap, &tm, 128);
  for (int i = 0; i < 4; i++) box->len[i] = len[i];
  box->cols = RVS_TMA_COLS_HOST;
  box->rows = RVS_TMA_ROWS_HOST;
  return 0;
}

extern "C" void rvs_profile_enable(int on) { rvs::g_prof = on != 0; }
extern "C" int rvs_profile_active(void) { return rvs::g_prof ? 1 : 0; }
extern "C" int rvs_profile_read(double *ms_total, int64_t *launches, int nstage) {
  using namespace rvs;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < nstage; i++) { ms_total[i] = 0; launches[i] = 0; }
  for (auto &r : g_prof_rec) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (r.stage < nstage) { ms_total[r.stage] += ms; launches[r.stage]++; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof_rec.clear();
  return ST_COUNT;
}

extern "C" void rvs_spline_construct(double *xs, double *ys, int N, double *A, double *B,
                                     double *C, double *D, double *hout) {
  using namespace rvs;
  if (N < 3) return;
  std::vector<double> h(N - 1), hinv(N - 1), cp(N - 2), winv(N - 2);
  rvs_knot_tables(xs, N, h.data(), hinv.data(), cp.data(), winv.data());
  // device layout: y | h | hinv | cp | winv | d | z | A | B | C | D
  const size_t n = N;
  DevBuf buf(11 * n);
  if (!buf.p) { set_error("rvs_spline_construct: cudaMalloc failed"); return; }
  double *dy = buf.p, *dh = dy + n, *dhi = dh + n, *dcp = dhi + n, *dw = dcp + n, *dd = dw + n,
         *dz = dd + n, *dA = dz + n, *dB = dA + n, *dC = dB + n, *dD = dC + n;
  cudaMemcpy(dy, ys, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dh, h.data(), (n - 1) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dhi, hinv.data(), (n - 1) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dcp, cp.data(), (n - 2) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, winv.data(), (n - 2) * 8, cudaMemcpyHostToDevice);
  const int nthreads = (N - 2 + SPL_CHUNK - 1) / SPL_CHUNK;
  const int nb = (nthreads + 127) / 128;
  spline_forward_kernel<<<nb, 128>>>(dy, dh, dhi, dw, N, dd);
  spline_backward_kernel<<<nb, 128>>>(dy, dh, dhi, dcp, dd, N, dz);
  spline_coeff_kernel<<<(N + 255) / 256, 256>>>(dy, dz, dh, dhi, N, dA, dB, dC, dD);
  count_launch(3);
  cudaMemcpy(A, dA, (n - 1) * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(B, dB, (n - 1) * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(C, dC, (n - 1) * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(D, dD, (n - 1) * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < N - 1; i++) hout[i] = h[i];
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) set_error("rvs_spline_construct: %s", cudaGetErrorString(e));
}

extern "C" int rvs_spline_eval(double *evalx, int nevalx, int N, double *xs, double *hs,
                               double *As, double *Bs, double *Cs, double *Ds, int log_step,
                               double *ret) {
  using namespace rvs;
  (void)hs;
  if (nevalx <= 0) return 0;
  // status codes of the reference evaler (spliner.c:76-96)
  const double x0 = xs[0], xlast = xs[N - 1];
  if (evalx[0] < x0 || evalx[nevalx - 1] < x0) return -1;
  if (evalx[0] >= xlast || evalx[nevalx - 1] >= xlast) return -1;
  rvs_knots kn;
  if (rvs_knot_info(xs, N, log_step, &kn) != 0) return -2;
  const size_t n = N, ne = nevalx;
  DevBuf buf(5 * n + 2 * ne);
  if (!buf.p) { set_error("rvs_spline_eval: cudaMalloc failed"); return RVS_E_CUDA; }
  double *dx = buf.p, *dA = dx + n, *dB = dA + n, *dC = dB + n, *dD = dC + n, *de = dD + n,
         *dout = de + ne;
  cudaMemcpy(dx, xs, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dA, As, (n - 1) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bs, (n - 1) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, Cs, (n - 1) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, Ds, (n - 1) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(de, evalx, ne * 8, cudaMemcpyHostToDevice);
  const double qstep = log_step ? log(xs[1] / x0) : xs[1] - x0;
  spline_eval_kernel<<<(nevalx + 255) / 256, 256>>>(de, nevalx, N, dx, dA, dB, dC, dD, log_step,
                                                    kn.q0, qstep, dout);
  count_launch();
  cudaMemcpy(ret, dout, ne * 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("rvs_spline_eval: %s", cudaGetErrorString(e)); return RVS_E_CUDA; }
  return 0;
}
