// One quarter of the npoly instantiations of the stage-B (Gram) kernel; compiled
// four times with -DRVS_NP_GROUP=0..3 so that the translation units build in parallel.
#include <algorithm>

#include "gram_kernel.cuh"
#include "gram_mma.cuh"
#include "scan_mma.cuh"

#ifndef RVS_NP_GROUP
#error "compile with -DRVS_NP_GROUP=0..3"
#endif
#define RVS_CAT2(a, b) a##b
#define RVS_CAT(a, b) RVS_CAT2(a, b)

namespace rvs {
template <int NP>
static int launch_gram_np(const GramArgs &a, int K, cudaStream_t st) {
  const size_t smem = sizeof(double) * GR_DEPTH * GR_THREADS * gram_slot(NP);
  RVS_REQUIRE(a.npp == ((NP + 1) & ~1), RVS_E_ARG, "gram: basis rows of %d doubles, expected %d",
              a.npp, (NP + 1) & ~1);
  RVS_CUDA_OK(ensure_dyn_smem(gram_kernel<NP>, smem));
  gram_kernel<NP><<<K, GR_THREADS, smem, st>>>(a);
  RVS_LAUNCH_OK();
  return 0;
}

template <int NP>
static int launch_gram_mma_np(GramMmaArgsM m, int narm, cudaStream_t st) {
  constexpr int NT = NP <= 10 ? 2 : 1;
  constexpr int NI = 8 * NT;
  const GramMmaArgs &a = m.a[0];      // K and the basis layout are the same for every arm
  RVS_REQUIRE(a.npp == ((NP + 1) & ~1), RVS_E_ARG, "gram: basis rows of %d doubles, expected %d",
              a.npp, (NP + 1) & ~1);
  const int groups = (a.K + NI - 1) / NI;
  // pixel splits: a function of the item count only, so that an arm's partial sums are
  // taken in the same order whether it is launched alone or with the other arms
  const int KS = std::min(GM_MAX_KS, std::max(1, (512 + groups - 1) / groups));
  for (int i = 0; i < narm; i++) m.a[i].KS = KS;
  dim3 grid(groups, KS, narm);
  size_t tile_smem = sizeof(double) * GM_WARPS * GM_NSTG * gm_stage_doubles(a.npp, NI);
  tile_smem = std::max(tile_smem, sizeof(double) * GramTiles<NP>::ROWS * (NI + 1));  // s_red alias
  RVS_CUDA_OK(ensure_dyn_smem(gram_mma_kernel<NP, NT>, tile_smem));
  prof_begin(ST_GRAM, st);
  gram_mma_kernel<NP, NT><<<grid, GM_THREADS, tile_smem, st>>>(m);
  prof_end(ST_GRAM, st);
  RVS_LAUNCH_OK();
  prof_begin(ST_SOLVE, st);
  gram_solve_kernel<NP, NT><<<dim3((a.K + GM_WARPS - 1) / GM_WARPS, narm), GM_THREADS, 0, st>>>(m);
  prof_end(ST_SOLVE, st);
  RVS_LAUNCH_OK();
  prof_begin(ST_RESID, st);
  resid_mma_kernel<NP, NT><<<grid, GM_THREADS, 0, st>>>(m);
  prof_end(ST_RESID, st);
  RVS_LAUNCH_OK();
  return 0;
}

template <int NP>
static int launch_scan_mma_np(const ScanArgs &a, cudaStream_t st) {
  constexpr int NT = NP <= 10 ? 2 : 1;
  constexpr int NI = 8 * NT;
  RVS_REQUIRE(a.npp == ((NP + 1) & ~1), RVS_E_ARG, "scan: basis rows of %d doubles, expected %d",
              a.npp, (NP + 1) & ~1);
  const int by = (a.nv + NI - 1) / NI;
  RVS_REQUIRE((int64_t)by * a.K <= 0x7fffffffLL, RVS_E_LIMIT, "scan: %d items x %d trials", a.K,
              a.nv);
  ScanArgs b = a;
  b.nby = by;
  const unsigned grid = (unsigned)((int64_t)by * a.K);
  RVS_REQUIRE(!a.resol || a.resol_hw <= RS_HW, RVS_E_LIMIT,
              "scan: resolution matrix half-bandwidth %d > %d", a.resol_hw, RS_HW);
  if (a.resol)  // resampled template staged in shared memory
    chisq_scan_mma_kernel<NP, NT, true><<<grid, GM_THREADS, 0, st>>>(b);
  else
    chisq_scan_mma_kernel<NP, NT, false><<<grid, GM_THREADS, 0, st>>>(b);
  RVS_LAUNCH_OK();
  return 0;
}

int RVS_CAT(launch_scan_mma_group, RVS_NP_GROUP)(const ScanArgs &a, int npoly, cudaStream_t st) {
  switch (npoly) {
#define RVS_CASE(N) case N: return launch_scan_mma_np<N>(a, st);
#if RVS_NP_GROUP == 0
    RVS_CASE(1) RVS_CASE(2) RVS_CASE(3) RVS_CASE(4) RVS_CASE(5) RVS_CASE(6) RVS_CASE(7)
#elif RVS_NP_GROUP == 1
    RVS_CASE(8) RVS_CASE(9) RVS_CASE(10)
#elif RVS_NP_GROUP == 2
    RVS_CASE(11) RVS_CASE(12) RVS_CASE(13)
#else
    RVS_CASE(14) RVS_CASE(15) RVS_CASE(16)
#endif
#undef RVS_CASE
  }
  return RVS_E_ARG;
}

int RVS_CAT(launch_gram_mma_group, RVS_NP_GROUP)(const GramMmaArgsM &m, int narm, int npoly,
                                                 cudaStream_t st) {
  switch (npoly) {
#define RVS_CASE(N) case N: return launch_gram_mma_np<N>(m, narm, st);
#if RVS_NP_GROUP == 0
    RVS_CASE(1) RVS_CASE(2) RVS_CASE(3) RVS_CASE(4) RVS_CASE(5) RVS_CASE(6) RVS_CASE(7)
#elif RVS_NP_GROUP == 1
    RVS_CASE(8) RVS_CASE(9) RVS_CASE(10)
#elif RVS_NP_GROUP == 2
    RVS_CASE(11) RVS_CASE(12) RVS_CASE(13)
#else
    RVS_CASE(14) RVS_CASE(15) RVS_CASE(16)
#endif
#undef RVS_CASE
  }
  return RVS_E_ARG;
}

int RVS_CAT(launch_gram_group, RVS_NP_GROUP)(const GramArgs &a, int npoly, int K, cudaStream_t st) {
  switch (npoly) {
#define RVS_CASE(N) case N: return launch_gram_np<N>(a, K, st);
#if RVS_NP_GROUP == 0
    RVS_CASE(1) RVS_CASE(2) RVS_CASE(3) RVS_CASE(4) RVS_CASE(5) RVS_CASE(6) RVS_CASE(7)
#elif RVS_NP_GROUP == 1
    RVS_CASE(8) RVS_CASE(9) RVS_CASE(10)
#elif RVS_NP_GROUP == 2
    RVS_CASE(11) RVS_CASE(12) RVS_CASE(13)
#else
    RVS_CASE(14) RVS_CASE(15) RVS_CASE(16)
#endif
#undef RVS_CASE
  }
  return RVS_E_ARG;
}
}  // namespace rvs
