// One quarter of the npoly instantiations of the stage-B (Gram) kernel; compiled
// four times with -DRVS_NP_GROUP=0..3 so that the translation units build in parallel.
#include "gram_kernel.cuh"

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
  RVS_CUDA_OK(cudaFuncSetAttribute(gram_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  gram_kernel<NP><<<K, GR_THREADS, smem, st>>>(a);
  RVS_LAUNCH_OK();
  return 0;
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
