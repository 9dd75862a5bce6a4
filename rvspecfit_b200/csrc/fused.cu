// C-ABI entry point of the fused optimiser-phase evaluation: per-item preparation
// (prep_kernel: rotation taps, chunk geometry), stage A (chunk_kernel.cuh: template
// window -> T/sigma) and stage B (gram_mma.cuh / gram_kernel.cuh: continuum solve ->
// chi-square).  The spline never goes to HBM; the only intermediate is T/sigma
// (8 bytes per observed pixel, L2-resident between the two stages).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "chunk_kernel.cuh"
#include "gram_kernel.cuh"
#include "gram_mma.cuh"

namespace rvs {
int launch_gram_group0(const GramArgs &, int, int, cudaStream_t);
int launch_gram_group1(const GramArgs &, int, int, cudaStream_t);
int launch_gram_group2(const GramArgs &, int, int, cudaStream_t);
int launch_gram_group3(const GramArgs &, int, int, cudaStream_t);
int launch_gram_mma_group0(const GramMmaArgsM &, int, int, cudaStream_t);
int launch_gram_mma_group1(const GramMmaArgsM &, int, int, cudaStream_t);
int launch_gram_mma_group2(const GramMmaArgsM &, int, int, cudaStream_t);
int launch_gram_mma_group3(const GramMmaArgsM &, int, int, cudaStream_t);

template <typename GT, int NV, bool TMA = false>
static int launch_chunk(const ChunkArgsM &m, int narm, size_t smem, cudaStream_t st) {
  auto kern = chunk_kernel<GT, NV, TMA>;
  // raised once per size: keeps the call out of CUDA-graph captures after the first,
  // uncaptured evaluation
  RVS_CUDA_OK(ensure_dyn_smem(kern, smem));
  int64_t blocks = 0;
  for (int i = 0; i < narm; i++) {
    const int64_t warps = (int64_t)m.a[i].K * m.a[i].nch;
    blocks = std::max(blocks, (warps + CK_WARPS - 1) / CK_WARPS);
  }
  RVS_REQUIRE(blocks <= 0x7fffffffLL, RVS_E_LIMIT, "rvs_chisq_fused: %lld CTAs", (long long)blocks);
  prof_begin(ST_CHUNK, st);
  kern<<<dim3((unsigned)blocks, narm), CK_THREADS, smem, st>>>(m);
  prof_end(ST_CHUNK, st);
  RVS_LAUNCH_OK();
  return 0;
}

// knots per chunk: long enough that the 2 x (SPL_HALO + taps) halo stays a small
// fraction, short enough that two window buffers per warp leave room for >= 16
// resident warps per SM
// The small latency-bound kernels of an evaluation (preparation, continuum solve)
// run on a HIGH-PRIORITY helper stream forked from / joined to the caller's
// stream.  A chunk_kernel grid keeps every SM full for ~100 us, and the block
// scheduler serves pending grids oldest first: without priority the 13-40 us
// solve kernels of one arm wait for slots behind the chunk grids of the other
// arms (and of the next evaluation) and take 100-160 us each.  With it the slots
// that finishing chunk CTAs free go to them first.  One helper stream per caller
// stream, created on first use (RVS_NO_AUX=1 in the environment disables it).
struct AuxStream {
  cudaStream_t aux = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};
static std::mutex g_aux_mu;
static std::map<std::pair<int, cudaStream_t>, AuxStream> g_aux;
static AuxStream *aux_for(cudaStream_t st) {
  static const bool off = getenv("RVS_NO_AUX") != nullptr;
  if (off) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(g_aux_mu);
  auto key = std::make_pair(dev, st);
  auto it = g_aux.find(key);
  if (it != g_aux.end()) return &it->second;
  AuxStream a;
  int lo = 0, hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithPriority(&a.aux, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
  for (auto &e : a.ev)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  return &(g_aux[key] = a);
}
// after this, work enqueued on `to` waits for everything enqueued on `from` so far
static void hand_over(cudaStream_t from, cudaStream_t to, cudaEvent_t ev) {
  if (from == to) return;
  cudaEventRecord(ev, from);
  cudaStreamWaitEvent(to, ev, 0);
}

#ifndef RVS_CK_C
#define RVS_CK_C 384
#endif
static int chunk_knots(int tapcap) { return tapcap > 48 ? 512 : RVS_CK_C; }
}  // namespace rvs

extern "C" int rvs_fused_chunks(int npix_t, int tapcap) {
  const int C = rvs::chunk_knots(tapcap);
  return (npix_t + C - 1) / C;
}

namespace rvs {
// Layout of d_work (doubles): taps [K][tapcap+1] | rec [K][2] | irec (int32) [K][4] |
// pbound (int32) [K][nch+1] | scratch of the Gram GEMM kernels.  Every section
// starts 16-byte aligned.
struct FusedWork {
  int64_t taps, rec, irec, pbound, gram, total;
  FusedWork(int64_t K, int tapcap, int nch) {
    auto even = [](int64_t v) { return (v + 1) & ~int64_t(1); };
    taps = 0;
    rec = taps + even(K * (tapcap + 1));
    irec = rec + 2 * K;
    pbound = irec + 4 * K;
    gram = pbound + even((K * (nch + 1) + 1) / 2);
    total = gram + GramScratch(K).total();
  }
};
}  // namespace rvs

extern "C" int64_t rvs_fused_workspace(int K, int tapcap, int npix_t) {
  return rvs::FusedWork(K, tapcap, rvs_fused_chunks(npix_t, tapcap)).total;
}

namespace rvs {
// Everything one arm of an evaluation call needs: the argument records of its kernels and
// what decides whether it can share launches with the other arms.
struct ArmPlan {
  PrepArgs prep;
  ChunkArgs chunk;
  alignas(64) CUtensorMap tmap;
  GramMmaArgs gmm;     // shared pixel grid: FP64 GEMM Gram stage
  GramArgs gram;       // mixed pixel grids: per-item Gram stage
  ResolArgs resol;
  bool use_box, grid_f64, has_resol, shared_grid;
  int nvert, npoly;
  size_t smem;
  int64_t tn_stride;
  int K;
};

static int plan_arm(ArmPlan &pl, const rvs_fused_arm &arm, const int32_t *d_ids, const double *d_w,
                    int nvert, const double *d_vsini, double vsini_max, const double *d_vels,
                    int K) {
  const rvs_knots *knots = arm.knots;
  const rvs_obs *obs = arm.obs;
  TemplateArgs ta;
  ScanArgs sa;
  int rc = fill_template_args(ta, arm.d_grid, arm.ld, knots, d_ids, d_w, nvert, d_vsini,
                              arm.log_spec);
  if (rc) return rc;
  rc = fill_scan_args(sa, knots, obs);
  if (rc) return rc;
  RVS_REQUIRE(arm.d_oix && d_vels && arm.d_chisq && arm.d_status && arm.d_tn, RVS_E_ARG,
              "rvs_chisq_fused: null pointer");
  RVS_REQUIRE(knots->ratio > 0 && knots->ratio_dev < 1e-8, RVS_E_LIMIT,
              "rvs_chisq_fused: knot spacing deviates from a uniform (log-)grid by %g; use "
              "rvs_template_build + rvs_chisq_scan", knots->ratio_dev);
  const int n = knots->npix_t;
  int tapcap = 0;
  if (d_vsini && vsini_max > 0) {
    const double R = (vsini_max / RVS_C_KMS) / knots->lnstep;
    tapcap = (int)ceil(R + 1) + 1;
  }
  RVS_REQUIRE(tapcap <= RVS_MAX_FUSED_TAPS, RVS_E_LIMIT,
              "rvs_chisq_fused: vsini_max=%g needs %d taps (fused limit %d); use "
              "rvs_template_build + rvs_chisq_scan", vsini_max, tapcap, RVS_MAX_FUSED_TAPS);
  double *d_work = arm.d_work;
  RVS_REQUIRE(d_work, RVS_E_ARG, "rvs_chisq_fused: d_work is NULL");
  RVS_REQUIRE(((uintptr_t)d_work & 15) == 0, RVS_E_ARG, "rvs_chisq_fused: d_work alignment");
  const rvs_gridbox *box = arm.box;
  // copy-engine gather: dense 4-D fp32 grid, descriptor built for this tile width
  pl.use_box = box && !arm.grid_f64 && nvert == 16 && box->cols == TMA_COLS &&
               box->rows == TMA_ROWS;
  pl.grid_f64 = arm.grid_f64 != 0;
  pl.nvert = nvert;
  pl.npoly = obs->npoly;
  pl.shared_grid = obs->shared_grid != 0;
  pl.K = K;
  pl.tn_stride = arm.tn_stride;
  if (pl.use_box) memcpy(&pl.tmap, box->tmap, sizeof(pl.tmap));
  else memset(&pl.tmap, 0, sizeof(pl.tmap));
  ChunkArgs &a = pl.chunk;
  a.C = chunk_knots(tapcap);
  a.nch = (n + a.C - 1) / a.C;
  const FusedWork fw(K, tapcap, a.nch);
  a.tapstride = tapcap + 1;
  a.taps = d_work + fw.taps;
  a.rec = d_work + fw.rec;
  a.irec = reinterpret_cast<int32_t *>(d_work + fw.irec);
  a.pbound = reinterpret_cast<int32_t *>(d_work + fw.pbound);
  {
    PrepArgs &t = pl.prep;
    t.vsini = tapcap > 0 ? d_vsini : nullptr; t.lnstep = knots->lnstep; t.tapcap = tapcap;
    t.tapstride = tapcap + 1; t.K = K; t.taps = d_work + fw.taps;
    t.lam_t = knots->d_lam_t; t.npix_t = n; t.log_step = knots->log_step; t.x0 = knots->x0;
    t.xlast = knots->xlast; t.q0 = knots->q0; t.qstep_inv = knots->qstep_inv;
    t.qstep = 1.0 / knots->qstep_inv;
    t.lam = obs->d_lam; t.loglam = obs->d_loglam; t.off = obs->d_off; t.goff = obs->d_goff;
    t.oix = arm.d_oix; t.vels = d_vels; t.nch = a.nch; t.C = a.C;
    t.rec = d_work + fw.rec; t.irec = reinterpret_cast<int32_t *>(d_work + fw.irec);
    t.pbound = reinterpret_cast<int32_t *>(d_work + fw.pbound); t.status = arm.d_status;
    t.ids = d_ids; t.nvert = nvert; t.box = pl.use_box ? 1 : 0;
    for (int i = 0; i < 3; i++) t.blen[i] = pl.use_box ? box->len[i + 1] : 1;
  }
  a.grid = arm.d_grid; a.ld = arm.ld; a.npix_t = n; a.ids = d_ids; a.w = d_w; a.nvert = nvert;
  a.lam_t = knots->d_lam_t; a.hinv = knots->d_hinv; a.log_spec = arm.log_spec;
  a.log_step = knots->log_step; a.x0 = knots->x0; a.xlast = knots->xlast; a.q0 = knots->q0;
  a.qstep_inv = knots->qstep_inv;
  {  // constant-coefficient form of the spline system (chunk_kernel.cuh)
    const double r = knots->ratio;
    const double c1 = 2 * (1 + r) / (r * r), c2 = 1 / (r * r * r);
    a.rinv = 1 / r; a.r2inv = 1 / (r * r); a.c2 = c2;
    double om = c1;  // pivot of row 0 (s_0 = 0)
    for (int k = 0; k < CK_NT; k++) {
      a.wtab[k] = 1 / om;
      om = c1 - c2 / om;
    }
    for (int k = 0; k < 64; k++) om = c1 - c2 / om;
    a.winv_inf = 1 / om;
  }
  a.lam = obs->d_lam; a.loglam = obs->d_loglam; a.einv = obs->d_einv; a.off = obs->d_off;
  a.goff = obs->d_goff;
  // with resolution matrices the chunk kernel leaves T in the second half of d_tn
  pl.has_resol = obs->d_resol != nullptr;
  double *d_raw = arm.d_tn + (int64_t)K * arm.tn_stride;
  a.oix = arm.d_oix; a.vels = d_vels; a.tn = pl.has_resol ? d_raw : arm.d_tn;
  a.tn_stride = arm.tn_stride;
  a.status = arm.d_status; a.raw_out = pl.has_resol ? 1 : 0;
  a.K = K;
  a.wcap = (a.C + 2 * (CK_HALO + 3 + tapcap) + 12 + 3) & ~3;
  if (pl.use_box) {  // TMA destinations are 128-byte aligned
    a.wcap = (a.wcap + 15) & ~15;
    a.wcap1 = std::max(a.wcap, TMA_RING_DOUBLES);
  } else {
    a.wcap = std::max(a.wcap, CK_RING * 512 / 8);  // the gather's prefetch ring lives in B1
    a.wcap1 = a.wcap;
  }
  pl.smem = sizeof(double) * (size_t)(a.wcap + a.wcap1) * CK_WARPS;
  RVS_REQUIRE(pl.smem <= 200 * 1024, RVS_E_LIMIT, "rvs_chisq_fused: window needs %zu B smem",
              pl.smem);
  if (pl.has_resol) {
    ResolArgs &r = pl.resol;
    r.raw = d_raw; r.tn = arm.d_tn; r.tn_stride = arm.tn_stride; r.resol = obs->d_resol;
    r.einv = obs->d_einv; r.resol_offs = obs->d_resol_offs; r.nresol = obs->nresol;
    r.off = obs->d_off; r.oix = arm.d_oix;
  }
  double *w = d_work + fw.gram;
  if (pl.shared_grid) {  // one wavelength grid for all objects: Gram stage as an FP64 GEMM
    GramMmaArgs &m = pl.gmm;
    m.tn = arm.d_tn; m.tn_stride = arm.tn_stride; m.dn = obs->d_dn; m.sumlog2 = obs->d_sumlog2;
    m.off = obs->d_off; m.goff = obs->d_goff; m.oix = arm.d_oix; m.P = obs->d_P;
    m.npp = obs->npp; m.K = K; m.KS = 1; m.chisq = arm.d_chisq; m.status = arm.d_status;
    const GramScratch gs(K);
    m.part = w; w += gs.part();
    m.coef = w; w += gs.coef();
    m.logdet = w; w += gs.logdet();
    m.rpart = w; w += gs.rpart();
    m.ticket = reinterpret_cast<unsigned *>(w);
  } else {
    GramArgs &g = pl.gram;
    g.tn = arm.d_tn; g.tn_stride = arm.tn_stride; g.dn = obs->d_dn; g.sumlog2 = obs->d_sumlog2;
    g.off = obs->d_off; g.goff = obs->d_goff; g.oix = arm.d_oix; g.P = obs->d_P;
    g.npp = obs->npp; g.chisq = arm.d_chisq; g.status = arm.d_status;
  }
  return 0;
}

// arms that can share every launch: same kernel instantiations, same launch shapes
static bool can_merge(const ArmPlan &x, const ArmPlan &y) {
  return x.use_box == y.use_box && x.grid_f64 == y.grid_f64 && x.nvert == y.nvert &&
         x.npoly == y.npoly && x.shared_grid && y.shared_grid && !x.has_resol && !y.has_resol &&
         x.K == y.K;
}

// One launch of every kernel of the call for the arms pl[0..narm)
static int launch_arms(ArmPlan *const *pl, int narm, cudaStream_t st) {
  AuxStream *ax = aux_for(st);
  cudaStream_t s_aux = ax ? ax->aux : st;  // preparation and continuum solve
  const ArmPlan &p0 = *pl[0];
  const int K = p0.K;
  {
    PrepArgsM pm;
    for (int i = 0; i < narm; i++) pm.a[i] = pl[i]->prep;
    if (ax) hand_over(st, s_aux, ax->ev[0]);
    prof_begin(ST_PREP, s_aux);
    prep_kernel<<<dim3((K + 3) / 4, narm), 128, 0, s_aux>>>(pm);
    prof_end(ST_PREP, s_aux);
    RVS_LAUNCH_OK();
    if (ax) hand_over(s_aux, st, ax->ev[1]);
  }
  {
    ChunkArgsM cm;
    size_t smem = 0;
    for (int i = 0; i < narm; i++) {
      cm.a[i] = pl[i]->chunk;
      memcpy(&cm.tmap[i], &pl[i]->tmap, sizeof(CUtensorMap));
      smem = std::max(smem, pl[i]->smem);
    }
    // the arms' window capacities may differ (tap bound is common, chunk length too): the
    // launch takes the largest; every arm's record keeps its own layout inside it
    int rc;
    if (p0.use_box) rc = launch_chunk<float, 16, true>(cm, narm, smem, st);
    else if (p0.grid_f64) rc = launch_chunk<double, 0>(cm, narm, smem, st);
    else if (p0.nvert == 16) rc = launch_chunk<float, 16>(cm, narm, smem, st);
    else if (p0.nvert == 5) rc = launch_chunk<float, 5>(cm, narm, smem, st);
    else rc = launch_chunk<float, 0>(cm, narm, smem, st);
    if (rc) return rc;
  }
  for (int i = 0; i < narm; i++)
    if (pl[i]->has_resol) {
      dim3 grid((unsigned)K, (unsigned)((pl[i]->tn_stride + 255) / 256));
      resol_apply_kernel<<<grid, 256, 0, st>>>(pl[i]->resol);
      RVS_LAUNCH_OK();
    }
  if (ax) hand_over(st, s_aux, ax->ev[2]);
  int rc = 0;
  const int np = p0.npoly;
  if (p0.shared_grid) {
    GramMmaArgsM gm;
    for (int i = 0; i < narm; i++) gm.a[i] = pl[i]->gmm;
    if (np <= 7) rc = launch_gram_mma_group0(gm, narm, np, s_aux);
    else if (np <= 10) rc = launch_gram_mma_group1(gm, narm, np, s_aux);
    else if (np <= 13) rc = launch_gram_mma_group2(gm, narm, np, s_aux);
    else rc = launch_gram_mma_group3(gm, narm, np, s_aux);
  } else {
    for (int i = 0; i < narm && !rc; i++) {
      const GramArgs &g = pl[i]->gram;
      if (np <= 7) rc = launch_gram_group0(g, np, K, s_aux);
      else if (np <= 10) rc = launch_gram_group1(g, np, K, s_aux);
      else if (np <= 13) rc = launch_gram_group2(g, np, K, s_aux);
      else rc = launch_gram_group3(g, np, K, s_aux);
    }
  }
  if (ax) hand_over(s_aux, st, ax->ev[3]);
  return rc;
}
}  // namespace rvs

extern "C" int rvs_chisq_fused_multi(const rvs_fused_arm *arms, int narm, const int32_t *d_ids,
                                     const double *d_w, int nvert, const double *d_vsini,
                                     double vsini_max, const double *d_vels, int K, void *stream) {
  using namespace rvs;
  if (K == 0 || narm == 0) return 0;
  RVS_REQUIRE(arms && narm >= 1 && narm <= RVS_MAX_ARMS, RVS_E_ARG,
              "rvs_chisq_fused_multi: %d arms (1..%d)", narm, RVS_MAX_ARMS);
  static thread_local ArmPlan plans[RVS_MAX_ARMS];
  for (int i = 0; i < narm; i++) {
    const int rc = plan_arm(plans[i], arms[i], d_ids, d_w, nvert, d_vsini, vsini_max, d_vels, K);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  bool all = true;
  for (int i = 1; i < narm; i++) all = all && can_merge(plans[0], plans[i]);
  if (all) {
    ArmPlan *pl[RVS_MAX_ARMS];
    for (int i = 0; i < narm; i++) pl[i] = &plans[i];
    return launch_arms(pl, narm, st);
  }
  for (int i = 0; i < narm; i++) {      // arm by arm
    ArmPlan *pl[1] = {&plans[i]};
    const int rc = launch_arms(pl, 1, st);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int rvs_chisq_fused(const void *d_grid, int grid_f64, int64_t ld,
                               const rvs_knots *knots, const int32_t *d_ids, const double *d_w,
                               int nvert, const double *d_vsini, double vsini_max, int log_spec,
                               const rvs_obs *obs, const int32_t *d_oix, const double *d_vels,
                               int K, double *d_tn, int64_t tn_stride, double *d_work,
                               double *d_chisq, int32_t *d_status, const rvs_gridbox *box,
                               void *stream) {
  rvs_fused_arm arm;
  arm.d_grid = d_grid; arm.grid_f64 = grid_f64; arm.ld = ld; arm.knots = knots;
  arm.log_spec = log_spec; arm.obs = obs; arm.d_oix = d_oix; arm.d_tn = d_tn;
  arm.tn_stride = tn_stride; arm.d_work = d_work; arm.d_chisq = d_chisq; arm.d_status = d_status;
  arm.box = box;
  return rvs_chisq_fused_multi(&arm, 1, d_ids, d_w, nvert, d_vsini, vsini_max, d_vels, K, stream);
}
